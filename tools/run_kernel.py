#!/usr/bin/env python
"""Run one hot-path kernel on synthetic C4-like data through the C ABI (for ncu captures and quick timings).

    python tools/run_kernel.py score --users 37888 --items 100000 --mode 2 --reps 3
    python tools/run_kernel.py als --users 200000 --items 100000 --reps 3 --precision 0
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from devbuf import Dev, padded  # noqa: E402
from recsys_lite_b200 import _ffi, synth  # noqa: E402
from recsys_lite_b200.csr import transpose_pattern  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["score", "als"])
    ap.add_argument("--users", type=int, default=37888)
    ap.add_argument("--items", type=int, default=100000)
    ap.add_argument("--deg", type=int, default=50)
    ap.add_argument("--k", type=int, default=64)
    ap.add_argument("--mode", type=int, default=0)
    ap.add_argument("--precision", type=int, default=0)
    ap.add_argument("--ir", type=int, default=2)
    ap.add_argument("--iters", type=int, default=1, help="ALS iterations run before the scoring pass")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--every", type=int, default=0, help="also time a scoring pass after every N-th ALS iteration")
    args = ap.parse_args()
    h = _ffi.default_handle()
    nu, ni, k = args.users, args.items, args.k
    indptr, indices = synth.synth_pattern(nu, ni, args.deg, seed=4)
    indptr_t, indices_t = transpose_pattern(indptr, indices, ni)
    rng = np.random.default_rng(4)
    u0 = rng.random((nu, k), dtype=np.float32)
    v0 = rng.random((ni, k), dtype=np.float32)
    d_u, d_v = padded(h, u0, k), padded(h, v0, k)
    d_ip, d_ix, d_tp, d_tx = Dev(h, indptr), Dev(h, indices), Dev(h, indptr_t), Dev(h, indices_t)

    def als_iter():
        h.call("rl_als_half_step_dev", nu, ni, k, d_ip.ptr, d_ix.ptr, None, None, d_v.ptr, d_u.ptr, 0.01, 0.0, None,
               args.precision, args.ir, None)
        h.call("rl_als_half_step_dev", ni, nu, k, d_tp.ptr, d_tx.ptr, None, None, d_u.ptr, d_v.ptr, 0.01, 0.0, None,
               args.precision, args.ir, None)

    if args.what == "als":
        for r in range(args.reps):
            h.synchronize()
            t = time.perf_counter()
            als_iter()
            h.synchronize()
            print(f"als iteration {r}: {(time.perf_counter() - t) * 1e3:.3f} ms")
        return
    d_idx = Dev(h, shape=(nu, 10), dtype=np.int32)
    for it in range(args.iters):
        als_iter()
        if args.every and (it + 1) % args.every == 0 and it + 1 < args.iters:
            h.synchronize()
            t = time.perf_counter()
            h.call("rl_score_topn_dev", d_u.ptr, nu, d_v.ptr, ni, k, d_ip.ptr, d_ix.ptr, 0.0, None, None, 10, d_idx.ptr, None, args.mode)
            h.synchronize()
            print(f"after {it + 1} iterations: score_topn mode {args.mode}: {(time.perf_counter() - t) * 1e3:.3f} ms, overflow rows "
                  f"{h.lib.rl_last_overflow_rows(h.h)}")
    for r in range(args.reps):
        h.synchronize()
        t = time.perf_counter()
        seen = (None, None) if os.environ.get("RL_NO_SEEN") else (d_ip.ptr, d_ix.ptr)
        h.call("rl_score_topn_dev", d_u.ptr, nu, d_v.ptr, ni, k, seen[0], seen[1], 0.0, None, None, 10, d_idx.ptr, None,
               args.mode)
        h.synchronize()
        print(f"score_topn mode {args.mode} rep {r}: {(time.perf_counter() - t) * 1e3:.3f} ms, overflow rows "
              f"{h.lib.rl_last_overflow_rows(h.h)}")


if __name__ == "__main__":
    main()

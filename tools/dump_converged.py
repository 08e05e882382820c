#!/usr/bin/env python
"""Run T ALS iterations on a workload of bench.py and dump what the scoring kernel sees on converged factors.

    python tools/dump_converged.py --workload C4 --iters 25 --out gpurun_out/conv_c4.npz

Written: V (all items, float32) after the last iteration, the first --sample user rows of U with their seen lists, the
same for an early iteration, and per-iteration scoring time / overflow rows of the tensor-core path.  The file feeds the
offline analysis of the candidate window (tools/analyze_window.py); nothing in the product reads it.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import bench  # noqa: E402
from devbuf import Dev, padded, unpadded  # noqa: E402
from recsys_lite_b200 import _ffi  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="C4")
    ap.add_argument("--iters", type=int, default=25)
    ap.add_argument("--early", type=int, default=6)
    ap.add_argument("--sample", type=int, default=8192)
    ap.add_argument("--score-every", type=int, default=1)
    ap.add_argument("--out", default="gpurun_out/conv.npz")
    args = ap.parse_args()
    w = bench.WORKLOADS[args.workload]
    indptr, indices, indptr_t, indices_t, u0, v0 = bench.make_inputs(w)
    nu, ni, k = w["n_users"], w["n_items"], w["k"]
    h = _ffi.default_handle()
    d_u, d_v = padded(h, u0, k), padded(h, v0, k)
    d_ip, d_ix, d_tp, d_tx = Dev(h, indptr), Dev(h, indices), Dev(h, indptr_t), Dev(h, indices_t)
    d_idx = Dev(h, shape=(nu, 10), dtype=np.int32)
    ns = min(args.sample, nu)
    out = {"indptr_s": indptr[: ns + 1].copy(), "indices_s": indices[: indptr[ns]].copy()}
    log = []
    for it in range(1, args.iters + 1):
        h.synchronize()
        t0 = time.perf_counter()
        h.call("rl_als_half_step_dev", nu, ni, k, d_ip.ptr, d_ix.ptr, None, None, d_v.ptr, d_u.ptr, w["lam"], 0.0, None, 0, 1, None)
        h.call("rl_als_half_step_dev", ni, nu, k, d_tp.ptr, d_tx.ptr, None, None, d_u.ptr, d_v.ptr, w["lam"], 0.0, None, 0, 1, None)
        h.synchronize()
        t_als = time.perf_counter() - t0
        rec = {"iter": it, "als_ms": round(t_als * 1e3, 3)}
        if it % args.score_every == 0 or it == args.iters:
            t0 = time.perf_counter()
            h.call("rl_score_topn_dev", d_u.ptr, nu, d_v.ptr, ni, k, d_ip.ptr, d_ix.ptr, 0.0, None, None, 10, d_idx.ptr, None, 0)
            h.synchronize()
            rec["score_ms"] = round((time.perf_counter() - t0) * 1e3, 3)
            rec["overflow_rows"] = int(h.lib.rl_last_overflow_rows(h.h))
        log.append(rec)
        print(json.dumps(rec), flush=True)
        if it in (args.early, args.iters):
            tag = "early" if it == args.early and it != args.iters else "last"
            v = unpadded(h, d_v, k, np.float32)
            u = unpadded(h, d_u, k, np.float32)
            out[f"v_{tag}"] = v if tag == "last" else v[: ni // 2]
            out[f"u_{tag}"] = u[:ns]
            out[f"umean_{tag}"] = u.astype(np.float64).mean(axis=0)
            out[f"unorm_{tag}"] = np.linalg.norm(u[::97].astype(np.float64), axis=1)
            if tag == "last":
                out["idx_last"] = d_idx.get()[:ns]
            del u, v
    out["log"] = np.frombuffer(json.dumps(log).encode(), dtype=np.uint8)
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    np.savez(args.out, **out)
    print("wrote", args.out, os.path.getsize(args.out) >> 20, "MiB")


if __name__ == "__main__":
    main()

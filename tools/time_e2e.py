#!/usr/bin/env python
"""Time the two host-buffer C ABI calls of the end-to-end step separately (C4 shapes by default)."""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from recsys_lite_b200 import _ffi, synth  # noqa: E402
from recsys_lite_b200.csr import transpose_pattern  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--users", type=int, default=1000000)
    ap.add_argument("--items", type=int, default=100000)
    ap.add_argument("--k", type=int, default=64)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--device-transpose", action="store_true", help="pass NULL for the transposed pattern (built on the device)")
    args = ap.parse_args()
    h = _ffi.default_handle()
    nu, ni, k = args.users, args.items, args.k
    indptr, indices = synth.synth_pattern(nu, ni, 50, seed=4)
    indptr_t, indices_t = transpose_pattern(indptr, indices, ni)
    rng = np.random.default_rng(4)
    u0 = rng.random((nu, k), dtype=np.float32)
    v0 = rng.random((ni, k), dtype=np.float32)

    def pin(a):
        out = _ffi.pinned_empty(a.shape, a.dtype)
        out[...] = a
        return out
    p_ip, p_ix, p_tp, p_tx, p_u, p_v = map(pin, (indptr, indices, indptr_t, indices_t, u0, v0))
    idx = _ffi.pinned_empty((nu, 10), np.int32)
    for r in range(args.reps):
        p_u[:] = u0
        p_v[:] = v0
        h.synchronize()
        t0 = time.perf_counter()
        h.call("rl_als_fit_host", nu, ni, k, _ffi.ptr(p_ip), _ffi.ptr(p_ix), None if args.device_transpose else _ffi.ptr(p_tp),
               None if args.device_transpose else _ffi.ptr(p_tx), 1, 0.01, 0, 1, _ffi.ptr(p_u), _ffi.ptr(p_v), _ffi.RL_F32)
        t1 = time.perf_counter()
        h.call("rl_recommend_host", _ffi.ptr(p_u), _ffi.ptr(p_v), _ffi.RL_F32, nu, ni, k, _ffi.ptr(p_ip), _ffi.ptr(p_ix), 0.0, None,
               None, 10, _ffi.ptr(idx), None, 0)
        t2 = time.perf_counter()
        print(f"rep {r}: fit_host {(t1 - t0) * 1e3:.1f} ms, recommend_host {(t2 - t1) * 1e3:.1f} ms, overflow rows "
              f"{h.lib.rl_last_overflow_rows(h.h)}")


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Per-row accuracy of one ALS user half-step against float64 normal equations on sampled rows (C4-like or C5-like shapes).

    python tools/als_row_accuracy.py --k 64 --users 200000 --items 100000 --ir 1
"""
import argparse, os, sys
import numpy as np
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from devbuf import Dev, padded, unpadded  # noqa: E402
from recsys_lite_b200 import _ffi, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--k", type=int, default=64); ap.add_argument("--users", type=int, default=200000)
ap.add_argument("--items", type=int, default=100000); ap.add_argument("--ir", type=int, default=1)
ap.add_argument("--iters", type=int, default=1); ap.add_argument("--sample", type=int, default=3000)
a = ap.parse_args()
h = _ffi.default_handle()
nu, ni, k, lam = a.users, a.items, a.k, 0.01
indptr, indices = synth.synth_pattern(nu, ni, 50, seed=4)
from recsys_lite_b200.csr import transpose_pattern
tp, tx = transpose_pattern(indptr, indices, ni)
rng = np.random.default_rng(4)
u0 = rng.random((nu, k), dtype=np.float32); v0 = rng.random((ni, k), dtype=np.float32)
d_u, d_v = padded(h, u0, k), padded(h, v0, k)
d_ip, d_ix, d_tp, d_tx = Dev(h, indptr), Dev(h, indices), Dev(h, tp), Dev(h, tx)
for it in range(a.iters):
    V = unpadded(h, d_v, k, np.float32)
    h.call("rl_als_half_step_dev", nu, ni, k, d_ip.ptr, d_ix.ptr, None, None, d_v.ptr, d_u.ptr, lam, 0.0, None, 0, a.ir, None)
    U = unpadded(h, d_u, k, np.float32)
    if it == a.iters - 1:
        rows = rng.choice(nu, a.sample, replace=False)
        err = np.empty(len(rows)); deg = np.diff(indptr)[rows]
        for j, r in enumerate(rows):
            ys = V[indices[indptr[r]:indptr[r + 1]]].astype(np.float64)
            ref = np.linalg.solve(ys.T @ ys + lam * np.eye(k), ys.sum(axis=0))
            err[j] = np.linalg.norm(U[r] - ref) / np.linalg.norm(ref)
        print(f"k={k} ir={a.ir} after {a.iters} iteration(s): user rows rel err median {np.median(err):.2e} p99 {np.percentile(err, 99):.2e} max {err.max():.2e} "
              f"(deg of worst {deg[err.argmax()]}); rows with nnz<k: {np.mean(deg < k):.2f}")
    h.call("rl_als_half_step_dev", ni, nu, k, d_tp.ptr, d_tx.ptr, None, None, d_u.ptr, d_v.ptr, lam, 0.0, None, 0, a.ir, None)

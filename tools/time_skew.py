import sys, time, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from devbuf import Dev, padded
from recsys_lite_b200 import _ffi, synth
from recsys_lite_b200.csr import transpose_pattern
h = _ffi.default_handle()
nu, ni, k = 100000, 20000, 64
for skew in (False, True):
    indptr, indices = synth.synth_pattern(nu, ni, 50, seed=13, skew=skew)
    ti, tx = transpose_pattern(indptr, indices, ni)
    print("skew", skew, "nnz", len(indices), "max user deg", np.diff(indptr).max(), "max item deg", np.diff(ti).max())
    rng = np.random.default_rng(4)
    d_u, d_v = padded(h, rng.random((nu, k), dtype=np.float32), k), padded(h, rng.random((ni, k), dtype=np.float32), k)
    d_ip, d_ix, d_tp, d_tx = Dev(h, indptr), Dev(h, indices), Dev(h, ti), Dev(h, tx)
    order = Dev(h, np.argsort(-np.diff(ti), kind="stable").astype(np.int32))
    for r in range(3):
        h.synchronize(); t = time.perf_counter()
        h.call("rl_als_half_step_dev", nu, ni, k, d_ip.ptr, d_ix.ptr, None, None, d_v.ptr, d_u.ptr, 0.01, 0.0, None, 0, 1, None)
        h.synchronize(); t1 = time.perf_counter()
        h.call("rl_als_half_step_dev", ni, nu, k, d_tp.ptr, d_tx.ptr, None, None, d_u.ptr, d_v.ptr, 0.01, 0.0, None, 0, 1, None)
        h.synchronize(); t2 = time.perf_counter()
        h.call("rl_als_half_step_dev", ni, nu, k, d_tp.ptr, d_tx.ptr, None, None, d_u.ptr, d_v.ptr, 0.01, 0.0, None, 0, 1, order.ptr)
        h.synchronize(); t3 = time.perf_counter()
        print(f"  user {1e3*(t1-t):.2f} ms, item {1e3*(t2-t1):.2f} ms, item longest-first {1e3*(t3-t2):.2f} ms")

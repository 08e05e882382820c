import sys, time, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from devbuf import Dev, padded
from recsys_lite_b200 import _ffi, synth
h = _ffi.default_handle()
ni, k = 100000, 64
rng = np.random.default_rng(4)
v0 = rng.random((ni, k), dtype=np.float32)
d_v = padded(h, v0, k)
for nu in (1, 8, 64, 1000):
    indptr, indices = synth.synth_pattern(nu, ni, 50, seed=4)
    u0 = rng.random((nu, k), dtype=np.float32)
    d_u = padded(h, u0, k); d_ip, d_ix = Dev(h, indptr), Dev(h, indices)
    d_idx = Dev(h, shape=(nu, 10), dtype=np.int32)
    for r in range(3):
        h.synchronize(); t = time.perf_counter()
        h.call("rl_score_topn_dev", d_u.ptr, nu, d_v.ptr, ni, k, d_ip.ptr, d_ix.ptr, 0.0, None, None, 10, d_idx.ptr, None, 1)
        h.synchronize(); print(f"exact nu={nu} rep {r}: {(time.perf_counter()-t)*1e3:.3f} ms")

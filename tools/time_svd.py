import sys, time, os
sys.path.insert(0,'.')
import numpy as np
from recsys_lite_b200 import svd_reconstruct, _ffi, synth
m = synth.c2_readme_matrix().astype(np.float32)
h = _ffi.default_handle()
for rep in range(4):
    l0 = h.launches
    t = time.perf_counter(); r = svd_reconstruct(m, k=10, use_sparse=False); dt = time.perf_counter() - t
    print(os.environ.get("RL_SVD_GRAPH_SWEEPS"), "svd_reconstruct", round(dt*1e3,2), "ms, launches", h.launches - l0)

import numpy as np, time, ctypes as C, sys
sys.path.insert(0,'.')
from scipy import sparse
from recsys_lite_b200 import synth, _ffi
indptr, indices = synth.synth_pattern(1000000, 100000, 50, seed=4)
m = sparse.csr_matrix((np.ones(len(indices), np.float32), indices, indptr), shape=(1000000,100000))
flags=C.c_int(0)
lib=_ffi.load_library()
for nt in (1,2,4,8,16,32):
    t=time.perf_counter()
    rc=lib.rl_csr_check_host(_ffi.ptr(m.indptr), 2, _ffi.ptr(m.indices), _ffi.ptr(m.data), 0, m.shape[0], m.shape[1], nt, C.byref(flags))
    print('csr check threads', nt, rc, flags.value, round(1e3*(time.perf_counter()-t),2), 'ms')

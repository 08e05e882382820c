"""ctypes binding of librecsys_b200.so (include/recsys_b200.h).  No CPU fallback: loading or device
creation failures raise, they are never papered over with NumPy."""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RL_B200_LIB") or os.path.join(HERE, "librecsys_b200.so")   # RL_B200_LIB: kernel-variant experiments

RL_F32, RL_F64, RL_I32, RL_I64 = 0, 1, 2, 3
RL_PREC_FP32_IR, RL_PREC_FP64 = 0, 1
RL_SCORE_AUTO, RL_SCORE_EXACT, RL_SCORE_TENSOR, RL_SCORE_TENSOR_TF32X1, RL_SCORE_TENSOR_CENTRED = 0, 1, 2, 3, 4
ABI_VERSION = 1

_i64, _i32, _dbl, _vp = C.c_int64, C.c_int, C.c_double, C.c_void_p

# name -> (restype, argtypes); exactly the declarations of include/recsys_b200.h
SIGNATURES = {
    "rl_abi_version": (_i32, []),
    "rl_create": (_i32, [_i32, C.POINTER(_vp)]),
    "rl_destroy": (_i32, [_vp]),
    "rl_last_error": (C.c_char_p, [_vp]),
    "rl_set_stream": (_i32, [_vp, _vp]),
    "rl_synchronize": (_i32, [_vp]),
    "rl_device_info": (_i32, [_vp, C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i64)]),
    "rl_launch_count": (_i64, [_vp]),
    "rl_trim": (_i32, [_vp]),
    "rl_padded_k": (_i32, [_i32]),
    "rl_pinned_alloc": (_i32, [_i64, C.POINTER(_vp)]),
    "rl_pinned_free": (_i32, [_vp]),
    "rl_dev_alloc": (_i32, [_vp, _i64, C.POINTER(_vp)]),
    "rl_dev_free": (_i32, [_vp, _vp]),
    "rl_dev_alloc_async": (_i32, [_vp, _i64, C.POINTER(_vp)]),
    "rl_dev_free_async": (_i32, [_vp, _vp]),
    "rl_memcpy_h2d": (_i32, [_vp, _vp, _vp, _i64]),
    "rl_memcpy_d2h": (_i32, [_vp, _vp, _vp, _i64]),
    "rl_factors_upload": (_i32, [_vp, _vp, _i32, _i64, _i32, _vp]),
    "rl_factors_download": (_i32, [_vp, _vp, _i64, _i32, _vp, _i32]),
    "rl_als_half_step_dev": (_i32, [_vp, _i64, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _dbl, _dbl, _vp, _i32, _i32, _vp]),
    "rl_als_half_step_peers_dev": (_i32, [_vp, _i64, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _dbl, _dbl, _vp, _i32, _i32, _vp,
                                          _i32, _vp]),
    "rl_ipc_export": (_i32, [_vp, _vp, _vp, C.POINTER(_i64)]),
    "rl_ipc_open": (_i32, [_vp, _vp, _i64, C.POINTER(_vp)]),
    "rl_ipc_close": (_i32, [_vp, _vp]),
    "rl_gram_dev": (_i32, [_vp, _vp, _i64, _i32, _vp]),
    "rl_csr_transpose_dev": (_i32, [_vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp]),
    "rl_als_fit_host": (_i32, [_vp, _i64, _i64, _i32, _vp, _vp, _vp, _vp, _i32, _dbl, _i32, _i32, _vp, _vp, _i32]),
    "rl_als_fit_resident_host": (_i32, [_vp, _i64, _i64, _i32, _vp, _vp, _vp, _vp, _i32, _dbl, _i32, _i32, _vp, _vp, _i32, C.POINTER(_vp)]),
    "rl_csr_check_host": (_i32, [_vp, _i32, _vp, _vp, _i32, _i64, _i64, _i32, C.POINTER(_i32)]),
    "rl_recommend_resident_host": (_i32, [_vp, _vp, _i64, _vp, _i64, _i32, _vp, _vp, _i32, _vp, _i32]),
    "rl_score_topn_dev": (_i32, [_vp, _vp, _i64, _vp, _i64, _i32, _vp, _vp, _dbl, _vp, _vp, _i32, _vp, _vp, _i32]),
    "rl_score_dump_dev": (_i32, [_vp, _vp, _i64, _vp, _i64, _i32, _vp, _i64, _i32]),
    "rl_score_dump_centred_dev": (_i32, [_vp, _vp, _i64, _vp, _i64, _i32, _vp, _i64, _vp, _vp, _vp]),
    "rl_last_overflow_rows": (_i64, [_vp]),
    "rl_recommend_host": (_i32, [_vp, _vp, _vp, _i32, _i64, _i64, _i32, _vp, _vp, _dbl, _vp, _vp, _i32, _vp, _vp, _i32]),
    "rl_topn_dense_dev": (_i32, [_vp, _vp, _i32, _vp, _i32, _i64, _i64, _i32, _vp, _vp]),
    "rl_topn_dense_host": (_i32, [_vp, _vp, _i32, _vp, _i32, _i64, _i64, _i32, _vp, _vp]),
    "rl_svd_topk_dev": (_i32, [_vp, _vp, _i64, _i64, _i32, _vp, _vp, _vp]),
    "rl_lowrank_dev": (_i32, [_vp, _vp, _vp, _i32, _i64, _i64, _i32, _dbl, _vp, _vp, _vp, _i32]),
    "rl_svd_reconstruct_host": (_i32, [_vp, _vp, _i64, _i64, _i32, _vp, _vp, _vp, _vp]),
    "rl_svd_fit_host": (_i32, [_vp, _vp, _i64, _i64, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "rl_bias_centre_dev": (_i32, [_vp, _vp, _i64, _i64, _i32, _vp, _vp, _vp, _vp]),
    "rl_gram_tc_dev": (_i32, [_vp, _vp, _i32, _i32, _vp, _i32, _i32, _i64, _i32, _i32, _vp, _i32, _i32]),
    "rl_knn_item_sim_host": (_i32, [_vp, _vp, _i64, _i32, _vp]),
    "rl_knn_predict_host": (_i32, [_vp, _vp, _i64, _i32, _vp, _i32, _vp]),
    "rl_ranking_metrics_host": (_i32, [_vp, _vp, _i64, _i32, _i32, _vp, _vp, _vp]),
    "rl_predict_host": (_i32, [_vp, _vp, _vp, _i32, _i64, _i64, _i32, _dbl, _vp, _vp, _vp, _i32]),
    "rl_dense_error_host": (_i32, [_vp, _vp, _i32, _vp, _i32, _i64, _vp, _vp]),
    "rl_observed_error_dev": (_i32, [_vp, _vp, _vp, _i32, _i64, _vp, _vp, _vp, _vp]),
}


class RecsysB200Error(RuntimeError):
    """A CUDA / library failure surfaced through the C ABI (reference: LinAlgError -> RuntimeError, algo.py:527-530)."""


_lib = None
_lib_lock = threading.Lock()


def load_library(path: str | None = None):
    """dlopen the shared object and attach the prototypes.  Raises if it is missing (no fallback)."""
    global _lib
    with _lib_lock:
        if _lib is not None and path is None:
            return _lib
        p = path or LIB_PATH
        if not os.path.exists(p):
            raise RecsysB200Error(
                f"{p} not found: build it with `python -m recsys_lite_b200.build` (nvcc, sm_100a). "
                "recsys_lite_b200 has no CPU fallback."
            )
        lib = C.CDLL(p)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
            fn.restype, fn.argtypes = res, args
        if lib.rl_abi_version() != ABI_VERSION:
            raise RecsysB200Error(f"ABI mismatch: library {lib.rl_abi_version()} vs binding {ABI_VERSION}")
        if path is None:
            _lib = lib
        return lib


def ptr(a):
    """void* of a NumPy array (must be C-contiguous), an int address, or None."""
    if a is None:
        return None
    if isinstance(a, (int, np.integer)):
        return _vp(int(a))
    if not a.flags["C_CONTIGUOUS"]:
        raise ValueError("array handed to the C ABI must be C-contiguous")
    return _vp(a.ctypes.data)


class Handle:
    """One rl_handle = one device context + stream.  Not shared between threads (reference is single-threaded)."""

    def __init__(self, device: int = -1):
        self.lib = load_library()
        h = _vp()
        rc = self.lib.rl_create(device, C.byref(h))
        if rc != 0:
            msg = self.lib.rl_last_error(None).decode()
            raise RecsysB200Error(f"rl_create failed ({rc}): {msg}")
        self.h = h

    def check(self, rc: int, what: str = ""):
        if rc != 0:
            msg = self.lib.rl_last_error(self.h).decode()
            raise RecsysB200Error(f"{what or 'librecsys_b200'} failed ({rc}): {msg}")

    def call(self, name: str, *args):
        self.check(getattr(self.lib, name)(self.h, *args), name)

    def set_stream(self, stream_ptr: int | None):
        """stream_ptr: a cudaStream_t as int (0 = CUDA default stream); None = the handle's own stream."""
        self.call("rl_set_stream", _vp(-1 & 0xFFFFFFFFFFFFFFFF) if stream_ptr is None else _vp(int(stream_ptr)))

    def synchronize(self):
        self.call("rl_synchronize")

    @property
    def launches(self) -> int:
        return int(self.lib.rl_launch_count(self.h))

    def device_info(self):
        sm, ma, mi, mem = _i32(), _i32(), _i32(), _i64()
        self.call("rl_device_info", C.byref(sm), C.byref(ma), C.byref(mi), C.byref(mem))
        return {"sm_count": sm.value, "cc": (ma.value, mi.value), "total_mem": mem.value}

    def close(self):
        if getattr(self, "h", None):
            self.lib.rl_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default = None
_default_lock = threading.Lock()


def default_handle() -> Handle:
    """Process-wide handle used by the drop-in API (created on first use; raises without a B200)."""
    global _default
    with _default_lock:
        if _default is None:
            _default = Handle(-1)
        return _default


def pinned_empty(shape, dtype):
    """NumPy array backed by cudaMallocHost memory (kept alive by the array's base object)."""
    lib = load_library()
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    p = _vp()
    rc = lib.rl_pinned_alloc(n, C.byref(p))
    if rc != 0:
        raise RecsysB200Error(f"rl_pinned_alloc({n}) failed: {lib.rl_last_error(None).decode()}")

    class _Owner:
        def __init__(self, addr):
            self.addr = addr

        def __del__(self):
            lib.rl_pinned_free(_vp(self.addr))

    buf = (C.c_char * max(n, 1)).from_address(p.value)
    buf._owner = _Owner(p.value)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    return arr

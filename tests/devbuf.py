"""Tiny device-array helper for the -m gpu tests: device memory through the C ABI only (no torch)."""
import ctypes as C

import numpy as np

from recsys_lite_b200 import _ffi


class Dev:
    def __init__(self, h, arr=None, *, shape=None, dtype=None):
        self.h = h
        if arr is not None:
            arr = np.ascontiguousarray(arr)
            shape, dtype = arr.shape, arr.dtype
        self.shape, self.dtype = tuple(shape), np.dtype(dtype)
        self.nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        p = C.c_void_p()
        h.call("rl_dev_alloc", self.nbytes, C.byref(p))
        self.ptr = p
        if arr is not None and self.nbytes:
            h.call("rl_memcpy_h2d", self.ptr, _ffi.ptr(arr), self.nbytes)

    def get(self):
        out = np.empty(self.shape, dtype=self.dtype)
        if self.nbytes:
            self.h.call("rl_memcpy_d2h", _ffi.ptr(out), self.ptr, self.nbytes)
        else:
            self.h.synchronize()
        return out

    def __del__(self):
        try:
            if self.ptr:
                self.h.synchronize()
                self.h.call("rl_dev_free", self.ptr)
                self.ptr = None
        except Exception:
            pass


def padded(h, x, k):
    """host (rows x k) float -> device (rows x padded_k) float32 through rl_factors_upload."""
    x = np.ascontiguousarray(x)
    kp = h.lib.rl_padded_k(k)
    d = Dev(h, shape=(x.shape[0], kp), dtype=np.float32)
    h.call("rl_factors_upload", _ffi.ptr(x), _ffi.RL_F64 if x.dtype == np.float64 else _ffi.RL_F32,
           x.shape[0], k, d.ptr)
    return d


def unpadded(h, d, k, dtype=np.float64):
    out = np.empty((d.shape[0], k), dtype=dtype)
    h.call("rl_factors_download", d.ptr, d.shape[0], k, _ffi.ptr(out),
           _ffi.RL_F64 if dtype == np.float64 else _ffi.RL_F32)
    return out

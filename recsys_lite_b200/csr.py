"""Host-side pattern handling for the device kernels: observed cells -> CSR / transposed CSR.

The reference scans a dense matrix per user (`ratings[u] > 0`, algo.py:315) and per item
(`ratings[:, i] > 0`, algo.py:325).  The kernels instead take the observation pattern once as CSR
(rows = users) plus the CSR of its transpose (rows = items): indptr int64, indices int32 ascending.
"""
from __future__ import annotations

import numpy as np
from scipy import sparse


_DT = {np.dtype(np.float32): 0, np.dtype(np.float64): 1, np.dtype(np.int32): 2, np.dtype(np.int64): 3}


def _checked_fast(m):
    """CSR arrays of `m` as they are when the library's multi-threaded check (rl_csr_check_host: rows strictly ascending,
    columns in range, every stored value > 0) passes; None otherwise (the scipy route below repairs the matrix).
    scipy's own has_canonical_format + data.min() cost 60 ms per call at 50 M entries."""
    import ctypes as C
    from . import _ffi
    if m.indices.dtype != np.int32 or m.indptr.dtype not in _DT or m.data.dtype not in _DT:
        return None
    if not (m.indices.flags["C_CONTIGUOUS"] and m.indptr.flags["C_CONTIGUOUS"] and m.data.flags["C_CONTIGUOUS"]):
        return None
    flags = C.c_int(0)
    rc = _ffi.load_library().rl_csr_check_host(_ffi.ptr(m.indptr), _DT[m.indptr.dtype], _ffi.ptr(m.indices), _ffi.ptr(m.data),
                                               _DT[m.data.dtype], m.shape[0], m.shape[1], 0, C.byref(flags))
    if rc != 0 or flags.value != 3:
        return None
    return np.ascontiguousarray(m.indptr, dtype=np.int64), m.indices


def observed_pattern(ratings):
    """(indptr int64, indices int32) of the cells with value > 0 (negatives count as unobserved, algo.py:311)."""
    if sparse.issparse(ratings):
        m = sparse.csr_matrix(ratings)
        if m.nnz >= (1 << 16):
            fast = _checked_fast(m)
            if fast is not None:
                return fast
        if not m.has_canonical_format:
            m = m.copy()
            m.sum_duplicates()
        if m.nnz and not (m.data.min() > 0):          # one pass; the common case (all stored values positive) allocates nothing
            keep = m.data > 0
            m = sparse.csr_matrix((m.data * keep, m.indices.copy(), m.indptr.copy()), shape=m.shape)   # eliminate_zeros works in place
            m.eliminate_zeros()
        if not m.has_sorted_indices:
            m = m.sorted_indices()
        return np.ascontiguousarray(m.indptr, dtype=np.int64), np.ascontiguousarray(m.indices, dtype=np.int32)
    mask = np.asarray(ratings) > 0
    indptr = np.zeros(mask.shape[0] + 1, dtype=np.int64)
    np.cumsum(mask.sum(axis=1), out=indptr[1:])
    return indptr, np.ascontiguousarray(np.nonzero(mask)[1], dtype=np.int32)


def transpose_pattern(indptr, indices, n_cols):
    """CSR of the transposed pattern (ascending inside each row), via scipy's compiled csr->csc."""
    n_rows = len(indptr) - 1
    m = sparse.csr_matrix((np.ones(len(indices), dtype=np.int8), indices, indptr), shape=(n_rows, n_cols))
    t = m.tocsc()
    t.sort_indices()
    return np.ascontiguousarray(t.indptr, dtype=np.int64), np.ascontiguousarray(t.indices, dtype=np.int32)


def nnz_balanced_bounds(indptr, parts):
    """Row boundaries of `parts` contiguous row ranges with (nearly) equal nnz.

    Returns an int64 array of length parts+1; range p is rows [b[p], b[p+1]).
    """
    indptr = np.asarray(indptr)
    n_rows = len(indptr) - 1
    total = int(indptr[-1])
    targets = (np.arange(1, parts, dtype=np.float64) * total / parts)
    cuts = np.searchsorted(indptr[1:], targets, side="left") + 1 if total > 0 else \
        (np.arange(1, parts) * n_rows) // parts
    b = np.concatenate(([0], np.minimum(cuts, n_rows), [n_rows])).astype(np.int64)
    return np.maximum.accumulate(b)


def shard_rows(indptr, indices, lo, hi):
    """CSR of rows [lo, hi) with a local indptr starting at 0 (indices untouched)."""
    base = int(indptr[lo])
    return (np.ascontiguousarray(indptr[lo:hi + 1] - base, dtype=np.int64),
            np.ascontiguousarray(indices[base:int(indptr[hi])], dtype=np.int32))

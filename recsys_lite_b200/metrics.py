"""Ranking metrics of the reference (tools.py:50-110: precision@k, recall@k, NDCG@k) on the device.

Inputs as the reference takes them (``recs``: list of lists of item ids, ``actual``: list of sets) or as arrays
(``recs``: (n_users, n) int array padded with -1 -- what ``recommend`` returns; ``actual``: scipy sparse matrix whose
stored entries are the relevant items).  One kernel computes the three user averages (csrc/score.cu:
ranking_metrics_kernel).  Users without recommendations or without relevant items count as 0, as in the reference.
The reference's NDCG has a quirk that is kept: its "ideal" list counts every position (tools.py:100-103), so
IDCG = sum_{i < min(k, len(recs))} 1 / log2(i + 2); lists of length 1 follow the reference's arithmetic on the host.
"""
from __future__ import annotations

import numpy as np
from scipy import sparse

from . import _ffi


def _as_arrays(recs, actual):
    if isinstance(recs, np.ndarray):
        r = np.ascontiguousarray(recs, dtype=np.int32)
    else:
        width = max((len(x) for x in recs), default=0)
        r = np.full((len(recs), max(width, 1)), -1, dtype=np.int32)
        for i, x in enumerate(recs):
            r[i, : len(x)] = list(x)
    if sparse.issparse(actual):
        a = sparse.csr_matrix(actual)
        a.sort_indices()
        indptr, indices = a.indptr.astype(np.int64), a.indices.astype(np.int32)
    else:
        rows = [np.sort(np.fromiter(s, dtype=np.int32, count=len(s))) for s in actual]
        indptr = np.zeros(len(rows) + 1, dtype=np.int64)
        np.cumsum([len(x) for x in rows], out=indptr[1:])
        indices = np.concatenate(rows).astype(np.int32) if rows and indptr[-1] else np.zeros(0, dtype=np.int32)
    n = min(r.shape[0], len(indptr) - 1)         # the reference zips the two lists
    return np.ascontiguousarray(r[:n]), np.ascontiguousarray(indptr[: n + 1]), np.ascontiguousarray(indices)


def ranking_metrics(recs, actual, k: int):
    """(precision@k, recall@k, ndcg@k) in one device pass."""
    if k <= 0 or len(recs) == 0 or (not sparse.issparse(actual) and len(actual) == 0):
        return 0.0, 0.0, 0.0
    r, indptr, indices = _as_arrays(recs, actual)
    if r.shape[0] == 0:
        return 0.0, 0.0, 0.0
    out = np.zeros(3, dtype=np.float64)
    h = _ffi.default_handle()
    h.call("rl_ranking_metrics_host", _ffi.ptr(r), r.shape[0], r.shape[1], int(k), _ffi.ptr(indptr), _ffi.ptr(indices), _ffi.ptr(out))
    return float(out[0]), float(out[1]), float(out[2])


def precision_at_k(recs, actual, k: int) -> float:
    return ranking_metrics(recs, actual, k)[0]


def recall_at_k(recs, actual, k: int) -> float:
    return ranking_metrics(recs, actual, k)[1]


def ndcg_at_k(recs, actual, k: int) -> float:
    return ranking_metrics(recs, actual, k)[2]

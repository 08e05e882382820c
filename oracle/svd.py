"""Oracle: truncated-SVD reconstruct and the SVD class path (TEST INFRASTRUCTURE).

Follows /root/reference/src/recsys_lite/algo.py:412-530 (``svd_reconstruct``, dense branch :510-517),
:210-248 (``_fit_svd_dense``) and :352-375 (``_predict_svd``).

Third-party arithmetic: ``np.linalg.svd`` -> LAPACK sgesdd in float32 (algo.py:238,512).
The CUDA path reaches the same rank-k reconstruction through the identity
``U_k S_k V_k^T = M V_k V_k^T`` (V_k = top-k eigenvectors of M^T M, float64), so only
RECONSTRUCTIONS are comparable, never the factors (sign / rotation ambiguity).
"""

from __future__ import annotations

import numpy as np


def svd_reconstruct_dense(mat, k=None):
    """Dense, unchunked ``svd_reconstruct(mat, k=k, use_sparse=False)`` -> float32 ndarray."""
    mat = np.asarray(mat)
    if mat.ndim != 2:
        raise ValueError(f"Input matrix must be 2D (users x items), got {mat.ndim}D.")
    n_users, n_items = mat.shape
    if n_users == 0 or n_items == 0:
        raise ValueError("Input matrix cannot have zero dimensions.")
    k = max(1, min(n_users, n_items) // 4) if k is None else k
    if not 1 <= k <= min(n_users, n_items):
        raise ValueError(
            f"k must be between 1 and min(n_users, n_items)={min(n_users, n_items)}, got k={k}."
        )
    if np.all(mat == 0):
        return np.zeros_like(mat)
    u, s, vt = np.linalg.svd(mat.astype(np.float32), full_matrices=False)
    return u[:, :k] @ (s[:k, None] * vt[:k, :])


def reconstruct_f64(mat, k):
    """Float64 ground truth of the same estimator (used to bound both fp32 LAPACK and the CUDA path)."""
    u, s, vt = np.linalg.svd(np.asarray(mat, dtype=np.float64), full_matrices=False)
    return u[:, :k] @ (s[:k, None] * vt[:k, :])


def svd_biases(ratings):
    """global / user / item bias of ``_fit_svd_dense`` (algo.py:218-234), vectorised.

    global = mean of entries > 0; user/item bias = mean of that row's/column's entries > 0 minus
    global, 0 where there is none.  The reference computes these with float32 ``np.mean`` per
    row; the means here are float64 sums rounded once, so compare with rtol 1e-6.
    """
    r = np.asarray(ratings)
    pos = r > 0
    tot = pos.sum()
    gb = float(r[pos].astype(np.float64).sum() / tot) if tot else float("nan")
    rs = np.where(pos, r, 0).astype(np.float64)
    cu, ci = pos.sum(axis=1), pos.sum(axis=0)
    ub = np.where(cu > 0, rs.sum(axis=1) / np.maximum(cu, 1) - gb, 0.0)
    ib = np.where(ci > 0, rs.sum(axis=0) / np.maximum(ci, 1) - gb, 0.0)
    return gb, ub, ib


def fit_predict_svd_dense(ratings, k):
    """``RecommenderSystem('svd', use_sparse=False).fit(R, k).predict(R)`` in float64 (ground truth).

    centred = R - gb - ub[:,None] - ib over ALL cells (algo.py:235); rank-k reconstruction of the
    centred matrix plus the biases (algo.py:363-368).
    """
    r = np.asarray(ratings, dtype=np.float64)
    gb, ub, ib = svd_biases(ratings)
    centred = r - gb - ub[:, None] - ib[None, :]
    return reconstruct_f64(centred, k) + gb + ub[:, None] + ib[None, :]

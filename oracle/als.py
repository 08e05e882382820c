"""Oracle: CSR restatement of the reference's ALS sweeps (TEST INFRASTRUCTURE).

Follows /root/reference/src/recsys_lite/algo.py:300-337 (``RecommenderSystem._fit_als``):

  * init ``U = np.random.rand(n_users, f)`` then ``V = np.random.rand(n_items, f)`` from the
    GLOBAL NumPy RNG, float64 (algo.py:309-310);
  * any rating > 0 is an observation with value 1 (algo.py:311);
  * user sweep (algo.py:314-322): for every user with >= 1 observation
        x_u = solve(Y_S^T Y_S + lambda*I, Y_S^T 1)      Y_S = V[observed items, ascending]
    users without observations keep their previous row (algo.py:316-317);
  * item sweep (algo.py:324-332): same with roles swapped, using the ALREADY UPDATED U.

The reference needs a dense (n_users, n_items) matrix; here the observation pattern comes in
as CSR (rows = users) and CSR of the transpose (rows = items).  The arithmetic per row is the
same NumPy expression (same BLAS/LAPACK calls: dsyrk/dgemm, dgemv, dgesv), which is why the
result is bit-identical to the reference (checked in tests/test_oracle_pinning.py).

Third-party arithmetic: ``np.linalg.solve`` -> LAPACK dgesv (OpenBLAS bundled with numpy;
reference lock pins numpy 2.0.2/2.2.6, this image has 2.3.5).
"""

from __future__ import annotations

import numpy as np


def pattern_to_csr(ratings):
    """(indptr int64, indices int32) of the cells with rating > 0, rows ascending, cols ascending.

    Mirrors the boolean masks ``ratings[u] > 0`` (algo.py:315) / ``ratings[:, i] > 0`` (algo.py:325).
    Accepts a dense ndarray or any scipy.sparse matrix.
    """
    try:
        from scipy.sparse import issparse, csr_matrix
    except Exception:  # pragma: no cover
        issparse = lambda _x: False  # noqa: E731
    if issparse(ratings):
        m = csr_matrix(ratings, copy=True)
        m.sum_duplicates()
        m.data = (m.data > 0).astype(np.float32)
        m.eliminate_zeros()
        m.sort_indices()
        return m.indptr.astype(np.int64), m.indices.astype(np.int32)
    mask = np.asarray(ratings) > 0
    counts = mask.sum(axis=1)
    indptr = np.zeros(mask.shape[0] + 1, dtype=np.int64)
    np.cumsum(counts, out=indptr[1:])
    indices = np.nonzero(mask)[1].astype(np.int32)
    return indptr, indices


def transpose_pattern(indptr, indices, n_cols):
    """CSR of the transposed pattern; within each output row the indices ascend (stable)."""
    n_rows = len(indptr) - 1
    rows = np.repeat(np.arange(n_rows, dtype=np.int32), np.diff(indptr))
    order = np.argsort(indices, kind="stable")
    t_indices = rows[order]
    counts = np.bincount(indices, minlength=n_cols)
    t_indptr = np.zeros(n_cols + 1, dtype=np.int64)
    np.cumsum(counts, out=t_indptr[1:])
    return t_indptr, t_indices.astype(np.int32)


def half_step_csr(indptr, indices, other, mine, lam, rows=None):
    """One sweep (algo.py:314-322 or :324-332).  ``mine`` is updated IN PLACE (float64).

    other : (n_other, f) float64 -- the fixed factor matrix that is gathered
    mine  : (n_rows, f)  float64 -- rows with no observation are left untouched
    rows  : optional iterable restricting the sweep (used to time a bounded sample)
    """
    f = other.shape[1]
    it = range(len(indptr) - 1) if rows is None else rows
    for r in it:
        s, e = int(indptr[r]), int(indptr[r + 1])
        if e == s:
            continue
        g = other[indices[s:e]]
        rhs_w = np.ones(e - s, dtype=np.float32)  # ratings[u, rated] after binarisation
        mine[r] = np.linalg.solve(g.T @ g + lam * np.eye(f), g.T @ rhs_w)
    return mine


def init_factors(n_users, n_items, factors):
    """Draw order of algo.py:309-310 from the global RNG (caller seeds with np.random.seed)."""
    u = np.random.rand(n_users, factors)
    v = np.random.rand(n_items, factors)
    return u, v


def fit_csr(indptr_u, indices_u, indptr_i, indices_i, factors, iterations=10, lam=0.01,
            init=None):
    """Full ``_fit_als`` on a CSR pattern.  Returns the reference's model dict (algo.py:333-337)."""
    n_users, n_items = len(indptr_u) - 1, len(indptr_i) - 1
    if init is None:
        u, v = init_factors(n_users, n_items, factors)
    else:
        u, v = (np.array(init[0], dtype=np.float64, copy=True),
                np.array(init[1], dtype=np.float64, copy=True))
    for _ in range(iterations):
        half_step_csr(indptr_u, indices_u, v, u, lam)
        half_step_csr(indptr_i, indices_i, u, v, lam)
    return {"user_factors": u, "item_factors": v, "factors": factors}


def fit_dense(ratings, factors, iterations=10, lam=0.01, init=None):
    """Convenience: dense ratings -> CSR pattern -> fit_csr (same draws, same result)."""
    indptr_u, indices_u = pattern_to_csr(ratings)
    indptr_i, indices_i = transpose_pattern(indptr_u, indices_u, np.asarray(ratings).shape[1])
    return fit_csr(indptr_u, indices_u, indptr_i, indices_i, factors, iterations, lam, init)


def half_step_general(indptr, indices, other, mine, lam, w0=0.0, conf=None, pref=None):
    """General normal equations (SURVEY.md summary item 3) -- NOT reference parity when w0 != 0.

        A_r = w0*Y^T Y + sum_{i in S_r} (c_ri - w0) y_i y_i^T + lam*I
        b_r = sum_{i in S_r} c_ri * p_ri * y_i

    ``w0=0, conf=None, pref=None`` is the reference's arithmetic (up to summation order).
    The reference contains no Hu-Koren-Volinsky code; the ``w0=1`` branch is validated only
    against this small dense restatement and is labelled as such wherever it is reported.
    """
    f = other.shape[1]
    g0 = w0 * (other.T @ other) if w0 != 0.0 else np.zeros((f, f))
    for r in range(len(indptr) - 1):
        s, e = int(indptr[r]), int(indptr[r + 1])
        if e == s and w0 == 0.0:
            continue
        y = other[indices[s:e]]
        c = np.ones(e - s) if conf is None else np.asarray(conf[s:e], dtype=np.float64)
        p = np.ones(e - s) if pref is None else np.asarray(pref[s:e], dtype=np.float64)
        a = g0 + (y * (c - w0)[:, None]).T @ y + lam * np.eye(f)
        b = y.T @ (c * p)
        mine[r] = np.linalg.solve(a, b)
    return mine

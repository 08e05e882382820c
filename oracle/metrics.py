"""Oracle: RMSE / MAE over observed cells (TEST INFRASTRUCTURE).

Follows /root/reference/src/recsys_lite/algo.py:662-743 (``compute_rmse`` / ``compute_mae``):
shape check, inf / NaN rejection, error over cells with ``actual > 0``, 0.0 when there is none.
"""

from __future__ import annotations

import numpy as np


def _dense(x):
    return x.toarray() if hasattr(x, "toarray") else np.asarray(x)


def _check(pred, actual):
    if pred.shape != actual.shape:
        raise ValueError(
            "Predictions and actual matrices must have the same shape. "
            f"Got {pred.shape} and {actual.shape}."
        )
    p, a = _dense(pred), _dense(actual)
    if np.any(np.isinf(p)) or np.any(np.isinf(a)):
        raise ValueError("Input matrices must not contain infinite values.")
    if np.any(np.isnan(p)) or np.any(np.isnan(a)):
        raise ValueError("Input matrices must not contain NaN values.")
    return p, a


def rmse(pred, actual):
    p, a = _check(pred, actual)
    m = a > 0
    if not np.any(m):
        return 0.0
    d = p[m] - a[m]
    return float(np.sqrt(np.mean(d ** 2)))


def mae(pred, actual):
    p, a = _check(pred, actual)
    m = a > 0
    if not np.any(m):
        return 0.0
    return float(np.mean(np.abs(p[m] - a[m])))


def rmse_observed_csr(user_f, item_f, indptr, indices, values=None):
    """RMSE of ``U V^T`` over a CSR pattern without forming the score matrix (float64).

    Equals ``compute_rmse(U @ V.T, R)`` when R's positive cells are the pattern (values default 1).
    """
    user_f = np.asarray(user_f, dtype=np.float64)
    item_f = np.asarray(item_f, dtype=np.float64)
    nnz = int(indptr[-1])
    if nnz == 0:
        return 0.0
    rows = np.repeat(np.arange(len(indptr) - 1), np.diff(indptr))
    acc = 0.0
    step = 1 << 20
    for s in range(0, nnz, step):
        e = min(nnz, s + step)
        p = np.einsum("ij,ij->i", user_f[rows[s:e]], item_f[indices[s:e]])
        t = 1.0 if values is None else np.asarray(values[s:e], dtype=np.float64)
        acc += float(np.sum((p - t) ** 2))
    return float(np.sqrt(acc / nnz))

"""Oracle-side comparators with the tolerances stated once (TEST INFRASTRUCTURE).

Tolerances (BASELINE.json north_star):
  * ALS factors:  ||X - X_ref||_F / ||X_ref||_F <= 1e-4  and  max|X - X_ref| <= 1e-4 * max|X_ref|
  * RMSE on observed cells: |rmse - rmse_ref| <= 1e-5
  * top-N: index lists identical, except positions whose float64 reference scores are equal
    within ``TOPN_TIE_RTOL`` (4 * fp32 eps, relative to max(1,|score|)); inside such a tied set any
    permutation / substitution is accepted and counted.
"""

from __future__ import annotations

import numpy as np

FACTOR_RTOL = 1e-4
RMSE_ATOL = 1e-5
TOPN_TIE_RTOL = 4 * float(np.finfo(np.float32).eps)


def factor_error(x, x_ref):
    """(relative Frobenius error, max abs error / max|x_ref|)."""
    x = np.asarray(x, dtype=np.float64)
    x_ref = np.asarray(x_ref, dtype=np.float64)
    d = x - x_ref
    fro = float(np.linalg.norm(d) / max(np.linalg.norm(x_ref), 1e-300))
    mx = float(np.max(np.abs(d)) / max(np.max(np.abs(x_ref)), 1e-300)) if d.size else 0.0
    return fro, mx


def factors_match(x, x_ref, rtol=FACTOR_RTOL):
    fro, mx = factor_error(x, x_ref)
    return fro <= rtol and mx <= rtol


def topn_matches(idx_test, idx_ref, score_of_row, seen_of_row=None, rtol=TOPN_TIE_RTOL):
    """Compare two (rows, n) index matrices padded with -1.

    score_of_row(r) -> float64 vector of the exact scores of row r (all items);
    seen_of_row(r)  -> indices that must not appear (optional).
    Returns dict(exact=rows identical, tied=rows differing only inside tied score sets, bad=[row ids]).
    """
    idx_test = np.asarray(idx_test)
    idx_ref = np.asarray(idx_ref)
    res = {"exact": 0, "tied": 0, "bad": []}
    if idx_test.shape != idx_ref.shape:
        res["bad"] = list(range(max(idx_test.shape[0], idx_ref.shape[0])))
        return res
    for r in range(idx_ref.shape[0]):
        a, b = idx_test[r], idx_ref[r]
        if np.array_equal(a, b):
            res["exact"] += 1
            continue
        va, vb = a >= 0, b >= 0
        ok = np.array_equal(va, vb)
        if ok:
            aa, bb = a[va], b[vb]
            ok = len(set(aa.tolist())) == len(aa)
            if ok and seen_of_row is not None:
                ok = not np.isin(aa, np.asarray(seen_of_row(r))).any()
            if ok:
                s = np.asarray(score_of_row(r), dtype=np.float64)
                sa, sb = s[aa], s[bb]
                tol = rtol * np.maximum(1.0, np.abs(sb))
                ok = bool(np.all(np.abs(sa - sb) <= tol))
        if ok:
            res["tied"] += 1
        else:
            res["bad"].append(r)
    return res

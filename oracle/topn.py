"""Oracle: masked top-N selection (TEST INFRASTRUCTURE).

Follows /root/reference/src/recsys_lite/algo.py:533-659 (``_top_n_numba`` + ``top_n``) and the
``recommend`` call stack algo.py:172-208.

Semantics restated:
  * scores of items with ``known > 0`` are replaced by -inf (algo.py:567-569, :644);
  * ``n_actual = min(n, #items with known <= 0)`` (algo.py:571-575, :645-646);
  * the ``n_actual`` best indices, descending score, fill the row; other slots are -1 (algo.py:560,582);
  * the matrix is trimmed to the widest valid row (algo.py:636-639, :652-658);
  * every user fully rated -> shape (n_users, 0) (algo.py:621-623); zero-size -> (0, 0) (algo.py:616-618);
  * dtype: int32 on the reference's default (Numba) path (algo.py:560), ``int`` on the empty returns.

Tie order.  The reference's two own code paths order equal scores differently (SURVEY.md
Appendix A.6), so tie order is NOT part of the contract.  This oracle -- and the CUDA path --
use one documented rule: among equal scores the LOWER item index comes first.  Use
``oracle.compare.topn_matches`` to compare against the live reference; it accepts any
permutation/substitution inside a set of tied scores and nothing else.
"""

from __future__ import annotations

import numpy as np


def _dense(x):
    return x.toarray() if hasattr(x, "toarray") else np.asarray(x)


def select_row(scores, n_actual):
    """Indices of the n_actual largest scores: descending score, ascending index on ties."""
    if n_actual <= 0:
        return np.empty((0,), dtype=np.int64)
    scores = np.asarray(scores)
    if n_actual < scores.shape[0]:
        # cut first (argpartition), then widen to every score equal to the cut value so that the
        # tie rule is applied over the complete tied set
        kth = np.partition(scores, scores.shape[0] - n_actual)[scores.shape[0] - n_actual]
        cand = np.nonzero(scores >= kth)[0]
    else:
        cand = np.arange(scores.shape[0])
    order = np.lexsort((cand, -scores[cand]))
    return cand[order][:n_actual]


def top_n(est, known, n=10):
    """Restatement of ``top_n(est, known, n=n)`` (algo.py:587-659)."""
    if n <= 0:
        raise ValueError(f"n must be positive. Got n={n}.")
    est_d, known_d = _dense(est), _dense(known)
    if est_d.shape != known_d.shape:
        raise ValueError(
            "Estimated and known matrices must have the same shape. "
            f"Got {est_d.shape} and {known_d.shape}."
        )
    n_users, n_items = est_d.shape
    if n_users == 0 or n_items == 0:
        return np.empty((0, 0), dtype=int)
    seen = known_d > 0
    if np.all(seen):
        return np.empty((n_users, 0), dtype=int)
    out = np.full((n_users, n), -1, dtype=np.int32)
    widest = 0
    for r in range(n_users):
        row = np.array(est_d[r], dtype=np.float64, copy=True)
        row[seen[r]] = -np.inf
        n_actual = min(n, int(n_items - seen[r].sum()))
        if n_actual == 0:
            continue
        out[r, :n_actual] = select_row(row, n_actual)
        widest = max(widest, n_actual)
    if widest == 0:
        return np.empty((n_users, 0), dtype=int)
    return out[:, :widest]


def top_n_unmasked(est, n=10):
    """``recommend(..., exclude_rated=False)``: full argsort, int64 (algo.py:202-208).

    The reference takes ``argsort(pred)[:, -n:][:, ::-1]``; here ties use the documented rule.
    """
    est_d = _dense(est)
    n_users, n_items = est_d.shape
    take = min(n, n_items)
    out = np.empty((n_users, take), dtype=np.int64)
    for r in range(n_users):
        out[r] = select_row(np.asarray(est_d[r], dtype=np.float64), take)
    return out


def recommend_blocked(user_f, item_f, indptr, indices, n=10, users=None, block=2048,
                      exclude_rated=True, bias=None):
    """Row-blocked ``predict`` + ``top_n`` (algo.py:143-147 + :172-201) without the full score matrix.

    user_f (n_users,f), item_f (n_items,f): scored in float64 exactly like ``U @ V.T``.
    indptr/indices: CSR pattern of the rated items (what ``known > 0`` would mask).
    users: optional array of user rows to score (bounded samples for big configs).
    bias: optional (global, user_bias (n_users,), item_bias (n_items,)) -- SVD class path (algo.py:363-368).
    Returns (idx int32 (len(users), n) padded with -1, scores float64 same shape padded with -inf).
    """
    user_f = np.asarray(user_f, dtype=np.float64)
    item_f = np.asarray(item_f, dtype=np.float64)
    n_items = item_f.shape[0]
    users = np.arange(user_f.shape[0]) if users is None else np.asarray(users)
    idx = np.full((len(users), n), -1, dtype=np.int32)
    val = np.full((len(users), n), -np.inf, dtype=np.float64)
    for b0 in range(0, len(users), block):
        rows = users[b0:b0 + block]
        s = user_f[rows] @ item_f.T
        if bias is not None:
            s = s + bias[0] + np.asarray(bias[1], dtype=np.float64)[rows][:, None] \
                + np.asarray(bias[2], dtype=np.float64)[None, :]
        for j, r in enumerate(rows):
            row = s[j]
            n_seen = 0
            if exclude_rated:
                seg = indices[int(indptr[r]):int(indptr[r + 1])]
                row[seg] = -np.inf
                n_seen = len(seg)
            n_actual = min(n, n_items - n_seen)
            pick = select_row(row, n_actual)
            idx[b0 + j, :n_actual] = pick
            val[b0 + j, :n_actual] = row[pick]
    return idx, val

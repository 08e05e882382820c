"""Seeded synthetic inputs for the five BASELINE.json configs (bench / test inputs; no arithmetic of the path).

SURVEY.md section 8(d) table.  All generators are deterministic functions of their seed and produce host
NumPy arrays; nothing here touches a GPU.

  C1  1 000 x 1 000 dense binary, Bernoulli(0.05), RandomState(0); k=16, 10 iterations, lambda=0.01,
      factor init ``np.random.seed(123)`` then U, V drawn in that order (reference algo.py:309-310).
  C2  10 000 x 500: (a) README shape ``np.random.seed(0); rand(10000, 500)`` (reference README.md:111-116),
      (b) ``create_sample_ratings(10000, 500, sparsity=0.9, random_state=0)`` restated from
      reference src/recsys_lite/io.py:688-755.
  C3  100 000 x 20 000, ~5 M nnz, k=64      seed 3 (skewed variant seed 13)
  C4  1 000 000 x 100 000, ~50 M nnz, k=64  seed 4 (skewed variant seed 14)
  C5  10 000 000 x 1 000 000, ~500 M nnz, k=128  seed 5
CSR patterns: per-user degree ~ Poisson(mean) clipped to >= 1, item ids uniform (or Zipf(1.0) with
log-normal degrees in the skewed variant), duplicates within a row dropped (so nnz is ~0.03 % below
the nominal figure), indices ascending, int32.
"""

from __future__ import annotations

import numpy as np

CONFIGS = {
    "C1": dict(n_users=1_000, n_items=1_000, k=16, iterations=10, lam=0.01, seed=0, init_seed=123),
    "C2": dict(n_users=10_000, n_items=500, k=10, n=10, seed=0),
    "C3": dict(n_users=100_000, n_items=20_000, mean_deg=50, k=64, lam=0.01, seed=3, init_seed=3),
    "C4": dict(n_users=1_000_000, n_items=100_000, mean_deg=50, k=64, lam=0.01, seed=4, init_seed=4, n=10),
    "C5": dict(n_users=10_000_000, n_items=1_000_000, mean_deg=50, k=128, lam=0.01, seed=5, init_seed=5),
}


def c1_matrix(n_users=1000, n_items=1000, density=0.05, seed=0):
    rng = np.random.RandomState(seed)
    return (rng.random((n_users, n_items)) < density).astype(np.float32)


def c2_readme_matrix(n_users=10_000, n_items=500, seed=0):
    np.random.seed(seed)
    return np.random.rand(n_users, n_items)


def sample_ratings(n_users=100, n_items=50, sparsity=0.8, rating_range=(1.0, 5.0), random_state=None):
    """Restatement of ``create_sample_ratings`` (io.py:688-755): uniform ratings, Bernoulli zero mask."""
    rng = np.random.RandomState(random_state)
    m = rng.uniform(low=rating_range[0], high=rating_range[1], size=(n_users, n_items)).astype(np.float32)
    if sparsity > 0:
        m[rng.random((n_users, n_items)) < sparsity] = 0.0
    return m


def synth_pattern(n_users, n_items, mean_deg=50, seed=0, skew=False, row_offset=0):
    """CSR pattern (indptr int64, indices int32) of a synthetic implicit-feedback matrix.

    row_offset only changes the RNG stream (seed, row_offset) so shards can be generated per rank.
    """
    rng = np.random.default_rng([seed, row_offset])
    if skew:
        deg = rng.lognormal(mean=np.log(35.0), sigma=1.0, size=n_users)
        deg = np.maximum(1, np.rint(deg * (mean_deg * n_users / deg.sum()))).astype(np.int64)
        deg = np.minimum(deg, n_items)
    else:
        deg = np.maximum(1, rng.poisson(mean_deg, size=n_users)).astype(np.int64)
    total = int(deg.sum())
    rows = np.repeat(np.arange(n_users, dtype=np.int64), deg)
    if skew:
        # Zipf(s=1) popularity over a fixed random permutation of the item ids
        w = 1.0 / np.arange(1, n_items + 1, dtype=np.float64)
        cdf = np.cumsum(w / w.sum())
        ranks = np.minimum(np.searchsorted(cdf, rng.random(total)), n_items - 1)
        cols = rng.permutation(n_items)[ranks].astype(np.int64)
    else:
        cols = rng.integers(0, n_items, size=total, dtype=np.int64)
    key = rows * n_items + cols
    key.sort()
    keep = np.ones(total, dtype=bool)
    keep[1:] = key[1:] != key[:-1]
    key = key[keep]
    rows_k = key // n_items
    indices = (key - rows_k * n_items).astype(np.int32)
    counts = np.bincount(rows_k, minlength=n_users)
    indptr = np.zeros(n_users + 1, dtype=np.int64)
    np.cumsum(counts, out=indptr[1:])
    return indptr, indices


def seeded_factors(n_users, n_items, k, init_seed):
    """``np.random.seed(s)``; U then V, float64 -- the reference's draw order (algo.py:309-310)."""
    np.random.seed(init_seed)
    u = np.random.rand(n_users, k)
    v = np.random.rand(n_items, k)
    return u, v


def pattern_to_dense(indptr, indices, n_items):
    n_users = len(indptr) - 1
    r = np.zeros((n_users, n_items), dtype=np.float32)
    rows = np.repeat(np.arange(n_users), np.diff(indptr))
    r[rows, indices] = 1.0
    return r


def sharded_problem(n_users, n_items, mean_deg, seed, rank, world, n_blocks=None):
    """Rank `rank`'s share of a problem too large to build whole on every rank (C5: 10 M x 1 M, 500 M nnz): the user rows are
    generated in `n_blocks` (default `world`) independent blocks (`synth_pattern(..., row_offset=block)`), every rank walks all
    blocks, keeps its own as the user-side shard and, of the others, only the entries whose item falls into its item range --
    those become its rows of the transposed pattern.  Returns (user_bounds, item_bounds, up, ux, ip, ix): row boundaries of all
    ranks and this rank's two CSR shards with local indptr (ux: global item ids, ix: global user ids, ascending)."""
    from scipy import sparse
    n_blocks = world if n_blocks is None else n_blocks
    assert n_blocks % world == 0
    ub = (np.arange(n_blocks + 1, dtype=np.int64) * n_users) // n_blocks          # block boundaries (users)
    user_bounds = ub[:: n_blocks // world].copy()
    item_bounds = (np.arange(world + 1, dtype=np.int64) * n_items) // world
    ilo, ihi = int(item_bounds[rank]), int(item_bounds[rank + 1])
    mine_ptr, mine_idx, t_rows, t_cols = [], [], [], []
    for b in range(n_blocks):
        nb = int(ub[b + 1] - ub[b])
        ptr, idx = synth_pattern(nb, n_items, mean_deg, seed, row_offset=b)
        if user_bounds[rank] <= ub[b] < user_bounds[rank + 1]:
            mine_ptr.append(ptr)
            mine_idx.append(idx)
        keep = (idx >= ilo) & (idx < ihi)
        rows = np.repeat(np.arange(nb, dtype=np.int64), np.diff(ptr))[keep] + ub[b]
        t_rows.append(rows.astype(np.int32))
        t_cols.append((idx[keep] - ilo).astype(np.int32))
        del ptr, idx, keep, rows
    up = np.zeros(1, dtype=np.int64)
    for ptr in mine_ptr:
        up = np.concatenate([up, ptr[1:] + up[-1]])
    ux = np.concatenate(mine_idx)
    rows, cols = np.concatenate(t_rows), np.concatenate(t_cols)
    # transposed shard: rows = this rank's items, indices = global user ids ascending (users were walked in ascending order)
    t = sparse.csr_matrix((np.ones(len(rows), dtype=np.int8), (cols, rows)), shape=(ihi - ilo, n_users))
    t.sort_indices()
    return (user_bounds, item_bounds, np.ascontiguousarray(up), np.ascontiguousarray(ux, dtype=np.int32),
            np.ascontiguousarray(t.indptr, dtype=np.int64), np.ascontiguousarray(t.indices, dtype=np.int32))

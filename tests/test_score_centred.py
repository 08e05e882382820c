"""-m gpu tests of the centred single-product tensor-core scoring pass (recsys_lite_b200/csrc/score_tc2.cu):
the preprocessing (centring, norm classes, stable placement), the error bound the selection relies on, and
bit-identity with the exact kernel and the oracle on random, clustered ("converged") and degenerate factors.
Reference semantics: recommend = predict + top_n, algo.py:172-201 / :533-659."""
import ctypes as C
import struct

import numpy as np
import pytest

from oracle import als as o_als, compare, generators as gen, topn as o_topn

pytestmark = pytest.mark.gpu

EXACT, SPLIT3, CENTRED = 1, 2, 4
NCLS = 8


@pytest.fixture(scope="module")
def h():
    from recsys_lite_b200 import _ffi
    return _ffi.default_handle()


def _f32(x):
    return np.asarray(x).astype(np.float32).astype(np.float64)


def _clustered(rng, n_users, n_items, k, spread=0.01, outliers=0.08):
    """ALS-like converged factors: items around a common vector (a few per cent of them far from it), users positive."""
    vbar = rng.random(k) * 0.5 + 0.2
    v = vbar + rng.standard_normal((n_items, k)) * spread
    far = rng.random(n_items) < outliers
    v[far] += rng.standard_normal((int(far.sum()), k)) * rng.choice([0.05, 0.2, 1.0], size=(int(far.sum()), 1))
    u = rng.random((n_users, k)) * 0.06 + rng.standard_normal((n_users, k)) * 0.01
    return _f32(u), _f32(v)


def _parse_table(raw):
    ints = struct.unpack("11i", raw[:44])
    wc = struct.unpack("8f", raw[44:76])
    rc = struct.unpack("8f", raw[76:108])
    scale, vmax_raw = struct.unpack("2f", raw[108:116])
    return dict(first_tile=ints[:9], n_tiles=ints[9], wc=np.array(wc), rc=np.array(rc), scale=scale, vmax_raw=vmax_raw)


@pytest.mark.parametrize("k,kind", [(64, "clustered"), (40, "random"), (16, "clustered"), (64, "wide")])
def test_centred_raw_scores_and_class_table(h, k, kind):
    """rl_score_dump_centred_dev: (a) the position -> id map is a stable class-wise arrangement of all items, padded
    to whole tiles; (b) the class table bounds |w_i| s and the fp16 residual of every member; (c) every raw accumulator
    is within the window the kernel computes of s_u s u.(v_i - vbar)."""
    from devbuf import Dev, padded
    n_users, n_items = 300, 3000
    rng = np.random.default_rng(k + len(kind))
    if kind == "clustered":
        uf, vf = _clustered(rng, n_users, n_items, k)
    elif kind == "random":
        uf, vf = _f32(rng.standard_normal((n_users, k))), _f32(rng.standard_normal((n_items, k)))
    else:                                               # wide dynamic range across items and inside rows
        uf = _f32(rng.standard_normal((n_users, k)) * np.exp(rng.uniform(-10, 10, (n_users, 1))))
        vf = _f32(1.0 + rng.standard_normal((n_items, k)) * np.exp(rng.uniform(-12, 0, (n_items, 1))))
    ld = ((n_items + 127) // 128 + NCLS) * 128
    d_u, d_v = padded(h, uf, k), padded(h, vf, k)
    d_out = Dev(h, shape=(n_users, ld), dtype=np.float32)
    d_su = Dev(h, shape=(n_users,), dtype=np.float32)
    d_perm = Dev(h, shape=(ld,), dtype=np.int32)
    d_tab = Dev(h, shape=(256,), dtype=np.uint8)
    h.call("rl_score_dump_centred_dev", d_u.ptr, n_users, d_v.ptr, n_items, k, d_out.ptr, ld, d_su.ptr, d_perm.ptr, d_tab.ptr)
    raw, su, perm = d_out.get().astype(np.float64), d_su.get().astype(np.float64), d_perm.get()
    tab = _parse_table(d_tab.get().tobytes())
    n_pos = tab["n_tiles"] * 128
    assert n_pos <= ld and (perm[n_pos:] == -1).all()
    # (a) permutation, ascending ids inside each class, classes in whole tiles
    ids = perm[:n_pos]
    assert sorted(ids[ids >= 0].tolist()) == list(range(n_items))
    cls_of_pos = np.zeros(n_pos, dtype=int)
    for c in range(NCLS):
        lo, hi = tab["first_tile"][c] * 128, tab["first_tile"][c + 1] * 128
        cls_of_pos[lo:hi] = c
        seg = ids[lo:hi]
        real = seg[seg >= 0]
        assert (np.diff(real) > 0).all()
        assert (seg[len(real):] == -1).all()            # padding only at the end of a class
    # (b) class bounds
    s = tab["scale"]
    vbar = vf.mean(axis=0).astype(np.float32)
    w = (vf.astype(np.float32) - vbar).astype(np.float32).astype(np.float64)
    slack = np.abs(uf).sum(axis=1) * np.abs(vbar).max() * 2.0 ** -22   # the kernel's vbar may differ by an ulp per entry
    wn = np.linalg.norm(w, axis=1) * s
    wh = (w * s).astype(np.float16).astype(np.float64)
    res = np.linalg.norm(w * s - wh, axis=1)
    real_pos = np.nonzero(ids >= 0)[0]
    c_of_item = np.empty(n_items, dtype=int)
    c_of_item[ids[real_pos]] = cls_of_pos[real_pos]
    assert (wn <= tab["wc"][c_of_item] * 1.001 + np.abs(vbar).max() * s * 2.0 ** -20).all()
    assert (res <= tab["rc"][c_of_item] * 1.05 + 1e-3).all()
    sizes = np.diff(tab["first_tile"])
    wcs = np.sort(tab["wc"][sizes > 0])                             # factor-of-two norm bins, each at most once; their scan
    assert (wcs[1:] > wcs[:-1]).all()                               # order follows the sampled top-n density (class_scan_kernel)
    # (c) error bound per (user, item), in raw units, as the selection warps compute it
    us = uf * su[:, None]
    uh = us.astype(np.float16).astype(np.float64)
    e1, e2, e3 = np.linalg.norm(us - uh, axis=1), np.linalg.norm(uh, axis=1), np.linalg.norm(us, axis=1)
    acc = 64 * 2.0 ** -23
    eps_c = (e1[:, None] * tab["wc"] + e2[:, None] * tab["rc"] + acc * e2[:, None] * (tab["wc"] + tab["rc"]) +
             e3[:, None] * (6.1e-8 * tab["wc"] + 6.0e-14 * tab["vmax_raw"]))        # (users, classes)
    true = (us @ w.T) * s                                                              # (users, items)
    got = raw[:, real_pos]
    err = np.abs(got - true[:, ids[real_pos]])
    bound = eps_c[:, cls_of_pos[real_pos]] + (slack * su * s)[:, None]
    ratio = float(np.max(err / np.maximum(bound, 1e-300)))
    print(f"centred raw-score error / window: k={k} {kind}: {ratio:.3f}; classes (tiles) {np.diff(tab['first_tile'])}")
    assert ratio <= 1.0
    assert not raw[:, np.nonzero(ids < 0)[0]].any()     # padding positions are zero rows


def _run(h, uf, vf, indptr, indices, k, n, mode, want_scores=True):
    from devbuf import Dev, padded
    n_users = uf.shape[0]
    d_u, d_v = padded(h, uf, k), padded(h, vf, k)
    d_ip = Dev(h, indptr) if indptr is not None else None
    d_ix = Dev(h, indices) if indices is not None else None
    d_idx = Dev(h, shape=(n_users, n), dtype=np.int32)
    d_val = Dev(h, shape=(n_users, n), dtype=np.float64) if want_scores else None
    h.call("rl_score_topn_dev", d_u.ptr, n_users, d_v.ptr, vf.shape[0], k, d_ip.ptr if d_ip else None, d_ix.ptr if d_ix else None,
           0.0, None, None, n, d_idx.ptr, d_val.ptr if d_val else None, mode)
    return d_idx.get(), (d_val.get() if d_val else None), int(h.lib.rl_last_overflow_rows(h.h))


@pytest.mark.parametrize("k,n,n_users,n_items,kind", [
    (64, 10, 1000, 5000, "random"), (64, 10, 1500, 6000, "clustered"), (64, 1, 130, 1024, "random"),
    (32, 12, 257, 3000, "clustered"), (40, 10, 500, 2049, "random"), (16, 16, 300, 1200, "clustered"),
    (64, 10, 129, 20000, "clustered"), (64, 10, 40000, 1500, "random")])
def test_centred_topn_equals_exact_kernel(h, k, n, n_users, n_items, kind):
    """mode RL_SCORE_TENSOR_CENTRED must return exactly what RL_SCORE_EXACT returns (indices and float64 scores),
    both must match the oracle (algo.py:172-201) up to exact-score ties, and the window must be tight."""
    rng = np.random.default_rng(n_users + k)
    if kind == "random":
        uf, vf = _f32(rng.random((n_users, k)) * 0.4), _f32(rng.random((n_items, k)) * 0.4)
    else:
        uf, vf = _clustered(rng, n_users, n_items, k)
    indptr, indices = gen.synth_pattern(n_users, n_items, mean_deg=45, seed=n_items)
    ex_idx, ex_val, _ = _run(h, uf, vf, indptr, indices, k, n, EXACT)
    c_idx, c_val, over = _run(h, uf, vf, indptr, indices, k, n, CENTRED)
    assert np.array_equal(ex_idx, c_idx)
    assert np.array_equal(ex_val, c_val)
    assert over <= max(2, n_users // 1000), over
    a_idx, _, _ = _run(h, uf, vf, indptr, indices, k, n, 0)          # AUTO takes the same path
    assert np.array_equal(ex_idx, a_idx)
    if n_users <= 2000:
        ref_idx, _ = o_topn.recommend_blocked(uf, vf, indptr, indices, n=n)
        res = compare.topn_matches(c_idx, ref_idx, lambda row: uf[row] @ vf.T, lambda row: indices[indptr[row]:indptr[row + 1]])
        assert not res["bad"], res


def test_centred_topn_ragged_ties_and_no_seen_list(h):
    """users who have seen almost everything (ragged -1 rows), massive exact ties (candidate lists spill to pool blocks or
    go to the exact-kernel fallback), identical item vectors (zero-width classes) and no seen list at all."""
    k, n, n_users, n_items = 64, 10, 260, 2100
    rng = np.random.default_rng(11)
    uf = _f32(rng.integers(0, 2, (n_users, k)) * 0.5)
    vf = _f32(rng.integers(0, 2, (n_items, k)) * 0.5)
    dense = rng.random((n_users, n_items)) < 0.02
    dense[1, :] = True
    dense[2, :2095] = True
    dense[2, 2095:] = False
    indptr, indices = o_als.pattern_to_csr(dense.astype(np.float32))
    ex, _, _ = _run(h, uf, vf, indptr, indices, k, n, EXACT, False)
    ce, _, over = _run(h, uf, vf, indptr, indices, k, n, CENTRED, False)
    assert np.array_equal(ex, ce)
    assert (ce[1] == -1).all() and (ce[2, 5:] == -1).all() and (ce[2, :5] >= 2095).all()
    ce2, _, _ = _run(h, uf, vf, None, None, k, n, CENTRED, False)
    ref, _ = o_topn.recommend_blocked(uf, vf, indptr, indices, n=n, exclude_rated=False)
    assert np.array_equal(ce2, ref)
    vsame = np.tile(vf[:1], (n_items, 1))                             # every item identical: all scores tie, w = 0
    ex3, _, _ = _run(h, uf, vsame, indptr, indices, k, n, EXACT, False)
    ce3, _, _ = _run(h, uf, vsame, indptr, indices, k, n, CENTRED, False)
    assert np.array_equal(ex3, ce3)


def test_centred_topn_spilled_candidate_lists(h):
    """120 items share the best vector of every user: the candidate list (40 slots in shared memory) runs over and moves
    to pool blocks in global memory (up to 240 candidates) -- same result as the exact kernel (ties by ascending item id)
    without the exact-kernel fallback."""
    k, n, n_users, n_items = 64, 10, 700, 4000
    rng = np.random.default_rng(12)
    uf = _f32(rng.random((n_users, k)))
    vf = _f32(rng.random((n_items, k)))
    dup = rng.choice(n_items, 120, replace=False)
    vf[dup] = vf[dup[0]] * 2.0
    dense = rng.random((n_users, n_items)) < 0.01
    indptr, indices = o_als.pattern_to_csr(dense.astype(np.float32))
    ex, exv, _ = _run(h, uf, vf, indptr, indices, k, n, EXACT, True)
    ce, cev, over = _run(h, uf, vf, indptr, indices, k, n, CENTRED, True)
    assert np.array_equal(ex, ce) and np.array_equal(exv, cev)
    assert np.isin(ce, dup).all()
    assert over == 0


def test_centred_topn_on_converged_als_factors(h):
    """>= 15 ALS iterations on a 200K x 20K pattern (the regime the bench lives in: every score drifts towards 1,
    the top ten sit ~1e-5 apart): centred tensor path == exact kernel on all users, == float64 oracle ranking on
    sampled users, and (almost) no row needs the exact-kernel fallback."""
    from devbuf import Dev, padded, unpadded
    n_users, n_items, k, n, iters = 200_000, 20_000, 64, 10, 16
    indptr, indices = gen.synth_pattern(n_users, n_items, mean_deg=50, seed=21)
    tp, tx = o_als.transpose_pattern(indptr, indices, n_items)
    rng = np.random.default_rng(21)
    d_u = padded(h, rng.random((n_users, k), dtype=np.float32), k)
    d_v = padded(h, rng.random((n_items, k), dtype=np.float32), k)
    d_ip, d_ix, d_tp, d_tx = Dev(h, indptr), Dev(h, indices), Dev(h, tp), Dev(h, tx)
    for _ in range(iters):
        h.call("rl_als_half_step_dev", n_users, n_items, k, d_ip.ptr, d_ix.ptr, None, None, d_v.ptr, d_u.ptr, 0.01, 0.0, None, 0, 1, None)
        h.call("rl_als_half_step_dev", n_items, n_users, k, d_tp.ptr, d_tx.ptr, None, None, d_u.ptr, d_v.ptr, 0.01, 0.0, None, 0, 1, None)
    res = {}
    for mode in (EXACT, CENTRED):
        d_idx = Dev(h, shape=(n_users, n), dtype=np.int32)
        d_val = Dev(h, shape=(n_users, n), dtype=np.float64)
        h.call("rl_score_topn_dev", d_u.ptr, n_users, d_v.ptr, n_items, k, d_ip.ptr, d_ix.ptr, 0.0, None, None, n, d_idx.ptr,
               d_val.ptr, mode)
        res[mode] = (d_idx.get(), d_val.get(), int(h.lib.rl_last_overflow_rows(h.h)))
    assert np.array_equal(res[EXACT][0], res[CENTRED][0])
    assert np.array_equal(res[EXACT][1], res[CENTRED][1])
    over = res[CENTRED][2]
    print(f"converged factors ({iters} iterations): overflow rows {over} of {n_users}")
    assert over < n_users // 1000                                     # < 0.1 %
    uf, vf = unpadded(h, d_u, k, np.float64), unpadded(h, d_v, k, np.float64)
    rows = np.random.default_rng(5).choice(n_users, 300, replace=False)
    sub_ptr = np.zeros(len(rows) + 1, dtype=np.int64)
    sub_ix = []
    for i, r in enumerate(rows):
        sub_ix.append(indices[indptr[r]:indptr[r + 1]])
        sub_ptr[i + 1] = sub_ptr[i] + len(sub_ix[-1])
    sub_ix = np.concatenate(sub_ix).astype(np.int32)
    ref_idx, _ = o_topn.recommend_blocked(uf[rows], vf, sub_ptr, sub_ix, n=n)
    chk = compare.topn_matches(res[CENTRED][0][rows], ref_idx, lambda i: uf[rows[i]] @ vf.T,
                               lambda i: sub_ix[sub_ptr[i]:sub_ptr[i + 1]])
    assert not chk["bad"], chk
    spread = float(np.std(uf[rows[0]] @ vf.T))
    print(f"score spread of one user {spread:.2e}; oracle rows exact {chk['exact']} tied {chk['tied']}")

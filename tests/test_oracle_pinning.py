"""Pin the CPU oracle (oracle/) against the reference: committed golden outputs + live runs.

Golden files come from tests/golden/make_golden.py (the unmodified reference run in the build
container).  Live tests additionally require BITWISE equality for ALS when /root/reference is
importable.  KATs cite the reference test they restate.
"""
import os

import numpy as np
import pytest

from oracle import als, compare, generators as gen, metrics, svd, topn


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"))


def _ratings(g):
    shape = tuple(int(x) for x in g["shape"])
    return np.unpackbits(g["ratings"], axis=1)[:, : shape[1]].astype(np.float32)


@pytest.mark.parametrize("name", ["als_small_k8", "als_small_k64", "als_test_literal", "als_c1"])
def test_als_oracle_matches_golden(golden_dir, name):
    g = _load(golden_dir, name)
    r = _ratings(g)
    np.random.seed(int(g["init_seed"]))
    m = als.fit_dense(r, int(g["k"]), int(g["iterations"]), float(g["lam"]))
    # same arithmetic, possibly another CPU's BLAS kernels: 1e-8 (bitwise on the generating host)
    for key in ("user_factors", "item_factors"):
        fro, mx = compare.factor_error(m[key], g[key])
        assert fro < 1e-8 and mx < 1e-8, (key, fro, mx)
    pred = m["user_factors"] @ m["item_factors"].T
    assert abs(metrics.rmse(pred, r) - float(g["rmse"])) < 1e-9
    assert abs(metrics.rmse_observed_csr(m["user_factors"], m["item_factors"],
                                         *als.pattern_to_csr(r)) - float(g["rmse"])) < 1e-9


@pytest.mark.parametrize("name", ["als_small_k8", "als_small_k64", "als_c1"])
def test_recommend_oracle_matches_golden(golden_dir, name):
    g = _load(golden_dir, name)
    r = _ratings(g)
    u, v = g["user_factors"], g["item_factors"]
    indptr, indices = als.pattern_to_csr(r)
    idx, _ = topn.recommend_blocked(u, v, indptr, indices, n=10)
    ref_idx = g["recommend"]
    width = ref_idx.shape[1]
    assert (idx[:, width:] == -1).all()
    res = compare.topn_matches(idx[:, :width], ref_idx, lambda row: u[row] @ v.T,
                               lambda row: indices[indptr[row]:indptr[row + 1]])
    assert not res["bad"], res
    # and through the dense top_n restatement (same shape/dtype contract as the reference)
    out = topn.top_n(u @ v.T, r, n=10)
    assert out.shape == ref_idx.shape and out.dtype == ref_idx.dtype
    assert not compare.topn_matches(out, ref_idx, lambda row: u[row] @ v.T)["bad"]


def test_als_oracle_bitwise_vs_live_reference(live_reference):
    rng = np.random.RandomState(3)
    r = (rng.random((150, 90)) < 0.12).astype(np.float32)
    r[3, :] = 0
    r[:, 4] = 0
    np.random.seed(5)
    model = live_reference.RecommenderSystem(algorithm="als").fit(r, k=12, iterations=3, lambda_=0.01)
    np.random.seed(5)
    m = als.fit_dense(r, 12, 3, 0.01)
    assert np.array_equal(m["user_factors"], model._model["user_factors"])
    assert np.array_equal(m["item_factors"], model._model["item_factors"])


def test_transpose_pattern_matches_scipy():
    from scipy.sparse import csr_matrix
    indptr, indices = gen.synth_pattern(300, 70, mean_deg=6, seed=9)
    m = csr_matrix((np.ones(len(indices), np.float32), indices, indptr), shape=(300, 70))
    t = m.T.tocsr()
    t.sort_indices()
    ti, tj = als.transpose_pattern(indptr, indices, 70)
    assert np.array_equal(ti, t.indptr) and np.array_equal(tj, t.indices)
    # pattern_to_csr round trip from dense and from sparse
    d = gen.pattern_to_dense(indptr, indices, 70)
    a = als.pattern_to_csr(d)
    b = als.pattern_to_csr(m)
    assert np.array_equal(a[0], indptr) and np.array_equal(a[1], indices)
    assert np.array_equal(b[0], indptr) and np.array_equal(b[1], indices)


def test_general_half_step_reduces_to_reference():
    indptr, indices = gen.synth_pattern(40, 30, mean_deg=5, seed=2)
    u, v = gen.seeded_factors(40, 30, 6, 1)
    a = als.half_step_csr(indptr, indices, v, u.copy(), 0.01)
    b = als.half_step_general(indptr, indices, v, u.copy(), 0.01, w0=0.0)
    assert compare.factor_error(b, a)[0] < 1e-10


@pytest.mark.parametrize("case", ["lit", "order", "ragged", "rand_f64", "rand_f32"])
def test_topn_oracle_matches_golden(golden_dir, case):
    g = _load(golden_dir, "topn_cases")
    est, known, n, ref_out = g[case + "_est"], g[case + "_known"], int(g[case + "_n"]), g[case + "_out"]
    out = topn.top_n(est, known, n=n)
    assert out.shape == ref_out.shape and out.dtype == ref_out.dtype
    res = compare.topn_matches(out, ref_out, lambda row: est[row].astype(np.float64),
                               lambda row: np.nonzero(known[row] > 0)[0], rtol=0.0)
    assert not res["bad"], res
    if case != "ragged":
        assert np.array_equal(out, ref_out)   # distinct scores -> identical


def test_topn_reference_kats():
    # tests/test_algo.py:206-218 -- exact order for distinct scores
    est = np.array([[0.1, 0.5, 0.9, 0.7]], dtype=np.float32)
    known = np.array([[1, 0, 0, 0]], dtype=np.float32)
    assert topn.top_n(est, known, n=3).tolist() == [[2, 3, 1]]
    # tests/test_algo.py:774-780 -- all rated -> (n_users, 0)
    assert topn.top_n(np.ones((2, 3)), np.ones((2, 3)), n=2).shape == (2, 0)
    # tests/test_algo.py:183-189 -- empty
    assert topn.top_n(np.empty((0, 0)), np.empty((0, 0)), n=2).shape == (0, 0)
    with pytest.raises(ValueError, match="n must be positive"):
        topn.top_n(est, known, n=0)
    with pytest.raises(ValueError, match="must have the same shape"):
        topn.top_n(est, np.ones((2, 4)), n=1)
    # documented tie rule: lower index first (SURVEY Appendix A.6 input)
    est = np.array([[1, 2, 2, 2, .5, 2, 3]], dtype=np.float64)
    assert topn.top_n(est, np.zeros_like(est), n=3).tolist() == [[6, 1, 2]]


def test_topn_oracle_vs_live_reference(live_reference):
    rng = np.random.RandomState(21)
    est = rng.standard_normal((50, 40))
    known = (rng.random((50, 40)) < 0.3).astype(np.float64)
    known[7, :] = 1
    known[9, :38] = 1
    ref_out = live_reference.top_n(est, known, n=6)
    out = topn.top_n(est, known, n=6)
    assert out.shape == ref_out.shape and out.dtype == ref_out.dtype
    assert np.array_equal(out, ref_out)


@pytest.mark.parametrize("case", ["lit", "rank3", "rnd", "samp"])
def test_svd_oracle_matches_golden(golden_dir, case):
    g = _load(golden_dir, "svd_cases")
    m, k, ref_out = g[case + "_in"], int(g[case + "_k"]), g[case + "_out"]
    out = svd.svd_reconstruct_dense(m, k)
    assert out.dtype == np.float32 and out.shape == ref_out.shape
    assert np.max(np.abs(out - ref_out)) < 2e-5 * max(1.0, np.abs(ref_out).max())
    # float64 ground truth agrees with fp32 LAPACK to fp32 accuracy
    assert np.max(np.abs(svd.reconstruct_f64(m, k) - ref_out)) < 2e-5 * max(1.0, np.abs(ref_out).max())


def test_svd_quality_kat():
    # tests/test_algo.py:88-107 -- an exactly rank-3 matrix is reproduced (max |diff| < 1e-6 scale)
    rr = np.random.RandomState(0)
    m = (rr.standard_normal((10, 3)) @ rr.standard_normal((3, 8))).astype(np.float32)
    out = svd.svd_reconstruct_dense(m, 3)
    assert np.max(np.abs(out - m)) < 1e-5
    with pytest.raises(ValueError, match="k must be between"):
        svd.svd_reconstruct_dense(m, 0)
    with pytest.raises(ValueError, match="must be 2D"):
        svd.svd_reconstruct_dense(np.ones(3), 1)
    assert not svd.svd_reconstruct_dense(np.zeros((3, 4), np.float32), 2).any()


def test_svd_class_path_matches_golden(golden_dir):
    g = _load(golden_dir, "svd_class")
    m, k = g["ratings"], int(g["k"])
    gb, ub, ib = svd.svd_biases(m)
    assert abs(gb - float(g["global_bias"])) < 1e-5
    assert np.allclose(ub, g["user_bias"], atol=1e-5) and np.allclose(ib, g["item_bias"], atol=1e-5)
    pred = svd.fit_predict_svd_dense(m, k)
    assert np.max(np.abs(pred - g["pred"])) < 5e-5
    out = topn.top_n(pred, m, n=5)
    assert not compare.topn_matches(out, g["recommend"], lambda row: pred[row], rtol=1e-5)["bad"]


def test_metrics_match_golden_and_kats(golden_dir):
    g = _load(golden_dir, "metrics")
    assert abs(metrics.rmse(g["pred"], g["actual"]) - float(g["rmse"])) < 1e-6
    assert abs(metrics.mae(g["pred"], g["actual"]) - float(g["mae"])) < 1e-6
    # tests/test_algo.py:266-280 closed form; :330-339 == 1.0
    a = np.array([[1.0, 0.0], [3.0, 4.0]])
    p = a + np.array([[1.0, 5.0], [1.0, 1.0]])
    assert metrics.rmse(p, a) == 1.0 and metrics.mae(p, a) == 1.0
    assert metrics.rmse(p, np.zeros_like(a)) == 0.0
    with pytest.raises(ValueError, match="infinite"):
        metrics.rmse(np.array([[np.inf]]), np.array([[1.0]]))
    with pytest.raises(ValueError, match="NaN"):
        metrics.mae(np.array([[np.nan]]), np.array([[1.0]]))
    with pytest.raises(ValueError, match="must have the same shape"):
        metrics.rmse(np.ones((2, 2)), np.ones((2, 3)))


def test_generators_are_deterministic_and_sorted():
    a = gen.synth_pattern(2000, 500, mean_deg=20, seed=3)
    b = gen.synth_pattern(2000, 500, mean_deg=20, seed=3)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    indptr, indices = a
    assert indptr[0] == 0 and indptr[-1] == len(indices) and indices.dtype == np.int32
    assert (np.diff(indptr) >= 1).all()
    for r in (0, 17, 1999):
        seg = indices[indptr[r]:indptr[r + 1]]
        assert (np.diff(seg) > 0).all()
    s = gen.synth_pattern(2000, 500, mean_deg=20, seed=13, skew=True)
    deg_items = np.bincount(s[1], minlength=500)
    assert deg_items.max() > 8 * np.median(deg_items)      # Zipf head
    assert gen.c1_matrix().sum() == 49587                   # ~49.6 K nnz

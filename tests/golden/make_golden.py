#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (build container only).

    NUMBA_CACHE_DIR=/tmp/numba PYTHONPATH=/root/reference/src python tests/golden/make_golden.py

The reference (/root/reference, Lunexa-AI/recsys-lite) is pure Python and importable here but
does not exist on the GPU box, so its outputs on fixed seeded inputs are committed as small
fixtures.  Every fixture stores the inputs too, so the tests never need the reference.
Functions exercised: RecommenderSystem('als').fit (algo.py:300-337), .predict (:143-147),
.recommend (:172-208), top_n (:587-659), svd_reconstruct (:412-530), RecommenderSystem('svd')
fit/predict (:210-248, :352-375), compute_rmse / compute_mae (:662-743).
"""

import os
import sys

import numpy as np

os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache")
sys.path.insert(0, "/root/reference/src")
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))

import recsys_lite as ref  # noqa: E402
from oracle import generators as gen  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def als_case(name, ratings, k, iterations, lam, init_seed):
    np.random.seed(init_seed)
    model = ref.RecommenderSystem(algorithm="als").fit(ratings, k=k, iterations=iterations, lambda_=lam)
    u, v = model._model["user_factors"], model._model["item_factors"]
    pred = model.predict(ratings)
    rmse = ref.compute_rmse(pred, (ratings > 0).astype(np.float32))
    rec = model.recommend(ratings, n=10)
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"),
        ratings=np.packbits(ratings > 0, axis=1), shape=np.array(ratings.shape),
        k=k, iterations=iterations, lam=lam, init_seed=init_seed,
        user_factors=u, item_factors=v, rmse=rmse, recommend=rec,
    )
    print(name, ratings.shape, "rmse", rmse, "rec", rec.shape, rec.dtype)


def main():
    # ---- ALS -----------------------------------------------------------------------------
    rng = np.random.RandomState(7)
    small = (rng.random((60, 40)) < 0.15).astype(np.float32)
    small[5, :] = 0.0          # a user with no observation keeps its random row (algo.py:316-317)
    small[:, 9] = 0.0          # an item with no observation (algo.py:326-327)
    small[11, :] = 1.0         # a user who rated everything -> recommend row is all -1
    als_case("als_small_k8", small, k=8, iterations=3, lam=0.01, init_seed=11)
    als_case("als_small_k64", (rng.random((96, 80)) < 0.3).astype(np.float32), k=64, iterations=2,
             lam=0.01, init_seed=12)
    als_case("als_test_literal", np.array([[1, 0, 1], [0, 1, 0]], dtype=np.float32), k=2,
             iterations=10, lam=0.01, init_seed=13)          # tests/test_algo.py:700
    c1 = gen.CONFIGS["C1"]
    als_case("als_c1", gen.c1_matrix(), k=c1["k"], iterations=c1["iterations"], lam=c1["lam"],
             init_seed=c1["init_seed"])

    # ---- top_n -----------------------------------------------------------------------------
    cases = {}
    # tests/test_algo.py:141-142, :208-217, :193-196 style literals + seeded random cases
    est = np.array([[0.1, 0.4, 0.9, 0.7], [0.8, 0.2, 0.3, 0.6]], dtype=np.float32)
    known = np.array([[1, 0, 0, 0], [0, 0, 1, 0]], dtype=np.float32)
    cases["lit"] = (est, known, 3)
    est = np.array([[1.0, 2.0, 4.0, 3.0, 0.5]], dtype=np.float64)
    known = np.array([[0, 0, 0, 0, 1]], dtype=np.float64)
    cases["order"] = (est, known, 3)
    est = rng.random((7, 6)).astype(np.float32)
    known = (rng.random((7, 6)) < 0.5).astype(np.float32)
    known[2, :] = 1.0
    known[4, :5] = 1.0
    cases["ragged"] = (est, known, 4)
    est = rng.standard_normal((200, 300))
    known = (rng.random((200, 300)) < 0.2).astype(np.float32) * 3.0
    cases["rand_f64"] = (est, known, 10)
    est = rng.standard_normal((64, 1000)).astype(np.float32)
    known = (rng.random((64, 1000)) < 0.05).astype(np.float32)
    cases["rand_f32"] = (est, known, 10)
    blob = {}
    for name, (e, kn, n) in cases.items():
        out = ref.top_n(e, kn, n=n)
        blob[name + "_est"], blob[name + "_known"], blob[name + "_n"], blob[name + "_out"] = e, kn, n, out
        print("top_n", name, out.shape, out.dtype)
    np.savez_compressed(os.path.join(OUT, "topn_cases.npz"), **blob)

    # ---- svd_reconstruct -------------------------------------------------------------------
    blob = {}
    lit = np.array([[5, 3, 0, 1], [0, 0, 4, 5], [1, 1, 0, 0]], dtype=np.float32)  # tests/test_algo.py:31
    rr = np.random.RandomState(42)
    rank3 = (rr.standard_normal((10, 3)) @ rr.standard_normal((3, 8))).astype(np.float32)  # KAT :88-107 style
    rnd = rr.random((50, 20)).astype(np.float32)
    samp = gen.sample_ratings(300, 40, sparsity=0.9, random_state=0)
    for name, (m, k) in {"lit": (lit, 2), "rank3": (rank3, 3), "rnd": (rnd, 5), "samp": (samp, 10)}.items():
        out = ref.svd_reconstruct(m, k=k, use_sparse=False)
        blob[name + "_in"], blob[name + "_k"], blob[name + "_out"] = m, k, out
        print("svd", name, out.shape, out.dtype)
    np.savez_compressed(os.path.join(OUT, "svd_cases.npz"), **blob)

    # ---- SVD class path --------------------------------------------------------------------
    m = gen.sample_ratings(120, 30, sparsity=0.7, random_state=3)
    model = ref.RecommenderSystem(algorithm="svd", use_sparse=False).fit(m, k=6)
    pred = model.predict(m)
    rec = model.recommend(m, n=5)
    np.savez_compressed(
        os.path.join(OUT, "svd_class.npz"), ratings=m, k=6, pred=pred, recommend=rec,
        global_bias=model._model["global_bias"], user_bias=model._model["user_bias"],
        item_bias=model._model["item_bias"],
    )
    print("svd class", pred.shape, pred.dtype, rec.shape, rec.dtype)

    # ---- rmse / mae ------------------------------------------------------------------------
    p = rng.random((40, 30)).astype(np.float32) * 5
    a = gen.sample_ratings(40, 30, sparsity=0.6, random_state=5)
    np.savez_compressed(os.path.join(OUT, "metrics.npz"), pred=p, actual=a,
                        rmse=ref.compute_rmse(p, a), mae=ref.compute_mae(p, a))

    # ---- ranking metrics (tools.py:50-110) and the pipeline's recommend (tools.py:264-266) -------------------------
    from recsys_lite import tools as ref_tools
    rr = np.random.RandomState(11)
    recs = [list(rr.choice(60, size=int(rr.randint(0, 9)), replace=False)) for _ in range(200)]
    actual = [set(rr.choice(60, size=int(rr.randint(0, 12)), replace=False).tolist()) for _ in range(200)]
    recs[3], actual[3] = [5], {0, 5}                       # one-entry lists: the reference's ideal-list quirk
    recs[4], actual[4] = [5], {5, 40}
    width = max(len(x) for x in recs)
    rm = np.full((200, width), -1, dtype=np.int32)
    for i, x in enumerate(recs):
        rm[i, : len(x)] = x
    am = np.zeros((200, 60), dtype=np.int8)
    for i, a_set in enumerate(actual):
        am[i, list(a_set)] = 1
    vals = {k: (ref_tools.precision_at_k(recs, actual, k), ref_tools.recall_at_k(recs, actual, k), ref_tools.ndcg_at_k(recs, actual, k))
            for k in (1, 3, 5, 10)}
    np.savez_compressed(os.path.join(OUT, "ranking_metrics.npz"), recs=rm, actual=am,
                        ks=np.array(sorted(vals)), values=np.array([vals[k] for k in sorted(vals)]))
    print("ranking metrics", vals)
    m = gen.sample_ratings(90, 25, sparsity=0.7, random_state=8)
    model = ref.RecommenderSystem(algorithm="svd", use_sparse=False)
    pipe = ref_tools.RecsysPipeline([("model", model)]).fit(m, k=5)
    np.savez_compressed(os.path.join(OUT, "pipeline.npz"), ratings=m, k=5, recommend=np.array(pipe.recommend(m, n=4)),
                        predict=pipe.predict(m))

    # ---- KNN item-item cosine + neighbour-weighted prediction (algo.py:339-350, :148-169) ------------------------------
    rr = np.random.RandomState(21)
    m = (rr.random_sample((60, 35)) < 0.35) * rr.uniform(1.0, 5.0, (60, 35))
    m = m.astype(np.float32)
    model = ref.RecommenderSystem(algorithm="knn").fit(m, k=6)
    np.savez_compressed(os.path.join(OUT, "knn.npz"), ratings=m, neighbors=6, item_sim=model._model["item_sim"], pred=model.predict(m))
    print("knn", model._model["item_sim"].shape)


if __name__ == "__main__":
    main()

"""Factor error vs the float64 oracle as a function of the refinement steps (k = 64, ~50 nnz/user, lambda = 0.01)."""
import os, sys
import numpy as np
from scipy import sparse
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from oracle import als as o_als, compare
from recsys_lite_b200 import RecommenderSystem, synth

n_users, n_items, k, iters = 6000, 1500, 64, 3
indptr, indices = synth.synth_pattern(n_users, n_items, mean_deg=50, seed=3)
r = sparse.csr_matrix((np.ones(len(indices), np.float32), indices, indptr), shape=(n_users, n_items))
u0, v0 = synth.seeded_factors(n_users, n_items, k, 3)
ti, tx = o_als.transpose_pattern(indptr, indices, n_items)
ref = o_als.fit_csr(indptr, indices, ti, tx, k, iterations=iters, lam=0.01, init=(u0, v0))
for prec, irs in (("fp32_ir", (0, 1, 2, 3)), ("fp64", (0,))):
    for ir in irs:
        m = RecommenderSystem("als").fit(r, k=k, iterations=iters, lambda_=0.01, init=(u0, v0), precision=prec, ir_steps=ir)._model
        eu = compare.factor_error(m["user_factors"], ref["user_factors"])
        ev = compare.factor_error(m["item_factors"], ref["item_factors"])
        print(f"{prec} ir={ir}: U fro {eu[0]:.2e} max {eu[1]:.2e} | V fro {ev[0]:.2e} max {ev[1]:.2e}")

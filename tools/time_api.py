#!/usr/bin/env python
"""Time the drop-in Python surface at C4 size: RecommenderSystem("als").fit(R_csr) + .recommend(R_csr, n=10)."""
import os
import sys
import time

import numpy as np
from scipy import sparse

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from recsys_lite_b200 import RecommenderSystem, synth  # noqa: E402

nu, ni, k = 1_000_000, 100_000, 64
indptr, indices = synth.synth_pattern(nu, ni, 50, seed=4)
r = sparse.csr_matrix((np.ones(len(indices), np.float32), indices, indptr), shape=(nu, ni))
rng = np.random.default_rng(4)
u0, v0 = rng.random((nu, k)), rng.random((ni, k))
for rep in range(3):
    t0 = time.perf_counter()
    m = RecommenderSystem("als").fit(r, k=k, iterations=1, lambda_=0.01, init=(u0, v0))
    t1 = time.perf_counter()
    rec = m.recommend(r, n=10)
    t2 = time.perf_counter()
    print(f"rep {rep}: fit {t1 - t0:.3f} s, recommend {t2 - t1:.3f} s, recs {rec.shape} {rec.dtype}")

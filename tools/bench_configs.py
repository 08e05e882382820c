#!/usr/bin/env python
"""The other BASELINE.json configurations (bench.py measures C4): C1, C2, C3 and a k = 128 shape, GPU next to the CPU oracle.

    python tools/bench_configs.py > gpurun_out/configs.json

C1  implicit ALS k=16, 10 iterations, 1K x 1K dense binary (through RecommenderSystem, host arrays in / out)
C2  svd_reconstruct k=10 + top_n n=10 on 10K x 500 (README chunked-SVD shape), host arrays in / out
C3  implicit ALS k=64, 100K x 20K, ~5M nnz: one iteration on device-resident data; oracle on the first 20K users / items
K128 k=128 half-steps on a C3-sized pattern (general kernel: the tensor-core solve path is specialised for padded k = 64)
The oracle (CPU restatement of the reference, oracle/) is used here as the timed CPU baseline only.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import als as o_als, svd as o_svd, topn as o_topn  # noqa: E402
from recsys_lite_b200 import RecommenderSystem, _ffi, svd_reconstruct, synth, top_n  # noqa: E402
from recsys_lite_b200.csr import transpose_pattern  # noqa: E402


def timed(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return float(np.median(ts))


def main():
    from devbuf import Dev, padded
    h = _ffi.default_handle()
    out = {}
    # ---- C1
    r1 = synth.c1_matrix()
    u0, v0 = synth.seeded_factors(1000, 1000, 16, 123)
    gpu = timed(lambda: RecommenderSystem("als").fit(r1, k=16, iterations=10, lambda_=0.01, init=(u0, v0)))
    t0 = time.perf_counter()
    o_als.fit_dense(r1, 16, iterations=10, lam=0.01, init=(u0, v0))
    cpu = time.perf_counter() - t0
    out["C1"] = {"what": "ALS k=16, 10 iterations, 1000 x 1000 binary, host in/out", "gpu_s": gpu, "cpu_oracle_s": cpu, "speedup": cpu / gpu}
    # ---- C2
    m = synth.c2_readme_matrix().astype(np.float32)
    m2 = synth.sample_ratings(10_000, 500, sparsity=0.9, random_state=0)

    def c2_gpu():
        rec = svd_reconstruct(m2, k=10, use_sparse=False)
        return top_n(rec, m2, n=10)
    gpu = timed(c2_gpu)
    t0 = time.perf_counter()
    rec = o_svd.svd_reconstruct_dense(m2, k=10)
    o_topn.top_n(rec, m2, n=10)
    cpu = time.perf_counter() - t0
    gpu_svd = timed(lambda: svd_reconstruct(m, k=10, use_sparse=False))
    out["C2"] = {"what": "svd_reconstruct k=10 + top_n n=10, 10000 x 500, host in/out", "gpu_s": gpu, "cpu_oracle_s": cpu,
                 "speedup": cpu / gpu, "gpu_svd_reconstruct_readme_matrix_s": gpu_svd}
    # ---- C3 and k = 128
    for name, k in (("C3", 64), ("K128", 128)):
        w = synth.CONFIGS["C3"]
        nu, ni = w["n_users"], w["n_items"]
        indptr, indices = synth.synth_pattern(nu, ni, w["mean_deg"], w["seed"])
        ti, tx = transpose_pattern(indptr, indices, ni)
        rng = np.random.default_rng(3)
        uu = rng.random((nu, k), dtype=np.float32)
        vv = rng.random((ni, k), dtype=np.float32)
        d_u, d_v = padded(h, uu, k), padded(h, vv, k)
        d_ip, d_ix, d_tp, d_tx = Dev(h, indptr), Dev(h, indices), Dev(h, ti), Dev(h, tx)

        def it():
            h.call("rl_als_half_step_dev", nu, ni, k, d_ip.ptr, d_ix.ptr, None, None, d_v.ptr, d_u.ptr, 0.01, 0.0, None, 0, 1, None)
            h.call("rl_als_half_step_dev", ni, nu, k, d_tp.ptr, d_tx.ptr, None, None, d_u.ptr, d_v.ptr, 0.01, 0.0, None, 0, 1, None)
            h.synchronize()
        gpu = timed(it, reps=5, warm=2)
        su, si = 20_000, min(20_000, ni)
        uu64, vv64 = uu.astype(np.float64), vv.astype(np.float64)
        t0 = time.perf_counter()
        o_als.half_step_csr(indptr[:su + 1], indices[:indptr[su]], vv64, uu64[:su].copy(), 0.01)
        tu = (time.perf_counter() - t0) * nu / su
        t0 = time.perf_counter()
        o_als.half_step_csr(ti[:si + 1], tx[:ti[si]], uu64, vv64[:si].copy(), 0.01)
        tv = (time.perf_counter() - t0) * ni / si
        out[name] = {"what": f"ALS k={k}, one iteration, {nu} x {ni}, {len(indices)} nnz, device-resident",
                     "gpu_s": gpu, "cpu_oracle_s": tu + tv, "speedup": (tu + tv) / gpu,
                     "cpu_sample": f"first {su} users and {si} items, extrapolated linearly in rows"}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()

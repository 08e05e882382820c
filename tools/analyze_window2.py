#!/usr/bin/env python
"""Offline emulation of the class-wise candidate window (see tools/analyze_window.py): items are binned by the norm of
their centred vector w_i = v_i - vbar (factor-of-2 classes), every class has its own error bound
eps_c = |u - u_h| W_c + |u_h| WL_c + acc |u||W_c| ; a candidate is kept iff a_i + eps_c >= n-th largest (a_j - eps_c(j))."""
import sys
import numpy as np

f = np.load(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/conv_c4.npz")
tag = sys.argv[2] if len(sys.argv) > 2 else "last"
nsamp = int(sys.argv[3]) if len(sys.argv) > 3 else 512
ratio_cls = float(sys.argv[4]) if len(sys.argv) > 4 else 2.0
V = f[f"v_{tag}"].astype(np.float32)
U = f[f"u_{tag}"][:nsamp].astype(np.float32)
umean = f[f"umean_{tag}"].astype(np.float32)
indptr, indices = f["indptr_s"], f["indices_s"]
ni, k = V.shape
if ni < 100000:
    keep = indices < ni


def pow2_scale(mx):
    return 2.0 ** np.floor(14 - np.log2(np.maximum(mx, 1e-30)))


def run(name, centre_u, terms, acc):
    vbar = V.astype(np.float64).mean(axis=0).astype(np.float32)
    W = (V - vbar).astype(np.float32)                       # fp32 subtraction, as the kernel does it
    W64 = W.astype(np.float64)
    wn = np.linalg.norm(W64, axis=1)
    cls = np.minimum(np.floor(np.log(wn.max() / np.maximum(wn, 1e-300)) / np.log(ratio_cls)).astype(int), 7)
    sv = pow2_scale(np.abs(W64).max())
    wh = (W64 * sv).astype(np.float16).astype(np.float64)
    wl = ((W64 * sv) - wh).astype(np.float16).astype(np.float64)
    r1 = np.linalg.norm(W64 * sv - wh, axis=1)
    r2 = np.linalg.norm(W64 * sv - wh - wl, axis=1)
    ncls = cls.max() + 1
    Wc = np.array([wn[cls == c].max() * sv if (cls == c).any() else 0 for c in range(ncls)])
    R1c = np.array([r1[cls == c].max() if (cls == c).any() else 0 for c in range(ncls)])
    R2c = np.array([r2[cls == c].max() if (cls == c).any() else 0 for c in range(ncls)])
    print(name, "class sizes", np.bincount(cls), "class max norms", np.round(Wc / sv, 4))
    ub = umean if centre_u else np.zeros_like(umean)
    bias = W64 @ ub.astype(np.float64)                      # per-item exact term (added in the GEMM by extra columns)
    cnts = []
    for r in range(len(U)):
        u = (U[r] - ub).astype(np.float32).astype(np.float64)
        su = pow2_scale(np.abs(u).max())
        uh = (u * su).astype(np.float16).astype(np.float64)
        ul = (u * su - uh).astype(np.float16).astype(np.float64)
        un = np.linalg.norm(u * su)
        if terms == 1:
            a = wh @ uh
            eps_c = np.linalg.norm(u * su - uh) * Wc + np.linalg.norm(uh) * R1c
        else:
            a = wh @ uh + wh @ ul + wl @ uh
            eps_c = np.linalg.norm(u * su - uh - ul) * Wc + np.linalg.norm(uh) * R2c + np.linalg.norm(u * su - uh) * R1c
        eps_c = (eps_c + (acc + 2.0 ** -24) * un * Wc) / (su * sv)
        a = a / (su * sv) + bias
        eps = eps_c[cls]
        exact = W64 @ u + bias
        assert (np.abs(a - exact) <= eps * 1.000001 + 1e-300).all()
        seen = indices[indptr[r]:indptr[r + 1]]
        seen = seen[seen < ni]
        lo = a - eps
        lo[seen] = -np.inf
        hi = a + eps
        hi[seen] = -np.inf
        lth = np.partition(lo, -10)[-10]
        cnts.append(int((hi >= lth).sum()))
    c = np.array(cnts)
    print(f"{name:40s} candidates: median {np.median(c):5.0f} p90 {np.percentile(c, 90):6.0f} p99 {np.percentile(c, 99):6.0f} "
          f"p99.9 {np.percentile(c, 99.9):6.0f} max {c.max():6d}  >24: {100.0 * (c > 24).mean():5.2f}%  >32: {100.0 * (c > 32).mean():5.2f}%  >64: {100.0 * (c > 64).mean():5.2f}%")


K = V.shape[1]
run("1-term, V centred, classes", False, 1, K * 2.0 ** -23)
run("1-term, U+V centred, classes", True, 1, K * 2.0 ** -23)
run("3-term, V centred, classes", False, 3, (3 * K / 16 + 4) * 2.0 ** -23)

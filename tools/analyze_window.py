#!/usr/bin/env python
"""Offline (CPU) analysis of the candidate window of the tensor-core scoring pass on dumped factors
(tools/dump_converged.py): how many unseen items fall inside [n-th best approximate score - 2 eps, inf) per user for
different operand splits / centrings.  Pure NumPy emulation of the fp16 roundings; decides the kernel design."""
import sys
import numpy as np

f = np.load(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/conv_c4.npz")
tag = sys.argv[2] if len(sys.argv) > 2 else "last"
nsamp = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
V = f[f"v_{tag}"].astype(np.float64)
U = f[f"u_{tag}"][:nsamp].astype(np.float64)
umean = f[f"umean_{tag}"]
indptr, indices = f["indptr_s"], f["indices_s"]
ni, k = V.shape
print("items", ni, "k", k, "users", len(U))
vbar = V.mean(axis=0)
print("|vbar| %.4f  mean|v| %.4f max|v| %.4f  mean|v-vbar| %.4f max|v-vbar| %.4f" % (
    np.linalg.norm(vbar), np.linalg.norm(V, axis=1).mean(), np.linalg.norm(V, axis=1).max(),
    np.linalg.norm(V - vbar, axis=1).mean(), np.linalg.norm(V - vbar, axis=1).max()))
print("|ubar| %.4f  mean|u| %.4f max|u| %.4f  mean|u-ubar| %.4f max|u-ubar| %.4f" % (
    np.linalg.norm(umean), np.linalg.norm(U, axis=1).mean(), np.linalg.norm(U, axis=1).max(),
    np.linalg.norm(U - umean, axis=1).mean(), np.linalg.norm(U - umean, axis=1).max()))


def pow2_scale(mx):
    # largest power of two s with mx * s <= 2^14 (as common.cuh pow2_scale)
    return 2.0 ** np.floor(14 - np.log2(np.maximum(mx, 1e-30)))


def split16(x):
    hi = x.astype(np.float16).astype(np.float64)
    lo = (x - hi).astype(np.float16).astype(np.float64)
    return hi, lo


def analyze(name, Uq, Vq, bias, terms, err_extra=0.0):
    """Uq, Vq: operands given to the tensor core (fp64 copies of fp32 data); bias: per-item exact term added."""
    sv = pow2_scale(np.abs(Vq).max())
    vh, vl = split16(Vq * sv)
    vnorm_max = np.linalg.norm(Vq * sv, axis=1).max()
    vl_max = np.linalg.norm(Vq * sv - vh, axis=1).max()
    vl2_max = np.linalg.norm(Vq * sv - vh - vl, axis=1).max()
    cnts, ratio = [], []
    for r in range(len(Uq)):
        u = Uq[r]
        su = pow2_scale(np.abs(u).max())
        uh, ul = split16(u * su)
        if terms == 1:
            approx = vh @ uh
            eps = np.linalg.norm(u * su - uh) * vnorm_max + np.linalg.norm(uh) * vl_max
        else:
            approx = vh @ uh + vh @ ul + vl @ uh
            eps = np.linalg.norm(u * su - uh - ul) * vnorm_max + np.linalg.norm(uh) * vl2_max + np.linalg.norm(u * su - uh) * vl_max
        eps += err_extra * np.linalg.norm(u * su) * vnorm_max
        approx = approx / (su * sv)
        eps = eps / (su * sv)
        if bias is not None:
            approx = approx + bias
        exact = Uq[r] @ Vq.T + (bias if bias is not None else 0.0)
        assert np.abs(approx - exact).max() <= eps * 1.0000001 + 1e-300, (np.abs(approx - exact).max(), eps)
        seen = indices[indptr[r]:indptr[r + 1]]
        approx[seen] = -np.inf
        t10 = np.partition(approx, -10)[-10]
        cnts.append(int((approx > t10 - 2 * eps).sum()))
        ratio.append(np.abs(approx[approx > -np.inf] - exact[approx > -np.inf]).max() / eps)
    c = np.array(cnts)
    print(f"{name:34s} candidates: median {np.median(c):6.0f}  p90 {np.percentile(c, 90):7.0f}  p99 {np.percentile(c, 99):7.0f}  max {c.max():7d}"
          f"  >32: {100.0 * (c > 32).mean():5.1f}%  >64: {100.0 * (c > 64).mean():5.1f}%   (true err / bound: {np.max(ratio):.3f})")


ACC = 2.0 ** -22   # allowance for fp32 accumulation in the tensor core, relative to |u||v|
Vc = V - vbar
Uc = U - umean
bias_i = Vc @ umean
analyze("1-term plain", U, V, None, 1, ACC)
analyze("1-term, V centred", U, Vc, None, 1, ACC)
analyze("1-term, U and V centred + bias_i", Uc, Vc, bias_i, 1, ACC)
analyze("3-term plain (rigorous eps)", U, V, None, 3, ACC)
analyze("3-term, V centred", U, Vc, None, 3, ACC)
analyze("3-term, U,V centred + bias_i", Uc, Vc, bias_i, 3, ACC)
# the r1 window: 6e-6 |u| max|v|
cn = []
vmax = np.linalg.norm(V, axis=1).max()
for r in range(len(U)):
    s = V @ U[r]
    s[indices[indptr[r]:indptr[r + 1]]] = -np.inf
    t10 = np.partition(s, -10)[-10]
    cn.append(int((s > t10 - 2 * 6e-6 * np.linalg.norm(U[r]) * vmax).sum()))
cn = np.array(cn)
print(f"r1 window 6e-6|u|max|v|: median {np.median(cn):.0f} p90 {np.percentile(cn, 90):.0f} p99 {np.percentile(cn, 99):.0f} max {cn.max()} >32: {100.0 * (cn > 32).mean():.1f}%")
# score statistics
r = 0
s = V @ U[r]
print("user 0: scores mean %.6f std %.3e top10 %s" % (s.mean(), s.std(), np.sort(s)[-12:][::-1]))

"""CPU-side checks of the boundary: the shared object exports every symbol include/recsys_b200.h declares, the
ctypes table matches the header, validation happens before any device call, and the product never falls back
to the CPU (or to oracle/)."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _header_functions():
    text = open(os.path.join(ROOT, "include", "recsys_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rl_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    from recsys_lite_b200 import build, _ffi
    build.build()
    return _ffi.load_library()


def test_library_exports_every_declared_symbol(lib):
    from recsys_lite_b200 import _ffi
    names = _header_functions()
    assert len(names) >= 25
    assert set(names) == set(_ffi.SIGNATURES), set(names) ^ set(_ffi.SIGNATURES)
    for n in names:
        assert hasattr(lib, n), n
    assert lib.rl_abi_version() == _ffi.ABI_VERSION
    out = subprocess.run(["nm", "-D", "--defined-only", _ffi.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (\w+)", out))
    assert set(names) <= exported
    # nothing but the C ABI is exported (built with -fvisibility=hidden)
    assert all(e.startswith("rl_") or e in ("_init", "_fini") for e in exported), exported - set(names)


def test_padded_k(lib):
    assert [lib.rl_padded_k(k) for k in (1, 16, 17, 32, 33, 64, 65, 128, 129, 0, -3)] == \
        [16, 16, 32, 32, 64, 64, 128, 128, 0, 0, 0]


def test_sass_is_sm100a_only():
    from recsys_lite_b200 import _ffi
    out = subprocess.run(["cuobjdump", "-lelf", _ffi.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def _has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_cuda(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback_without_gpu(lib):
    """Without a device the product raises; it must not compute anything on the host."""
    import recsys_lite_b200 as r
    from recsys_lite_b200._ffi import RecsysB200Error
    x = np.random.default_rng(0).random((6, 5)).astype(np.float32)
    with pytest.raises(RecsysB200Error, match="no CUDA device|no CPU path"):
        r.top_n(x, np.zeros_like(x), n=2)
    with pytest.raises(RecsysB200Error):
        r.svd_reconstruct(x, k=2)
    with pytest.raises(RecsysB200Error):
        r.RecommenderSystem("als").fit(x, k=2)


def test_validation_precedes_device_calls():
    """error strings of the reference's tests (tests/test_algo.py:59-83, :167-181, :298-380, :423-478) are raised
    in Python before any C call, so they hold with or without a GPU."""
    import recsys_lite_b200 as r
    x = np.ones((3, 4), np.float32)
    with pytest.raises(ValueError, match="Unsupported algorithm: 'invalid'"):
        r.RecommenderSystem(algorithm="invalid")
    with pytest.raises(RuntimeError, match="Model must be fitted"):
        r.RecommenderSystem("svd").predict(x)
    with pytest.raises(ValueError, match="n must be positive"):
        r.RecommenderSystem("als").recommend(x, n=0)
    with pytest.raises(ValueError, match="n must be positive"):
        r.top_n(x, x, n=-1)
    with pytest.raises(ValueError, match="must have the same shape"):
        r.top_n(x, np.ones((3, 5)), n=1)
    with pytest.raises(ValueError, match=r"Input matrix must be 2D \(users x items\), got 1D\."):
        r.svd_reconstruct(np.ones(3), k=1)
    with pytest.raises(ValueError, match="cannot have zero dimensions"):
        r.svd_reconstruct(np.ones((0, 3)), k=1)
    with pytest.raises(ValueError, match="k must be between"):
        r.svd_reconstruct(x, k=9)
    assert r.top_n(np.empty((0, 0)), np.empty((0, 0)), n=3).shape == (0, 0)
    assert not r.svd_reconstruct(np.zeros((3, 4)), k=2).any()           # all-zero guard needs no device
    with pytest.raises(ValueError, match="must have the same shape"):
        r.compute_rmse(x, np.ones((2, 2)))
    m = r.RecommenderSystem("als")
    assert m.get_params() == {"algorithm": "als", "use_sparse": True} and not m.clone().is_fitted()


def test_product_never_imports_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may touch oracle/ (or /root/reference)."""
    pkg = os.path.join(ROOT, "recsys_lite_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "sys.path" not in text or "reference" not in text, f
    code = "import sys; import recsys_lite_b200; assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules)"
    subprocess.run([sys.executable, "-c", code], check=True, cwd=ROOT)


def test_csr_helpers():
    from recsys_lite_b200 import csr
    from oracle import als as o_als, generators as gen
    indptr, indices = gen.synth_pattern(500, 120, mean_deg=9, seed=1)
    dense = gen.pattern_to_dense(indptr, indices, 120)
    dense[3, 5] = -2.0                                   # negatives are unobserved (algo.py:311)
    a = csr.observed_pattern(dense)
    from scipy import sparse
    b = csr.observed_pattern(sparse.coo_matrix(dense))
    want = o_als.pattern_to_csr(dense)
    for got in (a, b):
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
        assert got[0].dtype == np.int64 and got[1].dtype == np.int32
    t = csr.transpose_pattern(*a, 120)
    wt = o_als.transpose_pattern(*want, 120)
    assert np.array_equal(t[0], wt[0]) and np.array_equal(t[1], wt[1])
    bounds = csr.nnz_balanced_bounds(a[0], 4)
    assert bounds[0] == 0 and bounds[-1] == 500 and (np.diff(bounds) >= 0).all()
    per = np.diff(a[0][bounds])
    assert per.max() - per.min() <= 2 * np.diff(a[0]).max()
    lo, hi = int(bounds[1]), int(bounds[2])
    sp, sx = csr.shard_rows(a[0], a[1], lo, hi)
    assert sp[0] == 0 and sp[-1] == len(sx) and np.array_equal(sx, a[1][a[0][lo]:a[0][hi]])


def test_csr_check_host_and_fast_pattern_path(lib):
    """rl_csr_check_host (multi-threaded host validation, no device): bit 0 = canonical rows (strictly ascending, in range),
    bit 1 = all stored values > 0; observed_pattern takes scipy's arrays as they are only when both hold and returns what
    the scipy route returns otherwise (algo.py:311 semantics)."""
    import ctypes as C
    from scipy import sparse
    from recsys_lite_b200 import csr, _ffi
    from oracle import generators as gen

    def flags(m, threads=0):
        f = C.c_int(-1)
        rc = lib.rl_csr_check_host(_ffi.ptr(m.indptr), csr._DT[m.indptr.dtype], _ffi.ptr(m.indices), _ffi.ptr(m.data),
                                   csr._DT[m.data.dtype], m.shape[0], m.shape[1], threads, C.byref(f))
        assert rc == 0
        return f.value

    indptr, indices = gen.synth_pattern(40000, 3000, mean_deg=40, seed=3)      # > 2^20 entries: the threaded path
    for dt in (np.float32, np.float64, np.int32, np.int64):
        m = sparse.csr_matrix((np.ones(len(indices), dt), indices, indptr), shape=(40000, 3000))
        assert flags(m) == 3 and flags(m, 1) == 3 and flags(m, 5) == 3
    m = sparse.csr_matrix((np.ones(len(indices), np.float32), indices.copy(), indptr), shape=(40000, 3000))
    fast = csr.observed_pattern(m)
    assert fast[1] is m.indices or np.shares_memory(fast[1], m.indices)     # no copy of the 4 nnz bytes
    assert np.array_equal(fast[0], indptr) and fast[0].dtype == np.int64
    row = 31234
    s, e = indptr[row], indptr[row + 1]
    m.data[s + 1] = -1.0                                                     # a negative value: bit 1 drops
    assert flags(m) == 1
    got = csr.observed_pattern(m)
    keep = np.ones(len(indices), bool); keep[s + 1] = False
    assert np.array_equal(got[1], indices[keep]) and got[0][-1] == len(indices) - 1
    m.data[s + 1] = 1.0
    m.indices[s], m.indices[s + 1] = m.indices[s + 1], m.indices[s]         # unsorted row: bit 0 drops
    assert flags(m) == 2
    m.has_sorted_indices = False
    got = csr.observed_pattern(m)
    assert np.array_equal(got[0], indptr) and np.array_equal(got[1], indices)
    m2 = sparse.csr_matrix((np.ones(len(indices), np.float32), indices, indptr), shape=(40000, 3000))
    m2.indices = m2.indices.copy()
    m2.indices[e - 1] = 3000                                                # column out of range
    assert flags(m2) == 2
    m2.indices[e - 1] = m2.indices[e - 2]                                   # duplicate column
    assert flags(m2) == 2
    empty = sparse.csr_matrix((5, 7), dtype=np.float32)
    assert flags(empty) == 3

"""Drop-in surface of the recsys-lite hot path on a B200: same names, arguments and error behaviour as
/root/reference/src/recsys_lite/algo.py, with the arithmetic done by librecsys_b200.so (sm_100a kernels).

Mirrors (reference file:line):
  RecommenderSystem.__init__/is_fitted   algo.py:50-79
  RecommenderSystem.fit                  algo.py:81-118   (als -> _fit_als :300-337, svd -> _fit_svd_dense :210-248)
  RecommenderSystem.predict              algo.py:120-170  (als :143-147, svd -> _predict_svd :352-375)
  RecommenderSystem.recommend            algo.py:172-208
  save / load / get_params / clone / reset   algo.py:377-409
  svd_reconstruct                        algo.py:412-530
  top_n                                  algo.py:587-659
  compute_rmse / compute_mae             algo.py:662-743

Validation, error strings and output post-processing stay in Python; there is no
NumPy fallback for the device work: without the shared library or a B200 the calls raise RecsysB200Error.

Documented deviations from the reference:
  * ALS accepts scipy.sparse input (the reference raises IndexError for it) and the extra keyword arguments
    ``precision`` ("fp32_ir" default | "fp64"), ``ir_steps`` and ``init=(U0, V0)``;
  * factors are kept in float32 on the device and returned as float64 (reference: float64 throughout):
    relative factor error ~1e-6 against the reference, inside the stated 1e-4 tolerance;
  * equal scores are ordered by ascending item index (the reference's tie order is implementation defined);
  * ``algorithm="knn"`` (outside the north-star path): item similarities and predictions are computed as
    masked Gram products on the tensor cores; neighbours with exactly equal similarity at the cut are chosen by ascending
    item index (the reference's argsort order is implementation defined);
  * limits of the fused kernels: ALS fit 1..128 factors (NotImplementedError beyond); ``recommend`` with more than 128
    factors (SVD models default to min(shape) // 4) or n > 128 takes the unfused device path (predict + top_n per row
    chunk); ``svd_reconstruct`` / the SVD fit need min(n_users, n_items) <= 4096 (ValueError beyond: the device
    eigen-solver is a one-CTA-per-column-pair Jacobi).
"""
from __future__ import annotations

import os
import sys
import time
from typing import Any, Dict, Optional

import numpy as np
from scipy import sparse

from . import _ffi
from .csr import observed_pattern

MAX_FACTORS = 128   # rl_padded_k
MAX_TOP_N = 128     # on-chip list capacity of the selection kernels

_PRECISIONS = {"fp32_ir": _ffi.RL_PREC_FP32_IR, "fp64": _ffi.RL_PREC_FP64}


def _dense(x):
    return x.toarray() if sparse.issparse(x) else x


def _as_float_matrix(x, allow=(np.float32, np.float64), default=np.float64):
    a = np.asarray(_dense(x))
    if a.dtype.type not in allow:
        a = a.astype(default)
    return np.ascontiguousarray(a)


def _rl_dtype(a):
    return _ffi.RL_F64 if a.dtype == np.float64 else _ffi.RL_F32


_TIMING = bool(os.environ.get("RL_API_TIMING"))


class _Tick:
    """RL_API_TIMING=1: host-side phase times of fit / recommend on stderr (where the adapter's time goes)."""

    def __init__(self, what):
        self.what, self.t, self.parts = what, time.perf_counter(), []

    def __call__(self, name):
        if _TIMING:
            t = time.perf_counter()
            self.parts.append(f"{name} {1e3 * (t - self.t):.1f}")
            self.t = t

    def done(self):
        if _TIMING:
            sys.stderr.write(f"[rl api] {self.what}: " + ", ".join(self.parts) + " ms\n")


def _finish_topn(idx, n_valid):
    """-1 padded (n_users, n) int32 + valid counts -> the reference's trimmed result (algo.py:629-639)."""
    widest = int(n_valid.max()) if n_valid.size else 0
    if widest == 0:
        return np.empty((idx.shape[0], 0), dtype=int)
    return idx[:, :widest]


class _DevBuf:
    """One device allocation through the C ABI, from the stream-ordered pool (rl_dev_alloc_async / rl_dev_free_async:
    no cudaMalloc / cudaFree and no device synchronisation per call); `adopt` takes over a pointer the library handed out."""

    def __init__(self, h, nbytes: int = 0, adopt=None):
        import ctypes as C
        self.h, self.nbytes = h, int(nbytes)
        if adopt is not None:
            self.ptr = C.c_void_p(adopt)
            return
        p = C.c_void_p()
        h.call("rl_dev_alloc_async", max(self.nbytes, 16), C.byref(p))
        self.ptr = p

    def free(self):
        if getattr(self, "ptr", None):
            try:
                self.h.call("rl_dev_free_async", self.ptr)
            except Exception:
                pass
            self.ptr = None

    def __del__(self):
        self.free()


class _ResidentALS:
    """ALS factors (float32, padded rows) and the training pattern kept in HBM between fit / recommend / predict, so that
    the drop-in surface moves the model over PCIe only when the caller asks for the NumPy arrays."""

    def __init__(self, h, n_users, n_items, k, adopt=None):
        self.h, self.n_users, self.n_items, self.k = h, int(n_users), int(n_items), int(k)
        self.kp = h.lib.rl_padded_k(self.k)
        self.ip = self.ix = self.tp = self.tx = None
        self.pattern_key = None
        self.min_row_nnz = None
        if adopt is not None:       # the six buffers of rl_als_fit_resident_host
            self.u, self.v, self.ip, self.ix, self.tp, self.tx = (_DevBuf(h, adopt=p) if p else None for p in adopt)
            return
        self.u = _DevBuf(h, 4 * self.n_users * self.kp)
        self.v = _DevBuf(h, 4 * self.n_items * self.kp)

    @staticmethod
    def key_of(ratings, indptr, indices):
        return (id(ratings), tuple(ratings.shape), int(indptr[-1]), int(indices[:64].sum()) if len(indices) else 0)

    def upload_pattern(self, indptr, indices, key, transpose=True):
        h = self.h
        nnz = int(indptr[-1])
        self.ip, self.ix = _DevBuf(h, indptr.nbytes), _DevBuf(h, max(indices.nbytes, 16))
        h.call("rl_memcpy_h2d", self.ip.ptr, _ffi.ptr(indptr), indptr.nbytes)
        if nnz:
            h.call("rl_memcpy_h2d", self.ix.ptr, _ffi.ptr(indices), indices.nbytes)
        if transpose:
            self.tp, self.tx = _DevBuf(h, 8 * (self.n_items + 1)), _DevBuf(h, max(4 * nnz, 16))
            h.call("rl_csr_transpose_dev", self.n_users, self.n_items, nnz, self.ip.ptr, self.ix.ptr, self.tp.ptr, self.tx.ptr)
        self.pattern_key = key

    def upload_factors(self, user_f, item_f):
        self.h.call("rl_factors_upload", _ffi.ptr(user_f), _rl_dtype(user_f), self.n_users, self.k, self.u.ptr)
        self.h.call("rl_factors_upload", _ffi.ptr(item_f), _rl_dtype(item_f), self.n_items, self.k, self.v.ptr)

    def iterate(self, iterations, lam, precision, ir_steps):
        h = self.h
        for _ in range(int(iterations)):
            h.call("rl_als_half_step_dev", self.n_users, self.n_items, self.k, self.ip.ptr, self.ix.ptr, None, None,
                   self.v.ptr, self.u.ptr, float(lam), 0.0, None, precision, int(ir_steps), None)
            h.call("rl_als_half_step_dev", self.n_items, self.n_users, self.k, self.tp.ptr, self.tx.ptr, None, None,
                   self.u.ptr, self.v.ptr, float(lam), 0.0, None, precision, int(ir_steps), None)

    def download(self):
        uf = np.empty((self.n_users, self.k), dtype=np.float64)
        vf = np.empty((self.n_items, self.k), dtype=np.float64)
        self.h.call("rl_factors_download", self.u.ptr, self.n_users, self.k, _ffi.ptr(uf), _ffi.RL_F64)
        self.h.call("rl_factors_download", self.v.ptr, self.n_items, self.k, _ffi.ptr(vf), _ffi.RL_F64)
        return uf, vf

    def free(self):
        for b in (self.u, self.v, self.ip, self.ix, self.tp, self.tx):
            if b is not None:
                b.free()


class _LazyALSModel(dict):
    """The reference's ALS model dict ({"user_factors", "item_factors", "factors"}, algo.py:333-337) whose two arrays stay
    in HBM until somebody reads them.  Reading any of them downloads both as float64 NumPy arrays and makes the HOST copy
    authoritative from then on (the owner drops its device copy: callers may edit the arrays in place)."""

    _LAZY = ("user_factors", "item_factors")

    def __init__(self, owner, factors):
        super().__init__(user_factors=None, item_factors=None, factors=factors)
        self._owner = owner
        self._pending = True

    def _materialise(self):
        if self._pending:
            self._pending = False
            uf, vf = self._owner._resident.download()
            super().__setitem__("user_factors", uf)
            super().__setitem__("item_factors", vf)
            self._owner._drop_resident()

    def __getitem__(self, key):
        if key in self._LAZY:
            self._materialise()
        return super().__getitem__(key)

    def get(self, key, default=None):
        if key in self._LAZY:
            self._materialise()
        return super().get(key, default)

    def items(self):
        self._materialise()
        return super().items()

    def values(self):
        self._materialise()
        return super().values()

    def copy(self):
        self._materialise()
        return dict(super().items())

    def __reduce__(self):           # pickled (joblib save) as a plain dict of NumPy arrays
        self._materialise()
        return (dict, (dict(super().items()),))


class RecommenderSystem:
    """B200 implementation of the reference's ``RecommenderSystem`` for ``algorithm="als" | "svd"``."""

    def __init__(self, algorithm: str = "svd", random_state: Optional[int] = None, use_sparse: bool = True) -> None:
        self.algorithm = algorithm
        self.random_state = random_state
        self.use_sparse = use_sparse
        self._model: Optional[Dict[str, Any]] = None
        self._fitted = False
        self._resident: Optional[_ResidentALS] = None
        if algorithm not in ("svd", "als", "knn"):
            raise ValueError(f"Unsupported algorithm: '{algorithm}'")

    def is_fitted(self) -> bool:
        return self._fitted

    # ------------------------------------------------------------------------------------------ fit
    def fit(self, ratings, k: Optional[int] = None, **kwargs) -> "RecommenderSystem":
        if self.algorithm == "als":
            factors = kwargs.get("factors", k or 10)
            self._model = self._fit_als(
                ratings, factors, kwargs.get("iterations", 10), kwargs.get("lambda_", 0.01),
                precision=kwargs.get("precision", "fp32_ir"), ir_steps=kwargs.get("ir_steps", 1),
                init=kwargs.get("init"),
            )
        elif self.algorithm == "svd":
            self._model = self._fit_svd(ratings, k)
        elif self.algorithm == "knn":
            self._model = self._fit_knn(ratings, kwargs.get("neighbors", k or 10))
        else:
            raise ValueError(f"Unsupported algorithm: {self.algorithm}")
        self._fitted = True
        return self

    def _fit_als(self, ratings, factors: int, iterations: int = 10, lambda_: float = 0.01, *,
                 precision: str = "fp32_ir", ir_steps: int = 1, init=None) -> Dict[str, Any]:
        """algo.py:300-337: init from the GLOBAL NumPy RNG (U then V), binarise, `iterations` sweeps."""
        if ratings.ndim != 2:
            raise ValueError(f"Input matrix must be 2D (users x items), got {ratings.ndim}D.")
        n_users, n_items = ratings.shape
        factors = int(factors)
        if not 1 <= factors <= MAX_FACTORS:
            raise NotImplementedError(f"factors={factors}: the device kernels cover 1..{MAX_FACTORS}")
        if precision not in _PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_PRECISIONS)}, got {precision!r}")
        if init is None:
            user_f = np.random.rand(n_users, factors)   # draw order of algo.py:309-310
            item_f = np.random.rand(n_items, factors)
        else:
            # read only (uploaded as they are: float32 or float64); copied to float64 only if they become the model
            user_f = np.ascontiguousarray(init[0], dtype=None if np.asarray(init[0]).dtype == np.float32 else np.float64)
            item_f = np.ascontiguousarray(init[1], dtype=None if np.asarray(init[1]).dtype == np.float32 else np.float64)
            if user_f.shape != (n_users, factors) or item_f.shape != (n_items, factors):
                raise ValueError("init factors must have shapes (n_users, factors) and (n_items, factors)")
        if item_f.dtype != user_f.dtype:           # one dtype for both in the C call
            item_f = np.ascontiguousarray(item_f, dtype=user_f.dtype)
        tick = _Tick("fit_als")
        indptr, indices = observed_pattern(ratings)
        tick("pattern")
        self._drop_resident()
        tick("drop")
        if n_users and n_items and iterations > 0:
            h = _ffi.default_handle()
            # The model stays in HBM (float32, padded rows): pattern and initial factors go up once, the transposed pattern
            # (users of every item) is built on the device, and nothing comes back until the caller reads the arrays.
            # One call (rl_als_fit_resident_host): uploads on the copy stream underneath the first half-steps, the item-side
            # pattern built on the device; it hands back the six device buffers of the model and its pattern.
            import ctypes as C
            dev = (C.c_void_p * 6)()
            h.call("rl_als_fit_resident_host", n_users, n_items, factors, _ffi.ptr(indptr), _ffi.ptr(indices), None, None,
                   int(iterations), float(lambda_), _PRECISIONS[precision], int(ir_steps), _ffi.ptr(user_f), _ffi.ptr(item_f),
                   _rl_dtype(user_f), dev)
            res = _ResidentALS(h, n_users, n_items, factors, adopt=[dev[i] for i in range(6)])
            res.pattern_key = _ResidentALS.key_of(ratings, indptr, indices)
            try:
                h.synchronize()
            except Exception:
                res.free()
                raise
            tick("fit")
            tick.done()
            self._resident = res
            return _LazyALSModel(self, factors)
        return {"user_factors": np.array(user_f, dtype=np.float64), "item_factors": np.array(item_f, dtype=np.float64),
                "factors": factors}

    def _drop_resident(self):
        if getattr(self, "_resident", None) is not None:
            self._resident.free()
            self._resident = None

    def __getstate__(self):          # joblib / pickle: host arrays only (the device copy cannot travel)
        state = dict(self.__dict__)
        if isinstance(state.get("_model"), _LazyALSModel):
            state["_model"] = state["_model"].copy()
        state["_resident"] = None
        return state

    def _fit_knn(self, ratings, neighbors: int) -> Dict[str, Any]:
        """algo.py:339-350: item-item cosine over the users who rated both items -- on the device three masked Gram products
        on the tensor cores (csrc/gram_tc.cu) instead of the reference's O(n_items^2) Python loop."""
        r = np.ascontiguousarray(np.asarray(_dense(ratings)), dtype=np.float32)
        if r.ndim != 2:
            raise ValueError(f"Input matrix must be 2D (users x items), got {r.ndim}D.")
        n_users, n_items = r.shape
        sim = np.zeros((n_items, n_items), dtype=np.float64)
        if n_users and n_items:
            h = _ffi.default_handle()
            h.call("rl_knn_item_sim_host", _ffi.ptr(r), n_users, n_items, _ffi.ptr(sim))
        return {"item_sim": sim, "neighbors": neighbors, "ratings": ratings}

    def _predict_knn(self, ratings) -> np.ndarray:
        """algo.py:148-169: neighbour-weighted mean of the user's ratings of the `neighbors` most similar items, for the
        cells the user has not rated (rated cells stay 0)."""
        m = self._need_model()
        r = np.ascontiguousarray(np.asarray(_dense(ratings)), dtype=np.float32)
        sim = np.ascontiguousarray(m["item_sim"], dtype=np.float64)
        if r.ndim != 2 or r.shape[1] != sim.shape[0]:
            raise ValueError(f"ratings must have {sim.shape[0]} item columns, got shape {r.shape}")
        out = np.zeros(r.shape, dtype=np.float64)
        if r.size:
            h = _ffi.default_handle()
            h.call("rl_knn_predict_host", _ffi.ptr(r), r.shape[0], r.shape[1], _ffi.ptr(sim), int(m["neighbors"]), _ffi.ptr(out))
        return out

    def _fit_svd(self, ratings, k: Optional[int]) -> Dict[str, Any]:
        """algo.py:210-248: biases over entries > 0, centre every cell, rank-k SVD of the centred matrix.

        Bias means (masked row / column means), centring and the factorisation all run on the device
        (rl_svd_fit_host: csrc/bias.cu + csrc/svd.cu).  Model keys as pinned by tests/test_algo.py:659-684.
        """
        r = np.asarray(_dense(ratings))
        n_users, n_items = r.shape
        if k is None:
            k = max(1, min(n_users, n_items) // 4)
        k_eff = max(1, min(int(k), n_users, n_items))  # slicing u[:, :k] tolerates k > rank (algo.py:240-243)
        # bias means, centring and the factorisation in ONE device call: the ratings cross PCIe once (algo.py:218-243)
        ftype = np.float32 if r.dtype == np.float32 else np.float64
        r32 = np.ascontiguousarray(r, dtype=np.float32)
        bias, status = np.empty(1, dtype=np.float64), np.zeros(1, dtype=np.int32)
        ub64, ib64 = np.empty(n_users, dtype=np.float64), np.empty(n_items, dtype=np.float64)
        s = np.empty(k_eff, dtype=np.float64)
        left = np.empty((n_users, k_eff), dtype=np.float64)
        right = np.empty((n_items, k_eff), dtype=np.float64)
        _check_svd_shape(r32.shape)
        h = _ffi.default_handle()
        h.call("rl_svd_fit_host", _ffi.ptr(r32), n_users, n_items, k_eff, 1 if ftype is np.float32 else 0, _ffi.ptr(bias),
               _ffi.ptr(ub64), _ffi.ptr(ib64), _ffi.ptr(s), _ffi.ptr(left), _ffi.ptr(right), _ffi.ptr(status))
        gb, ub, ib = ftype(bias[0]), ub64.astype(ftype), ib64.astype(ftype)
        safe = np.where(s > 0, s, 1.0)
        if n_users >= n_items:   # left = U S, right = V
            u, vt = left / safe, right.T
        else:                    # left = U, right = V S
            u, vt = left, (right / safe).T
        return {
            "u": u.astype(np.float32), "s": s.astype(np.float32), "vt": np.ascontiguousarray(vt).astype(np.float32),
            "k": k, "global_bias": gb, "user_bias": ub, "item_bias": ib,
        }

    # -------------------------------------------------------------------------------------- predict
    def _need_model(self) -> Dict[str, Any]:
        if not self._fitted:
            raise RuntimeError("Model must be fitted before making predictions. Call .fit() first.")
        if self._model is None:
            raise RuntimeError("Model not fitted. Call .fit() first.")
        return self._model

    def _factor_view(self):
        """(row factors, column factors, global bias, row bias, col bias, output dtype) of the fitted model."""
        m = self._need_model()
        if self.algorithm == "als":
            return m["user_factors"], m["item_factors"], 0.0, None, None, np.float64
        us = np.ascontiguousarray(m["u"].astype(np.float64) * m["s"].astype(np.float64)[None, :])
        v = np.ascontiguousarray(m["vt"].T.astype(np.float64))
        out_t = np.float32 if np.asarray(m["user_bias"]).dtype == np.float32 else np.float64
        return (us, v, float(m["global_bias"]), np.ascontiguousarray(m["user_bias"], dtype=np.float32),
                np.ascontiguousarray(m["item_bias"], dtype=np.float32), out_t)

    def predict(self, ratings):
        """als: U @ V.T float64, argument ignored (algo.py:143-147); svd: u@(s*vt)+biases (algo.py:363-368)."""
        if self.algorithm == "knn":
            return self._predict_knn(ratings)
        if self.algorithm == "als" and self._need_model() is not None and self._resident is not None:
            res = self._resident                      # U @ V.T from the device copy (float64 out, as the reference)
            out = np.empty((res.n_users, res.n_items), dtype=np.float64)
            d_out = _DevBuf(res.h, out.nbytes)
            res.h.call("rl_lowrank_dev", res.u.ptr, res.v.ptr, _ffi.RL_F32, res.n_users, res.n_items, res.k, 0.0, None, None,
                       d_out.ptr, _ffi.RL_F64)
            res.h.call("rl_memcpy_d2h", _ffi.ptr(out), d_out.ptr, out.nbytes)
            d_out.free()
            return out
        a, b, gb, rb, cb, out_t = self._factor_view()
        m, n = a.shape[0], b.shape[0]
        kk = a.shape[1]
        out = np.empty((m, n), dtype=out_t)
        if m and n:
            h = _ffi.default_handle()
            h.call("rl_predict_host", _ffi.ptr(np.ascontiguousarray(a)), _ffi.ptr(np.ascontiguousarray(b)),
                   _ffi.RL_F64, m, n, kk, gb, _ffi.ptr(rb), _ffi.ptr(cb), _ffi.ptr(out), _rl_dtype(out))
        if self.algorithm == "svd" and self.use_sparse:
            return sparse.csr_matrix(out)
        return out

    # ------------------------------------------------------------------------------------ recommend
    def recommend(self, ratings, n: int = 10, exclude_rated: bool = True) -> np.ndarray:
        """predict + top_n fused on the device: the score matrix is never formed (algo.py:172-208)."""
        if n <= 0:
            raise ValueError(f"n must be positive. Got n={n}.")
        if self.algorithm == "knn":            # the reference composes predict -> top_n (algo.py:172-201)
            pred = self._predict_knn(ratings)
            if not exclude_rated:
                return np.argsort(-pred, axis=1, kind="stable")[:, :n].astype(np.int64)
            return top_n(pred, np.asarray(_dense(ratings)), n=n)
        if self.algorithm == "als" and self._need_model() is not None and self._resident is not None:
            out = self._recommend_resident(ratings, int(n), exclude_rated)
            if out is not None:
                return out
        a, b, gb, rb, cb, _ = self._factor_view()
        n_users, n_items = a.shape[0], b.shape[0]
        if exclude_rated and tuple(ratings.shape) != (n_users, n_items):
            raise ValueError(
                "Estimated and known matrices must have the same shape. "
                f"Got {(n_users, n_items)} and {tuple(ratings.shape)}."
            )
        if n_users == 0 or n_items == 0:
            return np.empty((0, 0), dtype=int)
        n_dev = min(int(n), n_items)
        indptr = indices = None
        if exclude_rated:
            indptr, indices = observed_pattern(ratings)
        h = _ffi.default_handle()
        if a.shape[1] > MAX_FACTORS or n_dev > MAX_TOP_N:
            # outside the fused kernels (SVD models default to min(shape) // 4 factors; n > 128): score rows in chunks with
            # rl_predict_host and select with rl_topn_dense_host, MAX_TOP_N items per pass (selected items join the mask)
            idx = self._recommend_unfused(h, a, b, gb, rb, cb, indptr, indices, n_dev)
            if not exclude_rated:
                return idx.astype(np.int64)
            n_valid = np.minimum(n_dev, n_items - np.diff(indptr)).astype(np.int32)
            return _finish_topn(idx, n_valid)
        idx = np.empty((n_users, n_dev), dtype=np.int32)
        h.call("rl_recommend_host", _ffi.ptr(np.ascontiguousarray(a)), _ffi.ptr(np.ascontiguousarray(b)), _ffi.RL_F64,
               n_users, n_items, a.shape[1], _ffi.ptr(indptr), _ffi.ptr(indices), gb, _ffi.ptr(rb), _ffi.ptr(cb),
               n_dev, _ffi.ptr(idx), None, _ffi.RL_SCORE_AUTO)
        if not exclude_rated:
            return idx.astype(np.int64)   # reference: argsort result, int64 (algo.py:207)
        n_valid = np.minimum(n_dev, n_items - np.diff(indptr)).astype(np.int32)
        return _finish_topn(idx, n_valid)

    def _recommend_resident(self, ratings, n, exclude_rated):
        """Fused scoring straight from the device copy of the factors; the training pattern is reused when `ratings` is
        the matrix the model was fitted on.  Returns None when the shape needs the general path."""
        res = self._resident
        n_users, n_items = res.n_users, res.n_items
        if exclude_rated and tuple(ratings.shape) != (n_users, n_items):
            raise ValueError(
                "Estimated and known matrices must have the same shape. "
                f"Got {(n_users, n_items)} and {tuple(ratings.shape)}."
            )
        n_dev = min(n, n_items)
        if n_dev > MAX_TOP_N:
            return None
        h = res.h
        ip = ix = None
        indptr = None
        tmp = []
        reuse = False
        tick = _Tick("recommend")
        if exclude_rated:
            indptr, indices = observed_pattern(ratings)
            tick("pattern")
            reuse = res.pattern_key == _ResidentALS.key_of(ratings, indptr, indices)
            if reuse:
                ip, ix = res.ip, res.ix
            else:
                ip, ix = _DevBuf(h, indptr.nbytes), _DevBuf(h, max(indices.nbytes, 16))
                tmp = [ip, ix]
                h.call("rl_memcpy_h2d", ip.ptr, _ffi.ptr(indptr), indptr.nbytes)
                if len(indices):
                    h.call("rl_memcpy_h2d", ix.ptr, _ffi.ptr(indices), indices.nbytes)
        idx = np.empty((n_users, n_dev), dtype=np.int32)
        try:
            # scored in chunks, the rows of a chunk travel home underneath the next one
            h.call("rl_recommend_resident_host", res.u.ptr, n_users, res.v.ptr, n_items, res.k, ip.ptr if ip else None,
                   ix.ptr if ix else None, n_dev, _ffi.ptr(idx), _ffi.RL_SCORE_AUTO)
            tick("score + d2h")
        finally:
            for b in tmp:
                b.free()
        if not exclude_rated:
            return idx.astype(np.int64)
        # trimmed as the reference trims (algo.py:629-639): width = the most unseen items any row has, capped at n
        if reuse and res.min_row_nnz is not None:
            least = res.min_row_nnz
        else:
            least = int(np.diff(indptr).min()) if n_users else 0
            if reuse:
                res.min_row_nnz = least
        widest = max(0, min(n_dev, n_items - least))
        out = idx[:, :widest] if widest else np.empty((n_users, 0), dtype=int)
        tick("finish")
        tick.done()
        return out

    @staticmethod
    def _recommend_unfused(h, a, b, gb, rb, cb, indptr, indices, n_dev, max_cells: int = 1 << 26) -> np.ndarray:
        """predict + top_n as two device calls per row chunk (algo.py:172-201 literally); any factor count, any n."""
        n_users, n_items, kk = a.shape[0], b.shape[0], a.shape[1]
        a = np.ascontiguousarray(a, dtype=np.float64)
        b = np.ascontiguousarray(b, dtype=np.float64)
        idx = np.full((n_users, n_dev), -1, dtype=np.int32)
        rows = max(1, min(n_users, max_cells // max(1, n_items)))
        for r0 in range(0, n_users, rows):
            r1 = min(n_users, r0 + rows)
            est = np.empty((r1 - r0, n_items), dtype=np.float64)
            h.call("rl_predict_host", _ffi.ptr(a[r0:r1]), _ffi.ptr(b), _ffi.RL_F64, r1 - r0, n_items, kk, gb,
                   _ffi.ptr(None if rb is None else np.ascontiguousarray(rb[r0:r1])), _ffi.ptr(cb), _ffi.ptr(est), _ffi.RL_F64)
            known = np.zeros((r1 - r0, n_items), dtype=np.float32)
            if indptr is not None:
                for r in range(r0, r1):
                    known[r - r0, indices[indptr[r]:indptr[r + 1]]] = 1.0
            done = 0
            while done < n_dev:
                take = min(MAX_TOP_N, n_dev - done)
                part = np.empty((r1 - r0, take), dtype=np.int32)
                nv = np.empty(r1 - r0, dtype=np.int32)
                h.call("rl_topn_dense_host", _ffi.ptr(est), _ffi.RL_F64, _ffi.ptr(known), _ffi.RL_F32, r1 - r0, n_items, take,
                       _ffi.ptr(part), _ffi.ptr(nv))
                idx[r0:r1, done:done + take] = part
                done += take
                if done < n_dev:                      # the next pass must not return these again
                    rr, cc = np.nonzero(part >= 0)
                    known[rr, part[rr, cc]] = 1.0
        return idx

    # ---------------------------------------------------------------------------------- persistence
    def save(self, path: str) -> None:
        import joblib
        with open(path, "wb") as f:
            joblib.dump(self, f)

    @classmethod
    def load(cls, path: str) -> "RecommenderSystem":
        import joblib
        with open(path, "rb") as f:
            obj = joblib.load(f)
        if not isinstance(obj, cls):
            raise ValueError(f"Loaded object is not a {cls.__name__}")
        return obj

    def get_params(self) -> Dict[str, Any]:
        return {"algorithm": getattr(self, "algorithm", None), "use_sparse": getattr(self, "use_sparse", False)}

    def clone(self) -> "RecommenderSystem":
        return RecommenderSystem(algorithm=self.algorithm, random_state=self.random_state, use_sparse=self.use_sparse)

    def reset(self) -> None:
        self._drop_resident()
        self._model = None
        self._fitted = False

    def __del__(self):
        try:
            self._drop_resident()
        except Exception:
            pass


MAX_SVD_SMALL_DIM = 4096   # rl_svd_topk_dev: Jacobi eigen-solver on the Gram matrix of the smaller dimension


def _check_svd_shape(shape):
    if min(shape) > MAX_SVD_SMALL_DIM:
        raise ValueError(f"the device SVD needs min(n_users, n_items) <= {MAX_SVD_SMALL_DIM}, got shape {tuple(shape)}")


def _reconstruct_device(mat32: np.ndarray, k: int) -> np.ndarray:
    _check_svd_shape(mat32.shape)
    out = np.empty_like(mat32)
    h = _ffi.default_handle()
    h.call("rl_svd_reconstruct_host", _ffi.ptr(mat32), mat32.shape[0], mat32.shape[1], int(k), _ffi.ptr(out),
           None, None, None)
    return out


def svd_reconstruct(mat, *, k: Optional[int] = None, random_state: Optional[int] = None, use_sparse: bool = True,
                    chunk_size: Optional[int] = None):
    """Rank-k reconstruction of ``mat`` (algo.py:412-530) with the SVD done on the device.

    ``random_state`` is accepted for signature compatibility; the device factorisation is deterministic.
    """
    if mat.ndim != 2:
        raise ValueError(f"Input matrix must be 2D (users x items), got {mat.ndim}D.")
    n_users, n_items = mat.shape
    if n_users == 0 or n_items == 0:
        raise ValueError("Input matrix cannot have zero dimensions.")
    k = max(1, min(n_users, n_items) // 4) if k is None else k
    if not 1 <= k <= min(n_users, n_items):
        raise ValueError(
            f"k must be between 1 and min(n_users, n_items)={min(n_users, n_items)}, got k={k}."
        )
    dense = np.asarray(_dense(mat))
    if not dense.any():
        return np.zeros_like(dense)          # all-zero guard, input dtype (algo.py:489-490)
    if chunk_size:
        # per-chunk rank-k estimator of algo.py:491-505 (a different estimator from the unchunked call)
        out = np.zeros_like(dense)
        for start in range(0, n_users, chunk_size):
            chunk = dense[start:start + chunk_size]
            if not chunk.any():
                continue
            kc = min(int(k), min(chunk.shape))
            out[start:start + chunk.shape[0]] = _reconstruct_device(
                np.ascontiguousarray(chunk, dtype=np.float32), kc)
        return out
    recon = _reconstruct_device(np.ascontiguousarray(dense, dtype=np.float32), int(k))
    return sparse.csr_matrix(recon) if use_sparse else recon


def top_n(est, known, *, n: int = 10) -> np.ndarray:
    """Top-n unseen items per user from dense/sparse score and known matrices (algo.py:587-659)."""
    if n <= 0:
        raise ValueError(f"n must be positive. Got n={n}.")
    if tuple(est.shape) != tuple(known.shape):
        raise ValueError(
            f"Estimated and known matrices must have the same shape. Got {tuple(est.shape)} and {tuple(known.shape)}."
        )
    n_users, n_items = est.shape
    if n_users == 0 or n_items == 0:
        return np.empty((0, 0), dtype=int)
    est_d = _as_float_matrix(est)
    known_d = _as_float_matrix(known, default=np.float32)
    n_dev = min(int(n), n_items)
    if n_dev > MAX_TOP_N:
        raise NotImplementedError(f"n={n}: the selection kernels keep at most {MAX_TOP_N} items per user")
    idx = np.empty((n_users, n_dev), dtype=np.int32)
    n_valid = np.empty(n_users, dtype=np.int32)
    h = _ffi.default_handle()
    h.call("rl_topn_dense_host", _ffi.ptr(est_d), _rl_dtype(est_d), _ffi.ptr(known_d), _rl_dtype(known_d),
           n_users, n_items, n_dev, _ffi.ptr(idx), _ffi.ptr(n_valid))
    return _finish_topn(idx, n_valid)


def _dense_errors(predictions, actual):
    """(sum sq err, sum |err|, #observed) over cells with actual > 0, on the device (algo.py:662-743 semantics)."""
    if predictions.shape != actual.shape:
        raise ValueError(
            "Predictions and actual matrices must have the same shape. "
            f"Got {predictions.shape} and {actual.shape}."
        )
    p, a = _as_float_matrix(predictions), _as_float_matrix(actual)
    out3 = np.zeros(3, dtype=np.float64)
    flags = np.zeros(1, dtype=np.int32)
    if p.size:
        h = _ffi.default_handle()
        h.call("rl_dense_error_host", _ffi.ptr(p), _rl_dtype(p), _ffi.ptr(a), _rl_dtype(a), p.size, _ffi.ptr(out3),
               _ffi.ptr(flags))
    if flags[0] & 1:
        raise ValueError("Input matrices must not contain infinite values.")
    if flags[0] & 2:
        raise ValueError("Input matrices must not contain NaN values.")
    return out3


def _model_errors(model, actual):
    """(sum of squared errors, sum |err|, #cells) of an ALS model's predictions ``U V^T`` over the cells with ``actual > 0``,
    straight from the factors (SDDMM on the device, rl_observed_error_dev): the n_users x n_items prediction matrix the
    reference's ``compute_rmse(model.predict(R), R)`` needs (algo.py:143-147, :662-701) is never formed.  `actual`: dense or
    scipy.sparse."""
    if not isinstance(model, RecommenderSystem) or model.algorithm != "als" or not model.is_fitted():
        raise ValueError("the model form of compute_rmse / compute_mae needs a fitted RecommenderSystem('als')")
    if model._resident is None:
        m = model._need_model()
        n_users, n_items, k = m["user_factors"].shape[0], m["item_factors"].shape[0], m["user_factors"].shape[1]
    else:
        n_users, n_items, k = model._resident.n_users, model._resident.n_items, model._resident.k
    if tuple(actual.shape) != (n_users, n_items):
        raise ValueError(f"Shape mismatch: predictions {(n_users, n_items)} vs actual {tuple(actual.shape)}")
    if sparse.issparse(actual):
        a = sparse.csr_matrix(actual)
        if not a.has_canonical_format:
            a = a.copy()
            a.sum_duplicates()
        keep = a.data > 0
        if not keep.all():
            a = sparse.csr_matrix((a.data * keep, a.indices.copy(), a.indptr.copy()), shape=a.shape)
            a.eliminate_zeros()
        indptr = np.ascontiguousarray(a.indptr, dtype=np.int64)
        indices = np.ascontiguousarray(a.indices, dtype=np.int32)
        values = np.ascontiguousarray(a.data, dtype=np.float32)
    else:
        d = np.asarray(actual)
        if not np.isfinite(d).all():
            raise ValueError("Input matrices must not contain NaN values." if np.isnan(d).any()
                             else "Input matrices must not contain infinite values.")
        mask = d > 0
        indptr = np.zeros(n_users + 1, dtype=np.int64)
        np.cumsum(mask.sum(axis=1), out=indptr[1:])
        indices = np.ascontiguousarray(np.nonzero(mask)[1], dtype=np.int32)
        values = np.ascontiguousarray(d[mask], dtype=np.float32)
    cnt = int(indptr[-1])
    if cnt == 0:
        return 0.0, 0.0, 0
    h = _ffi.default_handle()
    res, own = model._resident, None
    if res is None:                        # host model (edited or loaded): factors go up for this call
        m = model._need_model()
        own = res = _ResidentALS(h, n_users, n_items, k)
        res.upload_factors(np.ascontiguousarray(m["user_factors"]), np.ascontiguousarray(m["item_factors"]))
    bufs = [_DevBuf(h, indptr.nbytes), _DevBuf(h, indices.nbytes), _DevBuf(h, values.nbytes), _DevBuf(h, 16)]
    out2 = np.zeros(2, dtype=np.float64)
    try:
        for b, arr in zip(bufs[:3], (indptr, indices, values)):
            h.call("rl_memcpy_h2d", b.ptr, _ffi.ptr(arr), arr.nbytes)
        h.call("rl_observed_error_dev", res.u.ptr, res.v.ptr, k, n_users, bufs[0].ptr, bufs[1].ptr, bufs[2].ptr, bufs[3].ptr)
        h.call("rl_memcpy_d2h", _ffi.ptr(out2), bufs[3].ptr, 16)
    finally:
        for b in bufs:
            b.free()
        if own is not None:
            own.free()
    return float(out2[0]), float(out2[1]), cnt


def compute_rmse(predictions, actual) -> float:
    """RMSE over cells with ``actual > 0`` (algo.py:662-701); 0.0 when there is none.  `predictions` may also be a fitted
    ``RecommenderSystem('als')``: the error of ``U V^T`` over the observed cells then comes straight from the factors
    (no dense prediction matrix; `actual` dense or scipy.sparse)."""
    if isinstance(predictions, RecommenderSystem):
        sse, _, cnt = _model_errors(predictions, actual)
    else:
        sse, _, cnt = _dense_errors(predictions, actual)
    return float(np.sqrt(sse / cnt)) if cnt > 0 else 0.0


def compute_mae(predictions, actual) -> float:
    """MAE over cells with ``actual > 0`` (algo.py:704-743); `predictions` as in compute_rmse."""
    if isinstance(predictions, RecommenderSystem):
        _, sae, cnt = _model_errors(predictions, actual)
    else:
        _, sae, cnt = _dense_errors(predictions, actual)
    return float(sae / cnt) if cnt > 0 else 0.0


__all__ = ["RecommenderSystem", "svd_reconstruct", "top_n", "compute_rmse", "compute_mae"]

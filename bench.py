#!/usr/bin/env python
"""bench.py -- headline benchmark of the recsys-lite hot path on B200 (contract: see the task statement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload C4|C3|tiny]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one full ALS iteration (user half-step, factor exchange, item half-step, factor exchange;
reference algo.py:312-332) followed by top-10 scoring of ALL users with seen items masked (algo.py:172-201), on
the BASELINE.json configuration the metric is quoted on: C4 = 1 M users x 100 K items, ~50 M nnz, k = 64
(fits one GPU).  Inputs are synthetic (recsys_lite_b200.synth, seed 4) and much larger than L2 (indices
2 x 200 MB, U 256 MB), so no L2 flush is needed between timed iterations.

value        whole-job iterations/s with every input resident in HBM (CUDA events, max over ranks)
e2e          the same step through the host-buffer C ABI (rl_als_fit_host + rl_recommend_host at N = 1; the
             same kernels with per-rank H2D / D2H at N > 1), host<->device copies inside the timed region
roofline     dominant kernel of the step (the scoring kernel); roofline_als: the two gather/solve launches
cpu_baseline the oracle's CSR restatement of the reference on a bounded sample, extrapolated linearly in rows
--impl reference: the CPU leg alone, as its own JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "ALS iterations/sec (both half-steps + top-10 recs for all users)"
UNIT = "iterations/s"

WORKLOADS = {
    "C4": dict(n_users=1_000_000, n_items=100_000, mean_deg=50, k=64, seed=4, init_seed=4, n=10, lam=0.01),
    "C3": dict(n_users=100_000, n_items=20_000, mean_deg=50, k=64, seed=3, init_seed=3, n=10, lam=0.01),
    # skewed variants of SURVEY 8(d): Zipf(1.0) item popularity, log-normal user degrees (load balance, long rows)
    "C4s": dict(n_users=1_000_000, n_items=100_000, mean_deg=50, k=64, seed=14, init_seed=4, n=10, lam=0.01, skew=True),
    "C3s": dict(n_users=100_000, n_items=20_000, mean_deg=50, k=64, seed=13, init_seed=3, n=10, lam=0.01, skew=True),
    "tiny": dict(n_users=20_000, n_items=4_000, mean_deg=30, k=64, seed=1, init_seed=1, n=10, lam=0.01),
    # BASELINE.json configs[4]: ALS only, k = 128, meant for 8 GPUs; every rank builds only its own shards (run_sharded_als).
    # `--gpus 1 --emulate-world 8` runs ONE rank's share of the 8-GPU problem on one GPU (no exchange).
    "C5": dict(n_users=10_000_000, n_items=1_000_000, mean_deg=50, k=128, seed=5, init_seed=5, n=0, lam=0.01, sharded=True),
    "C5tiny": dict(n_users=80_000, n_items=16_000, mean_deg=30, k=128, seed=5, init_seed=5, n=0, lam=0.01, sharded=True),
}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], tflops=d.get("bf16_tflops_sustained", d["bf16_tflops"]), source="measured")
    return dict(hbm_gbs=6650.0, tflops=1400.0, source="fallback")


def load_traffic():
    """dram bytes per launch / tensor-pipe activity from the COMMITTED ncu --set full capture (profiles/r2_traffic.json: taken at
    the last step of `bench.py --steps 20 --warmup 5`, see profiles/README.md), or {}.  Static profile data, not a measurement of
    the current run: the line says so next to every such number."""
    p = os.path.join(ROOT, "profiles", "r2_traffic.json")
    return json.load(open(p)) if os.path.exists(p) else {}


def sampled_parity(w, indptr, indices, indptr_t, indices_t, u, v, idx_rows, users, items, lam, n):
    """Size-independent checks of the device results against float64 NumPy on sampled rows (no oracle import: this runs in the
    timed arm).  (1) top-n of `users`: ranking of the exact float64 scores of the float32 factors, seen items excluded, ties
    (|ds| <= 4 eps32 max(1,|s|)) accepted as in oracle/compare.py; (2) every sampled user / item row solves its normal equations
    (Y^T Y + lam I) x = Y^T 1 (reference algo.py:318-322 / :328-332): relative residual.  The item rows are checked against the U
    they were solved with (the current one); the user rows only bound the drift since their solve (V has moved on by one item
    half-step), so (2) reports them separately."""
    u64, v64 = u.astype(np.float64), v.astype(np.float64)
    bad = tied = 0
    tol_rel = 4 * float(np.finfo(np.float32).eps)
    for r, got in zip(users, idx_rows):
        sc = v64 @ u64[r]
        seen = indices[indptr[r]:indptr[r + 1]]
        sc_m = sc.copy()
        sc_m[seen] = -np.inf
        order = np.lexsort((np.arange(len(sc_m)), -sc_m))[:n]
        order = order[np.isfinite(sc_m[order])]
        got = got[got >= 0]
        if np.array_equal(order, got):
            continue
        ok = len(got) == len(order) and len(set(got.tolist())) == len(got) and not np.isin(got, seen).any()
        if ok:
            ok = bool(np.all(np.abs(sc[got] - sc[order]) <= tol_rel * np.maximum(1.0, np.abs(sc[order]))))
        tied += ok
        bad += not ok
    res_items = 0.0
    for i in items:
        rows = indices_t[indptr_t[i]:indptr_t[i + 1]]
        if len(rows) == 0:
            continue
        y = u64[rows]
        a = y.T @ y + lam * np.eye(y.shape[1])
        b = y.sum(axis=0)
        res_items = max(res_items, float(np.linalg.norm(a @ v64[i] - b) / max(np.linalg.norm(b), 1e-300)))
    return {"topn_rows_checked": int(len(users)), "topn_rows_tied": int(tied), "topn_rows_bad": int(bad),
            "als_item_rows_checked": int(len(items)), "als_item_max_rel_residual": res_items,
            "what": "float64 NumPy on sampled rows of the final state: top-n ranking (seen items masked, ties within 4 eps32) and the "
                    "normal equations (Y^T Y + lam I) v_i = Y^T 1 of the last item half-step"}



def make_inputs(w):
    from recsys_lite_b200 import synth
    from recsys_lite_b200.csr import transpose_pattern
    indptr, indices = synth.synth_pattern(w["n_users"], w["n_items"], w["mean_deg"], w["seed"], skew=w.get("skew", False))
    indptr_t, indices_t = transpose_pattern(indptr, indices, w["n_items"])
    rng = np.random.default_rng(w["init_seed"])          # float32 uniform [0,1) like np.random.rand, drawn U then V
    u0 = rng.random((w["n_users"], w["k"]), dtype=np.float32)
    v0 = rng.random((w["n_items"], w["k"]), dtype=np.float32)
    return indptr, indices, indptr_t, indices_t, u0, v0


def als_bytes(n_rows, nnz, k):
    """algorithmic bytes of one half-step: gathered rows + indices + indptr + output, fp32 factors."""
    return nnz * k * 4 + nnz * 4 + (n_rows + 1) * 8 + n_rows * k * 4


# ------------------------------------------------------------------------------------------- clocks sampling
class ClockSampler:
    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.dev = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.dev, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake_slowdown",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.dev, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.dev)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.dev)
                for bit, nm in names.items():
                    if mask & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            self._stop.wait(0.02)

    def start(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ----------------------------------------------------------------------------------------------- CPU leg
def cpu_reference_leg(w, inputs, n_user_rows, n_item_rows, n_topn_users, target=None):
    """The oracle's CSR restatement of algo.py:314-332 + blocked predict/top_n (algo.py:143-147, :533-659) on a
    bounded contiguous sample, extrapolated linearly in rows.  Returns (iterations/s, seconds, description)."""
    from oracle import als as o_als, topn as o_topn       # the CPU legs are the only place bench.py touches oracle/
    indptr, indices, indptr_t, indices_t, u0, v0 = inputs
    nu, ni = w["n_users"], w["n_items"]
    tu, ti = target if target else (nu, ni)      # sizes the per-row costs are extrapolated to
    n_user_rows, n_item_rows, n_topn_users = min(n_user_rows, nu), min(n_item_rows, ni), min(n_topn_users, nu)
    u = u0.astype(np.float64)
    v = v0.astype(np.float64)
    t0 = time.perf_counter()
    o_als.half_step_csr(indptr, indices, v, u, w["lam"], rows=range(n_user_rows))
    t1 = time.perf_counter()
    o_als.half_step_csr(indptr_t, indices_t, u, v, w["lam"], rows=range(n_item_rows))
    t2 = time.perf_counter()
    if w["n"] > 0:
        o_topn.recommend_blocked(u, v, indptr, indices, n=w["n"], users=np.arange(n_topn_users), block=256)
    t3 = time.perf_counter()
    t_iter = (t1 - t0) * tu / n_user_rows + (t2 - t1) * ti / n_item_rows + (t3 - t2) * tu / n_topn_users
    desc = (f"oracle CSR restatement of algo.py:314-332 on the first {n_user_rows} users ({t1 - t0:.2f}s) and "
            f"{n_item_rows} items ({t2 - t1:.2f}s) + blocked U@V.T/top-{w['n']} for {n_topn_users} users ({t3 - t2:.2f}s), "
            f"extrapolated linearly in rows to {tu} x {ti}; single Python process, BLAS threads as configured")
    return 1.0 / t_iter, t3 - t0, desc


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
    except Exception:
        return os.cpu_count() or 1


class all_host_threads:
    """BLAS / OpenMP thread pools set to every host core for the CPU leg, whatever the launcher exported (torchrun sets
    OMP_NUM_THREADS=1 for its workers, which would throttle the reference arm at N > 1)."""

    def __enter__(self):
        self.ctx = None
        try:
            from threadpoolctl import threadpool_limits
            self.ctx = threadpool_limits(limits=os.cpu_count() or 1)
            self.ctx.__enter__()
        except Exception:
            self.ctx = None
        return self

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.ctx.__exit__(*exc)
        return False


def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    target = None
    if w.get("sharded"):
        # problems that are never built whole: the per-row costs are timed on a 1 / 50 scale model with the SAME row lengths
        # (users x items scaled together keep ~50 nnz per user and ~500 per item) and extrapolated to the full row counts
        target = (w["n_users"], w["n_items"])
        w = dict(w, n_users=w["n_users"] // 50, n_items=w["n_items"] // 50)
    inputs = make_inputs(w)
    # >= 20 K users / 20 K items per half-step and >= 2 K users for top-N, whatever --steps says; the sample is
    # repeated at most 3 times (+1 warm-up) so that the arm ends within a few minutes
    nu_s, ni_s, nt_s = 20_000, 20_000, 2_000
    reps, warm = max(1, min(args.steps, 3)), min(args.warmup, 1)
    vals, secs, desc = [], [], ""
    with all_host_threads():
        for i in range(warm + reps):
            v, s, desc = cpu_reference_leg(w, inputs, nu_s, ni_s, nt_s, target=target)
            if i >= warm:
                vals.append(v)
                secs.append(s)
        cores = blas_threads()
    value = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 / value, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, w), "where": "host CPU",
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "sample_seconds_per_step": float(np.mean(secs)), "sample_repeats": reps,
    }
    _emit(line)


def workload_config(args, w):
    return {"workload": f"{args.workload}: implicit ALS k={w['k']} one full iteration + top-{w['n']} for all users, "
                        f"{w['n_users']} users x {w['n_items']} items, ~{w['mean_deg']} nnz/user ({'skewed: Zipf item popularity, log-normal degrees' if w.get('skew') else 'uniform synthetic'}, seed {w['seed']})",
            "lambda": w["lam"],
            "l2": "inputs larger than L2 (indices 2x%d MB, U %d MB); no flush" % (
                w["n_users"] * w["mean_deg"] * 4 // 2**20, w["n_users"] * w["k"] * 4 // 2**20)}


# ----------------------------------------------------------------------------------------------- GPU leg
def run_b200(args, w):
    import torch
    import torch.distributed as dist
    from recsys_lite_b200 import _ffi
    from recsys_lite_b200.sharding import ShardedProblem, make_plan

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {args.gpus}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    h = _ffi.Handle(local)
    stream = torch.cuda.Stream(device=device)               # everything (kernels, events, NCCL sync) on one stream
    torch.cuda.set_stream(stream)
    h.set_stream(stream.cuda_stream)
    peaks = load_peaks()
    prec = _ffi.RL_PREC_FP64 if args.precision == "fp64" else _ffi.RL_PREC_FP32_IR

    inputs = make_inputs(w)
    indptr, indices, indptr_t, indices_t, u0, v0 = inputs
    nu, ni, k, n = w["n_users"], w["n_items"], w["k"], w["n"]
    plan = make_plan(indptr, indptr_t, rank, world)
    prob = ShardedProblem(h, plan, indptr, indices, indptr_t, indices_t, k, device)
    prob.set_factors(u0, v0)
    fused = prob.enable_peer_stores() if (world > 1 and not args.no_peer_stores) else False
    ulo, uhi = plan.users
    ilo, ihi = plan.items
    idx_out = torch.empty((uhi - ulo, n), dtype=torch.int32, device=device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ev = lambda: torch.cuda.Event(enable_timing=True)      # noqa: E731
    marks, overflow_rows = [], []

    def step(record):
        e = [ev() for _ in range(5)] if record else None
        if record:
            e[0].record()
        prob.user_half_step(w["lam"], prec, args.ir_steps)
        if record:
            e[1].record()
        prob.exchange_users()
        prob.item_half_step(w["lam"], prec, args.ir_steps)
        if record:
            e[2].record()
        prob.exchange_items()
        if record:
            e[3].record()
        prob.recommend(n, idx_out, args.score_mode)
        if record:
            e[4].record()
            marks.append(e)
            overflow_rows.append(int(h.lib.rl_last_overflow_rows(h.h)))

    for _ in range(args.warmup):
        step(False)
    barrier()
    clocks = ClockSampler(local)
    clocks.start()
    launches0 = h.launches
    t_begin, t_end = ev(), ev()
    barrier()
    t_begin.record()
    for _ in range(args.steps):
        step(True)
    t_end.record()
    barrier()
    clk = clocks.stop()
    launches = h.launches - launches0
    total_ms = t_begin.elapsed_time(t_end)
    seg = np.array([[e[i].elapsed_time(e[i + 1]) for i in range(4)] for e in marks])   # user(+0), xchg+item, xchg, score
    t_user, t_item_x, t_x2, t_score = seg.mean(axis=0)
    per_rank = None
    if world > 1:
        tt = torch.tensor([total_ms, t_user, t_item_x, t_x2, t_score, float(h.lib.rl_last_overflow_rows(h.h))],
                          dtype=torch.float64, device=device)
        allt = [torch.empty_like(tt) for _ in range(world)]
        dist.all_gather(allt, tt)
        # diagnostics (stragglers show up here): per rank [user, exchange+item, exchange, score] ms, overflow rows of the last scoring call
        per_rank = [[round(float(x), 3) for x in t.tolist()[1:]] for t in allt]
        tt = tt[:5]
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total_ms, t_user, t_item_x, t_x2, t_score = tt.tolist()
    ms_per_step = total_ms / args.steps
    value = 1000.0 / ms_per_step

    # ---- correctness of what was just timed (sampled, float64 NumPy; replicas compared across ranks) --------------
    parity = None
    if not args.no_parity:
        rs = np.random.default_rng(1234 + rank)
        users = np.sort(rs.choice(np.arange(ulo, uhi), size=min(64, uhi - ulo), replace=False))
        items = np.sort(rs.choice(np.arange(ilo, ihi), size=min(64, ihi - ilo), replace=False))
        u_host = prob.U[:, :k].cpu().numpy()
        v_host = prob.V[:, :k].cpu().numpy()
        idx_rows = idx_out[torch.from_numpy(users - ulo).to(device)].cpu().numpy()
        parity = sampled_parity(w, indptr, indices, indptr_t, indices_t, u_host, v_host, idx_rows, users, items, w["lam"], n)
        if world > 1:
            # every rank must hold the same replicas after the fused exchange (NVLink peer stores): compare checksums
            cs_ = torch.stack([prob.U[:, :k].double().sum(), prob.V[:, :k].double().sum(), prob.U[:, :k].double().abs().sum()])
            allc = [torch.empty_like(cs_) for _ in range(world)]
            dist.all_gather(allc, cs_)
            same = all(torch.equal(allc[0], c) for c in allc)
            agg = torch.tensor([float(parity["topn_rows_bad"]), parity["als_item_max_rel_residual"], 0.0 if same else 1.0],
                               dtype=torch.float64, device=device)
            dist.all_reduce(agg, op=dist.ReduceOp.MAX)
            parity["topn_rows_bad"] = int(agg[0].item())
            parity["als_item_max_rel_residual"] = float(agg[1].item())
            parity["replicas_identical_across_ranks"] = agg[2].item() == 0.0
            parity["topn_rows_checked"] *= world
            parity["als_item_rows_checked"] *= world
        parity["ok"] = bool(parity["topn_rows_bad"] == 0 and parity["als_item_max_rel_residual"] <= 1e-4
                            and parity.get("replicas_identical_across_ranks", True))
        del u_host, v_host

    # ---- end-to-end leg: host buffers, copies inside the timed region -------------------------------------
    e2e = run_e2e(args, w, h, prob, plan, inputs, prec, rank, world, device, barrier)

    api = run_api_leg(args, w, inputs, prob, k) if (world == 1 and not args.no_api) else None

    line = None
    if rank == 0:
        traffic = load_traffic() if (world == 1 and args.workload == "C4") else {}
        flops_score = 2.0 * (uhi - ulo) * ni * k
        b_user = als_bytes(uhi - ulo, prob.nnz_users, k)
        b_item = als_bytes(ihi - ilo, prob.nnz_items, k)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32 factors; ALS solve " + ("f64" if prec else f"f32 + {args.ir_steps} f64 refinement steps")
                     + "; scores f64",
            "data": "synthetic", "config": workload_config(args, w), "where": f"{world} x B200",
            "exchange": ("none (1 GPU)" if world == 1 else "fused: solve kernel stores rows into the peers' replicas over NVLink "
                         "(CUDA IPC) + one-element all-reduce as barrier" if fused else "NCCL broadcasts of the row ranges"),
            "als_iterations_per_s": 1000.0 / (t_user + t_item_x + t_x2),
            "top10_users_per_s": nu / (t_score / 1000.0),
            "ms": {"user_half_step": t_user, "exchange_U_plus_item_half_step": t_item_x, "exchange_V": t_x2, "score_topn": t_score},
            "per_step": {"score_topn_ms": [round(float(x), 3) for x in seg[:, 3]],
                         "overflow_rows": overflow_rows,
                         "what": "rank 0, every timed step (each step is one more ALS iteration on the same factors: the scoring "
                                 "kernel sees more and more converged factors); overflow_rows = rows finished by the exact kernel"},
            "roofline": {"kernel": "score_topn (rank 0 shard)", "bound": "tensor",
                         "achieved": flops_score / (t_score * 1e-3) / 1e12, "peak": peaks["tflops"], "unit": "TFLOP/s",
                         "frac": flops_score / (t_score * 1e-3) / 1e12 / peaks["tflops"], "traffic": traffic.get("score_topn"),
                         "static_profile": {"source": traffic.get("source"), "tensor_pipe_active_pct": (traffic.get("tensor_pipe_active_pct") or {}).get("score_topn"),
                                            "what": "traffic and tensor-pipe activity come from the committed ncu capture named in `source`, not from this run"},
                         "note": "algorithmic flops 2*n_users*n_items*k over the whole scoring call (preprocessing of V, the candidate pass "
                                 "score_tc2_kernel, exact re-scoring); one fp16 tcgen05 product per score on centred factors (score_tc2.cu), so "
                                 "issued MMA flops = algorithmic flops; the kernel is bound by the integer/compare pipe of the selection, not by the tensor pipe",
                         "peak_source": peaks["source"] + " bf16 sustained"},
            "roofline_als": {
                "bound": "hbm", "unit": "GB/s", "peak": peaks["hbm_gbs"], "peak_source": peaks["source"],
                "user_half_step": {"bytes": b_user, "achieved": b_user / (t_user * 1e-3) / 1e9,
                                   "frac": b_user / (t_user * 1e-3) / 1e9 / peaks["hbm_gbs"],
                                   "traffic": traffic.get("als_user_half_step"),
                                   "note": "V (25.6 MB) is L2-resident: the gather is served by L2, dram traffic is indices + output"},
                "item_half_step_incl_exchange": {"bytes": b_item, "achieved": b_item / (t_item_x * 1e-3) / 1e9,
                                                 "frac": b_item / (t_item_x * 1e-3) / 1e9 / peaks["hbm_gbs"],
                                                 "traffic": traffic.get("als_item_half_step")}},
            "clocks": clk, "gpu_launches": int(launches), "e2e": e2e, "e2e_api": api, "parity": parity, "ms_per_rank": per_rank,
        }
        if world == 1 and not args.no_cpu:
            with all_host_threads():
                v, s, desc = cpu_reference_leg(w, inputs, 60_000, 60_000, 6_000)
                cores = blas_threads()
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc, "seconds": s}
        _emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return line


def run_sharded_als(args, w):
    """ALS-only workloads whose pattern is never built whole (C5): one ALS iteration per step on this rank's shards, the factor
    exchange fused into the solve kernels (NVLink peer stores) when there is more than one rank."""
    import torch
    import torch.distributed as dist
    from recsys_lite_b200 import _ffi, synth
    from recsys_lite_b200.sharding import ShardedProblem, ShardPlan
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    emu = args.emulate_world if (world == 1 and args.emulate_world > 1) else world
    rank_e = args.emulate_rank if emu != world else rank
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    h = _ffi.Handle(local)
    stream = torch.cuda.Stream(device=device)
    torch.cuda.set_stream(stream)
    h.set_stream(stream.cuda_stream)
    peaks = load_peaks()
    nu, ni, k = w["n_users"], w["n_items"], w["k"]
    t_gen = time.perf_counter()
    ub, ib, up, ux, ip, ix = synth.sharded_problem(nu, ni, w["mean_deg"], w["seed"], rank_e, emu)
    t_gen = time.perf_counter() - t_gen
    plan = ShardPlan(rank_e, emu, ub, ib)
    prob = ShardedProblem(h, plan, None, None, None, None, k, device, shards=(up, ux, ip, ix), n_users=nu, n_items=ni)
    ulo, uhi = plan.users
    ilo, ihi = plan.items
    v0 = np.random.default_rng([w["init_seed"], 999]).random((ni, k), dtype=np.float32)
    u_rows = np.random.default_rng([w["init_seed"], 1000 + rank_e]).random((uhi - ulo, k), dtype=np.float32)
    if emu != world:       # one rank of a larger job on its own: the other ranks' user rows are random stand-ins (never exchanged)
        gen = torch.Generator(device=device)
        gen.manual_seed(1234)
        prob.U[:, :k].uniform_(0.0, 1.0, generator=gen)
    prob.set_factor_rows(u_rows, v0)
    fused = prob.enable_peer_stores() if (world > 1 and not args.no_peer_stores) else False
    real_exchange = world > 1

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ev = lambda: torch.cuda.Event(enable_timing=True)      # noqa: E731
    marks = []

    def step(record):
        e = [ev() for _ in range(3)] if record else None
        if record:
            e[0].record()
        prob.user_half_step(w["lam"], 0, args.ir_steps)
        if real_exchange:
            prob.exchange_users()
        if record:
            e[1].record()
        prob.item_half_step(w["lam"], 0, args.ir_steps)
        if real_exchange:
            prob.exchange_items()
        if record:
            e[2].record()
            marks.append(e)

    for _ in range(args.warmup):
        step(False)
    barrier()
    clocks = ClockSampler(local)
    clocks.start()
    launches0 = h.launches
    t_begin, t_end = ev(), ev()
    barrier()
    t_begin.record()
    for _ in range(args.steps):
        step(True)
    t_end.record()
    barrier()
    clk = clocks.stop()
    launches = h.launches - launches0
    total_ms = t_begin.elapsed_time(t_end)
    seg = np.array([[e[i].elapsed_time(e[i + 1]) for i in range(2)] for e in marks])
    t_user, t_item = seg.mean(axis=0)
    # sampled parity: the item rows of this rank solve their normal equations against the U they were computed from
    rs = np.random.default_rng(77 + rank_e)
    items = np.sort(rs.choice(ihi - ilo, size=min(32, ihi - ilo), replace=False))
    res = 0.0
    v_rows = prob.V[ilo:ihi, :k][torch.from_numpy(items).to(device)].double().cpu().numpy()
    for j, i in enumerate(items):
        rows = ix[ip[i]:ip[i + 1]]
        if len(rows) == 0:
            continue
        y = prob.U[:, :k][torch.from_numpy(rows.astype(np.int64)).to(device)].double().cpu().numpy()
        a = y.T @ y + w["lam"] * np.eye(k)
        b = y.sum(axis=0)
        res = max(res, float(np.linalg.norm(a @ v_rows[j] - b) / max(np.linalg.norm(b), 1e-300)))
    stats = torch.tensor([total_ms, t_user, t_item, res], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
    total_ms, t_user, t_item, res = stats.tolist()
    if rank == 0:
        ms_per_step = total_ms / args.steps
        b_user = als_bytes(uhi - ulo, len(ux), k)
        b_item = als_bytes(ihi - ilo, len(ix), k)
        line = {
            "metric": METRIC, "value": 1000.0 / ms_per_step, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": f"f32 factors; ALS solve f32 + {args.ir_steps} f64 refinement steps", "data": "synthetic",
            "config": {"workload": f"{args.workload}: implicit ALS k={k}, one full iteration per step (no top-N: BASELINE.json configs[4]), "
                                   f"{nu} users x {ni} items, ~{w['mean_deg']} nnz/user (uniform synthetic, seed {w['seed']}, generated shard-wise "
                                   f"on every rank)", "lambda": w["lam"],
                       "l2": "every rank's pattern shard and its factor replicas are far larger than L2; no flush"},
            "where": f"{world} x B200" + (f" (one GPU playing rank {rank_e} of {emu}: its shards of the {emu}-GPU job, no exchange)" if emu != world else ""),
            "exchange": ("none" if not real_exchange else "fused: solve kernel stores rows into the peers' replicas over NVLink (CUDA IPC)"
                         if fused else "NCCL broadcasts of the row ranges"),
            "ms": {"user_half_step_incl_exchange": t_user, "item_half_step_incl_exchange": t_item},
            "roofline": {"kernel": "als_k128_kernel, user half-step (this rank's shard)", "bound": "hbm", "unit": "GB/s", "peak": peaks["hbm_gbs"],
                         "achieved": b_user / (t_user * 1e-3) / 1e9, "frac": b_user / (t_user * 1e-3) / 1e9 / peaks["hbm_gbs"], "traffic": None,
                         "bytes": b_user, "peak_source": peaks["source"]},
            "roofline_als": {"bound": "hbm", "unit": "GB/s", "peak": peaks["hbm_gbs"],
                             "item_half_step": {"bytes": b_item, "achieved": b_item / (t_item * 1e-3) / 1e9,
                                                "frac": b_item / (t_item * 1e-3) / 1e9 / peaks["hbm_gbs"]}},
            "parity": {"als_item_rows_checked": int(len(items)) * world, "als_item_max_rel_residual": res, "ok": bool(res <= 1e-4),
                       "what": "float64 NumPy on sampled item rows: (Y^T Y + lam I) v_i = Y^T 1 against the U of the last user half-step"},
            "shard": {"users": [ulo, uhi], "items": [ilo, ihi], "nnz_user_side": int(len(ux)), "nnz_item_side": int(len(ix)),
                      "host_generation_s": round(t_gen, 1)},
            "clocks": clk, "gpu_launches": int(launches),
            "e2e": None, "cpu_baseline": None,
        }
        _emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_api_leg(args, w, inputs, prob, k):
    """The same step through the drop-in Python surface (what a user of the reference calls): RecommenderSystem("als")
    .fit(R_csr, k, iterations=1, init=...) + .recommend(R_csr, n) on pageable NumPy / SciPy inputs; the model stays in HBM between
    the two calls (api.py: _ResidentALS), the top-n lists come back as the reference's int array."""
    import torch
    from scipy import sparse
    from recsys_lite_b200 import RecommenderSystem
    indptr, indices = inputs[0], inputs[1]
    nu, ni, n = w["n_users"], w["n_items"], w["n"]
    r = sparse.csr_matrix((np.ones(len(indices), np.float32), indices, indptr), shape=(nu, ni))
    u_now = prob.U[:, :k].cpu().numpy().copy()
    v_now = prob.V[:, :k].cpu().numpy().copy()
    t_fit, t_rec = [], []
    steps = max(1, min(args.steps, 5))
    for i in range(1 + steps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        m = RecommenderSystem("als").fit(r, k=k, iterations=1, lambda_=w["lam"], init=(u_now, v_now))
        t1 = time.perf_counter()
        rec = m.recommend(r, n=n)
        t2 = time.perf_counter()
        if i > 0:
            t_fit.append(t1 - t0)
            t_rec.append(t2 - t1)
        del m
    # median over the steps: this leg runs on pageable host memory through Python, one page-fault storm or pool growth in a
    # single step (seen: 1 s) would otherwise be the whole number; the per-step times are in the line
    per_step = np.asarray(t_fit) + np.asarray(t_rec)
    sec = float(np.median(per_step))
    return {"value": 1.0 / sec, "unit": UNIT, "ms_per_step": sec * 1e3, "ms_fit": float(np.median(t_fit)) * 1e3,
            "ms_recommend": float(np.median(t_rec)) * 1e3, "statistic": "median over the steps",
            "ms_per_step_all": [round(float(x) * 1e3, 2) for x in per_step], "steps": steps, "result_shape": list(rec.shape),
            "api": "RecommenderSystem('als').fit(scipy CSR, k=64, iterations=1, init=float32 arrays) + .recommend(R, n=10); pageable "
                   "host inputs, model resident in HBM between the calls, int top-n array back on the host"}


def run_e2e(args, w, h, prob, plan, inputs, prec, rank, world, device, barrier):
    """The step through host buffers.  N = 1: the host-buffer C ABI calls a reference binding would make
    (rl_als_fit_host + rl_recommend_host) on pinned float32 buffers.  N > 1: per-rank H2D of its CSR / CSC ranges
    and the initial factors, the same kernels + exchanges, D2H of its factor ranges and top-n rows."""
    import torch
    from recsys_lite_b200 import _ffi
    indptr, indices, indptr_t, indices_t, u0, v0 = inputs
    nu, ni, k, n = w["n_users"], w["n_items"], w["k"], w["n"]
    steps = max(1, args.steps)
    # every end-to-end step starts from the factors the device-resident steps ended with (the converged state `value` was
    # measured on), not from the random initialisation: the scoring pass sees the same kind of data in both numbers
    u_now = prob.U[:, :k].cpu().numpy().copy()
    v_now = prob.V[:, :k].cpu().numpy().copy()
    if world == 1:
        pin = lambda a: _pinned_copy(a)                    # noqa: E731
        p_ip, p_ix = pin(indptr), pin(indices)      # the item-side pattern is built on the device (as RecommenderSystem.fit does)
        p_u, p_v = pin(u_now), pin(v_now)
        idx = _ffi.pinned_empty((nu, n), np.int32)
        # rl_als_fit_host sends the initial U only when some user row has no observation (every other row is overwritten by
        # the first user half-step before anything reads it); rl_recommend_host uploads the fitted U, V and the seen lists
        u_needed = bool((np.diff(indptr) == 0).any())
        h2d = (p_ip.nbytes + p_ix.nbytes + p_v.nbytes + (p_u.nbytes if u_needed else 0)) + (p_u.nbytes + p_v.nbytes + p_ip.nbytes + p_ix.nbytes)
        d2h = p_u.nbytes + p_v.nbytes + idx.nbytes
        times = []
        for i in range(1 + steps):
            p_u[:] = u_now
            p_v[:] = v_now
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            h.call("rl_als_fit_host", nu, ni, k, _ffi.ptr(p_ip), _ffi.ptr(p_ix), None, None, 1,
                   float(w["lam"]), prec, args.ir_steps, _ffi.ptr(p_u), _ffi.ptr(p_v), _ffi.RL_F32)
            h.call("rl_recommend_host", _ffi.ptr(p_u), _ffi.ptr(p_v), _ffi.RL_F32, nu, ni, k, _ffi.ptr(p_ip), _ffi.ptr(p_ix),
                   0.0, None, None, n, _ffi.ptr(idx), None, args.score_mode)
            torch.cuda.synchronize()
            if i > 0:
                times.append(time.perf_counter() - t0)
        sec = float(np.mean(times))
        return {"value": 1.0 / sec, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": sec * 1e3, "steps": steps,
                "state": "every step starts from the factors of the last device-resident step",
                "api": "rl_als_fit_host(iterations=1) + rl_recommend_host(n=10), pinned float32 host buffers"}
    import torch.distributed as dist
    ulo, uhi = plan.users
    ilo, ihi = plan.items
    hp = {kk: torch.from_numpy(v).pin_memory() for kk, v in prob.host.items()}
    hu, hv = torch.from_numpy(u_now).pin_memory(), torch.from_numpy(v_now).pin_memory()
    out_u = torch.empty((uhi - ulo, k), dtype=torch.float32).pin_memory()
    out_v = torch.empty((ihi - ilo, k), dtype=torch.float32).pin_memory()
    out_idx = torch.empty((uhi - ulo, n), dtype=torch.int32).pin_memory()
    idx_dev = torch.empty((uhi - ulo, n), dtype=torch.int32, device=device)
    # Each rank uploads the initial factors it needs: all of V (the first user half-step gathers from it) but only ITS
    # range of U -- every user row with observations is overwritten by its owner's half-step (and arrives through the
    # exchange) before anyone reads it, and rows without observations are read by nobody but their owner's scoring pass.
    # (the rank's rows of the initial U only travel when one of them has no observation: see rl_als_fit_host)
    u_needed = bool((np.diff(prob.host["up"]) == 0).any())
    h2d = sum(t.numel() * t.element_size() for t in hp.values()) + ((uhi - ulo) * k * 4 if u_needed else 0) + hv.numel() * 4
    d2h = out_u.numel() * 4 + out_v.numel() * 4 + out_idx.numel() * 4
    # Uploads and downloads run on a second stream: the user half-step starts on the first eighth of this rank's rows
    # while the rest of its pattern and factor rows travel (row ranges holding 1/8, 1/8, 1/4, 1/2 of its observations),
    # the item-side pattern arrives underneath it, and U / V go home underneath the item half-step / the scoring.
    main, cs = torch.cuda.current_stream(), torch.cuda.Stream(device=device)
    up_host = prob.host["up"]
    nloc, nnz_loc = uhi - ulo, int(up_host[-1])
    cuts = [0] + [int(np.searchsorted(up_host, f * nnz_loc)) for f in (0.125, 0.25, 0.5)] + [nloc]
    parts = [(a, b) for a, b in zip(cuts[:-1], cuts[1:]) if b > a] if nnz_loc >= (1 << 20) else [(0, nloc)]
    times = []
    for i in range(1 + steps):
        barrier()
        t0 = time.perf_counter()
        ev_parts = []
        cs.wait_stream(main)
        with torch.cuda.stream(cs):
            prob.V[:, :k].copy_(hv, non_blocking=True)
            prob.up.copy_(hp["up"], non_blocking=True)
            for r0, r1 in parts:
                o0, o1 = int(up_host[r0]), int(up_host[r1])
                prob.ux[o0:o1].copy_(hp["ux"][o0:o1], non_blocking=True)
                if u_needed:
                    prob.U[ulo + r0:ulo + r1, :k].copy_(hu[ulo + r0:ulo + r1], non_blocking=True)
                ev_parts.append(cs.record_event())
            prob.ip.copy_(hp["ip"], non_blocking=True)
            prob.ix.copy_(hp["ix"], non_blocking=True)
            ev_item = cs.record_event()
        for (r0, r1), ev in zip(parts, ev_parts):
            main.wait_event(ev)
            prob.user_half_step(w["lam"], prec, args.ir_steps, rows=(r0, r1))
        prob.exchange_users()
        ev_u = main.record_event()
        with torch.cuda.stream(cs):
            cs.wait_event(ev_u)
            out_u.copy_(prob.U[ulo:uhi, :k], non_blocking=True)
        main.wait_event(ev_item)
        prob.item_half_step(w["lam"], prec, args.ir_steps)
        prob.exchange_items()
        ev_v = main.record_event()
        with torch.cuda.stream(cs):
            cs.wait_event(ev_v)
            out_v.copy_(prob.V[ilo:ihi, :k], non_blocking=True)
        prob.recommend(n, idx_dev, args.score_mode)
        out_idx.copy_(idx_dev, non_blocking=True)
        main.wait_stream(cs)
        barrier()
        if i > 0:
            times.append(time.perf_counter() - t0)
    # the overlapped step must produce what the plain sequence (everything on one stream) produces, bit for bit
    barrier()
    prob.up.copy_(hp["up"]); prob.ux.copy_(hp["ux"]); prob.ip.copy_(hp["ip"]); prob.ix.copy_(hp["ix"])
    prob.U[ulo:uhi, :k].copy_(hu[ulo:uhi]); prob.V[:, :k].copy_(hv)
    barrier()
    prob.user_half_step(w["lam"], prec, args.ir_steps)
    prob.exchange_users()
    prob.item_half_step(w["lam"], prec, args.ir_steps)
    prob.exchange_items()
    prob.recommend(n, idx_dev, args.score_mode)
    barrier()
    same = (torch.equal(out_u, prob.U[ulo:uhi, :k].cpu()) and torch.equal(out_v, prob.V[ilo:ihi, :k].cpu())
            and torch.equal(out_idx, idx_dev.cpu()))
    ok = torch.tensor([1.0 if same else 0.0], dtype=torch.float64, device=device)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if ok.item() != 1.0:
        raise SystemExit("bench.py: the overlapped end-to-end step and the single-stream step disagree")
    sec = torch.tensor([float(np.mean(times))], dtype=torch.float64, device=device)
    dist.all_reduce(sec, op=dist.ReduceOp.MAX)
    sec = float(sec.item())
    tot = torch.tensor([float(h2d), float(d2h)], dtype=torch.float64, device=device)
    dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    return {"value": 1.0 / sec, "unit": UNIT, "h2d_bytes_per_step": int(tot[0].item()), "d2h_bytes_per_step": int(tot[1].item()),
            "ms_per_step": sec * 1e3, "steps": steps, "checked": "equal to the single-stream step (U, V ranges and top-n rows, bitwise)",
            "state": "every step starts from the factors of the last device-resident step",
            "api": "per-rank pinned H2D of its CSR/CSC ranges + factors on a copy stream (user half-step in 4 row ranges behind it), rl_als_half_step(_peers)_dev/rl_score_topn_dev + exchanges, D2H of its ranges"}


_REAL_STDOUT = None


def _quiet_stdout():
    """Everything a library prints on fd 1 (NCCL's version banner, ...) goes to stderr: stdout carries the one JSON line."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def _emit(line):
    sys.stdout.flush()
    if _REAL_STDOUT is not None:
        os.dup2(_REAL_STDOUT, 1)
    print(json.dumps(line), flush=True)


def _pinned_copy(a):
    from recsys_lite_b200 import _ffi
    out = _ffi.pinned_empty(a.shape, a.dtype)
    out[...] = a
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C4", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default="fp32_ir", choices=["fp32_ir", "fp64"])
    ap.add_argument("--ir-steps", type=int, default=1)
    ap.add_argument("--score-mode", type=int, default=0, help="0 auto, 1 exact CUDA-core, 2 tensor-core candidate pass")
    ap.add_argument("--no-peer-stores", action="store_true", help="N > 1: exchange with NCCL broadcasts instead of the fused NVLink stores")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--emulate-world", type=int, default=0, help="sharded workloads on one GPU: play one rank of a job of this size")
    ap.add_argument("--emulate-rank", type=int, default=0)
    ap.add_argument("--no-api", action="store_true", help="skip the e2e_api leg (the step through RecommenderSystem.fit / .recommend)")
    ap.add_argument("--no-parity", action="store_true", help="skip the sampled float64 parity check of the final state")
    args = ap.parse_args()
    _quiet_stdout()
    if args.warmup < 3 and args.impl == "b200" and not os.environ.get("RL_BENCH_ALLOW_SHORT_WARMUP"):
        args.warmup = 3
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, w)
    elif w.get("sharded"):
        run_sharded_als(args, w)
    else:
        run_b200(args, w)


if __name__ == "__main__":
    main()

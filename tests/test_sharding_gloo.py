"""World-size-2 CPU tests (gloo) of the multi-GPU plumbing: the nnz-balanced plan, the in-place row exchange and
the sharded iteration schedule (user half-step -> exchange U -> item half-step -> exchange V).  The kernels are
replaced by the CPU oracle here, so what is checked is exactly the host logic that bench.py / ShardedProblem use
on the GPUs: a sharded iteration must reproduce the unsharded one bit for bit."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

from oracle import als as o_als, generators as gen  # noqa: E402
from recsys_lite_b200 import sharding  # noqa: E402
from recsys_lite_b200.csr import shard_rows  # noqa: E402


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n_users, n_items, k = 400, 150, 8
        indptr, indices = gen.synth_pattern(n_users, n_items, mean_deg=7, seed=5, skew=True)
        ti, tx = o_als.transpose_pattern(indptr, indices, n_items)
        u0, v0 = gen.seeded_factors(n_users, n_items, k, 2)
        plan = sharding.make_plan(indptr, ti, rank, world)
        ulo, uhi = plan.users
        ilo, ihi = plan.items
        up, ux = shard_rows(indptr, indices, ulo, uhi)
        ip, ix = shard_rows(ti, tx, ilo, ihi)
        U, V = torch.from_numpy(u0.copy()), torch.from_numpy(v0.copy())
        for _ in range(2):
            o_als.half_step_csr(up, ux, V.numpy(), U.numpy()[ulo:uhi], 0.01)       # this rank's users only
            if rank == 1:
                U[:ulo] = -7.0                                                      # stale rows must be overwritten
            sharding.exchange_rows(U, plan.user_bounds, rank, world)
            o_als.half_step_csr(ip, ix, U.numpy(), V.numpy()[ilo:ihi], 0.01)
            sharding.exchange_rows(V, plan.item_bounds, rank, world)
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), U=U.numpy(), V=V.numpy(),
                 ub=plan.user_bounds, ib=plan.item_bounds)
    finally:
        dist.destroy_process_group()


def test_sharded_iteration_matches_unsharded(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    n_users, n_items, k = 400, 150, 8
    indptr, indices = gen.synth_pattern(n_users, n_items, mean_deg=7, seed=5, skew=True)
    ti, tx = o_als.transpose_pattern(indptr, indices, n_items)
    u0, v0 = gen.seeded_factors(n_users, n_items, k, 2)
    ref = o_als.fit_csr(indptr, indices, ti, tx, k, iterations=2, lam=0.01, init=(u0, v0))
    outs = [np.load(os.path.join(tmp_path, f"rank{r}.npz")) for r in range(world)]
    for o in outs:
        assert np.array_equal(o["U"], ref["user_factors"]) and np.array_equal(o["V"], ref["item_factors"])
        assert o["ub"][0] == 0 and o["ub"][-1] == n_users and o["ib"][-1] == n_items
    # the plan balances nnz, not rows (skewed degrees): the two user ranges differ in length but not in work
    ub = outs[0]["ub"]
    nnz = np.diff(indptr[ub])
    assert abs(int(nnz[0]) - int(nnz[1])) <= 2 * int(np.diff(indptr).max())


def test_plan_covers_rows_for_every_world_size():
    indptr, indices = gen.synth_pattern(1000, 300, mean_deg=9, seed=1, skew=True)
    ti, _ = o_als.transpose_pattern(indptr, indices, 300)
    for world in (1, 2, 3, 4, 8):
        plans = [sharding.make_plan(indptr, ti, r, world) for r in range(world)]
        covered = np.concatenate([np.arange(*p.users) for p in plans])
        assert np.array_equal(covered, np.arange(1000))
        covered = np.concatenate([np.arange(*p.items) for p in plans])
        assert np.array_equal(covered, np.arange(300))
        per = np.array([indptr[p.users[1]] - indptr[p.users[0]] for p in plans])
        assert per.max() - per.min() <= 2 * np.diff(indptr).max() + 1

"""Multi-GPU plumbing for the ALS iteration and the scoring pass: one process per GPU.

The reference has no parallelism (plain Python loops, algo.py:314,324).  The loops are independent per row, so:

  * user half-step: users are cut into `world` contiguous ranges balanced by nnz; rank g holds the CSR rows of
    its users and a full replica of V, solves its range and writes U[g];
  * the freshly solved range is exchanged once per half-step so every rank has the full U again before the item
    half-step, and likewise V afterwards.  On NVLink-connected GPUs the exchange is FUSED into the solve kernel
    (`enable_peer_stores`): every rank maps the other ranks' replicas (CUDA IPC) and the kernel stores each solved
    row into all of them, so the transfer rides under the solves and what is left of the "all-gather" is one
    one-element all-reduce as a barrier.  `exchange_rows` (NCCL broadcasts of the row ranges) is the portable path;
  * scoring shards users with V replicated: no collective.

torch is used here only for device memory, the stream and torch.distributed; all arithmetic goes through the
C ABI (librecsys_b200.so) on raw data_ptr()s.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .csr import nnz_balanced_bounds, shard_rows


@dataclass
class ShardPlan:
    """Host-side description of what one rank owns."""
    rank: int
    world: int
    user_bounds: np.ndarray   # (world+1,) row boundaries of the user ranges
    item_bounds: np.ndarray   # (world+1,)

    @property
    def users(self):
        return int(self.user_bounds[self.rank]), int(self.user_bounds[self.rank + 1])

    @property
    def items(self):
        return int(self.item_bounds[self.rank]), int(self.item_bounds[self.rank + 1])


def make_plan(indptr_u, indptr_i, rank: int, world: int) -> ShardPlan:
    return ShardPlan(rank, world, nnz_balanced_bounds(indptr_u, world), nnz_balanced_bounds(indptr_i, world))


def exchange_rows(full, bounds, rank: int, world: int, group=None):
    """All-gather of contiguous row ranges of `full` (a torch tensor, rows x ld) IN PLACE.

    Rank r contributes rows [bounds[r], bounds[r+1]); afterwards every rank holds all rows.  Ranges may have
    different lengths (nnz-balanced), so this is issued as `world` broadcasts of views of `full` inside one
    coalesced group where the backend supports it -- no staging copy, no padding.
    """
    if world == 1:
        return
    import torch.distributed as dist
    views = [full[int(bounds[r]):int(bounds[r + 1])] for r in range(world)]
    works = []
    for r, v in enumerate(views):
        if v.numel():
            works.append(dist.broadcast(v, src=dist.get_global_rank(group, r) if group is not None else r,
                                        group=group, async_op=True))
    for w in works:
        w.wait()


class ShardedProblem:
    """Device-resident ALS problem of one rank: its CSR / CSC row ranges and full factor replicas."""

    def __init__(self, handle, plan: ShardPlan, indptr_u, indices_u, indptr_i, indices_i, k: int, device, *, shards=None,
                 n_users=None, n_items=None):
        """indptr_* / indices_*: the GLOBAL CSR and transposed CSR (each rank cuts its ranges out), or -- problems too large to
        build whole on every rank -- shards=(up, ux, ip, ix): this rank's rows of both patterns with local indptr (indices stay
        global ids), together with n_users / n_items."""
        import torch
        self.h, self.plan, self.k = handle, plan, int(k)
        self.kp = handle.lib.rl_padded_k(self.k)
        self.device = device
        ulo, uhi = plan.users
        ilo, ihi = plan.items
        if shards is not None:
            up, ux, ip, ix = shards
            self.n_users, self.n_items = int(n_users), int(n_items)
            assert len(up) == uhi - ulo + 1 and len(ip) == ihi - ilo + 1
        else:
            self.n_users, self.n_items = len(indptr_u) - 1, len(indptr_i) - 1
            up, ux = shard_rows(indptr_u, indices_u, ulo, uhi)
            ip, ix = shard_rows(indptr_i, indices_i, ilo, ihi)
        self.host = {"up": up, "ux": ux, "ip": ip, "ix": ix}   # kept for the end-to-end (H2D inside) leg
        self.nnz_users, self.nnz_items = len(ux), len(ix)
        t = lambda a: torch.from_numpy(a).to(device)           # noqa: E731
        self.up, self.ux, self.ip, self.ix = t(up), t(ux), t(ip), t(ix)
        self.U = torch.zeros((self.n_users, self.kp), dtype=torch.float32, device=device)
        self.V = torch.zeros((self.n_items, self.kp), dtype=torch.float32, device=device)
        self._bound_stream = None
        self._bind_stream()

    def _bind_stream(self):
        """The handle's kernels (and their NVLink peer stores), torch's copies and the NCCL barrier must be ordered on ONE
        stream: the handle is (re)bound to torch's current stream before anything is enqueued.  (A default Handle owns a
        private non-blocking stream: with it the one-element all-reduce would not wait for the solve kernel, and peers
        would read half-written replicas.)"""
        import torch
        if getattr(self.device, "type", str(self.device)) != "cuda" and "cuda" not in str(self.device):
            return
        cur = torch.cuda.current_stream(self.device).cuda_stream
        if cur != self._bound_stream:
            self.h.set_stream(cur)
            self._bound_stream = cur

    # -- fused exchange over peer memory ---------------------------------------------------------------------
    def enable_peer_stores(self, group=None):
        """Map every other rank's U / V replica into this process (rl_ipc_export / rl_ipc_open, all_gather_object for
        the handles).  Afterwards the half-steps store their rows into all replicas (rl_als_half_step_peers_dev) and
        `iteration` only needs `peer_barrier` between them.  Returns False (and changes nothing) if any rank could
        not export or open a mapping; collective: every rank must call it."""
        import ctypes as C
        import torch
        import torch.distributed as dist
        from . import _ffi
        world, rank = self.plan.world, self.plan.rank
        if world == 1:
            return False
        if world - 1 > 15:
            return False
        ok, mine = True, None
        try:
            mine = []
            for t in (self.U, self.V):
                hb = C.create_string_buffer(64)
                off = C.c_int64(0)
                self.h.call("rl_ipc_export", t.data_ptr(), hb, C.byref(off))
                mine.append((hb.raw, int(off.value)))
        except _ffi.RecsysB200Error:
            ok = False
        gathered = [None] * world
        dist.all_gather_object(gathered, mine if ok else None, group=group)
        if any(g is None for g in gathered):
            return False
        peers_u, peers_v, opened = [], [], []
        try:
            for r in range(world):
                if r == rank:
                    continue
                for (raw, off), lst in zip(gathered[r], (peers_u, peers_v)):
                    out = C.c_void_p()
                    self.h.call("rl_ipc_open", C.create_string_buffer(raw, 64), off, C.byref(out))
                    opened.append(out.value)
                    lst.append(out.value)
        except _ffi.RecsysB200Error:
            ok = False
        flags = [None] * world
        dist.all_gather_object(flags, ok, group=group)
        if not all(flags):
            for ptr in opened:
                self.h.call("rl_ipc_close", ptr)
            return False
        self._peer_ptrs = opened
        ulo, _ = self.plan.users
        ilo, _ = self.plan.items
        row_bytes = 4 * self.kp
        arr = C.c_void_p * (world - 1)
        self._peers_u = arr(*[p + ulo * row_bytes for p in peers_u])     # where MY user range lives in each peer's U
        self._peers_v = arr(*[p + ilo * row_bytes for p in peers_v])
        self._barrier_buf = torch.zeros(1, dtype=torch.float32, device=self.device)
        self._group = group
        self.peer_stores = True
        return True

    def peer_barrier(self):
        """Every rank's kernel (and with it its NVLink stores into this replica) has finished once this one-element
        all-reduce, enqueued behind the kernel on the same stream, completes."""
        import torch.distributed as dist
        self._bind_stream()
        dist.all_reduce(self._barrier_buf, group=self._group)

    def close_peer_stores(self):
        if getattr(self, "peer_stores", False):
            self.h.synchronize()
            for ptr in self._peer_ptrs:
                self.h.call("rl_ipc_close", ptr)
            self.peer_stores = False

    def set_factor_rows(self, u_rows, v_host):
        """Initial factors of a sharded start: this rank's user rows and all of V (every other user row is written by its
        owner's first half-step and arrives through the exchange before anybody reads it)."""
        import torch
        self._bind_stream()
        ulo, uhi = self.plan.users
        self.U[ulo:uhi, : self.k].copy_(torch.from_numpy(np.ascontiguousarray(u_rows, dtype=np.float32)))
        self.V[:, : self.k].copy_(torch.from_numpy(np.ascontiguousarray(v_host, dtype=np.float32)))

    def set_factors(self, u_host, v_host):
        import torch
        self._bind_stream()
        self.U[:, : self.k].copy_(torch.from_numpy(np.ascontiguousarray(u_host, dtype=np.float32)))
        self.V[:, : self.k].copy_(torch.from_numpy(np.ascontiguousarray(v_host, dtype=np.float32)))

    # -- one ALS half-step on this rank's rows, then the exchange -------------------------------------------
    def user_half_step(self, lam, precision, ir_steps, rows=None):
        """rows = (r0, r1): only that part of this rank's user range (callers that stream the pattern in, part by part)."""
        ulo, uhi = self.plan.users
        r0, r1 = (0, uhi - ulo) if rows is None else rows
        if r1 <= r0:
            return
        self._bind_stream()
        if getattr(self, "peer_stores", False):
            import ctypes as C
            peers = self._peers_u
            if r0:
                peers = (C.c_void_p * (self.plan.world - 1))(*[p + r0 * 4 * self.kp for p in self._peers_u])
            self.h.call("rl_als_half_step_peers_dev", r1 - r0, self.n_items, self.k, self.up[r0:].data_ptr(), self.ux.data_ptr(),
                        None, None, self.V.data_ptr(), self.U[ulo + r0:].data_ptr(), float(lam), 0.0, None, precision,
                        ir_steps, None, self.plan.world - 1, peers)
        else:
            self.h.call("rl_als_half_step_dev", r1 - r0, self.n_items, self.k, self.up[r0:].data_ptr(), self.ux.data_ptr(),
                        None, None, self.V.data_ptr(), self.U[ulo + r0:].data_ptr(), float(lam), 0.0, None, precision,
                        ir_steps, None)

    def item_half_step(self, lam, precision, ir_steps):
        ilo, ihi = self.plan.items
        self._bind_stream()
        if getattr(self, "peer_stores", False):
            self.h.call("rl_als_half_step_peers_dev", ihi - ilo, self.n_users, self.k, self.ip.data_ptr(), self.ix.data_ptr(),
                        None, None, self.U.data_ptr(), self.V[ilo:].data_ptr(), float(lam), 0.0, None, precision,
                        ir_steps, None, self.plan.world - 1, self._peers_v)
        else:
            self.h.call("rl_als_half_step_dev", ihi - ilo, self.n_users, self.k, self.ip.data_ptr(), self.ix.data_ptr(),
                        None, None, self.U.data_ptr(), self.V[ilo:].data_ptr(), float(lam), 0.0, None, precision,
                        ir_steps, None)

    def exchange_users(self):
        """Make U whole on every rank after the user half-step (barrier only when the kernel already stored to the peers)."""
        if getattr(self, "peer_stores", False):
            self.peer_barrier()
        else:
            exchange_rows(self.U, self.plan.user_bounds, self.plan.rank, self.plan.world)

    def exchange_items(self):
        if getattr(self, "peer_stores", False):
            self.peer_barrier()
        else:
            exchange_rows(self.V, self.plan.item_bounds, self.plan.rank, self.plan.world)

    def iteration(self, lam=0.01, precision=0, ir_steps=1):
        self.user_half_step(lam, precision, ir_steps)
        self.exchange_users()
        self.item_half_step(lam, precision, ir_steps)
        self.exchange_items()

    def recommend(self, n, idx_out, mode=0):
        """top-n for this rank's users (seen items = its CSR rows) into idx_out (users_local x n int32 tensor)."""
        ulo, uhi = self.plan.users
        self._bind_stream()
        self.h.call("rl_score_topn_dev", self.U[ulo:].data_ptr(), uhi - ulo, self.V.data_ptr(), self.n_items, self.k,
                    self.up.data_ptr(), self.ux.data_ptr(), 0.0, None, None, int(n), idx_out.data_ptr(), None, mode)

"""Point-sharded evaluation over the GPUs of one node.

Host-map points are independent units; the only coupling between shards is the sum into the packed vector
[cost, g, triu(H)] (and three scalars of the outlier statistics).  Every rank owns a contiguous block of the points of
every pyramid level; images / spline / pattern are replicated.

Two ways to make the result global:
  * fused (default, `connect_shards`): the tracking kernel's last block exchanges the packed vector with all ranks
    through peer-mapped mailboxes over NVLink and sums them in rank order — one-shot all-reduce inside the kernel.
    torch.distributed only carries the 64-byte CUDA IPC handles once, at set-up.
  * NCCL (`ShardedEvaluator`, baseline): kernel -> ncclAllReduce of the packed vector (728 B for 2 knots) on the same
    stream -> D2H.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import numpy as np


def shard_bounds(num_points: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of `num_points` owned by `rank`: sizes differ by at most one, whole points only."""
    base, rem = divmod(num_points, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def reduce_packed(local: np.ndarray, all_reduce_sum: Callable[[np.ndarray], np.ndarray]) -> np.ndarray:
    """Sum of the per-shard packed vectors.  Each shard is already scaled by 1 / num_residuals_GLOBAL, so the plain
    sum is the global [cost, g, triu(H)]."""
    return all_reduce_sum(np.ascontiguousarray(local, dtype=np.float64))


def connect_shards(ctx, prob, rank: int, world: int, all_gather_bytes: Callable[[bytes], list]):
    """Upload this rank's shard of every level of `prob` into `ctx` and connect the fused all-reduce.
    all_gather_bytes(b) -> [b_0, ..., b_{world-1}] exchanges one bytes object per rank (any transport).
    Afterwards ctx.evaluate / gn_iteration / optimize_level / detect_outliers are collective calls."""
    ctx.set_frame_times(prob.cap, prob.exp)
    for l, lv in enumerate(prob.levels):
        lo, hi = shard_bounds(lv.P, rank, world)
        ctx.set_level(l, lv, slice(lo, hi))
    handle, _ = ctx.shard_export()
    handles = all_gather_bytes(handle)
    ctx.shard_connect(world, rank, handles=handles)
    for l, lv in enumerate(prob.levels):
        ctx.shard_set_global_points(l, lv.P)


class ShardedEvaluator:
    """One rank's view of a point-sharded tracker: a Context holding this rank's shard of every level."""

    def __init__(self, ctx, prob, rank: int, world: int, device_index: Optional[int] = None):
        import torch

        self.torch = torch
        self.ctx, self.prob, self.rank, self.world = ctx, prob, rank, world
        self.device = torch.device("cuda", torch.cuda.current_device() if device_index is None else device_index)
        self.bounds = [shard_bounds(lv.P, rank, world) for lv in prob.levels]
        ctx.set_frame_times(prob.cap, prob.exp)
        for l, lv in enumerate(prob.levels):
            lo, hi = self.bounds[l]
            ctx.set_level(l, lv, slice(lo, hi))
        self.packed = torch.zeros(ctx.packed_len(8), dtype=torch.float64, device=self.device)
        self.stream = torch.cuda.current_stream(self.device)
        ctx.set_stream(self.stream.cuda_stream)

    def evaluate(self, level: int, knots_t, knots_R, with_hessian: bool = True, num_bad_global: int = 0):
        """Global (cost, H, g) — identical on every rank."""
        import torch.distributed as dist

        prob, lv = self.prob, self.prob.levels[level]
        nres = (lv.P - num_bad_global) * prob.F * lv.S
        kmin, nk = self.ctx.evaluate_async(level, prob.k, prob.t0, prob.dt, knots_t, knots_R, prob.huber_a, with_hessian,
                                           nres, self.packed.data_ptr())
        E = self.ctx.packed_len(nk) if with_hessian else 1
        buf = self.packed[:E]
        if self.world > 1:
            dist.all_reduce(buf, op=dist.ReduceOp.SUM)
        host = buf.cpu().numpy()
        return self.ctx.unpack(host, kmin, nk, prob.n_knots, with_hessian)

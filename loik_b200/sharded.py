"""Batch-sharded multi-GPU driver.

Problem instances are independent, so each rank (one process per GPU, torchrun) owns a contiguous slice of
the global batch and runs the same kernels on it; no tensor ever crosses NVLink.  The only collective is the
all-reduce (SUM) of the still-active instance count that decides the *global* stop: one int32 per chunk of
ADMM iterations.  Converged instances are frozen on device, so running a few extra sweeps past a rank's own
convergence never changes a result.
"""
from __future__ import annotations

import torch


class _DevInt:
    """Zero-copy view of the library's device-resident active counter."""

    def __init__(self, ptr: int):
        self.__cuda_array_interface__ = {"shape": (1,), "typestr": "<i4", "data": (ptr, False), "version": 2}


def shard_range(global_batch: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous slice [lo, hi) of the global batch owned by `rank` (remainder spread over the first ranks)."""
    base, rem = divmod(global_batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def global_stop(active_local: torch.Tensor, world: int) -> int:
    """SUM all-reduce of the active count; returns the global number of still-active instances."""
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(active_local, op=dist.ReduceOp.SUM)
    return int(active_local.item())


class ShardedSolver:
    def __init__(self, solver, world: int, chunk: int = 4):
        self.S, self.world, self.chunk = solver, world, chunk
        self._active = None

    def solve(self) -> int:
        """Solve() on every rank's shard; returns the number of ADMM sweeps launched."""
        S = self.S
        if self.world == 1:
            S.Solve()
            return 0
        if self._active is None:
            self._active = torch.as_tensor(_DevInt(S.active_count_ptr()), device=f"cuda:{S.device}")
        S.SolveBegin()
        done = 0
        limit = self.max_iter
        while done < limit:
            k = min(self.chunk, limit - done)
            S.SolveChunk(k)
            done += k
            if global_stop(self._active, self.world) == 0:
                break
        return done

    @property
    def max_iter(self) -> int:
        return int(self.S.max_iter)

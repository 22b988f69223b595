"""Batch-sharded multi-GPU driver.

Problem instances are independent, so each rank (one process per GPU, torchrun) owns a contiguous slice of
the global batch and runs the same kernels on it; no tensor ever crosses NVLink.  Loop control is per
instance and lives on the device (converged instances are frozen), so a rank never has to wait for another
rank's instances.  The only collective is ONE all-reduce (SUM) per solve of four int64 -- the global stopping-
criterion outcome {#converged, #primal infeasible, #stopped at max_iter, total iterations} -- which every rank
needs to report the same global status.  A chunked variant (`solve_chunked`) that all-reduces the still-active
count every few sweeps and stops all ranks as soon as the global count reaches zero is kept for callers that
want the early global exit.
"""
from __future__ import annotations

import torch


class _DevArray:
    """Zero-copy view of a device buffer owned by the library."""

    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def shard_range(global_batch: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous slice [lo, hi) of the global batch owned by `rank` (remainder spread over the first ranks)."""
    base, rem = divmod(global_batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_reduce_sum(t: torch.Tensor, world: int) -> torch.Tensor:
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


class ShardedSolver:
    def __init__(self, solver, world: int, chunk: int = 8):
        self.S, self.world, self.chunk = solver, world, chunk
        self._active = None
        self._stats = None

    def solve(self) -> torch.Tensor:
        """Solve() on this rank's shard (asynchronous) + the single all-reduce of the global outcome.
        Returns a device tensor of 4 int64 (global counts); reading it synchronizes."""
        S = self.S
        S.Solve()
        ptr = S.reduce_stats_ptr()
        if self._stats is None:
            self._stats = torch.as_tensor(_DevArray(ptr, 4, "<i8"), device=f"cuda:{S.device}")
        return all_reduce_sum(self._stats, self.world)

    def solve_chunked(self) -> int:
        """Chunks of ADMM sweeps interleaved with the all-reduce of the active count; global early exit."""
        S = self.S
        if self._active is None:
            self._active = torch.as_tensor(_DevArray(S.active_count_ptr(), 1, "<i4"), device=f"cuda:{S.device}")
        S.SolveBegin()
        done, limit = 0, int(S.max_iter)
        while done < limit:
            k = min(self.chunk, limit - done)
            S.SolveChunk(k)
            done += k
            if int(all_reduce_sum(self._active, self.world).item()) == 0:
                break
        return done

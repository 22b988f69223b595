"""Batch-sharded multi-GPU driver.

Problem instances are independent, so each rank (one process per GPU, torchrun) owns a contiguous slice of
the global batch and runs the same kernels on it; no tensor ever crosses NVLink.  Loop control is per
instance and lives on the device (converged instances are frozen), so a rank never has to wait for another
rank's instances.  The only collective is ONE all-reduce (SUM) per solve of four int64 -- the global stopping-
criterion outcome {#converged, #primal infeasible, #stopped at max_iter, total iterations} -- which every rank
needs to report the same global status.  A chunked variant (`solve_chunked`) all-reduces the still-active
count every few sweeps and stops all ranks as soon as the global count reaches zero: the literal "all-reduce the
residual of the global stopping criterion" form of BASELINE.json's north_star.

The solver object only has to provide ``Solve() / stats_tensor()`` and ``SolveBegin() / SolveChunk(k) /
active_tensor() / max_iter`` (``loik_b200.solver.FirstOrderLoikOptimized`` does; the world-size-2 gloo test drives this
file's control flow with a CPU stand-in).
"""
from __future__ import annotations

import torch


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

    def solve(self) -> torch.Tensor:
        """Solve() on this rank's shard (asynchronous) + the single all-reduce of the global outcome.
        Returns a tensor of 4 int64 (global counts) owned by the caller; reading it synchronizes."""
        S = self.S
        S.Solve()
        # the library's buffer is rewritten by the next solve / stats call: reduce a private copy (same stream)
        return all_reduce_sum(S.stats_tensor().clone(), self.world)

    def solve_chunked(self) -> int:
        """Chunks of ADMM sweeps interleaved with the all-reduce of the active count; every rank leaves the loop in the
        same round, the first one after which no instance is active anywhere.  Returns the sweeps run."""
        S = self.S
        S.SolveBegin()
        done, limit = 0, int(S.max_iter)
        while done < limit:
            k = min(self.chunk, limit - done)
            S.SolveChunk(k)
            done += k
            if int(all_reduce_sum(S.active_tensor().clone(), self.world).item()) == 0:
                break
        return done

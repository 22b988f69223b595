"""Host-side logic of the batch-sharded multi-GPU mode, world_size 2 over gloo on CPU."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from loik_b200 import problems, robots, sharded


def test_shard_ranges_cover_the_batch():
    for B in (1, 7, 64, 131072, 131073):
        for W in (1, 2, 3, 8):
            spans = [sharded.shard_range(B, r, W) for r in range(W)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(W - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_shards_are_bit_identical_to_the_whole():
    """Counter-based generation: rank r's shard == rows [lo, hi) of the global batch, for any world size."""
    model = robots.talos()
    B = 10000
    whole = problems.random_batch(model, B, seed=0)
    for W in (2, 8):
        for r in range(W):
            lo, hi = sharded.shard_range(B, r, W)
            part = problems.random_batch(model, hi - lo, seed=0, first_index=lo)
            np.testing.assert_array_equal(part["q"], whole["q"][lo:hi])
            np.testing.assert_array_equal(part["bis"], whole["bis"][lo:hi])


class _FakeSolver:
    """CPU stand-in for one rank's solver: a scripted sequence of still-active counts per chunk, and per-rank stats."""

    def __init__(self, active_per_chunk, stats, max_iter=200):
        self.script, self.stats, self.max_iter = list(active_per_chunk), stats, max_iter
        self.chunks, self.solves = [], 0

    def Solve(self):
        self.solves += 1

    def stats_tensor(self):
        return torch.tensor(self.stats, dtype=torch.int64)

    def SolveBegin(self):
        self.chunks = []

    def SolveChunk(self, k):
        self.chunks.append(k)

    def active_tensor(self):
        i = len(self.chunks) - 1
        return torch.tensor([self.script[i] if i < len(self.script) else 0], dtype=torch.int32)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # ShardedSolver's own control flow over gloo.  solve(): the global stopping-criterion outcome, 4 int64 per rank, SUM
    fake = _FakeSolver([[5, 3, 0, 0, 0], [2, 2, 2, 0, 0]][rank], [10 + rank, rank, 1, 100 * (rank + 1)])
    drv = sharded.ShardedSolver(fake, world, chunk=8)
    total = drv.solve()
    # the library buffer (here: the fake's tensor) must not be reduced in place: a second solve reports the same totals
    total2 = drv.solve()
    # solve_chunked(): everybody stops in the same round, only when the global active count is zero
    sweeps = drv.solve_chunked()
    q.put((rank, total.tolist(), total2.tolist(), sweeps, list(fake.chunks)))
    dist.destroy_process_group()


def test_global_stop_all_reduce_gloo():
    world, port = 2, 29531 + os.getpid() % 500
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, total, total2, sweeps, chunks in res:
        assert total == [21, 1, 2, 300] and total2 == total
        assert sweeps == 32 and chunks == [8, 8, 8, 8]  # rank 0 is idle after 3 chunks but leaves with rank 1 after the 4th


def test_chunked_budget_is_max_iter():
    """No global convergence: the chunks add up to max_iter exactly (the last one is shortened)."""
    fake = _FakeSolver([1] * 100, [0, 0, 0, 0], max_iter=20)
    assert sharded.ShardedSolver(fake, 1, chunk=8).solve_chunked() == 20
    assert fake.chunks == [8, 8, 4]

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


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # the global stopping-criterion outcome: 4 int64 per rank, SUM all-reduce (sharded.ShardedSolver.solve)
    local = torch.tensor([10 + rank, rank, 1, 100 * (rank + 1)], dtype=torch.int64)
    out = sharded.all_reduce_sum(local.clone(), world)
    # the chunked stop rule: everybody stops in the same round, only when the global active count is zero
    active = [[5, 3, 0, 0], [2, 2, 2, 0]][rank]
    rounds = 0
    for a in active:
        rounds += 1
        if int(sharded.all_reduce_sum(torch.tensor([a], dtype=torch.int32), world).item()) == 0:
            break
    q.put((rank, out.tolist(), rounds))
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
    for rank, total, rounds in res:
        assert total == [21, 1, 2, 300]
        assert rounds == 4

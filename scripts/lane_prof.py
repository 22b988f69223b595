#!/usr/bin/env python
"""Development aid: a short run for ncu captures of one iteration kernel.

  python scripts/lane_prof.py <robot> <lane_after> [iters]     # fixed-iteration launches, then one Solve()
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from loik_b200 import problems, robots, solver as lk  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "panda"
la = int(sys.argv[2]) if len(sys.argv) > 2 else 0
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 4
B = int(os.environ.get("BATCH", {"panda": 65536, "ur10": 262144, "talos": 16384}.get(name, 16384)))
model = robots.get_robot(name)
pb = problems.random_batch(model, B, seed=0)
S = lk.make_solver(model, problems.bench_params(len(pb["ids"])), B)
S.set_schedule(lane_after=la, lane_groups_per_instance=int(os.environ.get("GPI", "0")))
S.SolveInit(pb["q"], pb["H_ref"], pb["v_ref"], pb["ids"], pb["Ais"], pb["bis"], pb["lb"], pb["ub"])
for _ in range(3):
    S.IterateFixed(iters)
torch.cuda.synchronize()
S.Solve()
torch.cuda.synchronize()
print(S.stats())
S.close()

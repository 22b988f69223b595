#!/usr/bin/env python
"""One warm-up solve + one solve of a BASELINE batch (for `ncu --metrics gpu__time_duration.sum` launch lists)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from loik_b200 import problems, robots, solver as lk  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "panda"
B = int(os.environ.get("BATCH", {"panda": 65536, "ur10": 262144, "talos": 16384, "talos_ff": 16384}[name]))
model = robots.get_robot(name)
pb = problems.random_batch(model, B, seed=0)
S = lk.make_solver(model, problems.bench_params(len(pb["ids"])), B)
S.SolveInit(pb["q"], pb["H_ref"], pb["v_ref"], pb["ids"], pb["Ais"], pb["bis"], pb["lb"], pb["ub"])
S.Solve()
torch.cuda.synchronize()
print("MARK second solve starts after", S.launch_count(), "launches")
S.Solve()
torch.cuda.synchronize()
print("total launches", S.launch_count())

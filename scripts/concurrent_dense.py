#!/usr/bin/env python
"""Development aid: aggregate rate of fixed-iteration dense launches issued on 1 / 2 / 4 / 8 streams (independent solver handles) --
does the GPU overlap the tile kernels of different solves?   python scripts/concurrent_dense.py talos"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
from loik_b200 import problems, robots, solver as lk  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "talos"
B = {"panda": 65536, "ur10": 262144, "talos": 16384}[name]
model = robots.get_robot(name)
pb = problems.random_batch(model, B, seed=0)
params = problems.bench_params(len(pb["ids"]))
N = 8
Ss = [lk.make_solver(model, params, B) for _ in range(N)]
streams = [torch.cuda.Stream() for _ in range(N)]
for S in Ss:
    S.set_schedule(lane_after=-1)
    S.SolveInit(pb["q"], pb["H_ref"], pb["v_ref"], pb["ids"], pb["Ais"], pb["bis"], pb["lb"], pb["ub"])
    S.IterateFixed(2)
torch.cuda.synchronize()
K = 20
for n in (1, 2, 4, 8):
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(n):
            with torch.cuda.stream(streams[i]):
                Ss[i].IterateFixed(K, reset=False)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    print(f"{name}: {n} streams x {K} dense iterations: {dt * 1e6 / K:.1f} us per round of {n} launches = {n * B * K / dt / 1e6:.0f} M instance-iterations/s", flush=True)

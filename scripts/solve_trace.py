#!/usr/bin/env python
"""Per-launch device timeline of ONE unpipelined batched solve (torch.profiler / CUPTI): development aid."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from loik_b200 import problems, robots, solver as lk  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "panda"
B = int(os.environ.get("BATCH", {"panda": 65536, "ur10": 262144, "talos": 16384, "talos_ff": 16384}[name]))
model = robots.get_robot(name)
pb = problems.random_batch(model, B, seed=0)
S = lk.make_solver(model, problems.bench_params(len(pb["ids"])), B)
stream = torch.cuda.Stream()
with torch.cuda.stream(stream):
    S.SolveInit(pb["q"], pb["H_ref"], pb["v_ref"], pb["ids"], pb["Ais"], pb["bis"], pb["lb"], pb["ub"])
    for _ in range(3):
        S.Solve()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    with torch.cuda.stream(stream):
        S.Solve()
    torch.cuda.synchronize()
ev = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA), key=lambda e: e.time_range.start)
t0 = ev[0].time_range.start
print(f"{name} B={B}: {len(ev)} device activities, {(ev[-1].time_range.end - t0) / 1e3:.3f} ms")
tot = {}
for e in ev:
    nm = e.name.split("<")[0].split("(")[0].replace("void loik::", "")
    d = e.time_range.end - e.time_range.start
    tot[nm] = tot.get(nm, 0.0) + d
    if not nm.startswith("Memset"):
        print(f"  +{(e.time_range.start - t0):9.1f} us  {d:8.1f} us  {nm}")
print({k: round(v, 1) for k, v in tot.items()})

#!/usr/bin/env python
"""Kernel timeline of the pipelined solve via torch.profiler (CUPTI): who runs when, how much overlaps."""
import collections
import os
import sys

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from loik_b200 import problems, robots, solver as lk  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "panda"
D = int(os.environ.get("DEPTH", 16))
B = int(os.environ.get("BATCH", {"panda": 65536, "ur10": 262144, "talos": 16384, "talos_ff": 16384}[name]))
model = robots.get_robot(name)
pb = problems.random_batch(model, B, seed=0)
P = problems.bench_params(len(pb["ids"]))
Ss = [lk.make_solver(model, P, B) for _ in range(D)]
st = [torch.cuda.Stream() for _ in range(D)]
for i, S in enumerate(Ss):
    with torch.cuda.stream(st[i]):
        S.SolveInit(pb["q"], pb["H_ref"], pb["v_ref"], pb["ids"], pb["Ais"], pb["bis"], pb["lb"], pb["ub"])
        S.Solve()
torch.cuda.synchronize()
for i in range(2 * D):
    with torch.cuda.stream(st[i % D]):
        Ss[i % D].Solve()
torch.cuda.synchronize()
n = 3 * D
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(n):
        with torch.cuda.stream(st[i % D]):
            Ss[i % D].Solve()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ks = [(e.name.split("<")[0].split("(")[0].replace("void loik::", ""), e.time_range.start, e.time_range.end) for e in ev]
t0 = min(k[1] for k in ks); t1 = max(k[2] for k in ks)
wall = (t1 - t0)
print(f"{name} depth {D}: {len(ks)} kernels/memsets over {wall/1e3:.2f} ms wall = {wall/1e3/n:.3f} ms/solve")
agg = collections.defaultdict(lambda: [0, 0.0])
for nm, a, b in ks:
    agg[nm][0] += 1; agg[nm][1] += b - a
for nm, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"  {nm:28s} n={c:5d} sum={t/1e3:9.2f} ms  ({t/1e3/n:.3f} ms/solve)  avg={t/c:8.1f} us")
# concurrency profile: sweep line
pts = []
for nm, a, b in ks:
    pts.append((a, 1)); pts.append((b, -1))
pts.sort()
cur = 0; last = t0; hist = collections.Counter()
for t, d in pts:
    hist[min(cur, 20)] += t - last
    last = t; cur += d
print("  concurrency (kernels in flight -> fraction of wall):", {k: round(v / wall, 3) for k, v in sorted(hist.items())})
# k_iterate duration buckets
buck = collections.defaultdict(lambda: [0, 0.0])
for nm, a, b in ks:
    if nm.startswith("k_iterate"):
        d = b - a
        key = "<60us" if d < 60 else "<150us" if d < 150 else "<400us" if d < 400 else "<1ms" if d < 1000 else ">=1ms"
        buck[key][0] += 1; buck[key][1] += d
print("  k_iterate by duration:", {k: (v[0], round(v[1] / 1e3 / n, 3)) for k, v in buck.items()})

import os, sys, time
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from loik_b200 import problems, robots, solver as lk
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
best = 1e9
for rep in range(3):
    n = 4 * D
    t0 = time.perf_counter()
    for i in range(n):
        with torch.cuda.stream(st[i % D]):
            Ss[i % D].Solve()
    torch.cuda.synchronize()
    best = min(best, (time.perf_counter() - t0) / n)
with torch.cuda.stream(st[0]):
    t0 = time.perf_counter(); Ss[0].Solve(); st[0].synchronize(); lat = time.perf_counter() - t0
print(f"{name} depth {D} {dict((k, v) for k, v in os.environ.items() if k.startswith('LOIK_'))}: {best*1e3:.3f} ms/solve = {B/best/1e6:.1f} M solves/s; single-solve latency {lat*1e3:.2f} ms")

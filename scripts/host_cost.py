import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from loik_b200 import problems, robots, solver as lk
name = sys.argv[1] if len(sys.argv) > 1 else "panda"
B = int(os.environ.get("BATCH", 65536))
model = robots.get_robot(name)
pb = problems.random_batch(model, B, seed=0)
D = 8
Ss = [lk.make_solver(model, problems.bench_params(len(pb["ids"])), B) for _ in range(D)]
st = [torch.cuda.Stream() for _ in range(D)]
for S in Ss:
    S.SolveInit(pb["q"], pb["H_ref"], pb["v_ref"], pb["ids"], pb["Ais"], pb["bis"], pb["lb"], pb["ub"])
    S.Solve()
torch.cuda.synchronize()
for rep in range(2):
    t0 = time.perf_counter()
    for i in range(32):
        with torch.cuda.stream(st[i % D]):
            Ss[i % D].Solve()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"B={B}: enqueue {1e3*(t1-t0)/32:.3f} ms/solve, total {1e3*(t2-t0)/32:.3f} ms/solve, launches/solve {(Ss[0].launch_count())}")

import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from loik_b200 import problems, robots, solver as lk
name = sys.argv[1] if len(sys.argv) > 1 else "panda"
B = int(os.environ.get("BATCH", 65536))
model = robots.get_robot(name)
pb = problems.random_batch(model, B, seed=0)
for max_iter in (5, 6, 7, 9, 13, 17, 33, 65, 200):
    for D in (16,):
        P = problems.bench_params(len(pb["ids"]), max_iter=max_iter)
        Ss = [lk.make_solver(model, P, B) for _ in range(D)]
        st = [torch.cuda.Stream() for _ in range(D)]
        for i, S in enumerate(Ss):
            with torch.cuda.stream(st[i]):
                S.SolveInit(pb["q"], pb["H_ref"], pb["v_ref"], pb["ids"], pb["Ais"], pb["bis"], pb["lb"], pb["ub"])
                S.Solve()
        torch.cuda.synchronize()
        n = 4 * D
        t0 = time.perf_counter()
        for i in range(n):
            with torch.cuda.stream(st[i % D]):
                Ss[i % D].Solve()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / n
        print(f"max_iter {max_iter:4d} depth {D}: {dt*1e3:.3f} ms/solve")
        for S in Ss:
            S.close()

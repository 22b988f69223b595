import os, sys, torch
sys.path.insert(0, '/root/repo')
from loik_b200 import problems, robots, solver as lk
model = robots.get_robot("talos"); B = 1
pb = problems.random_batch(model, B, seed=0)
S = lk.make_solver(model, problems.bench_params(2), B)
S.set_schedule(lane_after=0, lane_groups_per_instance=int(sys.argv[1]))
S.SolveInit(pb["q"], pb["H_ref"], pb["v_ref"], pb["ids"], pb["Ais"], pb["bis"], pb["lb"], pb["ub"])
for _ in range(2): S.IterateFixed(50)
torch.cuda.synchronize(); S.close()

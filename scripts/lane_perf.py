#!/usr/bin/env python
"""Development aid: timings of the launch schedule's switch point (loik_schedule.lane_after) on the BASELINE batches.

  python scripts/lane_perf.py panda,ur10,talos            # fixed-iteration rate, one solve alone, pipelined solves
Env: LANE_AFTERS="-1,0,4,8,16" (schedules to try), DEPTHS="1,4,32" (handles in flight), BATCH.
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
from loik_b200 import problems, robots, solver as lk  # noqa: E402

BATCHES = {"panda": 65536, "ur10": 262144, "talos": 16384, "talos_ff": 16384, "panda9": 65536}


def main():
    names = sys.argv[1].split(",") if len(sys.argv) > 1 else ["panda"]
    afters = [int(x) for x in os.environ.get("LANE_AFTERS", "-1,0,4,8,16").split(",")]
    depths = [int(x) for x in os.environ.get("DEPTHS", "1,4,32").split(",")]
    extra = {}
    for kv in os.environ.get("SCHED", "").split(","):
        if "=" in kv:
            k, v = kv.split("=")
            extra[k] = float(v) if "." in v else int(v)
    for name in names:
        model = robots.get_robot(name)
        B = int(os.environ.get("BATCH", BATCHES[name]))
        pb = problems.random_batch(model, B, seed=0)
        nc = len(pb["ids"])
        params = problems.bench_params(nc)
        bpi = 8 * (143 * model.nb + 42 * nc)
        for la in afters:
            D = max(depths)
            Ss = [lk.make_solver(model, params, B) for _ in range(D)]
            for S in Ss:
                S.set_schedule(lane_after=la, **extra)
                S.SolveInit(pb["q"], pb["H_ref"], pb["v_ref"], pb["ids"], pb["Ais"], pb["bis"], pb["lb"], pb["ub"])
            S = Ss[0]
            sc = S.get_schedule()
            K = 20
            S.IterateFixed(3)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); S.IterateFixed(K); e1.record(); torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / K
            line = (f"{name} B={B} lane_after={la} (W={sc['lane_warps_chosen']} gpi={sc['lane_groups_chosen']} ctas={sc['lane_ctas']} smem={sc['lane_smem_bytes']}): "
                    f"fixed {us:.1f} us/iter = {B/us:.0f} M inst-it/s = {bpi*B/us/1e3/6547.5:.3f} of HBM (algorithmic)")
            for _ in range(2):
                S.Solve()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(3):
                S.Solve()
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / 3
            st = S.stats()
            line += f"; one solve {dt*1e3:.2f} ms ({B/dt/1e6:.1f} M/s, mean iters {st['total_iters']/B:.2f})"
            streams = [torch.cuda.Stream() for _ in range(D)]
            for d in depths:
                steps = max(4 * d, 8)
                for rep in range(2):
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    for i in range(steps):
                        with torch.cuda.stream(streams[i % d]):
                            Ss[i % d].Solve()
                    torch.cuda.synchronize()
                    dt = time.perf_counter() - t0
                line += f"; depth {d}: {B*steps/dt/1e6:.1f} M/s"
            print(line, flush=True)
            for S in Ss:
                S.close()


if __name__ == "__main__":
    main()

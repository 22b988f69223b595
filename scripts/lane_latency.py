#!/usr/bin/env python
"""Development aid: per-iteration LATENCY of the iteration kernels on a handful of instances (what the straggler tail of a solve
pays): fixed-iteration launches at batch 4 / 32 / 592, lane kernel (lane_after = 0) against the tile kernel (lane_after = -1).

  python scripts/lane_latency.py panda,ur10,talos
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from loik_b200 import problems, robots, solver as lk  # noqa: E402


def main():
    names = sys.argv[1].split(",") if len(sys.argv) > 1 else ["panda"]
    batches = [int(x) for x in os.environ.get("BATCHES", "4,32,592").split(",")]
    K = int(os.environ.get("ITERS", "200"))
    for name in names:
        model = robots.get_robot(name)
        for B in batches:
            pb = problems.random_batch(model, B, seed=0)
            params = problems.bench_params(len(pb["ids"]))
            out = []
            for la, gpi in ((0, 1), (0, 4), (-1, 0)):
                S = lk.make_solver(model, params, B)
                S.set_schedule(lane_after=la, lane_groups_per_instance=gpi)
                S.SolveInit(pb["q"], pb["H_ref"], pb["v_ref"], pb["ids"], pb["Ais"], pb["bis"], pb["lb"], pb["ub"])
                sc = S.get_schedule()
                if la == 0 and (not sc["lane_available"] or (gpi == 4 and sc["lane_groups_chosen"] != 4)):
                    S.close()
                    continue
                S.IterateFixed(K)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); S.IterateFixed(K); e1.record(); torch.cuda.synchronize()
                us = e0.elapsed_time(e1) * 1e3 / K
                out.append(f"{'lane gpi=%d' % sc['lane_groups_chosen'] if la == 0 else 'tile kernel (1 launch / iteration)'}: {us:.2f} us/iter")
                S.close()
            print(f"{name} nb={model.nb} B={B}: " + "; ".join(out), flush=True)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Quick device timing of the iteration kernel and of full solves (development aid, not the bench)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from loik_b200 import problems, robots, solver as lk  # noqa: E402


def main():
    names = sys.argv[1].split(",") if len(sys.argv) > 1 else ["panda"]
    batches = {"panda": 65536, "ur10": 262144, "talos": 16384, "talos_ff": 16384}
    for name in names:
        model = robots.get_robot(name)
        B = int(os.environ.get("BATCH", batches[name]))
        pb = problems.random_batch(model, B, seed=0)
        nc = len(pb["ids"])
        S = lk.make_solver(model, problems.bench_params(nc), B)
        S.SolveInit(pb["q"], pb["H_ref"], pb["v_ref"], pb["ids"], pb["Ais"], pb["bis"], pb["lb"], pb["ub"])
        S.IterateFixed(5)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); S.IterateFixed(50); e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / 50
        bpi = 8 * (143 * model.nb + 42 * nc)
        print(f"{name} B={B}: {us:.1f} us/iter, {B/us:.1f} M inst-it/s, "
              f"{bpi*B/us/1e3:.0f} GB/s algorithmic = {bpi*B/us/1e3/6547.5:.3f} of HBM")
        for _ in range(2):
            S.Solve()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            S.Solve()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 3
        st = S.stats()
        print(f"   solve: {dt*1e3:.2f} ms, {B/dt/1e6:.2f} M solves/s, mean iters {st['total_iters']/B:.2f}")
        S.close()


if __name__ == "__main__":
    main()


def pipelined(name="panda", depths=tuple(int(x) for x in os.environ.get("DEPTHS", "1,2,4,8").split(","))):
    """Throughput with D solver handles in flight on D streams (tail of one batch overlaps the bulk of the next)."""
    model = robots.get_robot(name)
    B = int(os.environ.get("BATCH", {"panda": 65536, "ur10": 262144, "talos": 16384, "talos_ff": 16384}[name]))
    pb = problems.random_batch(model, B, seed=0)
    nc = len(pb["ids"])
    for D in depths:
        Ss = [lk.make_solver(model, problems.bench_params(nc), B) for _ in range(D)]
        streams = [torch.cuda.Stream() for _ in range(D)]
        for S in Ss:
            S.SolveInit(pb["q"], pb["H_ref"], pb["v_ref"], pb["ids"], pb["Ais"], pb["bis"], pb["lb"], pb["ub"])
        torch.cuda.synchronize()
        steps = 4 * D
        for rep in range(2):
            t0 = time.perf_counter()
            for i in range(steps):
                with torch.cuda.stream(streams[i % D]):
                    Ss[i % D].Solve()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        print(f"{name} pipeline depth {D}: {dt/steps*1e3:.3f} ms/solve, {B*steps/dt/1e6:.2f} M solves/s")
        for S in Ss:
            S.close()


if __name__ == "__main__" and os.environ.get("PIPE"):
    for nm in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["panda"]):
        pipelined(nm)

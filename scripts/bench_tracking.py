#!/usr/bin/env python
"""Trajectory-tracking benchmark (the reference's real-time use, hpp:596-695).

Per solver handle: one full Solve, then T steps of { loik_integrate(dt) ; Solve(q = device-resident, c_id, A, b_t) } with the
state warm-started (warm_start = true) and only the per-instance target twist b_t (48 B per instance) changing, resident in HBM.
D handles (different trajectories of the same robot) are kept in flight on D streams.  Prints one JSON line: tracking
solves/s (device-timed, CUDA events), mean ADMM iterations per tracking solve, and the CPU restatement driven the same way
(lo_batch_track, all host cores) on a bounded sample.
"""
import argparse
import json
import os
import sys
import time

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from loik_b200 import problems, robots, solver as lk  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--robot", default="panda")
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--steps", type=int, default=50, help="tracking steps per handle inside the timed region")
    ap.add_argument("--pipeline", type=int, default=16)
    ap.add_argument("--dt", type=float, default=0.01)
    ap.add_argument("--cold", action="store_true", help="warm_start = false (every tracking solve starts from zero state)")
    ap.add_argument("--cpu-sample", type=int, default=4096)
    args = ap.parse_args()
    model = robots.get_robot(args.robot)
    B, D, T = args.batch, args.pipeline, args.steps
    c_id = int(robots.TASK_JOINTS[args.robot][0])
    pbs = [problems.random_batch(model, B, seed=100 + k) for k in range(D)]
    nxt = [problems.random_batch(model, B, seed=200 + k) for k in range(D)]
    nc = len(pbs[0]["ids"])
    slot = list(pbs[0]["ids"]).index(c_id)
    params = dict(problems.bench_params(nc), warm_start=not args.cold)
    dev = torch.device("cuda", 0)
    solvers = [lk.make_solver(model, params, B) for _ in range(D)]
    streams = [torch.cuda.Stream() for _ in range(D)]
    b0 = [torch.as_tensor(p["bis"][:, slot].copy(), device=dev) for p in pbs]
    b1 = [torch.as_tensor(p["bis"][:, slot].copy(), device=dev) for p in nxt]
    A = pbs[0]["Ais"][slot]
    bt = [[((1.0 - (t + 1) / T) * b0[k] + ((t + 1) / T) * b1[k]).contiguous() for t in range(T)] for k in range(D)] if B * T * D * 48 < 8e9 else None
    assert bt is not None, "reduce --steps/--pipeline"
    for k, S in enumerate(solvers):
        p = pbs[k]
        with torch.cuda.stream(streams[k]):
            S.Solve(p["q"], p["H_ref"], p["v_ref"], p["ids"], p["Ais"], p["bis"], p["lb"], p["ub"])
            for t in range(3):  # warm-up: graph capture of the tailored solve
                S.Integrate(0.0)
                S.Solve(None, c_id, A, b0[k])
            # start every trajectory from the state a fresh solver object has after its first full Solve
            S.set_warm_start(False)
            S.Solve(p["q"], p["H_ref"], p["v_ref"], p["ids"], p["Ais"], p["bis"], p["lb"], p["ub"])
            S.set_warm_start(not args.cold)
    torch.cuda.synchronize()
    it_sum = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cur = torch.cuda.current_stream()
    e0.record(cur)
    for st in streams:
        st.wait_event(e0)
    for t in range(T):
        for k, S in enumerate(solvers):
            with torch.cuda.stream(streams[k]):
                S.Integrate(args.dt)
                S.Solve(None, c_id, A, bt[k][t])
    for st in streams:
        ev = torch.cuda.Event(); ev.record(st); cur.wait_event(ev)
    e1.record(cur)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    it_last = np.mean([S.get_iter().mean() for S in solvers])
    line = {"metric": "tracking IK solves/s (Integrate + warm-started Solve(q, c_id, A, b), device-timed)", "value": B * D * T / (ms * 1e-3),
            "unit": "IK solves/s", "robot": args.robot, "batch": B, "handles": D, "steps": T, "dt": args.dt, "warm_start": not args.cold,
            "ms_per_tracking_step_per_handle": ms / T / D, "mean_iters_last_step": float(it_last)}
    # CPU restatement driven the same way (bounded sample)
    from oracle import recursion
    n = min(args.cpu_sample, B)
    p, q = pbs[0], nxt[0]
    cores = os.cpu_count() or 1
    t0 = time.perf_counter()
    o = recursion.batch_track(model, params, p["q"][:n], p["H_ref"], p["v_ref"], p["ids"], p["Ais"], p["bis"][:n], q["bis"][:n], p["lb"],
                              p["ub"], c_id=c_id, dt=args.dt, steps=T, warm=not args.cold, nthreads=cores)
    dt = time.perf_counter() - t0
    line["cpu_baseline"] = {"value": n * T / dt, "unit": "IK solves/s (incl. the initial full solve of each instance)", "cores": cores, "kind": "port",
                            "sample": f"{n} instances x {T} steps", "mean_iters_per_tracking_solve": float(o["step_iters"].mean())}
    # parity of the sample: same trajectory on the GPU handle 0
    zs, qs = solvers[0].z[:n], solvers[0].q[:n]
    its = solvers[0].get_iter()[:n]
    rq = np.abs(qs - o["q"]).max(axis=1) / np.maximum(1e-12, np.abs(o["q"]).max(axis=1))
    same = (its == o["step_iters"][:, -1]) & (rq < 1e-9)  # a decision that flips on rounding at some step forks that trajectory
    rel = np.abs(zs - o["z"]).max(axis=1) / np.maximum(1e-12, np.abs(o["z"]).max(axis=1))
    line["parity_sample"] = {"same_trajectory_fraction": float(same.mean()),
                             "worst_rel_inf_z_on_same_trajectories": float(rel[same].max()) if same.any() else None}
    print(json.dumps(line))
    for S in solvers:
        S.close()


if __name__ == "__main__":
    main()

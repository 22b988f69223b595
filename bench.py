#!/usr/bin/env python
"""bench.py -- headline benchmark of the batched LoIK hot path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload panda|ur10|talos] [--batch B]
  python bench.py --impl reference ...      # the CPU restatement of loik-loid-optimized on the host cores

A "step" is one full batched solve (SolveInit's device part excluded, Solve() = ResetRecursion +
ResetSolver + the ADMM loop, every instance to its own convergence / infeasibility tail / max_iter) of the
synthetic batch BASELINE.json names; `value` = IK solves per second over all ranks with inputs resident in HBM;
`e2e` = the same through the public API with HOST buffers (q, b in; z, iteration counts out) inside the timed
region.  `roofline` is quoted on the ADMM-iteration kernel in fixed-iteration mode (every instance active,
one launch per iteration): algorithmic bytes per launch = 8*(143 n + 42 nc) * batch.

For N > 1 (torchrun) the batch is sharded across ranks (weak scaling: per-GPU batch fixed); loop control is per
instance on the device, and the only collective is one all-reduce of four int64 per solve: the global stopping-
criterion outcome (#converged, #infeasible, #max_iter, total iterations).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# The pipelined solver handles each own a stream (+ a high-priority one); with the default of 8 hardware queues
# streams alias and serialise.  Must be set before CUDA initialises.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

from loik_b200 import problems, robots  # noqa: E402

WORKLOADS = {  # BASELINE.json configs[1..3]: (robot, per-GPU batch)
    "panda": ("panda", 65536),
    "ur10": ("ur10", 262144),
    "talos": ("talos", 16384),
    "talos_ff": ("talos_ff", 16384),  # floating base (free-flyer root, nv = 38): not a BASELINE config, SURVEY.md 8(f) rank 4
}
FIXED_ITERS = 50
METRIC = "IK solves/sec (batch, device-timed)"


def algorithmic_bytes_per_instance_iteration(n, nc):
    return 8 * (143 * n + 42 * nc)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def cpu_baseline(model, pb, params, seconds_target=12.0, lib=None):
    """Oracle B (CPU restatement of loik-loid-optimized) on the host cores, bounded sample of the same batch."""
    from oracle import recursion
    cores = os.cpu_count() or 1
    B = pb["q"].shape[0]
    probe = min(B, 256 * cores)
    sub = dict(pb, q=pb["q"][:probe], bis=pb["bis"][:probe])
    t0 = time.perf_counter()
    recursion.batch_solve(model, params, sub["q"], sub["H_ref"], sub["v_ref"], sub["ids"], sub["Ais"], sub["bis"], sub["lb"],
                          sub["ub"], nthreads=cores, want_outputs=False, lib=lib)
    rate = probe / max(time.perf_counter() - t0, 1e-6)
    n = int(min(B, max(probe, rate * seconds_target)))
    sub = dict(pb, q=pb["q"][:n], bis=pb["bis"][:n])
    t0 = time.perf_counter()
    out = recursion.batch_solve(model, params, sub["q"], sub["H_ref"], sub["v_ref"], sub["ids"], sub["Ais"], sub["bis"],
                                sub["lb"], sub["ub"], nthreads=cores, want_outputs=False, lib=lib)
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": "IK solves/s", "cores": cores, "kind": "port",
            "sample": f"first {n} of {B} instances of the same batch, {cores} threads, one solver per thread, "
                      f"{out['total_iters'] / n:.2f} iterations/solve",
            "iters_per_s": out["total_iters"] / dt}


def native_oracle_lib():
    """Rebuild the oracle with -march=native on this host when gcc is present (fairer CPU baseline)."""
    from oracle import recursion
    try:
        out = os.path.join("/tmp", f"libloik_oracle_native_{os.getpid()}.so")
        recursion.build(out=out, march="native")
        return recursion.load(out)
    except Exception:
        return recursion.load()


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path (restated: the reference cannot be built offline) on the host cores."""
    if rank != 0:
        return
    robot, batch = WORKLOADS[args.workload]
    batch = args.batch or batch
    model = robots.get_robot(robot)
    pb = problems.random_batch(model, batch, seed=0)
    params = problems.bench_params(len(pb["ids"]))
    lib = native_oracle_lib()
    from oracle import recursion
    cores = os.cpu_count() or 1
    # size one step to ~ (120 s / (steps+warmup)) of CPU work
    probe = min(batch, 128 * cores)
    t0 = time.perf_counter()
    recursion.batch_solve(model, params, pb["q"][:probe], pb["H_ref"], pb["v_ref"], pb["ids"], pb["Ais"], pb["bis"][:probe],
                          pb["lb"], pb["ub"], nthreads=cores, want_outputs=False, lib=lib)
    rate = probe / max(time.perf_counter() - t0, 1e-6)
    per_step = int(min(batch, max(probe, rate * 120.0 / (args.steps + args.warmup))))
    times = []
    for i in range(args.warmup + args.steps):
        lo = (i * per_step) % max(batch - per_step + 1, 1)
        t0 = time.perf_counter()
        recursion.batch_solve(model, params, pb["q"][lo:lo + per_step], pb["H_ref"], pb["v_ref"], pb["ids"], pb["Ais"],
                              pb["bis"][lo:lo + per_step], pb["lb"], pb["ub"], nthreads=cores, want_outputs=False, lib=lib)
        if i >= args.warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    value = per_step * len(times) / total
    sample = f"{per_step} of {batch} instances per step, {cores} threads, one solver per thread (CPU restatement of loik-loid-optimized; reference not buildable offline)"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "IK solves/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{robot} batch {batch} per GPU (BASELINE.json configs)", "robot": robot, "n_dof": model.nb,
                       "n_tasks": len(pb["ids"]), "batch_per_gpu": batch, "max_iter": params["max_iter"]},
            "cpu_baseline": {"value": value, "unit": "IK solves/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "IK solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="loik_b200", choices=["loik_b200", "reference"])
    ap.add_argument("--workload", default="panda", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: the BASELINE config's)")
    ap.add_argument("--pipeline", type=int, default=32,
                    help="solver handles (each on its own stream) kept in flight; step i uses handle i %% depth")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    from loik_b200 import sharded
    from loik_b200 import solver as lk

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    robot, batch = WORKLOADS[args.workload]
    batch = args.batch or batch
    model = robots.get_robot(robot)
    n, nc = model.nb, len(robots.TASK_JOINTS[robot])
    pb = problems.random_batch(model, batch, seed=0, first_index=rank * batch)  # this rank's shard of the global batch
    params = problems.bench_params(nc)
    # each handle owns a home arena + two re-pack arenas; keep the pipeline within ~60 GB of HBM
    rows_est = 48 + 61 * n + 24 * nc + 33 + 14 * n
    bytes_per_handle = 3 * ((batch + 31) // 32) * rows_est * 256
    D = max(1, min(args.pipeline, int(60e9 // bytes_per_handle)))
    solvers = [lk.make_solver(model, params, batch, device=local_rank) for _ in range(D)]
    drivers = [sharded.ShardedSolver(S, world) for S in solvers]
    streams = [torch.cuda.Stream(device=dev) for _ in range(D)]

    # resident inputs (value) and pinned host inputs/outputs (e2e)
    q_d = torch.as_tensor(pb["q"], device=dev)
    b_d = torch.as_tensor(pb["bis"], device=dev)
    q_h = torch.as_tensor(pb["q"]).pin_memory()
    b_h = torch.as_tensor(pb["bis"]).pin_memory()
    z_h = [torch.empty(batch, n, dtype=torch.float64).pin_memory() for _ in range(D)]
    it_h = [torch.empty(batch, dtype=torch.int32).pin_memory() for _ in range(D)]
    prob = (pb["H_ref"], pb["v_ref"], pb["ids"], pb["Ais"])

    def step_resident(i):
        k = i % D
        with torch.cuda.stream(streams[k]):
            drivers[k].solve()

    def step_e2e(i):
        # the reference-facing call sequence with HOST buffers: SolveInit(q, ..., b, ...) -> Solve() -> read z, iter
        k = i % D
        S = solvers[k]
        with torch.cuda.stream(streams[k]):
            S.SolveInit(q_h, prob[0], prob[1], prob[2], prob[3], b_h, pb["lb"], pb["ub"])  # pinned host -> HBM inside
            drivers[k].solve()
            S.get(lk.F_Z, out=z_h[k])        # HBM -> pinned host inside
            S.get(lk.F_ITER, out=it_h[k])

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        cur = torch.cuda.current_stream()
        e0.record(cur)
        for st in streams:
            st.wait_event(e0)
        for i in range(steps):
            fn(i)
        for st in streams:
            done = torch.cuda.Event()
            done.record(st)
            cur.wait_event(done)
        e1.record(cur)
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()  # clocks / throttle reasons are sampled from the warm-up to the end of the last timed region
    for S in solvers:
        S.SolveInit(q_d, prob[0], prob[1], prob[2], prob[3], b_d, pb["lb"], pb["ub"])
    for i in range(max(args.warmup, D)):  # every handle allocates its re-pack arenas and captures its graph once
        step_resident(i)
    barrier()
    launches0 = sum(S.launch_count() for S in solvers)
    ms_total = timed(step_resident, args.steps)
    launches = sum(S.launch_count() for S in solvers) - launches0
    stats = solvers[0].stats()
    mean_iters = stats["total_iters"] / batch
    # latency of one un-pipelined solve (one handle, one stream)
    ms_single = timed(lambda i: step_resident(0), 3) / 3

    # fixed-iteration mode: the roofline kernel (one launch = one ADMM iteration of the whole batch, all active)
    S0 = solvers[0]
    S0.IterateFixed(3)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    S0.IterateFixed(FIXED_ITERS)
    e1.record()
    barrier()
    ms_fixed = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms_fixed, op=dist.ReduceOp.MAX)
    ms_iter = float(ms_fixed.item()) / FIXED_ITERS

    # e2e
    for i in range(max(2, D)):
        step_e2e(i)
    ms_e2e = timed(step_e2e, args.steps)
    clocks = sampler.stop() if rank == 0 else None

    if rank == 0:
        hbm, how = peaks()
        bpi = algorithmic_bytes_per_instance_iteration(n, nc)
        achieved = bpi * batch / (ms_iter * 1e-3) / 1e9
        value = world * batch * args.steps / (ms_total * 1e-3)
        e2e_v = world * batch * args.steps / (ms_e2e * 1e-3)
        line = {
            "metric": METRIC, "value": value, "unit": "IK solves/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{robot} batch {batch} per GPU (BASELINE.json configs)", "robot": robot, "n_dof": n,
                       "n_tasks": nc, "batch_per_gpu": batch, "global_batch": world * batch, "max_iter": params["max_iter"],
                       "l2": "inputs larger than L2: per-iteration working set %.0f MB" % (bpi * batch / 2 ** 20),
                       "parallelism": f"batch-sharded x{world}", "pipeline_depth": D,
                       "pipeline": "step i runs on solver handle i % depth (own HBM state, own stream): the latency-bound "
                                   "tail of one batch overlaps the bulk of the next",
                       "mean_iters_per_solve": mean_iters},
            "ms_per_solve_unpipelined": ms_single,
            "iters_per_s": world * batch / (ms_iter * 1e-3),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
                         "traffic": None, "peak_source": how,
                         "kernel": "k_iterate (1 ADMM iteration / launch, all instances active)",
                         "algorithmic_bytes_per_launch": bpi * batch, "us_per_launch": ms_iter * 1e3,
                         "frac_of_nominal_8TBs": achieved / 8000.0},
            "e2e": {"value": e2e_v, "unit": "IK solves/s", "h2d_bytes_per_step": int(q_h.numel() * 8 + b_h.numel() * 8),
                    "d2h_bytes_per_step": int(z_h[0].numel() * 8 + it_h[0].numel() * 4), "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches), "clocks": clocks,
            "solve_stats": {k: int(v) for k, v in stats.items()},
        }
        traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(traffic_file):
            try:
                with open(traffic_file) as f:
                    line["roofline"]["traffic"] = json.load(f).get(robot)
            except Exception:
                pass
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(model, pb, params, lib=native_oracle_lib())
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line))
    for S in solvers:
        S.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

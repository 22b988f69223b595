#!/usr/bin/env python
"""bench.py -- headline benchmark of the batched LoIK hot path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload panda|ur10|talos|all] [--batch B]
  python bench.py --impl reference ...      # the CPU restatement of loik-loid-optimized on the host cores

A "step" is one full batched solve (Solve() = ResetRecursion + ResetSolver + the ADMM loop, every instance to its own
convergence / infeasibility tail / max_iter) of a synthetic batch BASELINE.json names.  The parsed headline is
BASELINE configs[1] (Panda 7-DoF x 65 536); the default run (`--workload all`) also measures configs[2] (UR10 x 262 144)
and configs[3] (Talos x 16 384; under torchrun this is the per-GPU shard of configs[4], Talos x 131 072 on 8 GPUs) and
reports them as sub-records under `extra.workloads`, each with its own value / e2e / roofline / mean iterations.

  value      IK solves/s over all ranks, inputs resident in HBM, steps round-robin over `pipeline_depth` solver handles
             (own HBM state, own stream).  The depth is capped at steps // 2 so that every handle runs at least two timed
             solves: the timed region is a steady state, not a burst of first solves.
  e2e        the same through the public API with pinned HOST buffers (q, b in; z, iteration counts out) inside the
             timed region.
  roofline   the ADMM-iteration kernel in fixed-iteration mode (every instance active, one launch per iteration):
             algorithmic bytes per launch = 8*(143 n + 42 nc) * batch / CUDA-event time.
  extra      un-pipelined latency of one solve, the rate at pipeline depth 4, the converged-only rate, the lane-parallel
             kernel's fixed-iteration rate, the launch schedule in use.

For N > 1 (torchrun) the batch is sharded across ranks (weak scaling: per-GPU batch fixed); loop control is per
instance on the device, and the only collective is one all-reduce of four int64 per solve: the global stopping-
criterion outcome (#converged, #infeasible, #max_iter, total iterations).
"""
from __future__ import annotations

import argparse
import hashlib
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
SAT_ITERS = 10     # dense iterations per handle and round in the launches-in-flight measurement
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


def source_sha():
    """Hash of the kernel sources: ties profiles/traffic.json (ncu DRAM bytes) to the build it was captured on."""
    h = hashlib.sha256()
    for f in ("loik_device.cuh", "loik_lane.cuh", "loik_solver.cu"):
        with open(os.path.join(ROOT, "loik_b200", "csrc", f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


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


def native_oracle_lib():
    """Rebuild the oracle with -march=native on this host when gcc is present (fairer CPU baseline)."""
    from oracle import recursion
    try:
        out = os.path.join("/tmp", f"libloik_oracle_native_{os.getpid()}.so")
        recursion.build(out=out, march="native")
        return recursion.load(out)
    except Exception:
        return recursion.load()


def cpu_rate(model, pb, params, lib, per_pass, passes, warm):
    """Oracle B (CPU restatement of loik-loid-optimized) on all host cores: `warm` untimed + `passes` timed passes over
    windows of `per_pass` instances of the batch.  Returns (solves/s, iterations/s, mean iterations)."""
    from oracle import recursion
    cores = os.cpu_count() or 1
    B = pb["q"].shape[0]
    tot_t, tot_n, tot_it = 0.0, 0, 0
    for i in range(warm + passes):
        lo = (i * per_pass) % max(B - per_pass + 1, 1)
        t0 = time.perf_counter()
        out = recursion.batch_solve(model, params, pb["q"][lo:lo + per_pass], pb["H_ref"], pb["v_ref"], pb["ids"], pb["Ais"],
                                    pb["bis"][lo:lo + per_pass], pb["lb"], pb["ub"], nthreads=cores, want_outputs=False, lib=lib)
        dt = time.perf_counter() - t0
        if i >= warm:
            tot_t += dt; tot_n += per_pass; tot_it += out["total_iters"]
    return tot_n / tot_t, tot_it / tot_t, tot_it / max(tot_n, 1)


def cpu_pass_size(model, pb, params, lib, seconds_per_pass):
    from oracle import recursion
    cores = os.cpu_count() or 1
    B = pb["q"].shape[0]
    probe = min(B, 128 * cores)
    recursion.batch_solve(model, params, pb["q"][:probe], pb["H_ref"], pb["v_ref"], pb["ids"], pb["Ais"], pb["bis"][:probe],
                          pb["lb"], pb["ub"], nthreads=cores, want_outputs=False, lib=lib)  # (first call: thread start-up, page faults)
    t0 = time.perf_counter()
    recursion.batch_solve(model, params, pb["q"][:probe], pb["H_ref"], pb["v_ref"], pb["ids"], pb["Ais"], pb["bis"][:probe],
                          pb["lb"], pb["ub"], nthreads=cores, want_outputs=False, lib=lib)
    rate = probe / max(time.perf_counter() - t0, 1e-6)
    return int(min(B, max(probe, rate * seconds_per_pass)))


def cpu_baseline(model, pb, params, lib, seconds_target=12.0):
    """The reported in-line CPU baseline: warmed, several passes, same protocol as `--impl reference`."""
    cores = os.cpu_count() or 1
    B = pb["q"].shape[0]
    passes = 20  # (a pass over the whole Panda batch takes ~70 ms on 16 cores: ~1.5 s of wall clock, ~20 core-seconds in all)
    n = cpu_pass_size(model, pb, params, lib, seconds_target / (passes + 1))
    v, its, mean_it = cpu_rate(model, pb, params, lib, n, passes, warm=1)
    return {"value": v, "unit": "IK solves/s", "cores": cores, "kind": "port",
            "sample": f"{passes} timed passes (1 warm-up) over windows of {n} of the {B} instances of the same batch, {cores} threads, "
                      f"one solver per thread, {mean_it:.2f} iterations/solve",
            "iters_per_s": its}


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path (restated: the reference cannot be built offline) on the host cores."""
    if rank != 0:
        return
    name = "panda" if args.workload == "all" else args.workload
    robot, batch = WORKLOADS[name]
    batch = args.batch or batch
    model = robots.get_robot(robot)
    pb = problems.random_batch(model, batch, seed=0)
    params = problems.bench_params(len(pb["ids"]))
    lib = native_oracle_lib()
    cores = os.cpu_count() or 1
    per_step = cpu_pass_size(model, pb, params, lib, 120.0 / (args.steps + args.warmup))
    value, its, mean_it = cpu_rate(model, pb, params, lib, per_step, args.steps, warm=args.warmup)
    sample = (f"{per_step} of {batch} instances per step, {cores} threads, one solver per thread (CPU restatement of "
              f"loik-loid-optimized; reference not buildable offline), {mean_it:.2f} iterations/solve")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "IK solves/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * per_step / value, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{robot} batch {batch} per GPU (BASELINE.json configs)", "robot": robot, "n_dof": model.nb,
                       "n_tasks": len(pb["ids"]), "batch_per_gpu": batch, "max_iter": params["max_iter"]},
            "cpu_baseline": {"value": value, "unit": "IK solves/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "IK solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def run_workload(name, args, rank, world, local_rank, dev, headline):
    """All measurements of one BASELINE workload on this rank's shard; returns the record (rank 0) or None."""
    import torch
    import torch.distributed as dist
    from loik_b200 import sharded
    from loik_b200 import solver as lk

    robot, batch = WORKLOADS[name]
    batch = args.batch or batch
    model = robots.get_robot(robot)
    n, nc = model.nb, len(robots.TASK_JOINTS[robot])
    pb = problems.random_batch(model, batch, seed=0, first_index=rank * batch)  # this rank's shard of the global batch
    params = problems.bench_params(nc)
    steps = args.steps if headline else max(8, min(args.steps, 32))
    warmup = args.warmup
    # each handle owns a home arena + two re-pack arenas; keep the pipeline within ~60 GB of HBM.  Depth <= steps // 2:
    # every handle runs at least two timed solves (steady state)
    rows_est = 48 + 61 * n + 24 * nc + 33 + 14 * n
    bytes_per_handle = 3 * ((batch + 31) // 32) * rows_est * 256
    D = max(1, min(args.pipeline, int(60e9 // bytes_per_handle), max(1, steps // 2)))
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

    def step_resident(i, depth=D):
        k = i % depth
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

    def timed(fn, nsteps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        cur = torch.cuda.current_stream()
        e0.record(cur)
        for st in streams:
            st.wait_event(e0)
        for i in range(nsteps):
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

    for S in solvers:
        S.SolveInit(q_d, prob[0], prob[1], prob[2], prob[3], b_d, pb["lb"], pb["ub"])
    schedule = solvers[0].get_schedule()
    for i in range(max(warmup, 2 * D)):  # every handle allocates its re-pack arenas and captures its graph, then runs warm once
        step_resident(i)
    barrier()
    launches0 = sum(S.launch_count() for S in solvers)
    ms_total = timed(step_resident, steps)
    launches = sum(S.launch_count() for S in solvers) - launches0
    stats = solvers[0].stats()
    mean_iters = stats["total_iters"] / batch
    # share of the instance-iterations of a solve that the roofline kernel (k_iterate) executes: everything up to the switch to the
    # lane-parallel kernel (an ncu launch list shows device TIME, where the latency-bound tail of the lane kernel weighs far more)
    it_all = solvers[0].get_iter().astype(np.int64)
    la = schedule["lane_after"] if (schedule["lane_available"] and schedule["lane_after"] >= 0) else 10 ** 9
    work_share = float(np.minimum(it_all, la).sum()) / max(1, int(it_all.sum()))
    # latency of one un-pipelined solve (one handle, one stream), and the rate with 4 handles in flight
    ms_single = timed(lambda i: step_resident(0, 1), 3) / 3
    d4 = min(4, D)
    n4 = max(8, 2 * d4)
    for i in range(d4):
        step_resident(i, d4)
    ms_d4 = timed(lambda i: step_resident(i, d4), n4)

    # fixed-iteration mode: the roofline kernel (one launch = one ADMM iteration of the whole batch, all active)
    S0 = solvers[0]
    saved_lane_after = schedule["lane_after"]

    def fixed_us(lane_after, iters):
        S0.set_schedule(lane_after=lane_after)
        S0.IterateFixed(3)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        S0.IterateFixed(iters)
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) * 1e3 / iters

    # the latency-oriented schedule (the defaults favour pipelined throughput): the lane-parallel kernel takes over early
    ms_single_lane, lat_after = None, None
    if schedule["lane_available"]:
        lat_after = 8 if schedule["lane_after"] < 0 else 4
        S0.set_schedule(lane_after=lat_after)
        step_resident(0, 1)
        ms_single_lane = timed(lambda i: step_resident(0, 1), 3) / 3
        S0.set_schedule(lane_after=saved_lane_after)
    us_iter = fixed_us(-1, FIXED_ITERS)                     # k_iterate, one launch per iteration
    us_lane = fixed_us(0, 20) if schedule["lane_available"] else None  # k_iterate_lane, 20 iterations per instance in one launch
    S0.set_schedule(lane_after=saved_lane_after)
    # the same kernel with launches of several solver handles in flight (what the pipelined solves run as): one launch alone
    # leaves warp slots empty when the batch is small (Talos-16 384: 512 warps on 1 184 slots)
    nsat = min(8, D)
    for S in solvers[:nsat]:
        S.set_schedule(lane_after=-1)
        S.IterateFixed(1)  # (reset: every instance active again; fixed mode keeps them active)

    def sat_round(i):
        for k in range(nsat):
            with torch.cuda.stream(streams[k]):
                solvers[k].IterateFixed(SAT_ITERS, reset=False)
    sat_round(0)
    us_sat = timed(sat_round, 2) * 1e3 / (2 * SAT_ITERS * nsat)  # per launch-equivalent
    for S in solvers[:nsat]:
        S.set_schedule(lane_after=saved_lane_after)
        S.SolveInit(q_d, prob[0], prob[1], prob[2], prob[3], b_d, pb["lb"], pb["ub"])

    # e2e
    for i in range(max(2, D)):
        step_e2e(i)
    ms_e2e = timed(step_e2e, steps)

    rec = None
    if rank == 0:
        hbm, how = peaks()
        bpi = algorithmic_bytes_per_instance_iteration(n, nc)
        achieved = bpi * batch / (us_iter * 1e-6) / 1e9
        value = world * batch * steps / (ms_total * 1e-3)
        e2e_v = world * batch * steps / (ms_e2e * 1e-3)
        rec = {
            "value": value, "unit": "IK solves/s", "steps": steps, "ms_per_step": ms_total / steps,
            "config": {"workload": f"{robot} batch {batch} per GPU (BASELINE.json configs)", "robot": robot, "n_dof": n,
                       "n_tasks": nc, "batch_per_gpu": batch, "global_batch": world * batch, "max_iter": params["max_iter"],
                       "l2": "inputs larger than L2: per-iteration working set %.0f MB" % (bpi * batch / 2 ** 20),
                       "parallelism": f"batch-sharded x{world}", "pipeline_depth": D,
                       "pipeline": "step i runs on solver handle i % depth (own HBM state, own stream): the latency-bound "
                                   "tail of one batch overlaps the bulk of the next; depth <= steps // 2 (steady state)",
                       "mean_iters_per_solve": mean_iters},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
                         "traffic": None, "peak_source": how,
                         "kernel": "k_iterate (1 ADMM iteration / launch, all instances active)",
                         "algorithmic_bytes_per_launch": bpi * batch, "us_per_launch": us_iter,
                         "share_of_instance_iterations": work_share,
                         "frac_of_nominal_8TBs": achieved / 8000.0},
            "e2e": {"value": e2e_v, "unit": "IK solves/s", "h2d_bytes_per_step": int(q_h.numel() * 8 + b_h.numel() * 8),
                    "d2h_bytes_per_step": int(z_h[0].numel() * 8 + it_h[0].numel() * 4), "ms_per_step": ms_e2e / steps},
            "gpu_launches": int(launches),
            "extra": {
                "ms_per_solve_unpipelined": ms_single,
                "ms_per_solve_unpipelined_latency_schedule": None if ms_single_lane is None else {"lane_after": lat_after, "ms": ms_single_lane},
                "value_pipeline_4": world * batch * n4 / (ms_d4 * 1e-3),
                "converged_only_solves_per_s": value * stats["converged"] / batch,
                "iters_per_s": world * batch / (us_iter * 1e-6),
                "dense_kernel_with_launches_in_flight": {"handles": nsat, "us_per_launch_equivalent": us_sat, "iters_per_s": world * batch / (us_sat * 1e-6),
                                                         "algorithmic_frac_of_hbm": bpi * batch / (us_sat * 1e-6) / 1e9 / hbm},
                "instance_iterations_per_s_whole_solves": value * mean_iters,
                "lane_kernel": None if us_lane is None else {
                    "kernel": "k_iterate_lane (8 lanes per instance, state resident in shared memory, 20 iterations per launch)",
                    "us_per_batch_iteration": us_lane, "iters_per_s": world * batch / (us_lane * 1e-6),
                    "algorithmic_frac_of_hbm": bpi * batch / (us_lane * 1e-6) / 1e9 / hbm},
                "schedule": schedule,
                "solve_stats": {k: int(v) for k, v in stats.items()},
            },
        }
        traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(traffic_file):
            try:
                with open(traffic_file) as f:
                    tj = json.load(f)
                rec["roofline"]["traffic"] = tj.get(robot)
                rec["roofline"]["traffic_source"] = ("ncu --set full dram__bytes_read.sum + dram__bytes_write.sum of one k_iterate launch, "
                                                     "scripts/roundend_gpu.sh -> profiles/traffic.json; captured on kernel sources "
                                                     f"{tj.get('source_sha')}, this build {source_sha()}")
            except Exception:
                pass
    for S in solvers:
        S.close()
    del solvers, drivers, q_d, b_d, q_h, b_h, z_h, it_h
    torch.cuda.empty_cache()
    return rec, (model, pb, params)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="loik_b200", choices=["loik_b200", "reference"])
    ap.add_argument("--workload", default="all", choices=sorted(WORKLOADS) + ["all"],
                    help="all (default): Panda is the headline line, UR10 and Talos are reported under extra.workloads")
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

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    head_name = "panda" if args.workload == "all" else args.workload
    others = ["ur10", "talos"] if args.workload == "all" else []

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()  # clocks / throttle reasons are sampled from the warm-up to the end of the last timed region
    head, (model, pb, params) = run_workload(head_name, args, rank, world, local_rank, dev, headline=True)
    subs = {}
    for nm in others:
        rec, _ = run_workload(nm, args, rank, world, local_rank, dev, headline=False)
        if rank == 0:
            subs[nm] = rec
    clocks = sampler.stop() if rank == 0 else None

    if rank == 0:
        line = {
            "metric": METRIC, "value": head["value"], "unit": "IK solves/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": head["config"],
            "ms_per_solve_unpipelined": head["extra"]["ms_per_solve_unpipelined"],
            "iters_per_s": head["extra"]["iters_per_s"],
            "roofline": head["roofline"], "e2e": head["e2e"], "gpu_launches": head["gpu_launches"], "clocks": clocks,
            "solve_stats": head["extra"]["solve_stats"],
            "extra": dict(head["extra"], workloads=subs),
        }
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(model, pb, params, native_oracle_lib())
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Turn the ncu artefacts brought back in gpurun_out/ into the small text/JSON summaries kept under profiles/."""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.environ.get("PROFILES_OUT", os.path.join(ROOT, "profiles"))  # (on the GPU box: gpurun_out/profiles_out, the only directory that travels back)
TAG = sys.argv[1] if len(sys.argv) > 1 else "r2"

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
        "sm__cycles_elapsed.max", "lts__t_bytes.sum", "l1tex__t_bytes.sum"]


def launch_list():
    src = os.path.join(ROOT, "gpurun_out", f"launches_{TAG}.csv")
    if not os.path.exists(src):
        return
    rows = []
    with open(src) as f:
        lines = [ln for ln in f if ln.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ki, vi, mi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    for r in rd:
        if r[mi] == "gpu__time_duration.sum":
            rows.append((r[ki].split("(")[0], float(r[vi].replace(",", ""))))
    agg = collections.OrderedDict()
    for k, v in rows:
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(OUT, f"{TAG}_launches.txt"), "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised): compare SHARES, not absolutes\n")
        f.write(f"# command: python bench.py --workload panda --steps 2 --warmup 3 --pipeline 1 --no-cpu-baseline ({len(rows)} launches captured)\n")
        f.write(f"{'kernel':60s} {'launches':>9s} {'total_us':>12s} {'share':>7s} {'avg_us':>10s}\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k[:60]:60s} {n:9d} {t / 1e3:12.1f} {t / tot:7.3f} {t / n / 1e3:10.2f}\n")
    print(open(os.path.join(OUT, f"{TAG}_launches.txt")).read())


def full(name):
    rep = os.path.join(ROOT, "gpurun_out", f"prof_{TAG}_{name}.ncu-rep")
    if not os.path.exists(rep):
        return None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for k in KEEP:
            if k in hdr:
                d[k] = f"{r[hdr.index(k)]} {units[hdr.index(k)]}".strip()
        out.append(d)
    det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
    with open(os.path.join(OUT, f"{TAG}_ncu_full_{name}.txt"), "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on -k regex:k_iterate (scripts/roundend_gpu.sh; {name})\n")
        for d in out:
            f.write(json.dumps(d, indent=1) + "\n")
        f.write("\n# ---- details page of the first captured launch ----\n")
        keep = False
        n = 0
        for ln in det.splitlines():
            if "k_iterate" in ln and "Context" in ln:
                n += 1
                keep = n == 1
            if keep and ln.strip() and not ln.strip().startswith(("OPT", "INF")) and "----" not in ln:
                f.write(ln.rstrip() + "\n")
    return out


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    launch_list()
    traffic = {}
    full("lane_panda")
    full("lane_talos")
    for nm in ("panda", "talos", "ur10"):
        o = full(nm)
        if o:
            d = o[0]
            rd = float(d["dram__bytes_read.sum"].split()[0].replace(",", ""))
            wr = float(d["dram__bytes_write.sum"].split()[0].replace(",", ""))
            unit = d["dram__bytes_read.sum"].split()[1]
            scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[unit]
            traffic[nm] = (rd + wr) * scale
            print(nm, {k: d[k] for k in ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
                                        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
                                        "sm__warps_active.avg.pct_of_peak_sustained_active",
                                        "smsp__issue_active.avg.pct_of_peak_sustained_active",
                                        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
                                        "smsp__inst_executed.sum")})
    if traffic:
        sys.path.insert(0, ROOT)
        import bench
        traffic["source_sha"] = bench.source_sha()  # the kernel sources these captures were taken on (bench.py reports a mismatch)
        traffic["tag"] = TAG
        with open(os.path.join(OUT, "traffic.json"), "w") as f:
            json.dump(traffic, f)

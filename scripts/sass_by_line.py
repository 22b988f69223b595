#!/usr/bin/env python
"""Development aid: join an ncu SASS-level source page with nvdisasm line info and aggregate executed warp
instructions / stall samples per source line of one file.

  ncu -i rep.ncu-rep --page source --csv > sass.csv
  python scripts/sass_by_line.py sass.csv <lib.so> <kernel substring> <file substring> [per_unit_divisor]
"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

sass_csv, lib, kern, fsub = sys.argv[1:5]
div = float(sys.argv[5]) if len(sys.argv) > 5 else 1.0
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
# locate the kernel's section
start = next(i for i, ln in enumerate(dis) if ln.startswith("//---") and kern in ln and ".text." in ln)
# nvdisasm repeats the annotation block before every group of instructions it applies to; rebuild with reset semantics
addr2line = {}
cur, fresh = [], True
for ln in dis[start + 1:]:
    if ln.startswith("//---"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        if not fresh:
            cur, fresh = [], True
        cur.append((m.group(1), int(m.group(2))))
        continue
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(\S.*);", ln)
    if m:
        fresh = False
        pick = None
        for f, n in cur:  # innermost frame first: keep the innermost one inside the file of interest
            if fsub in f:
                pick = n
                break
        addr2line[int(m.group(1), 16)] = (pick, cur[0] if cur else None, m.group(2))
rows = list(csv.reader(open(sass_csv)))
hdr = next(r for r in rows if "Address" in r and "Source" in r)
ix = {n: i for i, n in enumerate(hdr)}
data = rows[rows.index(hdr) + 1:]
base = None
by = collections.defaultdict(lambda: [0, 0])
tot = [0, 0]
stall_cols = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
by_stall = collections.defaultdict(lambda: collections.Counter())
for r in data:
    if len(r) < len(hdr):
        continue
    try:
        a = int(r[ix["Address"]], 16) if not r[ix["Address"]].isdigit() else int(r[ix["Address"]])
    except ValueError:
        continue
    if base is None:
        base = a
    off = a - base
    n = int(float(r[ix["Instructions Executed"]] or 0))
    s = int(float(r[ix["# Samples"]] or 0))
    line = addr2line.get(off, (None, None, ""))[0]
    by[line][0] += n
    by[line][1] += s
    tot[0] += n
    tot[1] += s
    for c in stall_cols:
        v = r[ix[c]]
        if v:
            by_stall[line][c] += int(float(v))
import glob
cands = [p for p in glob.glob(os.path.join(os.path.dirname(os.path.abspath(lib)), "**", "*"), recursive=True) if fsub in os.path.basename(p)]
lines = open(cands[0]).read().splitlines() if cands else []
print(f"total: {tot[0] / div:.1f} instructions per unit, {tot[1]} samples")
for line, (n, s) in sorted(by.items(), key=lambda kv: (kv[0] is None, kv[0] or 0)):
    if n == 0 and s == 0:
        continue
    top = ",".join(f"{k[6:]}:{v}" for k, v in by_stall[line].most_common(3))
    text = lines[line - 1].strip()[:90] if line and line <= len(lines) else ""
    print(f"{str(line):>5} {n / div:9.1f} {100.0 * s / max(tot[1], 1):6.2f}%  {top:40s} | {text}")

#!/usr/bin/env python
"""Development aid: the instances of a full-size batch whose decision trace (iteration count, final mu, status) differs
between the CUDA path and the oracle -- how far apart the two actually are there."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from loik_b200 import problems, robots, solver as lk  # noqa: E402
from oracle import recursion  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "ur10"
B = int(sys.argv[2]) if len(sys.argv) > 2 else {"panda": 65536, "ur10": 262144, "talos": 16384}[name]
model = robots.get_robot(name)
pb = problems.random_batch(model, B, seed=0)
params = problems.bench_params(len(pb["ids"]))
ref = recursion.batch_solve(model, params, pb["q"], pb["H_ref"], pb["v_ref"], pb["ids"], pb["Ais"], pb["bis"], pb["lb"], pb["ub"], nthreads=os.cpu_count())
for la in (-1, 0):
    G = lk.make_solver(model, params, B)
    G.set_schedule(lane_after=la)
    G.SolveInit(pb["q"], pb["H_ref"], pb["v_ref"], pb["ids"], pb["Ais"], pb["bis"], pb["lb"], pb["ub"])
    G.Solve()
    it, mu, st, z = G.get_iter(), G.get_mu(), G.get_status(), G.z
    same = (it == ref["iters"]) & (mu == ref["mu"]) & ((st & 3) == (ref["status"] & 3))
    bad = np.nonzero(~same)[0]
    print(f"{name} x {B}, lane_after {la}: {len(bad)} diverged decision traces")
    dz = np.abs(z[bad] - ref["z"][bad]).max(axis=1) / np.maximum(1e-9, np.abs(ref["z"][bad]).max(axis=1))
    dit = it[bad] - ref["iters"][bad]
    print("  iteration-count differences:", dict(zip(*np.unique(dit, return_counts=True))))
    print("  status (gpu, oracle) pairs:", dict(zip(*np.unique(np.stack([st[bad] & 7, ref["status"][bad] & 7], 1), axis=0, return_counts=True))) if False else "")
    print("  oracle iterations of the diverged: median", np.median(ref["iters"][bad]), "max", ref["iters"][bad].max(), "; share at max_iter:", np.mean(ref["iters"][bad] >= params["max_iter"] - 1))
    print("  rel-inf(z) between the two on the diverged: median %.2e, 90%% %.2e, max %.2e" % (np.median(dz), np.quantile(dz, 0.9), dz.max()))
    for i in bad[:8]:
        print(f"    #{i}: gpu it {it[i]} mu {mu[i]:g} st {st[i]} | oracle it {ref['iters'][i]} mu {ref['mu'][i]:g} st {ref['status'][i]} | rel-inf z {np.abs(z[i]-ref['z'][i]).max()/max(1e-9,np.abs(ref['z'][i]).max()):.2e}")
    G.close()

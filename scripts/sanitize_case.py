#!/usr/bin/env python
"""Small end-to-end case for compute-sanitizer (memcheck / racecheck / initcheck / synccheck)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from loik_b200 import problems, robots, solver as lk  # noqa: E402

for name, B in (("panda9", 100), ("talos", 70)):
    model = robots.get_robot(name)
    pb = problems.random_batch(model, B, seed=0)
    S = lk.make_solver(model, problems.bench_params(len(pb["ids"]), max_iter=40), B)
    S.SolveInit(pb["q"], pb["H_ref"], pb["v_ref"], pb["ids"], pb["Ais"], pb["bis"], pb["lb"], pb["ub"])
    S.Solve()
    z = S.z
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        S.Solve()          # graph path
        side.synchronize()
        assert np.array_equal(S.z, z)
    S.Integrate(0.01)
    S.Solve(None, int(pb["ids"][0]), pb["Ais"][0], pb["bis"][:, 0])
    S.set_debug(True)
    S.ResetRecursion(); S.StepBackward(); S.StepForward(); S.StepResidual()
    _ = S.His, S.norms(), S.liMi, S.stats()
    S.close()
    # the lane-parallel shared-memory kernel: whole solves, the hand-over from a packed arena, four groups per instance,
    # per-instance task matrices, the backward-pass workspace brought home
    rng = np.random.default_rng(0)
    for sched in (dict(lane_after=0), dict(lane_after=6), dict(lane_after=0, lane_groups_per_instance=4)):
        S = lk.make_solver(model, problems.bench_params(len(pb["ids"]), max_iter=40), B)
        S.set_schedule(**sched)
        S.set_keep_workspace(True)
        A = np.eye(6)[None, None] + 0.1 * rng.standard_normal((B, len(pb["ids"]), 6, 6))
        S.SolveInit(pb["q"], pb["H_ref"], pb["v_ref"], pb["ids"], A, pb["bis"], pb["lb"], pb["ub"])
        S.Solve()
        _ = S.His, S.z
        S.IterateFixed(3)
        S.close()
    print(name, "ok")

# multi-DoF trees (free-flyer / spherical / SphericalZYX ...), per-instance references (v_ref and H_ref), the workspace brought home,
# the target-only task update, the solver log
for seed, zyx in ((3, 0.0), (5, 0.5)):
    model = robots.random_tree(11, 200 + seed, multidof=0.4, zyx=zyx)
    B = 70
    rng = np.random.default_rng(seed)
    ids = np.array(sorted(rng.choice(np.arange(1, model.nj), size=2, replace=False)), np.int32)
    pb = dict(q=model.normalize(rng.uniform(model.q_min, model.q_max, size=(B, model.nq))), H_ref=np.eye(6), v_ref=np.zeros(6), ids=ids,
              Ais=np.tile(np.eye(6), (2, 1, 1)), bis=rng.uniform(-0.5, 0.5, size=(B, 2, 6)), lb=-model.v_max, ub=model.v_max)
    params = dict(problems.FIXTURE_PARAMS, max_iter=40, num_eq_c=2, tol_abs=1e-3, tol_rel=1e-3)
    S = lk.make_solver(model, params, B)
    S.set_keep_workspace(True)
    S.SolveInit(pb["q"], pb["H_ref"], pb["v_ref"], pb["ids"], pb["Ais"], pb["bis"], pb["lb"], pb["ub"])
    W = rng.normal(size=(B, model.nj, 6, 6))
    S.UpdateReferences(np.eye(6)[None, None] + 0.05 * (W + W.transpose(0, 1, 3, 2)), 0.05 * rng.normal(size=(B, model.nj, 6)))
    S.Solve(); S.Solve()
    _ = S.His, S.z
    S.Integrate(0.01)
    S.Solve(None, int(ids[0]), None, pb["bis"][:, 0])
    S.close()
    S = lk.make_solver(model, dict(params, logging=True), B)
    S.SolveInit(pb["q"], pb["H_ref"], pb["v_ref"], pb["ids"], pb["Ais"], pb["bis"], pb["lb"], pb["ub"])
    S.UpdateReferences(np.tile(np.eye(6), (model.nj, 1, 1)), 0.05 * rng.normal(size=(B, model.nj, 6)))
    S.Solve()
    _ = S.history()
    S.close()
    print("tree", seed, "ok")

"""The lane-parallel shared-memory kernel (``loik_b200/csrc/loik_lane.cuh``) vs oracle B, through the C ABI.

It is a second CUDA implementation of the same iteration (8 lanes per instance instead of one thread), selected by
``loik_set_schedule(lane_after = N)``: N = 0 runs whole solves in it, N > 0 hands it the instances still active after
N sweeps of the tile kernels.  Tests: (1) iteration by iteration against the oracle driven method by method (the
reference's own test shape, ``tests/loik-loid.cpp:305-556``), every field incl. the backward-pass workspace;
(2) full solves for several switch points incl. decision traces; (3) the hand-over from a packed arena;
(4) warm-started tailored solves, per-instance bounds, ragged batches.
"""
import numpy as np
import pytest

from loik_b200 import problems, robots
from tests.helpers import check_abs_or_rel, ctor_kwargs, instance, rel_inf

pytestmark = pytest.mark.gpu

STEP_TOL = 1e-10


def _gpu(model, params, batch, **schedule):
    from loik_b200 import solver
    G = solver.make_solver(model, params, batch)
    sc = G.get_schedule()
    assert sc["lane_available"] == 1, "the lane-parallel kernel must be available for trees of 1-DoF joints"
    G.set_schedule(**schedule)
    return G


def _oracle(model, params):
    from oracle import recursion
    return recursion.FirstOrderLoikOptimized(model, **ctor_kwargs(params))


def _oracle_batch(model, params, pb, nthreads=8, **kw):
    from oracle import recursion
    return recursion.batch_solve(model, params, pb["q"], pb["H_ref"], pb["v_ref"], pb["ids"], pb["Ais"], pb["bis"],
                                 pb["lb"], pb["ub"], nthreads=nthreads, **kw)


def _solve_init(G, pb):
    G.SolveInit(pb["q"], pb["H_ref"], pb["v_ref"], pb["ids"], pb["Ais"], pb["bis"], pb["lb"], pb["ub"])


def _oracle_iteration(o, it, fixed=True):
    """One iteration of the main loop of Solve() (loik-loid-optimized.hpp:377-454) method by method."""
    o.UpdatePrev(); o.ResetInfNorms(); o.FwdPass1(); o.BwdPassOptimizedVisitor()
    o.FwdPass2OptimizedVisitor(); o.BoxProj(); o.DualUpdate()
    o.ComputeResiduals(); o.CheckConvergence()
    if it > 1:
        o.CheckFeasibility()
    o.UpdateMu()


@pytest.mark.parametrize("name,gpi", [("panda", 1), ("ur10", 1), ("talos", 1), ("panda9", 1), ("ur10c", 1),
                                      ("talos", 4), ("panda9", 4), ("ur10", 4)])
def test_lane_iterations_vs_oracle_method_by_method(name, gpi):
    """k iterations in fixed-iteration mode (stopping disabled) == the oracle's public methods called k times.
    gpi = 1: one 8-lane group per instance sweeps the whole tree; gpi = 4: the four groups of a warp sweep different
    chains of ONE instance level by level (Talos: 5 chains, panda9: the two fingers; ur10: the degenerate single chain)."""
    model = robots.get_robot(name)
    B = 24
    pb = problems.random_batch(model, B, seed=11)
    params = problems.bench_params(len(pb["ids"]))
    G = _gpu(model, params, B, lane_after=0, lane_groups_per_instance=gpi)
    assert G.get_schedule()["lane_groups_chosen"] == gpi
    G.set_keep_workspace(True)
    _solve_init(G, pb)
    O = []
    for i in range(B):
        o = _oracle(model, params)
        o.SolveInit(*instance(pb, i))
        o.ResetSolver()
        O.append(o)
    for it in range(1, 5):
        for o in O:
            _oracle_iteration(o, it)
        G.IterateFixed(it, reset=True)  # from scratch: `it` iterations inside ONE launch, state never leaves the SM
        tag = f"{name} it{it}"
        fields = dict(nu=G.nu, v=G.vis, f=G.fis, z=G.z, w=G.w, y=G.yis, Aty=G.Aty, H=G.His, p=G.pis, UD=G.UDinv, Di=G.Dinv,
                      r=G.r, F=G.fis_diff_plus_Aty, T=G.Stf_plus_w, res=G.get(18), mu=G.get_mu())
        assert (G.get_iter() == it).all()
        for i, o in enumerate(O):
            fscale = max(1.0, np.abs(o.fis).max())
            check_abs_or_rel(fields["H"][i], o.His[1:], STEP_TOL, tag + " His")
            check_abs_or_rel(fields["p"][i], o.pis[1:], STEP_TOL, tag + " pis")
            check_abs_or_rel(fields["UD"][i], o.UDinv[1:], STEP_TOL, tag + " UDinv")
            check_abs_or_rel(fields["Di"][i], o.Dinv[1:], STEP_TOL, tag + " Dinv")
            check_abs_or_rel(fields["r"][i], o.r, STEP_TOL, tag + " r")
            check_abs_or_rel(fields["nu"][i], o.nu, STEP_TOL, tag + " nu")
            check_abs_or_rel(fields["v"][i], o.vis[1:], STEP_TOL, tag + " vis")
            check_abs_or_rel(fields["f"][i], o.fis[1:], STEP_TOL, tag + " fis")
            check_abs_or_rel(fields["z"][i], o.z, STEP_TOL, tag + " z")
            check_abs_or_rel(fields["w"][i], o.w, STEP_TOL, tag + " w")
            check_abs_or_rel(fields["y"][i], o.yis, STEP_TOL, tag + " yis")
            check_abs_or_rel(fields["Aty"][i], o.Aty, STEP_TOL, tag + " Aty")
            assert np.abs(fields["F"][i] - o.fis_diff_plus_Aty[1:]).max() < 1e-11 * fscale, tag + " fis_diff_plus_Aty"
            assert np.abs(fields["T"][i] - o.Stf_plus_w).max() < 1e-11 * fscale, tag + " Stf_plus_w"
            check_abs_or_rel(fields["res"][i, 0], o.get_primal_residual(), STEP_TOL, tag + " primal_residual")
            assert abs(fields["res"][i, 1] - o.get_dual_residual()) < 1e-11 * fscale, tag + " dual_residual"
            check_abs_or_rel(fields["res"][i, 2], o.get_tol_primal(), STEP_TOL, tag + " tol_primal")
            check_abs_or_rel(fields["res"][i, 3], o.get_tol_dual(), 1e-9, tag + " tol_dual")
            assert fields["mu"][i] == o.get_mu(), tag + " mu"
    G.close()


def _compare_solves(model, params, pb, what, schedule, tol=1e-6, max_diverged_frac=0.0, state_tol=1e-8):
    B = pb["q"].shape[0]
    G = _gpu(model, params, B, **schedule)
    _solve_init(G, pb)
    G.Solve()
    ref = _oracle_batch(model, params, pb)
    it, mu, st = G.get_iter(), G.get_mu(), G.get_status()
    same = (it == ref["iters"]) & (mu == ref["mu"]) & ((st & 3) == (ref["status"] & 3))
    diverged = int((~same).sum())
    z, nu, w, y = G.z, G.nu, G.w, G.yis
    worst = 0.0
    for i in np.nonzero(same)[0]:
        worst = max(worst, rel_inf(z[i], ref["z"][i]), rel_inf(nu[i], ref["nu"][i]), rel_inf(w[i], ref["w"][i]),
                    rel_inf(y[i], ref["y"][i]))
    print(f"[{what} {schedule}] B={B} diverged decision traces: {diverged}/{B}; worst rel-inf over z,nu,w,y: {worst:.3e}; "
          f"mean iters {it.mean():.2f}")
    assert worst < tol, f"{what}: rel-inf {worst:.3e} >= {tol}"
    assert diverged <= max_diverged_frac * B, f"{what}: {diverged} instances with a different iteration count / mu / status"
    s = G.stats()
    assert s["total_iters"] == int(it.sum())
    # every other state row against the tile kernels' result of the same solve (those are pinned to the oracle field by
    # field in test_gpu_parity.py): catches a mis-routed row that z, nu, w, y would not show
    out = dict(v=G.vis, f=G.fis, F=G.fis_diff_plus_Aty, T=G.Stf_plus_w, Aty=G.Aty, res=G.get(18))
    G.set_schedule(lane_after=-1)
    G.Solve()
    same2 = (G.get_iter() == it) & (G.get_mu() == mu)
    assert same2.mean() >= 1.0 - max_diverged_frac
    ref2 = dict(v=G.vis, f=G.fis, F=G.fis_diff_plus_Aty, T=G.Stf_plus_w, Aty=G.Aty, res=G.get(18))
    for k in out:
        a, b = out[k][same2], ref2[k][same2]
        scale = max(1.0, np.abs(b).max())
        assert np.abs(a - b).max() < state_tol * scale, f"{what}: field {k} differs between the lane and the tile kernels"
    G.close()


@pytest.mark.parametrize("name,B,gpi", [("panda", 4096, 1), ("ur10", 4096, 1), ("talos", 1024, 1), ("panda9", 1024, 1), ("ur10c", 2048, 1),
                                        ("talos", 1024, 4), ("panda9", 1024, 4)])
@pytest.mark.parametrize("lane_after", [0, 3, 7, 40])
def test_lane_full_solves(name, B, gpi, lane_after):
    """Full solves (max_iter 200): the whole solve in the lane kernel (0), the hand-over from the home arena after the
    dense sweeps (3: list mode, no origin map) and from a packed scratch arena (7, 40: list + origin map); one group
    per instance and, for the branching trees, four groups per instance."""
    model = robots.get_robot(name)
    pb = problems.random_batch(model, B, seed=0)
    _compare_solves(model, problems.bench_params(len(pb["ids"])), pb, name, dict(lane_after=lane_after, lane_groups_per_instance=gpi),
                    max_diverged_frac=0.002)


@pytest.mark.parametrize("seed", [0, 1, 2, 3, 4, 5])
def test_lane_random_trees(seed):
    """Random trees of every 1-DoF joint type (aligned / unaligned / unbounded revolute, prismatic), two tasks, random
    branching: both geometries of the lane kernel against the oracle, iteration by iteration decisions included."""
    model = robots.random_tree(12 + 3 * seed, seed=seed, continuous=0.3 if seed % 2 else 0.0)
    tasks = [model.nj - 1, max(1, model.nj // 2)]
    pb = problems.random_batch(model, 256, seed=seed, task_joints=tasks)
    params = dict(problems.FIXTURE_PARAMS, max_iter=200, num_eq_c=2)
    for gpi in (1, 4):
        _compare_solves(model, params, pb, f"tree{seed} gpi{gpi}", dict(lane_after=0, lane_groups_per_instance=gpi), max_diverged_frac=0.004)


@pytest.mark.parametrize("B", [1, 3, 4, 5, 31, 33, 100])
def test_lane_ragged_batches(B):
    model = robots.panda()
    pb = problems.random_batch(model, B, seed=3)
    _compare_solves(model, problems.bench_params(1), pb, f"panda B={B}", dict(lane_after=0), max_diverged_frac=0.0)


@pytest.mark.parametrize("max_iter", [1, 2, 3, 5])
def test_lane_tiny_iteration_budgets(max_iter):
    """max_iter - 1 iterations run (hpp:377); max_iter < 2: none."""
    model = robots.ur10()
    pb = problems.random_batch(model, 128, seed=4)
    params = problems.bench_params(1, max_iter=max_iter)
    for la in (0, 1):
        G = _gpu(model, params, 128, lane_after=la)
        _solve_init(G, pb)
        G.Solve()
        ref = _oracle_batch(model, params, pb)
        np.testing.assert_array_equal(G.get_iter(), ref["iters"])
        assert rel_inf(G.z, ref["z"]) < 1e-9
        G.close()


def test_lane_infeasible_targets_take_the_tail_path():
    """Targets far outside the velocity limits: most instances leave through InfeasibilityTailSolve (hpp:271-319)."""
    model = robots.panda()
    B = 1024
    pb = problems.random_batch(model, B, seed=2, b_scale=3.0)
    params = problems.bench_params(1)
    G = _gpu(model, params, B, lane_after=0)
    _solve_init(G, pb)
    G.Solve()
    ref = _oracle_batch(model, params, pb)
    st = G.get_status()
    assert ((st >> 1) & 1).mean() > 0.5
    same = (G.get_iter() == ref["iters"]) & (G.get_mu() == ref["mu"]) & ((st & 3) == (ref["status"] & 3))
    assert same.mean() >= 0.998
    assert rel_inf(G.z[same], ref["z"][same]) < 1e-6
    G.close()


def test_lane_per_instance_bounds_warm_start_and_tailored_solve():
    """Per-instance lb / ub (rows of the instance record instead of the parameter block) and the tailored, warm-started
    Solve(q, c_id, Ai, bi) (hpp:596-695) with the whole solve in the lane kernel."""
    from oracle import recursion
    model = robots.panda()
    B = 96
    rng = np.random.default_rng(5)
    pb = problems.random_batch(model, B, seed=8)
    pb["ub"] = np.tile(pb["ub"], (B, 1)) * rng.uniform(0.5, 1.5, (B, model.nv))
    pb["lb"] = -pb["ub"] * rng.uniform(0.5, 1.0, (B, model.nv))
    params = dict(problems.bench_params(1), warm_start=True)
    G = _gpu(model, params, B, lane_after=0)
    _solve_init(G, pb)
    G.Solve()
    b2 = pb["bis"][:, 0] + 0.05 * rng.standard_normal((B, 6))
    A2 = np.eye(6) + 0.05 * rng.standard_normal((6, 6))
    G.Solve(pb["q"], int(pb["ids"][0]), A2, b2)
    z, it = G.z, G.get_iter()
    for i in range(B):
        o = recursion.FirstOrderLoikOptimized(model, **ctor_kwargs(params))
        o.SolveInit(pb["q"][i], pb["H_ref"], pb["v_ref"], pb["ids"], pb["Ais"], pb["bis"][i], pb["lb"][i], pb["ub"][i])
        o.Solve()
        o.Solve(pb["q"][i], int(pb["ids"][0]), A2, b2[i])
        assert it[i] == o.get_iter(), f"instance {i}: iteration count"
        assert rel_inf(z[i], o.z) < 1e-6
    G.close()


def test_lane_per_joint_references():
    """UpdateReferences (ik-id-description-optimized.hpp:103-121): a different symmetric H_ref / v_ref per joint travels
    to the lane kernel through the per-CTA constants in shared memory."""
    from oracle import recursion
    model = robots.ur10()
    B = 64
    rng = np.random.default_rng(9)
    pb = problems.random_batch(model, B, seed=9)
    params = problems.bench_params(1)
    H_refs = np.zeros((model.nj, 6, 6))
    for j in range(model.nj):
        a = rng.standard_normal((6, 6))
        H_refs[j] = np.eye(6) + 0.1 * (a @ a.T)
    v_refs = 0.1 * rng.standard_normal((model.nj, 6))
    G = _gpu(model, params, B, lane_after=0)
    _solve_init(G, pb)
    G.UpdateReferences(H_refs, v_refs)
    G.Solve()
    z, it = G.z, G.get_iter()
    for i in range(B):
        o = recursion.FirstOrderLoikOptimized(model, **ctor_kwargs(params))
        o.SolveInit(*instance(pb, i))
        o.UpdateReferences(H_refs, v_refs)
        o.Solve()
        assert it[i] == o.get_iter()
        assert rel_inf(z[i], o.z) < 1e-6
    G.close()

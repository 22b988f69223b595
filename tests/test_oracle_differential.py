"""Oracle B (recursion, C) == oracle A (dense, numpy), step by step and end to end.

Mirrors the differential structure of the reference's own tests, which pin the optimized solver against
the non-optimized one on one fixture (``/root/reference/tests/loik-loid.cpp``):
  test_1st_order_loik_optimized_correctness_component_wise  :305-556
  test_1st_order_loik_optimized_correctness                 :559-671
  test_1st_order_loik_optimized_reset_component_wise        :674-865
  test_1st_order_loik_optimized_reset                       :868-984
  test_loik_solve_split                                     :261-303
This is what pins the oracle (PARITY UNPINNED by known answers: the reference has no golden vectors).
"""
import os

import numpy as np
import pytest

from loik_b200 import problems, robots
from oracle import dense, recursion
from tests.helpers import check_abs_or_rel, ctor_kwargs, instance, prob_args

TOL = 1e-10


def make_pair(model, params):
    kw = ctor_kwargs(params)
    return dense.FirstOrderLoik(model, **kw), recursion.FirstOrderLoikOptimized(model, **kw)


def compare_state(A, B, model, what, fis_tol=TOL):
    c_ids = [int(c) for c in A.task_ids]
    check_abs_or_rel(B.nu, A.nu, TOL, what + " nu")
    check_abs_or_rel(B.z, A.z, TOL, what + " z")
    check_abs_or_rel(B.w, A.w, TOL, what + " w")
    check_abs_or_rel(B.vis[1:], A.vis[1:], TOL, what + " vis")
    check_abs_or_rel(B.fis[1:], A.fis[1:], fis_tol, what + " fis")
    for k, c in enumerate(c_ids):
        check_abs_or_rel(B.yis[k], A.yis[c], TOL, what + f" yis[{k}]")


def step_once_and_compare(A, B, model, it, what):
    """One ADMM iteration, comparing after every public step (tests/loik-loid.cpp:340-478)."""
    A.iter = it
    A.UpdatePrev()
    B.UpdatePrev()
    B.ResetInfNorms()
    A.FwdPass1()
    B.FwdPass1()
    check_abs_or_rel(B.His[1:], A.His[1:], TOL, what + " FwdPass1 His")
    check_abs_or_rel(B.His_aba[1:], B.His[1:], TOL, what + " FwdPass1 His_aba")
    check_abs_or_rel(B.pis[1:], A.pis[1:], TOL, what + " FwdPass1 pis")
    for i in range(1, model.nj):
        iv, n = model.idx_v(i), model.nv_joint(i)
        check_abs_or_rel(B.R[iv:iv + n], np.diag(A.Ris[i]), TOL, what + " R")
        check_abs_or_rel(B.r[iv:iv + n], A.ris[i], TOL, what + " r")
    A.BwdPass()
    B.BwdPassOptimizedVisitor()
    check_abs_or_rel(B.His[1:], A.His[1:], TOL, what + " BwdPass His")
    check_abs_or_rel(B.pis[1:], A.pis[1:], TOL, what + " BwdPass pis")
    Dfull = B.Dinv_full
    for i in range(1, model.nj):  # D^-1 and the projector agree with calc_aba's Dinv / UDinv
        n = model.nv_joint(i)
        check_abs_or_rel(Dfull[i][:n, :n], A.Di_invs[i], TOL, what + " Dinv")
    A.FwdPass2()
    B.FwdPass2OptimizedVisitor()
    check_abs_or_rel(B.nu, A.nu, TOL, what + " FwdPass2 nu")
    check_abs_or_rel(B.vis[1:], A.vis[1:], TOL, what + " FwdPass2 vis")
    check_abs_or_rel(B.fis[1:], A.fis[1:], TOL, what + " FwdPass2 fis")
    A.BoxProj()
    B.BoxProj()
    check_abs_or_rel(B.z, A.z, TOL, what + " BoxProj z")
    A.DualUpdate()
    B.DualUpdate()
    compare_state(A, B, model, what + " DualUpdate")
    A.UpdateQPADMMSolveLoopUtility()
    A.ComputeResiduals()
    B.ComputeResiduals()
    check_abs_or_rel(B.get_primal_residual_vec(), A.primal_residual_vec, TOL, what + " primal_residual_vec")
    # the dense dual residual is P x + q + A^T y (loik-loid.hxx:280); entries are differences of O(|f|) terms
    scale = max(1.0, np.abs(A.fis).max())
    assert np.abs(B.get_dual_residual_vec() - A.dual_residual_vec).max() < 1e-12 * scale * 10, what + " dual_residual_vec"
    check_abs_or_rel(B.get_primal_residual(), A.primal_residual, TOL, what + " primal_residual")
    assert abs(B.get_dual_residual() - A.dual_residual) < 1e-11 * scale, what + " dual_residual"
    A.CheckConvergence()
    B.CheckConvergence()
    assert A.tol_primal != 0.0 and B.get_tol_primal() != 0.0
    assert A.tol_dual != 0.0 and B.get_tol_dual() != 0.0
    # tol_dual: both are tol_abs + tol_rel * max(|Px|, |A^T y|, |q|)
    check_abs_or_rel(B.get_tol_dual(), A.tol_dual, 1e-9, what + " tol_dual")
    assert A.converged == B.get_convergence_status(), what + " converged"
    if it > 1:
        A.CheckFeasibility()
        B.CheckFeasibility()
        check_abs_or_rel(B.get_delta_y_qp_inf_norm(), A.delta_y_qp_inf_norm, TOL, what + " delta_y_qp")
        assert abs(B.get_A_qp_T_delta_y_qp_inf_norm() - A.A_qp_T_delta_y_qp_inf_norm) < 1e-11 * scale
        check_abs_or_rel(B.get_ub_qp_T_delta_y_qp_plus(), A.ub_qp_T_delta_y_qp_plus, 1e-9, what + " ub^T dy+")
        check_abs_or_rel(B.get_lb_qp_T_delta_y_qp_minus(), A.lb_qp_T_delta_y_qp_minus, 1e-9, what + " lb^T dy-")
        assert A.primal_infeasibility_cond_1 == B.get_primal_infeasibility_cond_1()
        assert A.primal_infeasibility_cond_2 == B.get_primal_infeasibility_cond_2()
        assert A.primal_infeasible == B.get_primal_infeasibility_status()
        check_abs_or_rel(B.get_delta_x_qp_inf_norm(), np.abs(A.delta_x_qp).max(), TOL, what + " delta_x_qp")
        check_abs_or_rel(B.get_delta_z_qp_inf_norm(), np.abs(A.delta_z_qp).max(), TOL, what + " delta_z_qp")
    A.UpdateMu()
    B.UpdateMu()
    assert A.mu == B.get_mu(), what + " mu"


ROBOT_CASES = [("talos", 1.0), ("panda", 1.0), ("panda9", 2.0), ("ur10", 1.0), ("talos_ff", 1.0), ("ur10c", 1.0)]


@pytest.mark.parametrize("name,bound", ROBOT_CASES)
def test_optimized_correctness_component_wise(name, bound):
    """tests/loik-loid.cpp:305-556 -- bounds +-1, every public step compared."""
    model = robots.get_robot(name)
    params = dict(problems.FIXTURE_PARAMS, max_iter=2)
    pr = problems.fixture_problem(model, bound)
    A, B = make_pair(model, params)
    A.SolveInit(*prob_args(pr))
    B.SolveInit(*prob_args(pr))
    for it in (1, 2, 3, 4):
        step_once_and_compare(A, B, model, it, f"{name} it{it}")


@pytest.mark.parametrize("seed,continuous,multidof,zyx", [(0, 0.0, 0.0, 0.0), (1, 0.0, 0.0, 0.0), (2, 0.0, 0.0, 0.0), (3, 0.0, 0.0, 0.0), (4, 0.5, 0.0, 0.0),
                                                          (5, 0.5, 0.0, 0.0), (6, 1.0, 0.0, 0.0), (7, 0.0, 0.3, 0.0), (8, 0.3, 0.3, 0.0), (9, 0.0, 0.6, 0.0),
                                                          (10, 0.0, 0.0, 0.4), (11, 0.3, 0.3, 0.4), (12, 0.0, 0.0, 1.0)])
def test_component_wise_random_trees(seed, continuous, multidof, zyx):
    """Same, on seeded random trees with every joint type (aligned/unaligned, revolute/prismatic/unbounded revolute,
    and -- oracles only so far -- spherical / translation joints and free-flyers anywhere in the tree), branching."""
    model = robots.random_tree(12, seed, continuous=continuous, multidof=multidof, zyx=zyx)  # (zyx: JointModelSphericalZYX, S depends on q)
    rng = np.random.default_rng(100 + seed)
    params = dict(problems.FIXTURE_PARAMS, max_iter=2, num_eq_c=2)
    ids = np.array(sorted(rng.choice(np.arange(1, model.nj), size=2, replace=False)), np.int32)
    As = np.stack([np.eye(6) + 0.3 * rng.normal(size=(6, 6)) for _ in range(2)])
    Hs = rng.normal(size=(6, 6))
    pr = dict(q=model.normalize(rng.uniform(model.q_min, model.q_max)), H_ref=np.eye(6) + 0.1 * (Hs + Hs.T),
              v_ref=0.1 * rng.normal(size=6), ids=ids, Ais=As, bis=rng.uniform(-0.5, 0.5, size=(2, 6)), lb=-model.v_max,
              ub=model.v_max)
    A, B = make_pair(model, params)
    A.SolveInit(*prob_args(pr))
    B.SolveInit(*prob_args(pr))
    for it in (1, 2, 3):
        step_once_and_compare(A, B, model, it, f"tree{seed} it{it}")


@pytest.mark.parametrize("name,bound", [("talos", 2.0), ("panda", 2.0), ("ur10", 2.0), ("talos_ff", 2.0), ("ur10c", 2.0)])
def test_optimized_correctness_end_to_end(name, bound):
    """tests/loik-loid.cpp:559-671 -- max_iter = 8, bounds +-2: SolveInit + Solve() == dense Solve(args)."""
    model = robots.get_robot(name)
    params = dict(problems.FIXTURE_PARAMS, max_iter=8)
    pr = problems.fixture_problem(model, bound)
    A, B = make_pair(model, params)
    A.Solve(*prob_args(pr))
    B.SolveInit(*prob_args(pr))
    B.Solve()
    compare_state(A, B, model, name)
    assert A.iter == B.get_iter()
    assert A.mu == B.get_mu()
    assert A.converged == B.get_convergence_status()
    assert A.primal_infeasible == B.get_primal_infeasibility_status()
    check_abs_or_rel(B.get_primal_residual(), A.primal_residual, TOL, "primal_residual")


@pytest.mark.parametrize("name", ["panda", "ur10", "talos", "talos_ff", "ur10c"])
def test_end_to_end_random_instances(name):
    """Full solves (max_iter = 200) on seeded random instances of the BASELINE configs: same iterates, same
    iteration count, same mu history, same flags."""
    model = robots.get_robot(name)
    n = 6 if name.startswith("talos") else 12
    pb = problems.random_batch(model, n, seed=7)
    params = problems.bench_params(len(pb["ids"]), max_iter=60)
    for i in range(n):
        A, B = make_pair(model, params)
        A.Solve(*instance(pb, i))
        B.Solve(*instance(pb, i))
        assert A.iter == B.get_iter(), f"{name}[{i}] iter"
        np.testing.assert_array_equal(np.array(A.hist_mu), B.hist_mu)
        assert A.converged == B.get_convergence_status()
        assert A.primal_infeasible == B.get_primal_infeasibility_status()
        # a long solve amplifies the rounding differences of two different formulations
        compare_state_loose(A, B, f"{name}[{i}]")


def compare_state_loose(A, B, what, tol=1e-7):
    for nm in ("nu", "z", "w"):
        a, b = getattr(A, nm), getattr(B, nm)
        assert np.abs(a - b).max() <= tol * max(1.0, np.abs(a).max()), f"{what} {nm}"
    assert np.abs(A.vis[1:] - B.vis[1:]).max() <= tol * max(1.0, np.abs(A.vis).max()), what + " vis"


def test_solve_split():
    """tests/loik-loid.cpp:261-303 -- Solve(args) == SolveInit(args) + Solve(); bounds +-5, max_iter = 200."""
    model = robots.talos()
    params = dict(problems.FIXTURE_PARAMS, max_iter=200)
    pr = problems.fixture_problem(model, 5.0)
    kw = ctor_kwargs(params)
    B1 = recursion.FirstOrderLoikOptimized(model, **kw)
    B2 = recursion.FirstOrderLoikOptimized(model, **kw)
    B1.Solve(*prob_args(pr))
    B2.SolveInit(*prob_args(pr))
    B2.Solve()
    for nm in ("nu", "z", "w", "vis", "fis", "yis"):
        np.testing.assert_array_equal(getattr(B1, nm), getattr(B2, nm))
    assert B1.get_iter() == B2.get_iter()


def test_reset_repeated_solves():
    """tests/loik-loid.cpp:868-984 -- 5 repeated Solve(args) on the same objects == ground truth incl. iteration count."""
    model = robots.talos()
    params = dict(problems.FIXTURE_PARAMS, max_iter=100)
    pr = problems.fixture_problem(model, 2.0)
    A, B = make_pair(model, params)
    for rep in range(5):
        A.Solve(*prob_args(pr))
        B.Solve(*prob_args(pr))
        compare_state(A, B, model, f"rep{rep}", fis_tol=1e-9)
        assert A.iter == B.get_iter()
        assert A.mu == B.get_mu()


def test_reset_component_wise():
    """tests/loik-loid.cpp:674-865 -- a second Solve() after ResetRecursion/ResetSolver reproduces the first."""
    model = robots.talos()
    params = dict(problems.FIXTURE_PARAMS, max_iter=100)
    pr = problems.fixture_problem(model, 1.5)
    B = recursion.FirstOrderLoikOptimized(model, **ctor_kwargs(params))
    B.SolveInit(*prob_args(pr))
    B.Solve()
    first = {nm: getattr(B, nm).copy() for nm in ("nu", "z", "w", "vis", "fis", "yis")}
    it1 = B.get_iter()
    B.Solve()
    for nm, a in first.items():
        np.testing.assert_array_equal(getattr(B, nm), a)
    assert B.get_iter() == it1


def test_tailored_solve_matches_full_solve():
    """Solve(q, c_id, Ai, bi) (hpp:596-695) after a SolveInit == Solve(args) with the updated constraint."""
    model = robots.panda()
    params = problems.bench_params(1, max_iter=50)
    pb = problems.random_batch(model, 2, seed=3)
    kw = ctor_kwargs(params)
    B1 = recursion.FirstOrderLoikOptimized(model, **kw)
    B2 = recursion.FirstOrderLoikOptimized(model, **kw)
    B1.SolveInit(*instance(pb, 0))
    B1.Solve(pb["q"][1], int(pb["ids"][0]), pb["Ais"][0], pb["bis"][1][0])
    args = instance(pb, 1)
    B2.Solve(*args)
    # bis_inf_norm only grows under UpdateEqConstraint (quirk 9) so tolerances may differ; iterates agree when
    # the first problem's |b| is not larger
    if np.abs(pb["bis"][0]).max() <= np.abs(pb["bis"][1]).max():
        np.testing.assert_allclose(B1.z, B2.z, rtol=0, atol=1e-12)
        assert B1.get_iter() == B2.get_iter()
    assert B1.scalar("bis_inf_norm") == max(np.abs(pb["bis"][0]).max(), np.abs(pb["bis"][1]).max())


def test_error_paths():
    model = robots.panda()
    kw = ctor_kwargs(problems.FIXTURE_PARAMS)
    with pytest.raises(RuntimeError):  # eq_c_dim != 6 (ik-id-description-optimized.hpp:41-44)
        recursion.FirstOrderLoikOptimized(model, **dict(kw, eq_c_dim=3))
    B = recursion.FirstOrderLoikOptimized(model, **kw)
    pr = problems.fixture_problem(model)
    bad = dict(pr, ids=np.array([3, 5], np.int32), Ais=np.tile(np.eye(6), (2, 1, 1)), bis=np.zeros((2, 6)))
    with pytest.raises(RuntimeError):  # number of constraints != num_eq_c (:142-145)
        B.SolveInit(*prob_args(bad))
    with pytest.raises(RuntimeError):  # lb size != nv (:333-335)
        B.SolveInit(*prob_args(dict(pr, lb=np.zeros(3), ub=np.zeros(3))))
    B.SolveInit(*prob_args(pr))
    with pytest.raises(RuntimeError):  # UpdateEqConstraint on a joint without a constraint (:184-186)
        B.Solve(pr["q"], 2, np.eye(6), np.zeros(6))
    Bo = recursion.FirstOrderLoikOptimized(model, **dict(kw, mu_update_strat=1, max_iter=5))
    with pytest.raises(RuntimeError):  # OSQP strategy throws (hxx:632-635)
        Bo.Solve(*prob_args(pr))


def test_unbounded_revolute_equals_bounded_twin():
    """JointModelRUB*: q = (cos, sin) used as given -- the solve equals the one of the bounded joint at the same angle
    (same S, same M up to the rounding of sin/cos), and pinocchio's SO(2) integrate keeps (cos, sin) on the circle."""
    m, m0 = robots.get_robot("ur10c"), robots.get_robot("ur10")
    pb = problems.random_batch(m, 16, seed=5)
    q0 = np.zeros((16, 6))
    q0[:, 0] = np.arctan2(pb["q"][:, 1], pb["q"][:, 0]); q0[:, 1:5] = pb["q"][:, 2:6]; q0[:, 5] = np.arctan2(pb["q"][:, 7], pb["q"][:, 6])
    P = problems.bench_params(1)
    a = recursion.batch_solve(m, P, pb["q"], pb["H_ref"], pb["v_ref"], pb["ids"], pb["Ais"], pb["bis"], pb["lb"], pb["ub"], nthreads=1)
    b = recursion.batch_solve(m0, P, q0, pb["H_ref"], pb["v_ref"], pb["ids"], pb["Ais"], pb["bis"], pb["lb"], pb["ub"], nthreads=1)
    np.testing.assert_array_equal(a["iters"], b["iters"])
    assert np.abs(a["z"] - b["z"]).max() < 1e-9
    q1 = m.integrate(pb["q"], 0.05 * a["z"])
    th = q0 + 0.05 * b["z"]
    assert np.abs(q1[:, 0] - np.cos(th[:, 0])).max() < 1e-6 and np.abs(q1[:, 7] - np.sin(th[:, 5])).max() < 1e-6
    assert np.abs(np.hypot(q1[:, 0], q1[:, 1]) - 1.0).max() < 1e-6
    np.testing.assert_allclose(q1[:, 2:6], th[:, 1:5], rtol=0, atol=1e-15)


def test_multidof_step_tolerance_is_rounding_sensitivity(tmp_path):
    """Why the GPU step tests of trees with multi-DoF joints use 5e-9 instead of the reference comparator's 1e-10
    (tests/test_gpu_parity.py::test_multi_dof_joints_anywhere_step_by_step): the SAME oracle source built once without
    and once with FMA contraction -- two roundings of one formula -- already differs by that much on those trees (below a
    free-flyer, H - H (H + mu I)^-1 H is a difference of nearly equal matrices: mu = 1e-2 next to task weights of 1e2),
    while on trees of 1-DoF joints the two builds agree to 1e-10.  No formula differs; the arithmetic is that sensitive."""
    import shutil
    import subprocess

    from oracle import recursion
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    have_fma = "fma" in open("/proc/cpuinfo").read() if os.path.exists("/proc/cpuinfo") else False
    if not have_fma:
        pytest.skip("host CPU has no FMA: cannot build a second rounding of the oracle")
    lib_a = recursion.load(recursion.build(out=str(tmp_path / "a.so"), march="native", fp_contract="off"))
    lib_b = recursion.load(recursion.build(out=str(tmp_path / "b.so"), march="native", fp_contract="fast"))

    def worst_step_deviation(model, pb, params):
        worst = 0.0
        for i in range(pb["q"].shape[0]):
            sols = []
            for lib in (lib_a, lib_b):
                o = recursion.FirstOrderLoikOptimized(model, **ctor_kwargs(params), lib=lib)
                o.SolveInit(pb["q"][i], pb["H_ref"], pb["v_ref"], pb["ids"], pb["Ais"], pb["bis"][i], pb["lb"], pb["ub"])
                o.ResetSolver()
                sols.append(o)
            for it in range(1, 4):
                for o in sols:
                    o.UpdatePrev(); o.ResetInfNorms(); o.FwdPass1(); o.BwdPassOptimizedVisitor()
                    o.FwdPass2OptimizedVisitor(); o.BoxProj(); o.DualUpdate(); o.ComputeResiduals(); o.CheckConvergence()
                    o.UpdateMu()
                if sols[0].get_mu() != sols[1].get_mu():
                    break
                for f in ("His", "pis", "vis", "fis", "nu", "z", "w", "yis"):
                    a, b = getattr(sols[0], f), getattr(sols[1], f)
                    d = np.abs(a - b).max()
                    scale = max(np.abs(a).max(), np.abs(b).max(), 1.0)
                    worst = max(worst, d / scale)
        return worst

    params = dict(problems.FIXTURE_PARAMS, max_iter=200, num_eq_c=2)
    rng = np.random.default_rng(5)
    md, plain = 0.0, 0.0
    for seed in range(5):
        multidof = (0.3, 0.3, 0.6, 1.0, 0.4)[seed]
        model = robots.random_tree(9 + seed, 40 + seed, multidof=multidof)
        ids = np.array(sorted(np.random.default_rng(300 + seed).choice(np.arange(1, model.nj), size=2, replace=False)), np.int32)
        Hs = rng.normal(size=(6, 6))
        pb = dict(q=model.normalize(rng.uniform(model.q_min, model.q_max, size=(12, model.nq))), H_ref=np.eye(6) + 0.1 * (Hs + Hs.T),
                  v_ref=0.1 * rng.normal(size=6), ids=ids, Ais=np.stack([np.eye(6) + 0.3 * rng.normal(size=(6, 6)) for _ in range(2)]),
                  bis=rng.uniform(-0.5, 0.5, size=(12, 2, 6)), lb=-model.v_max, ub=model.v_max)
        md = max(md, worst_step_deviation(model, pb, params))
        model1 = robots.random_tree(9 + seed, 40 + seed, multidof=0.0)
        pb1 = dict(pb, q=model1.normalize(rng.uniform(model1.q_min, model1.q_max, size=(12, model1.nq))), lb=-model1.v_max, ub=model1.v_max,
                   ids=np.array([model1.nj - 1, max(1, model1.nj // 2)], np.int32))
        plain = max(plain, worst_step_deviation(model1, pb1, params))
    print(f"two roundings of the oracle: trees with multi-DoF joints differ by {md:.2e}, trees of 1-DoF joints by {plain:.2e}")
    assert plain < 1e-10, "1-DoF trees: the reference comparator's tolerance holds between two roundings"
    assert md < 5e-9, "the GPU tests' multi-DoF step tolerance covers the rounding sensitivity"
    assert md > 10 * plain or md > 1e-11, "multi-DoF trees are measurably more sensitive (otherwise tighten the GPU tolerance)"


@pytest.mark.parametrize("name", ["panda", "talos", "talos_ff"])
def test_no_task_constraints(name):
    """num_eq_c = 0 (ik-id-description-optimized.hpp:30-59 accepts it; UpdateEqConstraints with empty lists): only the reference
    term and the box remain.  A == B, and the answer is the closed form of the box-constrained regularisation."""
    model = robots.get_robot(name)
    rng = np.random.default_rng(3)
    params = dict(problems.FIXTURE_PARAMS, max_iter=60, num_eq_c=0)
    q = model.normalize(rng.uniform(model.q_min, model.q_max))
    args = (q, np.eye(6), 0.3 * rng.normal(size=6), np.zeros(0, np.int32), np.zeros((0, 6, 6)), np.zeros((0, 6)), -model.v_max, model.v_max)
    A, B = make_pair(model, params)
    A.Solve(*args)
    B.Solve(*args)
    assert A.iter == B.get_iter() and A.mu == B.get_mu()
    check_abs_or_rel(B.z, A.z, 1e-9, "z")
    check_abs_or_rel(B.nu, A.nu, 1e-9, "nu")
    assert np.abs(B.z).max() > 1e-3

"""CUDA path (through the C ABI) vs oracle B on the same seeded inputs.  Run on the B200 box: -m gpu.

Structure follows the reference's tests (``/root/reference/tests/loik-loid.cpp``): component-wise after each
(fused) step :305-556, end-to-end :559-671, reset / repeated solves :674-984, with the CUDA solver in the role
of the optimized solver and the oracle in the role of the ground truth.

Tolerances: the reference compares its two solvers at 1e-10 abs-or-rel (:39-83); north_star's gate is 1e-6
rel-inf on z, nu, w, y plus identical iteration counts / mu / flags.  Per-step comparisons use 1e-10
(scaled by the magnitude of the quantity), full solves use 1e-6 rel-inf and report the number of instances
whose decision trace (iteration count, final mu, status) diverged.
"""
import numpy as np
import pytest

from loik_b200 import problems, robots
from tests.helpers import check_abs_or_rel, ctor_kwargs, instance, prob_args, rel_inf

pytestmark = pytest.mark.gpu

STEP_TOL = 1e-10


def _oracle(model, params):
    from oracle import recursion
    return recursion.FirstOrderLoikOptimized(model, **ctor_kwargs(params))


def _gpu(model, params, batch):
    from loik_b200 import solver
    return solver.make_solver(model, params, batch)


def _oracle_batch(model, params, pb, nthreads=8, **kw):
    from oracle import recursion
    return recursion.batch_solve(model, params, pb["q"], pb["H_ref"], pb["v_ref"], pb["ids"], pb["Ais"], pb["bis"],
                                 pb["lb"], pb["ub"], nthreads=nthreads, **kw)


def _solve_init(G, pb):
    G.SolveInit(pb["q"], pb["H_ref"], pb["v_ref"], pb["ids"], pb["Ais"], pb["bis"], pb["lb"], pb["ub"])


def _compare_steps(model, params, pb, n_iters, what, tol=None):
    """Step-by-step: after each fused CUDA step compare every field the oracle exposes at the same point."""
    B = pb["q"].shape[0]
    STEP_TOL = tol if tol is not None else globals()["STEP_TOL"]
    G = _gpu(model, params, B)
    G.set_debug(True)
    _solve_init(G, pb)
    G.ResetRecursion()
    O = []
    for i in range(B):
        o = _oracle(model, params)
        o.SolveInit(*instance(pb, i))
        o.ResetSolver()
        O.append(o)
    nb = model.nb
    for it in range(1, n_iters + 1):
        tag = f"{what} it{it}"
        # ---- backward: UpdatePrev + ResetInfNorms + FwdPass1 + BwdPass
        for o in O:
            o.UpdatePrev(); o.ResetInfNorms(); o.FwdPass1(); o.BwdPassOptimizedVisitor()
        G.StepBackward()
        H, p, UD, Di, r = G.His, G.pis, G.UDinv, G.Dinv, G.r
        for i, o in enumerate(O):
            check_abs_or_rel(H[i], o.His[1:], STEP_TOL, tag + " His")
            check_abs_or_rel(p[i], o.pis[1:], STEP_TOL, tag + " pis")
            one = np.array([model.nv_joint(j) == 1 for j in range(1, model.nj)])  # the K x K Dinv / 6 x K UDinv of multi-DoF joints are not exposed per joint
            check_abs_or_rel(UD[i][one], o.UDinv[1:][one], STEP_TOL, tag + " UDinv")
            check_abs_or_rel(Di[i][one], o.Dinv[1:][one], STEP_TOL, tag + " Dinv")
            check_abs_or_rel(r[i], o.r, STEP_TOL, tag + " r")
        # ---- forward: FwdPass2 + BoxProj + DualUpdate (+ primal residuals)
        for o in O:
            o.FwdPass2OptimizedVisitor(); o.BoxProj(); o.DualUpdate()
        G.StepForward()
        nu, v, f, z, w, y, Aty = G.nu, G.vis, G.fis, G.z, G.w, G.yis, G.Aty
        for i, o in enumerate(O):
            check_abs_or_rel(nu[i], o.nu, STEP_TOL, tag + " nu")
            check_abs_or_rel(v[i], o.vis[1:], STEP_TOL, tag + " vis")
            check_abs_or_rel(f[i], o.fis[1:], STEP_TOL, tag + " fis")
            check_abs_or_rel(z[i], o.z, STEP_TOL, tag + " z")
            check_abs_or_rel(w[i], o.w, STEP_TOL, tag + " w")
            check_abs_or_rel(y[i], o.yis, STEP_TOL, tag + " yis")
            check_abs_or_rel(Aty[i], o.Aty, STEP_TOL, tag + " Aty")
        # ---- residual: ComputeResiduals + CheckConvergence + CheckFeasibility + UpdateMu
        for o in O:
            o.ComputeResiduals(); o.CheckConvergence()
            if it > 1:
                o.CheckFeasibility()
        G.StepResidual()
        F, T, res, nrm = G.fis_diff_plus_Aty, G.Stf_plus_w, G.get(18), G.norms()
        prv, drv = G.get_primal_residual_vec(), G.get_dual_residual_vec()
        for i, o in enumerate(O):
            fscale = max(1.0, np.abs(o.fis).max())
            assert np.abs(F[i] - o.fis_diff_plus_Aty[1:]).max() < (STEP_TOL * 1e-2) * fscale * 10, tag + " fis_diff_plus_Aty"
            assert np.abs(T[i] - o.Stf_plus_w).max() < (STEP_TOL * 1e-2) * fscale * 10, tag + " Stf_plus_w"
            check_abs_or_rel(prv[i], o.get_primal_residual_vec(), STEP_TOL, tag + " primal_residual_vec")
            assert np.abs(drv[i] - o.get_dual_residual_vec()).max() < (STEP_TOL * 1e-2) * fscale * 10, tag + " dual_residual_vec"
            check_abs_or_rel(res[i, 0], o.get_primal_residual(), STEP_TOL, tag + " primal_residual")
            assert abs(res[i, 1] - o.get_dual_residual()) < (STEP_TOL * 1e-2) * fscale * 10, tag + " dual_residual"
            check_abs_or_rel(res[i, 2], o.get_tol_primal(), STEP_TOL, tag + " tol_primal")
            check_abs_or_rel(res[i, 3], o.get_tol_dual(), max(1e-9, 10 * STEP_TOL), tag + " tol_dual")
            for nm in ("Av_inf_norm", "nu_inf_norm", "delta_vis_inf_norm", "delta_z_inf_norm", "delta_fis_inf_norm",
                       "delta_yis_inf_norm", "delta_w_inf_norm", "bT_delta_y_plus", "bT_delta_y_minus",
                       "primal_residual_task", "primal_residual_slack"):
                ref = o.scalar(nm)
                assert abs(nrm[nm][i] - ref) <= max(1e-9, 10 * STEP_TOL) * max(1.0, abs(ref)) * (fscale if "fis" in nm else 1.0), f"{tag} {nm}"
            if it > 1:
                assert bool(nrm["primal_infeasibility_cond_1"][i]) == o.get_primal_infeasibility_cond_1(), tag
                assert bool(nrm["primal_infeasibility_cond_2"][i]) == o.get_primal_infeasibility_cond_2(), tag
        # loop control: oracle instances that stopped are frozen exactly like the CUDA ones
        st = G.get_status()
        for i, o in enumerate(O):
            stopped = o.get_convergence_status() or o.get_primal_infeasibility_status()
            if not stopped:
                o.UpdateMu()
            assert bool(st[i] & 1) == o.get_convergence_status(), tag + " converged flag"
        mu = G.get_mu()
        for i, o in enumerate(O):
            assert mu[i] == o.get_mu(), tag + " mu"
        if any(o.get_convergence_status() or o.get_primal_infeasibility_status() for o in O):
            break  # per-step driving of stopped instances is the end-to-end tests' job
    G.close()


@pytest.mark.parametrize("name,bound", [("talos", 1.0), ("panda", 1.0), ("panda9", 2.0), ("ur10", 1.0)])
def test_component_wise_fixture(name, bound):
    """tests/loik-loid.cpp:305-556 on the reference's fixture shape (neutral q, A = I, b = (0,0,.5,0,0,0))."""
    model = robots.get_robot(name)
    pr = problems.fixture_problem(model, bound)
    pb = dict(pr, q=pr["q"][None], bis=pr["bis"][None])
    _compare_steps(model, dict(problems.FIXTURE_PARAMS, max_iter=200), pb, 3, name)


@pytest.mark.parametrize("name", ["panda", "ur10", "talos", "panda9", "talos_ff", "ur10c"])
def test_component_wise_random_batch(name):
    model = robots.get_robot(name)
    pb = problems.random_batch(model, 40, seed=11)
    _compare_steps(model, problems.bench_params(len(pb["ids"])), pb, 4, name)


@pytest.mark.parametrize("seed,continuous", [(0, 0.0), (1, 0.0), (2, 0.0), (3, 0.0), (4, 0.0), (5, 0.0), (6, 0.5), (7, 0.5), (8, 1.0)])
def test_component_wise_random_trees(seed, continuous):
    """Every joint type (incl. unbounded revolute, q = (cos, sin)), random branching, non-identity A and H_ref, non-zero
    v_ref, two tasks."""
    model = robots.random_tree(10 + 3 * min(seed, 5), seed, continuous=continuous)
    rng = np.random.default_rng(100 + seed)
    B = 33
    ids = np.array(sorted(rng.choice(np.arange(1, model.nj), size=2, replace=False)), np.int32)
    Hs = rng.normal(size=(6, 6))
    pb = dict(q=model.normalize(rng.uniform(model.q_min, model.q_max, size=(B, model.nq))), H_ref=np.eye(6) + 0.1 * (Hs + Hs.T),
              v_ref=0.1 * rng.normal(size=6), ids=ids,
              Ais=np.stack([np.eye(6) + 0.3 * rng.normal(size=(6, 6)) for _ in range(2)]),
              bis=rng.uniform(-0.5, 0.5, size=(B, 2, 6)), lb=-model.v_max, ub=model.v_max)
    _compare_steps(model, dict(problems.FIXTURE_PARAMS, max_iter=200, num_eq_c=2), pb, 3, f"tree{seed}")


def _compare_solves(model, params, pb, what, tol=1e-6, max_diverged_frac=0.0):
    B = pb["q"].shape[0]
    G = _gpu(model, params, B)
    _solve_init(G, pb)
    G.Solve()
    ref = _oracle_batch(model, params, pb)
    it, mu, st = G.get_iter(), G.get_mu(), G.get_status()
    ref_st = ref["status"]
    same = (it == ref["iters"]) & (mu == ref["mu"]) & ((st & 3) == (ref_st & 3))
    diverged = int((~same).sum())
    z, nu, w, y = G.z, G.nu, G.w, G.yis
    worst = 0.0
    for i in np.nonzero(same)[0]:
        worst = max(worst, rel_inf(z[i], ref["z"][i]), rel_inf(nu[i], ref["nu"][i]), rel_inf(w[i], ref["w"][i]),
                    rel_inf(y[i], ref["y"][i]))
    print(f"[{what}] B={B} diverged decision traces: {diverged}/{B}; worst rel-inf over z,nu,w,y: {worst:.3e}; "
          f"mean iters {it.mean():.2f}")
    assert worst < tol, f"{what}: rel-inf {worst:.3e} >= {tol}"
    assert diverged <= max_diverged_frac * B, f"{what}: {diverged} instances with a different iteration count / mu / status"
    s = G.stats()
    assert s["total_iters"] == int(it.sum())
    assert s["converged"] == int((st & 1).sum())
    G.close()


@pytest.mark.parametrize("name,bound", [("talos", 2.0), ("panda", 2.0), ("ur10", 2.0)])
def test_end_to_end_fixture(name, bound):
    """tests/loik-loid.cpp:559-671 (max_iter = 8, bounds +-2)."""
    model = robots.get_robot(name)
    pr = problems.fixture_problem(model, bound)
    pb = dict(pr, q=pr["q"][None], bis=pr["bis"][None])
    _compare_solves(model, dict(problems.FIXTURE_PARAMS, max_iter=8), pb, name + " fixture")


@pytest.mark.parametrize("name,B", [("panda", 4096), ("ur10", 4096), ("talos", 1024), ("panda9", 1024), ("talos_ff", 1024),
                                    ("ur10c", 2048)])
def test_end_to_end_random_batch(name, B):
    """Full solves (max_iter = 200) of the BASELINE configs at a size the oracle finishes in seconds."""
    model = robots.get_robot(name)
    pb = problems.random_batch(model, B, seed=0)
    _compare_solves(model, problems.bench_params(len(pb["ids"])), pb, name, max_diverged_frac=0.002)


def test_solve_full_and_repeat():
    """Solve(args) == SolveInit + Solve(), and repeated solves reproduce themselves (tests/loik-loid.cpp:261-303,868-984)."""
    model = robots.panda()
    pb = problems.random_batch(model, 512, seed=5)
    params = problems.bench_params(1)
    G = _gpu(model, params, 512)
    _solve_init(G, pb)
    G.Solve()
    z1, it1 = G.z, G.get_iter()
    G.Solve()
    np.testing.assert_array_equal(G.z, z1)
    np.testing.assert_array_equal(G.get_iter(), it1)
    G.Solve(pb["q"], pb["H_ref"], pb["v_ref"], pb["ids"], pb["Ais"], pb["bis"], pb["lb"], pb["ub"])
    np.testing.assert_array_equal(G.z, z1)
    np.testing.assert_array_equal(G.get_iter(), it1)
    G.close()


def test_device_pointer_inputs_and_outputs():
    """Same results when q / b arrive as CUDA tensors and z is written to a CUDA tensor."""
    import torch
    model = robots.ur10()
    pb = problems.random_batch(model, 300, seed=9)
    params = problems.bench_params(1)
    G = _gpu(model, params, 300)
    _solve_init(G, pb)
    G.Solve()
    z_host = G.z
    G2 = _gpu(model, params, 300)
    G2.SolveInit(torch.as_tensor(pb["q"]).cuda(), pb["H_ref"], pb["v_ref"], pb["ids"], pb["Ais"],
                 torch.as_tensor(pb["bis"]).cuda(), pb["lb"], pb["ub"])
    G2.Solve()
    out = torch.empty(300, model.nv, dtype=torch.float64, device="cuda")
    from loik_b200 import solver
    G2.get(solver.F_Z, out=out)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(out.cpu().numpy(), z_host)
    G.close(); G2.close()


def test_per_instance_bounds_and_shared_b():
    model = robots.panda()
    B = 64
    pb = problems.random_batch(model, B, seed=2)
    rng = np.random.default_rng(0)
    ub = model.v_max[None] * rng.uniform(0.05, 1.0, size=(B, model.nv))
    pb2 = dict(pb, lb=-ub, ub=ub, bis=pb["bis"][0])
    params = problems.bench_params(1)
    G = _gpu(model, params, B)
    _solve_init(G, pb2)
    G.Solve()
    z, it = G.z, G.get_iter()
    for i in range(B):
        o = _oracle(model, params)
        o.SolveInit(pb["q"][i], pb["H_ref"], pb["v_ref"], pb["ids"], pb["Ais"], pb["bis"][0], -ub[i], ub[i])
        o.Solve()
        assert o.get_iter() == it[i]
        assert rel_inf(z[i], o.z) < 1e-6
    G.close()


def test_tailored_solve_and_warm_start():
    """Solve(q, c_id, Ai, bi) (hpp:596-695), cold and warm-started, against the oracle's same call sequence."""
    model = robots.panda()
    B = 128
    pb = problems.random_batch(model, B, seed=4)
    pb_next = problems.random_batch(model, B, seed=5)
    for warm in (False, True):
        params = dict(problems.bench_params(1), warm_start=warm)
        G = _gpu(model, params, B)
        G.Solve(pb["q"], pb["H_ref"], pb["v_ref"], pb["ids"], pb["Ais"], pb["bis"], pb["lb"], pb["ub"])
        q2 = pb["q"] + 0.01 * (pb_next["q"] - pb["q"])
        b2 = pb_next["bis"][:, 0]
        G.Solve(q2, int(pb["ids"][0]), pb["Ais"][0], b2)
        z, it, mu = G.z, G.get_iter(), G.get_mu()
        bad = 0
        for i in range(B):
            o = _oracle(model, params)
            o.Solve(*instance(pb, i))
            o.Solve(q2[i], int(pb["ids"][0]), pb["Ais"][0], b2[i])
            if o.get_iter() != it[i] or o.get_mu() != mu[i]:
                bad += 1
                continue
            assert rel_inf(z[i], o.z) < 1e-6, f"warm={warm} instance {i}"
        assert bad == 0, f"warm={warm}: {bad} diverged decision traces"
        G.close()


def test_fixed_iteration_mode_matches_oracle():
    """Fixed-iteration (throughput) mode == the oracle driven the same way.  5 iterations: once an instance has
    converged its residuals are rounding noise and the mu-update ratio tests (hxx:617-628) become arbitrary."""
    model = robots.panda()
    B = 256
    pb = problems.random_batch(model, B, seed=6)
    params = problems.bench_params(1)
    G = _gpu(model, params, B)
    _solve_init(G, pb)
    G.IterateFixed(5)
    ref = _oracle_batch(model, params, pb, mode=1, fixed_iters=5)
    assert (G.get_iter() == 5).all()
    same = G.get_mu() == ref["mu"]
    assert same.mean() > 0.99
    z = G.z
    for i in np.nonzero(same)[0]:
        assert rel_inf(z[i], ref["z"][i]) < 1e-6
    G.close()


def test_error_paths():
    """Same error conditions as the reference's std::runtime_error sites, surfaced as RuntimeError."""
    from loik_b200 import solver
    model = robots.panda()
    p = problems.bench_params(1)
    with pytest.raises(RuntimeError, match="equality constraint dimension is not 6"):
        solver.make_solver(model, dict(p, eq_c_dim=3), 4)
    G = _gpu(model, p, 4)
    pb = problems.random_batch(model, 4, seed=0)
    with pytest.raises(RuntimeError, match="call loik_solve_init first"):
        G.Solve()
    with pytest.raises(RuntimeError, match="number of equality constraints"):
        G.SolveInit(pb["q"], pb["H_ref"], pb["v_ref"], np.array([3, 5], np.int32), np.tile(np.eye(6), (2, 1, 1)),
                    np.zeros((4, 2, 6)), pb["lb"], pb["ub"])
    with pytest.raises(RuntimeError, match="dimension"):
        G.SolveInit(pb["q"], pb["H_ref"], pb["v_ref"], pb["ids"], pb["Ais"], pb["bis"], np.zeros(3), np.zeros(3))
    _solve_init(G, pb)
    with pytest.raises(RuntimeError, match="doesn't yet exist"):
        G.Solve(pb["q"], 2, np.eye(6), np.zeros((4, 6)))
    G.close()
    Go = _gpu(model, dict(p, mu_update_strat=1), 4)
    _solve_init(Go, pb)
    with pytest.raises(RuntimeError, match="not yet implemented"):
        Go.Solve()
    Go.close()


def test_graph_replay_on_side_stream_matches_direct_launches():
    """On a non-default stream the (reset + schedule) is replayed from a CUDA graph: same results as the direct launches
    on the default stream, across repeated solves, a changed problem (graph re-capture) and a changed max_iter."""
    import torch
    model = robots.panda()
    B = 2048
    pb = problems.random_batch(model, B, seed=8)
    pb2 = problems.random_batch(model, B, seed=9)
    params = problems.bench_params(1)
    G0 = _gpu(model, params, B)
    _solve_init(G0, pb)
    G0.Solve()
    z0, it0 = G0.z, G0.get_iter()
    G1 = _gpu(model, params, B)
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        _solve_init(G1, pb)
        for _ in range(3):
            G1.Solve()
        side.synchronize()
        np.testing.assert_array_equal(G1.z, z0)
        np.testing.assert_array_equal(G1.get_iter(), it0)
        # new problem data, same batch-uniform block: the graph is reused
        _solve_init(G1, pb2)
        G1.Solve()
        side.synchronize()
        z1 = G1.z
    _solve_init(G0, pb2)
    G0.Solve()
    np.testing.assert_array_equal(z1, G0.z)
    # changed batch-uniform data (A) and max_iter: re-capture
    with torch.cuda.stream(side):
        A2 = pb["Ais"] * 1.5
        G1.set_max_iter(20)
        G1.SolveInit(pb["q"], pb["H_ref"], pb["v_ref"], pb["ids"], A2, pb["bis"], pb["lb"], pb["ub"])
        G1.Solve()
        side.synchronize()
        z2, it2 = G1.z, G1.get_iter()
    G0.set_max_iter(20)
    G0.SolveInit(pb["q"], pb["H_ref"], pb["v_ref"], pb["ids"], A2, pb["bis"], pb["lb"], pb["ub"])
    G0.Solve()
    np.testing.assert_array_equal(z2, G0.z)
    np.testing.assert_array_equal(it2, G0.get_iter())
    assert it2.max() <= 20
    G0.close(); G1.close()


def test_outer_ik_loop_on_device():
    """integrate q <- q + dt z on the device, then the tailored Solve for the next target,
    against the same sequence driven through the oracle (user-side integration + Solve(q, c_id, Ai, bi))."""
    model = robots.panda()
    B = 256
    pb = problems.random_batch(model, B, seed=12)
    nxt = problems.random_batch(model, B, seed=13)
    params = dict(problems.bench_params(1), warm_start=True)
    c_id, A, dt = int(pb["ids"][0]), pb["Ais"][0], 0.05
    G = _gpu(model, params, B)
    G.Solve(pb["q"], pb["H_ref"], pb["v_ref"], pb["ids"], pb["Ais"], pb["bis"], pb["lb"], pb["ub"])
    z0 = G.z
    G.Integrate(dt)
    np.testing.assert_allclose(G.q, pb["q"] + dt * z0, rtol=0, atol=1e-15)
    b1 = 0.5 * (pb["bis"][:, 0] + nxt["bis"][:, 0])
    G.Solve(None, c_id, A, b1)
    z1, it1, mu1 = G.z, G.get_iter(), G.get_mu()
    bad = 0
    for i in range(B):
        o = _oracle(model, params)
        o.Solve(*instance(pb, i))
        q1 = pb["q"][i] + dt * o.z
        o.Solve(q1, c_id, A, b1[i])
        if o.get_iter() != it1[i] or o.get_mu() != mu1[i]:
            bad += 1
            continue
        assert rel_inf(z1[i], o.z) < 1e-6
    assert bad <= 1
    G.close()


def test_outer_ik_loop_unbounded_revolute():
    """The same outer loop on a robot with URDF `continuous` joints (JointModelRUBZ / RUBY, q = (cos, sin)): the
    device-side integration is pinocchio's SO(2) update (rotate, first-order renormalise), checked against
    RobotModel.integrate, and the q getter returns (cos, sin) pairs."""
    model = robots.get_robot("ur10c")
    B = 192
    pb = problems.random_batch(model, B, seed=31)
    nxt = problems.random_batch(model, B, seed=32)
    params = dict(problems.bench_params(1), warm_start=True)
    c_id, A, dt = int(pb["ids"][0]), pb["Ais"][0], 0.05
    G = _gpu(model, params, B)
    G.Solve(pb["q"], pb["H_ref"], pb["v_ref"], pb["ids"], pb["Ais"], pb["bis"], pb["lb"], pb["ub"])
    np.testing.assert_array_equal(G.q, pb["q"])
    z0 = G.z
    G.Integrate(dt)
    q1 = model.integrate(pb["q"], dt * z0)
    np.testing.assert_allclose(G.q, q1, rtol=0, atol=2e-15)
    b1 = 0.5 * (pb["bis"][:, 0] + nxt["bis"][:, 0])
    G.Solve(None, c_id, A, b1)
    z1, it1, mu1 = G.z, G.get_iter(), G.get_mu()
    bad = 0
    for i in range(B):
        o = _oracle(model, params)
        o.Solve(*instance(pb, i))
        o.Solve(model.integrate(pb["q"][i], dt * o.z), c_id, A, b1[i])
        if o.get_iter() != it1[i] or o.get_mu() != mu1[i]:
            bad += 1
            continue
        assert rel_inf(z1[i], o.z) < 1e-6
    assert bad <= 1
    G.close()


@pytest.mark.parametrize("name,warm", [("panda", True), ("panda", False), ("ur10c", True), ("talos_ff", True)])
def test_tracking_loop_vs_oracle(name, warm):
    """The trajectory-tracking loop (full Solve, then T x { Integrate(dt); Solve(q on device, c_id, A, b_t) }) against
    the oracle driven the same way (lo_batch_track): iteration count of every instance at the last step, final z, and
    the integrated configuration."""
    from oracle import recursion
    model = robots.get_robot(name)
    B, T, dt = (512 if model.nb < 20 else 128), 6, 0.02
    pb = problems.random_batch(model, B, seed=51)
    nx = problems.random_batch(model, B, seed=52)
    params = dict(problems.bench_params(len(pb["ids"])), warm_start=warm)
    c_id, A = int(pb["ids"][0]), pb["Ais"][0]
    G = _gpu(model, params, B)
    G.Solve(pb["q"], pb["H_ref"], pb["v_ref"], pb["ids"], pb["Ais"], pb["bis"], pb["lb"], pb["ub"])
    for t in range(1, T + 1):
        a = t / T
        G.Integrate(dt)
        G.Solve(None, c_id, A, (1.0 - a) * pb["bis"][:, 0] + a * nx["bis"][:, 0])
    ref = recursion.batch_track(model, params, pb["q"], pb["H_ref"], pb["v_ref"], pb["ids"], pb["Ais"], pb["bis"], nx["bis"], pb["lb"],
                                pb["ub"], c_id=c_id, dt=dt, steps=T, warm=warm, nthreads=4)
    same = G.get_iter() == ref["step_iters"][:, -1]
    # a decision that flips on rounding at an earlier step changes the trajectory of that instance from there on
    assert same.mean() >= 0.99, f"{(~same).sum()} of {B} instances ended with a different iteration count"
    z, q = G.z, G.q
    for i in np.nonzero(same)[0]:
        if rel_inf(q[i], ref["q"][i]) < 1e-9:
            assert rel_inf(z[i], ref["z"][i]) < 1e-6
    close_q = np.array([rel_inf(q[i], ref["q"][i]) < 1e-9 for i in range(B)])
    assert close_q.mean() >= 0.99
    G.close()


def test_update_references_per_joint():
    """problem_.UpdateReferences(H_refs, v_refs) (ik-id-description-optimized.hpp:103-121): per-joint symmetric weights and
    reference velocities after SolveInit, against the oracle driven the same way."""
    model = robots.panda(fingers=True)
    B = 96
    rng = np.random.default_rng(21)
    pb = problems.random_batch(model, B, seed=14)
    H_refs = np.zeros((model.nj, 6, 6))
    v_refs = 0.05 * rng.normal(size=(model.nj, 6))
    for i in range(model.nj):
        M = rng.normal(size=(6, 6))
        H_refs[i] = np.eye(6) * rng.uniform(0.5, 2.0) + 0.05 * (M + M.T)
    params = problems.bench_params(1, max_iter=60)
    G = _gpu(model, params, B)
    _solve_init(G, pb)
    G.UpdateReferences(H_refs, v_refs)
    G.Solve()
    z, it, mu = G.z, G.get_iter(), G.get_mu()
    bad = 0
    for i in range(B):
        o = _oracle(model, params)
        o.SolveInit(*instance(pb, i))
        o.UpdateReferences(H_refs, v_refs)
        o.Solve()
        if o.get_iter() != it[i] or o.get_mu() != mu[i]:
            bad += 1
            continue
        assert rel_inf(z[i], o.z) < 1e-6, i
    assert bad == 0
    with pytest.raises(RuntimeError, match="symmetric"):
        Hn = H_refs.copy(); Hn[2, 0, 1] += 0.3
        G.UpdateReferences(Hn, v_refs)
    with pytest.raises(RuntimeError, match="wrong size"):
        G.UpdateReferences(H_refs[:-1], v_refs[:-1])
    G.close()


@pytest.mark.parametrize("name,B,kw,h_per", [("panda9", 160, {}, False), ("talos", 96, {}, False), ("talos_ff", 64, {}, False), ("panda", 2048, dict(max_iter=200), False),
                                             ("panda9", 160, {}, True), ("talos", 96, {}, True), ("talos_ff", 64, {}, True), ("panda", 2048, dict(max_iter=200), True)])
def test_update_references_per_instance(name, B, kw, h_per):
    """loik_update_references_batch: per-joint weights shared by the batch, a reference velocity of its own for every instance
    and joint (UpdateReferences applied to each instance of the batch, ik-id-description-optimized.hpp:103-121), through
    dense sweeps, migrating re-pack launches and (Talos) the segment kernel; fused steps at 1e-10 and full solves against
    the oracle driven instance by instance.  The second solve runs through the cached launch graph."""
    model = robots.get_robot(name)
    rng = np.random.default_rng(5)
    pb = problems.random_batch(model, B, seed=23)
    nc = len(pb["ids"])
    H_refs = np.zeros((model.nj, 6, 6))
    for i in range(model.nj):
        M = rng.normal(size=(6, 6))
        H_refs[i] = np.eye(6) * rng.uniform(0.5, 2.0) + 0.05 * (M + M.T)
    v_refs = 0.05 * rng.normal(size=(B, model.nj, 6))
    if h_per:  # every instance its own symmetric weights too
        W = rng.normal(size=(B, model.nj, 6, 6))
        H_refs = H_refs[None] * rng.uniform(0.7, 1.4, size=(B, model.nj, 1, 1)) + 0.03 * (W + W.transpose(0, 1, 3, 2))
    Hof = (lambda i: H_refs[i]) if h_per else (lambda i: H_refs)
    params = problems.bench_params(nc, **(kw or dict(max_iter=60)))
    G = _gpu(model, params, B)
    _solve_init(G, pb)
    G.UpdateReferences(H_refs, v_refs)
    # two fused iterations, field by field (incl. tol_dual, which sees this instance's |H_ref v_ref|inf)
    D = _gpu(model, params, B)
    D.set_debug(True)
    _solve_init(D, pb)
    D.UpdateReferences(H_refs, v_refs)
    D.ResetRecursion()
    pick = list(range(0, B, max(1, B // 8)))
    O = []
    for i in pick:
        o = _oracle(model, params)
        o.SolveInit(*instance(pb, i))
        o.UpdateReferences(Hof(i), v_refs[i])
        o.ResetSolver()
        O.append(o)
    for itn in (1, 2):
        D.StepBackward(); D.StepForward(); D.StepResidual()
        vis, fis, nu, res = D.vis, D.fis, D.nu, D.get(18)
        for i, o in zip(pick, O):
            o.UpdatePrev(); o.ResetInfNorms(); o.FwdPass1(); o.BwdPassOptimizedVisitor(); o.FwdPass2OptimizedVisitor(); o.BoxProj()
            o.DualUpdate(); o.ComputeResiduals(); o.CheckConvergence()
            if itn > 1:
                o.CheckFeasibility()
            o.UpdateMu()
            tol = 1e-8 if any(model.nv_joint(j) > 1 for j in range(1, model.nj)) else 1e-10
            check_abs_or_rel(vis[i], o.vis[1:], tol, "vis")
            check_abs_or_rel(fis[i], o.fis[1:], tol * max(1.0, np.abs(o.fis).max()), "fis")
            check_abs_or_rel(nu[i], o.nu, tol, "nu")
            check_abs_or_rel(res[i, 3], o.get_tol_dual(), 1e-8, "tol_dual")
    D.close()
    G.set_keep_workspace(True)  # (instances that retire from a re-packed arena bring their workspace home: the reference rows behind it must stay)
    for rep in range(2):
        G.Solve()
    z, it, mu = G.z, G.get_iter(), G.get_mu()
    bad, worst = 0, 0.0
    for i in range(0, B, max(1, B // 160)):
        o = _oracle(model, params)
        o.SolveInit(*instance(pb, i))
        o.UpdateReferences(Hof(i), v_refs[i])
        o.Solve()
        if o.get_iter() != it[i] or o.get_mu() != mu[i]:
            bad += 1
            continue
        worst = max(worst, rel_inf(z[i], o.z))
    assert bad == 0 and worst < 1e-6, (bad, worst)
    # SolveInit puts the batch back on one shared reference
    _solve_init(G, pb)
    G.Solve()
    o = _oracle(model, params)
    o.SolveInit(*instance(pb, 3))
    o.Solve()
    assert o.get_iter() == G.get_iter()[3] and rel_inf(G.z[3], o.z) < 1e-6
    with pytest.raises(RuntimeError, match="wrong size"):
        G.UpdateReferences(H_refs, v_refs[:, :-1])
    if h_per:
        with pytest.raises(RuntimeError, match="symmetric"):
            Hn = H_refs.copy(); Hn[1, 2, 0, 1] += 0.3
            G.UpdateReferences(Hn, v_refs)
    G.close()


@pytest.mark.parametrize("name", ["panda", "ur10", "talos", "panda9", "ur10c", "tree_zyx"])
def test_against_committed_golden_fixtures(name):
    """CUDA path vs tests/golden/random_*.npz (frozen oracle-pair outputs, scripts/make_golden.py) -- no oracle call."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", f"random_{name}.npz"))
    model = robots.get_robot(name)
    n = int(g["n"])
    pb = problems.random_batch(model, n, seed=int(g["seed"]))
    params = problems.bench_params(len(pb["ids"]))
    G = _gpu(model, params, n)
    _solve_init(G, pb)
    G.Solve()
    np.testing.assert_array_equal(G.get_iter(), g["iter"])
    np.testing.assert_array_equal(G.get_mu(), g["mu"])
    np.testing.assert_array_equal(G.get_convergence_status(), g["converged"])
    np.testing.assert_array_equal(G.get_primal_infeasibility_status(), g["primal_infeasible"])
    z, nu, w, y = G.z, G.nu, G.w, G.yis
    for i in range(n):
        assert rel_inf(z[i], g["z"][i]) < 1e-6 and rel_inf(nu[i], g["nu"][i]) < 1e-6
        assert rel_inf(w[i], g["w"][i]) < 1e-6 and rel_inf(y[i], g["yis"][i]) < 1e-6
    G.close()


@pytest.mark.parametrize("name", ["talos", "panda", "ur10"])
def test_fixture_golden_on_gpu(name):
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", f"fixture_{name}.npz"))
    model = robots.get_robot(name)
    pr = problems.fixture_problem(model, float(g["bound"]))
    G = _gpu(model, dict(problems.FIXTURE_PARAMS, max_iter=int(g["max_iter"])), 1)
    G.Solve(pr["q"][None], pr["H_ref"], pr["v_ref"], pr["ids"], pr["Ais"], pr["bis"][None], pr["lb"], pr["ub"])
    assert int(G.get_iter()[0]) == int(g["iter"]) and float(G.get_mu()[0]) == float(g["mu"])
    assert bool(G.get_convergence_status()[0]) == bool(g["converged"])
    assert bool(G.get_primal_infeasibility_status()[0]) == bool(g["primal_infeasible"])
    assert rel_inf(G.z[0], g["z"]) < 1e-6 and rel_inf(G.w[0], g["w"]) < 1e-6 and rel_inf(G.vis[0], g["vis"][1:]) < 1e-6
    G.close()


def test_free_flyer_root_joint():
    """SURVEY.md section 8(f) rank 4 (first step): a JointModelFreeFlyer root (nq 7, nv 6; floating-base Talos, nv = 38),
    per-instance bounds, a task on the root joint itself, sub-batch invariance and the unsupported corners."""
    model = robots.talos(floating=True)
    B = 200
    pb = problems.random_batch(model, B, seed=31)
    rng = np.random.default_rng(5)
    ub = model.v_max[None] * rng.uniform(0.2, 1.0, size=(B, model.nv))
    ids = np.array([1, 22, 30], np.int32)                     # root joint + both wrists
    pb = dict(pb, ids=ids, Ais=np.tile(np.eye(6), (3, 1, 1)), bis=rng.uniform(-0.3, 0.3, size=(B, 3, 6)), lb=-ub, ub=ub)
    params = problems.bench_params(3, max_iter=80)
    G = _gpu(model, params, B)
    _solve_init(G, pb)
    G.Solve()
    z, nu, w, y, it, mu = G.z, G.nu, G.w, G.yis, G.get_iter(), G.get_mu()
    assert z.shape == (B, 38) and G.q.shape == (B, 39)
    np.testing.assert_array_equal(G.q, pb["q"])
    bad = 0
    for i in range(B):
        o = _oracle(model, params)
        o.SolveInit(pb["q"][i], pb["H_ref"], pb["v_ref"], ids, pb["Ais"], pb["bis"][i], -ub[i], ub[i])
        o.Solve()
        if o.get_iter() != it[i] or o.get_mu() != mu[i]:
            bad += 1
            continue
        assert rel_inf(z[i], o.z) < 1e-6 and rel_inf(nu[i], o.nu) < 1e-6 and rel_inf(w[i], o.w) < 1e-6
        assert rel_inf(y[i], o.yis) < 1e-6
    assert bad <= 1
    G2 = _gpu(model, params, 37)
    sub = dict(pb, q=pb["q"][50:87], bis=pb["bis"][50:87], lb=pb["lb"][50:87], ub=pb["ub"][50:87])
    _solve_init(G2, sub)
    G2.Solve()
    np.testing.assert_array_equal(G2.z, z[50:87])
    G.Integrate(0.01)  # M * exp6(dt z) for the floating base (SpecialEuclideanOperationTpl<3>::integrate_impl)
    np.testing.assert_allclose(G.q, model.integrate(pb["q"], 0.01 * z), rtol=0, atol=5e-15)
    G.set_debug(True)
    with pytest.raises(RuntimeError, match="multi-DoF"):
        G.FwdPass1()
    G.close(); G2.close()
    from loik_b200 import solver
    bad_model = robots.talos(floating=True)
    bad_model.jtype = bad_model.jtype.copy(); bad_model.jtype[5] = 99
    with pytest.raises(RuntimeError, match="unsupported joint type"):
        solver.make_solver(bad_model, params, 4)


def _multidof_problem(model, B, seed):
    rng = np.random.default_rng(300 + seed)
    ids = np.array(sorted(rng.choice(np.arange(1, model.nj), size=2, replace=False)), np.int32)
    Hs = rng.normal(size=(6, 6))
    return dict(q=model.normalize(rng.uniform(model.q_min, model.q_max, size=(B, model.nq))), H_ref=np.eye(6) + 0.1 * (Hs + Hs.T),
                v_ref=0.1 * rng.normal(size=6), ids=ids, Ais=np.stack([np.eye(6) + 0.3 * rng.normal(size=(6, 6)) for _ in range(2)]),
                bis=rng.uniform(-0.5, 0.5, size=(B, 2, 6)), lb=-model.v_max, ub=model.v_max)


@pytest.mark.parametrize("seed,multidof,continuous", [(0, 0.3, 0.0), (1, 0.3, 0.3), (2, 0.6, 0.0), (3, 1.0, 0.0), (4, 0.4, 0.0)])
def test_multi_dof_joints_anywhere_step_by_step(seed, multidof, continuous):
    """spherical / translation joints and free-flyers anywhere in a random tree (next to
    every 1-DoF type), two tasks that may sit on the multi-DoF joints themselves: every fused step against the oracle."""
    model = robots.random_tree(9 + seed, 40 + seed, continuous=continuous, multidof=multidof)
    assert any(model.nv_joint(i) > 1 for i in range(1, model.nj))
    pb = _multidof_problem(model, 33, seed)
    # several free-flyers in series: H - H (H + mu I)^-1 H below each of them is a difference of nearly equal matrices
    # (mu = 1e-2 next to task weights of 1e2): rounding differences between two orders of evaluation reach a few 1e-10
    _compare_steps(model, dict(problems.FIXTURE_PARAMS, max_iter=200, num_eq_c=2), pb, 3, f"mdtree{seed}", tol=5e-9)


@pytest.mark.parametrize("seed,multidof", [(0, 0.3), (1, 0.5), (2, 1.0)])
def test_multi_dof_joints_anywhere_full_solves(seed, multidof):
    """Full solves (per-instance loop control, migrating launches, retire) on trees with multi-DoF joints; per-instance
    bounds on one of them; the getters' per-dof layout (z, nu, w of 3- and 6-dof joints at idx_v)."""
    model = robots.random_tree(10 + 2 * seed, 60 + seed, multidof=multidof)
    B = 300
    pb = _multidof_problem(model, B, 10 + seed)
    if seed == 1:
        rng = np.random.default_rng(7)
        ub = model.v_max[None] * rng.uniform(0.3, 1.0, size=(B, model.nv))
        pb = dict(pb, lb=-ub, ub=ub)
    _compare_solves(model, dict(problems.FIXTURE_PARAMS, max_iter=120, num_eq_c=2, tol_abs=1e-3, tol_rel=1e-3), pb, f"mdsolve{seed}",
                    max_diverged_frac=0.01)


@pytest.mark.parametrize("seed,multidof,zyx", [(0, 0.0, 0.4), (1, 0.3, 0.4), (2, 0.0, 1.0)])
def test_spherical_zyx_joints_step_by_step(seed, multidof, zyx):
    """JointModelSphericalZYX anywhere in a random tree: the motion subspace S = [0; E(q)] differs per instance (rows FR_S of
    the multi-DoF block, written by FwdPassInit), U = H S is a product instead of a column selection: every fused step
    against the oracle."""
    model = robots.random_tree(9 + seed, 141 + seed, multidof=multidof, zyx=zyx)
    assert (model.jtype == robots.ZYX).any()
    pb = _multidof_problem(model, 33, seed)
    _compare_steps(model, dict(problems.FIXTURE_PARAMS, max_iter=200, num_eq_c=2), pb, 3, f"zyxtree{seed}", tol=5e-9)


@pytest.mark.parametrize("seed,multidof", [(0, 0.0), (1, 0.4)])
def test_spherical_zyx_joints_full_solves_and_integrate(seed, multidof):
    """Full solves (migrating launches carry the S rows along) on trees with SphericalZYX joints, then the device-side
    outer loop: q <- q + dt z for the Euler angles, M and S of the new configuration, warm-started tracking solve."""
    model = robots.random_tree(10 + 2 * seed, 160 + seed, multidof=multidof, zyx=0.5)
    assert (model.jtype == robots.ZYX).any()
    B = 300
    pb = _multidof_problem(model, B, 20 + seed)
    params = dict(problems.FIXTURE_PARAMS, max_iter=120, num_eq_c=2, tol_abs=1e-3, tol_rel=1e-3)
    _compare_solves(model, params, pb, f"zyxsolve{seed}", max_diverged_frac=0.01)
    G = _gpu(model, params, B)
    _solve_init(G, pb)
    np.testing.assert_array_equal(G.q, pb["q"])
    L = G.liMi
    for i in range(0, B, 37):
        o = _oracle(model, params)
        o.SolveInit(*instance(pb, i))
        np.testing.assert_allclose(L[i][:, :9].reshape(-1, 3, 3), o.liMi_R[1:], rtol=0, atol=1e-14)
    G.Solve()
    zz = G.z
    G.Integrate(0.05)
    q1 = model.integrate(pb["q"], 0.05 * zz)
    np.testing.assert_allclose(G.q, q1, rtol=0, atol=5e-15)
    c_id = int(pb["ids"][0])
    G.Solve(None, c_id, pb["Ais"][0], pb["bis"][:, 0])
    z2, it2 = G.z, G.get_iter()
    bad = 0
    for i in range(0, B, 11):
        o = _oracle(model, params)
        o.Solve(*instance(pb, i))
        o.Solve(q1[i], c_id, pb["Ais"][0], pb["bis"][i, 0])
        if o.get_iter() != it2[i]:
            bad += 1
            continue
        assert rel_inf(z2[i], o.z) < 1e-6, i
    assert bad <= 1
    G.close()


def test_multi_dof_fwd_pass_init_and_integrate():
    """FwdPassInit for multi-DoF joints (liMi = placement * M(q): quaternion of free-flyer / spherical joints, offset of
    translation joints) against the oracle, the q getter layout, and the device-side integrate of a tree whose only
    multi-DoF joints are translation joints (vector space) next to unbounded revolute ones."""
    model = robots.random_tree(12, 77, multidof=0.5)
    B = 40
    pb = _multidof_problem(model, B, 3)
    params = dict(problems.FIXTURE_PARAMS, max_iter=50, num_eq_c=2)
    G = _gpu(model, params, B)
    _solve_init(G, pb)
    np.testing.assert_array_equal(G.q, pb["q"])
    L = G.liMi
    for i in range(0, B, 7):
        o = _oracle(model, params)
        o.SolveInit(*instance(pb, i))
        np.testing.assert_allclose(L[i][:, :9].reshape(-1, 3, 3), o.liMi_R[1:], rtol=0, atol=1e-14)
        np.testing.assert_allclose(L[i][:, 9:], o.liMi_p[1:], rtol=0, atol=1e-14)
    G.Solve()
    zz = G.z
    G.Integrate(0.05)  # quaternion * exp3 / M * exp6 for the spherical joints and free-flyers of the tree
    np.testing.assert_allclose(G.q, model.integrate(pb["q"], 0.05 * zz), rtol=0, atol=5e-15)
    L = G.liMi
    for i in range(0, B, 9):
        o = _oracle(model, params)
        o.SolveInit(model.integrate(pb["q"][i], 0.05 * zz[i]), *instance(pb, i)[1:])
        np.testing.assert_allclose(L[i][:, :9].reshape(-1, 3, 3), o.liMi_R[1:], rtol=0, atol=1e-13)
        np.testing.assert_allclose(L[i][:, 9:], o.liMi_p[1:], rtol=0, atol=1e-13)
    G.close()
    # translation + unbounded revolute joints: integrate on the device == RobotModel.integrate, then a tailored solve
    J = [("j1", 0, "R", "z", (0, 0, 0.1), (0, 0, 0), -2, 2, 2.0), ("t2", 1, "T", None, (0.1, 0, 0.2), (0.3, -0.2, 0.5), None, None, 1.5),
         ("c3", 2, "C", "y", (0, 0.1, 0.1), (0, 0.4, 0), None, None, 2.0), ("j4", 3, "R", (1.0, 1.0, 0.0), (0.2, 0, 0), (0, 0, 0.7), -2, 2, 2.0),
         ("t5", 2, "T", None, (0, -0.2, 0.1), (0, 0, 0), None, None, 1.0)]
    model = robots._build("tra_tree", J)
    pb = _multidof_problem(model, B, 4)
    pb = dict(pb, ids=np.array([4, 5], np.int32))
    params = dict(problems.bench_params(2), warm_start=True)
    G = _gpu(model, params, B)
    G.Solve(pb["q"], pb["H_ref"], pb["v_ref"], pb["ids"], pb["Ais"], pb["bis"], pb["lb"], pb["ub"])
    z0, dt = G.z, 0.03
    G.Integrate(dt)
    q1 = model.integrate(pb["q"], dt * z0)
    np.testing.assert_allclose(G.q, q1, rtol=0, atol=2e-15)
    b1 = 0.7 * pb["bis"][:, 0]
    G.Solve(None, 4, pb["Ais"][0], b1)
    z1, it1, mu1 = G.z, G.get_iter(), G.get_mu()
    bad = 0
    for i in range(B):
        o = _oracle(model, params)
        o.Solve(*instance(pb, i))
        o.Solve(model.integrate(pb["q"][i], dt * o.z), 4, pb["Ais"][0], b1[i])
        if o.get_iter() != it1[i] or o.get_mu() != mu1[i]:
            bad += 1
            continue
        assert rel_inf(z1[i], o.z) < 1e-6
    assert bad <= 1
    G.close()


@pytest.mark.parametrize("name,B", [("panda", 192), ("talos", 96), ("talos_ff", 64)])
def test_workspace_after_solve(name, B):
    """The reference leaves the last backward pass in the caller's data after Solve() (His, pis: tests/loik-loid.cpp:597-615;
    jdata UDinv / Dinv, r).  With set_keep_workspace the batched solve does too, although its instances finish in
    re-packed arenas; without it those getters fail loudly instead of returning stale rows."""
    model = robots.get_robot(name)
    pb = problems.random_batch(model, B, seed=21)
    params = problems.bench_params(len(pb["ids"]))
    G = _gpu(model, params, B)
    _solve_init(G, pb)
    G.Solve()
    with pytest.raises(RuntimeError, match="loik_set_keep_workspace"):
        G.His
    z0, it0 = G.z, G.get_iter()
    assert it0.max() > 8  # some instances finish after the dense sweeps, i.e. away from their home slot
    G.set_keep_workspace(True)
    G.Solve()  # (the captured graph is rebuilt: keep_ws is a kernel parameter)
    np.testing.assert_array_equal(G.z, z0)
    np.testing.assert_array_equal(G.get_iter(), it0)
    H, p, UD, Di, r = G.His, G.pis, G.UDinv, G.Dinv, G.r
    one = np.array([model.nv_joint(j) == 1 for j in range(1, model.nj)])
    checked = 0
    for i in range(B):
        o = _oracle(model, params)
        o.Solve(*instance(pb, i))
        if o.get_iter() != it0[i]:
            continue  # a diverged decision trace (allowed fraction: see _compare_solves)
        checked += 1
        tag = f"{name} #{i} after Solve ({it0[i]} iterations)"
        check_abs_or_rel(H[i], o.His[1:], 1e-6, tag + " His")
        check_abs_or_rel(p[i], o.pis[1:], 1e-6, tag + " pis")
        check_abs_or_rel(UD[i][one], o.UDinv[1:][one], 1e-6, tag + " UDinv")
        check_abs_or_rel(Di[i][one], o.Dinv[1:][one], 1e-6, tag + " Dinv")
        check_abs_or_rel(r[i], o.r, 1e-6, tag + " r")
    assert checked >= B - 1
    # the step-by-step interface works in place: always readable, also with keep_workspace off
    G.set_keep_workspace(False)
    G.Solve()
    G.ResetRecursion()
    G.StepBackward()
    assert np.isfinite(G.His).all()
    G.close()

"""ABI surface beyond the solves: base-class setters / getters between solves (the cached launch graph must follow),
ResetSolver() vs ResetRecursion(), the chunked global-stop path of the sharded driver, launch-schedule knobs.  -m gpu."""
import numpy as np
import pytest

from loik_b200 import problems, robots, sharded
from tests.helpers import ctor_kwargs, instance, rel_inf

pytestmark = pytest.mark.gpu


def _gpu(model, params, batch):
    from loik_b200 import solver
    return solver.make_solver(model, params, batch)


def _solve_init(G, pb):
    G.SolveInit(pb["q"], pb["H_ref"], pb["v_ref"], pb["ids"], pb["Ais"], pb["bis"], pb["lb"], pb["ub"])


def _oracle_batch(model, params, pb, **kw):
    from oracle import recursion
    return recursion.batch_solve(model, params, pb["q"], pb["H_ref"], pb["v_ref"], pb["ids"], pb["Ais"], pb["bis"],
                                 pb["lb"], pb["ub"], nthreads=8, **kw)


def _check(G, model, params, pb, what):
    ref = _oracle_batch(model, params, pb)
    it, mu, st = G.get_iter(), G.get_mu(), G.get_status()
    same = (it == ref["iters"]) & (mu == ref["mu"]) & ((st & 3) == (ref["status"] & 3))
    assert same.mean() >= 0.998, f"{what}: {int((~same).sum())} diverged decision traces"
    assert rel_inf(G.z[same], ref["z"][same]) < 1e-6, what


@pytest.mark.parametrize("lane_after", [-1, 0])
def test_setters_between_solves_follow_into_the_cached_graph(lane_after):
    """task-solver-base.hpp:105-141: every setter changes the NEXT solve (the launch graph is re-captured when the
    parameter block changes) and agrees with an oracle constructed with the new value."""
    import torch
    model = robots.panda()
    B = 512
    pb = problems.random_batch(model, B, seed=31)
    params = problems.bench_params(1)
    G = _gpu(model, params, B)
    G.set_schedule(lane_after=lane_after)
    _solve_init(G, pb)
    stream = torch.cuda.Stream()  # a capturable stream: the graph path
    changes = [("tol_abs", 1e-5), ("tol_rel", 1e-5), ("tol_primal_inf", 1e-3), ("rho", 1e-4), ("mu", 1e-1),
               ("mu_equality_scale_factor", 1e3), ("tol_tail_solve", 1e-3), ("max_iter", 60), ("tol_dual_inf", 1e-3)]
    with torch.cuda.stream(stream):
        G.Solve()
        stream.synchronize()
        _check(G, model, params, pb, "initial")
        for name, val in changes:
            getattr(G, "set_" + name)(val)
            params = dict(params, **{name: val})
            assert G.get_params()[name] == val
            G.Solve()
            stream.synchronize()
            _check(G, model, params, pb, f"after set_{name}({val})")
    assert G.get_rho() == 1e-4 and G.get_max_iter() == 60 and G.get_tol_primal_inf() == 1e-3 and G.get_tol_dual_inf() == 1e-3
    G.close()


def test_reset_solver_keeps_the_state_reset_recursion_clears_it():
    """ResetSolver() (hpp:168-186) resets iter / flags / mu only; ResetRecursion() (data hxx:138-154) zeroes w, z, vis, fis,
    yis, Aty as well.  The oracle's two methods are the reference."""
    from oracle import recursion
    model = robots.ur10()
    B = 32
    pb = problems.random_batch(model, B, seed=12)
    params = problems.bench_params(1)
    G = _gpu(model, params, B)
    _solve_init(G, pb)
    G.Solve()
    z1, w1, v1, y1, mu1 = G.z, G.w, G.vis, G.yis, G.get_mu()
    assert np.abs(z1).max() > 0 and (G.get_iter() > 0).all()
    G.ResetSolver()
    np.testing.assert_array_equal(G.z, z1)
    np.testing.assert_array_equal(G.w, w1)
    np.testing.assert_array_equal(G.vis, v1)
    np.testing.assert_array_equal(G.yis, y1)
    assert (G.get_iter() == 0).all() and (G.get_mu() == params["mu"]).all() and (G.get_status() == 0).all()
    # stepping on from the kept state == the oracle after Solve(); ResetSolver(); one iteration
    G.StepBackward(); G.StepForward(); G.StepResidual()
    for i in range(B):
        o = recursion.FirstOrderLoikOptimized(model, **ctor_kwargs(params))
        o.SolveInit(*instance(pb, i))
        o.Solve()
        o.ResetSolver()
        o.UpdatePrev(); o.ResetInfNorms(); o.FwdPass1(); o.BwdPassOptimizedVisitor()
        o.FwdPass2OptimizedVisitor(); o.BoxProj(); o.DualUpdate()
        assert rel_inf(G.z[i], o.z) < 1e-9 and rel_inf(G.nu[i], o.nu) < 1e-9 and rel_inf(G.w[i], o.w) < 1e-9
    G.ResetRecursion()
    assert np.abs(G.z).max() == 0 and np.abs(G.w).max() == 0 and np.abs(G.vis).max() == 0 and np.abs(G.yis).max() == 0
    G.close()


@pytest.mark.parametrize("name,B", [("panda", 2048), ("talos", 256)])
def test_chunked_global_stop_path_matches_solve(name, B):
    """ShardedSolver.solve_chunked (world 1): chunks of sweeps + the active-count read-back until nothing is active ==
    Solve(), instance by instance, and stops before max_iter when everything has finished."""
    model = robots.get_robot(name)
    pb = problems.random_batch(model, B, seed=14)
    params = problems.bench_params(len(pb["ids"]))
    G = _gpu(model, params, B)
    G.set_schedule(lane_after=-1)
    _solve_init(G, pb)
    G.Solve()
    z, it, mu, st = G.z, G.get_iter(), G.get_mu(), G.get_status()
    drv = sharded.ShardedSolver(G, 1, chunk=8)
    sweeps = drv.solve_chunked()
    np.testing.assert_array_equal(G.get_iter(), it)
    np.testing.assert_array_equal(G.get_mu(), mu)
    np.testing.assert_array_equal(G.get_status(), st)
    np.testing.assert_array_equal(G.z, z)
    assert sweeps >= it.max() and sweeps - it.max() < 8 and sweeps % 8 == 0 or sweeps == params["max_iter"]
    tot = drv.solve()
    assert int(tot[3].item()) == int(it.sum()) and int(tot[0].item()) == int((st & 1).sum())
    G.close()


def test_solve_chunk_needs_a_problem():
    model = robots.panda()
    G = _gpu(model, problems.bench_params(1), 8)
    with pytest.raises(RuntimeError, match="loik_solve_init"):
        G.SolveChunk(4)
    G.close()


def test_get_into_a_wrong_buffer_raises():
    """get(out=...) writes in place: a buffer of the wrong dtype / shape / layout is an error, not a silent copy."""
    import torch
    from loik_b200 import solver as lk
    model = robots.panda()
    B = 16
    pb = problems.random_batch(model, B, seed=1)
    G = _gpu(model, problems.bench_params(1), B)
    _solve_init(G, pb)
    G.Solve()
    good = torch.empty(B, model.nv, dtype=torch.float64, device="cuda")
    G.get(lk.F_Z, out=good)
    np.testing.assert_array_equal(good.cpu().numpy(), G.z)
    for bad in (torch.empty(B, model.nv, dtype=torch.float32, device="cuda"), torch.empty(model.nv, B, dtype=torch.float64, device="cuda").t(),
                torch.empty(B, model.nv + 1, dtype=torch.float64, device="cuda"), np.empty((B, model.nv), np.float32)):
        with pytest.raises(RuntimeError, match="get\\(out=\\)"):
            G.get(lk.F_Z, out=bad)
    G.close()


def test_schedule_knobs_do_not_change_results():
    """Every launch schedule is the same algorithm: dense sweeps, re-pack growth, segment kernel on / off, lane switch
    point, graph on / off give identical iteration counts and z within rounding."""
    model = robots.talos()
    B = 512
    pb = problems.random_batch(model, B, seed=17)
    params = problems.bench_params(len(pb["ids"]))
    G = _gpu(model, params, B)
    _solve_init(G, pb)
    G.Solve()
    z0, it0 = G.z, G.get_iter()
    for sc in (dict(dense_sweeps=0), dict(dense_sweeps=7, repack_reps=1, repack_growth=3.0), dict(seg_warps=1), dict(seg_warps=2, seg_after=0),
               dict(use_graph=0), dict(lane_after=5), dict(lane_after=0, lane_warps_per_cta=1), dict(drop_workspace=0),
               dict(hi_priority_after=-1, small_after=16, small_grid=64)):
        G.set_schedule(**sc)
        G.Solve()
        same = G.get_iter() == it0
        assert same.mean() >= 0.998, sc
        assert rel_inf(G.z[same], z0[same]) < 1e-8, sc
    with pytest.raises(RuntimeError, match="loik_set_schedule"):
        G.set_schedule(repack_reps=0)
    with pytest.raises(KeyError):
        G.set_schedule(lane_ctas=3)
    G.close()


def test_large_model_100_joints_16_tasks():
    """The reference has no size limit (state sized from model.njoints, loik-loid-data-optimized.hxx:40-104).  A random tree
    of 100 joints with 16 tasks (more than the 8 task matrices the parameter block holds: A goes to the task rows) and
    per-joint references (UpdateReferences, one (H_ref, v_ref) pair per joint up to the table size): fused steps and full
    solves against the oracle."""
    from oracle import recursion
    from tests.helpers import check_abs_or_rel
    model = robots.random_tree(100, seed=123, branching=0.15)
    assert model.nj == 101
    rng = np.random.default_rng(7)
    nc = 16
    tasks = sorted(rng.choice(np.arange(5, model.nj), size=nc, replace=False).tolist())
    B = 40
    pb = problems.random_batch(model, B, seed=3, task_joints=tasks)
    pb["Ais"] = np.eye(6)[None] + 0.2 * rng.standard_normal((nc, 6, 6))
    params = dict(problems.FIXTURE_PARAMS, max_iter=60, num_eq_c=nc)
    # 20 distinct references spread over the joints (the table holds 33)
    H_pool = [np.eye(6) + 0.1 * (lambda a: a @ a.T)(rng.standard_normal((6, 6))) for _ in range(20)]
    v_pool = 0.05 * rng.standard_normal((20, 6))
    pick = rng.integers(0, 20, size=model.nj)
    H_refs, v_refs = np.stack([H_pool[k] for k in pick]), v_pool[pick]
    G = _gpu(model, params, B)
    assert G.get_schedule()["lane_available"] in (0, 1)
    G.set_debug(True)
    _solve_init(G, pb)
    G.UpdateReferences(H_refs, v_refs)
    G.ResetRecursion()
    O = []
    for i in range(B):
        o = recursion.FirstOrderLoikOptimized(model, **ctor_kwargs(params))
        o.SolveInit(pb["q"][i], pb["H_ref"], pb["v_ref"], pb["ids"], pb["Ais"], pb["bis"][i], pb["lb"], pb["ub"])
        o.UpdateReferences(H_refs, v_refs)
        o.ResetSolver()
        O.append(o)
    for it in (1, 2):
        stopped = []
        for o in O:
            o.UpdatePrev(); o.ResetInfNorms(); o.FwdPass1(); o.BwdPassOptimizedVisitor(); o.FwdPass2OptimizedVisitor(); o.BoxProj(); o.DualUpdate()
            o.ComputeResiduals(); o.CheckConvergence()
            if it > 1:
                o.CheckFeasibility()
            stopped.append(o.get_convergence_status() or o.get_primal_infeasibility_status())
            if not stopped[-1]:
                o.UpdateMu()  # (the loop of Solve() leaves before UpdateMu when a flag is raised, hpp:421-446)
        G.StepBackward(); G.StepForward(); G.StepResidual()
        H, nu, z, y, mu = G.His, G.nu, G.z, G.yis, G.get_mu()
        for i, o in enumerate(O):
            check_abs_or_rel(H[i], o.His[1:], 1e-10, f"it{it} His")
            check_abs_or_rel(nu[i], o.nu, 1e-10, f"it{it} nu")
            check_abs_or_rel(z[i], o.z, 1e-10, f"it{it} z")
            check_abs_or_rel(y[i], o.yis, 1e-10, f"it{it} yis")
            assert mu[i] == o.get_mu()
        if any(stopped):
            break
    G.set_debug(False)
    G.Solve()
    it, zz = G.get_iter(), G.z
    for i, o in enumerate(O):
        o.Solve()
        assert it[i] == o.get_iter(), f"instance {i}"
        assert rel_inf(zz[i], o.z) < 1e-6
    G.close()
    too_big = robots.random_tree(104, seed=1)
    with pytest.raises(RuntimeError, match="njoints out of range"):
        _gpu(too_big, problems.bench_params(1), 4)


@pytest.mark.parametrize("a_per", [False, True])
def test_update_eq_constraint_target_only(a_per):
    """problem_.UpdateEqConstraint(c_id, bi) (ik-id-description-optimized.hpp:224-240) through Solve(q, c_id, None, bi): a new
    target, the task keeps the matrix it was given at SolveInit (batch-shared in the parameter block, or one per instance
    in the task rows); an unknown link id fails with the reference's message."""
    from oracle import recursion
    model = robots.panda()
    B = 96
    rng = np.random.default_rng(8)
    pb = problems.random_batch(model, B, seed=31)
    A = np.eye(6) + 0.2 * rng.standard_normal((B, 1, 6, 6)) if a_per else (np.eye(6) + 0.2 * rng.standard_normal((1, 6, 6)))
    pb = dict(pb, Ais=A)
    params = dict(problems.bench_params(1), warm_start=True)
    G = _gpu(model, params, B)
    G.Solve(pb["q"], pb["H_ref"], pb["v_ref"], pb["ids"], pb["Ais"], pb["bis"], pb["lb"], pb["ub"])
    b2 = problems.random_batch(model, B, seed=32)["bis"][:, 0]
    c_id = int(pb["ids"][0])
    G.Solve(pb["q"], c_id, None, b2)
    z, it = G.z, G.get_iter()
    bad = 0
    for i in range(B):
        Ai = A[i] if a_per else A
        o = recursion.FirstOrderLoikOptimized(model, **ctor_kwargs(params))
        o.Solve(pb["q"][i], pb["H_ref"], pb["v_ref"], pb["ids"], Ai, pb["bis"][i], pb["lb"], pb["ub"])
        o.Solve(pb["q"][i], c_id, Ai[0], b2[i])
        if o.get_iter() != it[i]:
            bad += 1
            continue
        assert rel_inf(z[i], o.z) < 1e-6, i
    assert bad == 0
    with pytest.raises(RuntimeError, match="constraint doesn't yet exist"):
        G.Solve(pb["q"], c_id - 1, None, b2)
    G.close()


@pytest.mark.parametrize("name,via_ctor", [("panda", False), ("talos", True)])
def test_logging_history_matches_the_oracle_log(name, via_ctor):
    """logging_ / LoikSolverInfo (loik-loid-optimized.hpp:406-420 and :290-306): the per-iteration primal / dual residual and mu
    of every instance, main loop and tail solve, against the log of the oracle driven the same way."""
    from oracle import recursion
    model = robots.get_robot(name)
    B = 64
    pb = problems.random_batch(model, B, seed=12)
    params = dict(problems.bench_params(len(pb["ids"]), max_iter=80), logging=via_ctor)
    G = _gpu(model, params, B)
    if not via_ctor:
        G.set_logging(True)
    _solve_init(G, pb)
    G.Solve()
    H, it = G.history(), G.get_iter()
    assert H.shape == (B, 80, 8)
    tails = 0
    for i in range(B):
        o = recursion.FirstOrderLoikOptimized(model, **ctor_kwargs(params))
        o.Solve(*instance(pb, i))
        n = o.get_iter()
        assert n == it[i] and int(o.scalar("hist_len")) == n
        h = H[i, :n]
        np.testing.assert_array_equal(h[:, 4], o.hist_mu)
        np.testing.assert_allclose(np.maximum(h[:, 0], h[:, 1]), o.hist_primal_residual, rtol=1e-8, atol=1e-12)
        np.testing.assert_allclose(np.maximum(h[:, 2], h[:, 3]), o.hist_dual_residual, rtol=1e-6, atol=1e-9)
        tails += int(h[:, 7].sum() > 0)
        assert (np.diff(h[:, 7]) >= 0).all()  # once on the infeasibility tail, an instance stays there
    assert tails > 0
    G.close()


@pytest.mark.parametrize("name,lane_after", [("panda", 32), ("panda", 0), ("talos", -1), ("talos_ff", -1)])
def test_no_task_constraints(name, lane_after):
    """num_eq_c = 0: no task block is ever touched (the record still holds one, unused); tile kernels and the lane kernel."""
    from oracle import recursion
    model = robots.get_robot(name)
    B = 200
    rng = np.random.default_rng(4)
    params = dict(problems.FIXTURE_PARAMS, max_iter=60, num_eq_c=0)
    q = model.normalize(rng.uniform(model.q_min, model.q_max, size=(B, model.nq)))
    v_ref = 0.3 * rng.normal(size=6)
    G = _gpu(model, params, B)
    if G.get_schedule()["lane_available"]:
        G.set_schedule(lane_after=lane_after)
    G.SolveInit(q, np.eye(6), v_ref, np.zeros(0, np.int32), np.zeros((0, 6, 6)), np.zeros((B, 0, 6)), -model.v_max, model.v_max)
    G.Solve()
    z, it = G.z, G.get_iter()
    for i in range(0, B, 9):
        o = recursion.FirstOrderLoikOptimized(model, **ctor_kwargs(params))
        o.Solve(q[i], np.eye(6), v_ref, np.zeros(0, np.int32), np.zeros((0, 6, 6)), np.zeros((0, 6)), -model.v_max, model.v_max)
        assert o.get_iter() == it[i], i
        assert rel_inf(z[i], o.z) < 1e-6, i
    G.close()


def test_infinite_bounds_behave_like_the_reference():
    """lb = -inf, ub = +inf (no box): BoxProj is the identity, ub^T max(dw, 0) is inf * 0 = NaN in the reference's CheckFeasibility
    (hxx:587-590) and both of its comparisons are false -- the CUDA path must take the same decisions (never "primal infeasible")."""
    from oracle import recursion
    model = robots.panda()
    B = 128
    pb = problems.random_batch(model, B, seed=17)
    inf = np.full(model.nv, np.inf)
    params = problems.bench_params(1, max_iter=60)
    G = _gpu(model, params, B)
    G.SolveInit(pb["q"], pb["H_ref"], pb["v_ref"], pb["ids"], pb["Ais"], pb["bis"], -inf, inf)
    G.Solve()
    z, it, st = G.z, G.get_iter(), G.get_status()
    assert not (st & 2).any()
    for i in range(0, B, 5):
        o = recursion.FirstOrderLoikOptimized(model, **ctor_kwargs(params))
        o.Solve(pb["q"][i], pb["H_ref"], pb["v_ref"], pb["ids"], pb["Ais"], pb["bis"][i], -inf, inf)
        assert o.get_iter() == it[i] and not o.get_primal_infeasibility_status(), i
        assert rel_inf(z[i], o.z) < 1e-6, i
    G.close()

"""Per-instance task matrices (``A_per_instance``): every instance of the batch has its own A_k, as every reference solver
object has its own problem (``UpdateEqConstraints`` / ``UpdateEqConstraint``, ik-id-description-optimized.hpp:127-218) --
e.g. a world-frame end-effector task, whose A depends on q.  Against oracle B, instance by instance: fused steps
(debug mode, every field), full solves through the tile kernels (in place, migrating) and the lane kernel, the tailored
warm-started Solve(q, c_id, Ai, bi) with a per-instance Ai.  -m gpu."""
import numpy as np
import pytest

from loik_b200 import problems, robots
from tests.helpers import check_abs_or_rel, ctor_kwargs, rel_inf

pytestmark = pytest.mark.gpu


def _gpu(model, params, batch, **schedule):
    from loik_b200 import solver
    G = solver.make_solver(model, params, batch)
    if schedule:
        G.set_schedule(**schedule)
    return G


def _oracle(model, params):
    from oracle import recursion
    return recursion.FirstOrderLoikOptimized(model, **ctor_kwargs(params))


def _problem(model, B, seed, nc):
    rng = np.random.default_rng(seed)
    pb = problems.random_batch(model, B, seed=seed, task_joints=robots.TASK_JOINTS.get(model.name, [model.nj - 1])[:nc])
    assert len(pb["ids"]) == nc
    A = np.eye(6)[None, None] + 0.3 * rng.standard_normal((B, nc, 6, 6))  # general (non-symmetric) matrices, one per instance and task
    return dict(pb, Ais=A)


@pytest.mark.parametrize("name,nc", [("panda", 1), ("talos", 2), ("talos_ff", 2)])
def test_fused_steps_with_per_instance_A(name, nc):
    model = robots.get_robot(name)
    B = 20
    pb = _problem(model, B, 41, nc)
    params = problems.bench_params(nc)
    G = _gpu(model, params, B)
    G.set_debug(True)
    G.SolveInit(pb["q"], pb["H_ref"], pb["v_ref"], pb["ids"], pb["Ais"], pb["bis"], pb["lb"], pb["ub"])
    G.ResetRecursion()
    O = []
    for i in range(B):
        o = _oracle(model, params)
        o.SolveInit(pb["q"][i], pb["H_ref"], pb["v_ref"], pb["ids"], pb["Ais"][i], pb["bis"][i], pb["lb"], pb["ub"])
        o.ResetSolver()
        O.append(o)
    tol = 1e-10 if name != "talos_ff" else 5e-9
    for it in range(1, 4):
        for o in O:
            o.UpdatePrev(); o.ResetInfNorms(); o.FwdPass1(); o.BwdPassOptimizedVisitor()
        G.StepBackward()
        H, p = G.His, G.pis
        for i, o in enumerate(O):
            check_abs_or_rel(H[i], o.His[1:], tol, f"it{it} His")
            check_abs_or_rel(p[i], o.pis[1:], tol, f"it{it} pis")
        for o in O:
            o.FwdPass2OptimizedVisitor(); o.BoxProj(); o.DualUpdate()
        G.StepForward()
        nu, v, z, w, y, Aty = G.nu, G.vis, G.z, G.w, G.yis, G.Aty
        for i, o in enumerate(O):
            check_abs_or_rel(nu[i], o.nu, tol, f"it{it} nu")
            check_abs_or_rel(v[i], o.vis[1:], tol, f"it{it} vis")
            check_abs_or_rel(z[i], o.z, tol, f"it{it} z")
            check_abs_or_rel(w[i], o.w, tol, f"it{it} w")
            check_abs_or_rel(y[i], o.yis, tol, f"it{it} yis")
            check_abs_or_rel(Aty[i], o.Aty, tol, f"it{it} Aty")
        for o in O:
            o.ComputeResiduals(); o.CheckConvergence()
            if it > 1:
                o.CheckFeasibility()
        G.StepResidual()
        res, mu = G.get(18), None
        for i, o in enumerate(O):
            check_abs_or_rel(res[i, 0], o.get_primal_residual(), tol, f"it{it} primal residual")
            check_abs_or_rel(res[i, 2], o.get_tol_primal(), tol, f"it{it} tol_primal")
            if not (o.get_convergence_status() or o.get_primal_infeasibility_status()):
                o.UpdateMu()
        mu = G.get_mu()
        if any(o.get_convergence_status() or o.get_primal_infeasibility_status() for o in O):
            break
        for i, o in enumerate(O):
            assert mu[i] == o.get_mu()
    G.close()


@pytest.mark.parametrize("name,nc,B,schedule", [
    ("panda", 1, 768, dict(lane_after=-1)),                    # tile kernels: in place, then migrating launches
    ("panda", 1, 768, dict(lane_after=0)),                     # lane kernel: task matrices in the instance record
    ("panda", 1, 768, dict(lane_after=6)),                     # hand-over from a packed arena
    ("talos", 2, 256, dict(lane_after=-1)),                    # segment kernel for the late rounds
    ("talos", 2, 256, dict(lane_after=0, lane_groups_per_instance=4)),
    ("ur10", 1, 512, dict(lane_after=-1, dense_sweeps=0)),
])
def test_full_solves_with_per_instance_A(name, nc, B, schedule):
    model = robots.get_robot(name)
    pb = _problem(model, B, 43, nc)
    params = problems.bench_params(nc)
    G = _gpu(model, params, B, **schedule)
    G.SolveInit(pb["q"], pb["H_ref"], pb["v_ref"], pb["ids"], pb["Ais"], pb["bis"], pb["lb"], pb["ub"])
    G.Solve()
    z, it, mu, st, y = G.z, G.get_iter(), G.get_mu(), G.get_status(), G.yis
    diverged = 0
    for i in range(B):
        o = _oracle(model, params)
        o.Solve(pb["q"][i], pb["H_ref"], pb["v_ref"], pb["ids"], pb["Ais"][i], pb["bis"][i], pb["lb"], pb["ub"])
        if o.get_iter() != it[i] or o.get_mu() != mu[i]:
            diverged += 1
            continue
        assert rel_inf(z[i], o.z) < 1e-6 and rel_inf(y[i], o.yis) < 1e-6, f"instance {i}"
        assert bool(st[i] & 1) == o.get_convergence_status()
    assert diverged <= 0.004 * B, f"{diverged} diverged decision traces"
    assert it.max() > 12, "some instances must outlive the dense sweeps (migrating / lane path exercised)"
    # a second Solve() reproduces itself (the per-instance matrices survive the re-packs)
    G.Solve()
    np.testing.assert_array_equal(G.get_iter(), it)
    np.testing.assert_array_equal(G.z, z)
    G.close()


@pytest.mark.parametrize("lane_after", [-1, 0])
def test_tailored_solve_with_per_instance_Ai(lane_after):
    """Solve(q, c_id, Ai, bi) (hpp:596-695) warm-started, Ai [B, 6, 6]: the trajectory-tracking call of a world-frame task."""
    model = robots.panda()
    B = 64
    rng = np.random.default_rng(8)
    pb = _problem(model, B, 44, 1)
    params = dict(problems.bench_params(1), warm_start=True)
    G = _gpu(model, params, B, lane_after=lane_after)
    G.SolveInit(pb["q"], pb["H_ref"], pb["v_ref"], pb["ids"], pb["Ais"], pb["bis"], pb["lb"], pb["ub"])
    G.Solve()
    A2 = pb["Ais"][:, 0] + 0.05 * rng.standard_normal((B, 6, 6))
    b2 = pb["bis"][:, 0] + 0.05 * rng.standard_normal((B, 6))
    q2 = pb["q"] + 0.01 * rng.standard_normal(pb["q"].shape)
    G.Solve(q2, int(pb["ids"][0]), A2, b2)
    z, it = G.z, G.get_iter()
    for i in range(B):
        o = _oracle(model, params)
        o.SolveInit(pb["q"][i], pb["H_ref"], pb["v_ref"], pb["ids"], pb["Ais"][i], pb["bis"][i], pb["lb"], pb["ub"])
        o.Solve()
        o.Solve(q2[i], int(pb["ids"][0]), A2[i], b2[i])
        assert it[i] == o.get_iter(), f"instance {i}"
        assert rel_inf(z[i], o.z) < 1e-6
    # mixing the two forms on one handle is an error, not a silent reinterpretation
    with pytest.raises(RuntimeError, match="per instance exactly when"):
        G.Solve(q2, int(pb["ids"][0]), np.eye(6), b2)
    G.close()


def test_device_resident_per_instance_A():
    """q, b and A as CUDA tensors (no host staging), results equal to the host-array call."""
    import torch
    model = robots.ur10()
    B = 300
    pb = _problem(model, B, 45, 1)
    params = problems.bench_params(1)
    G = _gpu(model, params, B)
    G.SolveInit(pb["q"], pb["H_ref"], pb["v_ref"], pb["ids"], pb["Ais"], pb["bis"], pb["lb"], pb["ub"])
    G.Solve()
    z, it = G.z, G.get_iter()
    dev = lambda a: torch.as_tensor(a, device="cuda")
    G.Solve(dev(pb["q"]), pb["H_ref"], pb["v_ref"], pb["ids"], dev(pb["Ais"]), dev(pb["bis"]), pb["lb"], pb["ub"])
    np.testing.assert_array_equal(G.get_iter(), it)
    np.testing.assert_array_equal(G.z, z)
    G.close()

"""The reference's component-wise test, method by method (/root/reference/tests/loik-loid.cpp:305-556), with the CUDA
solver in the role of the optimized solver and the oracle as the ground truth: after each of
FwdPass1 / BwdPassOptimizedVisitor / FwdPass2OptimizedVisitor / BoxProj / DualUpdate / ComputeResiduals /
CheckConvergence / CheckFeasibility / UpdateMu the same fields the reference test compares must agree.  -m gpu."""
import numpy as np
import pytest

from loik_b200 import problems, robots
from tests.helpers import check_abs_or_rel, ctor_kwargs, instance

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _run(model, params, pb, n_iters, what):
    from loik_b200 import solver
    from oracle import recursion
    B = pb["q"].shape[0]
    G = solver.make_solver(model, params, B)
    G.set_debug(True)
    G.SolveInit(pb["q"], pb["H_ref"], pb["v_ref"], pb["ids"], pb["Ais"], pb["bis"], pb["lb"], pb["ub"])
    G.ResetRecursion()
    O = []
    for i in range(B):
        o = recursion.FirstOrderLoikOptimized(model, **ctor_kwargs(params))
        o.SolveInit(*instance(pb, i))
        o.ResetSolver()
        O.append(o)
    for it in range(1, n_iters + 1):
        tag = f"{what} it{it}"
        G.UpdatePrev(); G.ResetInfNorms()
        for o in O:
            o.UpdatePrev(); o.ResetInfNorms()
        # FwdPass1 (:340-362)
        G.FwdPass1()
        H, p, r = G.His, G.pis, G.r
        for i, o in enumerate(O):
            o.FwdPass1()
            check_abs_or_rel(H[i], o.His[1:], TOL, tag + " FwdPass1 His")
            check_abs_or_rel(p[i], o.pis[1:], TOL, tag + " FwdPass1 pis")
            check_abs_or_rel(r[i], o.r, TOL, tag + " FwdPass1 r")
        # BwdPass (:366-372)
        G.BwdPassOptimizedVisitor()
        H, p, r, Di, UD = G.His, G.pis, G.r, G.Dinv, G.UDinv
        for i, o in enumerate(O):
            o.BwdPassOptimizedVisitor()
            check_abs_or_rel(H[i], o.His[1:], TOL, tag + " BwdPass His")
            check_abs_or_rel(p[i], o.pis[1:], TOL, tag + " BwdPass pis")
            check_abs_or_rel(r[i], o.r, TOL, tag + " BwdPass r")
            check_abs_or_rel(Di[i], o.Dinv[1:], TOL, tag + " Dinv")
            check_abs_or_rel(UD[i], o.UDinv[1:], TOL, tag + " UDinv")
        # FwdPass2 (:377-386)
        G.FwdPass2OptimizedVisitor()
        nu, v, f, nrm = G.nu, G.vis, G.fis, G.norms()
        for i, o in enumerate(O):
            o.FwdPass2OptimizedVisitor()
            check_abs_or_rel(nu[i], o.nu, TOL, tag + " FwdPass2 nu")
            check_abs_or_rel(v[i], o.vis[1:], TOL, tag + " FwdPass2 vis")
            check_abs_or_rel(f[i], o.fis[1:], TOL, tag + " FwdPass2 fis")
            for nm in ("nu_inf_norm", "delta_vis_inf_norm", "delta_nu_inf_norm", "Href_v_inf_norm"):
                check_abs_or_rel(nrm[nm][i], o.scalar(nm), 1e-9, f"{tag} {nm}")
        # BoxProj (:389-395)
        G.BoxProj()
        z, w = G.z, G.w
        for i, o in enumerate(O):
            o.BoxProj()
            check_abs_or_rel(z[i], o.z, TOL, tag + " BoxProj z")
            check_abs_or_rel(w[i], o.w, TOL, tag + " BoxProj w (unchanged)")
        # DualUpdate (:398-406)
        G.DualUpdate()
        w, y = G.w, G.yis
        for i, o in enumerate(O):
            o.DualUpdate()
            check_abs_or_rel(w[i], o.w, TOL, tag + " DualUpdate w")
            check_abs_or_rel(y[i], o.yis, TOL, tag + " DualUpdate yis")
        # ComputeResiduals (:410-418)
        G.ComputeResiduals()
        prv, drv, res = G.get_primal_residual_vec(), G.get_dual_residual_vec(), G.get(18)
        for i, o in enumerate(O):
            o.ComputeResiduals()
            fscale = 10 * max(1.0, np.abs(o.fis).max(), np.abs(o.Aty).max(), np.abs(o.yis).max())  # the dual residual is a difference of these
            check_abs_or_rel(prv[i], o.get_primal_residual_vec(), TOL, tag + " primal_residual_vec")
            assert np.abs(drv[i] - o.get_dual_residual_vec()).max() < 1e-11 * fscale, tag + " dual_residual_vec"
            check_abs_or_rel(res[i, 0], o.get_primal_residual(), TOL, tag + " primal_residual")
            assert abs(res[i, 1] - o.get_dual_residual()) < 1e-11 * fscale, tag + " dual_residual"
        # CheckConvergence (:420-433)
        G.CheckConvergence()
        res, nrm = G.get(18), G.norms()
        for i, o in enumerate(O):
            o.CheckConvergence()
            assert res[i, 2] != 0.0 and res[i, 3] != 0.0
            check_abs_or_rel(res[i, 2], o.get_tol_primal(), TOL, tag + " tol_primal")
            check_abs_or_rel(res[i, 3], o.get_tol_dual(), 1e-9, tag + " tol_dual")
            assert bool(nrm["converged"][i]) == o.get_convergence_status(), tag + " converged"
        # CheckFeasibility (:436-472)
        if it > 1:
            G.CheckFeasibility()
            nrm = G.norms()
            for i, o in enumerate(O):
                o.CheckFeasibility()
                fscale = 10 * max(1.0, np.abs(o.fis).max(), np.abs(o.Aty).max(), np.abs(o.yis).max())
                check_abs_or_rel(nrm["delta_y_qp_inf_norm"][i], o.get_delta_y_qp_inf_norm(), 1e-9 * fscale, tag + " delta_y_qp")
                assert abs(nrm["A_qp_T_delta_y_qp_inf_norm"][i] - o.get_A_qp_T_delta_y_qp_inf_norm()) < 1e-11 * fscale
                check_abs_or_rel(nrm["ub_qp_T_delta_y_qp_plus"][i], o.get_ub_qp_T_delta_y_qp_plus(), 1e-9, tag + " ub^T dy+")
                check_abs_or_rel(nrm["lb_qp_T_delta_y_qp_minus"][i], o.get_lb_qp_T_delta_y_qp_minus(), 1e-9, tag + " lb^T dy-")
                assert bool(nrm["primal_infeasibility_cond_1"][i]) == o.get_primal_infeasibility_cond_1(), tag
                assert bool(nrm["primal_infeasibility_cond_2"][i]) == o.get_primal_infeasibility_cond_2(), tag
                assert bool(nrm["primal_infeasible"][i]) == o.get_primal_infeasibility_status(), tag
                check_abs_or_rel(nrm["delta_x_qp_inf_norm"][i], o.get_delta_x_qp_inf_norm(), 1e-9, tag + " delta_x_qp")
                check_abs_or_rel(nrm["delta_z_inf_norm"][i], o.get_delta_z_qp_inf_norm(), 1e-9, tag + " delta_z_qp")
        # UpdateMu (:475-478)
        G.UpdateMu()
        mu = G.get_mu()
        for i, o in enumerate(O):
            o.UpdateMu()
            assert mu[i] == o.get_mu(), tag + " mu"
    G.close()


@pytest.mark.parametrize("name,bound", [("talos", 1.0), ("panda", 1.0), ("ur10", 1.0)])
def test_per_method_fixture(name, bound):
    model = robots.get_robot(name)
    pr = problems.fixture_problem(model, bound)
    _run(model, dict(problems.FIXTURE_PARAMS, max_iter=200), dict(pr, q=pr["q"][None], bis=pr["bis"][None]), 4, name)


@pytest.mark.parametrize("name", ["panda9", "talos"])
def test_per_method_random_batch(name):
    model = robots.get_robot(name)
    pb = problems.random_batch(model, 37, seed=17)
    _run(model, problems.bench_params(len(pb["ids"])), pb, 3, name)


def test_per_method_steps_need_debug_mode():
    from loik_b200 import solver
    model = robots.panda()
    pb = problems.random_batch(model, 4, seed=0)
    G = solver.make_solver(model, problems.bench_params(1), 4)
    G.SolveInit(pb["q"], pb["H_ref"], pb["v_ref"], pb["ids"], pb["Ais"], pb["bis"], pb["lb"], pb["ub"])
    with pytest.raises(RuntimeError, match="loik_set_debug"):
        G.FwdPass1()
    G.close()


def test_fwd_pass_init_alone():
    """FwdPassInit(q) (hxx:253-283): liMi = jointPlacement * M(q) for a new configuration."""
    from loik_b200 import solver
    from oracle import recursion
    model = robots.panda(fingers=True)
    B = 9
    pb = problems.random_batch(model, B, seed=1)
    pb2 = problems.random_batch(model, B, seed=2)
    params = problems.bench_params(1)
    G = solver.make_solver(model, params, B)
    G.SolveInit(pb["q"], pb["H_ref"], pb["v_ref"], pb["ids"], pb["Ais"], pb["bis"], pb["lb"], pb["ub"])
    G.FwdPassInit(pb2["q"])
    L = G.liMi
    for i in range(B):
        o = recursion.FirstOrderLoikOptimized(model, **ctor_kwargs(params))
        o.SolveInit(*instance(pb, i))
        o.FwdPassInit(pb2["q"][i])
        np.testing.assert_allclose(L[i][:, :9].reshape(-1, 3, 3), o.liMi_R[1:], rtol=0, atol=1e-14)
        np.testing.assert_allclose(L[i][:, 9:], o.liMi_p[1:], rtol=0, atol=1e-14)
    G.close()

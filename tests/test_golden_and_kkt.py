"""Oracle vs the committed golden fixtures (tests/golden, made by scripts/make_golden.py) and KKT checks.

The reference's tests contain no expected-output numbers; the golden files freeze the
oracle pair's agreed outputs, and the KKT checks pin converged solutions against the QP itself
(``ik-id-description.hpp:411-491``), independently of either ADMM implementation.
"""
import os

import numpy as np
import pytest

from loik_b200 import problems, robots
from oracle import dense, recursion
from tests.helpers import ctor_kwargs, instance, prob_args, rel_inf

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", ["talos", "panda", "ur10"])
def test_fixture_golden(name):
    g = np.load(os.path.join(GOLD, f"fixture_{name}.npz"))
    model = robots.get_robot(name)
    B = recursion.FirstOrderLoikOptimized(model, **ctor_kwargs(dict(problems.FIXTURE_PARAMS, max_iter=int(g["max_iter"]))))
    B.Solve(*prob_args(problems.fixture_problem(model, float(g["bound"]))))
    assert B.get_iter() == int(g["iter"]) and B.get_mu() == float(g["mu"])
    assert B.get_convergence_status() == bool(g["converged"])
    assert B.get_primal_infeasibility_status() == bool(g["primal_infeasible"])
    for nm in ("z", "nu", "w", "yis", "vis"):
        assert rel_inf(getattr(B, nm), g[nm]) < 1e-9, nm


@pytest.mark.parametrize("name", ["panda", "ur10", "talos", "panda9", "ur10c", "tree_zyx"])
def test_random_golden(name):
    g = np.load(os.path.join(GOLD, f"random_{name}.npz"))
    model = robots.get_robot(name)
    n = int(g["n"])
    pb = problems.random_batch(model, n, seed=int(g["seed"]))
    np.testing.assert_array_equal(pb["q"], g["q"])      # the generator itself is pinned
    np.testing.assert_array_equal(pb["bis"], g["bis"])
    params = problems.bench_params(len(pb["ids"]))
    out = recursion.batch_solve(model, params, pb["q"], pb["H_ref"], pb["v_ref"], pb["ids"], pb["Ais"], pb["bis"], pb["lb"],
                                pb["ub"], nthreads=2)
    np.testing.assert_array_equal(out["iters"], g["iter"])
    np.testing.assert_array_equal(out["mu"], g["mu"])
    np.testing.assert_array_equal(out["status"] & 1, g["converged"].astype(int))
    for i in range(n):
        assert rel_inf(out["z"][i], g["z"][i]) < 1e-7
        assert rel_inf(out["w"][i], g["w"][i]) < 1e-7


@pytest.mark.parametrize("name", ["panda", "ur10", "talos"])
def test_kkt_of_converged_solutions(name):
    """A converged point satisfies the KKT system of the QP to the solver's tolerance scale."""
    model = robots.get_robot(name)
    pb = problems.random_batch(model, 24, seed=21)
    params = dict(problems.bench_params(len(pb["ids"]), max_iter=400), tol_abs=1e-7, tol_rel=1e-7)
    n_checked = 0
    for i in range(24):
        B = recursion.FirstOrderLoikOptimized(model, **ctor_kwargs(params))
        B.Solve(*instance(pb, i))
        if not B.get_convergence_status():
            continue
        n_checked += 1
        liMi = list(zip(B.liMi_R, B.liMi_p))
        rep = dense.kkt_report(model, liMi, pb["ids"], pb["Ais"], pb["bis"][i], pb["H_ref"], pb["v_ref"], pb["lb"], pb["ub"],
                               B.vis, B.nu, B.z, B.fis, B.yis, B.w)
        scale = max(1.0, np.abs(B.fis).max(), np.abs(B.yis).max())
        assert rep["kinematics"] < 1e-10, rep
        assert rep["task"] < 1e-5, rep
        assert rep["slack"] < 1e-5, rep
        assert rep["box"] == 0.0, rep
        assert rep["stationarity_v"] < 1e-5 * scale, rep
        assert rep["stationarity_nu"] < 1e-5 * scale, rep
        assert rep["complementarity"] < 1e-4 * scale, rep
    assert n_checked >= 6


def test_kkt_cross_solve_scipy():
    """Small case: the converged z equals an independent dense QP solve (scipy SLSQP on the reduced problem in nu)."""
    from scipy.optimize import minimize
    model = robots.ur10()
    pb = problems.random_batch(model, 6, seed=33)
    params = dict(problems.bench_params(1, max_iter=2000), tol_abs=1e-9, tol_rel=1e-9)
    done = 0
    for i in range(6):
        B = recursion.FirstOrderLoikOptimized(model, **ctor_kwargs(params))
        B.Solve(*instance(pb, i))
        if not B.get_convergence_status():
            continue
        # Jacobians J_i (v_i = J_i nu) from the kinematics recursion
        n = model.nv
        J = np.zeros((model.nj, 6, n))
        for j in range(1, model.nj):
            X = np.linalg.inv(dense.action_matrix(B.liMi_R[j], B.liMi_p[j]))
            J[j] = X @ J[int(model.parent[j])]
            J[j][:, j - 1] += dense.joint_subspace(int(model.jtype[j]), model.axis[j])[:, 0]
        c = int(pb["ids"][0])
        A_eq, b_eq = pb["Ais"][0] @ J[c], pb["bis"][i][0]
        Hs = sum(J[j].T @ J[j] for j in range(1, model.nj))
        res = minimize(lambda x: 0.5 * x @ Hs @ x, np.zeros(n), jac=lambda x: Hs @ x, method="SLSQP",
                       bounds=list(zip(pb["lb"], pb["ub"])), constraints=[{"type": "eq", "fun": lambda x: A_eq @ x - b_eq,
                                                                          "jac": lambda x: A_eq}],
                       options={"ftol": 1e-14, "maxiter": 500})
        if not res.success:
            continue
        assert np.abs(res.x - B.z).max() < 1e-5, (res.x, B.z)
        done += 1
    assert done >= 1

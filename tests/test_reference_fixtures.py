"""Golden vectors produced by the REFERENCE BINARY (oracle/ref_recipe/dump_fixtures.cpp, run on a machine that has
Pinocchio) against oracle B and, on a GPU box, the CUDA path.  The files cannot be produced in the offline build
container: while they are absent these tests are skipped with a loud reason and parity stays "unpinned" (DESIGN.md 6)."""
import json
import os

import numpy as np
import pytest

from loik_b200 import problems, robots
from tests.helpers import check_abs_or_rel, ctor_kwargs

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["ref_talos_fixture", "ref_panda_q"]
NAME2CODE = {"JointModelRX": 0, "JointModelRY": 1, "JointModelRZ": 2, "JointModelPX": 3, "JointModelPY": 4, "JointModelPZ": 5,
             "JointModelRevoluteUnaligned": 6, "JointModelPrismaticUnaligned": 7, "JointModelFreeFlyer": 8, "JointModelRUBX": 9,
             "JointModelRUBY": 10, "JointModelRUBZ": 11, "JointModelRevoluteUnboundedUnaligned": 12, "JointModelSpherical": 13,
             "JointModelTranslation": 14, "JointModelPlanar": 15}


def _load(case):
    path = os.path.join(GOLDEN, case + ".json")
    if not os.path.exists(path):
        pytest.skip(f"reference fixtures absent: parity unpinned ({path} is produced by oracle/ref_recipe on a machine with Pinocchio)")
    with open(path) as f:
        return json.load(f)


def _model_and_problem(d):
    nj = d["njoints"]
    model = robots.RobotModel.from_tables(
        name="ref", parent=np.array(d["parents"], np.int32), jtype=np.array([0] + [NAME2CODE[s] for s in d["joint_shortnames"][1:]], np.int32),
        axis=np.array(d["joint_axes"]).reshape(nj, 3), placement_R=np.array(d["placement_R"]).reshape(nj, 3, 3),
        placement_p=np.array(d["placement_p"]).reshape(nj, 3))
    assert model.nq == d["nq"] and model.nv == d["nv"]
    pr = dict(q=np.array(d["q"]), H_ref=np.eye(6), v_ref=np.zeros(6), ids=np.array([d["task_joint"]], np.int32), Ais=np.eye(6)[None],
              bis=np.array(d["b"])[None], lb=np.array(d["lb"]), ub=np.array(d["ub"]))
    params = dict(problems.FIXTURE_PARAMS, **{k: d[k] for k in ("tol_abs", "tol_rel", "tol_primal_inf", "tol_dual_inf", "tol_tail_solve",
                                                                   "rho", "mu", "mu_equality_scale_factor")})
    return model, pr, params


def _compare(d, p, got, what):
    nb = d["njoints"] - 1
    assert got["iter"] == d[p + "iter"], what + " iteration count"
    assert got["mu"] == d[p + "mu"], what + " mu"
    for k, shape in (("z", None), ("nu", None), ("w", None), ("yis", (1, 6)), ("vis", (nb, 6)), ("fis", (nb, 6)), ("His", (nb, 6, 6)),
                     ("pis", (nb, 6))):
        ref = np.array(d[p + k])
        check_abs_or_rel(np.asarray(got[k]).reshape(ref.shape), ref, 1e-10, f"{what} {k}")
    check_abs_or_rel(got["primal_residual"], d[p + "primal_residual"], 1e-10, what + " primal residual")


@pytest.mark.parametrize("case", CASES)
def test_oracle_vs_reference_binary(case):
    from oracle import recursion
    d = _load(case)
    model, pr, params = _model_and_problem(d)
    for m in d["max_iters"]:
        o = recursion.FirstOrderLoikOptimized(model, **ctor_kwargs(dict(params, max_iter=m)))
        o.Solve(pr["q"], pr["H_ref"], pr["v_ref"], pr["ids"], pr["Ais"], pr["bis"][0], pr["lb"], pr["ub"])
        got = dict(iter=o.get_iter(), mu=o.get_mu(), z=o.z, nu=o.nu, w=o.w, yis=o.yis, vis=o.vis[1:], fis=o.fis[1:], His=o.His[1:],
                   pis=o.pis[1:], primal_residual=o.get_primal_residual())
        _compare(d, f"m{m}_", got, f"{case} max_iter {m}: oracle")
        R, t = o.liMi_R[1:], o.liMi_p[1:]
        check_abs_or_rel(R, np.array(d[f"m{m}_liMi_R"]).reshape(R.shape), 1e-12, case + " liMi rotation (Pinocchio conventions)")
        check_abs_or_rel(t, np.array(d[f"m{m}_liMi_p"]).reshape(t.shape), 1e-12, case + " liMi translation")


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_cuda_vs_reference_binary(case):
    from loik_b200 import solver
    d = _load(case)
    model, pr, params = _model_and_problem(d)
    for m in d["max_iters"]:
        G = solver.make_solver(model, dict(params, max_iter=m), 1)
        G.set_keep_workspace(True)
        G.Solve(pr["q"][None], pr["H_ref"], pr["v_ref"], pr["ids"], pr["Ais"], pr["bis"][None], pr["lb"], pr["ub"])
        got = dict(iter=int(G.get_iter()[0]), mu=float(G.get_mu()[0]), z=G.z[0], nu=G.nu[0], w=G.w[0], yis=G.yis[0], vis=G.vis[0], fis=G.fis[0],
                   His=G.His[0], pis=G.pis[0], primal_residual=float(G.get_primal_residual()[0]))
        _compare(d, f"m{m}_", got, f"{case} max_iter {m}: CUDA")
        G.close()

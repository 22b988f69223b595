#!/usr/bin/env python
"""Generate tests/golden/*.npz: frozen input/output vectors of the oracle pair.

The reference holds no golden vectors and cannot be built or imported offline, so
these fixtures freeze the outputs of oracle B (oracle/loik_oracle.c) -- each one cross-checked here against
the dense oracle A (oracle/dense.py) before it is written.  They pin the oracle against silent drift and give
the GPU tests a reference that does not depend on the build host's libm/gcc.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from loik_b200 import problems, robots  # noqa: E402
from oracle import dense, recursion  # noqa: E402
from tests.helpers import ctor_kwargs, instance, prob_args  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def solve_both(model, params, args):
    kw = ctor_kwargs(params)
    A = dense.FirstOrderLoik(model, **kw)
    B = recursion.FirstOrderLoikOptimized(model, **kw)
    A.Solve(*args)
    B.Solve(*args)
    assert A.iter == B.get_iter() and A.mu == B.get_mu()
    assert np.abs(A.z - B.z).max() < 1e-8 and np.abs(A.nu - B.nu).max() < 1e-8
    return B


ONLY = set(sys.argv[1:])  # optional: regenerate only the named random_* fixtures


def main():
    os.makedirs(OUT, exist_ok=True)
    # 1. the reference's fixture (tests/loik-loid.cpp:87-165) on the synthetic tables, bounds as in :559-671
    for name in ("talos", "panda", "ur10"):
        if ONLY:
            continue
        model = robots.get_robot(name)
        pr = problems.fixture_problem(model, 2.0)
        params = dict(problems.FIXTURE_PARAMS, max_iter=8)
        B = solve_both(model, params, prob_args(pr))
        np.savez(os.path.join(OUT, f"fixture_{name}.npz"), robot=name, bound=2.0, max_iter=8, z=B.z, nu=B.nu, w=B.w, yis=B.yis,
                 vis=B.vis, fis=B.fis, iter=B.get_iter(), mu=B.get_mu(), converged=B.get_convergence_status(),
                 primal_infeasible=B.get_primal_infeasibility_status(), primal_residual=B.get_primal_residual(),
                 dual_residual=B.get_dual_residual())
    # 2. seeded random instances of the BASELINE configs, full solves (max_iter = 200)
    for name, n in (("panda", 48), ("ur10", 48), ("talos", 16), ("panda9", 16), ("ur10c", 32), ("tree_zyx", 24)):
        if ONLY and name not in ONLY:
            continue
        model = robots.get_robot(name)
        pb = problems.random_batch(model, n, seed=1234)
        params = problems.bench_params(len(pb["ids"]))
        rec = {k: [] for k in ("z", "nu", "w", "yis", "iter", "mu", "converged", "primal_infeasible")}
        for i in range(n):
            B = solve_both(model, params, instance(pb, i))
            rec["z"].append(B.z); rec["nu"].append(B.nu); rec["w"].append(B.w); rec["yis"].append(B.yis)
            rec["iter"].append(B.get_iter()); rec["mu"].append(B.get_mu())
            rec["converged"].append(B.get_convergence_status()); rec["primal_infeasible"].append(B.get_primal_infeasibility_status())
        np.savez(os.path.join(OUT, f"random_{name}.npz"), robot=name, seed=1234, n=n, q=pb["q"], bis=pb["bis"],
                 **{k: np.array(v) for k, v in rec.items()})
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()

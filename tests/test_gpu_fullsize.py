"""Full-size (BASELINE.json) parity, plus ragged / edge-case batches.  -m gpu.

At 65 536 / 262 144 / 16 384 instances the CUDA path is checked (1) instance by instance against the oracle (all of them:
the C oracle solves a full-size batch in about a second on the host cores) and (2) through properties that hold for
every instance regardless of size:
permutation equivariance (an instance's result does not depend on its slot, its tile neighbours or the re-packing
order), sub-batch invariance, idempotence of repeated solves, feasibility of z (box) and the KKT-style residual
identities the iterates satisfy by construction.
"""
import numpy as np
import pytest

from loik_b200 import problems, robots
from tests.helpers import ctor_kwargs, rel_inf, rel_inf_rows

pytestmark = pytest.mark.gpu

FULL = [("panda", 65536), ("ur10", 262144), ("talos", 16384), ("talos_ff", 8192)]


def _gpu(model, params, batch):
    from loik_b200 import solver
    return solver.make_solver(model, params, batch)


def _solve(G, pb):
    G.SolveInit(pb["q"], pb["H_ref"], pb["v_ref"], pb["ids"], pb["Ais"], pb["bis"], pb["lb"], pb["ub"])
    G.Solve()
    return dict(z=G.z, nu=G.nu, w=G.w, y=G.yis, it=G.get_iter(), mu=G.get_mu(), st=G.get_status())


@pytest.mark.parametrize("name,B", FULL)
def test_full_size_every_instance_vs_oracle_and_properties(name, B):
    from oracle import recursion
    model = robots.get_robot(name)
    pb = problems.random_batch(model, B, seed=0)
    params = problems.bench_params(len(pb["ids"]))
    G = _gpu(model, params, B)
    r = _solve(G, pb)
    # every instance ended in a terminal state (converged = 1, primal infeasible = 2, both flags raised in the final iteration = 3,
    # stopped at max_iter = 4) and within the iteration budget
    assert ((r["st"] == 1) | (r["st"] == 2) | (r["st"] == 3) | (r["st"] == 4)).all()
    assert r["it"].min() >= 1 and r["it"].max() <= params["max_iter"]
    assert (r["it"][r["st"] == 4] == params["max_iter"] - 1).all()
    # z is the box projection: always inside the bounds, and equal to nu wherever w vanished
    assert (r["z"] <= pb["ub"] + 0).all() and (r["z"] >= pb["lb"] - 0).all()
    # complementarity sign of the slack multiplier at converged instances: w > 0 only at the upper bound, < 0 at the lower
    conv = (r["st"] & 1) > 0
    tol = 5e-3
    up = (r["w"] > tol) & conv[:, None]
    lo = (r["w"] < -tol) & conv[:, None]
    assert (np.abs(r["z"] - pb["ub"])[up] < 5e-2).all()
    assert (np.abs(r["z"] - pb["lb"])[lo] < 5e-2).all()
    # (1) EVERY instance of the full-size batch against the oracle (a second on the host cores): identical decision
    # traces (iteration count, final mu, status) and z, nu, w, y within 1e-6 rel-inf
    rng = np.random.default_rng(1)
    ref = recursion.batch_solve(model, params, pb["q"], pb["H_ref"], pb["v_ref"], pb["ids"], pb["Ais"], pb["bis"], pb["lb"], pb["ub"],
                                nthreads=8)
    same = (r["it"] == ref["iters"]) & (r["mu"] == ref["mu"]) & ((r["st"] & 3) == (ref["status"] & 3))
    print(f"[{name} x {B}] diverged decision traces: {int((~same).sum())}")
    assert (~same).sum() <= max(2, B // 20000), f"{(~same).sum()} diverged decision traces of {B}"
    worst = max(rel_inf_rows(r[k][same], ref[k][same]).max() for k in ("z", "nu", "w", "y"))
    print(f"[{name} x {B}] worst rel-inf over z, nu, w, y of {int(same.sum())} instances: {worst:.3e}")
    assert worst < 1e-6
    # (2) idempotence: the same handle solving again reproduces itself bit for bit
    G.Solve()
    np.testing.assert_array_equal(G.z, r["z"])
    np.testing.assert_array_equal(G.get_iter(), r["it"])
    # (3) permutation equivariance (different slots, tiles, re-pack order): bit-identical per instance
    perm = rng.permutation(B)
    pbp = dict(pb, q=pb["q"][perm], bis=pb["bis"][perm])
    rp = _solve(G, pbp)
    np.testing.assert_array_equal(rp["z"], r["z"][perm])
    np.testing.assert_array_equal(rp["w"], r["w"][perm])
    np.testing.assert_array_equal(rp["it"], r["it"][perm])
    np.testing.assert_array_equal(rp["mu"], r["mu"][perm])
    G.close()
    # (4) sub-batch invariance: a ragged slice solved alone gives the same bits
    lo_, hi_ = 1000, 1000 + 4099
    G2 = _gpu(model, params, hi_ - lo_)
    r2 = _solve(G2, dict(pb, q=pb["q"][lo_:hi_], bis=pb["bis"][lo_:hi_]))
    np.testing.assert_array_equal(r2["z"], r["z"][lo_:hi_])
    np.testing.assert_array_equal(r2["y"], r["y"][lo_:hi_])
    np.testing.assert_array_equal(r2["it"], r["it"][lo_:hi_])
    G2.close()


@pytest.mark.parametrize("B", [1, 2, 31, 32, 33, 65, 100])
def test_ragged_batches(B):
    """Batches that do not fill a tile / a CTA (the last tile is partially empty)."""
    from oracle import recursion
    model = robots.panda(fingers=True)
    pb = problems.random_batch(model, B, seed=3)
    params = problems.bench_params(1)
    G = _gpu(model, params, B)
    r = _solve(G, pb)
    ref = recursion.batch_solve(model, params, pb["q"], pb["H_ref"], pb["v_ref"], pb["ids"], pb["Ais"], pb["bis"], pb["lb"],
                                pb["ub"], nthreads=2)
    np.testing.assert_array_equal(r["it"], ref["iters"])
    np.testing.assert_array_equal(r["mu"], ref["mu"])
    for i in range(B):
        assert rel_inf(r["z"][i], ref["z"][i]) < 1e-6
    s = G.stats()
    assert s["converged"] + s["primal_infeasible"] + s["max_iter"] == B
    G.close()


@pytest.mark.parametrize("max_iter", [1, 2, 3, 5])
def test_tiny_iteration_budgets(max_iter):
    """max_iter = 2 is the reference's timing protocol: exactly one iteration per Solve() (tests/loik-loid.cpp:987-1032);
    max_iter = 1 runs none."""
    from oracle import recursion
    model = robots.ur10()
    B = 257
    pb = problems.random_batch(model, B, seed=4)
    params = problems.bench_params(1, max_iter=max_iter)
    G = _gpu(model, params, B)
    r = _solve(G, pb)
    ref = recursion.batch_solve(model, params, pb["q"], pb["H_ref"], pb["v_ref"], pb["ids"], pb["Ais"], pb["bis"], pb["lb"],
                                pb["ub"], nthreads=2)
    np.testing.assert_array_equal(r["it"], ref["iters"])
    assert r["it"].max() <= max(max_iter - 1, 0) or max_iter >= 3
    for i in range(B):
        assert rel_inf(r["z"][i], ref["z"][i]) < 1e-6
        assert rel_inf(r["w"][i], ref["w"][i]) < 1e-6
    G.close()


def test_infeasible_targets_take_the_tail_path():
    """Targets far outside the velocity limits: primal infeasibility is flagged and the tail solve runs, exactly as in
    the oracle (InfeasibilityTailSolve, loik-loid-optimized.hpp:271-319)."""
    from oracle import recursion
    model = robots.panda()
    B = 512
    pb = problems.random_batch(model, B, seed=5, b_scale=25.0)
    params = problems.bench_params(1)
    G = _gpu(model, params, B)
    r = _solve(G, pb)
    ref = recursion.batch_solve(model, params, pb["q"], pb["H_ref"], pb["v_ref"], pb["ids"], pb["Ais"], pb["bis"], pb["lb"],
                                pb["ub"], nthreads=4)
    assert ((ref["status"] & 2) > 0).mean() > 0.5          # the case really exercises the infeasible path
    same = (r["it"] == ref["iters"]) & (r["mu"] == ref["mu"]) & ((r["st"] & 3) == (ref["status"] & 3))
    assert same.mean() >= 0.99
    for i in np.nonzero(same)[0]:
        assert rel_inf(r["z"][i], ref["z"][i]) < 1e-6
    G.close()

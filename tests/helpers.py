"""Shared helpers for the parity tests."""
import numpy as np

from loik_b200 import problems

CTOR_KEYS = ("max_iter", "tol_abs", "tol_rel", "tol_primal_inf", "tol_dual_inf", "rho", "mu", "mu_equality_scale_factor",
             "mu_update_strat", "num_eq_c", "eq_c_dim", "warm_start", "tol_tail_solve")
PROB_KEYS = ("q", "H_ref", "v_ref", "ids", "Ais", "bis", "lb", "ub")


def ctor_kwargs(params):
    return {k: params[k] for k in CTOR_KEYS}


def prob_args(pr):
    return [pr[k] for k in PROB_KEYS]


def instance(pb, i):
    """Instance i of a random_batch as single-instance SolveInit arguments."""
    return [pb["q"][i], pb["H_ref"], pb["v_ref"], pb["ids"], pb["Ais"], pb["bis"][i], pb["lb"], pb["ub"]]


REL_FLOOR = 1e-9


def rel_inf(a, b):
    """|a - b|_inf / |b|_inf -- the rel-inf distance north_star quotes (1e-6 gate).  Truly relative: the denominator is the
    reference's own norm; only below |b|_inf = 1e-9 (quantities that are zero up to rounding, e.g. the multiplier w of a
    bound that never became active) does it turn into an absolute test at 1e-9 * gate."""
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.abs(a - b).max() / max(REL_FLOOR, np.abs(b).max())) if a.size else 0.0


def rel_inf_rows(a, b):
    """rel_inf of every instance (row) of two batched arrays at once."""
    a, b = np.asarray(a, float).reshape(len(a), -1), np.asarray(b, float).reshape(len(b), -1)
    return np.abs(a - b).max(axis=1) / np.maximum(REL_FLOOR, np.abs(b).max(axis=1))


def check_abs_or_rel(a, b, tol=1e-10, what=""):
    """tests/loik-loid.cpp:39-83: |a-b| < tol  or  |a-b| < tol * max(|a|,|b|), element-wise on the inf norm."""
    a, b = np.asarray(a, float), np.asarray(b, float)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    if a.size == 0:
        return
    d = np.abs(a - b).max()
    scale = max(np.abs(a).max(), np.abs(b).max())
    assert d < tol or d < tol * scale, f"{what}: |a-b|inf={d:.3e}, scale={scale:.3e}"

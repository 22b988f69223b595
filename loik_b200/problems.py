"""Synthetic problem generators.

``fixture_problem`` is the reference's only test fixture (``tests/loik-loid.cpp:87-165``) on a flat
model table; ``random_batch`` is the seeded batch both the CPU baseline and the CUDA path consume.
"""
from __future__ import annotations

import numpy as np

from .robots import TASK_JOINTS, RobotModel

# tests/loik-loid.cpp:91-105
FIXTURE_PARAMS = dict(max_iter=2, tol_abs=1e-3, tol_rel=1e-3, tol_primal_inf=1e-2, tol_dual_inf=1e-2, rho=1e-5, mu=1e-2,
                      mu_equality_scale_factor=1e4, mu_update_strat=0, num_eq_c=1, eq_c_dim=6, warm_start=False,
                      tol_tail_solve=1e-1)


def fixture_problem(model: RobotModel, bound_magnitude: float = 4.0) -> dict:
    """q = neutral, H_ref = I, v_ref = 0, one task at the last joint with A = I, b = (0,0,.5,0,0,0)."""
    b = np.zeros(6)
    b[2] = 0.5
    return dict(q=model.neutral(), H_ref=np.eye(6), v_ref=np.zeros(6), ids=np.array([model.nj - 1], np.int32),
                Ais=np.eye(6)[None], bis=b[None], lb=-bound_magnitude * np.ones(model.nv),
                ub=bound_magnitude * np.ones(model.nv))


def bench_params(nc: int, max_iter: int = 200) -> dict:
    p = dict(FIXTURE_PARAMS)
    p.update(max_iter=max_iter, num_eq_c=nc)
    return p


def random_batch(model: RobotModel, batch: int, seed: int = 0, task_joints=None, first_index: int = 0,
                 b_scale: float = 0.5) -> dict:
    """Seeded batch: q ~ U(q_min, q_max), b ~ U(-b_scale, b_scale)^6 per instance per task, A = I,
    H_ref = I, v_ref = 0, ub = -lb = joint velocity limits (shared across the batch).

    Instance ``i`` of the batch is generated from ``(seed, first_index + i)`` alone, so a shard of a
    larger batch is bit-identical to the same rows generated in one piece.
    """
    if task_joints is None:
        task_joints = TASK_JOINTS.get(model.name, [model.nj - 1])
    nc = len(task_joints)
    nq = model.nq
    # counter-based generation: one Philox stream per block of 4096 instances, independent of sharding
    BLK = 4096
    q = np.empty((batch, nq))
    b = np.empty((batch, nc, 6))
    lo = first_index
    hi = first_index + batch
    blk0 = lo // BLK
    blk1 = (hi + BLK - 1) // BLK
    for blk in range(blk0, blk1):
        rng = np.random.Generator(np.random.Philox(key=seed, counter=[0, 0, 0, blk]))
        uq = rng.random((BLK, nq))
        ub_ = rng.random((BLK, nc, 6))
        s = max(lo, blk * BLK)
        e = min(hi, (blk + 1) * BLK)
        q[s - lo:e - lo] = model.q_min + uq[s - blk * BLK:e - blk * BLK] * (model.q_max - model.q_min)
        b[s - lo:e - lo] = (2.0 * ub_[s - blk * BLK:e - blk * BLK] - 1.0) * b_scale
    for a, b_ in model.quaternion_slices():  # unit quaternions (x, y, z, w) of free-flyer / spherical joints
        quat = q[:, a:b_]
        quat /= np.linalg.norm(quat, axis=1, keepdims=True)
    for a, b_ in model.unit_pair_slices():  # (cos, sin) of unbounded revolute joints / of the heading of planar joints
        cs = q[:, a:b_]
        cs /= np.linalg.norm(cs, axis=1, keepdims=True)
    return dict(q=q, H_ref=np.eye(6), v_ref=np.zeros(6), ids=np.asarray(task_joints, np.int32),
                Ais=np.tile(np.eye(6), (nc, 1, 1)), bis=b, lb=-model.v_max.copy(), ub=model.v_max.copy())

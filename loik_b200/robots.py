"""Flat kinematic-tree tables for the benchmark robots.

The reference takes a ``pinocchio::Model`` built from example-robot-data URDFs
(`tests/loik-loid.cpp:108-111,208`).  Neither Pinocchio nor any URDF exists
offline, so the robots are described here by the only things the LoIK hot path
reads from the model (`loik-loid-optimized.hxx:46-47,258-265`): ``njoints``,
``parents``, the joint type/axis, and ``jointPlacements``.  The kinematic
constants are synthetic stand-ins recalled from the public URDFs; parity is
always oracle-vs-CUDA on the *same* table.

Joint type codes (shared with ``include/loik_b200.h`` and ``oracle/loik_oracle.c``)::

    0,1,2  revolute about +x,+y,+z   (pinocchio JointModelRX/RY/RZ)
    3,4,5  prismatic along +x,+y,+z  (JointModelPX/PY/PZ)
    6      revolute, unaligned axis  (JointModelRevoluteUnaligned)
    7      prismatic, unaligned axis (JointModelPrismaticUnaligned)
    8      free-flyer (JointModelFreeFlyer, nq = 7 = x y z qx qy qz qw, nv = 6)
    9,10,11 unbounded revolute about +x,+y,+z (JointModelRUBX/RUBY/RUBZ: URDF ``continuous`` joints; nq = 2, q = (cos, sin))
    12     unbounded revolute, unaligned axis (JointModelRevoluteUnboundedUnaligned, nq = 2)
    13     spherical (JointModelSpherical, nq = 4 = unit quaternion x y z w, nv = 3, S = [0; I3])
    14     translation (JointModelTranslation, nq = nv = 3, S = [I3; 0])
    15     planar (JointModelPlanar, nq = 4 = x y cos sin, nv = 3 = vx vy wz: S selects components 0, 1, 5)
    (the multi-DoF types 8, 13, 14, 15 may sit anywhere in the tree; the CUDA kernels take up to 8 of them per model)

Other 1-DoF joints have ``nq = nv = 1``; ``idx_q`` / ``idx_v`` follow pinocchio (cumulative over the joints in id order).
"""
from __future__ import annotations

import dataclasses
import math

import numpy as np

RX, RY, RZ, PX, PY, PZ, RU, PU, FF, RUBX, RUBY, RUBZ, RUBU, SPH, TRA, PLA, ZYX = range(17)
_AXES = {"x": (1.0, 0.0, 0.0), "y": (0.0, 1.0, 0.0), "z": (0.0, 0.0, 1.0)}


def rpy_to_matrix(r: float, p: float, y: float) -> np.ndarray:
    """URDF fixed-axis roll/pitch/yaw -> rotation matrix (Rz(y) Ry(p) Rx(r))."""
    cr, sr, cp, sp, cy, sy = math.cos(r), math.sin(r), math.cos(p), math.sin(p), math.cos(y), math.sin(y)
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]], dtype=np.float64)
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]], dtype=np.float64)
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]], dtype=np.float64)
    return Rz @ Ry @ Rx


def _quat_mul(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Hamilton product of quaternions stored (x, y, z, w), broadcasting over leading axes."""
    ax, ay, az, aw = (a[..., k] for k in range(4))
    bx, by, bz, bw = (b[..., k] for k in range(4))
    return np.stack([aw * bx + ax * bw + ay * bz - az * by, aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw, aw * bw - ax * bx - ay * by - az * bz], axis=-1)


def _quat_rotate(qt: np.ndarray, p: np.ndarray) -> np.ndarray:
    """R(q) p for a quaternion (x, y, z, w): the rotation matrix of joint_transform applied to p."""
    x, y, z, w = (qt[..., k] for k in range(4))
    R = np.stack([np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], -1),
                  np.stack([2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)], -1),
                  np.stack([2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], -1)], -2)
    return np.einsum("...ij,...j->...i", R, p)


@dataclasses.dataclass
class RobotModel:
    """What the LoIK hot path needs from ``pinocchio::Model`` (joint 0 = universe)."""

    name: str
    parent: np.ndarray        # [nj] int32, parent[0] = 0
    jtype: np.ndarray         # [nj] int32 joint type code (entry 0 unused)
    axis: np.ndarray          # [nj,3] unit axis in the joint frame
    placement_R: np.ndarray   # [nj,3,3] jointPlacements[i].rotation()
    placement_p: np.ndarray   # [nj,3]   jointPlacements[i].translation()
    q_min: np.ndarray         # [nv]
    q_max: np.ndarray         # [nv]
    v_max: np.ndarray         # [nv] joint velocity limits (ub = -lb)
    joint_names: list

    @classmethod
    def from_tables(cls, name, parent, jtype, axis, placement_R, placement_p, q_min=None, q_max=None, v_max=None, joint_names=None):
        """A model from flat tables (e.g. the dump of a pinocchio::Model, oracle/ref_recipe); limits default to +-pi / 2."""
        m = cls(name, np.asarray(parent, np.int32), np.asarray(jtype, np.int32), np.asarray(axis, float), np.asarray(placement_R, float),
                np.asarray(placement_p, float), None, None, None, joint_names or [f"j{i}" for i in range(len(parent))])
        m.q_min = -np.pi * np.ones(m.nq) if q_min is None else np.asarray(q_min, float)
        m.q_max = np.pi * np.ones(m.nq) if q_max is None else np.asarray(q_max, float)
        m.v_max = 2.0 * np.ones(m.nv) if v_max is None else np.asarray(v_max, float)
        return m

    @property
    def nj(self) -> int:
        return int(self.parent.shape[0])

    @property
    def nb(self) -> int:
        return self.nj - 1

    @property
    def has_free_flyer(self) -> bool:
        """A free-flyer root joint (floating-base robots)."""
        return self.nj > 1 and int(self.jtype[1]) == FF

    @property
    def nv(self) -> int:
        return sum(self.nv_joint(i) for i in range(1, self.nj))

    @property
    def nq(self) -> int:
        return sum(self.nq_joint(i) for i in range(1, self.nj))

    def nv_joint(self, i: int) -> int:
        jt = int(self.jtype[i])
        return 6 if jt == FF else (3 if jt in (SPH, TRA, PLA, ZYX) else 1)

    def nq_joint(self, i: int) -> int:
        jt = int(self.jtype[i])
        return {FF: 7, SPH: 4, TRA: 3, PLA: 4, ZYX: 3}.get(jt, 2 if RUBX <= jt <= RUBU else 1)

    def quaternion_slices(self):
        """(start, stop) of every unit quaternion inside q (free-flyer: q[iq+3:iq+7], spherical: q[iq:iq+4])."""
        out = []
        for i in range(1, self.nj):
            jt, iq = int(self.jtype[i]), self.idx_q(i)
            if jt == FF:
                out.append((iq + 3, iq + 7))
            elif jt == SPH:
                out.append((iq, iq + 4))
        return out

    def idx_v(self, i: int) -> int:
        return sum(self.nv_joint(k) for k in range(1, i))

    def idx_q(self, i: int) -> int:
        return sum(self.nq_joint(k) for k in range(1, i))

    def is_unbounded(self, i: int) -> bool:
        return RUBX <= int(self.jtype[i]) <= RUBU

    def unit_pair_slices(self):
        """(start, stop) of every (cos, sin) pair inside q: unbounded revolute joints, the heading of planar joints."""
        out = []
        for i in range(1, self.nj):
            if self.is_unbounded(i):
                out.append((self.idx_q(i), self.idx_q(i) + 2))
            elif int(self.jtype[i]) == PLA:
                out.append((self.idx_q(i) + 2, self.idx_q(i) + 4))
        return out

    def neutral(self) -> np.ndarray:
        q = np.zeros(self.nq)
        for a, b in self.quaternion_slices():
            q[b - 1] = 1.0  # unit quaternion (x, y, z, w)
        for a, b in self.unit_pair_slices():
            q[a] = 1.0  # (cos, sin) = (1, 0)
        return q

    def normalize(self, q: np.ndarray) -> np.ndarray:
        """pinocchio::normalize: unit quaternion of a free-flyer, unit (cos, sin) of the unbounded revolute joints."""
        q = np.array(q, np.float64)
        for a, b in self.quaternion_slices():
            q[..., a:b] /= np.linalg.norm(q[..., a:b], axis=-1, keepdims=True)
        for a, b in self.unit_pair_slices():
            q[..., a:b] /= np.linalg.norm(q[..., a:b], axis=-1, keepdims=True)
        return q

    def integrate(self, q: np.ndarray, v: np.ndarray) -> np.ndarray:
        """pinocchio::integrate(model, q, v): q + v for the vector-space joints; the SO(2) update of
        JointModelRevoluteUnbounded* (rotate (cos, sin) by v, then the first-order renormalisation
        ``out *= (3 - |out|^2) / 2``); quaternion * exp3(omega) for spherical joints and M * exp6(v) for free-flyers
        (body-frame velocities, [linear; angular]), each followed by the same first-order renormalisation of the
        quaternion.  ``q`` is [..., nq], ``v`` is [..., nv]."""
        q = np.asarray(q, np.float64)
        v = np.asarray(v, np.float64)
        out = np.empty_like(q)
        for i in range(1, self.nj):
            iq, iv, jt = self.idx_q(i), self.idx_v(i), int(self.jtype[i])
            if jt in (FF, SPH):
                o = 3 if jt == FF else 0
                w = v[..., iv + o:iv + o + 3]
                quat = q[..., iq + o:iq + o + 4]
                th2 = np.sum(w * w, axis=-1)
                th = np.sqrt(th2)
                small = th < 1e-4
                ths = np.where(small, 1.0, th)
                k = np.where(small, 0.5 - th2 / 48.0, np.sin(ths / 2) / ths)            # sin(th/2)/th
                dq = np.concatenate([k[..., None] * w, np.where(small, 1.0 - th2 / 8.0, np.cos(ths / 2))[..., None]], axis=-1)
                res = _quat_mul(quat, dq)
                if jt == FF:
                    vl = v[..., iv:iv + 3]
                    a_v = np.where(small, 1.0 - th2 / 6.0, np.sin(ths) / ths)           # sin(th)/th
                    a_wxv = np.where(small, 0.5 - th2 / 24.0, (1.0 - np.cos(ths)) / np.where(small, 1.0, th2))
                    a_w = np.where(small, 1.0 / 6.0 - th2 / 120.0, (1.0 - a_v) / np.where(small, 1.0, th2)) * np.sum(w * vl, axis=-1)
                    p = a_v[..., None] * vl + a_w[..., None] * w + a_wxv[..., None] * np.cross(w, vl)
                    out[..., iq:iq + 3] = q[..., iq:iq + 3] + _quat_rotate(quat, p)
                    res = np.where((np.sum(res * quat, axis=-1) < 0.0)[..., None], -res, res)
                res = res * ((3.0 - np.sum(res * res, axis=-1)) / 2.0)[..., None]        # quaternion::firstOrderNormalize
                out[..., iq + o:iq + o + 4] = res
            elif jt in (TRA, ZYX):  # vector spaces
                out[..., iq:iq + 3] = q[..., iq:iq + 3] + v[..., iv:iv + 3]
            elif jt == PLA:  # SpecialEuclideanOperationTpl<2>::integrate_impl: (R0, t0) * exp(v)
                c0, s0 = q[..., iq + 2], q[..., iq + 3]
                vx, vy, om = v[..., iv], v[..., iv + 1], v[..., iv + 2]
                cv, sv = np.cos(om), np.sin(om)
                big = np.abs(om) > 1e-14
                oms = np.where(big, om, 1.0)
                ax, ay = -vy / oms, vx / oms
                tx = np.where(big, ax - (cv * ax - sv * ay), vx)
                ty = np.where(big, ay - (sv * ax + cv * ay), vy)
                out[..., iq] = q[..., iq] + (c0 * tx - s0 * ty)
                out[..., iq + 1] = q[..., iq + 1] + (s0 * tx + c0 * ty)
                out[..., iq + 2] = c0 * cv - s0 * sv
                out[..., iq + 3] = s0 * cv + c0 * sv
            elif self.is_unbounded(i):
                ca, sa, om = q[..., iq], q[..., iq + 1], v[..., iv]
                co, so = np.cos(om), np.sin(om)
                c, s_ = co * ca - so * sa, so * ca + co * sa
                k = (3.0 - (c * c + s_ * s_)) / 2.0
                out[..., iq], out[..., iq + 1] = c * k, s_ * k
            else:
                out[..., iq] = q[..., iq] + v[..., iv]
        return out

    def validate(self) -> None:
        assert self.parent[0] == 0
        for i in range(1, self.nj):
            assert 0 <= self.parent[i] < i, "joints must be numbered parent < child"
            assert 0 <= self.jtype[i] <= ZYX
            assert abs(np.linalg.norm(self.axis[i]) - 1.0) < 1e-12
            R = self.placement_R[i]
            assert np.allclose(R @ R.T, np.eye(3), atol=1e-12)


def _build(name, joints) -> RobotModel:
    """joints: list of (name, parent_id, type, axis, xyz, rpy, qmin, qmax, vmax), ids from 1."""
    nj = len(joints) + 1
    parent = np.zeros(nj, np.int32)
    jtype = np.zeros(nj, np.int32)
    axis = np.zeros((nj, 3))
    axis[0] = (0.0, 0.0, 1.0)
    R = np.tile(np.eye(3), (nj, 1, 1))
    p = np.zeros((nj, 3))
    qmin, qmax, vmax, names = [], [], [], ["universe"]
    for i, (jn, par, jt, ax, xyz, rpy, lo, hi, vm) in enumerate(joints, start=1):
        parent[i] = par
        if jt in ("FF", "S", "T", "PL", "ZYX"):
            jtype[i] = {"FF": FF, "S": SPH, "T": TRA, "PL": PLA, "ZYX": ZYX}[jt]
            axis[i] = (0.0, 0.0, 1.0)
            R[i] = rpy_to_matrix(*rpy)
            p[i] = xyz
            nqj, nvj = {"FF": (7, 6), "S": (4, 3), "T": (3, 3), "PL": (4, 3), "ZYX": (3, 3)}[jt]  # (ZYX: three Euler angles, away from the gimbal lock at +-pi/2)
            qmin += [-1.0] * nqj   # position box; the quaternion part is normalised by the samplers
            qmax += [1.0] * nqj
            vmax += [vm] * nvj
            names.append(jn)
            continue
        if isinstance(ax, str):
            a = np.array(_AXES[ax])
            code = {"x": 0, "y": 1, "z": 2}[ax] + {"R": 0, "P": 3, "C": RUBX}[jt]
        else:
            a = np.asarray(ax, np.float64)
            a = a / np.linalg.norm(a)
            code = {"R": RU, "P": PU, "C": RUBU}[jt]
        jtype[i] = code
        axis[i] = a
        R[i] = rpy_to_matrix(*rpy)
        p[i] = xyz
        if jt == "C":  # URDF `continuous`: q = (cos, sin); the samplers draw both in [-1, 1] and normalise the pair
            qmin += [-1.0, -1.0]
            qmax += [1.0, 1.0]
        else:
            qmin.append(lo)
            qmax.append(hi)
        vmax.append(vm)
        names.append(jn)
    m = RobotModel(name, parent, jtype, axis, R, p, np.array(qmin), np.array(qmax), np.array(vmax), names)
    m.validate()
    return m


_H = math.pi / 2


def panda(fingers: bool = False) -> RobotModel:
    """Franka Panda arm, 7 revolute-z joints (BASELINE.json "Panda 7-DoF").

    ``fingers=True`` adds the two prismatic finger joints of the real URDF (nv = 9,
    `tests/loik-loid.cpp:214-215`): finger 1 slides along +y (PY), finger 2 along -y
    (unaligned prismatic); both hang off joint 7, which makes the tree branch.
    """
    J = [
        ("panda_joint1", 0, "R", "z", (0, 0, 0.333), (0, 0, 0), -2.8973, 2.8973, 2.175),
        ("panda_joint2", 1, "R", "z", (0, 0, 0), (-_H, 0, 0), -1.7628, 1.7628, 2.175),
        ("panda_joint3", 2, "R", "z", (0, -0.316, 0), (_H, 0, 0), -2.8973, 2.8973, 2.175),
        ("panda_joint4", 3, "R", "z", (0.0825, 0, 0), (_H, 0, 0), -3.0718, -0.0698, 2.175),
        ("panda_joint5", 4, "R", "z", (-0.0825, 0.384, 0), (-_H, 0, 0), -2.8973, 2.8973, 2.61),
        ("panda_joint6", 5, "R", "z", (0, 0, 0), (_H, 0, 0), -0.0175, 3.7525, 2.61),
        ("panda_joint7", 6, "R", "z", (0.088, 0, 0), (_H, 0, 0), -2.8973, 2.8973, 2.61),
    ]
    if fingers:
        # joint8 (0,0,0.107) * hand Rz(-pi/4) * (0,0,0.0584)
        J += [
            ("panda_finger_joint1", 7, "P", "y", (0, 0, 0.1654), (0, 0, -math.pi / 4), 0.0, 0.04, 0.2),
            ("panda_finger_joint2", 7, "P", (0.0, -1.0, 0.0), (0, 0, 0.1654), (0, 0, -math.pi / 4), 0.0, 0.04, 0.2),
        ]
    return _build("panda9" if fingers else "panda", J)


def ur10(continuous: bool = False) -> RobotModel:
    """UR10, 6-DoF serial chain (BASELINE.json "UR10 6-DoF").  ``continuous=True``: shoulder pan and wrist 3 as URDF
    ``continuous`` joints (pinocchio JointModelRUBZ / RUBY, nq = 8), the way the e-series wrist 3 is modelled."""
    tp = 2 * math.pi
    J = [
        ("shoulder_pan_joint", 0, "R", "z", (0, 0, 0.1273), (0, 0, 0), -tp, tp, 2.16),
        ("shoulder_lift_joint", 1, "R", "y", (0, 0.220941, 0), (0, _H, 0), -tp, tp, 2.16),
        ("elbow_joint", 2, "R", "y", (0, -0.1719, 0.612), (0, 0, 0), -tp, tp, 3.15),
        ("wrist_1_joint", 3, "R", "y", (0, 0, 0.5723), (0, _H, 0), -tp, tp, 3.2),
        ("wrist_2_joint", 4, "R", "z", (0, 0.1149, 0), (0, 0, 0), -tp, tp, 3.2),
        ("wrist_3_joint", 5, "R", "y", (0, 0, 0.1157), (0, 0, 0), -tp, tp, 3.2),
    ]
    if continuous:
        J[0] = J[0][:2] + ("C",) + J[0][3:]
        J[5] = J[5][:2] + ("C",) + J[5][3:]
    return _build("ur10c" if continuous else "ur10", J)


def talos(floating: bool = False) -> RobotModel:
    """Talos humanoid, 32 revolute joints, five branches; fixed base as in the reference fixture
    (`tests/loik-loid.cpp:110-111`), or ``floating=True``: a free-flyer root joint (nv = 38).

    Topology is the structural part (the base has three children, torso_2 three);
    link offsets are plausible synthetic values.
    """
    J = []
    base = 0
    if floating:
        J.append(("root_joint", 0, "FF", None, (0, 0, 0), (0, 0, 0), None, None, 2.0))
        base = 1

    def leg(side, sgn, root_parent):
        base = len(J)
        ax = ["z", "x", "y", "y", "y", "x"]
        xyz = [(-0.02, sgn * 0.085, -0.27105), (0, 0, 0), (0, 0, 0), (0, 0, -0.38), (0, 0, -0.325), (0, 0, 0)]
        lim = [(-0.35, 1.57), (-0.52, 0.52), (-2.10, 0.70), (0.0, 2.62), (-1.27, 0.68), (-0.52, 0.52)]
        vm = [3.87, 5.8, 5.8, 7.0, 5.8, 4.8]
        for k in range(6):
            par = root_parent if k == 0 else base + k
            J.append((f"leg_{side}_{k+1}_joint", par, "R", ax[k], xyz[k], (0, 0, 0), lim[k][0], lim[k][1], vm[k]))

    leg("left", +1.0, base)       # joints 1..6   (+1 each with a floating base)
    leg("right", -1.0, base)      # joints 7..12
    J.append(("torso_1_joint", base, "R", "z", (0, 0, 0.0722), (0, 0, 0), -1.26, 1.26, 5.4))          # 13
    J.append(("torso_2_joint", 13 + base, "R", "y", (0, 0, 0), (0, 0, 0), -0.23, 0.73, 5.4))          # 14

    def arm(side, sgn):
        base = len(J)
        ax = ["z", "x", "z", "y", "z", "x", "y", "y"]
        xyz = [(0, sgn * 0.1575, 0.232), (0.00493, sgn * 0.1365, 0.04673), (0, 0, 0), (0.02, 0, -0.273),
               (-0.02, 0, -0.2643), (0, 0, 0), (0, 0, 0), (0, 0, -0.09)]
        lim = [(-1.57, 0.52), (0.01, 2.86), (-2.43, 2.43), (-2.23, 0.0), (-2.51, 2.51), (-1.37, 1.37), (-0.68, 0.68), (-0.96, 0.0)]
        vm = [2.7, 3.66, 4.58, 4.58, 1.95, 1.76, 1.76, 1.0]
        names = [f"arm_{side}_{k+1}_joint" for k in range(7)] + [f"gripper_{side}_joint"]
        for k in range(8):
            par = torso2 if k == 0 else base + k
            J.append((names[k], par, "R", ax[k], xyz[k], (0, 0, 0), lim[k][0], lim[k][1], vm[k]))

    torso2 = 14 + base
    arm("left", +1.0)             # joints 15..22
    arm("right", -1.0)            # joints 23..30
    J.append(("head_1_joint", torso2, "R", "y", (0, 0, 0.316), (0, 0, 0), -0.21, 0.79, 3.0))          # 31
    J.append(("head_2_joint", 31 + base, "R", "z", (0.039, 0, 0), (0, 0, 0), -1.31, 1.31, 3.0))       # 32
    m = _build("talos_ff" if floating else "talos", J)
    assert m.nj == 33 + base
    return m


def random_tree(nb: int, seed: int, branching: float = 0.3, unaligned: float = 0.3, prismatic: float = 0.25,
                continuous: float = 0.0, multidof: float = 0.0, max_multidof: int = 8, zyx: float = 0.0) -> RobotModel:
    """Seeded random kinematic tree covering every joint type (parity stress tests)."""
    rng = np.random.default_rng(seed)
    J = []
    n_md = 0  # (the CUDA kernels take up to kMaxMd = 16; the default of 8 keeps existing seeds unchanged multi-DoF joints per model)
    for i in range(1, nb + 1):
        par = i - 1 if (i == 1 or rng.random() > branching) else int(rng.integers(0, i))
        kind = "P" if rng.random() < prismatic else "R"
        if continuous > 0.0 and kind == "R" and rng.random() < continuous:
            kind = "C"
        if multidof > 0.0 and rng.random() < multidof and n_md < max_multidof:  # spherical / translation / free-flyer anywhere in the tree
            kind = ("S", "T", "FF", "PL")[int(rng.integers(0, 4))]
            n_md += 1
        if zyx > 0.0 and rng.random() < zyx and n_md < max_multidof and kind not in ("S", "T", "FF", "PL"):  # JointModelSphericalZYX: S depends on q
            kind = "ZYX"
            n_md += 1
        if rng.random() < unaligned:
            ax = rng.normal(size=3)
        else:
            ax = "xyz"[int(rng.integers(0, 3))]
        xyz = tuple(rng.uniform(-0.4, 0.4, size=3))
        rpy = tuple(rng.uniform(-math.pi, math.pi, size=3))
        lo, hi = (-0.3, 0.3) if kind == "P" else (-2.5, 2.5)
        J.append((f"j{i}", par, kind, ax, xyz, rpy, lo, hi, float(rng.uniform(1.0, 4.0))))
    return _build(f"random{nb}_s{seed}" + ("c" if continuous > 0.0 else "") + ("m" if multidof > 0.0 else "") + ("z" if zyx > 0.0 else ""), J)


ROBOTS = {"panda": panda, "panda9": lambda: panda(True), "ur10": ur10, "ur10c": lambda: ur10(True), "talos": talos,
          "talos_ff": lambda: talos(True),
          # a seeded tree with every multi-DoF joint type incl. SphericalZYX (golden fixture tests/golden/random_tree_zyx.npz)
          "tree_zyx": lambda: dataclasses.replace(random_tree(11, 205, multidof=0.4, zyx=0.5), name="tree_zyx")}

# End-effector task joints used by the BASELINE.json configs.
TASK_JOINTS = {"panda": [7], "panda9": [7], "ur10": [6], "ur10c": [6], "talos": [21, 29], "talos_ff": [22, 30], "tree_zyx": [6, 11]}


def get_robot(name: str) -> RobotModel:
    return ROBOTS[name]()

"""The kinematic conventions of the restated Pinocchio primitives (P3-P5 of SURVEY.md section 8(a)) against first
principles.

Both oracles (and the CUDA kernels) share the *restated* conventions of pinocchio 3.0.0 -- liMi = placement * M(q),
body-frame spatial velocities ordered [linear; angular], v_i = liMi.actInv(v_parent) + S nu_i -- so agreement between
them cannot expose a wrong convention.  This test can: the link velocities `vis` a solve returns must equal the body
velocities obtained by *numerically differentiating an independent forward kinematics* (plain rotation matrices,
written here, nothing imported from oracle/ or loik_b200/csrc) along the joint velocity `nu` the same solve returns.
"""
import numpy as np
import pytest

from loik_b200 import problems, robots
from oracle import recursion
from tests.helpers import ctor_kwargs


def _rot(axis, th):
    a = np.asarray(axis, float)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * (K @ K)


def _fk(model, q):
    """World placement (R, p) of every joint frame: oMi = oM_parent * jointPlacement_i * M_i(q_i)."""
    R = [np.eye(3)] * model.nj
    p = [np.zeros(3)] * model.nj
    for i in range(1, model.nj):
        jt, iq = int(model.jtype[i]), model.idx_q(i)
        ax = np.eye(3)[jt % 3] if jt <= 5 else (np.eye(3)[jt - 9] if 9 <= jt <= 11 else model.axis[i])
        if jt in (8, 13):  # free-flyer (x y z | quaternion) / spherical (quaternion x y z w)
            x, y, z, w = q[iq + 3:iq + 7] if jt == 8 else q[iq:iq + 4]
            Rj = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                           [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                           [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
            pj = q[iq:iq + 3].copy() if jt == 8 else np.zeros(3)
        elif jt == 14:
            Rj, pj = np.eye(3), q[iq:iq + 3].copy()
        elif jt == 16:  # SphericalZYX: yaw about z, then pitch about the new y, then roll about the new x
            Rj, pj = _rot([0, 0, 1.0], q[iq]) @ _rot([0, 1.0, 0], q[iq + 1]) @ _rot([1.0, 0, 0], q[iq + 2]), np.zeros(3)
        elif jt == 15:  # planar: x, y, heading (cos, sin)
            Rj, pj = _rot([0, 0, 1.0], np.arctan2(q[iq + 3], q[iq + 2])), np.array([q[iq], q[iq + 1], 0.0])
        elif jt <= 2 or jt == 6:
            Rj, pj = _rot(ax, q[iq]), np.zeros(3)
        elif 9 <= jt <= 12:
            Rj, pj = _rot(ax, np.arctan2(q[iq + 1], q[iq])), np.zeros(3)
        else:
            Rj, pj = np.eye(3), ax * q[iq]
        par = int(model.parent[i])
        Rl = model.placement_R[i] @ Rj
        pl = model.placement_p[i] + model.placement_R[i] @ pj
        R[i] = R[par] @ Rl
        p[i] = p[par] + R[par] @ pl
    return R, p


def _quat_mul(a, b):  # (x, y, z, w)
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by, aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw, aw * bw - ax * bx - ay * by - az * bz])


def _move(model, q, v):
    """q (+) v with the exact exponential maps, written here (body-frame velocities: q_R <- q_R * exp(omega),
    p <- p + R v_lin for the free-flyer)."""
    out = np.array(q, float)
    for i in range(1, model.nj):
        jt, iq, iv = int(model.jtype[i]), model.idx_q(i), model.idx_v(i)
        if jt in (8, 13):
            w = v[iv + 3:iv + 6] if jt == 8 else v[iv:iv + 3]
            th = np.linalg.norm(w)
            dq = np.append(np.sin(th / 2) * w / th, np.cos(th / 2)) if th > 0 else np.array([0, 0, 0, 1.0])
            qs = slice(iq + 3, iq + 7) if jt == 8 else slice(iq, iq + 4)
            if jt == 8:  # first order in |v| is all the finite difference needs
                x, y, z, ww = q[qs]
                R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * ww), 2 * (x * z + y * ww)],
                              [2 * (x * y + z * ww), 1 - 2 * (x * x + z * z), 2 * (y * z - x * ww)],
                              [2 * (x * z - y * ww), 2 * (y * z + x * ww), 1 - 2 * (x * x + y * y)]])
                out[iq:iq + 3] = q[iq:iq + 3] + R @ v[iv:iv + 3]
            out[qs] = _quat_mul(q[qs], dq)
        elif jt in (14, 16):  # vector spaces: translation, Euler angles
            out[iq:iq + 3] = q[iq:iq + 3] + v[iv:iv + 3]
        elif jt == 15:  # body-frame (vx, vy, wz), first order in |v|
            th = np.arctan2(q[iq + 3], q[iq + 2])
            out[iq] = q[iq] + np.cos(th) * v[iv] - np.sin(th) * v[iv + 1]
            out[iq + 1] = q[iq + 1] + np.sin(th) * v[iv] + np.cos(th) * v[iv + 1]
            out[iq + 2], out[iq + 3] = np.cos(th + v[iv + 2]), np.sin(th + v[iv + 2])
        elif 9 <= jt <= 12:
            th = np.arctan2(q[iq + 1], q[iq]) + v[iv]
            out[iq], out[iq + 1] = np.cos(th), np.sin(th)
        else:
            out[iq] = q[iq] + v[iv]
    return out


def _body_velocities(model, q, nu, eps=1e-6):
    """Central differences of the forward kinematics along nu, expressed in each joint's own frame, [linear; angular]."""
    Rp, pp = _fk(model, _move(model, q, eps * nu))
    Rm, pm = _fk(model, _move(model, q, -eps * nu))
    R0, _ = _fk(model, q)
    out = np.zeros((model.nj, 6))
    for i in range(1, model.nj):
        W = R0[i].T @ (Rp[i] - Rm[i]) / (2 * eps)          # R^T Rdot = [omega]x in the body frame
        out[i, :3] = R0[i].T @ (pp[i] - pm[i]) / (2 * eps)
        out[i, 3:] = [W[2, 1] - W[1, 2], W[0, 2] - W[2, 0], W[1, 0] - W[0, 1]]
        out[i, 3:] *= 0.5
    return out


def _check(model, pr, what):
    params = dict(problems.FIXTURE_PARAMS, max_iter=30, num_eq_c=len(pr["ids"]))
    B = recursion.FirstOrderLoikOptimized(model, **ctor_kwargs(params))
    B.Solve(pr["q"], pr["H_ref"], pr["v_ref"], pr["ids"], pr["Ais"], pr["bis"], pr["lb"], pr["ub"])
    assert np.abs(B.nu).max() > 1e-3, what + ": degenerate solve"
    fd = _body_velocities(model, np.asarray(pr["q"], float), B.nu)
    err = np.abs(fd[1:] - B.vis[1:]).max()
    assert err < 1e-6 * max(1.0, np.abs(B.vis).max()), f"{what}: link velocities differ from d/dt FK by {err:.2e}"


@pytest.mark.parametrize("name", ["panda", "panda9", "ur10", "ur10c", "talos", "talos_ff"])
def test_link_velocities_are_time_derivatives_of_forward_kinematics(name):
    model = robots.get_robot(name)
    pb = problems.random_batch(model, 3, seed=41)
    for k in range(3):
        pr = dict(pb, q=pb["q"][k], bis=pb["bis"][k])
        _check(model, pr, f"{name}[{k}]")


@pytest.mark.parametrize("seed,continuous,multidof,zyx", [(0, 0.0, 0.0, 0.0), (1, 0.0, 0.0, 0.0), (2, 0.5, 0.0, 0.0), (3, 1.0, 0.0, 0.0), (4, 0.0, 0.4, 0.0),
                                                          (5, 0.3, 0.5, 0.0), (6, 0.0, 0.0, 0.4), (7, 0.2, 0.3, 0.4), (8, 0.0, 0.0, 1.0)])
def test_link_velocities_random_trees(seed, continuous, multidof, zyx):
    """Every joint type (aligned / unaligned, revolute / prismatic / unbounded revolute / spherical / translation /
    free-flyer / SphericalZYX -- whose motion subspace depends on q -- anywhere), random placements, branching."""
    model = robots.random_tree(11, seed, continuous=continuous, multidof=multidof, zyx=zyx)
    if zyx > 0.0:
        assert (model.jtype == 16).any()
    rng = np.random.default_rng(500 + seed)
    ids = np.array(sorted(rng.choice(np.arange(1, model.nj), size=2, replace=False)), np.int32)
    pr = dict(q=model.normalize(rng.uniform(model.q_min, model.q_max)), H_ref=np.eye(6), v_ref=np.zeros(6), ids=ids,
              Ais=np.tile(np.eye(6), (2, 1, 1)), bis=rng.uniform(-0.5, 0.5, size=(2, 6)), lb=-model.v_max, ub=model.v_max)
    _check(model, pr, f"tree{seed}")


def test_integrate_is_the_group_exponential():
    """RobotModel.integrate / the oracle's tracking driver for free-flyer and spherical joints against the matrix
    exponential (scipy.linalg.expm of the 4x4 twist / 3x3 skew matrix): M1 = M0 expm(v^), R1 = R0 expm(omega^)."""
    import scipy.linalg as sl

    def skew(a):
        return np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])

    def rq(q):
        x, y, z, w = q
        return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                         [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                         [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])

    J = [("ff", 0, "FF", None, (0, 0, 0), (0, 0, 0), None, None, 1.0), ("s", 1, "S", None, (0.1, 0, 0), (0, 0, 0), None, None, 1.0),
         ("r", 2, "R", "z", (0, 0, 0.2), (0, 0, 0), -2, 2, 1.0), ("pl", 3, "PL", None, (0, 0.1, 0), (0, 0, 0), None, None, 1.0)]
    m = robots._build("ff_sph", J)
    rng = np.random.default_rng(0)
    for scale in (1e-7, 1e-3, 0.3, 2.0):
        for _ in range(20):
            q = m.normalize(rng.normal(size=m.nq))
            v = scale * rng.normal(size=m.nv)
            q1 = m.integrate(q, v)
            M0 = np.eye(4); M0[:3, :3] = rq(q[3:7]); M0[:3, 3] = q[:3]
            X = np.zeros((4, 4)); X[:3, :3] = skew(v[3:6]); X[:3, 3] = v[:3]
            M1 = M0 @ sl.expm(X)
            assert np.abs(rq(q1[3:7]) - M1[:3, :3]).max() < 1e-12 and np.abs(q1[:3] - M1[:3, 3]).max() < 1e-12
            assert np.abs(rq(q1[7:11]) - rq(q[7:11]) @ sl.expm(skew(v[6:9]))).max() < 1e-12
            assert abs(q1[11] - (q[11] + v[9])) < 1e-15
            # planar joint: SE(2) as a 3x3 homogeneous matrix
            P0 = np.array([[q[14], -q[15], q[12]], [q[15], q[14], q[13]], [0, 0, 1.0]])
            Xp = np.array([[0, -v[12], v[10]], [v[12], 0, v[11]], [0, 0, 0.0]])
            P1 = P0 @ sl.expm(Xp)
            assert np.abs(np.array([[q1[14], -q1[15], q1[12]], [q1[15], q1[14], q1[13]], [0, 0, 1.0]]) - P1).max() < 1e-12
    # the C tracking driver integrates the same way (one step, dt = 0.05): compare the configurations it ends with
    B = 16
    pb = dict(q=m.normalize(rng.normal(size=(B, m.nq))), H_ref=np.eye(6), v_ref=np.zeros(6), ids=np.array([3], np.int32),
              Ais=np.eye(6)[None], bis=rng.uniform(-0.3, 0.3, size=(B, 1, 6)), lb=-m.v_max, ub=m.v_max)
    pb["ids"] = np.array([4], np.int32)
    params = dict(problems.bench_params(1), warm_start=True)
    trk = recursion.batch_track(m, params, pb["q"], pb["H_ref"], pb["v_ref"], pb["ids"], pb["Ais"], pb["bis"], pb["bis"], pb["lb"],
                                pb["ub"], c_id=4, dt=0.05, steps=1, warm=True)
    full = recursion.batch_solve(m, params, pb["q"], pb["H_ref"], pb["v_ref"], pb["ids"], pb["Ais"], pb["bis"], pb["lb"], pb["ub"])
    np.testing.assert_allclose(trk["q"], m.integrate(pb["q"], 0.05 * full["z"]), rtol=0, atol=1e-14)

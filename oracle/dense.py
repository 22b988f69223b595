"""Oracle A: dense / explicit numpy restatement of the reference's *non-optimized* solver.

TEST INFRASTRUCTURE ONLY.  Restates ``FirstOrderLoikTpl`` (``loik-loid.hxx``, ``loik-loid.hpp``) and
``IkProblemStandardQPFormulation`` (``ik-id-description.hpp:341-565``) with explicit 6x6 action
matrices, explicit joint subspace ``S``, explicit ``D = R + S^T H S`` / ``P = I - H S D^-1 S^T`` and the
dense OSQP-form QP ``(P_qp, q_qp, A_qp, lb_qp, ub_qp)`` for the residuals.  It shares no code with
oracle B (``loik_oracle.c``): agreement of the two to 1e-10, step by step and end to end, is what pins
the recursion oracle -- the same differential structure the reference's own tests use
(``tests/loik-loid.cpp:305-984``).  PARITY UNPINNED by known answers: the reference has none.

Paths cited are relative to /root/reference/include/loik/.
"""
from __future__ import annotations

import numpy as np


def skew(t):
    return np.array([[0.0, -t[2], t[1]], [t[2], 0.0, -t[0]], [-t[1], t[0], 0.0]])


def action_matrix(R, t):
    """SE3::toActionMatrix():  [[R, t^ R], [0, R]]  acting on Motion [lin; ang]."""
    X = np.zeros((6, 6))
    X[:3, :3] = R
    X[:3, 3:] = skew(t) @ R
    X[3:, 3:] = R
    return X


def dual_action_matrix(R, t):
    """SE3::toDualActionMatrix():  [[R, 0], [t^ R, R]]  acting on Force [lin; ang]."""
    X = np.zeros((6, 6))
    X[:3, :3] = R
    X[3:, :3] = skew(t) @ R
    X[3:, 3:] = R
    return X


def joint_subspace(jtype, axis, q=None):
    """jdata.S(); `q` (the joint's configuration) is only needed where the subspace depends on it (SphericalZYX)."""
    if jtype == 16:  # JointModelSphericalZYX (pinocchio joint-spherical-ZYX.hpp): w_body = E(q) qdot for R = Rz(q0) Ry(q1) Rx(q2)
        if q is None:
            return np.zeros((6, 3))
        c1, s1, c2, s2 = np.cos(q[1]), np.sin(q[1]), np.cos(q[2]), np.sin(q[2])
        S = np.zeros((6, 3))
        S[3:] = [[-s1, 0.0, 1.0], [c1 * s2, c2, 0.0], [c1 * c2, -s2, 0.0]]
        return S
    if jtype == 8:  # free-flyer
        return np.eye(6)
    if jtype == 13:  # spherical: S = [0; I3]
        return np.vstack([np.zeros((3, 3)), np.eye(3)])
    if jtype == 14:  # translation: S = [I3; 0]
        return np.vstack([np.eye(3), np.zeros((3, 3))])
    if jtype == 15:  # planar: (vx, vy, wz)
        S = np.zeros((6, 3)); S[0, 0] = S[1, 1] = S[5, 2] = 1.0
        return S
    if 9 <= jtype <= 12:  # unbounded revolute: the subspace of RX/RY/RZ/RU
        jtype = jtype - 9 if jtype < 12 else 6
    S = np.zeros((6, 1))
    if jtype <= 2:
        S[3 + jtype, 0] = 1.0
    elif jtype <= 5:
        S[jtype - 3, 0] = 1.0
    elif jtype == 6:
        S[3:, 0] = axis
    else:
        S[:3, 0] = axis
    return S


def joint_transform(jtype, axis, q):
    """jmodel.calc -> jdata.M(): (R, p).  q: scalar for 1-DoF joints, (x, y, z, qx, qy, qz, qw) for the free-flyer."""
    if jtype == 14:
        return np.eye(3), np.asarray(q[:3], float).copy()
    if jtype == 16:  # SphericalZYX: three elementary rotations multiplied out (independent of the closed form in oracle B)
        def rot(k, a):
            c, s_ = np.cos(a), np.sin(a)
            i, j = (k + 1) % 3, (k + 2) % 3
            M = np.eye(3)
            M[i, i] = c; M[i, j] = -s_; M[j, i] = s_; M[j, j] = c
            return M
        return rot(2, float(q[0])) @ rot(1, float(q[1])) @ rot(0, float(q[2])), np.zeros(3)
    if jtype == 15:  # planar: q = (x, y, cos, sin)
        c, s_ = float(q[2]), float(q[3])
        return np.array([[c, -s_, 0.0], [s_, c, 0.0], [0.0, 0.0, 1.0]]), np.array([float(q[0]), float(q[1]), 0.0])
    if jtype in (8, 13):
        x, y, z, w = q[3:7] if jtype == 8 else q[0:4]
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                      [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                      [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        return R, (np.asarray(q[:3], float).copy() if jtype == 8 else np.zeros(3))
    if 9 <= jtype <= 12:  # JointModelRevoluteUnbounded*: q = (cos, sin), used as given
        c, s_ = float(q[0]), float(q[1])
        a = np.eye(3)[jtype - 9] if jtype < 12 else np.asarray(axis, float)
        K = skew(a)
        return np.eye(3) + s_ * K + (1.0 - c) * (K @ K), np.zeros(3)
    q = float(np.asarray(q).reshape(-1)[0])
    if jtype <= 2:
        a = np.eye(3)[jtype]
        rev = True
    elif jtype <= 5:
        a = np.eye(3)[jtype - 3]
        rev = False
    else:
        a = np.asarray(axis, float)
        rev = jtype == 6
    if rev:
        K = skew(a)
        R = np.eye(3) + np.sin(q) * K + (1.0 - np.cos(q)) * (K @ K)
        return R, np.zeros(3)
    return np.eye(3), a * q


class FirstOrderLoik:
    """Dense ground truth.  Ctor mirrors ``loik-loid.hpp:123-147``."""

    def __init__(self, model, max_iter, tol_abs, tol_rel, tol_primal_inf, tol_dual_inf, rho, mu,
                 mu_equality_scale_factor, mu_update_strat=0, num_eq_c=1, eq_c_dim=6, warm_start=False,
                 tol_tail_solve=1e-1):
        if eq_c_dim != 6:
            raise RuntimeError("equality constraint dimension is not 6")
        self.model = model
        self.nj, self.nb, self.nv = model.nj, model.nb, model.nv
        self.max_iter, self.tol_abs, self.tol_rel = max_iter, tol_abs, tol_rel
        self.tol_primal_inf, self.tol_dual_inf = tol_primal_inf, tol_dual_inf
        self.rho, self.mu0, self.mu = rho, mu, mu
        self.mu_equality_scale_factor = mu_equality_scale_factor
        self.mu_update_strat = mu_update_strat
        self.nc, self.m = num_eq_c, eq_c_dim
        self.warm_start, self.tol_tail_solve = warm_start, tol_tail_solve
        nj, nv = self.nj, self.nv
        # IkIdDataTpl (loik-loid-data.hpp): yis indexed by *joint id*
        self.nu = np.zeros(nv); self.nu_prev = np.zeros(nv)
        self.w = np.zeros(nv); self.z = np.zeros(nv); self.z_prev = np.zeros(nv)
        self.vis = np.zeros((nj, 6)); self.vis_prev = np.zeros((nj, 6))
        self.fis = np.zeros((nj, 6)); self.yis = np.zeros((nj, 6))
        self.liMi = [(np.eye(3), np.zeros(3)) for _ in range(nj)]
        self.oMi = [(np.eye(3), np.zeros(3)) for _ in range(nj)]
        self.S = [joint_subspace(int(model.jtype[i]), model.axis[i]) for i in range(nj)]
        self.ResetSolver()

    # ---- loik-loid.hpp:153-183 + loik-loid-data.hxx:99-160 -------------------------------------
    def ResetSolver(self):
        nj, nv, nb, m = self.nj, self.nv, self.nb, self.m
        self.iter = 0
        self.converged = self.primal_infeasible = self.dual_infeasible = False
        self.mu = self.mu0
        if not self.warm_start:
            self.nu[:] = 0; self.nu_prev[:] = 0; self.w[:] = 0; self.z[:] = 0; self.z_prev[:] = 0
            self.vis[:] = 0; self.vis_prev[:] = 0; self.fis[:] = 0; self.yis[:] = 0
        else:
            self.nu_prev[:] = self.nu; self.z_prev[:] = self.z; self.vis_prev[:] = self.vis; self.fis[:] = 0
        self.His = np.zeros((nj, 6, 6)); self.pis = np.zeros((nj, 6))
        self.Ris = [np.zeros((1, 1)) for _ in range(nj)]; self.ris = [np.zeros(1) for _ in range(nj)]
        self.Dis = [None] * nj; self.Di_invs = [None] * nj; self.Pis = [None] * nj
        self.tail_solve_iter = 0
        self.primal_residual = self.dual_residual = np.inf
        self.primal_residual_vec = np.zeros(m * nb + nv)
        self.dual_residual_vec = np.zeros(6 * nb + nv)
        self.mu_eq = self.mu_equality_scale_factor * self.mu
        self.mu_ineq = self.mu
        # IkProblemStandardQPFormulation::Reset (ik-id-description.hpp:372-403)
        self.ncon = 6 * nb + m * nb + nv
        self.nvar = 6 * nb + nv
        self.A_qp = np.zeros((self.ncon, self.nvar)); self.P_qp = np.zeros((self.nvar, self.nvar))
        self.q_qp = np.zeros(self.nvar); self.lb_qp = np.zeros(self.ncon); self.ub_qp = np.zeros(self.ncon)
        self.x_qp = np.zeros(self.nvar); self.z_qp = np.zeros(self.ncon); self.y_qp = np.zeros(self.ncon)
        self.delta_x_qp = np.zeros(self.nvar); self.delta_z_qp = np.zeros(self.ncon); self.delta_y_qp = np.zeros(self.ncon)
        self.delta_y_qp_plus = np.zeros(self.ncon); self.delta_y_qp_minus = np.zeros(self.ncon)
        self.hist_mu, self.hist_primal_residual, self.hist_dual_residual = [], [], []

    # ---- loik-loid.hxx:14-33 ------------------------------------------------------------------
    def FwdPassInit(self, q):
        mdl = self.model
        for i in range(1, self.nj):
            iq = mdl.idx_q(i)
            MR, Mp = joint_transform(int(mdl.jtype[i]), mdl.axis[i], q[iq:iq + mdl.nq_joint(i)])
            if int(mdl.jtype[i]) == 16:
                self.S[i] = joint_subspace(16, mdl.axis[i], q[iq:iq + 3])
            R = mdl.placement_R[i] @ MR
            p = mdl.placement_p[i] + mdl.placement_R[i] @ Mp
            self.liMi[i] = (R, p)
            oR, op = self.oMi[int(mdl.parent[i])]
            self.oMi[i] = (oR @ R, op + oR @ p)

    # ---- ik-id-description.hpp:411-491 --------------------------------------------------------
    def _update_qp_init(self, H_ref, v_ref, ids, Ais, bis, lb, ub):
        nb, nv, m = self.nb, self.nv, self.m
        self.H_refs = np.tile(np.asarray(H_ref, float).reshape(6, 6), (self.nj, 1, 1))
        self.v_refs = np.tile(np.asarray(v_ref, float).reshape(6), (self.nj, 1))
        self.task_ids = [int(c) for c in ids]
        self.Ais = np.asarray(Ais, float).reshape(-1, 6, 6).copy()
        self.bis = np.asarray(bis, float).reshape(-1, 6).copy()
        if not (len(self.task_ids) == len(self.Ais) == len(self.bis)):
            raise RuntimeError("task_constraint_ids, Ais, and bis have different size !!!")
        if len(self.task_ids) != self.nc:
            raise RuntimeError("number of equality constraints doesn't match initialization!!!")
        self.lb, self.ub = np.asarray(lb, float).copy(), np.asarray(ub, float).copy()
        off_task, off_box = 6 * nb, 6 * nb + m * nb
        self.lb_qp[off_box:] = self.lb; self.ub_qp[off_box:] = self.ub
        self.A_qp[:6 * nb, :6 * nb] = -np.eye(6 * nb)
        self.A_qp[off_box:, 6 * nb:] = np.eye(nv)
        for i in range(1, self.nj):
            r0 = (i - 1) * 6
            self.P_qp[r0:r0 + 6, r0:r0 + 6] = self.H_refs[i]
            self.q_qp[r0:r0 + 6] = -self.H_refs[i].T @ self.v_refs[i]
            iv, nvj = self.model.idx_v(i), self.model.nv_joint(i)
            self.A_qp[r0:r0 + 6, 6 * nb + iv:6 * nb + iv + nvj] = self.S[i]
            par = int(self.model.parent[i])
            if par > 0:
                # iMo * oMp as action matrices (ik-id-description.hpp:458)
                oR, op = self.oMi[i]
                iMo = action_matrix(oR.T, -oR.T @ op)
                oMp = action_matrix(*self.oMi[par])
                self.A_qp[r0:r0 + 6, (par - 1) * 6:(par - 1) * 6 + 6] = iMo @ oMp
            self.A_qp[r0:r0 + 6, r0:r0 + 6] = -np.eye(6)
        for k, c in enumerate(self.task_ids):
            r0 = off_task + (c - 1) * m
            self.A_qp[r0:r0 + m, (c - 1) * 6:(c - 1) * 6 + 6] = self.Ais[k]
            self.lb_qp[r0:r0 + m] = self.bis[k]; self.ub_qp[r0:r0 + m] = self.bis[k]
        self.z_qp[off_task:off_box] = self.ub_qp[off_task:off_box]

    # ---- ik-id-description.hpp:499-539 --------------------------------------------------------
    def UpdateQPADMMSolveLoopUtility(self):
        nb, nv, m = self.nb, self.nv, self.m
        x_prev, z_prev, y_prev = self.x_qp.copy(), self.z_qp.copy(), self.y_qp.copy()
        for i in range(1, self.nj):
            r0 = (i - 1) * 6
            self.x_qp[r0:r0 + 6] = self.vis[i]
            self.y_qp[r0:r0 + 6] = self.fis[i]
            ry = 6 * nb + (i - 1) * m
            self.y_qp[ry:ry + m] = self.yis[i]
        self.x_qp[6 * nb:] = self.nu
        self.y_qp[6 * nb + nb * m:] = self.w
        self.z_qp[6 * nb + nb * m:] = self.z
        self.delta_x_qp = self.x_qp - x_prev
        self.delta_y_qp = self.y_qp - y_prev
        self.delta_z_qp = self.z_qp - z_prev
        self.delta_y_qp_plus = np.maximum(self.delta_y_qp, 0.0)
        self.delta_y_qp_minus = np.minimum(self.delta_y_qp, 0.0)

    # ---- loik-loid-data.hxx UpdatePrev --------------------------------------------------------
    def UpdatePrev(self):
        self.vis_prev[:] = self.vis; self.nu_prev[:] = self.nu; self.z_prev[:] = self.z

    # ---- loik-loid.hxx:39-76 ------------------------------------------------------------------
    def FwdPass1(self):
        for i in range(1, self.nj):
            iv, nvj = self.model.idx_v(i), self.model.nv_joint(i)
            self.Ris[i] = self.mu_ineq * np.eye(nvj)
            self.ris[i] = self.w[iv:iv + nvj] - self.mu_ineq * self.z[iv:iv + nvj]
            self.His[i] = self.rho * np.eye(6) + self.H_refs[i]
            self.pis[i] = -self.rho * self.vis_prev[i] - self.H_refs[i].T @ self.v_refs[i]
        for k, c in enumerate(self.task_ids):
            Ai, bi = self.Ais[k], self.bis[k]
            self.His[c] = self.His[c] + self.mu_eq * Ai.T @ Ai
            self.pis[c] = self.pis[c] + Ai.T @ self.yis[c] - self.mu_eq * Ai.T @ bi

    # ---- loik-loid.hxx:82-113 -----------------------------------------------------------------
    def BwdPass(self):
        for i in range(self.nj - 1, 0, -1):
            par = int(self.model.parent[i])
            R, t = self.liMi[i]
            Hi, pi, Si = self.His[i], self.pis[i], self.S[i]
            Di = self.Ris[i] + Si.T @ Hi @ Si
            Di_inv = np.linalg.inv(Di)
            Pi = np.eye(6) - Hi @ Si @ Di_inv @ Si.T
            self.Dis[i], self.Di_invs[i], self.Pis[i] = Di, Di_inv, Pi
            Xd = dual_action_matrix(R, t)
            Xinv = np.linalg.inv(action_matrix(R, t))
            self.His[par] = self.His[par] + Xd @ (Pi @ Hi) @ Xinv
            self.pis[par] = self.pis[par] + Xd @ (Pi @ pi - Hi @ Si @ Di_inv @ self.ris[i])

    # ---- loik-loid.hxx:120-151 ----------------------------------------------------------------
    def FwdPass2(self):
        for i in range(1, self.nj):
            par = int(self.model.parent[i])
            R, t = self.liMi[i]
            vp = np.linalg.inv(action_matrix(R, t)) @ self.vis[par]
            Hi, pi, Si = self.His[i], self.pis[i], self.S[i]
            nu_i = -self.Di_invs[i] @ (Si.T @ (Hi @ vp + pi) + self.ris[i])
            iv, nvj = self.model.idx_v(i), self.model.nv_joint(i)
            self.nu[iv:iv + nvj] = nu_i
            self.vis[i] = vp + (Si @ nu_i)
            self.fis[i] = Hi @ self.vis[i] + pi

    # ---- loik-loid.hxx:158-189 ----------------------------------------------------------------
    def BoxProj(self):
        self.z = np.minimum(self.ub, np.maximum(self.lb, self.nu + (1.0 / self.mu_ineq) * self.w))

    def DualUpdate(self):
        for k, c in enumerate(self.task_ids):
            self.yis[c] = self.yis[c] + self.mu_eq * (self.Ais[k] @ self.vis[c] - self.bis[k])
        self.w = self.w + self.mu_ineq * (self.nu - self.z)

    # ---- loik-loid.hxx:207-295 ----------------------------------------------------------------
    def ComputeResiduals(self):
        nb, m = self.nb, self.m
        for k, c in enumerate(self.task_ids):
            self.primal_residual_vec[m * (c - 1):m * c] = self.Ais[k] @ self.vis[c] - self.bis[k]
        self.primal_residual_vec[m * nb:] = self.nu - self.z
        self.primal_residual = np.abs(self.primal_residual_vec).max()
        self.primal_residual_task = np.abs(self.primal_residual_vec[:m * nb]).max()
        self.primal_residual_slack = np.abs(self.primal_residual_vec[m * nb:]).max()
        # final dual residual is the dense expression (loik-loid.hxx:280)
        self.dual_residual_vec = self.P_qp @ self.x_qp + self.q_qp + self.A_qp.T @ self.y_qp
        self.dual_residual = np.abs(self.dual_residual_vec).max()
        self.dual_residual_v = np.abs(self.dual_residual_vec[:6 * nb]).max()
        self.dual_residual_nu = np.abs(self.dual_residual_vec[6 * nb:]).max()

    # ---- loik-loid.hxx:302-324 ----------------------------------------------------------------
    def CheckConvergence(self):
        self.tol_primal = self.tol_abs + self.tol_rel * max(np.abs(self.A_qp @ self.x_qp).max(), np.abs(self.z_qp).max())
        self.tol_dual = self.tol_abs + self.tol_rel * max(max(np.abs(self.P_qp @ self.x_qp).max(),
                                                              np.abs(self.A_qp.T @ self.y_qp).max()),
                                                          np.abs(self.q_qp).max())
        if self.primal_residual < self.tol_primal and self.dual_residual < self.tol_dual:
            self.converged = True

    # ---- loik-loid.hxx:331-367 (dual infeasibility included, as in the dense reference) --------
    def CheckFeasibility(self):
        dy_inf = np.abs(self.delta_y_qp).max()
        self.A_qp_T_delta_y_qp_inf_norm = np.abs(self.A_qp.T @ self.delta_y_qp).max()
        self.delta_y_qp_inf_norm = dy_inf
        self.primal_infeasibility_cond_1 = bool(self.A_qp_T_delta_y_qp_inf_norm <= self.tol_primal_inf * dy_inf)
        self.ub_qp_T_delta_y_qp_plus = float(self.ub_qp @ self.delta_y_qp_plus)
        self.lb_qp_T_delta_y_qp_minus = float(self.lb_qp @ self.delta_y_qp_minus)
        self.primal_infeasibility_cond_2 = bool(self.ub_qp_T_delta_y_qp_plus + self.lb_qp_T_delta_y_qp_minus
                                                <= self.tol_primal_inf * dy_inf)
        if self.primal_infeasibility_cond_1 and self.primal_infeasibility_cond_2:
            self.primal_infeasible = True
        dx_inf = np.abs(self.delta_x_qp).max()
        c1 = np.abs(self.P_qp @ self.delta_x_qp).max() <= self.tol_dual_inf * dx_inf
        c2 = float(self.q_qp @ self.delta_x_qp) <= self.tol_dual_inf * dx_inf
        if c1 and c2:
            Adx = self.A_qp @ self.delta_x_qp
            if (Adx >= -self.tol_dual_inf * dx_inf).all() and (Adx <= self.tol_dual_inf * dx_inf).all():
                self.dual_infeasible = True

    # ---- loik-loid.hxx:374-402 ----------------------------------------------------------------
    def UpdateMu(self):
        if self.mu_update_strat != 0:
            raise RuntimeError("[FirstOrderLoik::UpdateMu]: mu update strategy not yet implemented")
        if self.primal_residual > 10 * self.dual_residual:
            self.mu *= 10
        elif self.dual_residual > 10 * self.primal_residual:
            self.mu *= 0.1
        else:
            return
        self.mu_eq = self.mu_equality_scale_factor * self.mu
        self.mu_ineq = self.mu

    def _one_iteration(self):
        self.UpdatePrev()
        self.FwdPass1(); self.BwdPass(); self.FwdPass2(); self.BoxProj(); self.DualUpdate()
        self.UpdateQPADMMSolveLoopUtility()
        self.ComputeResiduals()
        self.hist_mu.append(self.mu); self.hist_primal_residual.append(self.primal_residual)
        self.hist_dual_residual.append(self.dual_residual)

    # ---- loik-loid.hpp:253-350 ----------------------------------------------------------------
    def InfeasibilityTailSolve(self):
        self.tail_solve_iter = 0
        while (np.abs(self.delta_x_qp).max() >= self.tol_tail_solve
               or np.abs(self.delta_z_qp).max() >= self.tol_tail_solve):
            if self.iter >= self.max_iter:
                return
            self.iter += 1
            self.tail_solve_iter += 1
            self._one_iteration()

    # ---- loik-loid.hpp:366-381 ----------------------------------------------------------------
    def SolveInit(self, q, H_ref, v_ref, ids, Ais, bis, lb, ub):
        self.ResetSolver()
        self.FwdPassInit(np.asarray(q, float))
        self._update_qp_init(H_ref, v_ref, ids, Ais, bis, lb, ub)

    # ---- loik-loid.hpp:387-455 / :470-560 -----------------------------------------------------
    def Solve(self, *args):
        if len(args) == 8:
            self.SolveInit(*args)
        elif len(args) != 0:
            raise TypeError("Solve() takes 0 or 8 arguments")
        for i in range(1, self.max_iter):
            self.iter = i
            self._one_iteration()
            self.CheckConvergence()
            if self.iter > 1:
                self.CheckFeasibility()
            if self.converged:
                break
            elif self.primal_infeasible or self.dual_infeasible:
                self.InfeasibilityTailSolve()
                break
            self.UpdateMu()


def kkt_report(model, liMi, task_ids, Ais, bis, H_ref, v_ref, lb, ub, v, nu, z, f, y, w, q=None):
    """KKT residuals of the QP of SURVEY.md section 0 at a primal-dual point (independent of ADMM):
    kinematics, task, stationarity in v and nu, box feasibility, complementarity of w."""
    nj = model.nj
    H_ref = np.asarray(H_ref, float).reshape(6, 6)
    v_ref = np.asarray(v_ref, float).reshape(6)
    out = {}
    kin = 0.0
    stat_v = np.zeros((nj, 6))
    for i in range(1, nj):
        par = int(model.parent[i])
        R, t = liMi[i]
        S = joint_subspace(int(model.jtype[i]), model.axis[i], None if q is None else q[model.idx_q(i):model.idx_q(i) + model.nq_joint(i)])
        iv, nvj = model.idx_v(i), model.nv_joint(i)
        vp = np.linalg.inv(action_matrix(R, t)) @ (v[par] if par > 0 else np.zeros(6))
        kin = max(kin, np.abs(-v[i] + vp + S @ nu[iv:iv + nvj]).max())
        stat_v[i] += H_ref @ (v[i] - v_ref) - f[i]
        if par > 0:
            stat_v[par] += dual_action_matrix(R, t) @ f[i]
    task = 0.0
    for k, c in enumerate(task_ids):
        A = np.asarray(Ais[k]).reshape(6, 6)
        stat_v[c] += A.T @ y[k]
        task = max(task, np.abs(A @ v[c] - bis[k]).max())
    stat_nu = np.concatenate([joint_subspace(int(model.jtype[i]), model.axis[i], None if q is None else q[model.idx_q(i):model.idx_q(i) + model.nq_joint(i)]).T @ f[i]
                              + w[model.idx_v(i):model.idx_v(i) + model.nv_joint(i)] for i in range(1, nj)])
    out["kinematics"] = kin
    out["task"] = task
    out["stationarity_v"] = np.abs(stat_v[1:]).max()
    out["stationarity_nu"] = np.abs(stat_nu).max()
    out["slack"] = np.abs(nu - z).max()
    out["box"] = max(0.0, (lb - z).max(), (z - ub).max())
    # w in the normal cone of the box at z: w<=0 at lb... sign: w_i > 0 only if z_i == ub_i, < 0 only if z_i == lb_i
    comp = 0.0
    for k in range(len(z)):
        if w[k] > 0:
            comp = max(comp, min(w[k], ub[k] - z[k]))
        elif w[k] < 0:
            comp = max(comp, min(-w[k], z[k] - lb[k]))
    out["complementarity"] = comp
    return out

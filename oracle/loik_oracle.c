/*
 * oracle/loik_oracle.c -- CPU restatement ("oracle B") of LoIK's FirstOrderLoikOptimizedTpl.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (loik_b200/) may call, link or
 * load this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, and only as the checker / the CPU baseline.
 *
 * PARITY UNPINNED: the reference (header-only C++ on Pinocchio 3.0.0 / Eigen 3.4 / Boost 1.84,
 * pins in /root/reference/pixi.lock:167,104,94) cannot be compiled offline and its tests hold no
 * golden vectors.  This file is pinned (1) against the dense/explicit
 * restatement oracle/loik_dense.py of the reference's *other* solver, exactly the way the
 * reference's own tests pin the optimized path (tests/loik-loid.cpp:305-984), and (2) by KKT
 * checks of converged solutions (tests/test_oracle_kkt.py).
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference/include/loik/).  Plain C99, fp64, one solver object = one problem
 * instance on one thread, as in the reference.  Pinocchio primitives the hot path calls
 * (calc_aba, SE3actOn, SE3::act / actInv, jmodel.calc) are restated from Pinocchio 3.0.0's
 * published algorithms for 1-DoF joints.
 *
 * Layout conventions: spatial vectors are [linear(0:3); angular(3:6)] (pinocchio Motion/Force),
 * 6x6 matrices are row-major double[36], rotations row-major double[9].
 */
#define _POSIX_C_SOURCE 200809L /* clock_gettime (lo_time_solve) */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define LO_API __attribute__((visibility("default")))

enum { JT_RX = 0, JT_RY, JT_RZ, JT_PX, JT_PY, JT_PZ, JT_RU, JT_PU, JT_FF, JT_RUBX, JT_RUBY, JT_RUBZ, JT_RUBU, JT_SPH, JT_TRA, JT_PLA, JT_ZYX };
static int jt_nv(int jt) { return jt == JT_FF ? 6 : ((jt == JT_SPH || jt == JT_TRA || jt == JT_PLA || jt == JT_ZYX) ? 3 : 1); }
static int jt_nq(int jt) { return jt == JT_FF ? 7 : ((jt == JT_SPH || jt == JT_PLA) ? 4 : ((jt == JT_TRA || jt == JT_ZYX) ? 3 : ((jt >= JT_RUBX && jt <= JT_RUBU) ? 2 : 1))); }

typedef struct lo_solver {
  /* ---- model (what the hot path reads from pinocchio::Model) ---- */
  int nj, nb, nv, nq, nc;
  int *parent, *jtype;
  int *nvj, *idxv, *idxq; /* jmodel.nv(), jmodel.idx_v(), jmodel.idx_q() */
  double *axis;      /* [nj*3] */
  double *plR, *plp; /* jointPlacements */
  /* ---- IkIdSolverBaseTpl (task-solver-base.hpp:146-170) ---- */
  double rho, mu0, mu, mu_equality_scale_factor;
  int mu_update_strat, max_iter, iter, converged;
  double tol_abs, tol_rel, tol_primal, tol_dual, tol_primal_inf, tol_dual_inf;
  int primal_infeasible, dual_infeasible;
  double primal_residual, dual_residual;
  /* ---- FirstOrderLoikOptimizedTpl (loik-loid-optimized.hpp:768-803) ---- */
  int tail_solve_iter, warm_start;
  double primal_residual_task, primal_residual_slack, dual_residual_v, dual_residual_nu;
  double *primal_residual_vec, *dual_residual_vec; /* [6nb+nv] */
  double delta_x_qp_inf_norm, delta_y_qp_inf_norm, A_qp_T_delta_y_qp_inf_norm;
  double ub_qp_T_delta_y_qp_plus, lb_qp_T_delta_y_qp_minus;
  int primal_infeasibility_cond_1, primal_infeasibility_cond_2;
  double mu_eq, mu_ineq, tol_tail_solve;
  /* ---- IkProblemFormulationOptimized (ik-id-description-optimized.hpp:342-362) ---- */
  double *H_refs, *v_refs, *Hv; /* [nj*36],[nj*6],[nj*6] */
  int *task_ids;                /* active_task_constraint_ids_ [nc] */
  double *Ais, *bis, *AtA, *Atb;
  double *lb, *ub;
  double bis_inf_norm, Hv_inf_norm;
  /* ---- IkIdDataTypeOptimizedTpl (loik-loid-data-optimized.hpp:109-329) ---- */
  double *oMi_R, *oMi_p, *liMi_R, *liMi_p;
  double *nu, *nu_prev, *vis, *vis_prev;
  double *His, *His_aba, *pis, *pis_aba, *R, *r;
  double *fis, *delta_fis, *yis, *delta_yis, *w, *delta_w, *z, *z_prev;
  double *Aty, *fis_diff_plus_Aty, *delta_fis_diff_plus_Aty, *Href_v, *Av_minus_b;
  double *Stf_plus_w, *delta_Stf_plus_w;
  double bT_delta_y_plus, bT_delta_y_minus, Av_inf_norm, nu_inf_norm, Href_v_inf_norm;
  double fis_diff_plus_Aty_inf_norm, Stf_plus_w_inf_norm, delta_fis_diff_plus_Aty_inf_norm;
  double delta_Stf_plus_w_inf_norm, delta_vis_inf_norm, delta_nu_inf_norm, delta_z_inf_norm;
  double delta_fis_inf_norm, delta_yis_inf_norm, delta_w_inf_norm;
  /* pinocchio JointData per joint: S (6 x nvj), U (6 x nvj), Dinv (nvj x nvj), UDinv (6 x nvj); row-major, row stride 6 */
  double *jS, *jU, *jDinv, *jUDinv;
  /* history (LoikSolverInfo, loik-loid-optimized.hpp:47-127) -- always recorded here */
  int hist_len, hist_cap;
  double *hist_mu, *hist_pres, *hist_dres;
  char err[256];
} lo_solver;

/* ------------------------------------------------------------------------------------------ */
/* small dense helpers                                                                          */
/* ------------------------------------------------------------------------------------------ */
static double *dalloc(size_t n) { return (double *)calloc(n ? n : 1, sizeof(double)); }
static double inf6(const double *v) {
  double m = 0.0;
  for (int i = 0; i < 6; ++i) { double a = fabs(v[i]); if (a > m) m = a; }
  return m;
}
static double infn(const double *v, int n) {
  double m = 0.0;
  for (int i = 0; i < n; ++i) { double a = fabs(v[i]); if (a > m) m = a; }
  return m;
}
static void mat6_vec(const double *M, const double *v, double *out) {
  for (int i = 0; i < 6; ++i) {
    double s = 0.0;
    for (int j = 0; j < 6; ++j) s += M[6 * i + j] * v[j];
    out[i] = s;
  }
}
static void mat6T_vec(const double *M, const double *v, double *out) {
  for (int i = 0; i < 6; ++i) {
    double s = 0.0;
    for (int j = 0; j < 6; ++j) s += M[6 * j + i] * v[j];
    out[i] = s;
  }
}
static void m3mul(const double *A, const double *B, double *C) { /* C = A B */
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
static void m3mulT(const double *A, const double *B, double *C) { /* C = A B^T */
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) C[3 * i + j] = A[3 * i] * B[3 * j] + A[3 * i + 1] * B[3 * j + 1] + A[3 * i + 2] * B[3 * j + 2];
}
static void cross3(const double *a, const double *b, double *c) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}
static void get_blk(const double *M, int r0, int c0, double *B) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) B[3 * i + j] = M[6 * (r0 + i) + c0 + j];
}
static void set_blk(double *M, int r0, int c0, const double *B) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) M[6 * (r0 + i) + c0 + j] = B[3 * i + j];
}

/* P3: pinocchio SE3::act(Force) -- [R f_lin ; R f_ang + p x (R f_lin)]  (call sites hxx:74,212) */
static void se3_act_force(const double *R, const double *p, const double *f, double *out) {
  double l[3], a[3], c[3];
  for (int i = 0; i < 3; ++i) {
    l[i] = R[3 * i] * f[0] + R[3 * i + 1] * f[1] + R[3 * i + 2] * f[2];
    a[i] = R[3 * i] * f[3] + R[3 * i + 1] * f[4] + R[3 * i + 2] * f[5];
  }
  cross3(p, l, c);
  for (int i = 0; i < 3; ++i) { out[i] = l[i]; out[3 + i] = a[i] + c[i]; }
}
/* P3: pinocchio SE3::actInv(Motion) -- [R^T (v_lin - p x v_ang) ; R^T v_ang]  (call site hxx:125) */
static void se3_actinv_motion(const double *R, const double *p, const double *v, double *out) {
  double c[3], t[3];
  cross3(p, v + 3, c);
  for (int i = 0; i < 3; ++i) t[i] = v[i] - c[i];
  for (int i = 0; i < 3; ++i) {
    out[i] = R[i] * t[0] + R[3 + i] * t[1] + R[6 + i] * t[2];
    out[3 + i] = R[i] * v[3] + R[3 + i] * v[4] + R[6 + i] * v[5];
  }
}
/* P2: pinocchio::impl::internal::SE3actOn<Scalar>::run(M, I) (call site hxx:66):  X* I X*^T with
 * X* = [[R,0],[p^ R,R]].  Like pinocchio it reads only the LL, LA and AA 3x3 blocks of I, i.e. it
 * assumes I symmetric. */
static void se3_act_on(const double *R, const double *t, const double *I, double *res) {
  double Ai[9], Bi[9], Di[9], tmp[9], Ao[9], Bo[9], Co[9], Do[9], col[3], cr[3];
  get_blk(I, 0, 0, Ai); get_blk(I, 0, 3, Bi); get_blk(I, 3, 3, Di);
  m3mul(R, Ai, tmp); m3mulT(tmp, R, Ao);
  m3mul(R, Bi, tmp); m3mulT(tmp, R, Bo);
  m3mul(R, Di, tmp); m3mulT(tmp, R, Do);
  for (int k = 0; k < 3; ++k) { /* Do.row(k) += t x Bo.col(k) */
    col[0] = Bo[k]; col[1] = Bo[3 + k]; col[2] = Bo[6 + k];
    cross3(t, col, cr);
    for (int j = 0; j < 3; ++j) Do[3 * k + j] += cr[j];
  }
  for (int k = 0; k < 3; ++k) { /* Co.col(k) = t x Ao.col(k) */
    col[0] = Ao[k]; col[1] = Ao[3 + k]; col[2] = Ao[6 + k];
    cross3(t, col, cr);
    for (int j = 0; j < 3; ++j) Co[3 * j + k] = cr[j];
  }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) Co[3 * i + j] += Bo[3 * j + i]; /* Co += Bo^T */
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) Bo[3 * i + j] = Co[3 * j + i]; /* Bo = Co^T */
  for (int k = 0; k < 3; ++k) { /* Do.col(k) += t x Bo.col(k) */
    col[0] = Bo[k]; col[1] = Bo[3 + k]; col[2] = Bo[6 + k];
    cross3(t, col, cr);
    for (int j = 0; j < 3; ++j) Do[3 * j + k] += cr[j];
  }
  set_blk(res, 0, 0, Ao); set_blk(res, 0, 3, Bo); set_blk(res, 3, 0, Co); set_blk(res, 3, 3, Do);
}

/* P4: jdata.S() for 1-DoF joints: S = [0;axis] (revolute) or [axis;0] (prismatic) */
static void joint_S(int jt, const double *axis, double *S) { /* S[6*row + col], col < nvj */
  memset(S, 0, 36 * sizeof(double));
  switch (jt) {
    case JT_RX: case JT_RY: case JT_RZ: S[6 * (3 + jt)] = 1.0; break;
    case JT_RUBX: case JT_RUBY: case JT_RUBZ: S[6 * (3 + jt - JT_RUBX)] = 1.0; break; /* same subspace as RX/RY/RZ */
    case JT_PX: case JT_PY: case JT_PZ: S[6 * (jt - 3)] = 1.0; break;
    case JT_RU: case JT_RUBU: S[18] = axis[0]; S[24] = axis[1]; S[30] = axis[2]; break;
    case JT_PU: S[0] = axis[0]; S[6] = axis[1]; S[12] = axis[2]; break;
    case JT_SPH: for (int k = 0; k < 3; ++k) S[6 * (3 + k) + k] = 1.0; break; /* JointModelSpherical: S = [0; I3] */
    case JT_TRA: for (int k = 0; k < 3; ++k) S[6 * k + k] = 1.0; break;       /* JointModelTranslation: S = [I3; 0] */
    case JT_PLA: S[6 * 0 + 0] = 1.0; S[6 * 1 + 1] = 1.0; S[6 * 5 + 2] = 1.0; break; /* JointModelPlanar: (vx, vy, wz) */
    case JT_ZYX: break; /* JointModelSphericalZYX: S depends on q, filled by joint_S_of_q (FwdPassInit) */
    default: for (int k = 0; k < 6; ++k) S[7 * k] = 1.0; break; /* free-flyer: identity */
  }
}

/* JointModelSphericalZYX::calc (pinocchio joint-spherical-ZYX.hpp): the motion subspace depends on the configuration,
 * S = [0; E(q)] with the body angular velocity w = E(q) qdot for R = Rz(q0) Ry(q1) Rx(q2). */
static void joint_S_of_q(int jt, const double *qv, double *S) {
  if (jt != JT_ZYX) return;
  const double c1 = cos(qv[1]), s1 = sin(qv[1]), c2 = cos(qv[2]), s2 = sin(qv[2]);
  memset(S, 0, 36 * sizeof(double));
  S[6 * 3 + 0] = -s1;     S[6 * 3 + 1] = 0.0; S[6 * 3 + 2] = 1.0;
  S[6 * 4 + 0] = c1 * s2; S[6 * 4 + 1] = c2;  S[6 * 4 + 2] = 0.0;
  S[6 * 5 + 0] = c1 * c2; S[6 * 5 + 1] = -s2; S[6 * 5 + 2] = 0.0;
}

/* P5: jmodel.calc(jdata, q) -> jdata.M(): revolute M = (Rot(axis,q), 0), prismatic M = (I, axis q).
 * Axis-aligned revolute joints fill exact 0/1 entries (pinocchio TransformRevoluteTpl); the unaligned
 * one uses pinocchio's toRotationMatrix(axis, cos, sin): R = c I + s [a]x + (1-c) a a^T. */
static void joint_M(int jt, const double *axis, const double *qv, double *MR, double *Mp) {
  const double q = qv[0];
  for (int i = 0; i < 9; ++i) MR[i] = (i % 4 == 0) ? 1.0 : 0.0;
  Mp[0] = Mp[1] = Mp[2] = 0.0;
  if (jt >= JT_RUBX && jt <= JT_RUBU) { /* JointModelRevoluteUnbounded*::calc: ca = q[0], sa = q[1], no sin/cos call */
    const double c = qv[0], s = qv[1];
    if (jt != JT_RUBU) {
      const int k = jt - JT_RUBX, a = (k + 1) % 3, b = (k + 2) % 3;
      MR[3 * a + a] = c; MR[3 * a + b] = -s;
      MR[3 * b + a] = s; MR[3 * b + b] = c;
    } else {
      const double *a = axis;
      const double v = 1.0 - c;
      MR[0] = c + v * a[0] * a[0];        MR[1] = v * a[0] * a[1] - s * a[2]; MR[2] = v * a[0] * a[2] + s * a[1];
      MR[3] = v * a[1] * a[0] + s * a[2]; MR[4] = c + v * a[1] * a[1];        MR[5] = v * a[1] * a[2] - s * a[0];
      MR[6] = v * a[2] * a[0] - s * a[1]; MR[7] = v * a[2] * a[1] + s * a[0]; MR[8] = c + v * a[2] * a[2];
    }
    return;
  }
  if (jt == JT_TRA) { Mp[0] = qv[0]; Mp[1] = qv[1]; Mp[2] = qv[2]; return; }
  if (jt == JT_PLA) { /* JointModelPlanar::calc: rotation about z from (cos, sin) = (q[2], q[3]) as given, translation (x, y, 0) */
    MR[0] = qv[2]; MR[1] = -qv[3]; MR[3] = qv[3]; MR[4] = qv[2];
    Mp[0] = qv[0]; Mp[1] = qv[1];
    return;
  }
  if (jt == JT_ZYX) { /* JointModelSphericalZYX::calc: R = Rz(q0) Ry(q1) Rx(q2) */
    const double c0 = cos(qv[0]), s0 = sin(qv[0]), c1 = cos(qv[1]), s1 = sin(qv[1]), c2 = cos(qv[2]), s2 = sin(qv[2]);
    MR[0] = c0 * c1; MR[1] = c0 * s1 * s2 - s0 * c2; MR[2] = c0 * s1 * c2 + s0 * s2;
    MR[3] = s0 * c1; MR[4] = s0 * s1 * s2 + c0 * c2; MR[5] = s0 * s1 * c2 - c0 * s2;
    MR[6] = -s1;     MR[7] = c1 * s2;                MR[8] = c1 * c2;
    return;
  }
  if (jt == JT_FF || jt == JT_SPH) { /* q = (x, y, z, qx, qy, qz, qw): M = (R(quat), p); spherical: q = (qx, qy, qz, qw), p = 0 */
    const int o = jt == JT_FF ? 3 : 0;
    const double x = qv[o], y = qv[o + 1], z = qv[o + 2], w = qv[o + 3];
    MR[0] = 1 - 2 * (y * y + z * z); MR[1] = 2 * (x * y - z * w);     MR[2] = 2 * (x * z + y * w);
    MR[3] = 2 * (x * y + z * w);     MR[4] = 1 - 2 * (x * x + z * z); MR[5] = 2 * (y * z - x * w);
    MR[6] = 2 * (x * z - y * w);     MR[7] = 2 * (y * z + x * w);     MR[8] = 1 - 2 * (x * x + y * y);
    if (jt == JT_FF) { Mp[0] = qv[0]; Mp[1] = qv[1]; Mp[2] = qv[2]; }
  } else if (jt <= JT_RZ) {
    double s = sin(q), c = cos(q);
    int a = (jt + 1) % 3, b = (jt + 2) % 3; /* rotation in the (a,b) plane */
    MR[3 * a + a] = c; MR[3 * a + b] = -s;
    MR[3 * b + a] = s; MR[3 * b + b] = c;
  } else if (jt <= JT_PZ) {
    Mp[jt - 3] = q;
  } else if (jt == JT_RU) {
    const double *a = axis;
    double s = sin(q), c = cos(q), v = 1.0 - c;
    MR[0] = c + v * a[0] * a[0];        MR[1] = v * a[0] * a[1] - s * a[2]; MR[2] = v * a[0] * a[2] + s * a[1];
    MR[3] = v * a[1] * a[0] + s * a[2]; MR[4] = c + v * a[1] * a[1];        MR[5] = v * a[1] * a[2] - s * a[0];
    MR[6] = v * a[2] * a[0] - s * a[1]; MR[7] = v * a[2] * a[1] + s * a[0]; MR[8] = c + v * a[2] * a[2];
  } else {
    Mp[0] = axis[0] * q; Mp[1] = axis[1] * q; Mp[2] = axis[2] * q;
  }
}

/* inverse of a small SPD matrix (n <= 6, row stride 6) by Cholesky: pinocchio's PerformStYSInversion does
 * Dinv.setIdentity(); StU.llt().solveInPlace(Dinv) */
static void spd_inverse(const double *A, int n, double *Ainv) {
  double L[36] = {0}, Li[36] = {0};
  for (int i = 0; i < n; ++i)
    for (int j = 0; j <= i; ++j) {
      double sum = A[6 * i + j];
      for (int k = 0; k < j; ++k) sum -= L[6 * i + k] * L[6 * j + k];
      L[6 * i + j] = (i == j) ? sqrt(sum) : sum / L[6 * j + j];
    }
  for (int c = 0; c < n; ++c) /* Li = L^-1 by forward substitution */
    for (int i = 0; i < n; ++i) {
      double sum = (i == c) ? 1.0 : 0.0;
      for (int k = 0; k < i; ++k) sum -= L[6 * i + k] * Li[6 * k + c];
      Li[6 * i + c] = sum / L[6 * i + i];
    }
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      double sum = 0.0;
      for (int k = 0; k < n; ++k) sum += Li[6 * k + i] * Li[6 * k + j];
      Ainv[6 * i + j] = sum;
    }
}

/* P1: JointModel*::calc_aba(jdata, armature, I, update_I) (call site hxx:60-63), generic form:
 * U = I S; StU = S^T U + diag(armature); Dinv = StU^-1; UDinv = U Dinv; if update_I: I -= UDinv U^T.
 * For a 1-DoF aligned joint this is pinocchio's U = I.col(k), Dinv = 1/(I(k,k) + armature). */
static void joint_calc_aba(lo_solver *s, int i, const double *armature, double *I, int update_I) {
  const int n = s->nvj[i];
  const double *S = s->jS + 36 * i;
  double *U = s->jU + 36 * i, *UDinv = s->jUDinv + 36 * i, *Dinv = s->jDinv + 36 * i;
  double StU[36] = {0};
  for (int a = 0; a < 6; ++a)
    for (int c = 0; c < n; ++c) {
      double sum = 0.0;
      for (int k = 0; k < 6; ++k) sum += I[6 * a + k] * S[6 * k + c];
      U[6 * a + c] = sum;
    }
  for (int r = 0; r < n; ++r)
    for (int c = 0; c < n; ++c) {
      double sum = 0.0;
      for (int k = 0; k < 6; ++k) sum += S[6 * k + r] * U[6 * k + c];
      StU[6 * r + c] = sum + (r == c ? armature[r] : 0.0);
    }
  if (n == 1) Dinv[0] = 1.0 / StU[0];
  else spd_inverse(StU, n, Dinv);
  for (int a = 0; a < 6; ++a)
    for (int c = 0; c < n; ++c) {
      double sum = 0.0;
      for (int k = 0; k < n; ++k) sum += U[6 * a + k] * Dinv[6 * k + c];
      UDinv[6 * a + c] = sum;
    }
  if (update_I)
    for (int a = 0; a < 6; ++a)
      for (int b = 0; b < 6; ++b) {
        double sum = 0.0;
        for (int k = 0; k < n; ++k) sum += UDinv[6 * a + k] * U[6 * b + k];
        I[6 * a + b] -= sum;
      }
}

/* ------------------------------------------------------------------------------------------ */
/* construction                                                                                 */
/* ------------------------------------------------------------------------------------------ */
static void lo_reset_solver(lo_solver *s);
static void problem_reset(lo_solver *s);

/* FirstOrderLoikOptimizedTpl ctor (loik-loid-optimized.hpp:129-162) + IkIdDataTypeOptimizedTpl ctor
 * (loik-loid-data-optimized.hxx:40-104) + IkProblemFormulationOptimized ctor
 * (ik-id-description-optimized.hpp:30-59; throws unless eq_c_dim == 6). */
LO_API lo_solver *lo_create(int nj, const int *parent, const int *jtype, const double *axis, const double *plR,
                            const double *plp, int max_iter, double tol_abs, double tol_rel, double tol_primal_inf,
                            double tol_dual_inf, double rho, double mu, double mu_equality_scale_factor,
                            int mu_update_strat, int num_eq_c, int eq_c_dim, int warm_start, double tol_tail_solve) {
  if (eq_c_dim != 6 || nj < 2 || num_eq_c < 0) return NULL;
  lo_solver *s = (lo_solver *)calloc(1, sizeof(lo_solver));
  s->nj = nj; s->nb = nj - 1; s->nc = num_eq_c;
  s->parent = (int *)calloc(nj, sizeof(int)); s->jtype = (int *)calloc(nj, sizeof(int));
  memcpy(s->parent, parent, nj * sizeof(int)); memcpy(s->jtype, jtype, nj * sizeof(int));
  s->nvj = (int *)calloc(nj, sizeof(int)); s->idxv = (int *)calloc(nj, sizeof(int)); s->idxq = (int *)calloc(nj, sizeof(int));
  s->nv = 0; s->nq = 0;
  for (int i = 1; i < nj; ++i) {
    s->idxv[i] = s->nv; s->idxq[i] = s->nq;
    s->nvj[i] = jt_nv(jtype[i]);
    s->nv += s->nvj[i]; s->nq += jt_nq(jtype[i]);
  }
  s->axis = dalloc(3 * nj); memcpy(s->axis, axis, 3 * nj * sizeof(double));
  s->plR = dalloc(9 * nj); memcpy(s->plR, plR, 9 * nj * sizeof(double));
  s->plp = dalloc(3 * nj); memcpy(s->plp, plp, 3 * nj * sizeof(double));
  s->rho = rho; s->mu0 = mu; s->mu = mu; s->mu_equality_scale_factor = mu_equality_scale_factor;
  s->mu_update_strat = mu_update_strat; s->max_iter = max_iter;
  s->tol_abs = tol_abs; s->tol_rel = tol_rel; s->tol_primal_inf = tol_primal_inf; s->tol_dual_inf = tol_dual_inf;
  s->warm_start = warm_start; s->tol_tail_solve = tol_tail_solve;
  int nres = 6 * s->nb + s->nv, nc = num_eq_c > 0 ? num_eq_c : 1;
  s->primal_residual_vec = dalloc(nres); s->dual_residual_vec = dalloc(nres);
  s->H_refs = dalloc(36 * nj); s->v_refs = dalloc(6 * nj); s->Hv = dalloc(6 * nj);
  s->task_ids = (int *)calloc(nc, sizeof(int));
  s->Ais = dalloc(36 * nc); s->bis = dalloc(6 * nc); s->AtA = dalloc(36 * nc); s->Atb = dalloc(6 * nc);
  s->lb = dalloc(s->nv); s->ub = dalloc(s->nv);
  s->oMi_R = dalloc(9 * nj); s->oMi_p = dalloc(3 * nj); s->liMi_R = dalloc(9 * nj); s->liMi_p = dalloc(3 * nj);
  for (int i = 0; i < nj; ++i)
    for (int k = 0; k < 3; ++k) s->oMi_R[9 * i + 4 * k] = s->liMi_R[9 * i + 4 * k] = 1.0;
  s->nu = dalloc(s->nv); s->nu_prev = dalloc(s->nv); s->vis = dalloc(6 * nj); s->vis_prev = dalloc(6 * nj);
  s->His = dalloc(36 * nj); s->His_aba = dalloc(36 * nj);
  for (int i = 0; i < nj; ++i)
    for (int k = 0; k < 6; ++k) s->His[36 * i + 7 * k] = s->His_aba[36 * i + 7 * k] = 1.0; /* Mat6x6::Identity() */
  s->pis = dalloc(6 * nj); s->pis_aba = dalloc(6 * nj); s->R = dalloc(s->nv); s->r = dalloc(s->nv);
  s->fis = dalloc(6 * nj); s->delta_fis = dalloc(6 * nj); s->yis = dalloc(6 * nc); s->delta_yis = dalloc(6 * nc);
  s->w = dalloc(s->nv); s->delta_w = dalloc(s->nv); s->z = dalloc(s->nv); s->z_prev = dalloc(s->nv);
  s->Aty = dalloc(6 * nc); s->fis_diff_plus_Aty = dalloc(6 * nj); s->delta_fis_diff_plus_Aty = dalloc(6 * nj);
  s->Href_v = dalloc(6 * nj); s->Av_minus_b = dalloc(6 * nc);
  s->Stf_plus_w = dalloc(s->nv); s->delta_Stf_plus_w = dalloc(s->nv);
  s->jS = dalloc(36 * nj); s->jU = dalloc(36 * nj); s->jDinv = dalloc(36 * nj); s->jUDinv = dalloc(36 * nj);
  for (int i = 1; i < nj; ++i) joint_S(s->jtype[i], s->axis + 3 * i, s->jS + 36 * i);
  s->hist_cap = max_iter + 2; s->hist_mu = dalloc(s->hist_cap); s->hist_pres = dalloc(s->hist_cap);
  s->hist_dres = dalloc(s->hist_cap);
  problem_reset(s);
  lo_reset_solver(s);
  return s;
}

LO_API void lo_destroy(lo_solver *s) {
  if (!s) return;
  void *ptrs[] = {s->nvj, s->idxv, s->idxq, s->parent, s->jtype, s->axis, s->plR, s->plp, s->primal_residual_vec, s->dual_residual_vec, s->H_refs,
                  s->v_refs, s->Hv, s->task_ids, s->Ais, s->bis, s->AtA, s->Atb, s->lb, s->ub, s->oMi_R, s->oMi_p,
                  s->liMi_R, s->liMi_p, s->nu, s->nu_prev, s->vis, s->vis_prev, s->His, s->His_aba, s->pis, s->pis_aba,
                  s->R, s->r, s->fis, s->delta_fis, s->yis, s->delta_yis, s->w, s->delta_w, s->z, s->z_prev, s->Aty,
                  s->fis_diff_plus_Aty, s->delta_fis_diff_plus_Aty, s->Href_v, s->Av_minus_b, s->Stf_plus_w,
                  s->delta_Stf_plus_w, s->jS, s->jU, s->jDinv, s->jUDinv, s->hist_mu, s->hist_pres, s->hist_dres};
  for (size_t i = 0; i < sizeof(ptrs) / sizeof(ptrs[0]); ++i) free(ptrs[i]);
  free(s);
}

/* ------------------------------------------------------------------------------------------ */
/* problem formulation (ik-id-description-optimized.hpp)                                        */
/* ------------------------------------------------------------------------------------------ */
/* Reset(): :61-72, ResetReferences :369-383, ResetEqConstraints :388-411, ResetIneqConstraints :416-420 */
static void problem_reset(lo_solver *s) {
  memset(s->H_refs, 0, 36 * s->nj * sizeof(double));
  memset(s->v_refs, 0, 6 * s->nj * sizeof(double));
  memset(s->Hv, 0, 6 * s->nj * sizeof(double));
  s->Hv_inf_norm = 0.0;
  for (int k = 0; k < s->nc; ++k) s->task_ids[k] = 0;
  memset(s->Ais, 0, 36 * s->nc * sizeof(double)); memset(s->bis, 0, 6 * s->nc * sizeof(double));
  memset(s->AtA, 0, 36 * s->nc * sizeof(double)); memset(s->Atb, 0, 6 * s->nc * sizeof(double));
  s->bis_inf_norm = 0.0;
  memset(s->lb, 0, s->nv * sizeof(double)); memset(s->ub, 0, s->nv * sizeof(double));
}
/* UpdateReference(H_ref, v_ref): :78-97 -- broadcast to every joint incl. 0; Hv_inf_norm = |Hv[0]|inf */
static void problem_update_reference(lo_solver *s, const double *H_ref, const double *v_ref) {
  for (int i = 0; i < s->nj; ++i) {
    memcpy(s->H_refs + 36 * i, H_ref, 36 * sizeof(double));
    memcpy(s->v_refs + 6 * i, v_ref, 6 * sizeof(double));
    mat6_vec(s->H_refs + 36 * i, s->v_refs + 6 * i, s->Hv + 6 * i);
  }
  s->Hv_inf_norm = inf6(s->Hv);
}
/* UpdateReferences(H_refs, v_refs): :103-121 -- per joint; Hv_inf_norm only grows */
LO_API void lo_update_references(lo_solver *s, const double *H_refs, const double *v_refs) {
  memcpy(s->H_refs, H_refs, 36 * s->nj * sizeof(double));
  memcpy(s->v_refs, v_refs, 6 * s->nj * sizeof(double));
  for (int i = 0; i < s->nj; ++i) {
    mat6_vec(s->H_refs + 36 * i, s->v_refs + 6 * i, s->Hv + 6 * i);
    double n = inf6(s->Hv + 6 * i);
    if (n > s->Hv_inf_norm) s->Hv_inf_norm = n;
  }
}
/* UpdateIneqConstraints: :325-339 */
static void problem_update_ineq(lo_solver *s, const double *lb, const double *ub) {
  memcpy(s->lb, lb, s->nv * sizeof(double)); memcpy(s->ub, ub, s->nv * sizeof(double));
}
static void ata_atb(const double *A, const double *b, double *AtA, double *Atb) {
  for (int i = 0; i < 6; ++i) {
    for (int j = 0; j < 6; ++j) {
      double sum = 0.0;
      for (int k = 0; k < 6; ++k) sum += A[6 * k + i] * A[6 * k + j];
      AtA[6 * i + j] = sum;
    }
  }
  mat6T_vec(A, b, Atb);
}
/* UpdateEqConstraints(ids, Ais, bis): :127-171 */
static int problem_update_eq(lo_solver *s, int n_ids, const int *ids, const double *Ais, const double *bis) {
  if (n_ids != s->nc) {
    snprintf(s->err, sizeof s->err, "[IkProblemFormulation::UpdateEqConstraints]: number of equality constraints doesn't match initialization!!!");
    return -1;
  }
  memcpy(s->task_ids, ids, s->nc * sizeof(int));
  memcpy(s->Ais, Ais, 36 * s->nc * sizeof(double)); memcpy(s->bis, bis, 6 * s->nc * sizeof(double));
  s->bis_inf_norm = 0.0;
  for (int k = 0; k < s->nc; ++k) {
    ata_atb(s->Ais + 36 * k, s->bis + 6 * k, s->AtA + 36 * k, s->Atb + 6 * k);
    double n = inf6(s->bis + 6 * k);
    if (n > s->bis_inf_norm) s->bis_inf_norm = n;
  }
  return 0;
}
/* UpdateEqConstraint(c_id, Ai, bi): :178-218 -- bis_inf_norm only grows (quirk 9) */
static int problem_update_eq_one(lo_solver *s, int c_id, const double *Ai, const double *bi) {
  int found = -1, count = 0;
  for (int k = 0; k < s->nc; ++k)
    if (s->task_ids[k] == c_id) { if (found < 0) found = k; ++count; }
  if (found < 0) {
    snprintf(s->err, sizeof s->err, "[IkProblemFormulation::UpdateEqConstraint]: constraint doesn't yet exist at link 'c_id' !!! ");
    return -1;
  }
  if (count > 1) {
    snprintf(s->err, sizeof s->err, "[IkProblemFormulation::UpdateEqConstraint]: multiple constraint specification for the same link id, not supported, terminating !!!");
    return -1;
  }
  memcpy(s->Ais + 36 * found, Ai, 36 * sizeof(double)); memcpy(s->bis + 6 * found, bi, 6 * sizeof(double));
  ata_atb(Ai, bi, s->AtA + 36 * found, s->Atb + 6 * found);
  double n = inf6(bi);
  if (n > s->bis_inf_norm) s->bis_inf_norm = n;
  return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* state resets (loik-loid-data-optimized.hxx)                                                  */
/* ------------------------------------------------------------------------------------------ */
/* Reset(warm_start): :114-127 */
static void data_reset(lo_solver *s, int warm_start) {
  if (!warm_start) {
    memset(s->w, 0, s->nv * sizeof(double)); memset(s->z, 0, s->nv * sizeof(double));
    memset(s->nu, 0, s->nv * sizeof(double));
    memset(s->vis, 0, 6 * s->nj * sizeof(double)); memset(s->fis, 0, 6 * s->nj * sizeof(double));
    memset(s->fis_diff_plus_Aty, 0, 6 * s->nj * sizeof(double));
  }
}
/* ResetRecursion(): :138-154 -- NOT nu, NOT Stf_plus_w (quirk 4) */
static void data_reset_recursion(lo_solver *s) {
  memset(s->w, 0, s->nv * sizeof(double)); memset(s->z, 0, s->nv * sizeof(double));
  memset(s->vis, 0, 6 * s->nj * sizeof(double)); memset(s->fis, 0, 6 * s->nj * sizeof(double));
  memset(s->fis_diff_plus_Aty, 0, 6 * s->nj * sizeof(double));
  memset(s->yis, 0, 6 * s->nc * sizeof(double)); memset(s->Aty, 0, 6 * s->nc * sizeof(double));
}
/* ResetInfNorms(): :165-182 */
LO_API void lo_reset_inf_norms(lo_solver *s) {
  s->bT_delta_y_plus = s->bT_delta_y_minus = 0.0;
  s->Av_inf_norm = s->nu_inf_norm = s->Href_v_inf_norm = 0.0;
  s->fis_diff_plus_Aty_inf_norm = s->Stf_plus_w_inf_norm = 0.0;
  s->delta_fis_diff_plus_Aty_inf_norm = s->delta_Stf_plus_w_inf_norm = 0.0;
  s->delta_vis_inf_norm = s->delta_nu_inf_norm = s->delta_z_inf_norm = 0.0;
  s->delta_fis_inf_norm = s->delta_yis_inf_norm = s->delta_w_inf_norm = 0.0;
}
/* UpdatePrev(): :192-197 */
LO_API void lo_update_prev(lo_solver *s) {
  memcpy(s->vis_prev, s->vis, 6 * s->nj * sizeof(double));
  memcpy(s->nu_prev, s->nu, s->nv * sizeof(double));
  memcpy(s->z_prev, s->z, s->nv * sizeof(double));
}
/* ResetSolver(): loik-loid-optimized.hpp:168-186 + Base::Reset task-solver-base.hpp:73-84 */
static void lo_reset_solver(lo_solver *s) {
  s->iter = 0; s->converged = 0; s->primal_infeasible = 0; s->dual_infeasible = 0;
  s->mu = s->mu0;
  s->tail_solve_iter = 0;
  s->delta_x_qp_inf_norm = s->delta_y_qp_inf_norm = s->A_qp_T_delta_y_qp_inf_norm = 0.0;
  s->ub_qp_T_delta_y_qp_plus = s->lb_qp_T_delta_y_qp_minus = 0.0;
  s->primal_infeasibility_cond_1 = s->primal_infeasibility_cond_2 = 0;
  s->mu_eq = s->mu_equality_scale_factor * s->mu;
  s->mu_ineq = s->mu;
  s->hist_len = 0;
}
LO_API void lo_reset_solver_public(lo_solver *s) { lo_reset_solver(s); }

/* ------------------------------------------------------------------------------------------ */
/* the per-iteration steps (loik-loid-optimized.hxx)                                            */
/* ------------------------------------------------------------------------------------------ */
/* FwdPassInit(q): hxx:253-283 */
LO_API void lo_fwd_pass_init(lo_solver *s, const double *q) {
  for (int i = 1; i < s->nj; ++i) {
    double MR[9], Mp[3], t[3];
    int par = s->parent[i];
    joint_M(s->jtype[i], s->axis + 3 * i, q + s->idxq[i], MR, Mp);
    joint_S_of_q(s->jtype[i], q + s->idxq[i], s->jS + 36 * i); /* jmodel.calc also sets jdata.S for configuration-dependent subspaces */
    /* liMi = jointPlacements[i] * M : (R1 R2, p1 + R1 p2) */
    m3mul(s->plR + 9 * i, MR, s->liMi_R + 9 * i);
    for (int a = 0; a < 3; ++a)
      s->liMi_p[3 * i + a] = s->plp[3 * i + a] + s->plR[9 * i + 3 * a] * Mp[0] + s->plR[9 * i + 3 * a + 1] * Mp[1] +
                             s->plR[9 * i + 3 * a + 2] * Mp[2];
    m3mul(s->oMi_R + 9 * par, s->liMi_R + 9 * i, s->oMi_R + 9 * i);
    for (int a = 0; a < 3; ++a) {
      t[a] = s->oMi_R[9 * par + 3 * a] * s->liMi_p[3 * i] + s->oMi_R[9 * par + 3 * a + 1] * s->liMi_p[3 * i + 1] +
             s->oMi_R[9 * par + 3 * a + 2] * s->liMi_p[3 * i + 2];
      s->oMi_p[3 * i + a] = s->oMi_p[3 * par + a] + t[a];
    }
  }
  if (!s->warm_start) {
    memset(s->yis, 0, 6 * s->nc * sizeof(double));
    memset(s->Aty, 0, 6 * s->nc * sizeof(double));
  }
}

/* FwdPass1(): hxx:290-338 */
LO_API void lo_fwd_pass1(lo_solver *s) {
  for (int k = 0; k < s->nv; ++k) { s->R[k] = 1.0; s->R[k] *= s->mu_ineq; }
  for (int k = 0; k < s->nv; ++k) s->r[k] = s->w[k] - s->mu_ineq * s->z[k];
  for (int i = 1; i < s->nj; ++i) {
    double *H = s->His + 36 * i;
    for (int a = 0; a < 36; ++a) H[a] = (a % 7 == 0) ? 1.0 : 0.0;
    for (int a = 0; a < 36; ++a) H[a] *= s->rho;
    for (int a = 0; a < 36; ++a) H[a] += s->H_refs[36 * i + a];
    memcpy(s->His_aba + 36 * i, H, 36 * sizeof(double));
    for (int a = 0; a < 6; ++a) {
      s->pis[6 * i + a] = -s->rho * s->vis_prev[6 * i + a];
      s->pis[6 * i + a] -= s->Hv[6 * i + a];
    }
    memcpy(s->pis_aba + 6 * i, s->pis + 6 * i, 6 * sizeof(double));
  }
  for (int k = 0; k < s->nc; ++k) {
    int c = s->task_ids[k];
    for (int a = 0; a < 36; ++a) {
      s->His[36 * c + a] += s->mu_eq * s->AtA[36 * k + a];
      s->His_aba[36 * c + a] += s->mu_eq * s->AtA[36 * k + a];
    }
    for (int a = 0; a < 6; ++a) s->pis[6 * c + a] += s->Aty[6 * k + a] - s->mu_eq * s->Atb[6 * k + a];
    memcpy(s->pis_aba + 6 * c, s->pis + 6 * c, 6 * sizeof(double));
  }
}

/* BwdPassOptimizedVisitor(): hxx:345-354 -> LoikBackwardStepVisitor::algo hxx:31-81 */
LO_API void lo_bwd_pass(lo_solver *s) {
  for (int i = s->nj - 1; i > 0; --i) {
    int par = s->parent[i];
    double *Hi_aba = s->His_aba + 36 * i, *pi_aba = s->pis_aba + 6 * i;
    const double *pi = s->pis + 6 * i, *S = s->jS + 36 * i;
    const int n = s->nvj[i], iv = s->idxv[i];
    double tmp36[36], tmp6[6];
    joint_calc_aba(s, i, s->R + iv, Hi_aba, par > 0);                                     /* :60-63 */
    se3_act_on(s->liMi_R + 9 * i, s->liMi_p + 3 * i, Hi_aba, tmp36);                      /* :66 */
    for (int a = 0; a < 36; ++a) s->His_aba[36 * par + a] += tmp36[a];
    memcpy(s->His + 36 * par, s->His_aba + 36 * par, 36 * sizeof(double));                /* :67 */
    for (int c = 0; c < n; ++c) {
      double Stp = 0.0;
      for (int a = 0; a < 6; ++a) Stp += S[6 * a + c] * pi[a];
      s->r[iv + c] += Stp;                                                                /* :70 */
    }
    for (int a = 0; a < 6; ++a)
      for (int c = 0; c < n; ++c) pi_aba[a] -= s->jUDinv[36 * i + 6 * a + c] * s->r[iv + c]; /* :71-73 */
    se3_act_force(s->liMi_R + 9 * i, s->liMi_p + 3 * i, pi_aba, tmp6);                    /* :74 */
    for (int a = 0; a < 6; ++a) s->pis[6 * par + a] += tmp6[a];
    memcpy(s->pis_aba + 6 * par, s->pis + 6 * par, 6 * sizeof(double));                   /* :75 */
  }
}

/* FwdPass2OptimizedVisitor(): hxx:361-377 -> LoikForwardStep2Visitor::algo hxx:102-163 */
LO_API void lo_fwd_pass2(lo_solver *s) {
  memcpy(s->delta_fis_diff_plus_Aty, s->fis_diff_plus_Aty, 6 * s->nj * sizeof(double));   /* :364 */
  for (int i = 1; i < s->nj; ++i) {
    int par = s->parent[i];
    const double *Hi = s->His + 36 * i, *pi = s->pis + 6 * i, *S = s->jS + 36 * i;
    const int nj_ = s->nvj[i], iv = s->idxv[i];
    double vp[6], Hv6[6], d[6], n;
    se3_actinv_motion(s->liMi_R + 9 * i, s->liMi_p + 3 * i, s->vis + 6 * par, vp);        /* :125 */
    for (int c = 0; c < nj_; ++c) {
      double acc = 0.0, dr = 0.0;
      for (int a = 0; a < 6; ++a) acc += s->jUDinv[36 * i + 6 * a + c] * vp[a];
      for (int k = 0; k < nj_; ++k) dr += s->jDinv[36 * i + 6 * c + k] * s->r[iv + k];
      s->nu[iv + c] = -acc - dr;                                                          /* :127 */
    }
    n = infn(s->nu + iv, nj_);
    if (n > s->nu_inf_norm) s->nu_inf_norm = n;                                           /* :129-131 */
    for (int a = 0; a < 6; ++a) s->vis[6 * i + a] = vp[a];
    for (int a = 0; a < 6; ++a)
      for (int c = 0; c < nj_; ++c) s->vis[6 * i + a] += S[6 * a + c] * s->nu[iv + c];    /* :133-134 */
    memcpy(s->delta_fis + 6 * i, s->fis + 6 * i, 6 * sizeof(double));                     /* :137 */
    mat6_vec(Hi, s->vis + 6 * i, Hv6);
    for (int a = 0; a < 6; ++a) s->fis[6 * i + a] = Hv6[a] + pi[a];                       /* :139-140 */
    for (int a = 0; a < 6; ++a) s->delta_fis[6 * i + a] = s->fis[6 * i + a] - s->delta_fis[6 * i + a];
    n = inf6(s->delta_fis + 6 * i);
    if (n > s->delta_fis_inf_norm) s->delta_fis_inf_norm = n;                             /* :144-146 */
    mat6_vec(s->H_refs + 36 * i, s->vis + 6 * i, s->Href_v + 6 * i);                      /* :149 */
    n = inf6(s->Href_v + 6 * i);
    if (n > s->Href_v_inf_norm) s->Href_v_inf_norm = n;
    for (int a = 0; a < 6; ++a) d[a] = s->vis[6 * i + a] - s->vis_prev[6 * i + a];
    n = inf6(d);
    if (n > s->delta_vis_inf_norm) s->delta_vis_inf_norm = n;                             /* :156-158 */
    memset(s->fis_diff_plus_Aty + 6 * i, 0, 6 * sizeof(double));                          /* :370 */
  }
  double m = 0.0;
  for (int k = 0; k < s->nv; ++k) { double a = fabs(s->nu[k] - s->nu_prev[k]); if (a > m) m = a; }
  s->delta_nu_inf_norm = m;                                                               /* :375 */
}

/* BoxProj(): hxx:384-397 */
LO_API void lo_box_proj(lo_solver *s) {
  double m = 0.0;
  for (int k = 0; k < s->nv; ++k) {
    double t = s->nu[k] + (1.0 / s->mu_ineq) * s->w[k];
    t = s->lb[k] > t ? s->lb[k] : t; /* lb.cwiseMax(.) */
    t = s->ub[k] < t ? s->ub[k] : t; /* ub.cwiseMin(.) */
    s->z[k] = t;
    double a = fabs(s->z[k] - s->z_prev[k]);
    if (a > m) m = a;
    s->primal_residual_vec[6 * s->nb + k] = s->nu[k] - s->z[k];
  }
  s->delta_z_inf_norm = m;
}

/* DualUpdate(): hxx:404-461 */
LO_API void lo_dual_update(lo_solver *s) {
  for (int k = 0; k < s->nc; ++k) {
    int c = s->task_ids[k];
    const double *Ai = s->Ais + 36 * k, *bi = s->bis + 6 * k, *vi = s->vis + 6 * c;
    double Av[6], n, plus = 0.0, minus = 0.0;
    mat6_vec(Ai, vi, Av);
    for (int a = 0; a < 6; ++a) s->Av_minus_b[6 * k + a] = Av[a] - bi[a];                 /* :416 */
    for (int a = 0; a < 6; ++a) s->delta_yis[6 * k + a] = s->mu_eq * s->Av_minus_b[6 * k + a];
    for (int a = 0; a < 6; ++a) s->yis[6 * k + a] += s->delta_yis[6 * k + a];             /* :422 */
    mat6T_vec(Ai, s->yis + 6 * k, s->Aty + 6 * k);                                        /* :425 */
    n = inf6(s->delta_yis + 6 * k);
    if (n > s->delta_yis_inf_norm) s->delta_yis_inf_norm = n;
    memcpy(s->primal_residual_vec + 6 * (c - 1), s->Av_minus_b + 6 * k, 6 * sizeof(double)); /* :433 */
    memcpy(s->fis_diff_plus_Aty + 6 * c, s->Aty + 6 * k, 6 * sizeof(double));             /* :438-439 */
    for (int a = 0; a < 6; ++a) {
      double dy = s->delta_yis[6 * k + a];
      plus += bi[a] * (dy > 0.0 ? dy : 0.0);
      minus += bi[a] * (dy < 0.0 ? dy : 0.0);
    }
    s->bT_delta_y_plus += plus; s->bT_delta_y_minus += minus;                             /* :442-443 */
    n = inf6(Av);
    if (n > s->Av_inf_norm) s->Av_inf_norm = n;                                           /* :446-448 */
  }
  double m = 0.0;
  for (int k = 0; k < s->nv; ++k) {
    s->delta_w[k] = s->mu_ineq * (s->nu[k] - s->z[k]);                                    /* :454 */
    s->w[k] += s->delta_w[k];
    double a = fabs(s->delta_w[k]);
    if (a > m) m = a;
  }
  s->delta_w_inf_norm = m;
}

/* BwdPass2OptimizedVisitor(): hxx:468-487 -> LoikBackwardStep2Visitor::algo hxx:185-241 */
static void lo_bwd_pass2(lo_solver *s) {
  memcpy(s->delta_Stf_plus_w, s->Stf_plus_w, s->nv * sizeof(double));                     /* :471 */
  for (int i = s->nj - 1; i > 0; --i) {
    int par = s->parent[i];
    const double *fi = s->fis + 6 * i, *S = s->jS + 36 * i;
    const int nj_ = s->nvj[i], iv = s->idxv[i];
    double t6[6], n;
    for (int a = 0; a < 6; ++a) s->fis_diff_plus_Aty[6 * i + a] += -fi[a];                /* :210 */
    se3_act_force(s->liMi_R + 9 * i, s->liMi_p + 3 * i, fi, t6);
    for (int a = 0; a < 6; ++a) s->fis_diff_plus_Aty[6 * par + a] += t6[a];               /* :212 */
    for (int a = 0; a < 6; ++a)
      s->delta_fis_diff_plus_Aty[6 * i + a] = s->fis_diff_plus_Aty[6 * i + a] - s->delta_fis_diff_plus_Aty[6 * i + a];
    n = inf6(s->delta_fis_diff_plus_Aty + 6 * i);
    if (n > s->delta_fis_diff_plus_Aty_inf_norm) s->delta_fis_diff_plus_Aty_inf_norm = n; /* :218-220 */
    n = inf6(s->fis_diff_plus_Aty + 6 * i);
    if (n > s->fis_diff_plus_Aty_inf_norm) s->fis_diff_plus_Aty_inf_norm = n;             /* :223-225 */
    for (int a = 0; a < 6; ++a)
      s->dual_residual_vec[6 * (i - 1) + a] = s->Href_v[6 * i + a] - s->Hv[6 * i + a] + s->fis_diff_plus_Aty[6 * i + a];
    for (int c = 0; c < nj_; ++c) {
      double Stf = 0.0;
      for (int a = 0; a < 6; ++a) Stf += S[6 * a + c] * fi[a];
      s->Stf_plus_w[iv + c] = Stf + s->w[iv + c];                                         /* :231 */
    }
    n = infn(s->Stf_plus_w + iv, nj_);
    if (n > s->Stf_plus_w_inf_norm) s->Stf_plus_w_inf_norm = n;
  }
  double m = 0.0;
  for (int k = 0; k < s->nv; ++k) {
    s->delta_Stf_plus_w[k] = s->Stf_plus_w[k] - s->delta_Stf_plus_w[k];                   /* :482 */
    double a = fabs(s->delta_Stf_plus_w[k]);
    if (a > m) m = a;
    s->dual_residual_vec[6 * s->nb + k] = s->Stf_plus_w[k];                               /* :484 */
  }
  s->delta_Stf_plus_w_inf_norm = m;
}

/* ComputeResiduals(): hxx:529-533 = ComputePrimalResiduals :494-503 + ComputeDualResiduals :510-522 */
LO_API void lo_compute_residuals(lo_solver *s) {
  int nres = 6 * s->nb + s->nv;
  s->primal_residual = infn(s->primal_residual_vec, nres);
  s->primal_residual_task = infn(s->primal_residual_vec, 6 * s->nb);
  s->primal_residual_slack = infn(s->primal_residual_vec + 6 * s->nb, s->nv);
  lo_bwd_pass2(s);
  s->dual_residual = infn(s->dual_residual_vec, nres);
  s->dual_residual_v = infn(s->dual_residual_vec, 6 * s->nb);
  s->dual_residual_nu = infn(s->dual_residual_vec + 6 * s->nb, s->nv);
}

static double dmax(double a, double b) { return a < b ? b : a; } /* std::max */

/* CheckConvergence(): hxx:540-565 (nu_inf_norm appears twice, quirk 3) */
LO_API void lo_check_convergence(lo_solver *s) {
  s->tol_primal = s->tol_abs + s->tol_rel * dmax(dmax(s->Av_inf_norm, s->nu_inf_norm), dmax(s->bis_inf_norm, s->nu_inf_norm));
  s->tol_dual = s->tol_abs + s->tol_rel * dmax(dmax(s->Href_v_inf_norm, dmax(s->fis_diff_plus_Aty_inf_norm, s->Stf_plus_w_inf_norm)), s->Hv_inf_norm);
  if (s->primal_residual < s->tol_primal && s->dual_residual < s->tol_dual) s->converged = 1;
}

/* CheckFeasibility(): hxx:572-606 */
LO_API void lo_check_feasibility(lo_solver *s) {
  s->delta_y_qp_inf_norm = dmax(s->delta_fis_inf_norm, dmax(s->delta_yis_inf_norm, s->delta_w_inf_norm));
  s->A_qp_T_delta_y_qp_inf_norm = dmax(s->delta_fis_diff_plus_Aty_inf_norm, s->delta_Stf_plus_w_inf_norm);
  s->primal_infeasibility_cond_1 = s->A_qp_T_delta_y_qp_inf_norm <= s->tol_primal_inf * s->delta_y_qp_inf_norm;
  double up = 0.0, lm = 0.0;
  for (int k = 0; k < s->nv; ++k) {
    double dw = s->delta_w[k];
    up += s->ub[k] * (dw > 0.0 ? dw : 0.0);
    lm += s->lb[k] * (dw < 0.0 ? dw : 0.0);
  }
  s->ub_qp_T_delta_y_qp_plus = s->bT_delta_y_plus; s->ub_qp_T_delta_y_qp_plus += up;
  s->lb_qp_T_delta_y_qp_minus = s->bT_delta_y_minus; s->lb_qp_T_delta_y_qp_minus += lm;
  s->primal_infeasibility_cond_2 = (s->ub_qp_T_delta_y_qp_plus + s->lb_qp_T_delta_y_qp_minus) <= s->tol_primal_inf * s->delta_y_qp_inf_norm;
  if (s->primal_infeasibility_cond_1 && s->primal_infeasibility_cond_2) s->primal_infeasible = 1;
  s->delta_x_qp_inf_norm = dmax(s->delta_vis_inf_norm, s->delta_nu_inf_norm);
}

/* UpdateMu(): hxx:613-641 -- only DEFAULT (0) is implemented by the reference; others throw */
LO_API int lo_update_mu(lo_solver *s) {
  if (s->mu_update_strat != 0) {
    snprintf(s->err, sizeof s->err, "[FirstOrderLoikOptimizedTpl::UpdateMu]: mu update strategy not yet implemented");
    return -1;
  }
  if (s->primal_residual > 10 * s->dual_residual) {
    s->mu *= 10;
    s->mu_eq = s->mu_equality_scale_factor * s->mu; s->mu_ineq = s->mu;
  } else if (s->dual_residual > 10 * s->primal_residual) {
    s->mu *= 0.1;
    s->mu_eq = s->mu_equality_scale_factor * s->mu; s->mu_ineq = s->mu;
  }
  return 0;
}

static void hist_push(lo_solver *s) {
  if (s->hist_len < s->hist_cap) {
    s->hist_mu[s->hist_len] = s->mu; s->hist_pres[s->hist_len] = s->primal_residual;
    s->hist_dres[s->hist_len] = s->dual_residual; s->hist_len++;
  }
}

static void one_iteration(lo_solver *s) { /* the 8 calls shared by the main loop and the tail loop */
  lo_update_prev(s);
  lo_reset_inf_norms(s);
  lo_fwd_pass1(s);
  lo_bwd_pass(s);
  lo_fwd_pass2(s);
  lo_box_proj(s);
  lo_dual_update(s);
  lo_compute_residuals(s);
}

/* InfeasibilityTailSolve(): loik-loid-optimized.hpp:271-319 */
static void tail_solve(lo_solver *s) {
  s->tail_solve_iter = 0;
  while (s->delta_x_qp_inf_norm >= s->tol_tail_solve || s->delta_z_inf_norm >= s->tol_tail_solve) {
    if (s->iter >= s->max_iter) return;
    s->iter++;
    s->tail_solve_iter++;
    one_iteration(s);
    s->delta_x_qp_inf_norm = dmax(s->delta_vis_inf_norm, s->delta_nu_inf_norm);
    hist_push(s);
  }
}

/* the main loop shared by the three Solve overloads: hpp:377-454 / :502-579 / :616-693 */
static int main_loop(lo_solver *s) {
  for (int i = 1; i < s->max_iter; i++) {
    s->iter = i;
    one_iteration(s);
    hist_push(s);
    lo_check_convergence(s);
    if (s->iter > 1) lo_check_feasibility(s);
    if (s->converged) break;
    else if (s->primal_infeasible) { tail_solve(s); break; }
    else if (s->dual_infeasible) { tail_solve(s); break; }
    if (lo_update_mu(s)) return -1;
  }
  return 0;
}

/* SolveInit(q, H_ref, v_ref, ids, Ais, bis, lb, ub): hpp:335-361 */
LO_API int lo_solve_init(lo_solver *s, const double *q, const double *H_ref, const double *v_ref, int n_ids,
                         const int *ids, const double *Ais, const double *bis, const double *lb, const double *ub) {
  problem_reset(s);
  data_reset(s, s->warm_start);
  lo_reset_solver(s);
  problem_update_reference(s, H_ref, v_ref);
  problem_update_ineq(s, lb, ub);
  if (problem_update_eq(s, n_ids, ids, Ais, bis)) return -1;
  lo_fwd_pass_init(s, q);
  return 0;
}
/* Solve(): hpp:368-455 */
LO_API int lo_solve(lo_solver *s) {
  data_reset_recursion(s);
  lo_reset_solver(s);
  return main_loop(s);
}
/* Solve(q, H_ref, v_ref, ids, Ais, bis, lb, ub): hpp:475-580 */
LO_API int lo_solve_full(lo_solver *s, const double *q, const double *H_ref, const double *v_ref, int n_ids,
                         const int *ids, const double *Ais, const double *bis, const double *lb, const double *ub) {
  if (lo_solve_init(s, q, H_ref, v_ref, n_ids, ids, Ais, bis, lb, ub)) return -1;
  return main_loop(s);
}
/* Solve(q, c_id, Ai, bi): hpp:596-695 -- the tailored / trajectory-tracking form */
LO_API int lo_solve_task(lo_solver *s, const double *q, int c_id, const double *Ai, const double *bi) {
  data_reset(s, s->warm_start);
  lo_reset_solver(s);
  if (problem_update_eq_one(s, c_id, Ai, bi)) return -1;
  lo_fwd_pass_init(s, q);
  return main_loop(s);
}

/* ------------------------------------------------------------------------------------------ */
/* accessors for the ctypes harness                                                             */
/* ------------------------------------------------------------------------------------------ */
LO_API const char *lo_last_error(lo_solver *s) { return s->err; }
LO_API void lo_set_max_iter(lo_solver *s, int m) {
  s->max_iter = m;
  if (m + 2 > s->hist_cap) {
    s->hist_cap = m + 2;
    s->hist_mu = (double *)realloc(s->hist_mu, s->hist_cap * sizeof(double));
    s->hist_pres = (double *)realloc(s->hist_pres, s->hist_cap * sizeof(double));
    s->hist_dres = (double *)realloc(s->hist_dres, s->hist_cap * sizeof(double));
  }
}
LO_API void lo_set_warm_start(lo_solver *s, int ws) { s->warm_start = ws; }

#define FIELD(name, ptr, len) if (!strcmp(field, name)) { *n = (len); return (ptr); }
LO_API double *lo_array(lo_solver *s, const char *field, int *n) {
  int nj = s->nj, nv = s->nv, nc = s->nc;
  FIELD("liMi_R", s->liMi_R, 9 * nj) FIELD("liMi_p", s->liMi_p, 3 * nj) FIELD("oMi_R", s->oMi_R, 9 * nj)
  FIELD("oMi_p", s->oMi_p, 3 * nj) FIELD("nu", s->nu, nv) FIELD("nu_prev", s->nu_prev, nv) FIELD("vis", s->vis, 6 * nj)
  FIELD("vis_prev", s->vis_prev, 6 * nj) FIELD("His", s->His, 36 * nj) FIELD("His_aba", s->His_aba, 36 * nj)
  FIELD("pis", s->pis, 6 * nj) FIELD("pis_aba", s->pis_aba, 6 * nj) FIELD("R", s->R, nv) FIELD("r", s->r, nv)
  FIELD("fis", s->fis, 6 * nj) FIELD("delta_fis", s->delta_fis, 6 * nj) FIELD("yis", s->yis, 6 * nc)
  FIELD("delta_yis", s->delta_yis, 6 * nc) FIELD("w", s->w, nv) FIELD("delta_w", s->delta_w, nv) FIELD("z", s->z, nv)
  FIELD("z_prev", s->z_prev, nv) FIELD("Aty", s->Aty, 6 * nc) FIELD("fis_diff_plus_Aty", s->fis_diff_plus_Aty, 6 * nj)
  FIELD("delta_fis_diff_plus_Aty", s->delta_fis_diff_plus_Aty, 6 * nj) FIELD("Href_v", s->Href_v, 6 * nj)
  FIELD("Av_minus_b", s->Av_minus_b, 6 * nc) FIELD("Stf_plus_w", s->Stf_plus_w, nv)
  FIELD("delta_Stf_plus_w", s->delta_Stf_plus_w, nv) FIELD("U_full", s->jU, 36 * nj) FIELD("Dinv_full", s->jDinv, 36 * nj)
  FIELD("UDinv_full", s->jUDinv, 36 * nj) FIELD("S_full", s->jS, 36 * nj)
  FIELD("primal_residual_vec", s->primal_residual_vec, 6 * s->nb + nv)
  FIELD("dual_residual_vec", s->dual_residual_vec, 6 * s->nb + nv)
  FIELD("Hv", s->Hv, 6 * nj) FIELD("H_refs", s->H_refs, 36 * nj) FIELD("AtA", s->AtA, 36 * nc) FIELD("Atb", s->Atb, 6 * nc)
  FIELD("hist_mu", s->hist_mu, s->hist_len) FIELD("hist_primal_residual", s->hist_pres, s->hist_len)
  FIELD("hist_dual_residual", s->hist_dres, s->hist_len)
  *n = 0;
  return NULL;
}
#undef FIELD
#define SCAL(name, val) if (!strcmp(field, name)) return (double)(val);
LO_API double lo_scalar(lo_solver *s, const char *field) {
  SCAL("iter", s->iter) SCAL("converged", s->converged) SCAL("primal_infeasible", s->primal_infeasible)
  SCAL("dual_infeasible", s->dual_infeasible) SCAL("mu", s->mu) SCAL("mu_eq", s->mu_eq) SCAL("mu_ineq", s->mu_ineq)
  SCAL("rho", s->rho) SCAL("tol_primal", s->tol_primal) SCAL("tol_dual", s->tol_dual)
  SCAL("primal_residual", s->primal_residual) SCAL("dual_residual", s->dual_residual)
  SCAL("primal_residual_task", s->primal_residual_task) SCAL("primal_residual_slack", s->primal_residual_slack)
  SCAL("dual_residual_v", s->dual_residual_v) SCAL("dual_residual_nu", s->dual_residual_nu)
  SCAL("tail_solve_iter", s->tail_solve_iter) SCAL("delta_x_qp_inf_norm", s->delta_x_qp_inf_norm)
  SCAL("delta_y_qp_inf_norm", s->delta_y_qp_inf_norm) SCAL("A_qp_T_delta_y_qp_inf_norm", s->A_qp_T_delta_y_qp_inf_norm)
  SCAL("ub_qp_T_delta_y_qp_plus", s->ub_qp_T_delta_y_qp_plus) SCAL("lb_qp_T_delta_y_qp_minus", s->lb_qp_T_delta_y_qp_minus)
  SCAL("primal_infeasibility_cond_1", s->primal_infeasibility_cond_1)
  SCAL("primal_infeasibility_cond_2", s->primal_infeasibility_cond_2)
  SCAL("bis_inf_norm", s->bis_inf_norm) SCAL("Hv_inf_norm", s->Hv_inf_norm)
  SCAL("bT_delta_y_plus", s->bT_delta_y_plus) SCAL("bT_delta_y_minus", s->bT_delta_y_minus)
  SCAL("Av_inf_norm", s->Av_inf_norm) SCAL("nu_inf_norm", s->nu_inf_norm) SCAL("Href_v_inf_norm", s->Href_v_inf_norm)
  SCAL("fis_diff_plus_Aty_inf_norm", s->fis_diff_plus_Aty_inf_norm) SCAL("Stf_plus_w_inf_norm", s->Stf_plus_w_inf_norm)
  SCAL("delta_fis_diff_plus_Aty_inf_norm", s->delta_fis_diff_plus_Aty_inf_norm)
  SCAL("delta_Stf_plus_w_inf_norm", s->delta_Stf_plus_w_inf_norm) SCAL("delta_vis_inf_norm", s->delta_vis_inf_norm)
  SCAL("delta_nu_inf_norm", s->delta_nu_inf_norm) SCAL("delta_z_inf_norm", s->delta_z_inf_norm)
  SCAL("delta_fis_inf_norm", s->delta_fis_inf_norm) SCAL("delta_yis_inf_norm", s->delta_yis_inf_norm)
  SCAL("delta_w_inf_norm", s->delta_w_inf_norm) SCAL("hist_len", s->hist_len)
  return NAN;
}
#undef SCAL

/* ------------------------------------------------------------------------------------------ */
/* batch driver: the CPU baseline.  One solver+data per thread, as the    */
/* reference's threading model allows; instances [lo,hi) per thread.  Inputs are batch-major.    */
/* mode 0: per instance SolveInit + Solve() to convergence;  mode 1: fixed `fixed_iters`         */
/* iterations with stopping disabled (iteration-rate measurement).                               */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
  int nj; const int *parent, *jtype; const double *axis, *plR, *plp;
  int max_iter; double tol_abs, tol_rel, tol_primal_inf, tol_dual_inf, rho, mu, mu_eq_scale; int strat, nc;
  double tol_tail;
  const double *q, *H_ref, *v_ref; const int *ids; const double *Ais, *bis, *lb, *ub;
  int b_per_instance, bounds_per_instance;
  int lo, hi, mode, fixed_iters;
  double *z, *nu, *w, *y; int *iters, *status; double *mu_out;
  long total_iters;
} lo_batch_job;

static void *batch_worker(void *arg) {
  lo_batch_job *j = (lo_batch_job *)arg;
  int nc = j->nc;
  lo_solver *s = lo_create(j->nj, j->parent, j->jtype, j->axis, j->plR, j->plp, j->max_iter, j->tol_abs, j->tol_rel,
                           j->tol_primal_inf, j->tol_dual_inf, j->rho, j->mu, j->mu_eq_scale, j->strat, nc, 6, 0, j->tol_tail);
  const int nv = s->nv, nq = s->nq;
  long tot = 0;
  for (int b = j->lo; b < j->hi; ++b) {
    const double *bis = j->b_per_instance ? j->bis + (size_t)b * 6 * nc : j->bis;
    const double *lb = j->bounds_per_instance ? j->lb + (size_t)b * nv : j->lb;
    const double *ub = j->bounds_per_instance ? j->ub + (size_t)b * nv : j->ub;
    lo_solve_init(s, j->q + (size_t)b * nq, j->H_ref, j->v_ref, nc, j->ids, j->Ais, bis, lb, ub);
    if (j->mode == 0) {
      lo_solve(s);
    } else {
      data_reset_recursion(s);
      lo_reset_solver(s);
      for (int i = 1; i <= j->fixed_iters; ++i) {
        s->iter = i;
        one_iteration(s);
        lo_check_convergence(s);
        if (s->iter > 1) lo_check_feasibility(s);
        lo_update_mu(s);
      }
    }
    tot += s->iter;
    if (j->z) memcpy(j->z + (size_t)b * nv, s->z, nv * sizeof(double));
    if (j->nu) memcpy(j->nu + (size_t)b * nv, s->nu, nv * sizeof(double));
    if (j->w) memcpy(j->w + (size_t)b * nv, s->w, nv * sizeof(double));
    if (j->y) memcpy(j->y + (size_t)b * 6 * nc, s->yis, 6 * nc * sizeof(double));
    if (j->iters) j->iters[b] = s->iter;
    if (j->status) j->status[b] = (s->converged ? 1 : 0) | (s->primal_infeasible ? 2 : 0);
    if (j->mu_out) j->mu_out[b] = s->mu;
  }
  j->total_iters = tot;
  lo_destroy(s);
  return NULL;
}

LO_API long lo_batch_solve(int nj, const int *parent, const int *jtype, const double *axis, const double *plR,
                           const double *plp, int max_iter, double tol_abs, double tol_rel, double tol_primal_inf,
                           double tol_dual_inf, double rho, double mu, double mu_eq_scale, int strat, int nc,
                           double tol_tail, int batch, const double *q, const double *H_ref, const double *v_ref,
                           const int *ids, const double *Ais, const double *bis, int b_per_instance, const double *lb,
                           const double *ub, int bounds_per_instance, int mode, int fixed_iters, int nthreads, double *z,
                           double *nu, double *w, double *y, int *iters, int *status, double *mu_out) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > batch) nthreads = batch > 0 ? batch : 1;
  lo_batch_job *jobs = (lo_batch_job *)calloc(nthreads, sizeof(lo_batch_job));
  pthread_t *th = (pthread_t *)calloc(nthreads, sizeof(pthread_t));
  for (int t = 0; t < nthreads; ++t) {
    lo_batch_job *j = &jobs[t];
    j->nj = nj; j->parent = parent; j->jtype = jtype; j->axis = axis; j->plR = plR; j->plp = plp;
    j->max_iter = max_iter; j->tol_abs = tol_abs; j->tol_rel = tol_rel; j->tol_primal_inf = tol_primal_inf;
    j->tol_dual_inf = tol_dual_inf; j->rho = rho; j->mu = mu; j->mu_eq_scale = mu_eq_scale; j->strat = strat; j->nc = nc;
    j->tol_tail = tol_tail; j->q = q; j->H_ref = H_ref; j->v_ref = v_ref; j->ids = ids; j->Ais = Ais; j->bis = bis;
    j->lb = lb; j->ub = ub; j->b_per_instance = b_per_instance; j->bounds_per_instance = bounds_per_instance;
    j->lo = (int)((long)batch * t / nthreads); j->hi = (int)((long)batch * (t + 1) / nthreads);
    j->mode = mode; j->fixed_iters = fixed_iters;
    j->z = z; j->nu = nu; j->w = w; j->y = y; j->iters = iters; j->status = status; j->mu_out = mu_out;
    if (nthreads == 1) batch_worker(j);
    else pthread_create(&th[t], NULL, batch_worker, j);
  }
  long tot = 0;
  for (int t = 0; t < nthreads; ++t) {
    if (nthreads > 1) pthread_join(th[t], NULL);
    tot += jobs[t].total_iters;
  }
  free(jobs); free(th);
  return tot;
}

/* ------------------------------------------------------------------------------------------ */
/* trajectory-tracking driver (the reference's real-time use, hpp:596-695; SURVEY.md 8(f) ranks  */
/* 2-3): per instance one full Solve, then `steps` times { q <- integrate(q, dt z) (user side:  */
/* pinocchio::integrate for 1-DoF joints); Solve(q, c_id, A, b_t) } with                          */
/* b_t = (1 - t/steps) b0 + (t/steps) b1.  Outputs: final z and the iteration count of every step.*/
/* ------------------------------------------------------------------------------------------ */
typedef struct {
  lo_batch_job base;
  const double *bis1; double dt; int steps, warm, c_id;
  int *step_iters; /* [batch][steps] */
  double *q_out;   /* [batch][nq] */
} lo_track_job;

static void integrate_q(const lo_solver *s, double *q, const double *v, double dt) {
  for (int i = 1; i < s->nj; ++i) {
    const int jt = s->jtype[i], iq = s->idxq[i], iv = s->idxv[i];
    if (jt >= JT_RUBX && jt <= JT_RUBU) { /* SpecialOrthogonalOperationTpl<2>::integrate_impl */
      const double ca = q[iq], sa = q[iq + 1], om = dt * v[iv];
      const double co = cos(om), so = sin(om);
      double c = co * ca - so * sa, sn = so * ca + co * sa;
      const double k = (3.0 - (c * c + sn * sn)) / 2.0;
      q[iq] = c * k; q[iq + 1] = sn * k;
    } else if (jt == JT_FF || jt == JT_SPH) {
      /* SpecialOrthogonalOperationTpl<3> / SpecialEuclideanOperationTpl<3>::integrate_impl: quat * exp3(omega) (and
       * p + R(quat) * [translation of exp6(v)]), body-frame velocities, then quaternion::firstOrderNormalize */
      const int o = jt == JT_FF ? 3 : 0;
      double w[3] = {dt * v[iv + o], dt * v[iv + o + 1], dt * v[iv + o + 2]};
      double *qt = q + iq + o;
      const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2], th = sqrt(th2);
      const int small = th < 1e-4;
      const double k = small ? 0.5 - th2 / 48.0 : sin(th / 2) / th, cw = small ? 1.0 - th2 / 8.0 : cos(th / 2);
      const double d[4] = {k * w[0], k * w[1], k * w[2], cw};
      const double ax = qt[0], ay = qt[1], az = qt[2], aw = qt[3];
      double r[4] = {aw * d[0] + ax * d[3] + ay * d[2] - az * d[1], aw * d[1] - ax * d[2] + ay * d[3] + az * d[0],
                     aw * d[2] + ax * d[1] - ay * d[0] + az * d[3], aw * d[3] - ax * d[0] - ay * d[1] - az * d[2]};
      if (jt == JT_FF) {
        const double vl[3] = {dt * v[iv], dt * v[iv + 1], dt * v[iv + 2]};
        const double a_v = small ? 1.0 - th2 / 6.0 : sin(th) / th, a_wxv = small ? 0.5 - th2 / 24.0 : (1.0 - cos(th)) / th2;
        const double a_w = (small ? 1.0 / 6.0 - th2 / 120.0 : (1.0 - a_v) / th2) * (w[0] * vl[0] + w[1] * vl[1] + w[2] * vl[2]);
        double cr[3], p[3], MR[9], Mp[3];
        cross3(w, vl, cr);
        for (int c = 0; c < 3; ++c) p[c] = a_v * vl[c] + a_w * w[c] + a_wxv * cr[c];
        joint_M(JT_SPH, NULL, qt, MR, Mp); /* R(quat) */
        for (int c = 0; c < 3; ++c) q[iq + c] += MR[3 * c] * p[0] + MR[3 * c + 1] * p[1] + MR[3 * c + 2] * p[2];
        if (r[0] * ax + r[1] * ay + r[2] * az + r[3] * aw < 0.0) for (int c = 0; c < 4; ++c) r[c] = -r[c];
      }
      const double nrm = (3.0 - (r[0] * r[0] + r[1] * r[1] + r[2] * r[2] + r[3] * r[3])) / 2.0;
      for (int c = 0; c < 4; ++c) qt[c] = r[c] * nrm;
    } else if (jt == JT_PLA) { /* SpecialEuclideanOperationTpl<2>::integrate_impl: (R0, t0) * exp(v) */
      const double c0 = q[iq + 2], s0 = q[iq + 3], vx = dt * v[iv], vy = dt * v[iv + 1], om = dt * v[iv + 2];
      const double cv = cos(om), sv = sin(om);
      double tx = vx, ty = vy;
      if (fabs(om) > 1e-14) { const double ax = -vy / om, ay = vx / om; tx = ax - (cv * ax - sv * ay); ty = ay - (sv * ax + cv * ay); }
      q[iq] += c0 * tx - s0 * ty; q[iq + 1] += s0 * tx + c0 * ty;
      q[iq + 2] = c0 * cv - s0 * sv; q[iq + 3] = s0 * cv + c0 * sv;
    } else { /* vector-space joints (1-DoF, translation) */
      for (int k = 0; k < s->nvj[i]; ++k) q[iq + k] += dt * v[iv + k];
    }
  }
}

static void *track_worker(void *arg) {
  lo_track_job *t = (lo_track_job *)arg;
  lo_batch_job *j = &t->base;
  const int nc = j->nc;
  lo_solver *s = lo_create(j->nj, j->parent, j->jtype, j->axis, j->plR, j->plp, j->max_iter, j->tol_abs, j->tol_rel,
                           j->tol_primal_inf, j->tol_dual_inf, j->rho, j->mu, j->mu_eq_scale, j->strat, nc, 6, t->warm, j->tol_tail);
  const int nv = s->nv, nq = s->nq;
  double *q = (double *)malloc(nq * sizeof(double));
  int slot = 0;
  for (int k = 0; k < nc; ++k) if (j->ids[k] == t->c_id) slot = k;
  long tot = 0;
  for (int b = j->lo; b < j->hi; ++b) {
    const double *b0 = j->bis + (size_t)b * 6 * nc, *b1 = t->bis1 + (size_t)b * 6 * nc;
    memcpy(q, j->q + (size_t)b * nq, nq * sizeof(double));
    s->warm_start = 0; /* a fresh solver object per instance: its first Solve starts from zero state */
    lo_solve_full(s, q, j->H_ref, j->v_ref, nc, j->ids, j->Ais, b0, j->lb, j->ub);
    s->warm_start = t->warm;
    for (int st = 1; st <= t->steps; ++st) {
      const double a = (double)st / t->steps;
      double bt[6];
      for (int c = 0; c < 6; ++c) bt[c] = (1.0 - a) * b0[6 * slot + c] + a * b1[6 * slot + c];
      integrate_q(s, q, s->z, t->dt);
      lo_solve_task(s, q, t->c_id, j->Ais + 36 * slot, bt);
      tot += s->iter;
      if (t->step_iters) t->step_iters[(size_t)b * t->steps + st - 1] = s->iter;
    }
    if (j->z) memcpy(j->z + (size_t)b * nv, s->z, nv * sizeof(double));
    if (t->q_out) memcpy(t->q_out + (size_t)b * nq, q, nq * sizeof(double));
  }
  j->total_iters = tot;
  free(q);
  lo_destroy(s);
  return NULL;
}

LO_API long lo_batch_track(int nj, const int *parent, const int *jtype, const double *axis, const double *plR, const double *plp,
                           int max_iter, double tol_abs, double tol_rel, double tol_primal_inf, double tol_dual_inf, double rho,
                           double mu, double mu_eq_scale, int strat, int nc, double tol_tail, int batch, const double *q,
                           const double *H_ref, const double *v_ref, const int *ids, const double *Ais, const double *bis0,
                           const double *bis1, const double *lb, const double *ub, int c_id, double dt, int steps, int warm,
                           int nthreads, double *z, double *q_out, int *step_iters) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > batch) nthreads = batch > 0 ? batch : 1;
  lo_track_job *jobs = (lo_track_job *)calloc(nthreads, sizeof(lo_track_job));
  pthread_t *th = (pthread_t *)calloc(nthreads, sizeof(pthread_t));
  for (int t = 0; t < nthreads; ++t) {
    lo_batch_job *j = &jobs[t].base;
    j->nj = nj; j->parent = parent; j->jtype = jtype; j->axis = axis; j->plR = plR; j->plp = plp;
    j->max_iter = max_iter; j->tol_abs = tol_abs; j->tol_rel = tol_rel; j->tol_primal_inf = tol_primal_inf;
    j->tol_dual_inf = tol_dual_inf; j->rho = rho; j->mu = mu; j->mu_eq_scale = mu_eq_scale; j->strat = strat; j->nc = nc;
    j->tol_tail = tol_tail; j->q = q; j->H_ref = H_ref; j->v_ref = v_ref; j->ids = ids; j->Ais = Ais; j->bis = bis0;
    j->lb = lb; j->ub = ub;
    j->lo = (int)((long)batch * t / nthreads); j->hi = (int)((long)batch * (t + 1) / nthreads);
    j->z = z;
    jobs[t].bis1 = bis1; jobs[t].dt = dt; jobs[t].steps = steps; jobs[t].warm = warm; jobs[t].c_id = c_id;
    jobs[t].step_iters = step_iters; jobs[t].q_out = q_out;
    if (nthreads == 1) track_worker(&jobs[t]);
    else pthread_create(&th[t], NULL, track_worker, &jobs[t]);
  }
  long tot = 0;
  for (int t = 0; t < nthreads; ++t) {
    if (nthreads > 1) pthread_join(th[t], NULL);
    tot += jobs[t].base.total_iters;
  }
  free(jobs); free(th);
  return tot;
}

/* ------------------------------------------------------------------------------------------ */
/* the reference's own timing protocol (tests/loik-loid.cpp:987-1032): SolveInit once, then n x */
/* Solve() on one thread.  Returns seconds; *iters_out = iterations of the last solve.          */
/* ------------------------------------------------------------------------------------------ */
LO_API double lo_time_solve(lo_solver *s, int n, int *iters_out) {
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (int k = 0; k < n; ++k) lo_solve(s);
  clock_gettime(CLOCK_MONOTONIC, &t1);
  if (iters_out) *iters_out = s->iter;
  return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

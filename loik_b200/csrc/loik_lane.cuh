// loik_lane.cuh -- the lane-parallel, shared-memory-resident iteration kernel (k_iterate_lane).
//
// Mapping: ONE GROUP OF 8 LANES = ONE PROBLEM INSTANCE (4 instances per warp), the whole per-instance state of the
// ADMM loop resident in SHARED MEMORY for as many iterations as the instance needs, persistent CTAs pulling instances
// from a device-side work queue.  Where k_iterate (loik_device.cuh) gives one thread a whole instance and streams
// its state through HBM once per iteration, this kernel
//   * cuts the dependent chain of a joint step ~4x: lanes 0..5 own the columns of the 6x6 H (its rows, H being
//     symmetric), lane 6 owns p as a seventh column, so the congruence X* H X*^T (pinocchio SE3actOn, call site
//     hxx:66) is two column transforms (2 x 24 FMA per lane, the same act_force() the force vectors use) around a
//     6x6 transpose through shared memory instead of ~270 FMA in one thread; the 6x6 mat-vecs H v, Href v, A v, A^T y
//     are one dot product per lane;
//   * touches HBM twice per SOLVE (state in, results out) instead of ~8 KB per instance and iteration: the
//     backward->forward workspace (His, pis, UDinv, Dinv, r) never leaves the SM;
//   * has no lock-step with 31 unrelated instances: a group that finishes pulls the next instance from the queue, so
//     a launch needs no re-pack rounds and the stragglers of a batch (0.06 % of the Panda instances run all 199
//     iterations) cost 6.9 us per iteration instead of 27 us.
// The maths restates loik-loid-optimized.hxx exactly as loik_device.cuh does (same per-element expressions wherever
// an element is produced by one lane; reference file:line cited there); running inf-norms are kept as per-lane
// partial maxima (max is exactly associative) and combined once per iteration.
//
// Two geometries (template parameter GPI = groups per instance):
//   GPI = 1  four instances per warp, one 8-lane group each, every group sweeping the whole tree (chains: Panda, UR10);
//   GPI = 4  ONE instance per warp, its four groups sweeping different chains of a branching tree at the same time, step by
//            step through a host-built, list-scheduled table (build_wide_table, loik_solver.cu), the instance record shared in
//            shared memory, a __syncwarp() per step.  Talos: 10 steps per sweep instead of 32 (32 joint steps on 4 x 10 group
//            steps), 12.3 us per iteration of a lone instance instead of 31.5.
// Scope: trees of 1-DoF joints (every BASELINE robot); models with multi-DoF joints keep the k_iterate path.
#pragma once
#include "loik_device.cuh"

namespace loik {

constexpr int kLaneI = 4;  // instances per warp (8 lanes each)

// ---- instance record in shared memory (doubles) ------------------------------------------------------------------
enum : int { LS_MU = 0, LS_BINF = 1, LS_CTL = 2, LS_RES = 3, LS_ROWS = 8 };
// per joint: the first 26 entries mirror rows JR_V .. JR_UB of the tile record (loik_device.cuh), then the workspace of
// the backward sweep: Dinv, r, UDinv (8), and [H | p] as 6 rows of 8 (column c < 6: H(:, c), column 6: p); last liMi =
// jointPlacements[i] * M_i(q) (FwdPassInit, hxx:263-264), built once when the instance is loaded (q does not change
// during a solve) instead of three times per joint and iteration.
enum : int { LJ_V = 0, LJ_F = 6, LJ_FD = 12, LJ_NU = 18, LJ_Z = 19, LJ_W = 20, LJ_T = 21, LJ_SQ = 22, LJ_CQ = 23, LJ_LB = 24,
             LJ_UB = 25, LJ_COPY = 26, LJ_DINV = 26, LJ_R = 27, LJ_UD = 28, LJ_HP = 36, LJ_XF = 84, LJ_ROWS = 96,    // XF: liMi = (R 9, t 3)
             // GPI = 4: the four groups of a warp work on DIFFERENT joints of one record at the same time; a joint's block starts
             // 4 x (its group) doubles into a 108-double slot (JointC::loff), so that the groups' 16 B broadcast loads hit
             // four different bank quads and their 64 B per-lane accesses pair up into two wavefronts
             LJ_SLOT_WIDE = 108 };
static_assert((int)LJ_V == (int)JR_V && (int)LJ_F == (int)JR_F && (int)LJ_FD == (int)JR_FD && (int)LJ_NU == (int)JR_NU && (int)LJ_Z == (int)JR_Z && (int)LJ_W == (int)JR_W && (int)LJ_T == (int)JR_T &&
              (int)LJ_SQ == (int)JR_JQ && (int)LJ_LB == (int)JR_LB && (int)LJ_UB == (int)JR_UB, "the copied part of a joint record mirrors the tile rows");
enum : int { LT_Y = TR_Y, LT_ATY = TR_ATY, LT_B = TR_B, LT_ATB = TR_ATB, LT_ROWS = 24 };  // (the tile's task block also holds per-instance A, A^T A)
enum : int { LP_HP = 0, LP_F = 48, LP_ROWS = 56 };       // pending block of a tree edge: [H | p] contribution, F contribution
enum : int { LX_V = 0, LX_T = 16, LX_ROWS = 96,          // exchange scratch: 16 scalars, 8 rows of 10 (transposes)
             // GPI = 4: + two joint-block-sized dummies for the steps in which a group has no work: one that is only read
             // (zeros, bounds -1 / 1, liMi = identity: every norm term the step produces is exactly 0) and, right behind it,
             // one that takes the step's stores; stride = 8 (mod 16)
             LX_RDUMMY = 96, LX_WDUMMY = 192, LX_ROWS_WIDE = 296 };
// ---- per-CTA constants in shared memory ---------------------------------------------------------------------------
enum : int { CJ_HREFR = 0, CJ_HREF = 48, CJ_HV = 96, CJ_ROWS = 104 };  // per joint: [Href + rho I | -Hv] and Href as 6 rows of 8, Hv (8)
enum : int { CT_AR = 0, CT_ATR = 48, CT_ATA = 96, CT_ROWS = 144 };  // per task: A, A^T, A^T A as 6 rows of 8

enum : int { GC_STRIDE = 24, GC_TOT = 96, GC_ROWS = 120 };  // GPI = 4: the groups' partial norms (4 x 24) and their combination (24)

struct LaneDims {
  int joint0, task0, tmat0, pend0, xch, gc, stride;  // offsets inside an instance record, record stride
  int ctask0, csize, cz0, tab0, cpad;         // constants: first task block, size, a zero joint block and the step table of the wide sweeps, size padded to 128 B
  int xstride;                                // exchange scratch per group
};
__host__ __device__ inline LaneDims lane_dims(const int nb, const int nc, const int npend, const int href_uniform, const int gpi, const int a_per,
                                              const int wide_steps) {
  LaneDims D;
  D.joint0 = LS_ROWS;
  D.task0 = D.joint0 + (gpi > 1 ? LJ_SLOT_WIDE : LJ_ROWS) * nb;
  D.tmat0 = D.task0 + LT_ROWS * nc;          // per-instance task matrices (A, A^T, A^T A as in the constants), if any
  D.pend0 = D.tmat0 + (a_per ? CT_ROWS * nc : 0);
  D.xch = D.pend0 + LP_ROWS * npend;
  D.xstride = gpi > 1 ? LX_ROWS_WIDE : LX_ROWS;
  D.gc = D.xch + D.xstride * gpi;  // (one exchange scratch per group)
  int sz = (D.gc + (gpi > 1 ? GC_ROWS : 0) + 7) & ~7;
  if ((sz & 15) == 0) sz += 8;  // stride = 8 (mod 16) doubles: the records of two neighbouring groups cover different banks
  D.stride = sz;
  D.ctask0 = CJ_ROWS * (href_uniform ? 1 : nb);  // one [Href | -Hv] block when every joint shares the reference (UpdateReference)
  D.csize = D.ctask0 + CT_ROWS * nc;
  D.cz0 = (D.csize + 1) & ~1;                         // (the constants of a step without work)
  D.tab0 = D.cz0 + (gpi > 1 ? CJ_ROWS : 0);           // 4 x int4 = 8 doubles per step of the wide sweeps (backward order, then forward order, one spare step)
  D.cpad = (D.tab0 + (gpi > 1 ? 8 * (wide_steps + 1) : 0) + 15) & ~15;
  return D;
}
inline size_t lane_smem_bytes(const LaneDims& D, const int warps, const int gpi) { return ((size_t)D.cpad + (size_t)warps * (kLaneI / gpi) * D.stride) * sizeof(double); }

struct LaneP {
  const double* src;   // arena the instances live in when the kernel starts
  const int* list;     // optional: queue entry k is slot list[k] of `src` (the survivors of the previous launch)
  const int* n_list;   // device-resident queue length (with `list`), else `n`
  const int* n_back;   // optional: that many more entries at the END of `list` (list[cap - 1], list[cap - 2], ...), queued behind the first *n_list
  int cap;
  int n;
  const int* origin;   // optional: home slot of slot s of `src` (a packed arena); else the home slot is s
  double* home;        // arena the results go to
  int* queue;          // work-queue head (zeroed before the launch)
  int iters;           // fixed mode: iterations per instance
  int fixed;           // stopping disabled (throughput mode)
  int keep_ws;         // the workspace of the last backward pass goes home too (loik_set_keep_workspace)
  const int4* tab;     // GPI = 4: step table of the wide sweeps (ModelC::nsb backward steps, then ModelC::nsf forward steps, 4 entries each)
};

LOIK_DEV void lds6(const double* p, double (&v)[6]) {
  const double2 a = *reinterpret_cast<const double2*>(p), b = *reinterpret_cast<const double2*>(p + 2),
                c = *reinterpret_cast<const double2*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y;
}
LOIK_DEV void sts6(double* p, const double (&v)[6]) {
  *reinterpret_cast<double2*>(p) = make_double2(v[0], v[1]);
  *reinterpret_cast<double2*>(p + 2) = make_double2(v[2], v[3]);
  *reinterpret_cast<double2*>(p + 4) = make_double2(v[4], v[5]);
}
LOIK_DEV double2 lds2(const double* p) { return *reinterpret_cast<const double2*>(p); }

// per-lane partial maxima of the norms whose terms are produced one component per lane
struct LanePart { double dfis, dyis, Av, ptask, dF, Finf, Hrefv, dresv, dvis; };

// ---------------------------------------------------------------------------------------------
// state in / results out (8 lanes of one group, uncoalesced 8 B accesses: once per solve and instance)
// ---------------------------------------------------------------------------------------------
// (l, nl): index of this lane among the nl lanes that share the instance (8, or the whole warp with GPI = 4)
template <int PER, typename F>
LOIK_DEV void lane_copy_rows(const int total, const int l, const int nl, F&& body) {  // body(e, phase, slot): phase 0 = load, 1 = store
  for (int e0 = 0; e0 < total; e0 += nl * PER) {
#pragma unroll
    for (int u = 0; u < PER; ++u) { const int e = e0 + nl * u + l; if (e < total) body(e, 0, u); }
#pragma unroll
    for (int u = 0; u < PER; ++u) { const int e = e0 + nl * u + l; if (e < total) body(e, 1, u); }
  }
}
LOIK_DEV int lane_joff(const ModelC& M, const bool wide, const int j) { return wide ? M.j[j + 1].loff : LJ_ROWS * j; }  // block of joint j + 1
LOIK_DEV void lane_load(const ModelC& M, const LaneDims& D, const double* T, double* I, const int l, const int nl, const bool wide) {
  const Offs& O = M.off;
  double buf[8];
  lane_copy_rows<8>(LJ_COPY * M.nb, l, nl, [&](const int e, const int phase, const int u) {
    const int j = e / LJ_COPY, r = e - LJ_COPY * j;
    if (phase == 0) buf[u] = __ldcs(T + (size_t)(O.joint0 + JR_ROWS * j + r) * 32);
    else I[D.joint0 + lane_joff(M, wide, j) + r] = buf[u];
  });
  lane_copy_rows<8>(LT_ROWS * M.nc, l, nl, [&](const int e, const int phase, const int u) {
    const int k = e / LT_ROWS, r = e - LT_ROWS * k;
    if (phase == 0) buf[u] = __ldcs(T + (size_t)(O.task0 + TR_ROWS * k + r) * 32);
    else I[D.task0 + e] = buf[u];
  });
  if (M.a_per) {  // per-instance task matrices: A (36 rows) and the packed A^T A (21 rows) -> the [6][8] blocks of the constants' layout
    for (int e = l; e < CT_ROWS * M.nc; e += nl) {
      const int k = e / CT_ROWS, o = e - CT_ROWS * k, blk = o / 48, r = (o - 48 * blk) >> 3, c = o & 7;
      const double* Pk = T + (size_t)(O.task0 + TR_ROWS * k) * 32;
      double x = 0.0;
      if (c < 6) {
        if (blk == 0) x = ld(Pk, TR_A + 6 * r + c);
        else if (blk == 1) x = ld(Pk, TR_A + 6 * c + r);
        else x = ld(Pk, TR_ATA + (r < 3 ? (c < 3 ? si(r, c) : 6 + 3 * r + (c - 3)) : (c < 3 ? 6 + 3 * c + (r - 3) : 15 + si(r - 3, c - 3))));
      }
      I[D.tmat0 + e] = x;
    }
  }
  if (l < 7) {
    const int row = l == 0 ? GR_MU : (l == 1 ? GR_BINF : (l == 2 ? GR_CTL : GR_RES + (l - 3)));
    const int dst = l == 0 ? LS_MU : (l == 1 ? LS_BINF : (l == 2 ? LS_CTL : LS_RES + (l - 3)));
    I[dst] = T[(size_t)(O.glob + row) * 32];
  }
}
// what retire_rows (loik_solver.cu) sends home: v, f, F, nu, z, w, T | y, Aty | mu, control, residuals
LOIK_DEV void lane_retire(const ModelC& M, const LaneDims& D, const double* I, double* Th, const int l, const int nl, const bool wide) {
  const Offs& O = M.off;
  for (int e = l; e < JR_JQ * M.nb; e += nl) {
    const int j = e / JR_JQ, r = e - JR_JQ * j;
    Th[(size_t)(O.joint0 + JR_ROWS * j + r) * 32] = I[D.joint0 + lane_joff(M, wide, j) + r];
  }
  for (int e = l; e < TR_B * M.nc; e += nl) {
    const int k = e / TR_B, r = e - TR_B * k;
    Th[(size_t)(O.task0 + TR_ROWS * k + r) * 32] = I[D.task0 + LT_ROWS * k + r];
  }
  if (l < 6) {
    const int row = l == 0 ? GR_MU : (l == 1 ? GR_CTL : GR_RES + (l - 2));
    const int src = l == 0 ? LS_MU : (l == 1 ? LS_CTL : LS_RES + (l - 2));
    Th[(size_t)(O.glob + row) * 32] = I[src];
  }
}
// opt-in: His (21 packed scalars), pis, UDinv, Dinv, r of the last backward pass
LOIK_DEV_CALL void lane_retire_workspace(const ModelC& M, const int joint0, const double* I, double* Th, const int l, const int nl, const bool wide) {
  const Offs& O = M.off;
  for (int j = l; j < M.nb; j += nl) {
    const double* Pj = I + joint0 + lane_joff(M, wide, j);
    double* Pd = Th + (size_t)(O.joint0 + JR_ROWS * j) * 32;
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) {
        if (a <= b) { st(Pd, JR_H + si(a, b), Pj[LJ_HP + 8 * a + b]); st(Pd, JR_H + 15 + si(a, b), Pj[LJ_HP + 8 * (3 + a) + 3 + b]); }
        st(Pd, JR_H + 6 + 3 * a + b, Pj[LJ_HP + 8 * a + 3 + b]);
      }
    for (int c = 0; c < 6; ++c) { st(Pd, JR_P + c, Pj[LJ_HP + 8 * c + 6]); st(Pd, JR_UD + c, Pj[LJ_UD + c]); }
    st(Pd, JR_DINV, Pj[LJ_DINV]);
    st(Pd, JR_R, Pj[LJ_R]);
  }
}

LOIK_DEV void lds_xf(const double* p, double (&R)[9], double (&t)[3]) {  // liMi of a joint from its record
  const double2 a = lds2(p), b = lds2(p + 2), c = lds2(p + 4), d = lds2(p + 6), e = lds2(p + 8), f = lds2(p + 10);
  R[0] = a.x; R[1] = a.y; R[2] = b.x; R[3] = b.y; R[4] = c.x; R[5] = c.y; R[6] = d.x; R[7] = d.y; R[8] = e.x;
  t[0] = e.y; t[1] = f.x; t[2] = f.y;
}

// ---------------------------------------------------------------------------------------------
// Backward sweep: FwdPass1 (hxx:290-338) + BwdPassOptimizedVisitor (hxx:345-354, algo :31-81); cf. sweep_backward.
// Lane c < 6 holds column c of H, lanes 6 and 7 (a duplicate) hold p.
// ---------------------------------------------------------------------------------------------
// In-sweep synchronisation of the lanes that exchange data (one group).  GPI = 1: all four groups run the same joints in
// lock-step, the warp is converged and a full-mask __syncwarp() costs nothing; GPI = 4: the groups sweep different chains
// (different trip counts), so only the group's own lanes may be named.
template <int GPI>
LOIK_DEV void lane_sync(const unsigned gmask) {
  if (GPI == 1) __syncwarp();
  else __syncwarp(gmask);
}

// The sweeps work on the joints lo..hi of the tree (GPI = 1: the whole tree; GPI = 4: one chain, cf. SegC), X = the group's
// exchange scratch.
template <int GPI>
LOIK_DEV void lane_backward(const ModelC& M, const LaneDims& D, const double* CB, double* I, double* X, const int l, const unsigned gmask,
                            const double mu, const double mu_eq, const int lo, const int hi) {
  const double rho = M.rho;
  const int lc = l < 6 ? l : 6;
  const bool isp = l >= 6;
  const double facv = isp ? -rho : 0.0;
  double* XT = X + LX_T;
  const double* XTrow = XT + 10 * lc;  // row of the transposed exchange this lane reads back (row 6: p itself)
  double cc[6];  // contribution carried from child i+1: this lane's column of X* H X*^T, resp. X* p
  bool have_carry = false;
  double* Pj = I + D.joint0 + LJ_ROWS * hi;
  const int cstep = M.href_uniform ? 0 : CJ_ROWS;
  const double* Cj = CB + cstep * hi + lc;
  for (int i = hi; i >= lo; --i) {
    const JointC& J = M.j[i];
    Pj -= LJ_ROWS;
    Cj -= cstep;
    lane_sync<GPI>(gmask);
    double vold[6], col[6];
    lds6(Pj + LJ_V, vold);
    const double w_i = Pj[LJ_W], z_i = Pj[LJ_Z];
    // FwdPass1: H_i = rho I + Href_i (:304-306); p_i = -rho v_prev_i - Hv_i (:310-313)
#pragma unroll
    for (int r = 0; r < 6; ++r) col[r] = fma(facv, vold[r], Cj[CJ_HREFR + 8 * r]);
    if (J.task >= 0) {  // H_c += mu_eq AtA; p_c += Aty - mu_eq Atb (:327-330)
      const double* Ck = M.a_per ? I + D.tmat0 + CT_ROWS * J.task : CB + D.ctask0 + CT_ROWS * J.task;
      const double* Pk = I + D.task0 + LT_ROWS * J.task;
      double aty[6], atb[6];
      lds6(Pk + LT_ATY, aty);
      lds6(Pk + LT_ATB, atb);
#pragma unroll
      for (int r = 0; r < 6; ++r) {
        const double hcol = fma(mu_eq, Ck[CT_ATA + 8 * r + lc], col[r]);
        const double pcol = col[r] + (aty[r] - mu_eq * atb[r]);
        col[r] = isp ? pcol : hcol;
      }
    }
    // children's contributions: His[parent] += SE3actOn(...), pis[parent] += liMi.act(...) (:66,:74)
    for (int n = 0; n < J.npin; ++n) {
      const double* Pp = I + D.pend0 + LP_ROWS * J.pin[n];
#pragma unroll
      for (int r = 0; r < 6; ++r) col[r] += Pp[LP_HP + 8 * r + lc];
    }
    if (have_carry) {
#pragma unroll
      for (int r = 0; r < 6; ++r) col[r] += cc[r];
    }
    // hand H_i, p_i (un-projected) to the forward sweep
#pragma unroll
    for (int r = 0; r < 6; ++r) Pj[LJ_HP + 8 * r + l] = col[r];
    // calc_aba (:60-63): U = H S, Dinv = 1 / (S^T U + R_i) with armature R_i = mu_ineq (:294-295), UDinv = U Dinv;
    // r_i = w_i - mu_ineq z_i (:296) + S^T p_i (:70)
    double U[6], d, Stp, StU;
    const int k = J.sidx;
    if (k >= 0) {
      // aligned joint, S = e_k: U = H(:, k) = row k of the symmetric H, and S^T p = p_k sits next to it in the stored
      // [H | p] block; this lane's own U_c (lane 6: p_k) is entry lc of the same row
      lane_sync<GPI>(gmask);
      const double* row = Pj + LJ_HP + 8 * k;
      lds6(row, U);
      Stp = row[6];
      d = row[l];
      StU = row[k];
    } else {
      // unaligned joint: U_c = S^T H(:, c) (H symmetric), one dot product per lane; lane 6 gets S^T p
      d = St_dot(J, col);
      lane_sync<GPI>(gmask);
      X[LX_V + l] = d;
      lane_sync<GPI>(gmask);
      lds6(X + LX_V, U);
      Stp = X[LX_V + 6];
      StU = St_dot(J, U);
    }
    const double Dinv = 1.0 / (StU + mu);
    const double ri = (w_i - mu * z_i) + Stp;
    double UD[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) UD[c] = U[c] * Dinv;
    Pj[LJ_UD + l] = d * Dinv;
    Pj[LJ_DINV] = Dinv;
    Pj[LJ_R] = ri;
    have_carry = false;
    if (J.parent > 0) {
      // projection: H -= UDinv U^T (calc_aba update_I, :63), p -= UDinv r_i (:71-73)
      const double m = isp ? ri : d;
#pragma unroll
      for (int r = 0; r < 6; ++r) col[r] -= UD[r] * m;
      double R[9], t[3], y[6], row[6];
      lds_xf(Pj + LJ_XF, R, t);
      act_force(R, t, col, y);  // Y = X* H, column by column
#pragma unroll
      for (int r = 0; r < 6; ++r) XT[10 * r + l] = y[r];
      if (isp) sts6(XT + 60, col);  // (the p lanes: their single transform is the second one)
      lane_sync<GPI>(gmask);
      lds6(XTrow, row);
      act_force(R, t, row, cc);  // row c of Y X*^T = X* (row c of Y): column c of X* H X*^T (SE3actOn, :66); lane 6: liMi.act(p) (:74)
      if (J.carry) {
        have_carry = true;
      } else {
        double* Pp = I + D.pend0 + LP_ROWS * J.pout;
#pragma unroll
        for (int r = 0; r < 6; ++r) Pp[LP_HP + 8 * r + l] = cc[r];
      }
    }
  }
}

// select x[l] for a (lane-dependent or uniform) l in 0..5 (two levels of selects)
LOIK_DEV double pick6(const double (&x)[6], const int l) {
  const bool odd = l & 1;
  const double a = odd ? x[1] : x[0], b = odd ? x[3] : x[2], c = odd ? x[5] : x[4];
  return l < 2 ? a : (l < 4 ? b : c);
}

// ---------------------------------------------------------------------------------------------
// Forward sweep: FwdPass2OptimizedVisitor (hxx:361-377) + BoxProj (:384-397) + DualUpdate (:404-461) +
// ComputePrimalResiduals (:494-503); cf. sweep_forward.  The 6-vectors are computed by every lane (they are the
// chain), f = H v + p / A v / A^T y one component per lane.
// ---------------------------------------------------------------------------------------------
template <int GPI>
LOIK_DEV void lane_forward(const ModelC& M, const LaneDims& D, const double* CB, double* I, double* X, const int l, const unsigned gmask,
                           const double mu, const double mu_eq, Carry& cy, LanePart& pt, const int lo, const int hi) {
  const double inv_mu = 1.0 / mu;
  const int lc = l < 6 ? l : 5;  // lanes 6, 7 duplicate lane 5
  double v[6] = {0, 0, 0, 0, 0, 0};  // v of joint i-1 on entry of step i
  double* Pj = I + D.joint0 + LJ_ROWS * (lo - 2);
  for (int i = lo; i <= hi; ++i) {
    const JointC& J = M.j[i];
    Pj += LJ_ROWS;
    lane_sync<GPI>(gmask);
    double UD[6], R[9], t[3];
    const double vold_l = Pj[LJ_V + lc];
    const double2 dr = lds2(Pj + LJ_DINV);  // (Dinv, r)
    const double2 nz = lds2(Pj + LJ_NU);    // (nu, z) of the previous iterate
    const double w_old = Pj[LJ_W];
    double lb = J.lb, ub = J.ub;
    if (M.bounds_per_instance) { const double2 b2 = lds2(Pj + LJ_LB); lb = b2.x; ub = b2.y; }
    lds6(Pj + LJ_UD, UD);
    lds_xf(Pj + LJ_XF, R, t);
    if (i == lo || J.parent != i - 1) {  // not the joint this group has just swept: the universe (v = 0) or a joint swept earlier (its new v)
      if (J.parent == 0) {
#pragma unroll
        for (int c = 0; c < 6; ++c) v[c] = 0.0;
      } else {
        lds6(I + D.joint0 + LJ_ROWS * (J.parent - 1) + LJ_V, v);
      }
    }
    // f_i = H_i v_i + p_i needs row lc of H (= column lc), p_lc, and the old f_lc
    double hc[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) hc[c] = Pj[LJ_HP + 8 * c + lc];
    const double p_l = Pj[LJ_HP + 8 * lc + 6];
    const double fold_l = Pj[LJ_F + lc];
    {
      double vp[6];
      actinv_motion(R, t, v, vp);  // vi_parent (:125)
#pragma unroll
      for (int c = 0; c < 6; ++c) v[c] = vp[c];
    }
    double acc = 0.0;
#pragma unroll
    for (int c = 0; c < 6; ++c) acc += UD[c] * v[c];
    const double nu = -acc - dr.x * dr.y;  // nu_i = -UDinv^T vp - Dinv r_i (:127)
    cy.nu_inf = amax(cy.nu_inf, nu);       // (:129-131)
    S_axpy(J, nu, v);                      // v_i = vp + S nu_i (:133-134)
    cy.dnu_inf = amax(cy.dnu_inf, nu - nz.x);  // (:375)
    const double z = dmin(ub, dmax(lb, nu + inv_mu * w_old));  // BoxProj (:388)
    cy.dz_inf = amax(cy.dz_inf, z - nz.y);
    const double rp = nu - z;
    cy.pres_slack = amax(cy.pres_slack, rp);
    const double dw = mu * rp;  // (:454-458)
    cy.dw_inf = amax(cy.dw_inf, dw);
    cy.ubdw_p += ub * dmax(dw, 0.0);  // CheckFeasibility's dot products (:588,590)
    cy.lbdw_m += lb * dmin(dw, 0.0);
    const double f_l = hc[0] * v[0] + hc[1] * v[1] + hc[2] * v[2] + hc[3] * v[3] + hc[4] * v[4] + hc[5] * v[5] + p_l;  // (:139-140)
    pt.dfis = amax(pt.dfis, f_l - fold_l);  // (:137-146)
    lane_sync<GPI>(gmask);  // every lane has read this joint's previous iterate
    if (l == 0) sts6(Pj + LJ_V, v);
    Pj[LJ_F + lc] = f_l;
    lane_sync<GPI>(gmask);
    pt.dvis = amax(pt.dvis, Pj[LJ_V + lc] - vold_l);  // delta_vis_inf_norm (:156-158), one component per lane
    *reinterpret_cast<double2*>(Pj + LJ_NU) = make_double2(nu, z);
    Pj[LJ_W] = w_old + dw;
    if (J.task >= 0) {  // DualUpdate for the task on this joint (:410-451)
      const double* Ck = M.a_per ? I + D.tmat0 + CT_ROWS * J.task : CB + D.ctask0 + CT_ROWS * J.task;
      double* Pk = I + D.task0 + LT_ROWS * J.task;
      const double Av = Ck[CT_ATR + lc] * v[0] + Ck[CT_ATR + 8 + lc] * v[1] + Ck[CT_ATR + 16 + lc] * v[2] + Ck[CT_ATR + 24 + lc] * v[3] +
                        Ck[CT_ATR + 32 + lc] * v[4] + Ck[CT_ATR + 40 + lc] * v[5];
      const double e = Av - Pk[LT_B + lc];  // Av_minus_b (:416)
      const double dy = mu_eq * e;          // delta_yis (:419)
      const double y_l = Pk[LT_Y + lc] + dy;
      pt.dyis = amax(pt.dyis, dy);
      pt.Av = amax(pt.Av, Av);
      pt.ptask = amax(pt.ptask, e);
      X[LX_V + l] = dy;
      X[LX_V + 8 + l] = y_l;
      lane_sync<GPI>(gmask);
      double dys[6], ys[6], bk[6];
      lds6(X + LX_V, dys);
      lds6(X + LX_V + 8, ys);
      lds6(Pk + LT_B, bk);
      double plus = 0.0, minus = 0.0;
#pragma unroll
      for (int a = 0; a < 6; ++a) {
        plus += bk[a] * dmax(dys[a], 0.0);
        minus += bk[a] * dmin(dys[a], 0.0);
      }
      cy.bTdy_p += plus;
      cy.bTdy_m += minus;
      // Aty = A^T y (:425)
      const double aty = Ck[CT_AR + lc] * ys[0] + Ck[CT_AR + 8 + lc] * ys[1] + Ck[CT_AR + 16 + lc] * ys[2] + Ck[CT_AR + 24 + lc] * ys[3] +
                         Ck[CT_AR + 32 + lc] * ys[4] + Ck[CT_AR + 40 + lc] * ys[5];
      Pk[LT_Y + lc] = y_l;
      Pk[LT_ATY + lc] = aty;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Residual sweep: BwdPass2OptimizedVisitor (hxx:468-487, algo :185-241) + ComputeDualResiduals (:510-522); cf.
// sweep_residual.  F = fis_diff_plus_Aty one component per lane, liMi.act(f_i) by every lane.
// ---------------------------------------------------------------------------------------------
template <int GPI>
LOIK_DEV void lane_residual(const ModelC& M, const LaneDims& D, const double* CB, double* I, const int l, const unsigned gmask, Resid& rs,
                            LanePart& pt, const int lo, const int hi) {
  const int lc = l < 6 ? l : 5;
  double cF = 0.0;
  bool have_carry = false;
  double* Pj = I + D.joint0 + LJ_ROWS * hi;
  const int cstep = M.href_uniform ? 0 : CJ_ROWS;
  const double* Cj = CB + cstep * hi + lc;
  for (int i = hi; i >= lo; --i) {
    const JointC& J = M.j[i];
    Pj -= LJ_ROWS;
    Cj -= cstep;
    lane_sync<GPI>(gmask);
    double f[6], v[6];
    lds6(Pj + LJ_F, f);
    lds6(Pj + LJ_V, v);
    const double2 wt = lds2(Pj + LJ_W);  // (w_i new, T old)
    const double f_l = Pj[LJ_F + lc], Fold = Pj[LJ_FD + lc], Hv_l = Cj[CJ_HV];
    double hr[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) hr[c] = Cj[CJ_HREF + 8 * c];
    const int k = J.sidx;
    const double Stf = k >= 0 ? Pj[LJ_F + k] : St_dot(J, f);  // S^T f_i
    double F = 0.0;  // (:370)
    if (J.task >= 0) F = I[D.task0 + LT_ROWS * J.task + LT_ATY + lc];  // (:438-439)
    for (int n = 0; n < J.npin; ++n) F += I[D.pend0 + LP_ROWS * J.pin[n] + LP_F + lc];
    if (have_carry) F += cF;
    F += -f_l;  // (:210)
    // Href_v (fwd pass 2, :149-153), recomputed from v_i
    const double Hrv = hr[0] * v[0] + hr[1] * v[1] + hr[2] * v[2] + hr[3] * v[3] + hr[4] * v[4] + hr[5] * v[5];
    const double rd = Hrv - Hv_l + F;       // (:228)
    pt.dF = amax(pt.dF, F - Fold);          // (:215-220)
    pt.Finf = amax(pt.Finf, F);             // (:223-225)
    pt.Hrefv = amax(pt.Hrefv, Hrv);
    pt.dresv = amax(pt.dresv, rd);
    const double Tn = Stf + wt.x;           // Stf_plus_w (:231-236) and its delta (:471,:482-483)
    rs.T_inf = amax(rs.T_inf, Tn);
    rs.dT_inf = amax(rs.dT_inf, Tn - wt.y);
    lane_sync<GPI>(gmask);
    Pj[LJ_FD + lc] = F;
    Pj[LJ_T] = Tn;
    have_carry = false;
    if (J.parent > 0) {  // fis_diff_plus_Aty[parent] += liMi.act(f_i) (:212)
      double R[9], t[3], c6[6];
      lds_xf(Pj + LJ_XF, R, t);
      act_force(R, t, f, c6);
      cF = pick6(c6, lc);
      if (J.carry) have_carry = true;
      else I[D.pend0 + LP_ROWS * J.pout + LP_F + lc] = cF;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// GPI = 4: the same three sweeps with the four groups of a warp on different chains of ONE instance.  SIMT runs one
// instruction stream per warp, so the groups only work in parallel if they execute the SAME instructions: every sweep is
// one flat loop over the steps of a host-built table (build_wide_table, loik_solver.cu; WideStep), four entries per step:
// what group g does in that step -- the joint, whether the step is real (`valid`), whether it starts a chain, where its
// result goes.  A group without work in a step runs the same arithmetic on an all-zero joint block and all-zero constants
// (LX_RDUMMY, LaneDims::cz0: every value it produces, hence every norm term, is exactly 0) with its stores redirected to
// a scratch block, so the hot path has neither a branch nor a select that depends on `valid`; the per-joint branches that
// remain (task on the joint, contributions of several children, unaligned axis) are rare.  The kernel copies the table
// to shared memory once per CTA and patches in what depends on the problem (the task on a joint) and on the record
// layout (the dummy blocks).  All exchanges between lanes are synchronised at the top level of the loop (full-mask __syncwarp(), free in
// converged code).  The level order of the table (children before parents towards the root, parents first on the way
// out) makes the pending blocks and the parents' v rows valid when a step reads them: a step's stores are separated
// from the next step's loads by the __syncwarp() at its top.
// ---------------------------------------------------------------------------------------------
struct WideStep { int joint_parent, flags, loff_pout, sidx_ploff; };  // joint | parent << 16; WF_* | (task + 1) << 8 (in the kernel's copy); block offset | pout << 16; (aligned axis index or -1) & 0xffff | parent's block offset << 16
enum : int { WF_VALID = 1, WF_FIRST = 2, WF_GIVE = 4, WF_ROOT = 8, WF_PINS = 16 };  // FIRST: first step of a chain; GIVE: result goes to a pending block

LOIK_DEV void wide_backward(const ModelC& M, const LaneDims& D, const double* CB, double* I, double* X, const int l, const int g,
                            const double mu, const double mu_eq, const int4* tab, const int nsteps) {
  const double rho = M.rho;
  const int lc = l < 6 ? l : 6;
  const bool isp = l >= 6;
  const double facv = isp ? -rho : 0.0;
  double* XT = X + LX_T;
  double* XD = X + LX_WDUMMY;
  const double* CZ = CB + D.cz0 + lc;
  const double* XTrow = XT + 10 * lc;
  const int cstep = M.href_uniform ? 0 : CJ_ROWS;
  double cc[6] = {0, 0, 0, 0, 0, 0};
  int4 nxt = tab[g];
  for (int s = 0; s < nsteps; ++s) {
    const int4 ws = nxt;
    nxt = tab[4 * (s + 1) + g];  // (the table has a spare step behind the last one)
    const int i = ws.x & 0xffff;
    const bool valid = ws.y & WF_VALID, carry_in = valid && !(ws.y & WF_FIRST);
    double* Pj = I + D.joint0 + (ws.z & 0xffff);
    double* Pw = Pj + (valid ? 0 : LX_WDUMMY - LX_RDUMMY);
    const double* Cj = valid ? CB + cstep * (i - 1) + lc : CZ;
    __syncwarp();
    double vold[6], col[6];
    lds6(Pj + LJ_V, vold);
    const double w_i = Pj[LJ_W], z_i = Pj[LJ_Z];
#pragma unroll
    for (int r = 0; r < 6; ++r) col[r] = fma(facv, vold[r], Cj[CJ_HREFR + 8 * r]);
    const int task = ((ws.y >> 8) & 0xff) - 1;
    if (task >= 0) {
      const double* Ck = M.a_per ? I + D.tmat0 + CT_ROWS * task : CB + D.ctask0 + CT_ROWS * task;
      const double* Pk = I + D.task0 + LT_ROWS * task;
      double aty[6], atb[6];
      lds6(Pk + LT_ATY, aty);
      lds6(Pk + LT_ATB, atb);
#pragma unroll
      for (int r = 0; r < 6; ++r) {
        const double hcol = fma(mu_eq, Ck[CT_ATA + 8 * r + lc], col[r]);
        const double pcol = col[r] + (aty[r] - mu_eq * atb[r]);
        col[r] = isp ? pcol : hcol;
      }
    }
    if (ws.y & WF_PINS) {
      const JointC& J = M.j[i];
      for (int n = 0; n < J.npin; ++n) {
        const double* Pp = I + D.pend0 + LP_ROWS * J.pin[n];
#pragma unroll
        for (int r = 0; r < 6; ++r) col[r] += Pp[LP_HP + 8 * r + lc];
      }
    }
#pragma unroll
    for (int r = 0; r < 6; ++r) col[r] += carry_in ? cc[r] : 0.0;  // the child swept one step earlier (same chain)
    const int k = (short)(ws.w & 0xffff);
    double d_un = 0.0;
    if (k < 0) d_un = St_dot(M.j[i], col);  // unaligned joint: U_c = S^T H(:, c), lane 6: S^T p
#pragma unroll
    for (int r = 0; r < 6; ++r) Pw[LJ_HP + 8 * r + l] = col[r];
    X[LX_V + l] = d_un;
    __syncwarp();
    double U[6], d, Stp, StU;
    if (k >= 0) {
      // (a group without work reads the block of joint 1 while its owner may be writing it: finite values, results unused)
      const double* row = Pj + LJ_HP + 8 * k;
      lds6(row, U);
      Stp = row[6];
      d = row[l];
      StU = row[k];
    } else {
      lds6(X + LX_V, U);
      Stp = X[LX_V + 6];
      d = d_un;
      StU = St_dot(M.j[i], U);
    }
    const double Dinv = 1.0 / (StU + mu);
    const double ri = (w_i - mu * z_i) + Stp;
    double UD[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) UD[c] = U[c] * Dinv;
    Pw[LJ_UD + l] = d * Dinv;
    Pw[LJ_DINV] = Dinv;
    Pw[LJ_R] = ri;
    // projection + transformation to the parent frame, by every group (a root joint's result is simply not handed on)
    const double m = isp ? ri : d;
#pragma unroll
    for (int r = 0; r < 6; ++r) col[r] -= UD[r] * m;
    double R[9], tr[3], y[6], row[6];
    lds_xf(Pj + LJ_XF, R, tr);
    act_force(R, tr, col, y);
#pragma unroll
    for (int r = 0; r < 6; ++r) XT[10 * r + l] = y[r];
    if (isp) sts6(XT + 60, col);
    __syncwarp();
    lds6(XTrow, row);
    act_force(R, tr, row, cc);
    double* Pp = (ws.y & WF_GIVE) ? I + D.pend0 + LP_ROWS * (ws.z >> 16) : XD;
#pragma unroll
    for (int r = 0; r < 6; ++r) Pp[LP_HP + 8 * r + l] = cc[r];
  }
}

LOIK_DEV void wide_forward(const ModelC& M, const LaneDims& D, const double* CB, double* I, double* X, const int l, const int g,
                           const unsigned gmask, const double mu, const double mu_eq, Carry& cy, LanePart& pt, const int4* tab, const int nsteps) {
  const double inv_mu = 1.0 / mu;
  const int lc = l < 6 ? l : 5;
  double v[6] = {0, 0, 0, 0, 0, 0};
  int4 nxt = tab[g];
  for (int s = 0; s < nsteps; ++s) {
    const int4 ws = nxt;
    nxt = tab[4 * (s + 1) + g];
    const int i = ws.x & 0xffff, parent = ws.x >> 16;
    const bool valid = ws.y & WF_VALID;
    double* Pj = I + D.joint0 + (ws.z & 0xffff);
    double* Pw = Pj + (valid ? 0 : LX_WDUMMY - LX_RDUMMY);
    __syncwarp();
    double UD[6], R[9], tr[3];
    const double vold_l = Pj[LJ_V + lc];
    const double2 dr = lds2(Pj + LJ_DINV);
    const double2 nz = lds2(Pj + LJ_NU);
    const double w_old = Pj[LJ_W];
    const double2 b2 = lds2(Pj + LJ_LB);  // (batch-shared bounds are written into the record when the instance is loaded)
    const double lb = b2.x, ub = b2.y;
    lds6(Pj + LJ_UD, UD);
    lds_xf(Pj + LJ_XF, R, tr);
    if (ws.y & WF_FIRST) {  // first joint of a chain: the parent's v comes from its record (zero for a root joint)
      double vq[6];
      lds6(I + D.joint0 + (ws.w >> 16) + LJ_V, vq);
#pragma unroll
      for (int c = 0; c < 6; ++c) v[c] = parent > 0 ? vq[c] : 0.0;
    }
    double hc[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) hc[c] = Pj[LJ_HP + 8 * c + lc];
    const double p_l = Pj[LJ_HP + 8 * lc + 6];
    const double fold_l = Pj[LJ_F + lc];
    {
      double vp[6];
      actinv_motion(R, tr, v, vp);
#pragma unroll
      for (int c = 0; c < 6; ++c) v[c] = vp[c];
    }
    double acc = 0.0;
#pragma unroll
    for (int c = 0; c < 6; ++c) acc += UD[c] * v[c];
    const double nu = -acc - dr.x * dr.y;
    const int k = (short)(ws.w & 0xffff);
    if (k >= 0) {
#pragma unroll
      for (int c = 0; c < 6; ++c) v[c] += c == k ? nu : 0.0;  // v_i = vp + S nu_i (:133-134), aligned axis
    } else {
      S_axpy(M.j[i], nu, v);
    }
    const double z = dmin(ub, dmax(lb, nu + inv_mu * w_old));
    const double rp = nu - z;
    const double dw = mu * rp;
    const double f_l = hc[0] * v[0] + hc[1] * v[1] + hc[2] * v[2] + hc[3] * v[3] + hc[4] * v[4] + hc[5] * v[5] + p_l;
    {  // (a step without work: nu = z = dw = f = 0 exactly)
      cy.nu_inf = amax(cy.nu_inf, nu);
      cy.dnu_inf = amax(cy.dnu_inf, nu - nz.x);
      cy.dz_inf = amax(cy.dz_inf, z - nz.y);
      cy.pres_slack = amax(cy.pres_slack, rp);
      cy.dw_inf = amax(cy.dw_inf, dw);
      cy.ubdw_p += ub * dmax(dw, 0.0);
      cy.lbdw_m += lb * dmin(dw, 0.0);
      pt.dfis = amax(pt.dfis, f_l - fold_l);
    }
    __syncwarp();
    if (l == 0) sts6(Pw + LJ_V, v);
    Pw[LJ_F + lc] = f_l;
    *reinterpret_cast<double2*>(Pw + LJ_NU) = make_double2(nu, z);
    Pw[LJ_W] = w_old + dw;
    __syncwarp();
    pt.dvis = amax(pt.dvis, Pj[LJ_V + lc] - vold_l);  // (a step without work reads the zero block again)
    const int task = ((ws.y >> 8) & 0xff) - 1;
    if (task >= 0) {  // DualUpdate for the task on this joint (:410-451); only the group that owns the joint is here
      const double* Ck = M.a_per ? I + D.tmat0 + CT_ROWS * task : CB + D.ctask0 + CT_ROWS * task;
      double* Pk = I + D.task0 + LT_ROWS * task;
      const double Av = Ck[CT_ATR + lc] * v[0] + Ck[CT_ATR + 8 + lc] * v[1] + Ck[CT_ATR + 16 + lc] * v[2] + Ck[CT_ATR + 24 + lc] * v[3] +
                        Ck[CT_ATR + 32 + lc] * v[4] + Ck[CT_ATR + 40 + lc] * v[5];
      const double e = Av - Pk[LT_B + lc];
      const double dy = mu_eq * e;
      const double y_l = Pk[LT_Y + lc] + dy;
      pt.dyis = amax(pt.dyis, dy);
      pt.Av = amax(pt.Av, Av);
      pt.ptask = amax(pt.ptask, e);
      X[LX_V + l] = dy;
      X[LX_V + 8 + l] = y_l;
      __syncwarp(gmask);
      double dys[6], ys[6], bk[6];
      lds6(X + LX_V, dys);
      lds6(X + LX_V + 8, ys);
      lds6(Pk + LT_B, bk);
      double plus = 0.0, minus = 0.0;
#pragma unroll
      for (int a = 0; a < 6; ++a) {
        plus += bk[a] * dmax(dys[a], 0.0);
        minus += bk[a] * dmin(dys[a], 0.0);
      }
      cy.bTdy_p += plus;
      cy.bTdy_m += minus;
      const double aty = Ck[CT_AR + lc] * ys[0] + Ck[CT_AR + 8 + lc] * ys[1] + Ck[CT_AR + 16 + lc] * ys[2] + Ck[CT_AR + 24 + lc] * ys[3] +
                         Ck[CT_AR + 32 + lc] * ys[4] + Ck[CT_AR + 40 + lc] * ys[5];
      __syncwarp(gmask);
      Pk[LT_Y + lc] = y_l;
      Pk[LT_ATY + lc] = aty;
    }
  }
}

LOIK_DEV void wide_residual(const ModelC& M, const LaneDims& D, const double* CB, double* I, double* X, const int l, const int g, Resid& rs,
                            LanePart& pt, const int4* tab, const int nsteps) {
  const int lc = l < 6 ? l : 5;
  const int cstep = M.href_uniform ? 0 : CJ_ROWS;
  double* XD = X + LX_WDUMMY;
  const double* CZ = CB + D.cz0 + lc;
  double cF = 0.0;
  int4 nxt = tab[g];
  for (int s = 0; s < nsteps; ++s) {
    const int4 ws = nxt;
    nxt = tab[4 * (s + 1) + g];
    const int i = ws.x & 0xffff;
    const bool valid = ws.y & WF_VALID, carry_in = valid && !(ws.y & WF_FIRST);
    double* Pj = I + D.joint0 + (ws.z & 0xffff);
    double* Pw = Pj + (valid ? 0 : LX_WDUMMY - LX_RDUMMY);
    const double* Cj = valid ? CB + cstep * (i - 1) + lc : CZ;
    __syncwarp();
    double f[6], v[6];
    lds6(Pj + LJ_F, f);
    lds6(Pj + LJ_V, v);
    const double2 wt = lds2(Pj + LJ_W);
    const double f_l = Pj[LJ_F + lc], Fold = Pj[LJ_FD + lc], Hv_l = Cj[CJ_HV];
    double hr[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) hr[c] = Cj[CJ_HREF + 8 * c];
    const int k = (short)(ws.w & 0xffff);
    const double Stf = k >= 0 ? Pj[LJ_F + k] : St_dot(M.j[i], f);
    double F = 0.0;
    const int task = ((ws.y >> 8) & 0xff) - 1;
    if (task >= 0) F = I[D.task0 + LT_ROWS * task + LT_ATY + lc];
    if (ws.y & WF_PINS) {
      const JointC& J = M.j[i];
      for (int n = 0; n < J.npin; ++n) F += I[D.pend0 + LP_ROWS * J.pin[n] + LP_F + lc];
    }
    F += carry_in ? cF : 0.0;
    F += -f_l;
    const double Hrv = hr[0] * v[0] + hr[1] * v[1] + hr[2] * v[2] + hr[3] * v[3] + hr[4] * v[4] + hr[5] * v[5];
    const double rd = Hrv - Hv_l + F;
    const double Tn = Stf + wt.x;
    pt.dF = amax(pt.dF, F - Fold);  // (a step without work: f = v = F = T = 0 exactly)
    pt.Finf = amax(pt.Finf, F);
    pt.Hrefv = amax(pt.Hrefv, Hrv);
    pt.dresv = amax(pt.dresv, rd);
    rs.T_inf = amax(rs.T_inf, Tn);
    rs.dT_inf = amax(rs.dT_inf, Tn - wt.y);
    __syncwarp();
    Pw[LJ_FD + lc] = F;
    Pw[LJ_T] = Tn;
    double R[9], tr[3], c6[6];
    lds_xf(Pj + LJ_XF, R, tr);
    act_force(R, tr, f, c6);
    cF = pick6(c6, lc);
    double* Pp = (ws.y & WF_GIVE) ? I + D.pend0 + LP_ROWS * (ws.z >> 16) : XD;
    Pp[LP_F + lc] = cF;
  }
}

// ---------------------------------------------------------------------------------------------
// The kernel.  blockDim = 32 W; dynamic shared memory = constants + (4 / GPI) W instance records (lane_smem_bytes).
// ---------------------------------------------------------------------------------------------
template <int GPI>
__global__ void __launch_bounds__(256) k_iterate_lane(const __grid_constant__ ModelC c_model, const LaneP P) {
  extern __shared__ __align__(16) double lsm[];
  const ModelC& M = c_model;
  const LaneDims D = lane_dims(M.nb, M.nc, M.npend, M.href_uniform, GPI, M.a_per, M.nsb + M.nsf);
  double* CB = lsm;
  const int4* tabB = reinterpret_cast<const int4*>(lsm + D.tab0);
  const int4* tabF = tabB + 4 * M.nsb;
  if (GPI > 1) {
    for (int e = threadIdx.x; e < 4 * (M.nsb + M.nsf + 1); e += blockDim.x) {
      int4 ws = e < 4 * (M.nsb + M.nsf) ? P.tab[e] : make_int4(1, 0, 0, 0);
      if (ws.y & WF_VALID) ws.y |= (M.j[ws.x & 0xffff].task + 1) << 8;                                  // the task on the joint (SolveInit)
      else ws.z = D.xch + (e & 3) * LX_ROWS_WIDE + LX_RDUMMY - D.joint0;                                // this group's zero block
      reinterpret_cast<int4*>(lsm + D.tab0)[e] = ws;
    }
    for (int e = threadIdx.x; e < CJ_ROWS; e += blockDim.x) lsm[D.cz0 + e] = 0.0;
  }
  for (int e = threadIdx.x; e < D.csize; e += blockDim.x) {
    double x = 0.0;
    if (e < D.ctask0) {
      const int j = e / CJ_ROWS, o = e - CJ_ROWS * j;
      const JointC& J = M.j[j + 1];
    const HrefC& Hr = M.href[J.href];
      if (o < CJ_HV) {
        const int r = (o % 48) >> 3, c = o & 7;
        if (c < 6) x = Hel(Hr.A, Hr.B, Hr.D, r, c) + ((o < CJ_HREF && r == c) ? M.rho : 0.0);  // H_i = rho I + Href_i (:304-306)
        else if (c == 6 && o < CJ_HREF) x = -Hr.Hv[r];
      } else if (o - CJ_HV < 6) {
        x = Hr.Hv[o - CJ_HV];
      }
    } else {
      const int k = (e - D.ctask0) / CT_ROWS, o = (e - D.ctask0) - CT_ROWS * k;
      const TaskC& K = M.t[k];
      const int blk = o / 48, r = (o - 48 * blk) >> 3, c = o & 7;
      if (c < 6) x = blk == 0 ? K.A[6 * r + c] : (blk == 1 ? K.A[6 * c + r] : Hel(K.AtA_A, K.AtA_B, K.AtA_D, r, c));
    }
    CB[e] = x;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, l = lane & 7, g = lane >> 3, w = threadIdx.x >> 5;
  const unsigned gmask = 0xffu << (8 * g);
  // the lanes that share one instance: an 8-lane group, or the whole warp
  const int wl = GPI == 1 ? l : lane, nl = GPI == 1 ? 8 : 32;
  const unsigned imask = GPI == 1 ? gmask : 0xffffffffu;
  double* I = lsm + D.cpad + (size_t)(GPI == 1 ? w * kLaneI + g : w) * D.stride;
  double* X = I + D.xch + (GPI == 1 ? 0 : g * LX_ROWS_WIDE);
  if (GPI > 1) {  // the dummy blocks of this group (never written again: lane_load only touches joints, tasks and globals)
    for (int e = l; e < LX_ROWS_WIDE - LX_RDUMMY; e += 8) X[LX_RDUMMY + e] = 0.0;
    __syncwarp();
    if (l == 0) {
      double* Z = X + LX_RDUMMY;
      Z[LJ_LB] = -1.0; Z[LJ_UB] = 1.0;
      Z[LJ_XF] = 1.0; Z[LJ_XF + 4] = 1.0; Z[LJ_XF + 8] = 1.0;
    }
    __syncwarp();
  }
  const int nfront = P.list ? *P.n_list : P.n;
  const int nback = (P.list && P.n_back) ? *P.n_back : 0;
  const int limit = nfront + nback;
  int home_slot = -1;  // >= 0: these lanes hold an instance
  int status = ST_CONVERGED, it = 0, left = 0;
  double mu = 1.0;
  bool exhausted = false;
  for (;;) {
    if (home_slot < 0 && !exhausted) {  // pull the next instance from the queue
      __syncwarp(imask);  // the record is free: every lane is done with the previous instance
      int s = -1, k = 0;
      if (wl == 0) k = atomicAdd(P.queue, 1);
      k = __shfl_sync(imask, k, GPI == 1 ? 8 * g : 0);
      // (an ordered hand-over, LaneP::n_back: the far-from-done instances at the front of the list come first, then the rest from its end)
      if (k < limit) s = P.list ? P.list[k < nfront ? k : P.cap - 1 - (k - nfront)] : k;
      else exhausted = true;
      if (s >= 0) {
        const double* T = P.src + ((size_t)(s >> 5) * M.off.rows) * 32 + (s & 31);
        lane_load(M, D, T, I, wl, nl, GPI > 1);
        __syncwarp(imask);
        for (int j = wl; j < M.nb; j += nl) {  // liMi of every joint (FwdPassInit, hxx:263-264), once per instance
          double* Pj = I + D.joint0 + lane_joff(M, GPI > 1, j);
          double R[9], t[3];
          make_xf(M.j[j + 1], Pj[LJ_SQ], Pj[LJ_CQ], R, t);
#pragma unroll
          for (int c = 0; c < 9; ++c) Pj[LJ_XF + c] = R[c];
#pragma unroll
          for (int c = 0; c < 3; ++c) Pj[LJ_XF + 9 + c] = t[c];
          if (GPI > 1 && !M.bounds_per_instance) { Pj[LJ_LB] = M.j[j + 1].lb; Pj[LJ_UB] = M.j[j + 1].ub; }  // (the wide sweeps read the bounds from the record)
        }
        __syncwarp(imask);
        const int2 ctl = *reinterpret_cast<const int2*>(I + LS_CTL);
        status = ctl.x; it = ctl.y;
        mu = I[LS_MU];
        left = P.iters;
        if (status < ST_CONVERGED) home_slot = P.origin ? P.origin[s] : s;
      }
    }
    __syncwarp();
    const bool act = home_slot >= 0;
    if (!__any_sync(0xffffffffu, act)) break;
    // ---- one ADMM iteration of the instance(s) of this warp
    const double mu_eq = M.mu_scale * mu;
    Carry cy;
    Resid rs;
    LanePart pt = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    zero(cy);
    zero(rs);
    if (GPI == 1) {
      lane_backward<1>(M, D, CB, I, X, l, gmask, mu, mu_eq, 1, M.nb);
      lane_forward<1>(M, D, CB, I, X, l, gmask, mu, mu_eq, cy, pt, 1, M.nb);
      lane_residual<1>(M, D, CB, I, l, gmask, rs, pt, 1, M.nb);
    } else {
      // the four groups sweep different chains of the instance, step by step through the host-built table (wide_backward)
      wide_backward(M, D, CB, I, X, l, g, mu, mu_eq, tabB, M.nsb);
      wide_forward(M, D, CB, I, X, l, g, gmask, mu, mu_eq, cy, pt, tabF, M.nsf);
      wide_residual(M, D, CB, I, X, l, g, rs, pt, tabB, M.nsb);
      __syncwarp();
    }
    {  // combine the per-lane partial maxima of the group: rows = lanes, then one column per lane
      double* XT = X + LX_T;
      lane_sync<GPI>(gmask);
      double* row = XT + 10 * l;
      *reinterpret_cast<double2*>(row) = make_double2(pt.dfis, pt.dyis);
      *reinterpret_cast<double2*>(row + 2) = make_double2(pt.Av, pt.ptask);
      *reinterpret_cast<double2*>(row + 4) = make_double2(pt.dF, pt.Finf);
      *reinterpret_cast<double2*>(row + 6) = make_double2(pt.Hrefv, pt.dresv);
      row[8] = pt.dvis;
      lane_sync<GPI>(gmask);
      const double a0 = XT[l], a1 = XT[10 + l], a2 = XT[20 + l], a3 = XT[30 + l], a4 = XT[40 + l], a5 = XT[50 + l];
      X[LX_V + 8 + l] = dmax(dmax(dmax(a0, a1), dmax(a2, a3)), dmax(a4, a5));
      if (l == 0) X[LX_V + 7] = dmax(dmax(dmax(XT[8], XT[18]), dmax(XT[28], XT[38])), dmax(XT[48], XT[58]));
      lane_sync<GPI>(gmask);
      cy.dvis_inf = X[LX_V + 7];
      const double2 t0 = lds2(X + LX_V + 8), t1 = lds2(X + LX_V + 10), t2 = lds2(X + LX_V + 12), t3 = lds2(X + LX_V + 14);
      cy.dfis_inf = t0.x; cy.dyis_inf = t0.y; cy.Av_inf = t1.x; cy.pres_task = t1.y;
      rs.dF_inf = t2.x; rs.F_inf = t2.y; rs.Hrefv_inf = t3.x; rs.dres_v = t3.y;
    }
    if (GPI > 1) {
      // combine the groups' norms / sums in a fixed group order (as k_iterate_seg combines its warps): entries 8..11 of
      // Carry are sums, everything else an inf-norm; afterwards every lane holds the instance's values
      static_assert(sizeof(Carry) == kCarryRows * sizeof(double) && sizeof(Resid) == 7 * sizeof(double), "Carry / Resid are arrays of doubles");
      double* GC = I + D.gc;
      if (l == 0) {
        const double* c = reinterpret_cast<const double*>(&cy);
        const double* r = reinterpret_cast<const double*>(&rs);
        double* o = GC + GC_STRIDE * g;
#pragma unroll
        for (int q = 0; q < kCarryRows; q += 2) *reinterpret_cast<double2*>(o + q) = make_double2(c[q], c[q + 1]);
#pragma unroll
        for (int q = 0; q < 6; q += 2) *reinterpret_cast<double2*>(o + kCarryRows + q) = make_double2(r[q], r[q + 1]);
        o[kCarryRows + 6] = r[6];
      }
      __syncwarp();
      if (lane < kCarryRows + 7) {
        const double a = GC[lane], b = GC[GC_STRIDE + lane], c = GC[2 * GC_STRIDE + lane], d = GC[3 * GC_STRIDE + lane];
        const bool is_sum = lane >= 8 && lane <= 11;
        GC[GC_TOT + lane] = is_sum ? ((a + b) + c) + d : dmax(dmax(a, b), dmax(c, d));
      }
      __syncwarp();
      {
        double* c = reinterpret_cast<double*>(&cy);
        double* r = reinterpret_cast<double*>(&rs);
#pragma unroll
        for (int q = 0; q < kCarryRows; q += 2) { const double2 t = lds2(GC + GC_TOT + q); c[q] = t.x; c[q + 1] = t.y; }
#pragma unroll
        for (int q = 0; q < 6; q += 2) { const double2 t = lds2(GC + GC_TOT + kCarryRows + q); r[q] = t.x; r[q + 1] = t.y; }
        r[6] = GC[GC_TOT + kCarryRows + 6];
      }
    }
    if (act) {
      ++it;
      Verdict V;
      status = decide_core(M, status, it, P.fixed != 0, cy, rs, I[LS_BINF], mu, V);
      I[LS_RES + 0] = V.pres;
      I[LS_RES + 1] = V.dres;
      if (V.has_tol) { I[LS_RES + 2] = V.tol_p; I[LS_RES + 3] = V.tol_d; }
      const bool done = status >= ST_CONVERGED || (P.fixed && --left <= 0);
      if (done) {  // results go home; the lanes are free for the next instance
        *reinterpret_cast<int2*>(I + LS_CTL) = make_int2(status, it);
        I[LS_MU] = mu;
        __syncwarp(imask);
        double* Th = P.home + ((size_t)(home_slot >> 5) * M.off.rows) * 32 + (home_slot & 31);
        lane_retire(M, D, I, Th, wl, nl, GPI > 1);
        if (P.keep_ws) lane_retire_workspace(M, D.joint0, I, Th, wl, nl, GPI > 1);
        home_slot = -1;
      }
    }
  }
}

}  // namespace loik

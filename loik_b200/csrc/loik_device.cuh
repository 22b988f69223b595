// loik_device.cuh -- device-side data model and the three tree sweeps of one LoIK ADMM iteration.
//
// Mapping: ONE THREAD = ONE PROBLEM INSTANCE.  A warp therefore walks the same joint of 32 instances
// in lock-step: every global access is a fully coalesced 256 B line per field row, there is no
// divergence (joint types / tree topology are batch-uniform and live in __constant__ memory), and
// the chain dependency between a joint and its parent never leaves the thread's registers.
//
// HBM layout (tile-major, joint-major SoA inside a tile): instances are grouped in tiles of 32 (one
// warp).  A tile is one contiguous record of `rows` rows x 32 lanes of doubles; every per-instance
// quantity is a set of rows, row index = field offset + (joint-1)*width + component.  A warp reads a
// row as one 256 B line, and because the lane stride is fixed the row offsets inside a joint step are
// compile-time immediates of the load/store instructions (no per-access address arithmetic).
//
// The maths restates loik-loid-optimized.hxx (reference file:line cited per block); nothing here is
// a translation of the Eigen/Pinocchio templates: H is kept as three 3x3 blocks (LL sym, LA, AA sym
// = 21 scalars), liMi is rebuilt from (sin q, cos q) and the constant placement, and UpdatePrev /
// ResetInfNorms / the delta_* copies of the reference become "read the old value before
// overwriting it" inside the sweeps.
#pragma once
#include <cstdint>

namespace loik {

// Capacities of the batch-uniform parameter block (ModelC travels to every kernel BY VALUE: 32 764 bytes of parameter
// space).  The per-joint record is packed (references through an index, shorts) so that 104 joints fit.
constexpr int kMaxJoints = 104;  // model.njoints incl. the universe
constexpr int kMaxTasks = 8;     // tasks whose matrix A is shared by the batch (kept in the block); more tasks (up to kMaxTasksAll)
constexpr int kMaxTasksAll = 32; //   keep A per instance in the task rows of the tile record (ModelC::a_per)
constexpr int kMaxHref = 33;     // distinct (H_ref, v_ref) references: 1 after UpdateReference, one per joint after UpdateReferences
constexpr int kMaxPin = 6;       // children per joint that are not carried in registers
constexpr int kMaxSeg = 16;      // chains of the tree that can be swept by different warps
constexpr int kMaxMd = 16;       // multi-DoF joints (free-flyer, spherical, translation, planar) per model

// status of an instance (per-instance loop control of Solve()/InfeasibilityTailSolve())
// ST_CONVERGED_PINF: converged_ with primal_infeasible_ raised in the same iteration -- the reference evaluates both
// checks before it looks at either flag (hpp:421-432), stops on converged_ first and leaves primal_infeasible_ set for
// get_primal_infeasibility_status() (0.03 % of a UR10-262 144 batch end this way)
enum : int { ST_RUNNING = 0, ST_TAIL = 1, ST_CONVERGED = 2, ST_INFEASIBLE_DONE = 3, ST_MAXITER = 4, ST_CONVERGED_PINF = 5 };

struct HrefC {
  double A[6], B[9], D[6];          // problem_.H_refs_[i] as blocks LL (sym), LA, AA (sym)
  double Hv[6];                     // problem_.Hv[i] = H_ref v_ref
};

struct JointC {
  double plR[9], plp[3], axis[3];   // model.jointPlacements[i], joint axis
  double lb, ub;                    // problem_.lb_/ub_ for this joint's dof (when shared by the batch)
  int sel0;                         // multi-DoF joints: the components S selects (4 bits each, dof k in bits 4k..4k+3)
  short parent, jtype, task;        // task: slot of the task on this joint or -1
  short idxv, idxq;                 // jmodel.idx_v(), jmodel.idx_q()
  short carry;                      // contribution to the parent travels in registers (parent == i-1, only child)
  short pout;                       // else (parent > 0): the pending block this joint writes its contribution to
  short npin;                       // number of children that hand their contribution over through a pending block
  short qkind;                      // how q parametrises the joint: 0 = one scalar, 1 = (cos, sin) (unbounded revolute)
  short nvj, mblk;                  // multi-DoF joints: nv of the joint (3 / 6), index of its md block
  short sidx;                       // aligned 1-DoF joints: the component of a [lin; ang] 6-vector S selects (S = e_sidx); -1 otherwise
  short href;                       // this joint's entry of ModelC::href
  short loff;                       // k_iterate_lane<4>: offset of the joint's block in the shared-memory instance record (assign_segments)
  short pin[kMaxPin];               // the pending blocks read (one per tree edge: single writer, no read-modify-write)
};

struct TaskC {
  double A[36];                     // problem_.Ais_[k]
  double AtA_A[6], AtA_B[9], AtA_D[6];
};

// Tile record layout.  Rows are grouped so that everything one joint step touches is contiguous and
// addressed as  base(joint) + compile-time row * 256 B:
//   [ globals | joint 1 | joint 2 | ... | task 0 | ... | pending slots | debug vectors ]
enum : int {  // rows of a joint block
  JR_V = 0, JR_F = 6, JR_FD = 12, JR_NU = 18, JR_Z = 19, JR_W = 20, JR_T = 21,  // persistent state (22 rows)
  JR_JQ = 22, JR_LB = 24, JR_UB = 25, JR_Q = 26,                                  // per-instance problem data (5 rows)
  JR_H = 27, JR_P = 48, JR_UD = 54, JR_DINV = 60, JR_R = 61,                      // backward -> forward workspace (35 rows)
  JR_HV = 62,                                                                      // H_ref v_ref of this instance (only touched with per-instance references, ModelC::vref_per)
  JR_HREF = 68,                                                                    // H_ref of this instance, 21 scalars like H (only touched with ModelC::href_per)
  JR_ROWS = 89
};
// rows of a task block: state y, Aty | per-instance problem data b, A^T b | per-instance task matrix A (row-major 6x6)
// and A^T A (21 scalars: LL sym, LA, AA sym) when the batch does not share A (ModelC::a_per)
enum : int { TR_Y = 0, TR_ATY = 6, TR_B = 12, TR_ATB = 18, TR_A = 24, TR_ATA = 60, TR_ROWS = 81 };
enum : int { PR_H = 0, PR_F = 27, PR_ROWS = 33 };                                 // rows of a pending-accumulator block
enum : int { GR_MU = 0, GR_BINF = 1, GR_CTL = 2, GR_RES = 3, GR_CARRY = 7, GR_NORMS = 21, GR_ROWS = 49 };  // globals (norms: 28 rows)

// Multi-DoF joints whose motion subspace selects components (JointModelFreeFlyer: S = I6, nq 7; JointModelSpherical:
// S = [0; I3], nq 4; JointModelTranslation: S = [I3; 0], nq 3): their 6-vector quantities (v, f, F, H, p) use the rows of
// their joint block like every joint; what is per-dof (K = 3 / 6 instead of 1) lives in an extra block per such joint
// (sized for K = 6).
enum : int { FR_NU = 0, FR_Z = 6, FR_W = 12, FR_T = 18,   // state (24 rows)
             FR_LB = 24, FR_UB = 30, FR_Q = 36,          // problem data (19 rows): bounds, q (up to 7: x y z qx qy qz qw)
             FR_XF = 43,                                 // liMi = placement * M(q): rotation (9, row-major), translation (3)
             FR_DINV = 55, FR_R = 76, FR_UD = 82,        // workspace: Dinv (21, symmetric packed), r (6), UDinv (6 x K, [a][k])
             FR_S = 118,                                 // configuration-dependent motion subspace S = [0; E(q)] (SphericalZYX): E 3 x 3 row-major (k_set_q)
             FR_ROWS = 127 };

struct Offs {
  int glob, joint0, task0, pend0;  // first row of the globals, of joint 1, of task 0, of pending slot 0
  int ff0;                         // first row of the multi-DoF blocks (FR_ROWS each)
  int prv, drv, drows;             // debug arena (StateP::dbg, allocated by loik_set_debug): primal / dual residual vectors (6 nb + nv rows each), rows per tile
  int rows;                        // rows per tile record
};

// A segment = a maximal chain lo..hi (parent(i) == i-1, single child) of the tree.  Segments only exchange data
// through pending blocks / the parent's v row, so different warps can sweep them; `blevel` / `flevel` order them
// (children before parents on the way down to the root, parents before children on the way out).
struct SegC { short lo, hi, bwarp, blevel, fwarp, flevel; };
// A span = a maximal run lo..hi of consecutive 1-DoF joints, or one multi-DoF joint (lo == hi, md = its nv): the units
// the one-warp-per-tile kernel sweeps, so that the multi-DoF steps stay out of the register-carried joint loops.
struct SpanC { short lo, hi, md, pad; };
constexpr int kMaxSpan = 2 * kMaxMd + 1;

struct ModelC {
  int nj, nb, nc, npend;
  int max_iter, bounds_per_instance;
  int nseg, nblevel, nflevel, nwarp;
  int nmd, nv, nq, href_uniform;   // number of multi-DoF joints; model.nv, model.nq; every joint shares H_ref / v_ref (UpdateReference)
  SegC seg[kMaxSeg];
  int nsb, nsf;                    // steps of the wide sweeps of k_iterate_lane<4> (backward / forward order; build_wide_table)
  int vref_per, href_per;          // v_ref / also H_ref differ per instance (loik_update_references_batch): H_ref v_ref in rows JR_HV, not HrefC::Hv
  int nspan, a_per;                // a_per: A_k (and A_k^T A_k) differ per instance: rows TR_A / TR_ATA of the task blocks, not TaskC
  SpanC span[kMaxSpan];
  Offs off;
  double rho, mu0, mu_scale, tol_abs, tol_rel, tol_pinf, tol_dinf, tol_tail, Hv_inf;
  short task_joint[kMaxTasksAll];  // joint of every task slot (TaskC::joint only exists for the first kMaxTasks)
  HrefC href[kMaxHref];
  JointC j[kMaxJoints];
  TaskC t[kMaxTasks];
};
static_assert(sizeof(ModelC) <= 32764 - 256, "ModelC + StateP + scalars must fit the kernel parameter space");

// Sweep-to-sweep scalars of one iteration (the reference keeps them in IkIdData, data hpp:259-329).
struct Carry {
  double nu_inf, dfis_inf, dvis_inf, dnu_inf, dz_inf, dyis_inf, dw_inf, Av_inf;
  double bTdy_p, bTdy_m, ubdw_p, lbdw_m, pres_task, pres_slack;
};
constexpr int kCarryRows = 14;

// per-iteration record of the solver log (loik_get_history): what loik-loid-optimized.hpp:406-420 pushes into LoikSolverInfo after
// ComputeResiduals -- primal_residual_task, primal_residual_slack, dual_residual_v, dual_residual_nu, mu (the value the iteration
// ran with; mu_eq = mu_equality_scale_factor * mu, mu_ineq = mu) -- plus what the tail-solve lists hold (hpp:290-306):
// delta_x_qp_inf_norm, delta_z_inf_norm, and whether the iteration belonged to InfeasibilityTailSolve
constexpr int kHistCols = 8;
struct StateP {
  double* arena;      // tile records
  int n;              // slots in use (instances)
  const int* n_dev;   // if set: device-resident slot count (a packed arena), overrides n
  const int* list;    // optional compaction list: thread k works on slot list[k]
  const int* n_list;  // device-resident length of `list`
  int* n_active;      // device counter: instances still active after this launch
  // Migrating launch (the re-pack fused into the iteration): thread k < *n_list reads the instance in slot list[k]
  // of `arena` during its first iteration and writes everything to slot k of `dst` (dense prefix, full tiles),
  // where it then stays for the rest of the launch.  Survivors of the launch claim the slots of the next launch in
  // next_list / next_count; instances that finish copy their results to their slot of `home` (origin_* = home slot of
  // a packed slot; origin_src == nullptr: `arena` is the home arena).
  double* dst;
  const int* origin_src;
  int* origin_dst;
  int* next_list;
  int* next_count;
  // hand-over to the lane-parallel kernel: survivors that look far from done (main loop, residual > hard_ratio x its
  // tolerance) claim from the front of next_list as usual, the others from its END (next_back counts them), so that the
  // kernel's work queue starts with the instances whose remaining iterations bound the solve's latency
  int* next_back;
  int next_cap;
  double hard_ratio;
  double* home;
  int keep_ws;        // retiring instances also carry their backward->forward workspace home (loik_set_keep_workspace)
  int drop_ws;        // the forward sweep drops the consumed workspace lines from L2 (discard_workspace); never with keep_ws
  double* hist;       // logging (LoikSolverInfo, loik_set_logging): [slot][hist_cap][kHistCols] per-iteration records of debug-mode solves
  int hist_cap;
  double* dbg;        // debug mode only: tile records of the residual vectors (Offs::prv / drv rows, Offs::drows per tile), by home slot
};

// The batch-uniform block (model, problem constants, hyper-parameters) is passed to every kernel BY VALUE as a
// __grid_constant__ parameter: it lands in the constant bank (uniform LDC reads, like a __constant__ symbol) but is
// private to the launch, so any number of solvers can have kernels in flight on different streams concurrently
// (a __constant__ symbol is a per-module singleton; a device buffer read with LDG measured 1.8x slower).

#define LOIK_DEV __device__ __forceinline__
#define LOIK_DEV_CALL __device__ __noinline__  // rare paths kept out of the hot loops' register allocation

__host__ __device__ __forceinline__ constexpr int si(int i, int j) { return i <= j ? (i * (5 - i)) / 2 + j : (j * (5 - j)) / 2 + i; }
// T = this thread's lane inside its tile record; row r lives at T[r * 32].
LOIK_DEV double* tile_ptr(const StateP& S, const ModelC& c_model, int s) { return S.arena + ((size_t)(s >> 5) * c_model.off.rows) * 32 + (s & 31); }
LOIK_DEV double ld(const double* P, int row) { return P[row * 32]; }
LOIK_DEV void st(double* P, int row, double x) { P[row * 32] = x; }
// rows read once per iteration and rewritten by the same sweep (F = fis_diff_plus_Aty, T = Stf_plus_w): streaming hints
LOIK_DEV double ld_cs(const double* P, int row) { return __ldcs(P + row * 32); }
LOIK_DEV void st_cs(double* P, int row, double x) { __stcs(P + row * 32, x); }
// block base pointers: computed once per joint step, rows inside a block are immediates
LOIK_DEV double* joint_blk(double* T, const Offs& O, int ji) { return T + (size_t)(O.joint0 + JR_ROWS * ji) * 32; }
LOIK_DEV double* task_blk(double* T, const Offs& O, int k) { return T + (size_t)(O.task0 + TR_ROWS * k) * 32; }
// task matrices: batch-shared (parameter block) or per instance (rows of the task block Pk)
LOIK_DEV double task_A(const ModelC& M, const TaskC& K, const double* Pk, int i) { return M.a_per ? Pk[(TR_A + i) * 32] : K.A[i]; }
// H_ref v_ref of joint block P (problem_.Hv[i], ik-id-description-optimized.hpp:93,113): batch-shared or this instance's own
// VR: compiled in only for the general instantiations of the iteration kernels (template parameter MD: models with multi-DoF
// joints or per-instance references); the kernels the BASELINE robots run read the constant and carry no extra code
// (1.2 % of a dense Panda launch when it was a run-time test).  Plain load: a migrating launch writes these rows.
template <bool VR>
LOIK_DEV double hv_of(const ModelC& M, const HrefC& Hr, const double* P, int c) { return (VR && M.vref_per) ? ld(P, JR_HV + c) : Hr.Hv[c]; }
// per-instance weights (ModelC::href_per): rows JR_HREF of joint block P replace the reference table's blocks
LOIK_DEV void href_rows(const double* P, double (&A)[6], double (&B)[9], double (&D)[6]) {
#pragma unroll
  for (int c = 0; c < 6; ++c) { A[c] = ld(P, JR_HREF + c); D[c] = ld(P, JR_HREF + 15 + c); }
#pragma unroll
  for (int c = 0; c < 9; ++c) B[c] = ld(P, JR_HREF + 6 + c);
}
LOIK_DEV void href_times(const double* P, const double (&v)[6], double (&Hrv)[6]) {
  double A[6], B[9], D[6];
  href_rows(P, A, B, D);
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    Hrv[a] = A[si(a, 0)] * v[0] + A[si(a, 1)] * v[1] + A[si(a, 2)] * v[2] + B[3 * a] * v[3] + B[3 * a + 1] * v[4] + B[3 * a + 2] * v[5];
    Hrv[3 + a] = B[a] * v[0] + B[3 + a] * v[1] + B[6 + a] * v[2] + D[si(a, 0)] * v[3] + D[si(a, 1)] * v[4] + D[si(a, 2)] * v[5];
  }
}
LOIK_DEV double* pend_blk(double* T, const Offs& O, int k) { return T + (size_t)(O.pend0 + PR_ROWS * k) * 32; }
LOIK_DEV double* glob_blk(double* T, const Offs& O) { return T + (size_t)O.glob * 32; }
// this lane's record of the debug arena (instances never migrate in debug mode: slot = home slot)
LOIK_DEV double* dbg_ptr(const StateP& S, const ModelC& c_model, int s) { return S.dbg ? S.dbg + ((size_t)(s >> 5) * c_model.off.drows) * 32 + (s & 31) : nullptr; }
LOIK_DEV double* md_blk(double* T, const Offs& O, int m) { return T + (size_t)(O.ff0 + FR_ROWS * m) * 32; }
__host__ __device__ __forceinline__ constexpr int s6(int i, int j) { return i <= j ? i * 6 - (i * (i - 1)) / 2 + (j - i) : j * 6 - (j * (j - 1)) / 2 + (i - j); }
// Running inf-norms and the box projection are compare + select: sm_100a has no fp64 min/max instruction, and
// fmax()/fmin() expand to DSETP + FSEL + SEL + a NaN-quieting LOP3 + register-pair shuffles (8-9 instructions each,
// ~30 % of the dynamic instructions of an iteration with ~54 norm updates per joint).  Same values as fmax/fmin for
// ordered operands; a NaN second operand is ignored like fmax/fmin do (the comparison is false) and the running
// norms (first operand) never hold one.
LOIK_DEV double dmax(double a, double b) { return b > a ? b : a; }
LOIK_DEV double dmin(double a, double b) { return b < a ? b : a; }
#ifndef LOIK_AMAX_VARIANT
#define LOIK_AMAX_VARIANT 3
#endif
#if LOIK_AMAX_VARIANT == 1
LOIK_DEV double amax(double m, double x) { const double a = fabs(x); return a > m ? a : m; }
#else
LOIK_DEV double amax(double m, double x) {  // max(m, |x|): |x| is a sign-bit mask on the high word, not an fp64 operation
  const bool gt = fabs(x) > m;
  const int hi = gt ? (__double2hiint(x) & 0x7fffffff) : __double2hiint(m);
  const int lo = gt ? __double2loint(x) : __double2loint(m);
  return __hiloint2double(hi, lo);
}
#endif
// max(m, |x_0|, ..., |x_5|).  max is exactly associative, so the order of the comparisons does not change the value;
// variant 3 uses a tree (depth 4 instead of a serial chain of 6 dependent compare-selects).
LOIK_DEV double absmax2(double a, double b) {
  const bool gt = fabs(a) > fabs(b);
  const int hi = (gt ? __double2hiint(a) : __double2hiint(b)) & 0x7fffffff;
  const int lo = gt ? __double2loint(a) : __double2loint(b);
  return __hiloint2double(hi, lo);
}
LOIK_DEV double amax6(double m, const double (&x)[6]) {
#if LOIK_AMAX_VARIANT == 3
  const double a = absmax2(x[0], x[1]), b = absmax2(x[2], x[3]), c = absmax2(x[4], x[5]);
  return dmax(dmax(m, a), dmax(b, c));
#else
#pragma unroll
  for (int c = 0; c < 6; ++c) m = amax(m, x[c]);
  return m;
#endif
}

// liMi = jointPlacements[i] * M_i(q)   (FwdPassInit, hxx:263-264).  (a, b) = (sin q, cos q) for
// revolute joints, (q, -) for prismatic ones.
LOIK_DEV void make_xf(const JointC& J, double a, double b, double (&R)[9], double (&t)[3]) {
  const double* P = J.plR;
  t[0] = J.plp[0]; t[1] = J.plp[1]; t[2] = J.plp[2];
  switch (J.jtype) {
    case 0:  // RX: columns 1,2 rotate
#pragma unroll
      for (int i = 0; i < 3; ++i) { R[3 * i] = P[3 * i]; R[3 * i + 1] = b * P[3 * i + 1] + a * P[3 * i + 2]; R[3 * i + 2] = b * P[3 * i + 2] - a * P[3 * i + 1]; }
      break;
    case 1:  // RY: columns 0,2 rotate
#pragma unroll
      for (int i = 0; i < 3; ++i) { R[3 * i] = b * P[3 * i] - a * P[3 * i + 2]; R[3 * i + 1] = P[3 * i + 1]; R[3 * i + 2] = b * P[3 * i + 2] + a * P[3 * i]; }
      break;
    case 2:  // RZ: columns 0,1 rotate
#pragma unroll
      for (int i = 0; i < 3; ++i) { R[3 * i] = b * P[3 * i] + a * P[3 * i + 1]; R[3 * i + 1] = b * P[3 * i + 1] - a * P[3 * i]; R[3 * i + 2] = P[3 * i + 2]; }
      break;
    case 6: {  // revolute unaligned: Rodrigues (pinocchio toRotationMatrix)
      const double x = J.axis[0], y = J.axis[1], z = J.axis[2], v = 1.0 - b;
      double M[9];
      M[0] = b + v * x * x;     M[1] = v * x * y - a * z; M[2] = v * x * z + a * y;
      M[3] = v * y * x + a * z; M[4] = b + v * y * y;     M[5] = v * y * z - a * x;
      M[6] = v * z * x - a * y; M[7] = v * z * y + a * x; M[8] = b + v * z * z;
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int k = 0; k < 3; ++k) R[3 * i + k] = P[3 * i] * M[k] + P[3 * i + 1] * M[3 + k] + P[3 * i + 2] * M[6 + k];
      break;
    }
    default: {  // prismatic: R = placement, t = p + R axis q
#pragma unroll
      for (int i = 0; i < 9; ++i) R[i] = P[i];
      double ax, ay, az;
      if (J.jtype == 7) { ax = J.axis[0]; ay = J.axis[1]; az = J.axis[2]; }
      else { ax = J.jtype == 3 ? 1.0 : 0.0; ay = J.jtype == 4 ? 1.0 : 0.0; az = J.jtype == 5 ? 1.0 : 0.0; }
#pragma unroll
      for (int i = 0; i < 3; ++i) t[i] += (P[3 * i] * ax + P[3 * i + 1] * ay + P[3 * i + 2] * az) * a;
      break;
    }
  }
}

// S^T x for a 6-vector x = [lin; ang]   (jdata.S().transpose() * x, hxx:70,231)
LOIK_DEV double St_dot(const JointC& J, const double (&x)[6]) {
  switch (J.jtype) {
    case 0: return x[3];
    case 1: return x[4];
    case 2: return x[5];
    case 3: return x[0];
    case 4: return x[1];
    case 5: return x[2];
    case 6: return J.axis[0] * x[3] + J.axis[1] * x[4] + J.axis[2] * x[5];
    default: return J.axis[0] * x[0] + J.axis[1] * x[1] + J.axis[2] * x[2];
  }
}
// x += S * s   (jdata.S() * nu, hxx:134)
LOIK_DEV void S_axpy(const JointC& J, double s, double (&x)[6]) {
  switch (J.jtype) {
    case 0: x[3] += s; break;
    case 1: x[4] += s; break;
    case 2: x[5] += s; break;
    case 3: x[0] += s; break;
    case 4: x[1] += s; break;
    case 5: x[2] += s; break;
    case 6: x[3] += J.axis[0] * s; x[4] += J.axis[1] * s; x[5] += J.axis[2] * s; break;
    default: x[0] += J.axis[0] * s; x[1] += J.axis[1] * s; x[2] += J.axis[2] * s; break;
  }
}
// U = H S for H = [[A, B], [B^T, D]]   (calc_aba: U = I.col(k) for aligned joints, I*S otherwise; P1)
LOIK_DEV void H_times_S(const JointC& J, const double (&A)[6], const double (&B)[9], const double (&D)[6], double (&U)[6]) {
#define LOIK_COL_LIN(k) { U[0] = A[si(0, k)]; U[1] = A[si(1, k)]; U[2] = A[si(2, k)]; U[3] = B[3 * k]; U[4] = B[3 * k + 1]; U[5] = B[3 * k + 2]; }
#define LOIK_COL_ANG(k) { U[0] = B[k]; U[1] = B[3 + k]; U[2] = B[6 + k]; U[3] = D[si(0, k)]; U[4] = D[si(1, k)]; U[5] = D[si(2, k)]; }
  switch (J.jtype) {
    case 0: LOIK_COL_ANG(0) break;
    case 1: LOIK_COL_ANG(1) break;
    case 2: LOIK_COL_ANG(2) break;
    case 3: LOIK_COL_LIN(0) break;
    case 4: LOIK_COL_LIN(1) break;
    case 5: LOIK_COL_LIN(2) break;
    case 6: {
      const double x = J.axis[0], y = J.axis[1], z = J.axis[2];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        U[i] = B[3 * i] * x + B[3 * i + 1] * y + B[3 * i + 2] * z;
        U[3 + i] = D[si(i, 0)] * x + D[si(i, 1)] * y + D[si(i, 2)] * z;
      }
      break;
    }
    default: {
      const double x = J.axis[0], y = J.axis[1], z = J.axis[2];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        U[i] = A[si(i, 0)] * x + A[si(i, 1)] * y + A[si(i, 2)] * z;
        U[3 + i] = B[i] * x + B[3 + i] * y + B[6 + i] * z;
      }
      break;
    }
  }
#undef LOIK_COL_LIN
#undef LOIK_COL_ANG
}

// out = R M R^T for symmetric M (6 unique in, 6 unique out)
LOIK_DEV void rot_sym(const double (&R)[9], const double (&M)[6], double (&out)[6]) {
  double T[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) T[3 * i + j] = R[3 * i] * M[si(0, j)] + R[3 * i + 1] * M[si(1, j)] + R[3 * i + 2] * M[si(2, j)];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = i; j < 3; ++j) out[si(i, j)] = T[3 * i] * R[3 * j] + T[3 * i + 1] * R[3 * j + 1] + T[3 * i + 2] * R[3 * j + 2];
}
// out = R M R^T for general 3x3 M
LOIK_DEV void rot_gen(const double (&R)[9], const double (&M)[9], double (&out)[9]) {
  double T[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) T[3 * i + j] = R[3 * i] * M[j] + R[3 * i + 1] * M[3 + j] + R[3 * i + 2] * M[6 + j];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) out[3 * i + j] = T[3 * i] * R[3 * j] + T[3 * i + 1] * R[3 * j + 1] + T[3 * i + 2] * R[3 * j + 2];
}

// X* H X*^T with X* = [[R, 0], [t^ R, R]]  (pinocchio SE3actOn, call site hxx:66; P2):
//   A' = Ab;  B' = Bb + (t^ Ab)^T;  D' = Db + t^ B' + (t^ Bb)^T     with Xb = R X R^T.
LOIK_DEV void congruence(const double (&R)[9], const double (&t)[3], const double (&A)[6], const double (&B)[9],
                         const double (&D)[6], double (&Ao)[6], double (&Bo)[9], double (&Do)[6]) {
  double Bb[9], Db[6];
  rot_sym(R, A, Ao);
  rot_gen(R, B, Bb);
  rot_sym(R, D, Db);
  // TA = t^ Ab (column-wise cross products), B' = Bb + TA^T
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const double a0 = Ao[si(0, j)], a1 = Ao[si(1, j)], a2 = Ao[si(2, j)];
    Bo[3 * j + 0] = Bb[3 * j + 0] + (t[1] * a2 - t[2] * a1);
    Bo[3 * j + 1] = Bb[3 * j + 1] + (t[2] * a0 - t[0] * a2);
    Bo[3 * j + 2] = Bb[3 * j + 2] + (t[0] * a1 - t[1] * a0);
  }
  // M2 = t^ B' (need upper triangle), N = t^ Bb (need lower triangle): D'(i,j) = Db + M2(i,j) + N(j,i), i <= j
  double M2[9], N[9];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    M2[0 + j] = t[1] * Bo[6 + j] - t[2] * Bo[3 + j];
    M2[3 + j] = t[2] * Bo[0 + j] - t[0] * Bo[6 + j];
    M2[6 + j] = t[0] * Bo[3 + j] - t[1] * Bo[0 + j];
    N[0 + j] = t[1] * Bb[6 + j] - t[2] * Bb[3 + j];
    N[3 + j] = t[2] * Bb[0 + j] - t[0] * Bb[6 + j];
    N[6 + j] = t[0] * Bb[3 + j] - t[1] * Bb[0 + j];
  }
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = i; j < 3; ++j) Do[si(i, j)] = Db[si(i, j)] + M2[3 * i + j] + N[3 * j + i];
}

// SE3::act(Force): [R f_lin ; R f_ang + t x (R f_lin)]   (hxx:74,212; P3)
LOIK_DEV void act_force(const double (&R)[9], const double (&t)[3], const double (&f)[6], double (&o)[6]) {
  double l[3], a[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    l[i] = R[3 * i] * f[0] + R[3 * i + 1] * f[1] + R[3 * i + 2] * f[2];
    a[i] = R[3 * i] * f[3] + R[3 * i + 1] * f[4] + R[3 * i + 2] * f[5];
  }
  o[0] = l[0]; o[1] = l[1]; o[2] = l[2];
  o[3] = a[0] + (t[1] * l[2] - t[2] * l[1]);
  o[4] = a[1] + (t[2] * l[0] - t[0] * l[2]);
  o[5] = a[2] + (t[0] * l[1] - t[1] * l[0]);
}
// SE3::actInv(Motion): [R^T (v_lin - t x v_ang) ; R^T v_ang]   (hxx:125; P3)
LOIK_DEV void actinv_motion(const double (&R)[9], const double (&t)[3], const double (&v)[6], double (&o)[6]) {
  const double l0 = v[0] - (t[1] * v[5] - t[2] * v[4]);
  const double l1 = v[1] - (t[2] * v[3] - t[0] * v[5]);
  const double l2 = v[2] - (t[0] * v[4] - t[1] * v[3]);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    o[i] = R[i] * l0 + R[3 + i] * l1 + R[6 + i] * l2;
    o[3 + i] = R[i] * v[3] + R[3 + i] * v[4] + R[6 + i] * v[5];
  }
}

// ---------------------------------------------------------------------------------------------
// Memory-access discipline of the sweeps: every joint step is written as  LOAD PHASE -> maths ->
// STORE PHASE.  All global loads of a step are issued back to back before the first dependent
// instruction, so one warp keeps tens of 256 B requests in flight (the compiler cannot hoist a load
// above a store to a possibly-aliasing row, so interleaving them would serialise on DRAM latency).
// ---------------------------------------------------------------------------------------------
LOIK_DEV double ldc(const double* T, int row) { return __ldg(T + row * 32); }  // read-only data (never a row this launch writes)
// Software prefetch of the NEXT joint step's rows, issued at the top of the current step: the sweeps are chains of
// dependent steps, each starting with a batch of loads, and with 8 resident warps per SM the latency of that batch is
// exposed; a prefetch costs no register.  Target L1 (CCTL.E.PF1), and cover the rows this iteration wrote one sweep
// earlier as well (the backward sweep's workspace for the forward step, f / v / w for the residual step): they are
// still in L2, but with several solves in flight an L2 hit is slow enough to matter -- pipelined Panda 67.5 -> 69.2 M
// solves/s, UR10 110.7 -> 113.6 M, Talos 6.59 -> 6.73 M; a dense launch timed alone pays 2 % for the extra
// instructions (84.3 -> 86.0 us).  L1 instead of L2 with the old coverage: no difference; the L1 carve-out alone: no
// difference (profiles/r1_history.md).  -DLOIK_PF_BASIC = the old coverage, -DLOIK_PF_LEVEL=\"L2\" = the old target.
#ifndef LOIK_PF_LEVEL
#define LOIK_PF_LEVEL "L1"
#endif
#ifndef LOIK_PF_BASIC
#define LOIK_PF_EXT 1
#endif
LOIK_DEV void pf(const double* P, int row) { asm volatile("prefetch.global." LOIK_PF_LEVEL " [%0];" ::"l"(P + row * 32)); }
template <int N>
LOIK_DEV void pf_rows(const double* P, int row0) {
#pragma unroll
  for (int c = 0; c < N; ++c) pf(P, row0 + c);
}

// The backward->forward workspace of a joint (rows JR_H .. JR_R: His, pis, UDinv, Dinv, r) is dead once the forward step
// has consumed it -- the next backward sweep rewrites it before anything reads it -- yet the dirty lines would still be
// written back to HBM when L2 evicts them (128 of the 165 MB a Panda-65 536 launch writes).  discard.global.L2 drops a
// line from L2 without the write-back.  The lanes that are executing share the 70 lines of the tile's block (a warp
// owns a whole tile here); called after the step's loads have been consumed, i.e. have returned for every lane.
LOIK_DEV void discard_workspace(const double* Pj_lane) {
  const unsigned am = __activemask();
  const int lane = threadIdx.x & 31;
  const int rank = __popc(am & ((1u << lane) - 1u)), n = __popc(am);
  const char* base = reinterpret_cast<const char*>(Pj_lane - lane) + JR_H * 256;
  constexpr int kLines = (JR_HV - JR_H) * 2;  // (rows JR_HV.. behind the workspace are problem data)
  for (int ln = rank; ln < kLines; ln += n) asm volatile("discard.global.L2 [%0], 128;" ::"l"(base + ln * 128) : "memory");
}

// ---------------------------------------------------------------------------------------------
// Backward sweep: FwdPass1 (hxx:290-338) fused into BwdPassOptimizedVisitor (hxx:345-354, algo :31-81).
// Leaves for the forward sweep, per joint: H_i and p_i (accumulated over the subtree, un-projected,
// = His[i]/pis[i] after the reference's BwdPass), UDinv_i, Dinv_i, r_i.
// ---------------------------------------------------------------------------------------------
// Every sweep takes two tile pointers: Ts, where the previous iterate and the per-instance problem data are read, and
// Td, where everything is written (and where what this iteration has already written is read back).  Ts == Td except
// in the first iteration of a migrating launch, whose backward sweep also carries the problem data rows over
// (`migrate`), so that the forward and residual sweeps read them from Td.
template <bool VR = false>
LOIK_DEV void sweep_backward(const ModelC& c_model, const double* Ts, double* Td, const double mu, const double mu_eq,
                             const int lo, const int hi, const bool migrate = false) {
  const Offs& O = c_model.off;
  const double rho = c_model.rho;
  double cA[6], cB[9], cD[6], cp[6];  // contribution carried from child i+1
  bool have_carry = false;
  for (int i = hi; i >= lo; --i) {
    const JointC& J = c_model.j[i];
    const HrefC& Hr = c_model.href[J.href];
    double* Pj = joint_blk(Td, O, i - 1);
    const double* Ps = joint_blk(const_cast<double*>(Ts), O, i - 1);
    if (i > lo) {
      const double* Pn = joint_blk(const_cast<double*>(Ts), O, i - 2);
      pf_rows<6>(Pn, JR_V); pf(Pn, JR_W); pf(Pn, JR_Z); pf_rows<2>(Pn, JR_JQ);
      const int kt = c_model.j[i - 1].task;
      if (kt >= 0) { const double* Pk = task_blk(const_cast<double*>(Ts), O, kt); pf_rows<6>(Pk, TR_ATY); pf_rows<6>(Pk, TR_ATB); }
    } else if (i == 1) {
      pf_rows<6>(Ps, JR_F); pf(Ps, JR_NU);  // first rows of the forward sweep that this sweep has not touched
    }
    // ---- load phase
    double vold[6], aty[6], atb[6];
    const double w_i = ld(Ps, JR_W), z_i = ld(Ps, JR_Z);
    const double qa = ldc(Ps, JR_JQ), qb = ldc(Ps, JR_JQ + 1);
#pragma unroll
    for (int c = 0; c < 6; ++c) vold[c] = ld(Ps, JR_V + c);
    if (J.task >= 0) {
      const double* Pk = task_blk(const_cast<double*>(Ts), O, J.task);
#pragma unroll
      for (int c = 0; c < 6; ++c) { aty[c] = ld(Pk, TR_ATY + c); atb[c] = ldc(Pk, TR_ATB + c); }
    }
    double cst[3] = {0, 0, 0}, tb[6] = {0, 0, 0, 0, 0, 0};  // migrating: the problem data rows travel with the instance
    if (migrate) {
#pragma unroll
      for (int c = 0; c < 3; ++c) cst[c] = ldc(Ps, JR_LB + c);  // lb, ub, q
      if (J.task >= 0) {
        const double* Pk = task_blk(const_cast<double*>(Ts), O, J.task);
#pragma unroll
        for (int c = 0; c < 6; ++c) tb[c] = ldc(Pk, TR_B + c);
      }
    }
    // ---- FwdPass1: H_i = rho I + Href_i (:304-306); p_i = -rho v_prev_i - Hv_i (:310-313).  v still holds the
    // previous iterate here, which is the reference's vis_prev (UpdatePrev, data hxx:192-197).
    double A[6], B[9], D[6], p[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) p[c] = -rho * vold[c] - hv_of<VR>(c_model, Hr, Ps, c);
#pragma unroll
    for (int c = 0; c < 6; ++c) { A[c] = Hr.A[c]; D[c] = Hr.D[c]; }
#pragma unroll
    for (int c = 0; c < 9; ++c) B[c] = Hr.B[c];
    if (VR && c_model.href_per) href_rows(Ps, A, B, D);
    A[0] += rho; A[3] += rho; A[5] += rho; D[0] += rho; D[3] += rho; D[5] += rho;
    if (J.task >= 0) {  // H_c += mu_eq AtA; p_c += Aty - mu_eq Atb (:327-330)
      const TaskC& K = c_model.t[J.task];
      if (c_model.a_per) {  // per-instance A: A^T A from the task block (and, migrating, A and A^T A travel with the instance)
        const double* Pk = task_blk(const_cast<double*>(Ts), O, J.task);
#pragma unroll
        for (int c = 0; c < 6; ++c) { A[c] += mu_eq * ld(Pk, TR_ATA + c); D[c] += mu_eq * ld(Pk, TR_ATA + 15 + c); }
#pragma unroll
        for (int c = 0; c < 9; ++c) B[c] += mu_eq * ld(Pk, TR_ATA + 6 + c);
        if (migrate) {
          double* Pd = task_blk(Td, O, J.task);
          for (int c = TR_A; c < TR_ROWS; ++c) st(Pd, c, ld(Pk, c));
        }
      } else {
#pragma unroll
        for (int c = 0; c < 6; ++c) { A[c] += mu_eq * K.AtA_A[c]; D[c] += mu_eq * K.AtA_D[c]; }
#pragma unroll
        for (int c = 0; c < 9; ++c) B[c] += mu_eq * K.AtA_B[c];
      }
#pragma unroll
      for (int c = 0; c < 6; ++c) p[c] += aty[c] - mu_eq * atb[c];
    }
    // children's contributions: His[parent] += SE3actOn(...), pis[parent] += liMi.act(...) (:66,:74)
    for (int n = 0; n < J.npin; ++n) {
      const double* Pp = pend_blk(Td, O, J.pin[n]);
#pragma unroll
      for (int c = 0; c < 6; ++c) { A[c] += ld(Pp, PR_H + c); D[c] += ld(Pp, PR_H + 15 + c); p[c] += ld(Pp, PR_H + 21 + c); }
#pragma unroll
      for (int c = 0; c < 9; ++c) B[c] += ld(Pp, PR_H + 6 + c);
    }
    if (have_carry) {
#pragma unroll
      for (int c = 0; c < 6; ++c) { A[c] += cA[c]; D[c] += cD[c]; p[c] += cp[c]; }
#pragma unroll
      for (int c = 0; c < 9; ++c) B[c] += cB[c];
    }
    // calc_aba (:60-63): U = H S, Dinv = 1/(S^T U + R_i) with armature R_i = mu_ineq (:294-295), UDinv = U Dinv
    double U[6], UD[6];
    H_times_S(J, A, B, D, U);
    const double Dinv = 1.0 / (St_dot(J, U) + mu);
#pragma unroll
    for (int c = 0; c < 6; ++c) UD[c] = U[c] * Dinv;
    // r_i = w_i - mu_ineq z_i (:296) + S^T p_i (:70)
    const double ri = (w_i - mu * z_i) + St_dot(J, p);
    // ---- store phase 1: hand H_i, p_i, UDinv_i, Dinv_i, r_i to the forward sweep
#pragma unroll
    for (int c = 0; c < 6; ++c) { st(Pj, JR_H + c, A[c]); st(Pj, JR_H + 15 + c, D[c]); st(Pj, JR_P + c, p[c]); st(Pj, JR_UD + c, UD[c]); }
#pragma unroll
    for (int c = 0; c < 9; ++c) st(Pj, JR_H + 6 + c, B[c]);
    st(Pj, JR_DINV, Dinv);
    st(Pj, JR_R, ri);
    if (migrate) {
      st(Pj, JR_JQ, qa); st(Pj, JR_JQ + 1, qb);
#pragma unroll
      for (int c = 0; c < 3; ++c) st(Pj, JR_LB + c, cst[c]);
      if (VR && c_model.vref_per)
        for (int c = 0; c < (c_model.href_per ? JR_ROWS - JR_HV : 6); ++c) st(Pj, JR_HV + c, ldc(Ps, JR_HV + c));  // (H_ref v_ref, and H_ref behind it)
      if (J.task >= 0) {
        double* Pk = task_blk(Td, O, J.task);
#pragma unroll
        for (int c = 0; c < 6; ++c) { st(Pk, TR_B + c, tb[c]); st(Pk, TR_ATB + c, atb[c]); }
      }
    }
    have_carry = false;
    if (J.parent > 0) {
      // projection: H -= UDinv U^T (calc_aba update_I, :63), p -= UDinv r_i (:71-73)
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = a; b < 3; ++b) { A[si(a, b)] -= UD[a] * U[b]; D[si(a, b)] -= UD[3 + a] * U[3 + b]; }
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) B[3 * a + b] -= UD[a] * U[3 + b];
#pragma unroll
      for (int c = 0; c < 6; ++c) p[c] -= UD[c] * ri;
      double R[9], t[3];
      make_xf(J, qa, qb, R, t);
      congruence(R, t, A, B, D, cA, cB, cD);
      act_force(R, t, p, cp);
      if (J.carry) {
        have_carry = true;
      } else {
        double* Pp = pend_blk(Td, O, J.pout);
#pragma unroll
        for (int c = 0; c < 6; ++c) { st(Pp, PR_H + c, cA[c]); st(Pp, PR_H + 15 + c, cD[c]); st(Pp, PR_H + 21 + c, cp[c]); }
#pragma unroll
        for (int c = 0; c < 9; ++c) st(Pp, PR_H + 6 + c, cB[c]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Forward sweep: FwdPass2OptimizedVisitor (hxx:361-377, algo :102-163) + BoxProj (:384-397) +
// DualUpdate (:404-461) + ComputePrimalResiduals (:494-503), joint by joint, root to leaves.
// ---------------------------------------------------------------------------------------------
LOIK_DEV void zero(Carry& cy) {
  cy.nu_inf = cy.dfis_inf = cy.dvis_inf = cy.dnu_inf = cy.dz_inf = cy.dyis_inf = cy.dw_inf = cy.Av_inf = 0.0;
  cy.bTdy_p = cy.bTdy_m = cy.ubdw_p = cy.lbdw_m = cy.pres_task = cy.pres_slack = 0.0;
}
// Accumulates into `cy` (the caller zeroes it once per iteration).
template <bool DEBUG>
LOIK_DEV void sweep_forward(const ModelC& c_model, const double* Ts, double* Td, const double mu, const double mu_eq, Carry& cy,
                            const int lo, const int hi, const bool drop_ws = false, double* Dg = nullptr) {
  const Offs& O = c_model.off;
  const int nb = c_model.nb;
  const double inv_mu = 1.0 / mu;
  double vprev[6] = {0, 0, 0, 0, 0, 0};  // v of joint i-1
  for (int i = lo; i <= hi; ++i) {
    const JointC& J = c_model.j[i];
    const int ji = i - 1;
    double* Pj = joint_blk(Td, O, ji);
    const double* Ps = joint_blk(const_cast<double*>(Ts), O, ji);
    if (i < hi) {
      const double* Pn = joint_blk(const_cast<double*>(Ts), O, ji + 1);
      const double* Pnd = joint_blk(Td, O, ji + 1);
      pf_rows<6>(Pn, JR_V); pf_rows<6>(Pn, JR_F); pf(Pn, JR_NU); pf(Pn, JR_Z); pf(Pn, JR_W); pf_rows<2>(Pnd, JR_JQ);
#ifdef LOIK_PF_EXT
      pf_rows<35>(Pnd, JR_H);  // (L1 prefetch: also what is still in L2)
      if (c_model.bounds_per_instance) pf_rows<2>(Pnd, JR_LB);
#else
      if (nb > 16) pf_rows<35>(Pnd, JR_H);  // long trees: the workspace written by the backward sweep has left L2 by now
#endif
      const int kt = c_model.j[i + 1].task;
      if (kt >= 0) { pf_rows<6>(task_blk(Td, O, kt), TR_B); pf_rows<6>(task_blk(const_cast<double*>(Ts), O, kt), TR_Y); }
    }
    // ---- load phase A: what nu_i, v_i and the dof update need
    double vin[6], UD[6], vold[6];
    const double qa = ld(Pj, JR_JQ), qb = ld(Pj, JR_JQ + 1);
    const double Dinv = ld(Pj, JR_DINV), ri = ld(Pj, JR_R);
    const double w_old = ld(Ps, JR_W), nu_old = ld(Ps, JR_NU), z_old = ld(Ps, JR_Z);
    const double lb = c_model.bounds_per_instance ? ld(Pj, JR_LB) : J.lb, ub = c_model.bounds_per_instance ? ld(Pj, JR_UB) : J.ub;
#pragma unroll
    for (int c = 0; c < 6; ++c) { UD[c] = ld(Pj, JR_UD + c); vold[c] = ld(Ps, JR_V + c); }
    if (J.parent == 0) {
#pragma unroll
      for (int c = 0; c < 6; ++c) vin[c] = 0.0;
    } else if (J.parent == i - 1 && i > lo) {
#pragma unroll
      for (int c = 0; c < 6; ++c) vin[c] = vprev[c];
    } else {
      const double* Pq = joint_blk(Td, O, J.parent - 1);  // the parent's new v (this sweep)
#pragma unroll
      for (int c = 0; c < 6; ++c) vin[c] = ld(Pq, JR_V + c);
    }
    // ---- maths A
    double R[9], t[3], v[6];
    make_xf(J, qa, qb, R, t);
    actinv_motion(R, t, vin, v);  // vi_parent (:125)
    // nu_i = -UDinv^T vp - Dinv r_i (:127)
    double acc = 0.0;
#pragma unroll
    for (int c = 0; c < 6; ++c) acc += UD[c] * v[c];
    const double nu = -acc - Dinv * ri;
    cy.nu_inf = amax(cy.nu_inf, nu);  // (:129-131)
    S_axpy(J, nu, v);                 // v_i = vp + S nu_i (:133-134)
    {                                 // delta_vis_inf_norm vs the previous iterate (:156-158)
      double dv[6];
#pragma unroll
      for (int c = 0; c < 6; ++c) { dv[c] = v[c] - vold[c]; vprev[c] = v[c]; }
      cy.dvis_inf = amax6(cy.dvis_inf, dv);
    }
    // this joint's dof: delta_nu (:375), BoxProj (:388-394), w update (:454-458), CheckFeasibility's dot products (:588,590)
    cy.dnu_inf = amax(cy.dnu_inf, nu - nu_old);
    const double z = dmin(ub, dmax(lb, nu + inv_mu * w_old));
    cy.dz_inf = amax(cy.dz_inf, z - z_old);
    const double rp = nu - z;
    cy.pres_slack = amax(cy.pres_slack, rp);
    const double dw = mu * rp;
    cy.dw_inf = amax(cy.dw_inf, dw);
    cy.ubdw_p += ub * dmax(dw, 0.0);
    cy.lbdw_m += lb * dmin(dw, 0.0);
    // ---- load phase B: f_i = H_i v_i + p_i (:139-140), delta_fis (:137-146).  Issued before the stores of
    // phase A so the requests overlap the maths above.
    {
      double fold[6], p[6], A[6], B[9], D[6], f[6];
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        fold[c] = ld(Ps, JR_F + c); p[c] = ld(Pj, JR_P + c);
        A[c] = ld(Pj, JR_H + c); D[c] = ld(Pj, JR_H + 15 + c);
      }
#pragma unroll
      for (int c = 0; c < 9; ++c) B[c] = ld(Pj, JR_H + 6 + c);
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        f[a] = A[si(a, 0)] * v[0] + A[si(a, 1)] * v[1] + A[si(a, 2)] * v[2] + B[3 * a] * v[3] + B[3 * a + 1] * v[4] + B[3 * a + 2] * v[5] + p[a];
        f[3 + a] = B[a] * v[0] + B[3 + a] * v[1] + B[6 + a] * v[2] + D[si(a, 0)] * v[3] + D[si(a, 1)] * v[4] + D[si(a, 2)] * v[5] + p[3 + a];
      }
      {
        double df[6];
#pragma unroll
        for (int c = 0; c < 6; ++c) df[c] = f[c] - fold[c];
        cy.dfis_inf = amax6(cy.dfis_inf, df);
      }
      // ---- store phase
#pragma unroll
      for (int c = 0; c < 6; ++c) { st(Pj, JR_V + c, v[c]); st(Pj, JR_F + c, f[c]); }
    }
    st(Pj, JR_NU, nu);
    st(Pj, JR_Z, z);
    st(Pj, JR_W, w_old + dw);
    if (DEBUG && Dg) st(Dg, O.prv + 6 * nb + J.idxv, rp);
    if (J.task >= 0) {  // DualUpdate for the task on this joint (:410-451)
      const TaskC& K = c_model.t[J.task];
      double* Pk = task_blk(Td, O, J.task);
      const double* Pks = task_blk(const_cast<double*>(Ts), O, J.task);
      double y[6], bk[6], yk[6];
#pragma unroll
      for (int c = 0; c < 6; ++c) { bk[c] = ld(Pk, TR_B + c); yk[c] = ld(Pks, TR_Y + c); }
      double plus = 0.0, minus = 0.0;
#pragma unroll
      for (int a = 0; a < 6; ++a) {
        const double Av = task_A(c_model, K, Pk, 6 * a) * v[0] + task_A(c_model, K, Pk, 6 * a + 1) * v[1] + task_A(c_model, K, Pk, 6 * a + 2) * v[2] +
                          task_A(c_model, K, Pk, 6 * a + 3) * v[3] + task_A(c_model, K, Pk, 6 * a + 4) * v[4] + task_A(c_model, K, Pk, 6 * a + 5) * v[5];
        const double e = Av - bk[a];       // Av_minus_b (:416)
        const double dy = mu_eq * e;       // delta_yis (:419)
        y[a] = yk[a] + dy;
        cy.dyis_inf = amax(cy.dyis_inf, dy);
        cy.Av_inf = amax(cy.Av_inf, Av);
        cy.pres_task = amax(cy.pres_task, e);
        plus += bk[a] * dmax(dy, 0.0);
        minus += bk[a] * dmin(dy, 0.0);
        if (DEBUG && Dg) st(Dg, O.prv + 6 * ji + a, e);
      }
      cy.bTdy_p += plus;
      cy.bTdy_m += minus;
#pragma unroll
      for (int a = 0; a < 6; ++a) {
        st(Pk, TR_Y + a, y[a]);
        // Aty = A^T y (:425)
        st(Pk, TR_ATY + a, task_A(c_model, K, Pk, a) * y[0] + task_A(c_model, K, Pk, 6 + a) * y[1] + task_A(c_model, K, Pk, 12 + a) * y[2] +
                               task_A(c_model, K, Pk, 18 + a) * y[3] + task_A(c_model, K, Pk, 24 + a) * y[4] + task_A(c_model, K, Pk, 30 + a) * y[5]);
      }
    }
    if (drop_ws) discard_workspace(Pj);
  }
}

struct Resid {
  double dres_v, dres_nu, Hrefv_inf, F_inf, T_inf, dF_inf, dT_inf;
};

// ---------------------------------------------------------------------------------------------
// Residual sweep: BwdPass2OptimizedVisitor (hxx:468-487, algo :185-241) + ComputeDualResiduals (:510-522),
// leaves to root.  F = fis_diff_plus_Aty, T = Stf_plus_w.  The reference's "copy F to delta_F, zero F,
// set F_c = Aty" choreography (:364,:370,:438-439) reduces to: F_old is what is in memory, F_new is rebuilt.
// ---------------------------------------------------------------------------------------------
LOIK_DEV void zero(Resid& rs) { rs.dres_v = rs.dres_nu = rs.Hrefv_inf = rs.F_inf = rs.T_inf = rs.dF_inf = rs.dT_inf = 0.0; }
// Accumulates into `rs` (the caller zeroes it once per iteration and sets dres_nu = T_inf at the end, hxx:484).
template <bool DEBUG, bool VR = false>
LOIK_DEV void sweep_residual(const ModelC& c_model, const double* Ts, double* Td, Resid& rs, const int lo, const int hi, double* Dg = nullptr) {
  const Offs& O = c_model.off;
  const int nb = c_model.nb;
  double cF[6];
  bool have_carry = false;
  for (int i = hi; i >= lo; --i) {
    const JointC& J = c_model.j[i];
    const HrefC& Hr = c_model.href[J.href];
    const int ji = i - 1;
    double* Pj = joint_blk(Td, O, ji);
    const double* Ps = joint_blk(const_cast<double*>(Ts), O, ji);
    if (i > lo) {
      const double* Pn = joint_blk(const_cast<double*>(Ts), O, ji - 1);
      pf_rows<6>(Pn, JR_FD); pf(Pn, JR_T);
#ifdef LOIK_PF_EXT
      const double* Pnd = joint_blk(Td, O, ji - 1);
      pf_rows<6>(Pnd, JR_F); pf_rows<6>(Pnd, JR_V); pf(Pnd, JR_W); pf_rows<2>(Pnd, JR_JQ);
      const int kt = c_model.j[i - 1].task;
      if (kt >= 0) pf_rows<6>(task_blk(Td, O, kt), TR_ATY);
#endif
    }
    // ---- load phase
    double f[6], F[6], v[6], Fold[6];
    const double w_i = ld(Pj, JR_W), T_old = ld_cs(Ps, JR_T);
    const double qa = ld(Pj, JR_JQ), qb = ld(Pj, JR_JQ + 1);
#pragma unroll
    for (int c = 0; c < 6; ++c) { f[c] = ld(Pj, JR_F + c); v[c] = ld(Pj, JR_V + c); Fold[c] = ld_cs(Ps, JR_FD + c); }
    if (J.task >= 0) {
      const double* Pk = task_blk(Td, O, J.task);
#pragma unroll
      for (int c = 0; c < 6; ++c) F[c] = ld(Pk, TR_ATY + c);  // (:438-439)
    } else {
#pragma unroll
      for (int c = 0; c < 6; ++c) F[c] = 0.0;  // (:370)
    }
    for (int n = 0; n < J.npin; ++n) {
      const double* Pp = pend_blk(Td, O, J.pin[n]);
#pragma unroll
      for (int c = 0; c < 6; ++c) F[c] += ld(Pp, PR_F + c);
    }
    // ---- maths
    if (have_carry) {
#pragma unroll
      for (int c = 0; c < 6; ++c) F[c] += cF[c];
    }
#pragma unroll
    for (int c = 0; c < 6; ++c) F[c] += -f[c];  // (:210)
    // Href_v (fwd pass 2, :149-153) is recomputed here from v_i instead of being stored
    double Hrv[6], rd[6];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      Hrv[a] = Hr.A[si(a, 0)] * v[0] + Hr.A[si(a, 1)] * v[1] + Hr.A[si(a, 2)] * v[2] + Hr.B[3 * a] * v[3] + Hr.B[3 * a + 1] * v[4] + Hr.B[3 * a + 2] * v[5];
      Hrv[3 + a] = Hr.B[a] * v[0] + Hr.B[3 + a] * v[1] + Hr.B[6 + a] * v[2] + Hr.D[si(a, 0)] * v[3] + Hr.D[si(a, 1)] * v[4] + Hr.D[si(a, 2)] * v[5];
    }
    if (VR && c_model.href_per) href_times(Pj, v, Hrv);
    {
      double dF[6];
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        dF[c] = F[c] - Fold[c];
        rd[c] = Hrv[c] - hv_of<VR>(c_model, Hr, Pj, c) + F[c];  // (:228)
      }
      if (VR && c_model.vref_per) {  // |Hv|inf of this instance enters tol_dual next to |Href v|inf (:548-552): max is associative
#pragma unroll
        for (int c = 0; c < 6; ++c) rs.Hrefv_inf = amax(rs.Hrefv_inf, ld(Pj, JR_HV + c));
      }
      rs.dF_inf = amax6(rs.dF_inf, dF);             // (:215-220)
      rs.F_inf = amax6(rs.F_inf, F);                // (:223-225)
      rs.Hrefv_inf = amax6(rs.Hrefv_inf, Hrv);
      rs.dres_v = amax6(rs.dres_v, rd);
    }
    // Stf_plus_w (:231-236) and its delta (:471,:482-483)
    const double Tn = St_dot(J, f) + w_i;
    rs.T_inf = amax(rs.T_inf, Tn);
    rs.dT_inf = amax(rs.dT_inf, Tn - T_old);
    // ---- store phase
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      st_cs(Pj, JR_FD + c, F[c]);
      if (DEBUG && Dg) st(Dg, O.drv + 6 * ji + c, rd[c]);
    }
    st_cs(Pj, JR_T, Tn);
    if (DEBUG && Dg) st(Dg, O.drv + 6 * nb + J.idxv, Tn);
    have_carry = false;
    if (J.parent > 0) {  // fis_diff_plus_Aty[parent] += liMi.act(f_i) (:212)
      double R[9], t[3];
      make_xf(J, qa, qb, R, t);
      act_force(R, t, f, cF);
      if (J.carry) {
        have_carry = true;
      } else {
        double* Pp = pend_blk(Td, O, J.pout);
#pragma unroll
        for (int c = 0; c < 6; ++c) st(Pp, PR_F + c, cF[c]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Multi-DoF joints (K = J.nvj dofs, S selects the components sel[0..K) packed in J.sel0): the three sweeps' steps for one such joint,
// called from inside the joint loops.  calc_aba (P1, general form): U = H S, StU = S^T U + mu_ineq I,
// Dinv = StU^-1 (pinocchio: Cholesky, PerformStYSInversion), UDinv = U Dinv, and below a non-root joint
// H -= UDinv U^T.  They never carry to / from a neighbour in registers: every edge into or out of them is a pending
// block.  At the root (parent = universe) nothing above is read: no projection, v_parent = 0, nu = -Dinv r.
// ---------------------------------------------------------------------------------------------
template <int K>
__host__ __device__ __forceinline__ constexpr int sk(int i, int j) { return i <= j ? i * K - (i * (i - 1)) / 2 + (j - i) : j * K - (j * (j - 1)) / 2 + (i - j); }

template <int K>
LOIK_DEV void spd_inverse(double (&M)[K * K], double (&Minv)[K * (K + 1) / 2]) {  // M: full symmetric KxK, overwritten by its Cholesky factor
#pragma unroll
  for (int i = 0; i < K; ++i)
#pragma unroll
    for (int j = 0; j <= i; ++j) {
      double sum = M[K * i + j];
#pragma unroll
      for (int k = 0; k < j; ++k) sum -= M[K * i + k] * M[K * j + k];
      M[K * i + j] = (i == j) ? sqrt(sum) : sum / M[K * j + j];
    }
  double Li[K * K];  // L^-1 (lower triangular)
#pragma unroll
  for (int c = 0; c < K; ++c)
#pragma unroll
    for (int i = c; i < K; ++i) {
      double sum = (i == c) ? 1.0 : 0.0;
#pragma unroll
      for (int k = c; k < i; ++k) sum -= M[K * i + k] * Li[K * k + c];
      Li[K * i + c] = sum / M[K * i + i];
    }
#pragma unroll
  for (int i = 0; i < K; ++i)
#pragma unroll
    for (int j = i; j < K; ++j) {
      double sum = 0.0;
#pragma unroll
      for (int k = j; k < K; ++k) sum += Li[K * k + i] * Li[K * k + j];
      Minv[sk<K>(i, j)] = sum;
    }
}
// element (a, b) of H = [[A, B], [B^T, D]]
LOIK_DEV double Hel(const double (&A)[6], const double (&B)[9], const double (&D)[6], int a, int b) {
  return a < 3 ? (b < 3 ? A[si(a, b)] : B[3 * a + (b - 3)]) : (b < 3 ? B[3 * b + (a - 3)] : D[si(a - 3, b - 3)]);
}
LOIK_DEV void Hsub(double (&A)[6], double (&B)[9], double (&D)[6], int a, int b, double x) {  // H(a, b) -= x, a <= b
  if (b < 3) A[si(a, b)] -= x;
  else if (a < 3) B[3 * a + (b - 3)] -= x;
  else D[si(a - 3, b - 3)] -= x;
}
LOIK_DEV void md_load_xf(const double* Pf, double (&R)[9], double (&t)[3]) {
#pragma unroll
  for (int c = 0; c < 9; ++c) R[c] = ld(Pf, FR_XF + c);
#pragma unroll
  for (int c = 0; c < 3; ++c) t[c] = ld(Pf, FR_XF + 9 + c);
}

template <int K, bool DS = false>
LOIK_DEV_CALL void md_backward(const ModelC& c_model, const double* Ts, double* Td, const double mu, const double mu_eq, const int i,
                          const bool migrate) {
  const Offs& O = c_model.off;
  const JointC& J = c_model.j[i];
    const HrefC& Hr = c_model.href[J.href];
  int sel[K];
#pragma unroll
  for (int k = 0; k < K; ++k) sel[k] = (J.sel0 >> (4 * k)) & 7;
  double* Pj = joint_blk(Td, O, i - 1);
  double* Pf = md_blk(Td, O, J.mblk);
  const double* Pjs = joint_blk(const_cast<double*>(Ts), O, i - 1);
  const double* Pfs = md_blk(const_cast<double*>(Ts), O, J.mblk);
  const double rho = c_model.rho;
  double A[6], B[9], D[6], p[6], w[K], z[K];
  if (migrate) {  // the problem data rows travel with the instance
    for (int c = FR_LB; c < FR_DINV; ++c) st(Pf, c, ldc(Pfs, c));
    if (J.task >= 0) {
      const double* Pks = task_blk(const_cast<double*>(Ts), O, J.task);
      double* Pk = task_blk(Td, O, J.task);
      for (int c = TR_B; c < (c_model.a_per ? TR_ROWS : TR_A); ++c) st(Pk, c, ldc(Pks, c));
    }
    if (c_model.vref_per)
      for (int c = 0; c < (c_model.href_per ? JR_ROWS - JR_HV : 6); ++c) st(Pj, JR_HV + c, ldc(Pjs, JR_HV + c));
  }
#pragma unroll
  for (int c = 0; c < 6; ++c) { p[c] = -rho * ld(Pjs, JR_V + c) - hv_of<true>(c_model, Hr, Pjs, c); A[c] = Hr.A[c]; D[c] = Hr.D[c]; }
#pragma unroll
  for (int c = 0; c < K; ++c) { w[c] = ld(Pfs, FR_W + c); z[c] = ld(Pfs, FR_Z + c); }
#pragma unroll
  for (int c = 0; c < 9; ++c) B[c] = Hr.B[c];
  if (c_model.href_per) href_rows(Pjs, A, B, D);
  A[0] += rho; A[3] += rho; A[5] += rho; D[0] += rho; D[3] += rho; D[5] += rho;
  if (J.task >= 0) {
    const TaskC& Kt = c_model.t[J.task];
    const double* Pk = task_blk(const_cast<double*>(Ts), O, J.task);
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      A[c] += mu_eq * (c_model.a_per ? ld(Pk, TR_ATA + c) : Kt.AtA_A[c]); D[c] += mu_eq * (c_model.a_per ? ld(Pk, TR_ATA + 15 + c) : Kt.AtA_D[c]);
      p[c] += ld(Pk, TR_ATY + c) - mu_eq * ld(Pk, TR_ATB + c);
    }
#pragma unroll
    for (int c = 0; c < 9; ++c) B[c] += mu_eq * (c_model.a_per ? ld(Pk, TR_ATA + 6 + c) : Kt.AtA_B[c]);
  }
  for (int n = 0; n < J.npin; ++n) {
    const double* Pp = pend_blk(Td, O, J.pin[n]);
#pragma unroll
    for (int c = 0; c < 6; ++c) { A[c] += ld(Pp, PR_H + c); D[c] += ld(Pp, PR_H + 15 + c); p[c] += ld(Pp, PR_H + 21 + c); }
#pragma unroll
    for (int c = 0; c < 9; ++c) B[c] += ld(Pp, PR_H + 6 + c);
  }
  // U = H S: the selected columns of H (read in place below), or (DS: S = [0; E(q)], K = 3) the angular columns of H times E
  double E[DS ? 9 : 1], Ud[DS ? 6 : 1][K];
  if (DS) {
#pragma unroll
    for (int c = 0; c < 9; ++c) E[c] = ld(migrate ? Pfs : Pf, FR_S + c);
#pragma unroll
    for (int a = 0; a < 6; ++a)
#pragma unroll
      for (int k = 0; k < K; ++k) Ud[a][k] = Hel(A, B, D, a, 3) * E[k] + Hel(A, B, D, a, 4) * E[3 + k] + Hel(A, B, D, a, 5) * E[6 + k];
    if (migrate)
      for (int c = 0; c < 9; ++c) st(Pf, FR_S + c, E[c]);
  }
  double Mx[K * K], Dinv[K * (K + 1) / 2], r[K];
#pragma unroll
  for (int a = 0; a < K; ++a)
#pragma unroll
    for (int b = 0; b < K; ++b) {  // S^T U
      if (DS) Mx[K * a + b] = E[a] * Ud[3][b] + E[3 + a] * Ud[4][b] + E[6 + a] * Ud[5][b];
      else Mx[K * a + b] = Hel(A, B, D, sel[a], sel[b]);
    }
#pragma unroll
  for (int c = 0; c < K; ++c) Mx[(K + 1) * c] += mu;  // armature R = mu_ineq (hxx:294-295)
  spd_inverse<K>(Mx, Dinv);
#pragma unroll
  for (int c = 0; c < K; ++c) {  // r = w - mu z (:296) + S^T p (:70)
    if (DS) r[c] = (w[c] - mu * z[c]) + (E[c] * p[3] + E[3 + c] * p[4] + E[6 + c] * p[5]);
    else r[c] = (w[c] - mu * z[c]) + p[sel[c]];
  }
#pragma unroll
  for (int c = 0; c < 6; ++c) { st(Pj, JR_H + c, A[c]); st(Pj, JR_H + 15 + c, D[c]); st(Pj, JR_P + c, p[c]); }
#pragma unroll
  for (int c = 0; c < 9; ++c) st(Pj, JR_H + 6 + c, B[c]);
#pragma unroll
  for (int c = 0; c < K; ++c) st(Pf, FR_R + c, r[c]);
#pragma unroll
  for (int c = 0; c < K * (K + 1) / 2; ++c) st(Pf, FR_DINV + c, Dinv[c]);
  if (J.parent > 0) {
    // UDinv = U Dinv; H -= UDinv U^T (:63); p -= UDinv r (:71-73)
    double UD[6][K];
#pragma unroll
    for (int a = 0; a < 6; ++a)
#pragma unroll
      for (int k = 0; k < K; ++k) {
        double sum = 0.0;
#pragma unroll
        for (int l = 0; l < K; ++l) sum += (DS ? Ud[a][l] : Hel(A, B, D, a, sel[l])) * Dinv[sk<K>(l, k)];
        UD[a][k] = sum;
        st(Pf, FR_UD + K * a + k, sum);
      }
    double U[6][K];
#pragma unroll
    for (int a = 0; a < 6; ++a)
#pragma unroll
      for (int k = 0; k < K; ++k) U[a][k] = DS ? Ud[a][k] : Hel(A, B, D, a, sel[k]);
#pragma unroll
    for (int a = 0; a < 6; ++a)
#pragma unroll
      for (int b = a; b < 6; ++b) {
        if (a < 3 && b >= 3) continue;  // the LA block is general: done below
        double sum = 0.0;
#pragma unroll
        for (int k = 0; k < K; ++k) sum += UD[a][k] * U[b][k];
        Hsub(A, B, D, a, b, sum);
      }
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 3; b < 6; ++b) {
        double sum = 0.0;
#pragma unroll
        for (int k = 0; k < K; ++k) sum += UD[a][k] * U[b][k];
        B[3 * a + (b - 3)] -= sum;
      }
#pragma unroll
    for (int a = 0; a < 6; ++a) {
      double sum = 0.0;
#pragma unroll
      for (int k = 0; k < K; ++k) sum += UD[a][k] * r[k];
      p[a] -= sum;
    }
    double R[9], t[3], cA[6], cB[9], cD[6], cp[6];
    md_load_xf(migrate ? Pfs : Pf, R, t);
    congruence(R, t, A, B, D, cA, cB, cD);
    act_force(R, t, p, cp);
    double* Pp = pend_blk(Td, O, J.pout);
#pragma unroll
    for (int c = 0; c < 6; ++c) { st(Pp, PR_H + c, cA[c]); st(Pp, PR_H + 15 + c, cD[c]); st(Pp, PR_H + 21 + c, cp[c]); }
#pragma unroll
    for (int c = 0; c < 9; ++c) st(Pp, PR_H + 6 + c, cB[c]);
  }
}

template <bool DEBUG, int K, bool DS = false>
LOIK_DEV_CALL void md_forward(const ModelC& c_model, const double* Ts, double* Td, const double mu, const double mu_eq, Carry& cy, const int i, double* Dg) {
  const Offs& O = c_model.off;
  const JointC& J = c_model.j[i];
  const int nb = c_model.nb;
  int sel[K];
#pragma unroll
  for (int k = 0; k < K; ++k) sel[k] = (J.sel0 >> (4 * k)) & 7;
  double* Pj = joint_blk(Td, O, i - 1);
  double* Pf = md_blk(Td, O, J.mblk);
  const double* Pjs = joint_blk(const_cast<double*>(Ts), O, i - 1);
  const double* Pfs = md_blk(const_cast<double*>(Ts), O, J.mblk);
  const double inv_mu = 1.0 / mu;
  double Dinv[K * (K + 1) / 2], r[K], A[6], B[9], D[6], p[6], vold[6], fold[6], v[6], f[6], nu[K];
#pragma unroll
  for (int c = 0; c < K * (K + 1) / 2; ++c) Dinv[c] = ld(Pf, FR_DINV + c);
#pragma unroll
  for (int c = 0; c < K; ++c) r[c] = ld(Pf, FR_R + c);
#pragma unroll
  for (int c = 0; c < 6; ++c) {
    A[c] = ld(Pj, JR_H + c); D[c] = ld(Pj, JR_H + 15 + c); p[c] = ld(Pj, JR_P + c);
    vold[c] = ld(Pjs, JR_V + c); fold[c] = ld(Pjs, JR_F + c);
  }
#pragma unroll
  for (int c = 0; c < 9; ++c) B[c] = ld(Pj, JR_H + 6 + c);
  if (J.parent > 0) {  // vi_parent = liMi.actInv(v_parent) (:125)
    double R[9], t[3], vin[6];
    md_load_xf(Pf, R, t);
    const double* Pq = joint_blk(Td, O, J.parent - 1);
#pragma unroll
    for (int c = 0; c < 6; ++c) vin[c] = ld(Pq, JR_V + c);
    actinv_motion(R, t, vin, v);
  } else {
#pragma unroll
    for (int c = 0; c < 6; ++c) v[c] = 0.0;
  }
#pragma unroll
  for (int k = 0; k < K; ++k) {  // nu = -UDinv^T vp - Dinv r (:127)
    double acc = 0.0;
    if (J.parent > 0) {
#pragma unroll
      for (int a = 0; a < 6; ++a) acc += ld(Pf, FR_UD + K * a + k) * v[a];
    }
    double dr = 0.0;
#pragma unroll
    for (int l = 0; l < K; ++l) dr += Dinv[sk<K>(k, l)] * r[l];
    nu[k] = -acc - dr;
    cy.nu_inf = amax(cy.nu_inf, nu[k]);
  }
#pragma unroll
  for (int k = 0; k < K; ++k) {  // v_i = vp + S nu (:133-134)
    if (DS) {
#pragma unroll
      for (int a = 0; a < 3; ++a) v[3 + a] += ld(Pf, FR_S + 3 * a + k) * nu[k];
    } else {
      v[sel[k]] += nu[k];
    }
  }
#pragma unroll
  for (int c = 0; c < 6; ++c) cy.dvis_inf = amax(cy.dvis_inf, v[c] - vold[c]);
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    f[a] = A[si(a, 0)] * v[0] + A[si(a, 1)] * v[1] + A[si(a, 2)] * v[2] + B[3 * a] * v[3] + B[3 * a + 1] * v[4] + B[3 * a + 2] * v[5] + p[a];
    f[3 + a] = B[a] * v[0] + B[3 + a] * v[1] + B[6 + a] * v[2] + D[si(a, 0)] * v[3] + D[si(a, 1)] * v[4] + D[si(a, 2)] * v[5] + p[3 + a];
  }
#pragma unroll
  for (int c = 0; c < 6; ++c) {
    cy.dfis_inf = amax(cy.dfis_inf, f[c] - fold[c]);
    st(Pj, JR_V + c, v[c]); st(Pj, JR_F + c, f[c]);
  }
#pragma unroll
  for (int c = 0; c < K; ++c) {
    const double lb = ld(Pf, FR_LB + c), ub = ld(Pf, FR_UB + c);  // (rows of the md block: per instance, or the shared bounds replicated by k_set_bounds)
    const double w_old = ld(Pfs, FR_W + c);
    cy.dnu_inf = amax(cy.dnu_inf, nu[c] - ld(Pfs, FR_NU + c));
    const double z = dmin(ub, dmax(lb, nu[c] + inv_mu * w_old));
    cy.dz_inf = amax(cy.dz_inf, z - ld(Pfs, FR_Z + c));
    const double rp = nu[c] - z;
    cy.pres_slack = amax(cy.pres_slack, rp);
    const double dw = mu * rp;
    cy.dw_inf = amax(cy.dw_inf, dw);
    cy.ubdw_p += ub * dmax(dw, 0.0);
    cy.lbdw_m += lb * dmin(dw, 0.0);
    st(Pf, FR_NU + c, nu[c]); st(Pf, FR_Z + c, z); st(Pf, FR_W + c, w_old + dw);
    if (DEBUG && Dg) st(Dg, O.prv + 6 * nb + J.idxv + c, rp);
  }
  if (J.task >= 0) {  // DualUpdate for a task on this joint (:410-451)
    const TaskC& Kt = c_model.t[J.task];
    double* Pk = task_blk(Td, O, J.task);
    const double* Pks = task_blk(const_cast<double*>(Ts), O, J.task);
    double y[6], plus = 0.0, minus = 0.0;
#pragma unroll
    for (int a = 0; a < 6; ++a) {
      const double Av = task_A(c_model, Kt, Pk, 6 * a) * v[0] + task_A(c_model, Kt, Pk, 6 * a + 1) * v[1] + task_A(c_model, Kt, Pk, 6 * a + 2) * v[2] +
                        task_A(c_model, Kt, Pk, 6 * a + 3) * v[3] + task_A(c_model, Kt, Pk, 6 * a + 4) * v[4] + task_A(c_model, Kt, Pk, 6 * a + 5) * v[5];
      const double bi = ld(Pk, TR_B + a), e = Av - bi, dy = mu_eq * e;
      y[a] = ld(Pks, TR_Y + a) + dy;
      cy.dyis_inf = amax(cy.dyis_inf, dy); cy.Av_inf = amax(cy.Av_inf, Av); cy.pres_task = amax(cy.pres_task, e);
      plus += bi * dmax(dy, 0.0); minus += bi * dmin(dy, 0.0);
      if (DEBUG && Dg) st(Dg, O.prv + 6 * (i - 1) + a, e);
    }
    cy.bTdy_p += plus; cy.bTdy_m += minus;
#pragma unroll
    for (int a = 0; a < 6; ++a) {
      st(Pk, TR_Y + a, y[a]);
      st(Pk, TR_ATY + a, task_A(c_model, Kt, Pk, a) * y[0] + task_A(c_model, Kt, Pk, 6 + a) * y[1] + task_A(c_model, Kt, Pk, 12 + a) * y[2] +
                             task_A(c_model, Kt, Pk, 18 + a) * y[3] + task_A(c_model, Kt, Pk, 24 + a) * y[4] + task_A(c_model, Kt, Pk, 30 + a) * y[5]);
    }
  }
}

template <bool DEBUG, int K, bool DS = false>
LOIK_DEV_CALL void md_residual(const ModelC& c_model, const double* Ts, double* Td, Resid& rs, const int i, double* Dg) {
  const Offs& O = c_model.off;
  const JointC& J = c_model.j[i];
    const HrefC& Hr = c_model.href[J.href];
  const int nb = c_model.nb;
  int sel[K];
#pragma unroll
  for (int k = 0; k < K; ++k) sel[k] = (J.sel0 >> (4 * k)) & 7;
  double* Pj = joint_blk(Td, O, i - 1);
  double* Pf = md_blk(Td, O, J.mblk);
  const double* Pjs = joint_blk(const_cast<double*>(Ts), O, i - 1);
  const double* Pfs = md_blk(const_cast<double*>(Ts), O, J.mblk);
  double f[6], v[6], F[6], Fold[6];
#pragma unroll
  for (int c = 0; c < 6; ++c) { f[c] = ld(Pj, JR_F + c); v[c] = ld(Pj, JR_V + c); Fold[c] = ld(Pjs, JR_FD + c); F[c] = 0.0; }
  if (J.task >= 0) {
    const double* Pk = task_blk(Td, O, J.task);
#pragma unroll
    for (int c = 0; c < 6; ++c) F[c] = ld(Pk, TR_ATY + c);
  }
  for (int n = 0; n < J.npin; ++n) {
    const double* Pp = pend_blk(Td, O, J.pin[n]);
#pragma unroll
    for (int c = 0; c < 6; ++c) F[c] += ld(Pp, PR_F + c);
  }
  double Hrv[6];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    Hrv[a] = Hr.A[si(a, 0)] * v[0] + Hr.A[si(a, 1)] * v[1] + Hr.A[si(a, 2)] * v[2] + Hr.B[3 * a] * v[3] + Hr.B[3 * a + 1] * v[4] + Hr.B[3 * a + 2] * v[5];
    Hrv[3 + a] = Hr.B[a] * v[0] + Hr.B[3 + a] * v[1] + Hr.B[6 + a] * v[2] + Hr.D[si(a, 0)] * v[3] + Hr.D[si(a, 1)] * v[4] + Hr.D[si(a, 2)] * v[5];
  }
  if (c_model.href_per) href_times(Pj, v, Hrv);
#pragma unroll
  for (int c = 0; c < 6; ++c) {
    F[c] += -f[c];
    rs.dF_inf = amax(rs.dF_inf, F[c] - Fold[c]);
    rs.F_inf = amax(rs.F_inf, F[c]);
    rs.Hrefv_inf = amax(rs.Hrefv_inf, Hrv[c]);
    const double hv = hv_of<true>(c_model, Hr, Pj, c);
    if (c_model.vref_per) rs.Hrefv_inf = amax(rs.Hrefv_inf, hv);  // (this instance's |Hv|inf, see sweep_residual)
    const double rd = Hrv[c] - hv + F[c];
    rs.dres_v = amax(rs.dres_v, rd);
    st(Pj, JR_FD + c, F[c]);
    if (DEBUG && Dg) st(Dg, O.drv + 6 * (i - 1) + c, rd);
  }
#pragma unroll
  for (int c = 0; c < K; ++c) {
    const double Stf = DS ? ld(Pf, FR_S + c) * f[3] + ld(Pf, FR_S + 3 + c) * f[4] + ld(Pf, FR_S + 6 + c) * f[5] : f[sel[c]];
    const double Tn = Stf + ld(Pf, FR_W + c);  // S^T f + w (:231)
    rs.T_inf = amax(rs.T_inf, Tn);
    rs.dT_inf = amax(rs.dT_inf, Tn - ld(Pfs, FR_T + c));
    st(Pf, FR_T + c, Tn);
    if (DEBUG && Dg) st(Dg, O.drv + 6 * nb + J.idxv + c, Tn);
  }
  if (J.parent > 0) {  // fis_diff_plus_Aty[parent] += liMi.act(f_i) (:212)
    double R[9], t[3], cF[6];
    md_load_xf(Pf, R, t);
    act_force(R, t, f, cF);
    double* Pp = pend_blk(Td, O, J.pout);
#pragma unroll
    for (int c = 0; c < 6; ++c) st(Pp, PR_F + c, cF[c]);
  }
}

// The three sweeps over the joints lo..hi, multi-DoF joints included (lo == hi for those).  The running norms travel
// to the out-of-line multi-DoF steps through a local copy, so that they stay in registers in the joint loops.
LOIK_DEV void span_backward(const ModelC& c_model, const double* Ts, double* Td, const double mu, const double mu_eq, const int lo,
                            const int hi, const bool migrate) {
  const int k = c_model.j[lo].nvj;
  if (k == 1) sweep_backward<true>(c_model, Ts, Td, mu, mu_eq, lo, hi, migrate);
  else if (k == 3 && c_model.j[lo].jtype == LOIK_JOINT_SPHERICAL_ZYX) md_backward<3, true>(c_model, Ts, Td, mu, mu_eq, lo, migrate);
  else if (k == 3) md_backward<3>(c_model, Ts, Td, mu, mu_eq, lo, migrate);
  else md_backward<6>(c_model, Ts, Td, mu, mu_eq, lo, migrate);
}
template <bool DEBUG>
LOIK_DEV void span_forward(const ModelC& c_model, const double* Ts, double* Td, const double mu, const double mu_eq, Carry& cy,
                           const int lo, const int hi, const bool drop_ws = false, double* Dg = nullptr) {
  const int k = c_model.j[lo].nvj;
  if (k == 1) { sweep_forward<DEBUG>(c_model, Ts, Td, mu, mu_eq, cy, lo, hi, drop_ws, Dg); return; }
  Carry tmp = cy;
  if (k == 3 && c_model.j[lo].jtype == LOIK_JOINT_SPHERICAL_ZYX) md_forward<DEBUG, 3, true>(c_model, Ts, Td, mu, mu_eq, tmp, lo, Dg);
  else if (k == 3) md_forward<DEBUG, 3>(c_model, Ts, Td, mu, mu_eq, tmp, lo, Dg);
  else md_forward<DEBUG, 6>(c_model, Ts, Td, mu, mu_eq, tmp, lo, Dg);
  cy = tmp;
}
template <bool DEBUG>
LOIK_DEV void span_residual(const ModelC& c_model, const double* Ts, double* Td, Resid& rs, const int lo, const int hi, double* Dg = nullptr) {
  const int k = c_model.j[lo].nvj;
  if (k == 1) { sweep_residual<DEBUG, true>(c_model, Ts, Td, rs, lo, hi, Dg); return; }
  Resid tmp = rs;
  if (k == 3 && c_model.j[lo].jtype == LOIK_JOINT_SPHERICAL_ZYX) md_residual<DEBUG, 3, true>(c_model, Ts, Td, tmp, lo, Dg);
  else if (k == 3) md_residual<DEBUG, 3>(c_model, Ts, Td, tmp, lo, Dg);
  else md_residual<DEBUG, 6>(c_model, Ts, Td, tmp, lo, Dg);
  rs = tmp;
}

// ---------------------------------------------------------------------------------------------
// CheckConvergence (hxx:540-565) + CheckFeasibility (:572-606) + UpdateMu (:613-641) + the loop control
// of Solve() (hpp:377-454) and InfeasibilityTailSolve() (hpp:271-319), per instance.
// `fixed`: stopping disabled (throughput mode).  Returns the new status; updates mu.
// ---------------------------------------------------------------------------------------------
// decide_core: the decisions alone (no memory access), shared by every iteration kernel; decide<DEBUG>: + the stores
// into a tile record (`writer`: this thread stores the per-instance results -- one warp per tile does when several
// warps share it).
struct Verdict {
  double pres, dres, tol_p, tol_d;                   // primal / dual residual, their tolerances (tol_*: only if has_tol)
  double dyqp, ATdy, ubp, lbm, c1, c2, dx;           // CheckFeasibility's scalars (debug rows)
  bool has_tol;
};
LOIK_DEV int decide_core(const ModelC& M, const int status, const int it, const bool fixed, const Carry& cy, const Resid& rs,
                         const double binf, double& mu, Verdict& V) {
  V.pres = dmax(cy.pres_task, cy.pres_slack);  // (:498)
  V.dres = dmax(rs.dres_v, rs.T_inf);          // (:517); dual_residual_vec[6nb:] = Stf_plus_w (:484)
  const double pres = V.pres, dres = V.dres;
  int ns = status;
  V.dyqp = V.ATdy = V.ubp = V.lbm = V.c1 = V.c2 = 0.0;
  V.tol_p = V.tol_d = 0.0;
  V.has_tol = status == ST_RUNNING;
  const double dx = dmax(cy.dvis_inf, cy.dnu_inf);
  V.dx = dx;
  if (status == ST_RUNNING) {
    const double tol_p = M.tol_abs + M.tol_rel * dmax(dmax(cy.Av_inf, cy.nu_inf), dmax(binf, cy.nu_inf));            // (:544-546)
    const double tol_d = M.tol_abs + M.tol_rel * dmax(dmax(rs.Hrefv_inf, dmax(rs.F_inf, rs.T_inf)), M.Hv_inf);      // (:548-552)
    V.tol_p = tol_p; V.tol_d = tol_d;
    const bool converged = (pres < tol_p) && (dres < tol_d);                                                        // (:555)
    bool infeasible = false;
    if (it > 1) {                                                                                                   // (hpp:425-427)
      V.dyqp = dmax(cy.dfis_inf, dmax(cy.dyis_inf, cy.dw_inf));                                                     // (:576-578)
      V.ATdy = dmax(rs.dF_inf, rs.dT_inf);                                                                          // (:580-581)
      const bool cond1 = V.ATdy <= M.tol_pinf * V.dyqp;                                                             // (:583-584)
      V.ubp = cy.bTdy_p + cy.ubdw_p;                                                                                // (:587-588)
      V.lbm = cy.bTdy_m + cy.lbdw_m;                                                                                // (:589-590)
      const bool cond2 = (V.ubp + V.lbm) <= M.tol_pinf * V.dyqp;                                                    // (:592-593)
      infeasible = cond1 && cond2;
      V.c1 = cond1; V.c2 = cond2;
    }
    if (fixed) {
      if (pres > 10 * dres) mu *= 10; else if (dres > 10 * pres) mu *= 0.1;
    } else if (converged) {
      ns = infeasible ? ST_CONVERGED_PINF : ST_CONVERGED;
    } else if (infeasible) {
      // entering InfeasibilityTailSolve: first evaluation of its while-condition (hpp:275-284)
      if (dx >= M.tol_tail || cy.dz_inf >= M.tol_tail) ns = (it >= M.max_iter) ? ST_INFEASIBLE_DONE : ST_TAIL;
      else ns = ST_INFEASIBLE_DONE;
    } else {
      if (pres > 10 * dres) mu *= 10; else if (dres > 10 * pres) mu *= 0.1;                                         // (:617-628)
      if (it >= M.max_iter - 1) ns = ST_MAXITER;                                                                    // loop bound (hpp:377)
    }
  } else {  // ST_TAIL: one tail iteration just ran (hpp:286-308); re-evaluate the while-condition
    if (dx >= M.tol_tail || cy.dz_inf >= M.tol_tail) ns = (it >= M.max_iter) ? ST_INFEASIBLE_DONE : ST_TAIL;
    else ns = ST_INFEASIBLE_DONE;
  }
  return ns;
}
template <bool DEBUG>
LOIK_DEV int decide(const ModelC& c_model, double* __restrict__ T, const int status, const int it, const bool fixed, const Carry& cy,
                    const Resid& rs, double& mu, const bool writer = true) {
  const ModelC& M = c_model;
  double* G = glob_blk(T, M.off);
  Verdict V;
  const double binf = status == ST_RUNNING ? ld(G, GR_BINF) : 0.0;
  const int ns = decide_core(M, status, it, fixed, cy, rs, binf, mu, V);
  if (writer) {
    st(G, GR_RES + 0, V.pres);
    st(G, GR_RES + 1, V.dres);
    if (V.has_tol) {
      st(G, GR_RES + 2, V.tol_p);
      st(G, GR_RES + 3, V.tol_d);
    }
  }
  if (DEBUG && writer) {
    const int N = GR_NORMS;
    st(G, N + 0, cy.bTdy_p); st(G, N + 1, cy.bTdy_m); st(G, N + 2, cy.Av_inf); st(G, N + 3, cy.nu_inf);
    st(G, N + 4, rs.Hrefv_inf); st(G, N + 5, rs.F_inf); st(G, N + 6, rs.T_inf); st(G, N + 7, rs.dF_inf);
    st(G, N + 8, rs.dT_inf); st(G, N + 9, cy.dvis_inf); st(G, N + 10, cy.dnu_inf); st(G, N + 11, cy.dz_inf);
    st(G, N + 12, cy.dfis_inf); st(G, N + 13, cy.dyis_inf); st(G, N + 14, cy.dw_inf);
    st(G, N + 15, cy.pres_task); st(G, N + 16, cy.pres_slack); st(G, N + 17, rs.dres_v); st(G, N + 18, rs.T_inf);
    if (status == ST_RUNNING && it > 1) {
      st(G, N + 19, V.dyqp); st(G, N + 20, V.ATdy); st(G, N + 21, V.ubp); st(G, N + 22, V.lbm);
      st(G, N + 23, V.c1); st(G, N + 24, V.c2);
    }
    if (status == ST_TAIL || it > 1) st(G, N + 25, V.dx);
  }
  return ns;
}

// ---------------------------------------------------------------------------------------------
// The reference's public per-step methods one by one (loik-loid-optimized.hpp:192-264), for tests that drive the
// solver the way tests/loik-loid.cpp:340-478 does.  They work on the home arena, ignore the loop-control status and
// keep the reference's running norms in the GR_NORMS rows (index = loik_norm_index).  Same device maths as the
// fused sweeps; the production path never calls them.
// ---------------------------------------------------------------------------------------------
enum : int { N_BTDY_P = 0, N_BTDY_M, N_AV, N_NU, N_HREFV, N_F, N_T, N_DF, N_DT, N_DVIS, N_DNU, N_DZ, N_DFIS, N_DYIS, N_DW,
             N_PRES_TASK, N_PRES_SLACK, N_DRES_V, N_DRES_NU, N_DYQP, N_ATDY, N_UBP, N_LBM, N_C1, N_C2, N_DX, N_CONVERGED,
             N_PINFEASIBLE };

// ResetInfNorms (data hxx:165-182)
LOIK_DEV void fine_reset_inf_norms(const ModelC& M, double* __restrict__ T) {
  double* G = glob_blk(T, M.off);
  for (int k = 0; k <= N_DW; ++k) st(G, GR_NORMS + k, 0.0);
  st(G, GR_CARRY + 10, 0.0);  // ub^T max(delta_w, 0)
  st(G, GR_CARRY + 11, 0.0);  // lb^T min(delta_w, 0)
}
// FwdPass1 (hxx:290-338): His, pis (rows JR_H / JR_P), r = w - mu_ineq z (row JR_R); R = mu_ineq is implicit
LOIK_DEV void fine_fwdpass1(const ModelC& M, double* __restrict__ T, const double mu, const double mu_eq) {
  const Offs& O = M.off;
  for (int i = 1; i <= M.nb; ++i) {
    const JointC& J = M.j[i];
    const HrefC& Hr = M.href[J.href];
    double* Pj = joint_blk(T, O, i - 1);
    double A[6], B[9], D[6], p[6];
    for (int c = 0; c < 6; ++c) { p[c] = -M.rho * ld(Pj, JR_V + c) - hv_of<true>(M, Hr, Pj, c); A[c] = Hr.A[c]; D[c] = Hr.D[c]; }
    for (int c = 0; c < 9; ++c) B[c] = Hr.B[c];
    if (M.href_per) href_rows(Pj, A, B, D);
    A[0] += M.rho; A[3] += M.rho; A[5] += M.rho; D[0] += M.rho; D[3] += M.rho; D[5] += M.rho;
    if (J.task >= 0) {
      const TaskC& K = M.t[J.task];
      const double* Pk = task_blk(T, O, J.task);
      for (int c = 0; c < 6; ++c) {
        A[c] += mu_eq * (M.a_per ? ld(Pk, TR_ATA + c) : K.AtA_A[c]); D[c] += mu_eq * (M.a_per ? ld(Pk, TR_ATA + 15 + c) : K.AtA_D[c]);
        p[c] += ld(Pk, TR_ATY + c) - mu_eq * ld(Pk, TR_ATB + c);
      }
      for (int c = 0; c < 9; ++c) B[c] += mu_eq * (M.a_per ? ld(Pk, TR_ATA + 6 + c) : K.AtA_B[c]);
    }
    const double r0 = ld(Pj, JR_W) - mu * ld(Pj, JR_Z);
    for (int c = 0; c < 6; ++c) { st(Pj, JR_H + c, A[c]); st(Pj, JR_H + 15 + c, D[c]); st(Pj, JR_P + c, p[c]); }
    for (int c = 0; c < 9; ++c) st(Pj, JR_H + 6 + c, B[c]);
    st(Pj, JR_R, r0);
  }
}
// FwdPass2OptimizedVisitor (hxx:361-377, algo :102-163)
LOIK_DEV void fine_fwdpass2(const ModelC& M, double* __restrict__ T) {
  const Offs& O = M.off;
  double* G = glob_blk(T, O);
  double nu_inf = ld(G, GR_NORMS + N_NU), dfis = ld(G, GR_NORMS + N_DFIS), hrefv = ld(G, GR_NORMS + N_HREFV), dvis = ld(G, GR_NORMS + N_DVIS);
  double dnu = 0.0;
  for (int i = 1; i <= M.nb; ++i) {
    const JointC& J = M.j[i];
    const HrefC& Hr = M.href[J.href];
    double* Pj = joint_blk(T, O, i - 1);
    double vin[6], R[9], t[3], v[6];
    for (int c = 0; c < 6; ++c) vin[c] = J.parent == 0 ? 0.0 : ld(joint_blk(T, O, J.parent - 1), JR_V + c);
    make_xf(J, ld(Pj, JR_JQ), ld(Pj, JR_JQ + 1), R, t);
    actinv_motion(R, t, vin, v);
    double acc = 0.0;
    for (int c = 0; c < 6; ++c) acc += ld(Pj, JR_UD + c) * v[c];
    const double nu = -acc - ld(Pj, JR_DINV) * ld(Pj, JR_R);
    nu_inf = amax(nu_inf, nu);
    S_axpy(J, nu, v);
    double A[6], B[9], D[6];
    for (int c = 0; c < 6; ++c) { A[c] = ld(Pj, JR_H + c); D[c] = ld(Pj, JR_H + 15 + c); }
    for (int c = 0; c < 9; ++c) B[c] = ld(Pj, JR_H + 6 + c);
    double f[6], Hrv[6];
    for (int a = 0; a < 3; ++a) {
      f[a] = A[si(a, 0)] * v[0] + A[si(a, 1)] * v[1] + A[si(a, 2)] * v[2] + B[3 * a] * v[3] + B[3 * a + 1] * v[4] + B[3 * a + 2] * v[5] + ld(Pj, JR_P + a);
      f[3 + a] = B[a] * v[0] + B[3 + a] * v[1] + B[6 + a] * v[2] + D[si(a, 0)] * v[3] + D[si(a, 1)] * v[4] + D[si(a, 2)] * v[5] + ld(Pj, JR_P + 3 + a);
      Hrv[a] = Hr.A[si(a, 0)] * v[0] + Hr.A[si(a, 1)] * v[1] + Hr.A[si(a, 2)] * v[2] + Hr.B[3 * a] * v[3] + Hr.B[3 * a + 1] * v[4] + Hr.B[3 * a + 2] * v[5];
      Hrv[3 + a] = Hr.B[a] * v[0] + Hr.B[3 + a] * v[1] + Hr.B[6 + a] * v[2] + Hr.D[si(a, 0)] * v[3] + Hr.D[si(a, 1)] * v[4] + Hr.D[si(a, 2)] * v[5];
    }
    if (M.href_per) href_times(Pj, v, Hrv);
    for (int c = 0; c < 6; ++c) {
      dvis = amax(dvis, v[c] - ld(Pj, JR_V + c));
      dfis = amax(dfis, f[c] - ld(Pj, JR_F + c));
      hrefv = amax(hrefv, Hrv[c]);
    }
    dnu = amax(dnu, nu - ld(Pj, JR_NU));
    for (int c = 0; c < 6; ++c) { st(Pj, JR_V + c, v[c]); st(Pj, JR_F + c, f[c]); }
    st(Pj, JR_NU, nu);
  }
  st(G, GR_NORMS + N_NU, nu_inf); st(G, GR_NORMS + N_DFIS, dfis); st(G, GR_NORMS + N_HREFV, hrefv);
  st(G, GR_NORMS + N_DVIS, dvis); st(G, GR_NORMS + N_DNU, dnu);
}
// BoxProj (hxx:384-397)
LOIK_DEV void fine_boxproj(const ModelC& M, double* __restrict__ T, const double mu, double* Dg) {
  const Offs& O = M.off;
  double* G = glob_blk(T, O);
  double dz = 0.0, slack = 0.0;
  for (int i = 1; i <= M.nb; ++i) {
    const JointC& J = M.j[i];
    double* Pj = joint_blk(T, O, i - 1);
    const double lb = M.bounds_per_instance ? ld(Pj, JR_LB) : J.lb, ub = M.bounds_per_instance ? ld(Pj, JR_UB) : J.ub;
    const double nu = ld(Pj, JR_NU);
    const double z = dmin(ub, dmax(lb, nu + (1.0 / mu) * ld(Pj, JR_W)));
    dz = amax(dz, z - ld(Pj, JR_Z));
    slack = amax(slack, nu - z);
    st(Pj, JR_Z, z);
    if (Dg) st(Dg, O.prv + 6 * M.nb + J.idxv, nu - z);
  }
  st(G, GR_NORMS + N_DZ, dz);
  st(G, GR_NORMS + N_PRES_SLACK, slack);
}
// DualUpdate (hxx:404-461)
LOIK_DEV void fine_dualupdate(const ModelC& M, double* __restrict__ T, const double mu, const double mu_eq, double* Dg) {
  const Offs& O = M.off;
  double* G = glob_blk(T, O);
  double dyis = ld(G, GR_NORMS + N_DYIS), Av_inf = ld(G, GR_NORMS + N_AV), bp = ld(G, GR_NORMS + N_BTDY_P), bm = ld(G, GR_NORMS + N_BTDY_M);
  double ptask = 0.0;
  for (int k = 0; k < M.nc; ++k) {
    const TaskC& K = M.t[k];
    double* Pk = task_blk(T, O, k);
    const double* Pj = joint_blk(T, O, M.task_joint[k] - 1);
    double v[6], y[6], plus = 0.0, minus = 0.0;
    for (int c = 0; c < 6; ++c) v[c] = ld(Pj, JR_V + c);
    for (int a = 0; a < 6; ++a) {
      const double Av = task_A(M, K, Pk, 6 * a) * v[0] + task_A(M, K, Pk, 6 * a + 1) * v[1] + task_A(M, K, Pk, 6 * a + 2) * v[2] +
                        task_A(M, K, Pk, 6 * a + 3) * v[3] + task_A(M, K, Pk, 6 * a + 4) * v[4] + task_A(M, K, Pk, 6 * a + 5) * v[5];
      const double bi = ld(Pk, TR_B + a), e = Av - bi, dy = mu_eq * e;
      y[a] = ld(Pk, TR_Y + a) + dy;
      dyis = amax(dyis, dy); Av_inf = amax(Av_inf, Av); ptask = amax(ptask, e);
      plus += bi * dmax(dy, 0.0); minus += bi * dmin(dy, 0.0);
      if (Dg) st(Dg, O.prv + 6 * (M.task_joint[k] - 1) + a, e);
    }
    bp += plus; bm += minus;
    for (int a = 0; a < 6; ++a) {
      st(Pk, TR_Y + a, y[a]);
      st(Pk, TR_ATY + a, task_A(M, K, Pk, a) * y[0] + task_A(M, K, Pk, 6 + a) * y[1] + task_A(M, K, Pk, 12 + a) * y[2] + task_A(M, K, Pk, 18 + a) * y[3] +
                             task_A(M, K, Pk, 24 + a) * y[4] + task_A(M, K, Pk, 30 + a) * y[5]);
    }
  }
  double dw_inf = 0.0, ubdw = 0.0, lbdw = 0.0;
  for (int i = 1; i <= M.nb; ++i) {
    const JointC& J = M.j[i];
    double* Pj = joint_blk(T, O, i - 1);
    const double lb = M.bounds_per_instance ? ld(Pj, JR_LB) : J.lb, ub = M.bounds_per_instance ? ld(Pj, JR_UB) : J.ub;
    const double dw = mu * (ld(Pj, JR_NU) - ld(Pj, JR_Z));
    st(Pj, JR_W, ld(Pj, JR_W) + dw);
    dw_inf = amax(dw_inf, dw);
    ubdw += ub * dmax(dw, 0.0); lbdw += lb * dmin(dw, 0.0);
  }
  st(G, GR_NORMS + N_DYIS, dyis); st(G, GR_NORMS + N_AV, Av_inf); st(G, GR_NORMS + N_BTDY_P, bp); st(G, GR_NORMS + N_BTDY_M, bm);
  st(G, GR_NORMS + N_PRES_TASK, ptask); st(G, GR_NORMS + N_DW, dw_inf);
  st(G, GR_CARRY + 10, ubdw); st(G, GR_CARRY + 11, lbdw);
}
// ComputeResiduals (hxx:529-533)
LOIK_DEV void fine_compute_residuals(const ModelC& M, double* __restrict__ T, double* Dg) {
  double* G = glob_blk(T, M.off);
  st(G, GR_RES + 0, dmax(ld(G, GR_NORMS + N_PRES_TASK), ld(G, GR_NORMS + N_PRES_SLACK)));
  Resid rs;
  zero(rs);
  sweep_residual<true, true>(M, T, T, rs, 1, M.nb, Dg);
  st(G, GR_NORMS + N_F, rs.F_inf); st(G, GR_NORMS + N_T, rs.T_inf); st(G, GR_NORMS + N_DF, rs.dF_inf); st(G, GR_NORMS + N_DT, rs.dT_inf);
  st(G, GR_NORMS + N_DRES_V, rs.dres_v); st(G, GR_NORMS + N_DRES_NU, rs.T_inf);
  st(G, GR_RES + 1, dmax(rs.dres_v, rs.T_inf));
}
// CheckConvergence (hxx:540-565)
LOIK_DEV void fine_check_convergence(const ModelC& M, double* __restrict__ T) {
  double* G = glob_blk(T, M.off);
  const double nu_inf = ld(G, GR_NORMS + N_NU);
  const double tol_p = M.tol_abs + M.tol_rel * dmax(dmax(ld(G, GR_NORMS + N_AV), nu_inf), dmax(ld(G, GR_BINF), nu_inf));
  const double tol_d = M.tol_abs + M.tol_rel * dmax(dmax(ld(G, GR_NORMS + N_HREFV), dmax(ld(G, GR_NORMS + N_F), ld(G, GR_NORMS + N_T))), M.Hv_inf);
  st(G, GR_RES + 2, tol_p); st(G, GR_RES + 3, tol_d);
  if (ld(G, GR_RES + 0) < tol_p && ld(G, GR_RES + 1) < tol_d) st(G, GR_NORMS + N_CONVERGED, 1.0);
}
// CheckFeasibility (hxx:572-606)
LOIK_DEV void fine_check_feasibility(const ModelC& M, double* __restrict__ T) {
  double* G = glob_blk(T, M.off);
  const double dyqp = dmax(ld(G, GR_NORMS + N_DFIS), dmax(ld(G, GR_NORMS + N_DYIS), ld(G, GR_NORMS + N_DW)));
  const double ATdy = dmax(ld(G, GR_NORMS + N_DF), ld(G, GR_NORMS + N_DT));
  const bool c1 = ATdy <= M.tol_pinf * dyqp;
  const double ubp = ld(G, GR_NORMS + N_BTDY_P) + ld(G, GR_CARRY + 10), lbm = ld(G, GR_NORMS + N_BTDY_M) + ld(G, GR_CARRY + 11);
  const bool c2 = (ubp + lbm) <= M.tol_pinf * dyqp;
  st(G, GR_NORMS + N_DYQP, dyqp); st(G, GR_NORMS + N_ATDY, ATdy); st(G, GR_NORMS + N_UBP, ubp); st(G, GR_NORMS + N_LBM, lbm);
  st(G, GR_NORMS + N_C1, c1 ? 1.0 : 0.0); st(G, GR_NORMS + N_C2, c2 ? 1.0 : 0.0);
  if (c1 && c2) st(G, GR_NORMS + N_PINFEASIBLE, 1.0);
  st(G, GR_NORMS + N_DX, dmax(ld(G, GR_NORMS + N_DVIS), ld(G, GR_NORMS + N_DNU)));
}
// UpdateMu (hxx:613-641)
LOIK_DEV void fine_update_mu(const ModelC& M, double* __restrict__ T) {
  double* G = glob_blk(T, M.off);
  const double pres = ld(G, GR_RES + 0), dres = ld(G, GR_RES + 1);
  double mu = ld(G, GR_MU);
  if (pres > 10 * dres) mu *= 10; else if (dres > 10 * pres) mu *= 0.1;
  st(G, GR_MU, mu);
}

}  // namespace loik

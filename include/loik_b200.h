/*
 * loik_b200.h -- C ABI of libloik_b200.so: batched, B200-native (sm_100a) replacement for the hot
 * path of LoIK's FirstOrderLoikOptimizedTpl<double>.
 *
 * The reference has no FFI layer; its boundary is the public C++ surface of
 *   FirstOrderLoikOptimizedTpl      /root/reference/include/loik/loik-loid-optimized.hpp:22
 * exported from libloik.so by explicit instantiation (src/loik-loid-optimized.cpp:10-13).  Each entry
 * point below names the reference member it replaces.  One handle = one solver object over a batch
 * of `batch` independent problem instances that share the robot model and the hyper-parameters
 * (the reference: one solver object = one instance).
 *
 * Conventions
 *   - extern "C", opaque handle, int status (0 ok, <0 error; text via loik_last_error()).
 *     No exceptions cross the ABI.  The reference throws std::runtime_error for the same conditions
 *     (ik-id-description-optimized.hpp:38,42,133,143,185,198,329,334; loik-loid-optimized.hxx:634-639).
 *   - All floating point is IEEE double.  Spatial vectors are [linear(0:3); angular(3:6)]
 *     (pinocchio Motion/Force), 6x6 matrices row-major, rotations row-major 3x3.
 *   - Batched arrays are batch-major (instance-major): q[batch][nq], b[batch][nc][6], z[batch][nv] ...
 *     i.e. one contiguous row per problem instance, exactly what a caller looping over reference
 *     solver objects would hold.  The library keeps its own joint-major SoA copy in HBM.
 *   - Every batched pointer argument may be a HOST pointer or a DEVICE pointer; `loc` says which
 *     (LOIK_HOST / LOIK_DEVICE / LOIK_HOST_PINNED).  Pageable host buffers are staged through the library's
 *     pinned buffers and the call synchronizes; device and page-locked buffers are fully asynchronous: the caller
 *     keeps a LOIK_HOST_PINNED / LOIK_DEVICE buffer alive and unmodified until `stream` has passed the call (inputs),
 *     and synchronizes `stream` before reading a LOIK_HOST_PINNED output.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  All work of a call
 *     is enqueued on it; calls that return data to HOST memory synchronize that stream before returning.
 *   - Not thread-safe per handle (same as the reference: one solver+data pair per thread).
 *   - Joints (LOIK_JOINT_*): 1-DoF revolute / prismatic (aligned, unaligned, unbounded), free-flyer, spherical,
 *     translation and planar joints, each anywhere in the tree; idx_q / idx_v are cumulative in joint-id order as in
 *     pinocchio; JointModelSphericalZYX (configuration-dependent motion subspace).  Composite, universal, helical and
 *     mimic joints are rejected by loik_create with LOIK_ERR_UNSUPPORTED.
 *   - `logging` (LoikSolverInfo, loik-loid-optimized.hpp:47-127,406-420) keeps the per-iteration residuals / mu of every
 *     instance (loik_set_logging, loik_get_history) and, like the reference's, costs speed: such solves run on the debug
 *     instantiation of the iteration kernel.  `verbose` is accepted and ignored.
 */
#ifndef LOIK_B200_H_
#define LOIK_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define LOIK_API __attribute__((visibility("default")))
#else
#define LOIK_API
#endif

#define LOIK_MAX_JOINTS 104 /* njoints incl. universe (the batch-uniform model block travels to the kernels by value) */
#define LOIK_MAX_TASKS 32   /* 6-D tasks; up to 8 keep a batch-shared A in the parameter block, more keep A in HBM rows */

/* joint type codes: pinocchio JointModelRX/RY/RZ, PX/PY/PZ, RevoluteUnaligned, PrismaticUnaligned */
enum { LOIK_JOINT_RX = 0, LOIK_JOINT_RY, LOIK_JOINT_RZ, LOIK_JOINT_PX, LOIK_JOINT_PY, LOIK_JOINT_PZ,
       LOIK_JOINT_RU, LOIK_JOINT_PU,
       LOIK_JOINT_FF, /* JointModelFreeFlyer (nq 7 = x y z qx qy qz qw, nv 6), anywhere in the tree */
       /* JointModelRUBX/RUBY/RUBZ and JointModelRevoluteUnboundedUnaligned (URDF `continuous` joints): nq 2 = (cos, sin), nv 1 */
       LOIK_JOINT_RUBX, LOIK_JOINT_RUBY, LOIK_JOINT_RUBZ, LOIK_JOINT_RUBU,
       LOIK_JOINT_SPHERICAL,   /* JointModelSpherical (nq 4 = unit quaternion x y z w, nv 3, S = [0; I3]) */
       LOIK_JOINT_TRANSLATION, /* JointModelTranslation (nq = nv = 3, S = [I3; 0]) */
       LOIK_JOINT_PLANAR,      /* JointModelPlanar (nq 4 = x y cos sin, nv 3 = vx vy wz: S selects components 0, 1, 5) */
       LOIK_JOINT_SPHERICAL_ZYX /* JointModelSphericalZYX (nq = nv = 3: yaw, pitch, roll; R = Rz Ry Rx; the motion subspace
                                   S = [0; E(q)] depends on the configuration) */ };

/* where a caller buffer lives: pageable host memory (staged + synchronous), device memory (asynchronous),
 * or page-locked host memory (asynchronous DMA; the caller synchronizes the stream before reusing / reading it) */
enum { LOIK_HOST = 0, LOIK_DEVICE = 1, LOIK_HOST_PINNED = 2 };

/* error codes */
enum { LOIK_OK = 0, LOIK_ERR_INVALID = -1, LOIK_ERR_CUDA = -2, LOIK_ERR_UNSUPPORTED = -3, LOIK_ERR_STATE = -4 };

/* ADMMPenaltyUpdateStrat (task-solver-base.hpp:13-18); only DEFAULT is implemented by the reference */
enum { LOIK_MU_DEFAULT = 0, LOIK_MU_OSQP = 1, LOIK_MU_MAXEIGENVALUE = 3 };

/* What the hot path reads from pinocchio::Model (loik-loid-optimized.hxx:46-47,258-265). */
typedef struct loik_model_desc {
  int32_t njoints;            /* model.njoints, incl. universe joint 0 */
  const int32_t* parents;     /* [njoints] model.parents, parents[i] < i */
  const int32_t* joint_types; /* [njoints] LOIK_JOINT_* (entry 0 ignored) */
  const double* joint_axes;   /* [njoints][3] unit axis (used by the unaligned types) */
  const double* placement_R;  /* [njoints][9] model.jointPlacements[i].rotation(), row-major */
  const double* placement_p;  /* [njoints][3] model.jointPlacements[i].translation() */
} loik_model_desc;

/* Constructor arguments of FirstOrderLoikOptimizedTpl (loik-loid-optimized.hpp:129-134), same order. */
typedef struct loik_params {
  int32_t max_iter;
  double tol_abs, tol_rel, tol_primal_inf, tol_dual_inf;
  double rho, mu, mu_equality_scale_factor;
  int32_t mu_update_strat; /* LOIK_MU_* */
  int32_t num_eq_c;        /* number of 6-D task constraints (<= LOIK_MAX_TASKS) */
  int32_t eq_c_dim;        /* must be 6 (ik-id-description-optimized.hpp:41-44) */
  int32_t warm_start;
  double tol_tail_solve;
  int32_t verbose; /* accepted, ignored on device */
  int32_t logging; /* != 0: loik_set_logging(h, 1) at creation */
} loik_params;

typedef struct loik_solver loik_solver;

/* How a batched Solve() is laid out on the GPU (not a reference concept; results do not depend on it beyond the
 * rounding of mathematically equal expressions).  A solve is a fixed, host-sync-free sequence of launches:
 *   1. `dense_sweeps` ADMM iterations of every instance in place (one thread per instance, k_iterate);
 *   2. re-pack rounds of 1,1,2,2,4,4,... iterations (`repack_reps` rounds per size, size growing by `repack_growth`):
 *      the still-active instances migrate to the dense prefix of a scratch arena, finished ones go home;
 *      branching trees switch to the segment-parallel kernel (`seg_warps` warps per tile) after `seg_after` sweeps;
 *      rounds after `hi_priority_after` sweeps run on a high-priority stream (< 0: never);
 *   3. after `lane_after` sweeps (0: from the start; < 0: never) every instance still active is finished by the
 *      lane-parallel kernel (8 lanes per instance, state resident in shared memory, device-side work queue).
 * loik_get_schedule reports the current values; the `lane_*` read-only fields describe the geometry chosen at
 * creation (lane_available = 0: the model has multi-DoF joints or its record does not fit shared memory). */
typedef struct loik_schedule {
  int32_t dense_sweeps;
  int32_t repack_reps;
  double repack_growth;
  int32_t hi_priority_after;
  int32_t seg_after;
  int32_t seg_warps;   /* 0 = as many as the tree has parallel chains (at most 4); 1 = one warp per tile only */
  int32_t lane_after;
  int32_t use_graph;   /* replay the schedule from a CUDA graph when the stream can be captured */
  int32_t small_after; /* rounds after this many sweeps are launched with at most `small_grid` CTAs (grid-stride) */
  int32_t small_grid;
  int32_t lane_warps_per_cta; /* warps (of 4 instances each) per CTA of the lane-parallel kernel; 0 = chosen from the record size */
  int32_t lane_groups_per_instance; /* 8-lane groups per instance in that kernel: 1 (four instances per warp, each group sweeps the
                                       whole tree), 4 (one instance per warp, the groups sweep different chains of a branching tree
                                       level by level through a host-built step table), 0 = default (4 when the tree branches, else 1) */
  int32_t drop_workspace;     /* tile kernels: drop the consumed backward->forward workspace lines from L2 (discard.global.L2)
                                 instead of letting them be written back to HBM; never applied with loik_set_keep_workspace */
  double lane_hard_first_ratio; /* hand-over to the lane-parallel kernel: instances still in the main loop whose residual exceeds this many
                                   times its tolerance are queued first (they bound the latency of the solve); 0 = slot order */
  /* read-only (ignored by loik_set_schedule) */
  int32_t lane_available, lane_warps_chosen, lane_groups_chosen, lane_ctas, lane_smem_bytes;
} loik_schedule;

/* Per-instance fields readable with loik_get().  Shapes are per instance; the batch dimension leads. */
typedef enum loik_field {
  LOIK_F_Z = 0,        /* [nv]      ik_id_data.z   -- the answer (loik-loid-optimized.hpp:333) */
  LOIK_F_NU,           /* [nv]      ik_id_data.nu */
  LOIK_F_W,            /* [nv]      ik_id_data.w */
  LOIK_F_Y,            /* [nc][6]   ik_id_data.yis */
  LOIK_F_V,            /* [nb][6]   ik_id_data.vis[1..] */
  LOIK_F_F,            /* [nb][6]   ik_id_data.fis[1..] */
  LOIK_F_ATY,          /* [nc][6]   ik_id_data.Aty */
  LOIK_F_FDPA,         /* [nb][6]   ik_id_data.fis_diff_plus_Aty[1..] */
  LOIK_F_STF_PLUS_W,   /* [nv]      ik_id_data.Stf_plus_w */
  LOIK_F_H,            /* [nb][36]  ik_id_data.His[1..] after the backward pass (accumulated, un-projected) */
  LOIK_F_P,            /* [nb][6]   ik_id_data.pis[1..] */
  LOIK_F_UDINV,        /* [nb][6]   jdata.UDinv() */
  LOIK_F_DINV,         /* [nb]      jdata.Dinv() */
  LOIK_F_R,            /* [nv]      ik_id_data.r (after the backward pass) */
  LOIK_F_LIMI,         /* [nb][12]  ik_id_data.liMi[1..]: rotation (9, row-major) then translation (3) */
  LOIK_F_MU,           /* [1]       get_mu() */
  LOIK_F_ITER,         /* [1] int32 get_iter() */
  LOIK_F_STATUS,       /* [1] int32 bit0 converged_, bit1 primal_infeasible_ (both can be set: the reference evaluates both checks of an
                          iteration, hpp:421-432, and stops on converged_ first), bit2 stopped at max_iter */
  LOIK_F_RESIDUALS,    /* [4]       primal_residual, dual_residual, tol_primal, tol_dual */
  LOIK_F_NORMS,        /* [LOIK_NUM_NORMS] the running norms / sums of IkIdDataTypeOptimized + feasibility scalars
                          (only maintained when loik_set_debug(h,1)); order: loik_norm_index */
  LOIK_F_PRIMAL_RES_VEC, /* [6nb+nv] get_primal_residual_vec() (debug mode only) */
  LOIK_F_DUAL_RES_VEC,   /* [6nb+nv] get_dual_residual_vec()   (debug mode only) */
  LOIK_F_Q               /* [nq]     the configuration the kinematics (liMi) were last initialised with */
} loik_field;

/* index into LOIK_F_NORMS (names = members of IkIdDataTypeOptimizedTpl, loik-loid-data-optimized.hpp:259-329,
 * and the solver's feasibility scalars, loik-loid-optimized.hpp:781-787) */
typedef enum loik_norm_index {
  LOIK_N_BT_DELTA_Y_PLUS = 0, LOIK_N_BT_DELTA_Y_MINUS, LOIK_N_AV_INF, LOIK_N_NU_INF, LOIK_N_HREF_V_INF,
  LOIK_N_FDPA_INF, LOIK_N_STF_PLUS_W_INF, LOIK_N_DELTA_FDPA_INF, LOIK_N_DELTA_STF_PLUS_W_INF,
  LOIK_N_DELTA_VIS_INF, LOIK_N_DELTA_NU_INF, LOIK_N_DELTA_Z_INF, LOIK_N_DELTA_FIS_INF, LOIK_N_DELTA_YIS_INF,
  LOIK_N_DELTA_W_INF,
  LOIK_N_PRIMAL_RES_TASK, LOIK_N_PRIMAL_RES_SLACK, LOIK_N_DUAL_RES_V, LOIK_N_DUAL_RES_NU,
  LOIK_N_DELTA_Y_QP_INF, LOIK_N_AT_DELTA_Y_QP_INF, LOIK_N_UB_T_DELTA_Y_PLUS, LOIK_N_LB_T_DELTA_Y_MINUS,
  LOIK_N_PINF_COND_1, LOIK_N_PINF_COND_2, LOIK_N_DELTA_X_QP_INF,
  LOIK_N_CONVERGED, LOIK_N_PRIMAL_INFEASIBLE, /* flags raised by the per-method steps (LOIK_STEP_CHECK_*) */
  LOIK_NUM_NORMS
} loik_norm_index;

/* The fused steps of one ADMM iteration, for step-by-step parity tests (the reference exposes the
 * individual passes as public members, loik-loid-optimized.hpp:192-264, and its tests call them one
 * by one, tests/loik-loid.cpp:340-478). */
typedef enum loik_step_id {
  LOIK_STEP_BACKWARD = 0, /* UpdatePrev + ResetInfNorms + FwdPass1 + BwdPassOptimizedVisitor      (hxx:290-354) */
  LOIK_STEP_FORWARD,      /* FwdPass2OptimizedVisitor + BoxProj + DualUpdate + ComputePrimalResiduals (hxx:361-503) */
  LOIK_STEP_RESIDUAL,     /* ComputeDualResiduals + CheckConvergence + CheckFeasibility + UpdateMu
                             + the loop-control of Solve()/InfeasibilityTailSolve()               (hxx:510-641) */
  /* the reference's public methods one by one (hpp:192-264).  They ignore / do not advance the loop control, keep the
   * running norms in LOIK_F_NORMS and need loik_set_debug(h, 1). */
  LOIK_STEP_UPDATE_PREV,       /* ik_id_data_.UpdatePrev(): nothing to do (the sweeps read old values before overwriting) */
  LOIK_STEP_RESET_INF_NORMS,   /* ik_id_data_.ResetInfNorms()    (data hxx:165-182) */
  LOIK_STEP_FWD_PASS1,         /* FwdPass1()                     (hxx:290-338): His, pis, r = w - mu z */
  LOIK_STEP_BWD_PASS,          /* BwdPassOptimizedVisitor()      (hxx:345-354) */
  LOIK_STEP_FWD_PASS2,         /* FwdPass2OptimizedVisitor()     (hxx:361-377) */
  LOIK_STEP_BOX_PROJ,          /* BoxProj()                      (hxx:384-397) */
  LOIK_STEP_DUAL_UPDATE,       /* DualUpdate()                   (hxx:404-461) */
  LOIK_STEP_COMPUTE_RESIDUALS, /* ComputeResiduals()             (hxx:529-533) */
  LOIK_STEP_CHECK_CONVERGENCE, /* CheckConvergence()             (hxx:540-565) */
  LOIK_STEP_CHECK_FEASIBILITY, /* CheckFeasibility()             (hxx:572-606) */
  LOIK_STEP_UPDATE_MU          /* UpdateMu()                     (hxx:613-641) */
} loik_step_id;

/* ---- lifetime -------------------------------------------------------------------------------- */
/* FirstOrderLoikOptimizedTpl ctor (hpp:129-162) + IkIdDataTypeOptimizedTpl ctor (data hxx:40-104).
 * Copies the model (the reference holds `Model model_;` by value, hpp:762).  `device` = CUDA ordinal.
 * Any number of solvers may be alive and have work in flight on different streams at the same time. */
LOIK_API int loik_create(const loik_model_desc* model, const loik_params* params, int32_t batch, int32_t device,
                         loik_solver** out);
LOIK_API void loik_destroy(loik_solver* h);
/* Host-side bookkeeping loik_create derives from the model, without touching CUDA (tests / diagnostics; not a reference
 * entry point).  Writes to out[0..n): rows per tile record, #pending blocks, #segments, warps per tile of the
 * segment-parallel kernel, #backward levels, #forward levels, #spans, #multi-DoF joints; then per joint 1..njoints-1
 * {carry, pending block written (-1: none), #pending blocks read, multi-DoF block (-1: none)}; per segment {lo, hi,
 * backward warp, backward level, forward warp, forward level}; per span {lo, hi, nv of a multi-DoF joint or 0}.
 * Returns n, or < 0 (same validation and messages as loik_create; cap too small). */
LOIK_API int32_t loik_model_layout(const loik_model_desc* model, const loik_params* params, int32_t* out, int32_t cap);
/* The step table of the lane-parallel kernel's wide geometry (four 8-lane groups of a warp on different chains of one
 * instance), same conventions: out = {#backward steps, #forward steps}, then for every step of the backward order and
 * of the forward order four entries (group 0..3) of {joint, parent, flags (1 valid, 2 first of a chain, 4 result goes to a
 * pending block, 8 root joint, 16 reads pending blocks), offset of the joint's block in the shared-memory record, pending
 * block written, aligned axis index or -1, offset of the parent's block}. */
LOIK_API int32_t loik_wide_table(const loik_model_desc* model, const loik_params* params, int32_t* out, int32_t cap);
LOIK_API const char* loik_last_error(void);

/* ---- problem set-up -------------------------------------------------------------------------- */
/* SolveInit(q, H_ref, v_ref, ids, Ais, bis, lb, ub)  (hpp:335-361): problem_.Reset, ik_id_data_.Reset(warm_start),
 * ResetSolver, UpdateReference / UpdateIneqConstraints / UpdateEqConstraints, FwdPassInit(q).
 *   q        [batch][nq]
 *   H_ref    [36] one symmetric 6x6 broadcast to all joints (UpdateReference), v_ref [6]
 *   task_joint_ids [nc] joint ids carrying a task (distinct, in 1..njoints-1), shared by the batch
 *   A        [batch][nc][36] if A_per_instance (every instance its own task matrices: e.g. a world-frame end-effector
 *            task, whose A depends on q), else [nc][36] shared by the batch (HOST)
 *   b        [batch][nc][6] if b_per_instance else [nc][6]
 *   lb, ub   [batch][nv] if bounds_per_instance else [nv] (HOST)
 * H_ref, v_ref, task_joint_ids and the batch-shared forms of A and lb/ub are always HOST pointers (small, batch-uniform:
 * they travel to the kernels in the parameter block); `loc` applies to q, b and the per-instance forms of A and lb/ub.
 * Per-instance A costs 57 more rows per task in HBM (A and A^T A, read once per iteration each): 8 * (143 n + 99 nc)
 * algorithmic bytes per instance and iteration instead of 8 * (143 n + 42 nc). */
LOIK_API int loik_solve_init(loik_solver* h, const double* q, const double* H_ref, const double* v_ref, int32_t n_ids,
                             const int32_t* task_joint_ids, const double* A, int32_t A_per_instance, const double* b,
                             int32_t b_per_instance, const double* lb, const double* ub, int32_t bounds_per_instance, int32_t loc,
                             void* stream);

/* problem_.UpdateReferences(H_refs, v_refs)  (ik-id-description-optimized.hpp:103-121): per-joint references,
 * H_refs [njoints][36], v_refs [njoints][6], HOST pointers; call after loik_solve_init. */
LOIK_API int loik_update_references(loik_solver* h, const double* H_refs, const double* v_refs, void* stream);
/* The same with references of its own for every instance (a batch of independent problems: e.g. the previous solution of each
 * trajectory as v_ref): v_refs [batch][njoints][6] at `loc`; H_refs [njoints][36] HOST (per joint, shared by the batch), or, if
 * H_per_instance, [batch][njoints][36] at `loc` (symmetric; checked when the buffer is on the host).  Stays in force until the
 * next loik_solve_init.  Costs 6 (+21 with per-instance weights) more rows per joint in HBM, read twice per iteration:
 * 8 * (155 n + 42 nc) resp. 8 * (197 n + 42 nc) algorithmic bytes per instance and iteration; such solves run on the general
 * instantiations of the tile kernels (the lane-parallel kernel keeps the references in per-CTA constants). */
LOIK_API int loik_update_references_batch(loik_solver* h, const double* H_refs, int32_t H_per_instance, const double* v_refs, int32_t loc,
                                          void* stream);

/* ---- solving --------------------------------------------------------------------------------- */
/* Solve()  (hpp:368-455): ResetRecursion + ResetSolver + main loop, every instance to its own
 * convergence / infeasibility tail / max_iter.  ASYNCHRONOUS: the whole solve is enqueued on `stream` as a fixed
 * schedule of launches (loop control is per instance, on the device); results are valid once the stream has
 * been synchronized or through a later loik_get on the same stream. */
LOIK_API int loik_solve(loik_solver* h, void* stream);
/* Solve(q, H_ref, v_ref, ids, Ais, bis, lb, ub)  (hpp:475-580) = SolveInit + main loop (no ResetRecursion). */
LOIK_API int loik_solve_full(loik_solver* h, const double* q, const double* H_ref, const double* v_ref, int32_t n_ids,
                             const int32_t* task_joint_ids, const double* A, int32_t A_per_instance, const double* b,
                             int32_t b_per_instance, const double* lb, const double* ub, int32_t bounds_per_instance, int32_t loc,
                             void* stream);
/* Solve(q, c_id, Ai, bi)  (hpp:596-695): tailored / trajectory-tracking form: Reset(warm_start), ResetSolver,
 * UpdateEqConstraint(c_id, Ai, bi), FwdPassInit(q), main loop.  Ai [36] HOST, or [batch][36] at `loc` if A_per_instance;
 * Ai == NULL is the UpdateEqConstraint(c_id, bi) overload (ik-id-description-optimized.hpp:224): the task keeps its matrix.  Ai
 * (which must match loik_solve_init's); q [batch][nq], bi [batch][6] (or [6] if !b_per_instance) at `loc`.  q == NULL
 * keeps the device-resident configuration (see loik_integrate). */
LOIK_API int loik_solve_task(loik_solver* h, const double* q, int32_t c_id, const double* Ai, int32_t A_per_instance,
                             const double* bi, int32_t b_per_instance, int32_t loc, void* stream);

/* Outer IK loop on the device (the step after the hot path; README.md:5 of the reference: "differential IK ... to be
 * integrated"): q <- pinocchio::integrate(model, q, dt * z) -- q + dt z for vector-space joints, the SO(2) update of the
 * unbounded revolute joints, quat * exp3 / M * exp6 with the first-order quaternion renormalisation for spherical joints
 * and free-flyers -- and FwdPassInit(q) (hxx:253-283) for every instance.  Follow with loik_solve_task(h, NULL, c_id, Ai, bi, ...) for the next target.  Not a reference entry point. */
LOIK_API int loik_integrate(loik_solver* h, double dt, void* stream);

/* Fixed-iteration mode for throughput measurement: ResetRecursion + ResetSolver, then exactly `iters`
 * ADMM iterations on every instance with stopping disabled (convergence/feasibility still evaluated,
 * UpdateMu still applied).  Not a reference entry point. */
LOIK_API int loik_iterate_fixed(loik_solver* h, int32_t iters, int32_t reset, void* stream);

/* ---- step-by-step interface (parity tests) ---------------------------------------------------- */
/* FwdPassInit(q)  (hxx:253-283) alone: q [batch][nq] at `loc`. */
LOIK_API int loik_fwd_pass_init(loik_solver* h, const double* q, int32_t loc, void* stream);
/* ik_id_data_.ResetRecursion() + ResetSolver(): what Solve() does before its loop (hpp:370-374). */
LOIK_API int loik_reset_recursion(loik_solver* h, void* stream);
/* ResetSolver() alone (hpp:168-186): iteration counter, convergence / infeasibility flags, mu and the feasibility
 * scalars of every instance; the primal and dual state (nu, z, w, vis, fis, yis, Aty ...) is kept. */
LOIK_API int loik_reset_solver(loik_solver* h, void* stream);
/* One step of the current iteration on every instance (the fused ones: on every still-active instance). */
LOIK_API int loik_step(loik_solver* h, int32_t step_id, void* stream);
/* Keep the reference's running norms, feasibility scalars and residual vectors readable (slower). */
LOIK_API int loik_set_debug(loik_solver* h, int32_t on);
/* logging_ / LoikSolverInfo (loik-loid-optimized.hpp:47-127, filled at :406-420 and in InfeasibilityTailSolve :290-306): keep, for
 * every instance and iteration of the following solves, LOIK_HISTORY_COLS values -- primal_residual_task, primal_residual_slack,
 * dual_residual_v, dual_residual_nu, mu (the value the iteration ran with; mu_eq = mu_equality_scale_factor * mu, mu_ineq = mu),
 * delta_x_qp_inf_norm, delta_z_inf_norm, 1 if the iteration belonged to the tail solve.  Switches debug mode on (the log is written
 * by the debug instantiation of the iteration kernel: in place, slower).  Capacity = max_iter at the time of the call.
 * loik_get_history copies [batch][capacity][LOIK_HISTORY_COLS]; rows [0, get_iter()) of an instance are valid. */
#define LOIK_HISTORY_COLS 8
LOIK_API int loik_set_logging(loik_solver* h, int32_t on);
LOIK_API int32_t loik_history_capacity(loik_solver* h);
LOIK_API int loik_get_history(loik_solver* h, double* dst, int32_t loc, void* stream);
/* The reference leaves the workspace of the last backward pass in the caller's data after Solve(): His, pis
 * (ik_id_data, read by tests/loik-loid.cpp:597-615) and jdata.UDinv()/Dinv(), r.  A batched solve packs the
 * still-active instances into dense tiles as it goes, and by default only the state rows of a finished instance
 * return to its slot.  on != 0: the workspace rows return too (35 more rows per joint and solve, once), so
 * LOIK_F_H / _P / _UDINV / _DINV / _R are valid after loik_solve*.  Off (default), loik_get of those fields after a
 * solve fails with LOIK_ERR_STATE instead of returning stale rows; after the in-place steps (loik_step,
 * loik_iterate_fixed, loik_solve_chunk) they are always valid. */
LOIK_API int loik_set_keep_workspace(loik_solver* h, int32_t on);

/* ---- results --------------------------------------------------------------------------------- */
/* Copy a per-instance field of all `batch` instances to dst ([batch][field shape], batch-major).
 * dst is double* except LOIK_F_ITER / LOIK_F_STATUS (int32_t*). */
LOIK_API int loik_get(loik_solver* h, int32_t field, void* dst, int32_t loc, void* stream);
/* Aggregates of the last solve: out[0] = #converged, out[1] = #primal infeasible, out[2] = #stopped at
 * max_iter, out[3] = sum of per-instance iteration counts, out[4] = ADMM sweeps launched (kernel iterations). */
LOIK_API int loik_get_stats(loik_solver* h, int64_t out[5]);
/* Same aggregates, asynchronously: enqueues the reduction on `stream` and returns a device pointer to 4 int64
 * {#converged, #primal infeasible, #stopped at max_iter, sum of iteration counts} -- the buffer the batch-sharded
 * multi-GPU driver all-reduces (NCCL SUM) for the global stopping-criterion outcome. */
LOIK_API int loik_reduce_stats(loik_solver* h, void* stream, void** dev_ptr);
/* Number of CUDA kernels this handle has launched so far (bench.py's gpu_launches). */
LOIK_API int64_t loik_launch_count(loik_solver* h);

/* setters of the base class (task-solver-base.hpp:105-141: set_max_iter, set_tol_abs, set_tol_rel, set_tol_primal_inf,
 * set_tol_dual_inf, set_rho, set_mu, set_mu_equality_scale_factor) and of the solver (loik-loid-optimized.hpp:703
 * set_tol_tail_solve); they take effect at the next solve.  loik_get_params = the getters get_max_iter ... get_rho,
 * get_tol_primal_inf, get_tol_dual_inf (task-solver-base.hpp:87-141) in one call (the per-instance get_mu / get_iter /
 * residuals / tolerances are loik_get fields). */
LOIK_API int loik_set_max_iter(loik_solver* h, int32_t max_iter);
LOIK_API int loik_set_rho(loik_solver* h, double rho);
LOIK_API int loik_set_mu(loik_solver* h, double mu);
LOIK_API int loik_set_mu_equality_scale_factor(loik_solver* h, double factor);
LOIK_API int loik_set_tol_abs(loik_solver* h, double tol);
LOIK_API int loik_set_tol_rel(loik_solver* h, double tol);
LOIK_API int loik_set_tol_primal_inf(loik_solver* h, double tol);
LOIK_API int loik_set_tol_dual_inf(loik_solver* h, double tol);
LOIK_API int loik_set_tol_tail_solve(loik_solver* h, double tol);
LOIK_API int loik_set_warm_start(loik_solver* h, int32_t warm_start);
LOIK_API int loik_get_params(loik_solver* h, loik_params* out);
/* launch schedule (see loik_schedule) */
LOIK_API int loik_get_schedule(loik_solver* h, loik_schedule* out);
LOIK_API int loik_set_schedule(loik_solver* h, const loik_schedule* schedule);
/* Global stop criterion hook for the batch-sharded multi-GPU mode: number of still-active instances
 * on this rank after the last loik_solve_chunk (device-resident int32, for an NCCL all-reduce). */
LOIK_API int loik_active_count_device_ptr(loik_solver* h, void** dev_ptr);
/* Run at most `iters` iterations of the main loop (no reset); returns immediately (async).  Used by
 * the multi-GPU driver, which interleaves chunks with the stop-criterion all-reduce. */
LOIK_API int loik_solve_begin(loik_solver* h, void* stream);
LOIK_API int loik_solve_chunk(loik_solver* h, int32_t iters, void* stream);

LOIK_API int32_t loik_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* LOIK_B200_H_ */

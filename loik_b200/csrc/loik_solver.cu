// loik_solver.cu -- kernels and the C ABI of libloik_b200.so (see include/loik_b200.h).
//
// Product path only: no CPU fallback exists in this library.  If CUDA is unavailable every entry
// point fails with LOIK_ERR_CUDA.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/loik_b200.h"
#include "loik_device.cuh"

namespace loik {


constexpr int kBlock = 128;

// ---------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------

// One launch = up to `iters` ADMM iterations of every active instance (all three sweeps + decisions
// fused; instances are independent so no grid-wide synchronisation is needed between iterations).
template <bool DEBUG, int MINB>
__global__ void __launch_bounds__(kBlock, MINB) k_iterate(const StateP S, const int iters, const int fixed) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  bool active = false;
  if (s < S.n) {
    int status = S.status[s];
    if (status < ST_CONVERGED) {
      int it = S.iter[s];
      double mu = S.mu[s];
      for (int k = 0; k < iters; ++k) {
        ++it;
        const double mu_eq = c_model.mu_scale * mu;
        sweep_backward(S, s, mu, mu_eq);
        Carry cy;
        sweep_forward<DEBUG>(S, s, mu, mu_eq, cy);
        Resid rs;
        sweep_residual<DEBUG>(S, s, rs);
        status = decide<DEBUG>(S, s, status, it, fixed != 0, cy, rs, mu);
        if (status >= ST_CONVERGED) break;
      }
      S.status[s] = status;
      S.iter[s] = it;
      S.mu[s] = mu;
      active = status < ST_CONVERGED;
    }
  }
  const unsigned m = __ballot_sync(0xffffffffu, active);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(S.n_active, __popc(m));
}

// Step-by-step interface: the same sweeps, one per launch, scalars handed over through S.carry.
__global__ void __launch_bounds__(kBlock) k_step_backward(const StateP S) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S.n || S.status[s] >= ST_CONVERGED) return;
  const double mu = S.mu[s];
  sweep_backward(S, s, mu, c_model.mu_scale * mu);
}
__global__ void __launch_bounds__(kBlock) k_step_forward(const StateP S) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S.n || S.status[s] >= ST_CONVERGED) return;
  const double mu = S.mu[s];
  Carry cy;
  sweep_forward<true>(S, s, mu, c_model.mu_scale * mu, cy);
  const double* c = reinterpret_cast<const double*>(&cy);
  for (int k = 0; k < kCarryRows; ++k) st(S.carry, k, S.cap, s, c[k]);
  // ComputePrimalResiduals (hxx:494-503)
  st(S.res, 0, S.cap, s, fmax(cy.pres_task, cy.pres_slack));
  st(S.norms, 15, S.cap, s, cy.pres_task);
  st(S.norms, 16, S.cap, s, cy.pres_slack);
}
__global__ void __launch_bounds__(kBlock) k_step_residual(const StateP S, const int fixed) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S.n) return;
  int status = S.status[s];
  if (status >= ST_CONVERGED) return;
  Carry cy;
  double* c = reinterpret_cast<double*>(&cy);
  for (int k = 0; k < kCarryRows; ++k) c[k] = ld(S.carry, k, S.cap, s);
  Resid rs;
  sweep_residual<true>(S, s, rs);
  double mu = S.mu[s];
  const int it = S.iter[s] + 1;
  status = decide<true>(S, s, status, it, fixed != 0, cy, rs, mu);
  S.status[s] = status;
  S.iter[s] = it;
  S.mu[s] = mu;
}

enum : int { RST_WZ = 1, RST_NU = 2, RST_VFF = 4, RST_YATY = 8, RST_SOLVER = 16 };

// ik_id_data_.Reset / ResetRecursion (data hxx:114-154) + ResetSolver (hpp:168-186)
__global__ void k_reset(const StateP S, const int flags) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S.n) return;
  const int cap = S.cap, nb = c_model.nb, nc = c_model.nc;
  if (flags & RST_WZ)
    for (int r = 0; r < nb; ++r) { st(S.w, r, cap, s, 0.0); st(S.z, r, cap, s, 0.0); }
  if (flags & RST_NU)
    for (int r = 0; r < nb; ++r) st(S.nu, r, cap, s, 0.0);
  if (flags & RST_VFF)
    for (int r = 0; r < 6 * nb; ++r) { st(S.v, r, cap, s, 0.0); st(S.f, r, cap, s, 0.0); st(S.F, r, cap, s, 0.0); }
  if (flags & RST_YATY)
    for (int r = 0; r < 6 * nc; ++r) { st(S.y, r, cap, s, 0.0); st(S.Aty, r, cap, s, 0.0); }
  if (flags & RST_SOLVER) {
    S.status[s] = ST_RUNNING;
    S.iter[s] = 0;
    S.mu[s] = c_model.mu0;
  }
}

// FwdPassInit (hxx:253-283): the q-dependent part of liMi, kept as (sin q, cos q) / (q, 0) per joint.
// q is batch-major [n][nq].
__global__ void k_set_q(const StateP S, const double* __restrict__ q) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S.n) return;
  const int nb = c_model.nb;
  for (int i = 1; i <= nb; ++i) {
    const int jt = c_model.j[i].jtype;
    const double qi = q[(size_t)s * nb + (i - 1)];
    double a, b;
    if (jt <= 2 || jt == 6) sincos(qi, &a, &b);
    else { a = qi; b = 0.0; }
    st(S.jq, 2 * (i - 1), S.cap, s, a);
    st(S.jq, 2 * (i - 1) + 1, S.cap, s, b);
  }
}

// UpdateEqConstraints (ik-id-description-optimized.hpp:127-171), per-instance part: b, Atb = A^T b, |b|inf.
// task < 0: all tasks, bis_inf_norm reset; task >= 0: UpdateEqConstraint for that slot (:178-218), norm only grows.
__global__ void k_set_b(const StateP S, const double* __restrict__ b, const int per_instance, const int task) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S.n) return;
  const int nc = c_model.nc, cap = S.cap;
  double binf = task < 0 ? 0.0 : ld(S.binf, 0, cap, s);
  const int k0 = task < 0 ? 0 : task, k1 = task < 0 ? nc : task + 1;
  for (int k = k0; k < k1; ++k) {
    double bk[6];
    for (int a = 0; a < 6; ++a) {
      const size_t src = task < 0 ? (per_instance ? ((size_t)s * nc + k) * 6 + a : (size_t)k * 6 + a)
                                  : (per_instance ? (size_t)s * 6 + a : (size_t)a);
      bk[a] = b[src];
      st(S.b, 6 * k + a, cap, s, bk[a]);
      binf = fmax(binf, fabs(bk[a]));
    }
    const double* A = c_model.t[k].A;
    for (int a = 0; a < 6; ++a) {
      double acc = 0.0;
      for (int r = 0; r < 6; ++r) acc += A[6 * r + a] * bk[r];
      st(S.Atb, 6 * k + a, cap, s, acc);
    }
  }
  st(S.binf, 0, cap, s, binf);
}

__global__ void k_set_bounds(const StateP S, const double* __restrict__ lb, const double* __restrict__ ub) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S.n) return;
  const int nb = c_model.nb;
  for (int r = 0; r < nb; ++r) {
    st(S.lbv, r, S.cap, s, lb[(size_t)s * nb + r]);
    st(S.ubv, r, S.cap, s, ub[(size_t)s * nb + r]);
  }
}

// batch-major gather of `rows` SoA rows: dst[s][k] = src[map ? map[k] : k][s]
__global__ void k_gather(const double* __restrict__ src, const int cap, const int n, const int rows,
                         const int* __restrict__ map, double* __restrict__ dst) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)n * rows) return;
  const int s = (int)(idx / rows), k = (int)(idx % rows);
  const int r = map ? map[k] : k;
  dst[idx] = r < 0 ? 0.0 : src[(size_t)r * cap + s];
}
__global__ void k_gather_limi(const StateP S, double* __restrict__ dst) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S.n) return;
  const int nb = c_model.nb;
  for (int i = 1; i <= nb; ++i) {
    double R[9], t[3];
    make_xf(c_model.j[i], ld(S.jq, 2 * (i - 1), S.cap, s), ld(S.jq, 2 * (i - 1) + 1, S.cap, s), R, t);
    double* o = dst + ((size_t)s * nb + (i - 1)) * 12;
    for (int c = 0; c < 9; ++c) o[c] = R[c];
    for (int c = 0; c < 3; ++c) o[9 + c] = t[c];
  }
}
__global__ void k_status_flags(const StateP S, int* __restrict__ dst) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S.n) return;
  const int stt = S.status[s];
  dst[s] = (stt == ST_CONVERGED ? 1 : 0) | ((stt == ST_TAIL || stt == ST_INFEASIBLE_DONE) ? 2 : 0) | (stt == ST_MAXITER ? 4 : 0);
}
// out[0..2] = #converged, #infeasible, #maxiter; out[3] = sum iters
__global__ void k_stats(const StateP S, unsigned long long* __restrict__ out) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  int stt = -1, it = 0;
  if (s < S.n) { stt = S.status[s]; it = S.iter[s]; }
  const unsigned c = __ballot_sync(0xffffffffu, stt == ST_CONVERGED);
  const unsigned f = __ballot_sync(0xffffffffu, stt == ST_TAIL || stt == ST_INFEASIBLE_DONE);
  const unsigned m = __ballot_sync(0xffffffffu, stt == ST_MAXITER);
  for (int o = 16; o > 0; o >>= 1) it += __shfl_down_sync(0xffffffffu, it, o);
  if ((threadIdx.x & 31) == 0) {
    if (c) atomicAdd(out + 0, (unsigned long long)__popc(c));
    if (f) atomicAdd(out + 1, (unsigned long long)__popc(f));
    if (m) atomicAdd(out + 2, (unsigned long long)__popc(m));
    if (it) atomicAdd(out + 3, (unsigned long long)it);
  }
}

}  // namespace loik

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
using namespace loik;

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      return fail(LOIK_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));              \
  } while (0)

struct loik_solver {
  int device = 0, batch = 0, cap = 0;
  int nj = 0, nb = 0, nc = 0, npend = 0;
  loik_params prm{};
  ModelC mc{};          // host copy of the constant block
  bool const_dirty = true;
  bool problem_set = false;
  bool debug = false;
  bool bounds_per_instance = false;
  double* arena = nullptr;  // one allocation for every double row
  size_t arena_rows = 0;
  int* iarena = nullptr;    // status, iter
  int* d_n_active = nullptr;
  unsigned long long* d_stats = nullptr;
  int* h_n_active = nullptr;  // pinned
  unsigned long long* h_stats = nullptr;
  int* d_map = nullptr;  // gather map scratch (<= 36*64 ints)
  StateP S{};
  double *lbv = nullptr, *ubv = nullptr;
  // staging
  void* h_stage = nullptr; size_t h_stage_bytes = 0;
  void* d_stage = nullptr; size_t d_stage_bytes = 0;
  int64_t launches = 0;
  int64_t sweeps = 0;
  int chunk_it = 0;  // iterations issued in the current solve
  int minb = 3;
};

static uint64_t g_const_owner = 0;  // which solver's ModelC currently sits in c_model
static uint64_t g_next_id = 1;
struct SolverId { uint64_t id; };
static std::vector<std::pair<loik_solver*, uint64_t>> g_ids;
static uint64_t solver_id(loik_solver* h) {
  for (auto& p : g_ids) if (p.first == h) return p.second;
  g_ids.emplace_back(h, g_next_id++);
  return g_ids.back().second;
}

static int upload_consts(loik_solver* h, cudaStream_t st) {
  const uint64_t id = solver_id(h);
  if (h->const_dirty || g_const_owner != id) {
    CK(cudaMemcpyToSymbolAsync(c_model, &h->mc, sizeof(ModelC), 0, cudaMemcpyHostToDevice, st));
    h->const_dirty = false;
    g_const_owner = id;
  }
  return LOIK_OK;
}

static inline int grid_for(int n) { return (n + kBlock - 1) / kBlock; }

// One place that launches the fused iteration kernel.  `minb` (resident CTAs per SM the kernel is
// compiled for: 2 -> <=255 regs, 3 -> <=168, 4 -> <=128) is a tuning knob (env LOIK_MINB).
static void launch_iterate(loik_solver* h, cudaStream_t st, int iters, int fixed) {
  const int g = grid_for(h->batch);
#define LOIK_LAUNCH(DBG, MB) k_iterate<DBG, MB><<<g, kBlock, 0, st>>>(h->S, iters, fixed)
  if (h->debug) { LOIK_LAUNCH(true, 2); }
  else if (h->minb == 2) { LOIK_LAUNCH(false, 2); }
  else if (h->minb == 3) { LOIK_LAUNCH(false, 3); }
  else { LOIK_LAUNCH(false, 4); }
#undef LOIK_LAUNCH
}

static int ensure_stage(loik_solver* h, size_t bytes) {
  if (bytes > h->h_stage_bytes) {
    if (h->h_stage) cudaFreeHost(h->h_stage);
    h->h_stage = nullptr; h->h_stage_bytes = 0;
    CK(cudaMallocHost(&h->h_stage, bytes));
    h->h_stage_bytes = bytes;
  }
  if (bytes > h->d_stage_bytes) {
    if (h->d_stage) cudaFree(h->d_stage);
    h->d_stage = nullptr; h->d_stage_bytes = 0;
    CK(cudaMalloc(&h->d_stage, bytes));
    h->d_stage_bytes = bytes;
  }
  return LOIK_OK;
}

// Bring a caller buffer to the device (no-op for device pointers).  Host buffers are copied into the
// pinned staging area first so the H2D copy is a true async DMA; `off` lets several inputs share it.
static int to_device(loik_solver* h, const void* src, size_t bytes, int loc, size_t off, cudaStream_t st, const void** out) {
  if (loc == LOIK_DEVICE) { *out = src; return LOIK_OK; }
  std::memcpy((char*)h->h_stage + off, src, bytes);
  CK(cudaMemcpyAsync((char*)h->d_stage + off, (char*)h->h_stage + off, bytes, cudaMemcpyHostToDevice, st));
  *out = (char*)h->d_stage + off;
  return LOIK_OK;
}

static void sym_blocks(const double* M, double* A, double* B, double* D) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      if (i <= j) { A[si(i, j)] = M[6 * i + j]; D[si(i, j)] = M[6 * (3 + i) + 3 + j]; }
      B[3 * i + j] = M[6 * i + 3 + j];
    }
}
static bool is_symmetric(const double* M) {
  for (int i = 0; i < 6; ++i)
    for (int j = i + 1; j < 6; ++j) {
      const double a = M[6 * i + j], b = M[6 * j + i];
      if (std::fabs(a - b) > 1e-12 * std::max(1.0, std::max(std::fabs(a), std::fabs(b)))) return false;
    }
  return true;
}

extern "C" {

int32_t loik_abi_version(void) { return 1; }
const char* loik_last_error(void) { return g_err.c_str(); }

int loik_create(const loik_model_desc* model, const loik_params* params, int32_t batch, int32_t device, loik_solver** out) {
  if (!model || !params || !out) return fail(LOIK_ERR_INVALID, "loik_create: null argument");
  *out = nullptr;
  const int nj = model->njoints;
  // IkProblemFormulationOptimized ctor checks (ik-id-description-optimized.hpp:37-44)
  if (params->eq_c_dim != 6)
    return fail(LOIK_ERR_INVALID, "[IkProblemFormulation::IkProblemFormulation]: equality constraint dimension is not 6, problem formulation not supported !!!");
  if (nj < 2 || nj > LOIK_MAX_JOINTS) return fail(LOIK_ERR_INVALID, "loik_create: njoints out of range [2, LOIK_MAX_JOINTS]");
  if (params->num_eq_c < 0 || params->num_eq_c > LOIK_MAX_TASKS) return fail(LOIK_ERR_INVALID, "loik_create: num_eq_c out of range [0, LOIK_MAX_TASKS]");
  if (batch < 1) return fail(LOIK_ERR_INVALID, "loik_create: batch must be >= 1");
  for (int i = 1; i < nj; ++i) {
    if (model->parents[i] < 0 || model->parents[i] >= i) return fail(LOIK_ERR_INVALID, "loik_create: parents[i] must be < i");
    if (model->joint_types[i] < 0 || model->joint_types[i] > LOIK_JOINT_PU) return fail(LOIK_ERR_UNSUPPORTED, "loik_create: only 1-DoF revolute/prismatic joints are supported");
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(LOIK_ERR_CUDA, "loik_create: no CUDA device (libloik_b200 has no CPU fallback)");
  CK(cudaSetDevice(device));
  loik_solver* h = new loik_solver();
  h->device = device; h->batch = batch; h->cap = (batch + 31) / 32 * 32;
  h->nj = nj; h->nb = nj - 1; h->nc = params->num_eq_c; h->prm = *params;
  if (const char* e = std::getenv("LOIK_MINB")) { const int v = std::atoi(e); if (v >= 2 && v <= 4) h->minb = v; }
  ModelC& M = h->mc;
  std::memset(&M, 0, sizeof(M));
  M.nj = nj; M.nb = nj - 1; M.nc = h->nc;
  M.max_iter = params->max_iter; M.rho = params->rho; M.mu0 = params->mu; M.mu_scale = params->mu_equality_scale_factor;
  M.tol_abs = params->tol_abs; M.tol_rel = params->tol_rel; M.tol_pinf = params->tol_primal_inf; M.tol_dinf = params->tol_dual_inf;
  M.tol_tail = params->tol_tail_solve;
  // tree bookkeeping: which contributions travel in registers (parent == i-1) and which go through a pending slot
  std::vector<int> pend(nj, -1);
  int npend = 0;
  for (int i = nj - 1; i >= 1; --i) {
    JointC& J = M.j[i];
    J.parent = model->parents[i]; J.jtype = model->joint_types[i]; J.task = -1;
    for (int c = 0; c < 9; ++c) J.plR[c] = model->placement_R[9 * i + c];
    for (int c = 0; c < 3; ++c) { J.plp[c] = model->placement_p[3 * i + c]; J.axis[c] = model->joint_axes[3 * i + c]; }
    J.carry = (J.parent > 0 && J.parent == i - 1) ? 1 : 0;
    J.pfirst = 0; J.ppend = -1;
    if (J.parent > 0 && !J.carry) {
      if (pend[J.parent] < 0) { pend[J.parent] = npend++; J.pfirst = 1; }
      J.ppend = pend[J.parent];
    }
  }
  for (int i = 1; i < nj; ++i) M.j[i].pend = pend[i];
  M.npend = npend; h->npend = npend;
  // device memory: one arena of rows
  const int nb = h->nb, nc = std::max(h->nc, 1), cap = h->cap;
  size_t rows = 0;
  auto take = [&](size_t r) { size_t o = rows; rows += r; return o; };
  const size_t o_v = take(6 * nb), o_f = take(6 * nb), o_F = take(6 * nb), o_nu = take(nb), o_z = take(nb), o_w = take(nb), o_T = take(nb);
  const size_t o_y = take(6 * nc), o_Aty = take(6 * nc), o_jq = take(2 * nb), o_b = take(6 * nc), o_Atb = take(6 * nc), o_binf = take(1);
  const size_t o_lb = take(nb), o_ub = take(nb), o_mu = take(1), o_res = take(4);
  const size_t o_H = take(21 * nb), o_p = take(6 * nb), o_UD = take(6 * nb), o_Di = take(nb), o_r = take(nb);
  const size_t o_pH = take(27 * std::max(npend, 1)), o_pF = take(6 * std::max(npend, 1));
  const size_t o_cy = take(kCarryRows), o_norms = take(LOIK_NUM_NORMS), o_prv = take(7 * nb), o_drv = take(7 * nb);
  h->arena_rows = rows;
  if (cudaMalloc(&h->arena, rows * cap * sizeof(double)) != cudaSuccess) { delete h; return fail(LOIK_ERR_CUDA, "loik_create: cudaMalloc failed"); }
  cudaMemset(h->arena, 0, rows * cap * sizeof(double));
  cudaMalloc(&h->iarena, 2 * (size_t)cap * sizeof(int));
  cudaMemset(h->iarena, 0, 2 * (size_t)cap * sizeof(int));
  cudaMalloc(&h->d_n_active, sizeof(int));
  cudaMalloc(&h->d_stats, 4 * sizeof(unsigned long long));
  cudaMalloc(&h->d_map, 36 * LOIK_MAX_JOINTS * sizeof(int));
  cudaMallocHost(&h->h_n_active, sizeof(int));
  cudaMallocHost(&h->h_stats, 4 * sizeof(unsigned long long));
  auto P = [&](size_t o) { return h->arena + o * cap; };
  StateP& S = h->S;
  S.cap = cap; S.n = batch;
  S.v = P(o_v); S.f = P(o_f); S.F = P(o_F); S.nu = P(o_nu); S.z = P(o_z); S.w = P(o_w); S.T = P(o_T); S.y = P(o_y); S.Aty = P(o_Aty);
  S.jq = P(o_jq); S.b = P(o_b); S.Atb = P(o_Atb); S.binf = P(o_binf); S.lbv = nullptr; S.ubv = nullptr;
  h->lbv = P(o_lb); h->ubv = P(o_ub);
  S.mu = P(o_mu); S.res = P(o_res); S.status = h->iarena; S.iter = h->iarena + cap;
  S.H = P(o_H); S.p = P(o_p); S.UDinv = P(o_UD); S.Dinv = P(o_Di); S.r = P(o_r); S.pendH = P(o_pH); S.pendF = P(o_pF);
  S.carry = P(o_cy); S.norms = P(o_norms); S.prv = P(o_prv); S.drv = P(o_drv);
  S.n_active = h->d_n_active;
  if (cudaGetLastError() != cudaSuccess) { loik_destroy(h); return fail(LOIK_ERR_CUDA, "loik_create: allocation failed"); }
  *out = h;
  return LOIK_OK;
}

void loik_destroy(loik_solver* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  cudaFree(h->arena); cudaFree(h->iarena); cudaFree(h->d_n_active); cudaFree(h->d_stats); cudaFree(h->d_map);
  cudaFreeHost(h->h_n_active); cudaFreeHost(h->h_stats);
  if (h->h_stage) cudaFreeHost(h->h_stage);
  if (h->d_stage) cudaFree(h->d_stage);
  for (size_t i = 0; i < g_ids.size(); ++i)
    if (g_ids[i].first == h) { if (g_const_owner == g_ids[i].second) g_const_owner = 0; g_ids.erase(g_ids.begin() + i); break; }
  delete h;
}

static int launch_reset(loik_solver* h, int flags, cudaStream_t st) {
  int rc = upload_consts(h, st);
  if (rc) return rc;
  k_reset<<<grid_for(h->batch), kBlock, 0, st>>>(h->S, flags);
  h->launches++;
  CK(cudaGetLastError());
  return LOIK_OK;
}

// problem_.UpdateReference / UpdateIneqConstraints / UpdateEqConstraints: the batch-uniform part goes to the constant block
static int set_problem_consts(loik_solver* h, const double* H_ref, const double* v_ref, int n_ids, const int32_t* ids,
                              const double* A, const double* lb, const double* ub, bool bounds_shared) {
  ModelC& M = h->mc;
  if (n_ids != h->nc)
    return fail(LOIK_ERR_INVALID, "[IkProblemFormulation::UpdateEqConstraints]: number of equality constraints doesn't match initialization!!!");
  if (!is_symmetric(H_ref))
    return fail(LOIK_ERR_UNSUPPORTED, "loik_solve_init: H_ref must be symmetric (the optimized path's SE3actOn reads only the LL, LA, AA blocks)");
  double Hv[6];
  for (int i = 0; i < 6; ++i) { Hv[i] = 0; for (int j = 0; j < 6; ++j) Hv[i] += H_ref[6 * i + j] * v_ref[j]; }
  double hv_inf = 0; for (int i = 0; i < 6; ++i) hv_inf = std::max(hv_inf, std::fabs(Hv[i]));
  M.Hv_inf = hv_inf;  // = |Hv[0]|inf (ik-id-description-optimized.hpp:95)
  for (int i = 1; i < h->nj; ++i) {
    JointC& J = M.j[i];
    sym_blocks(H_ref, J.HrA, J.HrB, J.HrD);
    for (int c = 0; c < 6; ++c) J.Hv[c] = Hv[c];
    J.task = -1;
    if (bounds_shared) { J.lb = lb[i - 1]; J.ub = ub[i - 1]; }
  }
  for (int k = 0; k < n_ids; ++k) {
    const int c = ids[k];
    if (c < 1 || c >= h->nj) return fail(LOIK_ERR_INVALID, "loik_solve_init: task joint id out of range [1, njoints-1]");
    if (M.j[c].task >= 0)
      return fail(LOIK_ERR_UNSUPPORTED, "[IkProblemFormulation::UpdateEqConstraint]: multiple constraint specification for the same link id, not supported, terminating !!!");
    M.j[c].task = k;
    TaskC& T = M.t[k];
    T.joint = c;
    double AtA[36];
    for (int i = 0; i < 36; ++i) T.A[i] = A[36 * k + i];
    for (int i = 0; i < 6; ++i)
      for (int j = 0; j < 6; ++j) { double s = 0; for (int r = 0; r < 6; ++r) s += T.A[6 * r + i] * T.A[6 * r + j]; AtA[6 * i + j] = s; }
    sym_blocks(AtA, T.AtA_A, T.AtA_B, T.AtA_D);
  }
  h->const_dirty = true;
  return LOIK_OK;
}

int loik_solve_init(loik_solver* h, const double* q, const double* H_ref, const double* v_ref, int32_t n_ids,
                    const int32_t* ids, const double* A, const double* b, int32_t b_per_instance, const double* lb,
                    const double* ub, int32_t bounds_per_instance, int32_t loc, void* stream) {
  if (!h || !q || !H_ref || !v_ref || !lb || !ub || (n_ids > 0 && (!ids || !A || !b))) return fail(LOIK_ERR_INVALID, "loik_solve_init: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  CK(cudaSetDevice(h->device));
  const int B = h->batch, nb = h->nb, nc = h->nc;
  const size_t q_bytes = (size_t)B * nb * sizeof(double);
  const size_t b_bytes = (size_t)(b_per_instance ? B : 1) * nc * 6 * sizeof(double);
  const size_t bd_bytes = (size_t)(bounds_per_instance ? B : 1) * nb * sizeof(double);
  // shared bounds / shared b are small: read them on the host when they are host pointers
  std::vector<double> lbh(nb), ubh(nb);
  if (!bounds_per_instance) {
    if (loc == LOIK_HOST) { std::memcpy(lbh.data(), lb, bd_bytes); std::memcpy(ubh.data(), ub, bd_bytes); }
    else { CK(cudaMemcpyAsync(lbh.data(), lb, bd_bytes, cudaMemcpyDeviceToHost, st)); CK(cudaMemcpyAsync(ubh.data(), ub, bd_bytes, cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st)); }
  }
  int rc = set_problem_consts(h, H_ref, v_ref, n_ids, ids, A, lbh.data(), ubh.data(), !bounds_per_instance);
  if (rc) return rc;
  if (loc == LOIK_HOST) { rc = ensure_stage(h, q_bytes + b_bytes + 2 * bd_bytes + 64); if (rc) return rc; }
  rc = upload_consts(h, st);
  if (rc) return rc;
  const void *dq, *db = nullptr, *dlb = nullptr, *dub = nullptr;
  size_t off = 0;
  rc = to_device(h, q, q_bytes, loc, off, st, &dq); if (rc) return rc; off += q_bytes;
  if (nc > 0) { rc = to_device(h, b, b_bytes, loc, off, st, &db); if (rc) return rc; off += b_bytes; }
  if (bounds_per_instance) {
    rc = to_device(h, lb, bd_bytes, loc, off, st, &dlb); if (rc) return rc; off += bd_bytes;
    rc = to_device(h, ub, bd_bytes, loc, off, st, &dub); if (rc) return rc; off += bd_bytes;
  }
  // ik_id_data_.Reset(warm_start) + ResetSolver() + FwdPassInit's y/Aty wipe (hpp:346-359, hxx:270-278)
  const int flags = RST_SOLVER | (h->prm.warm_start ? 0 : (RST_WZ | RST_NU | RST_VFF | RST_YATY));
  k_reset<<<grid_for(B), kBlock, 0, st>>>(h->S, flags);
  k_set_q<<<grid_for(B), kBlock, 0, st>>>(h->S, (const double*)dq);
  h->launches += 2;
  if (nc > 0) { k_set_b<<<grid_for(B), kBlock, 0, st>>>(h->S, (const double*)db, b_per_instance, -1); h->launches++; }
  h->bounds_per_instance = bounds_per_instance != 0;
  if (bounds_per_instance) {
    h->S.lbv = h->lbv; h->S.ubv = h->ubv;
    k_set_bounds<<<grid_for(B), kBlock, 0, st>>>(h->S, (const double*)dlb, (const double*)dub);
    h->launches++;
  } else {
    h->S.lbv = nullptr; h->S.ubv = nullptr;
  }
  CK(cudaGetLastError());
  if (loc == LOIK_HOST) CK(cudaStreamSynchronize(st));  // the staging buffer may be reused by the next call
  h->problem_set = true;
  return LOIK_OK;
}

int loik_update_references(loik_solver* h, const double* H_refs, const double* v_refs, void* stream) {
  if (!h || !H_refs || !v_refs) return fail(LOIK_ERR_INVALID, "loik_update_references: null argument");
  ModelC& M = h->mc;
  for (int i = 0; i < h->nj; ++i) {
    if (!is_symmetric(H_refs + 36 * i)) return fail(LOIK_ERR_UNSUPPORTED, "loik_update_references: H_refs[i] must be symmetric");
    double Hv[6], n = 0;
    for (int a = 0; a < 6; ++a) { Hv[a] = 0; for (int c = 0; c < 6; ++c) Hv[a] += H_refs[36 * i + 6 * a + c] * v_refs[6 * i + c]; n = std::max(n, std::fabs(Hv[a])); }
    if (n > M.Hv_inf) M.Hv_inf = n;  // only grows (ik-id-description-optimized.hpp:115-117)
    if (i >= 1) { sym_blocks(H_refs + 36 * i, M.j[i].HrA, M.j[i].HrB, M.j[i].HrD); for (int a = 0; a < 6; ++a) M.j[i].Hv[a] = Hv[a]; }
  }
  h->const_dirty = true;
  return LOIK_OK;
}

// the main loop of Solve() (hpp:377-454) over the whole batch: launches fused iteration kernels until no
// instance is active.  `chunk` iterations per launch; the active counter is read back after every launch.
static int run_loop(loik_solver* h, cudaStream_t st, int max_sweeps, bool fixed) {
  int rc = upload_consts(h, st);
  if (rc) return rc;
  const int B = h->batch;
  int done = 0;
  while (done < max_sweeps) {
    const int chunk = std::min(fixed ? max_sweeps : 4, max_sweeps - done);
    CK(cudaMemsetAsync(h->d_n_active, 0, sizeof(int), st));
    launch_iterate(h, st, chunk, fixed ? 1 : 0);
    h->launches++;
    h->sweeps += chunk;
    done += chunk;
    if (!fixed) {
      CK(cudaMemcpyAsync(h->h_n_active, h->d_n_active, sizeof(int), cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      if (*h->h_n_active == 0) break;
    }
  }
  CK(cudaGetLastError());
  return LOIK_OK;
}

int loik_reset_recursion(loik_solver* h, void* stream) {
  if (!h) return fail(LOIK_ERR_INVALID, "null handle");
  if (!h->problem_set) return fail(LOIK_ERR_STATE, "loik_reset_recursion: call loik_solve_init first");
  CK(cudaSetDevice(h->device));
  return launch_reset(h, RST_WZ | RST_VFF | RST_YATY | RST_SOLVER, (cudaStream_t)stream);
}

int loik_solve(loik_solver* h, void* stream) {
  if (!h) return fail(LOIK_ERR_INVALID, "null handle");
  if (!h->problem_set) return fail(LOIK_ERR_STATE, "loik_solve: call loik_solve_init first");
  if (h->prm.mu_update_strat != LOIK_MU_DEFAULT)
    return fail(LOIK_ERR_UNSUPPORTED, "[FirstOrderLoikOptimizedTpl::UpdateMu]: mu update strategy not yet implemented");
  cudaStream_t st = (cudaStream_t)stream;
  CK(cudaSetDevice(h->device));
  int rc = launch_reset(h, RST_WZ | RST_VFF | RST_YATY | RST_SOLVER, st);  // ResetRecursion + ResetSolver (hpp:370-374)
  if (rc) return rc;
  if (h->prm.max_iter < 2) return LOIK_OK;
  return run_loop(h, st, h->prm.max_iter, false);
}

int loik_solve_full(loik_solver* h, const double* q, const double* H_ref, const double* v_ref, int32_t n_ids,
                    const int32_t* ids, const double* A, const double* b, int32_t b_per_instance, const double* lb,
                    const double* ub, int32_t bounds_per_instance, int32_t loc, void* stream) {
  if (h && h->prm.mu_update_strat != LOIK_MU_DEFAULT)
    return fail(LOIK_ERR_UNSUPPORTED, "[FirstOrderLoikOptimizedTpl::UpdateMu]: mu update strategy not yet implemented");
  int rc = loik_solve_init(h, q, H_ref, v_ref, n_ids, ids, A, b, b_per_instance, lb, ub, bounds_per_instance, loc, stream);
  if (rc) return rc;
  if (h->prm.max_iter < 2) return LOIK_OK;
  return run_loop(h, (cudaStream_t)stream, h->prm.max_iter, false);
}

int loik_solve_task(loik_solver* h, const double* q, int32_t c_id, const double* Ai, const double* bi,
                    int32_t b_per_instance, int32_t loc, void* stream) {
  if (!h || !q || !Ai || !bi) return fail(LOIK_ERR_INVALID, "loik_solve_task: null argument");
  if (!h->problem_set) return fail(LOIK_ERR_STATE, "loik_solve_task: call loik_solve_init first");
  cudaStream_t st = (cudaStream_t)stream;
  CK(cudaSetDevice(h->device));
  ModelC& M = h->mc;
  // problem_.UpdateEqConstraint(c_id, Ai, bi) (ik-id-description-optimized.hpp:178-218)
  int k = -1;
  for (int t = 0; t < h->nc; ++t) if (M.t[t].joint == c_id) k = t;
  if (k < 0) return fail(LOIK_ERR_INVALID, "[IkProblemFormulation::UpdateEqConstraint]: constraint doesn't yet exist at link 'c_id' !!! ");
  TaskC& T = M.t[k];
  double AtA[36];
  for (int i = 0; i < 36; ++i) T.A[i] = Ai[i];
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) { double s = 0; for (int r = 0; r < 6; ++r) s += T.A[6 * r + i] * T.A[6 * r + j]; AtA[6 * i + j] = s; }
  sym_blocks(AtA, T.AtA_A, T.AtA_B, T.AtA_D);
  h->const_dirty = true;
  const int B = h->batch, nb = h->nb;
  const size_t q_bytes = (size_t)B * nb * sizeof(double), b_bytes = (size_t)(b_per_instance ? B : 1) * 6 * sizeof(double);
  int rc;
  if (loc == LOIK_HOST) { rc = ensure_stage(h, q_bytes + b_bytes + 64); if (rc) return rc; }
  rc = upload_consts(h, st); if (rc) return rc;
  const void *dq, *db;
  rc = to_device(h, q, q_bytes, loc, 0, st, &dq); if (rc) return rc;
  rc = to_device(h, bi, b_bytes, loc, q_bytes, st, &db); if (rc) return rc;
  const int flags = RST_SOLVER | (h->prm.warm_start ? 0 : (RST_WZ | RST_NU | RST_VFF | RST_YATY));
  k_reset<<<grid_for(B), kBlock, 0, st>>>(h->S, flags);
  k_set_b<<<grid_for(B), kBlock, 0, st>>>(h->S, (const double*)db, b_per_instance, k);
  k_set_q<<<grid_for(B), kBlock, 0, st>>>(h->S, (const double*)dq);
  h->launches += 3;
  CK(cudaGetLastError());
  if (loc == LOIK_HOST) CK(cudaStreamSynchronize(st));
  if (h->prm.max_iter < 2) return LOIK_OK;
  return run_loop(h, st, h->prm.max_iter, false);
}

int loik_iterate_fixed(loik_solver* h, int32_t iters, int32_t reset, void* stream) {
  if (!h) return fail(LOIK_ERR_INVALID, "null handle");
  if (!h->problem_set) return fail(LOIK_ERR_STATE, "loik_iterate_fixed: call loik_solve_init first");
  cudaStream_t st = (cudaStream_t)stream;
  CK(cudaSetDevice(h->device));
  if (reset) { int rc = launch_reset(h, RST_WZ | RST_VFF | RST_YATY | RST_SOLVER, st); if (rc) return rc; }
  int rc = upload_consts(h, st);
  if (rc) return rc;
  // one launch per iteration: this is the quantity the roofline is quoted on
  for (int i = 0; i < iters; ++i) {
    launch_iterate(h, st, 1, 1);
  }
  h->launches += iters; h->sweeps += iters;
  CK(cudaGetLastError());
  return LOIK_OK;
}

int loik_solve_begin(loik_solver* h, void* stream) {
  if (!h) return fail(LOIK_ERR_INVALID, "null handle");
  if (!h->problem_set) return fail(LOIK_ERR_STATE, "loik_solve_begin: call loik_solve_init first");
  CK(cudaSetDevice(h->device));
  h->chunk_it = 0;
  return launch_reset(h, RST_WZ | RST_VFF | RST_YATY | RST_SOLVER, (cudaStream_t)stream);
}
int loik_solve_chunk(loik_solver* h, int32_t iters, void* stream) {
  if (!h) return fail(LOIK_ERR_INVALID, "null handle");
  cudaStream_t st = (cudaStream_t)stream;
  CK(cudaSetDevice(h->device));
  int rc = upload_consts(h, st);
  if (rc) return rc;
  CK(cudaMemsetAsync(h->d_n_active, 0, sizeof(int), st));
  launch_iterate(h, st, iters, 0);
  h->launches++; h->sweeps += iters; h->chunk_it += iters;
  CK(cudaGetLastError());
  return LOIK_OK;
}
int loik_solve_end(loik_solver* h, void* stream) { (void)h; (void)stream; return LOIK_OK; }
int loik_active_count_device_ptr(loik_solver* h, void** dev_ptr) {
  if (!h || !dev_ptr) return fail(LOIK_ERR_INVALID, "null argument");
  *dev_ptr = h->d_n_active;
  return LOIK_OK;
}

int loik_step(loik_solver* h, int32_t step_id, void* stream) {
  if (!h) return fail(LOIK_ERR_INVALID, "null handle");
  if (!h->problem_set) return fail(LOIK_ERR_STATE, "loik_step: call loik_solve_init first");
  cudaStream_t st = (cudaStream_t)stream;
  CK(cudaSetDevice(h->device));
  int rc = upload_consts(h, st);
  if (rc) return rc;
  const int g = grid_for(h->batch);
  switch (step_id) {
    case LOIK_STEP_BACKWARD: k_step_backward<<<g, kBlock, 0, st>>>(h->S); break;
    case LOIK_STEP_FORWARD: k_step_forward<<<g, kBlock, 0, st>>>(h->S); break;
    case LOIK_STEP_RESIDUAL: k_step_residual<<<g, kBlock, 0, st>>>(h->S, 0); h->sweeps++; break;
    default: return fail(LOIK_ERR_INVALID, "loik_step: unknown step id");
  }
  h->launches++;
  CK(cudaGetLastError());
  return LOIK_OK;
}

int loik_set_debug(loik_solver* h, int32_t on) {
  if (!h) return fail(LOIK_ERR_INVALID, "null handle");
  h->debug = on != 0;
  return LOIK_OK;
}

int loik_get(loik_solver* h, int32_t field, void* dst, int32_t loc, void* stream) {
  if (!h || !dst) return fail(LOIK_ERR_INVALID, "loik_get: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  CK(cudaSetDevice(h->device));
  int rc = upload_consts(h, st);
  if (rc) return rc;
  const int B = h->batch, nb = h->nb, nc = h->nc, cap = h->cap;
  const StateP& S = h->S;
  const double* src = nullptr;
  int rows = 0;
  std::vector<int> map;
  bool is_int = false;
  switch (field) {
    case LOIK_F_Z: src = S.z; rows = nb; break;
    case LOIK_F_NU: src = S.nu; rows = nb; break;
    case LOIK_F_W: src = S.w; rows = nb; break;
    case LOIK_F_Y: src = S.y; rows = 6 * nc; break;
    case LOIK_F_V: src = S.v; rows = 6 * nb; break;
    case LOIK_F_F: src = S.f; rows = 6 * nb; break;
    case LOIK_F_ATY: src = S.Aty; rows = 6 * nc; break;
    case LOIK_F_FDPA: src = S.F; rows = 6 * nb; break;
    case LOIK_F_STF_PLUS_W: src = S.T; rows = nb; break;
    case LOIK_F_P: src = S.p; rows = 6 * nb; break;
    case LOIK_F_UDINV: src = S.UDinv; rows = 6 * nb; break;
    case LOIK_F_DINV: src = S.Dinv; rows = nb; break;
    case LOIK_F_R: src = S.r; rows = nb; break;
    case LOIK_F_MU: src = S.mu; rows = 1; break;
    case LOIK_F_RESIDUALS: src = S.res; rows = 4; break;
    case LOIK_F_NORMS: src = S.norms; rows = LOIK_NUM_NORMS; break;
    case LOIK_F_PRIMAL_RES_VEC: src = S.prv; rows = 7 * nb; break;
    case LOIK_F_DUAL_RES_VEC: src = S.drv; rows = 7 * nb; break;
    case LOIK_F_H: {  // expand the 21 stored scalars of each joint to a full symmetric 6x6
      src = S.H; rows = 36 * nb; map.resize(rows);
      for (int j = 0; j < nb; ++j)
        for (int a = 0; a < 6; ++a)
          for (int c = 0; c < 6; ++c) {
            int r;
            if (a < 3 && c < 3) r = si(a, c);
            else if (a >= 3 && c >= 3) r = 15 + si(a - 3, c - 3);
            else if (a < 3) r = 6 + 3 * a + (c - 3);
            else r = 6 + 3 * c + (a - 3);
            map[36 * j + 6 * a + c] = 21 * j + r;
          }
      break;
    }
    case LOIK_F_LIMI: rows = 12 * nb; break;
    case LOIK_F_ITER: case LOIK_F_STATUS: is_int = true; rows = 1; break;
    default: return fail(LOIK_ERR_INVALID, "loik_get: unknown field");
  }
  const size_t bytes = (size_t)B * rows * (is_int ? sizeof(int) : sizeof(double));
  void* ddst = dst;
  if (loc == LOIK_HOST) { rc = ensure_stage(h, bytes); if (rc) return rc; ddst = h->d_stage; }
  if (field == LOIK_F_LIMI) {
    k_gather_limi<<<grid_for(B), kBlock, 0, st>>>(S, (double*)ddst);
  } else if (field == LOIK_F_ITER) {
    CK(cudaMemcpyAsync(ddst, S.iter, bytes, cudaMemcpyDeviceToDevice, st));
  } else if (field == LOIK_F_STATUS) {
    k_status_flags<<<grid_for(B), kBlock, 0, st>>>(S, (int*)ddst);
  } else {
    const int* dmap = nullptr;
    if (!map.empty()) { CK(cudaMemcpyAsync(h->d_map, map.data(), map.size() * sizeof(int), cudaMemcpyHostToDevice, st)); CK(cudaStreamSynchronize(st)); dmap = h->d_map; }
    const size_t total = (size_t)B * rows;
    k_gather<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(src, cap, B, rows, dmap, (double*)ddst);
  }
  h->launches++;
  CK(cudaGetLastError());
  if (loc == LOIK_HOST) {
    CK(cudaMemcpyAsync(h->h_stage, ddst, bytes, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    std::memcpy(dst, h->h_stage, bytes);
  }
  return LOIK_OK;
}

int loik_get_stats(loik_solver* h, int64_t out[5]) {
  if (!h || !out) return fail(LOIK_ERR_INVALID, "null argument");
  CK(cudaSetDevice(h->device));
  CK(cudaMemset(h->d_stats, 0, 4 * sizeof(unsigned long long)));
  k_stats<<<grid_for(h->cap), kBlock>>>(h->S, h->d_stats);
  h->launches++;
  CK(cudaMemcpy(h->h_stats, h->d_stats, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  for (int i = 0; i < 4; ++i) out[i] = (int64_t)h->h_stats[i];
  out[4] = h->sweeps;
  return LOIK_OK;
}

int64_t loik_launch_count(loik_solver* h) { return h ? h->launches : 0; }

int loik_set_max_iter(loik_solver* h, int32_t m) { if (!h) return LOIK_ERR_INVALID; h->prm.max_iter = m; h->mc.max_iter = m; h->const_dirty = true; return LOIK_OK; }
int loik_set_rho(loik_solver* h, double rho) { if (!h) return LOIK_ERR_INVALID; h->prm.rho = rho; h->mc.rho = rho; h->const_dirty = true; return LOIK_OK; }
int loik_set_mu(loik_solver* h, double mu) { if (!h) return LOIK_ERR_INVALID; h->prm.mu = mu; h->mc.mu0 = mu; h->const_dirty = true; return LOIK_OK; }
int loik_set_tol_tail_solve(loik_solver* h, double tol) { if (!h) return LOIK_ERR_INVALID; h->prm.tol_tail_solve = tol; h->mc.tol_tail = tol; h->const_dirty = true; return LOIK_OK; }
int loik_set_warm_start(loik_solver* h, int32_t ws) { if (!h) return LOIK_ERR_INVALID; h->prm.warm_start = ws; return LOIK_OK; }

}  // extern "C"

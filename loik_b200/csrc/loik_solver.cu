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
#include "loik_lane.cuh"

namespace loik {

#ifndef LOIK_KBLOCK
#define LOIK_KBLOCK 64
#endif
constexpr int kBlock = LOIK_KBLOCK;  // threads per CTA of the sweep kernels (2 warps = 2 tiles)

// ---------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------

LOIK_DEV int2 ld_ctl(const ModelC& c_model, const double* T) { return *reinterpret_cast<const int2*>(T + (size_t)(c_model.off.glob + GR_CTL) * 32); }
LOIK_DEV void st_ctl(const ModelC& c_model, double* T, int status, int iter) { *reinterpret_cast<int2*>(T + (size_t)(c_model.off.glob + GR_CTL) * 32) = make_int2(status, iter); }

LOIK_DEV double* tile_of(double* arena, const ModelC& c_model, int s) { return arena + ((size_t)(s >> 5) * c_model.off.rows) * 32 + (s & 31); }

// Results of a finished instance go from its packed slot (tile pointer Ts) to its home slot (Th): the state rows
// only (v, f, F, nu, z, w, T | y, Aty | mu, control, residuals); problem data and workspace stay behind.
LOIK_DEV void retire_rows(const ModelC& c_model, const double* Ts, double* Th) {
  const Offs& O = c_model.off;
  {
    const double* Gs = glob_blk(const_cast<double*>(Ts), O);
    double* Gd = glob_blk(Th, O);
#pragma unroll
    for (int r = 0; r < GR_CARRY; ++r)
      if (r != GR_BINF) st(Gd, r, ld(Gs, r));
  }
  const int nb = c_model.nb, nc = c_model.nc;
  for (int j = 0; j < nb; ++j) {
    const double* Ps = joint_blk(const_cast<double*>(Ts), O, j);
    double* Pd = joint_blk(Th, O, j);
    double tmp[JR_JQ];
#pragma unroll
    for (int r = 0; r < JR_JQ; ++r) tmp[r] = ld(Ps, r);  // v, f, F, nu, z, w, T (22 rows)
#pragma unroll
    for (int r = 0; r < JR_JQ; ++r) st(Pd, r, tmp[r]);
  }
  for (int m = 0; m < c_model.nmd; ++m) {
    const double* Ps = md_blk(const_cast<double*>(Ts), O, m);
    double* Pd = md_blk(Th, O, m);
    for (int r = 0; r < FR_LB; ++r) st(Pd, r, ld(Ps, r));  // nu, z, w, T
  }
  for (int t = 0; t < nc; ++t) {
    const double* Ps = task_blk(const_cast<double*>(Ts), O, t);
    double* Pd = task_blk(Th, O, t);
#pragma unroll
    for (int r = 0; r < TR_B; ++r) st(Pd, r, ld(Ps, r));  // y, Aty
  }
}

// Opt-in (loik_set_keep_workspace): the workspace of the last backward pass (His, pis, UDinv, Dinv, r -- what the
// reference leaves in ik_id_data / jdata after Solve()) follows a retiring instance home too.  Out of line: a rare
// path that must not take part in the register allocation of the sweeps.
LOIK_DEV_CALL void retire_workspace(const ModelC& c_model, const double* Ts, double* Th) {
  const Offs& O = c_model.off;
  for (int j = 0; j < c_model.nb; ++j) {
    const double* Ps = joint_blk(const_cast<double*>(Ts), O, j);
    double* Pd = joint_blk(Th, O, j);
    for (int r = JR_H; r < JR_HV; ++r) st(Pd, r, ld(Ps, r));  // (the rows behind the workspace are problem data: they never left home)
  }
  for (int m = 0; m < c_model.nmd; ++m) {
    const double* Ps = md_blk(const_cast<double*>(Ts), O, m);
    double* Pd = md_blk(Th, O, m);
    for (int r = FR_DINV; r < FR_S; ++r) st(Pd, r, ld(Ps, r));
  }
}

// First iteration of a migrating launch: the rows of the globals block that no sweep rewrites travel here.
LOIK_DEV void migrate_globals(const ModelC& c_model, const double* Ts, double* Td) {
  const double* Gs = glob_blk(const_cast<double*>(Ts), c_model.off);
  double* Gd = glob_blk(Td, c_model.off);
  st(Gd, GR_BINF, ld(Gs, GR_BINF));
#pragma unroll
  for (int r = 0; r < 4; ++r) st(Gd, GR_RES + r, ld(Gs, GR_RES + r));
}

// Survivors of a launch claim the slots of the next one (order inside a warp is preserved, so neighbours stay
// neighbours and the gather of the next launch's first iteration stays mostly coalesced).
LOIK_DEV void claim_next(const StateP& S, const bool active, const int slot_now, const bool hard = true) {
  const int lane = threadIdx.x & 31;
  const unsigned m = __ballot_sync(0xffffffffu, active && (hard || !S.next_back));
  if (m) {
    int base = 0;
    if (lane == 0) base = atomicAdd(S.next_count, __popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (m >> lane & 1u) S.next_list[base + __popc(m & ((1u << lane) - 1))] = slot_now;
  }
  if (S.next_back) {  // (hand-over to the lane-parallel kernel: the rest queues up from the end of the list)
    const unsigned e = __ballot_sync(0xffffffffu, active && !hard);
    if (e) {
      int base = 0;
      if (lane == 0) base = atomicAdd(S.next_back, __popc(e));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (e >> lane & 1u) S.next_list[S.next_cap - 1 - (base + __popc(e & ((1u << lane) - 1)))] = slot_now;
    }
  }
}
// "far from done": still in the main loop with a residual more than hard_ratio x its tolerance away (rows GR_RES, just written by
// decide).  After 4 Panda iterations a ratio of 20 flags a third of the running instances and every one that goes on for more
// than 50 iterations; instances on the infeasibility tail finish within a few iterations
LOIK_DEV bool looks_hard(const ModelC& c_model, const StateP& S, const double* Td, const int status) {
  if (status == ST_TAIL) return false;
  const double* G = glob_blk(const_cast<double*>(Td), c_model.off);
  return ld(G, GR_RES + 0) > S.hard_ratio * ld(G, GR_RES + 2) || ld(G, GR_RES + 1) > S.hard_ratio * ld(G, GR_RES + 3);
}

// One launch = up to `iters` ADMM iterations of every active instance (all three sweeps + decisions
// fused; instances are independent so no grid-wide synchronisation is needed between iterations).
// Dense mode: thread k = slot k.  List mode: thread k = slot list[k] for k < *n_list (compacted
// still-active instances; the grid is sized for the worst case and surplus CTAs exit at once), in place, or --
// migrating launch, S.dst set -- read there during the first iteration and written to slot k of S.dst: the physical
// re-pack of the still-active instances into full tiles costs no pass of its own.
// MINB = resident CTAs per SM the kernel is compiled for -> register cap 65536 / (64 MINB), rounded down to 8.  Only
// MINB = 4 (255 registers) is instantiated: 5 / 6 / 8 CTAs per SM (200 / 168 / 128 registers) spill and were 1.3-1.6x
// slower (profiles/r1_history.md).
// MD: the model has multi-DoF joints (span-level dispatch, out-of-line steps); the 1-DoF-only instantiation is the
// plain joint loops, untouched by that machinery (it is the kernel every BASELINE robot runs).
template <bool DEBUG, int MINB, bool MD = false>
__global__ void __launch_bounds__(kBlock) __maxnreg__(MINB <= 4 ? 255 : (65536 / (kBlock * MINB)) / 8 * 8) k_iterate(const __grid_constant__ ModelC c_model, const StateP S, const int iters, const int fixed) {
  const int limit = S.list ? *S.n_list : (S.n_dev ? *S.n_dev : S.n);
  const int stride = gridDim.x * blockDim.x;
  // the warp writes a whole tile (in place without a list, or the dense prefix of a migrating launch): its dead workspace
  // lines can be dropped from L2 (discard_workspace)
  const bool drop_ws = !DEBUG && S.drop_ws && (S.list == nullptr || S.dst != nullptr);
  // grid-stride over the slots: late rounds are launched with a small grid (the count lives on the device)
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k - (int)threadIdx.x % 32 < limit; k += stride) {
  const int s = k < limit ? (S.list ? S.list[k] : k) : -1;
  bool active = false, hard = true;
  if (s >= 0) {
    const bool MIG = S.dst != nullptr;
    double* Td = MIG ? tile_of(S.dst, c_model, k) : tile_ptr(S, c_model, s);
    const double* Ts = MIG ? tile_ptr(S, c_model, s) : Td;
    const int2 ctl = ld_ctl(c_model, Ts);
    int status = ctl.x;
    double* Dg = DEBUG ? dbg_ptr(S, c_model, s) : nullptr;  // (debug mode runs in place: s is the home slot)
    if (status < ST_CONVERGED) {
      int it = ctl.y;
      double mu = ld(glob_blk(const_cast<double*>(Ts), c_model.off), GR_MU);
      bool migrate = MIG;  // the first iteration of a migrating launch reads at Ts and writes at Td
      if (MIG) { migrate_globals(c_model, Ts, Td); S.origin_dst[k] = S.origin_src ? S.origin_src[s] : s; }
      for (int n = 0; n < iters; ++n) {
        ++it;
        const double mu_eq = c_model.mu_scale * mu;
        Carry cy;
        Resid rs;
        zero(cy);
        zero(rs);
        if (MD) {
          for (int g = c_model.nspan - 1; g >= 0; --g) span_backward(c_model, Ts, Td, mu, mu_eq, c_model.span[g].lo, c_model.span[g].hi, migrate);
          for (int g = 0; g < c_model.nspan; ++g) span_forward<DEBUG>(c_model, Ts, Td, mu, mu_eq, cy, c_model.span[g].lo, c_model.span[g].hi, drop_ws, Dg);
          for (int g = c_model.nspan - 1; g >= 0; --g) span_residual<DEBUG>(c_model, Ts, Td, rs, c_model.span[g].lo, c_model.span[g].hi, Dg);
        } else {
          const int nb = c_model.nb;
          sweep_backward(c_model, Ts, Td, mu, mu_eq, 1, nb, migrate);
          sweep_forward<DEBUG>(c_model, Ts, Td, mu, mu_eq, cy, 1, nb, drop_ws, Dg);
          sweep_residual<DEBUG>(c_model, Ts, Td, rs, 1, nb, Dg);
        }
        const double mu_used = mu;
        const int status_was = status;
        status = decide<DEBUG>(c_model, Td, status, it, fixed != 0, cy, rs, mu);
        if (DEBUG && S.hist && it <= S.hist_cap) {  // the solver log (debug mode runs in place: s is the home slot)
          double* Hh = S.hist + ((size_t)s * S.hist_cap + (it - 1)) * kHistCols;
          Hh[0] = cy.pres_task; Hh[1] = cy.pres_slack; Hh[2] = rs.dres_v; Hh[3] = rs.T_inf; Hh[4] = mu_used;
          Hh[5] = dmax(cy.dvis_inf, cy.dnu_inf); Hh[6] = cy.dz_inf; Hh[7] = status_was == ST_TAIL ? 1.0 : 0.0;
        }
        Ts = Td; migrate = false;
        if (status >= ST_CONVERGED) break;
      }
      st_ctl(c_model, Td, status, it);
      st(glob_blk(Td, c_model.off), GR_MU, mu);
      active = status < ST_CONVERGED;
      if (S.next_back && active) hard = looks_hard(c_model, S, Td, status);
      if (!active && S.home) {  // finished away from home: the results go to the home slot now
        const int* origin = MIG ? S.origin_dst : S.origin_src;
        if (origin) {
          double* Th = tile_of(S.home, c_model, origin[MIG ? k : s]);
          retire_rows(c_model, Td, Th);
          if (S.keep_ws) retire_workspace(c_model, Td, Th);
        }
      }
    }
  }
  if (S.n_active) {
    const unsigned m = __ballot_sync(0xffffffffu, active);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(S.n_active, __popc(m));
  }
  if (S.next_list) claim_next(S, active, S.dst ? k : s, hard);
  }
}

// Segment-parallel variant for branching trees: one CTA = one tile of 32 instances, NW warps.  Warp w sweeps the
// chains (segments) of the tree assigned to it; chains only exchange data through pending blocks / the parent's v
// row in HBM/L2, ordered by CTA barriers between the levels of the segment DAG.  The running norms are combined
// through shared memory in a fixed warp order, after which every warp takes the same decisions redundantly.
template <bool DEBUG, int NW, bool MD = false>
__global__ void __launch_bounds__(32 * NW, NW <= 2 ? 4 : 2)
    k_iterate_seg(const __grid_constant__ ModelC c_model, const StateP S, const int iters, const int fixed) {
  constexpr int NP = kCarryRows + 7;
  __shared__ double part[NW][NP][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int limit = S.list ? *S.n_list : (S.n_dev ? *S.n_dev : S.n);
  const bool MIG = S.dst != nullptr;
  const bool drop_ws = !DEBUG && S.drop_ws && (S.list == nullptr || MIG);
  for (int tile = blockIdx.x; tile * 32 < limit; tile += gridDim.x) {
    const int k = tile * 32 + lane;
    const int s = k < limit ? (S.list ? S.list[k] : k) : -1;
    double* Td = nullptr;
    int status = ST_CONVERGED, it = 0;
    double mu = 0.0;
    if (s >= 0) Td = MIG ? tile_of(S.dst, c_model, k) : tile_ptr(S, c_model, s);
    const double* Ts = (MIG && s >= 0) ? tile_ptr(S, c_model, s) : Td;  // (first iteration of a migrating launch)
    if (s >= 0) {
      const int2 ctl = ld_ctl(c_model, Ts);
      status = ctl.x; it = ctl.y;
      mu = ld(glob_blk(const_cast<double*>(Ts), c_model.off), GR_MU);
    }
    const bool was_active = status < ST_CONVERGED;
    bool migrate = MIG;
    if (MIG && was_active && w == 0) { migrate_globals(c_model, Ts, Td); S.origin_dst[k] = S.origin_src ? S.origin_src[s] : s; }
    for (int n = 0; n < iters; ++n) {
      const bool alive = status < ST_CONVERGED;
      if (!__any_sync(0xffffffffu, alive)) break;  // identical in every warp of the CTA (same lanes, same decisions)
      const double mu_eq = c_model.mu_scale * mu;
      Carry cy;
      Resid rs;
      zero(cy);
      zero(rs);
      for (int lv = 0; lv < c_model.nblevel; ++lv) {
        if (alive)
          for (int g = 0; g < c_model.nseg; ++g)
            if (c_model.seg[g].bwarp == w && c_model.seg[g].blevel == lv) {
              if (MD) span_backward(c_model, Ts, Td, mu, mu_eq, c_model.seg[g].lo, c_model.seg[g].hi, migrate);
              else sweep_backward(c_model, Ts, Td, mu, mu_eq, c_model.seg[g].lo, c_model.seg[g].hi, migrate);
            }
        __syncthreads();
      }
      for (int lv = 0; lv < c_model.nflevel; ++lv) {
        if (alive)
          for (int g = 0; g < c_model.nseg; ++g)
            if (c_model.seg[g].fwarp == w && c_model.seg[g].flevel == lv) {
              if (MD) span_forward<DEBUG>(c_model, Ts, Td, mu, mu_eq, cy, c_model.seg[g].lo, c_model.seg[g].hi, drop_ws);
              else sweep_forward<DEBUG>(c_model, Ts, Td, mu, mu_eq, cy, c_model.seg[g].lo, c_model.seg[g].hi, drop_ws);
            }
        __syncthreads();
      }
      for (int lv = 0; lv < c_model.nblevel; ++lv) {
        if (alive)
          for (int g = 0; g < c_model.nseg; ++g)
            if (c_model.seg[g].bwarp == w && c_model.seg[g].blevel == lv) {
              if (MD) span_residual<DEBUG>(c_model, Ts, Td, rs, c_model.seg[g].lo, c_model.seg[g].hi);
              else sweep_residual<DEBUG>(c_model, Ts, Td, rs, c_model.seg[g].lo, c_model.seg[g].hi);
            }
        __syncthreads();
      }
      {  // combine the per-warp partial norms / sums
        const double* c = reinterpret_cast<const double*>(&cy);
        const double* r = reinterpret_cast<const double*>(&rs);
#pragma unroll
        for (int q = 0; q < kCarryRows; ++q) part[w][q][lane] = c[q];
#pragma unroll
        for (int q = 0; q < 7; ++q) part[w][kCarryRows + q][lane] = r[q];
      }
      __syncthreads();
      {
        double* c = reinterpret_cast<double*>(&cy);
        double* r = reinterpret_cast<double*>(&rs);
#pragma unroll
        for (int q = 0; q < kCarryRows; ++q) {
          const bool is_sum = q >= 8 && q <= 11;  // bTdy_p, bTdy_m, ubdw_p, lbdw_m are sums, the rest are inf-norms
          double acc = part[0][q][lane];
#pragma unroll
          for (int ww = 1; ww < NW; ++ww) acc = is_sum ? acc + part[ww][q][lane] : dmax(acc, part[ww][q][lane]);
          c[q] = acc;
        }
#pragma unroll
        for (int q = 0; q < 7; ++q) {
          double acc = part[0][kCarryRows + q][lane];
#pragma unroll
          for (int ww = 1; ww < NW; ++ww) acc = dmax(acc, part[ww][kCarryRows + q][lane]);
          r[q] = acc;
        }
      }
      if (alive) {
        ++it;
        status = decide<DEBUG>(c_model, Td, status, it, fixed != 0, cy, rs, mu, w == 0);
      }
      Ts = Td; migrate = false;
      __syncthreads();  // part[] is rewritten in the next iteration
    }
    bool active = false, hard = true;
    if (was_active) {
      active = status < ST_CONVERGED;
      if (w == 0) {
        st_ctl(c_model, Td, status, it);
        st(glob_blk(Td, c_model.off), GR_MU, mu);
        if (S.next_back && active) hard = looks_hard(c_model, S, Td, status);
        // (the last iteration's stores of the other warps are ordered before this by the barrier that ends it)
        if (!active && S.home) {
          const int* origin = MIG ? S.origin_dst : S.origin_src;
          if (origin) {
            double* Th = tile_of(S.home, c_model, origin[MIG ? k : s]);
            retire_rows(c_model, Td, Th);
            if (S.keep_ws) retire_workspace(c_model, Td, Th);
          }
        }
      }
    }
    if (w == 0) {
      if (S.n_active) {
        const unsigned m = __ballot_sync(0xffffffffu, active);
        if (lane == 0 && m) atomicAdd(S.n_active, __popc(m));
      }
      if (S.next_list) claim_next(S, active, S.dst ? k : s, hard);
    }
    __syncthreads();  // the next tile's first sweeps must not overtake this tile's result stores of warp 0
  }
}

// Compaction: append the slots that are still active to list_out (order inside a CTA is preserved so
// neighbouring instances stay neighbours).  Input is the previous list, or all slots when list_in == nullptr.
__global__ void __launch_bounds__(256) k_compact(const __grid_constant__ ModelC c_model, const StateP S, const int* __restrict__ list_in, const int* __restrict__ n_in,
                                                 int* __restrict__ list_out, int* __restrict__ n_out) {
  __shared__ int warp_cnt[8];
  __shared__ int block_base;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int limit = list_in ? *n_in : (S.n_dev ? *S.n_dev : S.n);
  int s = -1;
  if (k < limit) s = list_in ? list_in[k] : k;
  bool active = false;
  if (s >= 0) active = ld_ctl(c_model, tile_ptr(S, c_model, s)).x < ST_CONVERGED;
  const unsigned m = __ballot_sync(0xffffffffu, active);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) warp_cnt[wid] = __popc(m);
  __syncthreads();
  if (threadIdx.x == 0) {
    int tot = 0;
    for (int w = 0; w < 8; ++w) { const int c = warp_cnt[w]; warp_cnt[w] = tot; tot += c; }
    block_base = tot ? atomicAdd(n_out, tot) : 0;
  }
  __syncthreads();
  if (active) list_out[block_base + warp_cnt[wid] + __popc(m & ((1u << lane) - 1))] = s;
}

// Step-by-step interface: the same sweeps, one per launch, scalars handed over through the carry rows.
__global__ void __launch_bounds__(kBlock) k_step_backward(const __grid_constant__ ModelC c_model, const StateP S) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S.n) return;
  double* T = tile_ptr(S, c_model, s);
  if (ld_ctl(c_model, T).x >= ST_CONVERGED) return;
  const double mu = ld(glob_blk(T, c_model.off), GR_MU);
  for (int g = c_model.nspan - 1; g >= 0; --g) span_backward(c_model, T, T, mu, c_model.mu_scale * mu, c_model.span[g].lo, c_model.span[g].hi, false);
}
__global__ void __launch_bounds__(kBlock) k_step_forward(const __grid_constant__ ModelC c_model, const StateP S) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S.n) return;
  double* T = tile_ptr(S, c_model, s);
  if (ld_ctl(c_model, T).x >= ST_CONVERGED) return;
  double* G = glob_blk(T, c_model.off);
  const double mu = ld(G, GR_MU);
  Carry cy;
  zero(cy);
  for (int g = 0; g < c_model.nspan; ++g) span_forward<true>(c_model, T, T, mu, c_model.mu_scale * mu, cy, c_model.span[g].lo, c_model.span[g].hi, false, dbg_ptr(S, c_model, s));
  const double* c = reinterpret_cast<const double*>(&cy);
  for (int k = 0; k < kCarryRows; ++k) st(G, GR_CARRY + k, c[k]);
  // ComputePrimalResiduals (hxx:494-503)
  st(G, GR_RES + 0, fmax(cy.pres_task, cy.pres_slack));
  st(G, GR_NORMS + 15, cy.pres_task);
  st(G, GR_NORMS + 16, cy.pres_slack);
}
__global__ void __launch_bounds__(kBlock) k_step_residual(const __grid_constant__ ModelC c_model, const StateP S, const int fixed) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S.n) return;
  double* T = tile_ptr(S, c_model, s);
  const int2 ctl = ld_ctl(c_model, T);
  int status = ctl.x;
  if (status >= ST_CONVERGED) return;
  double* G = glob_blk(T, c_model.off);
  Carry cy;
  double* c = reinterpret_cast<double*>(&cy);
  for (int k = 0; k < kCarryRows; ++k) c[k] = ld(G, GR_CARRY + k);
  Resid rs;
  zero(rs);
  for (int g = c_model.nspan - 1; g >= 0; --g) span_residual<true>(c_model, T, T, rs, c_model.span[g].lo, c_model.span[g].hi, dbg_ptr(S, c_model, s));
  double mu = ld(G, GR_MU);
  const int it = ctl.y + 1;
  status = decide<true>(c_model, T, status, it, fixed != 0, cy, rs, mu);
  st_ctl(c_model, T, status, it);
  st(G, GR_MU, mu);
}

// The reference's public methods one by one (LOIK_STEP_UPDATE_PREV ... LOIK_STEP_UPDATE_MU), every instance.
__global__ void __launch_bounds__(kBlock) k_fine(const __grid_constant__ ModelC c_model, const StateP S, const int which) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S.n) return;
  double* T = tile_ptr(S, c_model, s);
  const double mu = ld(glob_blk(T, c_model.off), GR_MU), mu_eq = c_model.mu_scale * mu;
  switch (which) {
    case LOIK_STEP_RESET_INF_NORMS: fine_reset_inf_norms(c_model, T); break;
    case LOIK_STEP_FWD_PASS1: fine_fwdpass1(c_model, T, mu, mu_eq); break;
    case LOIK_STEP_BWD_PASS: sweep_backward<true>(c_model, T, T, mu, mu_eq, 1, c_model.nb); break;
    case LOIK_STEP_FWD_PASS2: fine_fwdpass2(c_model, T); break;
    case LOIK_STEP_BOX_PROJ: fine_boxproj(c_model, T, mu, dbg_ptr(S, c_model, s)); break;
    case LOIK_STEP_DUAL_UPDATE: fine_dualupdate(c_model, T, mu, mu_eq, dbg_ptr(S, c_model, s)); break;
    case LOIK_STEP_COMPUTE_RESIDUALS: fine_compute_residuals(c_model, T, dbg_ptr(S, c_model, s)); break;
    case LOIK_STEP_CHECK_CONVERGENCE: fine_check_convergence(c_model, T); break;
    case LOIK_STEP_CHECK_FEASIBILITY: fine_check_feasibility(c_model, T); break;
    case LOIK_STEP_UPDATE_MU: fine_update_mu(c_model, T); break;
    default: break;
  }
}

enum : int { RST_WZ = 1, RST_NU = 2, RST_VFF = 4, RST_YATY = 8, RST_SOLVER = 16 };

// ik_id_data_.Reset / ResetRecursion (data hxx:114-154) + ResetSolver (hpp:168-186)
__global__ void k_reset(const __grid_constant__ ModelC c_model, const StateP S, const int flags) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S.n) return;
  double* T = tile_ptr(S, c_model, s);
  const Offs& O = c_model.off;
  const int nb = c_model.nb, nc = c_model.nc;
  for (int j = 0; j < nb; ++j) {
    double* Pj = joint_blk(T, O, j);
    if (flags & RST_WZ) { st(Pj, JR_W, 0.0); st(Pj, JR_Z, 0.0); }
    if (flags & RST_NU) st(Pj, JR_NU, 0.0);
    if (flags & RST_VFF)
      for (int c = 0; c < 6; ++c) { st(Pj, JR_V + c, 0.0); st(Pj, JR_F + c, 0.0); st(Pj, JR_FD + c, 0.0); }
  }
  for (int m = 0; m < c_model.nmd; ++m) {
    double* Pf = md_blk(T, O, m);
    for (int c = 0; c < 6; ++c) {
      if (flags & RST_WZ) { st(Pf, FR_W + c, 0.0); st(Pf, FR_Z + c, 0.0); }
      if (flags & RST_NU) st(Pf, FR_NU + c, 0.0);
    }
  }
  if (flags & RST_YATY)
    for (int k = 0; k < nc; ++k) {
      double* Pk = task_blk(T, O, k);
      for (int c = 0; c < 6; ++c) { st(Pk, TR_Y + c, 0.0); st(Pk, TR_ATY + c, 0.0); }
    }
  if (flags & RST_SOLVER) {
    st_ctl(c_model, T, ST_RUNNING, 0);
    st(glob_blk(T, O), GR_MU, c_model.mu0);
    st(glob_blk(T, O), GR_NORMS + N_CONVERGED, 0.0);
    st(glob_blk(T, O), GR_NORMS + N_PINFEASIBLE, 0.0);
  }
}

// JointModelSphericalZYX::calc (pinocchio joint-spherical-ZYX.hpp): M = Rz(q0) Ry(q1) Rx(q2) and the angular block E(q) of the
// motion subspace S = [0; E] (body angular velocity = E qdot)
LOIK_DEV void zyx_calc(const double q0, const double q1, const double q2, double (&M)[9], double (&E)[9]) {
  double s0, c0, s1, c1, s2, c2;
  sincos(q0, &s0, &c0); sincos(q1, &s1, &c1); sincos(q2, &s2, &c2);
  M[0] = c0 * c1; M[1] = c0 * s1 * s2 - s0 * c2; M[2] = c0 * s1 * c2 + s0 * s2;
  M[3] = s0 * c1; M[4] = s0 * s1 * s2 + c0 * c2; M[5] = s0 * s1 * c2 - c0 * s2;
  M[6] = -s1;     M[7] = c1 * s2;                M[8] = c1 * c2;
  E[0] = -s1;     E[1] = 0.0; E[2] = 1.0;
  E[3] = c1 * s2; E[4] = c2;  E[5] = 0.0;
  E[6] = c1 * c2; E[7] = -s2; E[8] = 0.0;
}

// FwdPassInit (hxx:253-283): the q-dependent part of liMi, kept as (sin q, cos q) / (q, 0) per joint.
// q is batch-major [n][nq]; it is staged through shared memory so both the read and the write coalesce.
__global__ void __launch_bounds__(kBlock) k_set_q(const __grid_constant__ ModelC c_model, const StateP S, const double* __restrict__ q) {
  extern __shared__ double sh[];
  const int nb = c_model.nb, nq = c_model.nq;
  const int s0 = blockIdx.x * blockDim.x;
  const int cnt = min((int)blockDim.x, S.n - s0);
  for (int e = threadIdx.x; e < cnt * nq; e += blockDim.x) sh[e] = q[(size_t)s0 * nq + e];
  __syncthreads();
  const int s = s0 + threadIdx.x;
  if (s >= S.n) return;
  double* T = tile_ptr(S, c_model, s);
  for (int i = 1; i <= nb; ++i) {
    const int jt = c_model.j[i].jtype;
    if (c_model.j[i].nvj > 1) {  // multi-DoF joint: keep q, liMi = placement * M(q) (hxx:263-264) into the md block
      const JointC& J = c_model.j[i];
      const double* qj = sh + threadIdx.x * nq + J.idxq;
      double* Pf = md_blk(T, c_model.off, J.mblk);
      const int nqj = jt == LOIK_JOINT_FF ? 7 : ((jt == LOIK_JOINT_TRANSLATION || jt == LOIK_JOINT_SPHERICAL_ZYX) ? 3 : 4);
      for (int c = 0; c < nqj; ++c) st(Pf, FR_Q + c, qj[c]);
      double M[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, pq[3] = {0, 0, 0};
      if (jt == LOIK_JOINT_PLANAR) {  // JointModelPlanar::calc: R = Rz from (cos, sin) = (q[2], q[3]) as given, p = (x, y, 0)
        M[0] = qj[2]; M[1] = -qj[3]; M[3] = qj[3]; M[4] = qj[2];
        pq[0] = qj[0]; pq[1] = qj[1];
      } else if (jt == LOIK_JOINT_SPHERICAL_ZYX) {
        double E[9];
        zyx_calc(qj[0], qj[1], qj[2], M, E);
        for (int c = 0; c < 9; ++c) st(Pf, FR_S + c, E[c]);
      } else if (jt != LOIK_JOINT_TRANSLATION) {
        const int o = jt == LOIK_JOINT_FF ? 3 : 0;
        const double x = qj[o], y = qj[o + 1], z = qj[o + 2], w = qj[o + 3];
        M[0] = 1 - 2 * (y * y + z * z); M[1] = 2 * (x * y - z * w); M[2] = 2 * (x * z + y * w);
        M[3] = 2 * (x * y + z * w); M[4] = 1 - 2 * (x * x + z * z); M[5] = 2 * (y * z - x * w);
        M[6] = 2 * (x * z - y * w); M[7] = 2 * (y * z + x * w); M[8] = 1 - 2 * (x * x + y * y);
      }
      if (jt == LOIK_JOINT_FF || jt == LOIK_JOINT_TRANSLATION) { pq[0] = qj[0]; pq[1] = qj[1]; pq[2] = qj[2]; }
      const double* P = J.plR;
      for (int a = 0; a < 3; ++a) {
        for (int b = 0; b < 3; ++b) st(Pf, FR_XF + 3 * a + b, P[3 * a] * M[b] + P[3 * a + 1] * M[3 + b] + P[3 * a + 2] * M[6 + b]);
        st(Pf, FR_XF + 9 + a, J.plp[a] + (P[3 * a] * pq[0] + P[3 * a + 1] * pq[1] + P[3 * a + 2] * pq[2]));
      }
      continue;
    }
    const double qi = sh[threadIdx.x * nq + c_model.j[i].idxq];
    double a, b;
    if (c_model.j[i].qkind == 1) { b = qi; a = sh[threadIdx.x * nq + c_model.j[i].idxq + 1]; }  // q = (cos, sin), used as given (JointModelRevoluteUnbounded*::calc)
    else if (jt <= 2 || jt == 6) sincos(qi, &a, &b);
    else { a = qi; b = 0.0; }
    double* Pj = joint_blk(T, c_model.off, i - 1);
    st(Pj, JR_JQ, a);
    st(Pj, JR_JQ + 1, b);
    st(Pj, JR_Q, qi);
  }
}

// Outer IK loop: q <- q + dt * z for 1-DoF joints (pinocchio::integrate), followed by
// FwdPassInit of the new configuration (hxx:253-283), all on the device.
__global__ void __launch_bounds__(128) k_integrate(const __grid_constant__ ModelC c_model, const StateP S, const double dt) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S.n) return;
  double* T = tile_ptr(S, c_model, s);
  const int nb = c_model.nb;
  for (int i = 1; i <= nb; ++i) {
    double* Pj = joint_blk(T, c_model.off, i - 1);
    const int jt = c_model.j[i].jtype;
    if (c_model.j[i].nvj > 1) {
      // multi-DoF joints.  Translation (vector space): q += dt z.  Spherical / free-flyer
      // (SpecialOrthogonalOperationTpl<3> / SpecialEuclideanOperationTpl<3>::integrate_impl): quat * exp3(omega), and
      // p + R(quat) * [translation of exp6(v)], body-frame velocities, then quaternion::firstOrderNormalize.
      // liMi = placement * M(q) (FR_XF) is rebuilt for the new configuration.
      const JointC& J = c_model.j[i];
      double* Pf = md_blk(T, c_model.off, J.mblk);
      double M[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, pq[3] = {0, 0, 0};
      if (jt == LOIK_JOINT_TRANSLATION) {
        for (int c = 0; c < 3; ++c) { pq[c] = ld(Pf, FR_Q + c) + dt * ld(Pf, FR_Z + c); st(Pf, FR_Q + c, pq[c]); }
      } else if (jt == LOIK_JOINT_SPHERICAL_ZYX) {  // three Euler angles: a vector space; M and S follow the new configuration
        double qn[3], E[9];
        for (int c = 0; c < 3; ++c) { qn[c] = ld(Pf, FR_Q + c) + dt * ld(Pf, FR_Z + c); st(Pf, FR_Q + c, qn[c]); }
        zyx_calc(qn[0], qn[1], qn[2], M, E);
        for (int c = 0; c < 9; ++c) st(Pf, FR_S + c, E[c]);
      } else if (jt == LOIK_JOINT_PLANAR) {
        // SpecialEuclideanOperationTpl<2>::integrate_impl: (R0, t0) * exp(v): t = vcross - R vcross with vcross = (-vy, vx) / omega
        const double c0 = ld(Pf, FR_Q + 2), s0 = ld(Pf, FR_Q + 3);
        const double vx = dt * ld(Pf, FR_Z), vy = dt * ld(Pf, FR_Z + 1), om = dt * ld(Pf, FR_Z + 2);
        double sv, cv;
        sincos(om, &sv, &cv);
        double tx = vx, ty = vy;
        if (fabs(om) > 1e-14) { const double ax = -vy / om, ay = vx / om; tx = ax - (cv * ax - sv * ay); ty = ay - (sv * ax + cv * ay); }
        pq[0] = ld(Pf, FR_Q) + (c0 * tx - s0 * ty); pq[1] = ld(Pf, FR_Q + 1) + (s0 * tx + c0 * ty);
        const double c1 = c0 * cv - s0 * sv, s1 = s0 * cv + c0 * sv;
        st(Pf, FR_Q, pq[0]); st(Pf, FR_Q + 1, pq[1]); st(Pf, FR_Q + 2, c1); st(Pf, FR_Q + 3, s1);
        M[0] = c1; M[1] = -s1; M[3] = s1; M[4] = c1;
      } else {
        const int o = jt == LOIK_JOINT_FF ? 3 : 0;
        const double w0 = dt * ld(Pf, FR_Z + o), w1 = dt * ld(Pf, FR_Z + o + 1), w2 = dt * ld(Pf, FR_Z + o + 2);
        const double ax = ld(Pf, FR_Q + o), ay = ld(Pf, FR_Q + o + 1), az = ld(Pf, FR_Q + o + 2), aw = ld(Pf, FR_Q + o + 3);
        const double th2 = w0 * w0 + w1 * w1 + w2 * w2, th = sqrt(th2);
        const bool small = th < 1e-4;
        const double k = small ? 0.5 - th2 / 48.0 : sin(th / 2) / th, cw = small ? 1.0 - th2 / 8.0 : cos(th / 2);
        const double d0 = k * w0, d1 = k * w1, d2 = k * w2;
        double r0 = aw * d0 + ax * cw + ay * d2 - az * d1, r1 = aw * d1 - ax * d2 + ay * cw + az * d0;
        double r2 = aw * d2 + ax * d1 - ay * d0 + az * cw, r3 = aw * cw - ax * d0 - ay * d1 - az * d2;
        if (jt == LOIK_JOINT_FF) {
          const double v0 = dt * ld(Pf, FR_Z), v1 = dt * ld(Pf, FR_Z + 1), v2 = dt * ld(Pf, FR_Z + 2);
          const double a_v = small ? 1.0 - th2 / 6.0 : sin(th) / th, a_wxv = small ? 0.5 - th2 / 24.0 : (1.0 - cos(th)) / th2;
          const double a_w = (small ? 1.0 / 6.0 - th2 / 120.0 : (1.0 - a_v) / th2) * (w0 * v0 + w1 * v1 + w2 * v2);
          const double p0 = a_v * v0 + a_w * w0 + a_wxv * (w1 * v2 - w2 * v1), p1 = a_v * v1 + a_w * w1 + a_wxv * (w2 * v0 - w0 * v2);
          const double p2 = a_v * v2 + a_w * w2 + a_wxv * (w0 * v1 - w1 * v0);
          const double Rq[9] = {1 - 2 * (ay * ay + az * az), 2 * (ax * ay - az * aw), 2 * (ax * az + ay * aw), 2 * (ax * ay + az * aw),
                                1 - 2 * (ax * ax + az * az), 2 * (ay * az - ax * aw), 2 * (ax * az - ay * aw), 2 * (ay * az + ax * aw),
                                1 - 2 * (ax * ax + ay * ay)};
          for (int c = 0; c < 3; ++c) { pq[c] = ld(Pf, FR_Q + c) + (Rq[3 * c] * p0 + Rq[3 * c + 1] * p1 + Rq[3 * c + 2] * p2); st(Pf, FR_Q + c, pq[c]); }
          if (r0 * ax + r1 * ay + r2 * az + r3 * aw < 0.0) { r0 = -r0; r1 = -r1; r2 = -r2; r3 = -r3; }
        }
        const double nrm = (3.0 - (r0 * r0 + r1 * r1 + r2 * r2 + r3 * r3)) / 2.0;
        const double x = r0 * nrm, y = r1 * nrm, z = r2 * nrm, w = r3 * nrm;
        st(Pf, FR_Q + o, x); st(Pf, FR_Q + o + 1, y); st(Pf, FR_Q + o + 2, z); st(Pf, FR_Q + o + 3, w);
        M[0] = 1 - 2 * (y * y + z * z); M[1] = 2 * (x * y - z * w); M[2] = 2 * (x * z + y * w);
        M[3] = 2 * (x * y + z * w); M[4] = 1 - 2 * (x * x + z * z); M[5] = 2 * (y * z - x * w);
        M[6] = 2 * (x * z - y * w); M[7] = 2 * (y * z + x * w); M[8] = 1 - 2 * (x * x + y * y);
      }
      const double* P = J.plR;
      for (int a = 0; a < 3; ++a) {
        for (int b = 0; b < 3; ++b) st(Pf, FR_XF + 3 * a + b, P[3 * a] * M[b] + P[3 * a + 1] * M[3 + b] + P[3 * a + 2] * M[6 + b]);
        st(Pf, FR_XF + 9 + a, J.plp[a] + (P[3 * a] * pq[0] + P[3 * a + 1] * pq[1] + P[3 * a + 2] * pq[2]));
      }
      continue;
    }
    if (c_model.j[i].qkind == 1) {
      // SpecialOrthogonalOperationTpl<2>::integrate_impl: rotate (cos, sin) by omega = dt z, then the first-order
      // renormalisation out *= (3 - |out|^2) / 2
      const double ca = ld(Pj, JR_JQ + 1), sa = ld(Pj, JR_JQ);
      double so, co;
      sincos(dt * ld(Pj, JR_Z), &so, &co);
      double c = co * ca - so * sa, s_ = so * ca + co * sa;
      const double k = (3.0 - (c * c + s_ * s_)) / 2.0;
      c *= k; s_ *= k;
      st(Pj, JR_Q, c);
      st(Pj, JR_JQ, s_);
      st(Pj, JR_JQ + 1, c);
      continue;
    }
    const double qi = ld(Pj, JR_Q) + dt * ld(Pj, JR_Z);
    double a, b;
    if (jt <= 2 || jt == 6) sincos(qi, &a, &b);
    else { a = qi; b = 0.0; }
    st(Pj, JR_Q, qi);
    st(Pj, JR_JQ, a);
    st(Pj, JR_JQ + 1, b);
  }
}

// UpdateEqConstraints (ik-id-description-optimized.hpp:127-171), per-instance part: b, Atb = A^T b, |b|inf, and -- when the
// batch does not share the task matrices (Ain != nullptr: [n][nc][36], or [n][36] for one task) -- A and AtA = A^T A
// (:162).  task < 0: all tasks, bis_inf_norm reset; task >= 0: UpdateEqConstraint for that slot (:178-218), norm only grows.
__global__ void k_set_b(const __grid_constant__ ModelC c_model, const StateP S, const double* __restrict__ b, const int per_instance, const int task,
                        const double* __restrict__ Ain, const int Ain_per_instance) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S.n) return;
  double* T = tile_ptr(S, c_model, s);
  const Offs& O = c_model.off;
  const int nc = c_model.nc;
  double* G = glob_blk(T, O);
  double binf = task < 0 ? 0.0 : ld(G, GR_BINF);
  const int k0 = task < 0 ? 0 : task, k1 = task < 0 ? nc : task + 1;
  for (int k = k0; k < k1; ++k) {
    double* Pk = task_blk(T, O, k);
    double bk[6];
    for (int a = 0; a < 6; ++a) {
      const size_t src = task < 0 ? (per_instance ? ((size_t)s * nc + k) * 6 + a : (size_t)k * 6 + a)
                                  : (per_instance ? (size_t)s * 6 + a : (size_t)a);
      bk[a] = b[src];
      st(Pk, TR_B + a, bk[a]);
      binf = fmax(binf, fabs(bk[a]));
    }
    if (Ain) {
      // (Ain_per_instance == 0: one set of matrices for the batch, kept in the rows because there are more tasks than TaskC slots)
      const double* As = Ain + (task < 0 ? ((size_t)(Ain_per_instance ? s : 0) * nc + k) * 36 : (size_t)(Ain_per_instance ? s : 0) * 36);
      for (int i = 0; i < 36; ++i) st(Pk, TR_A + i, As[i]);
      for (int i = 0; i < 6; ++i)
        for (int j = i; j < 6; ++j) {
          double acc = 0.0;
          for (int r = 0; r < 6; ++r) acc += As[6 * r + i] * As[6 * r + j];  // (same order as the host's A^T A of the shared case)
          // packed like TaskC: LL (sym, 6), LA (9), AA (sym, 6)
          if (j < 3) st(Pk, TR_ATA + si(i, j), acc);
          else if (i >= 3) st(Pk, TR_ATA + 15 + si(i - 3, j - 3), acc);
          else st(Pk, TR_ATA + 6 + 3 * i + (j - 3), acc);
        }
    }
    for (int a = 0; a < 6; ++a) {
      double acc = 0.0;
      for (int r = 0; r < 6; ++r) acc += task_A(c_model, c_model.t[k < kMaxTasks ? k : 0], Pk, 6 * r + a) * bk[r];
      st(Pk, TR_ATB + a, acc);
    }
  }
  st(G, GR_BINF, binf);
}

// UpdateIneqConstraints (ik-id-description-optimized.hpp:325-339), per-instance part: lb / ub [n][nv] (per_instance), or the
// batch-shared [nv] replicated into the rows of the multi-DoF joints (their bounds are always read from rows; the shared
// bounds of 1-DoF joints travel in the parameter block)
__global__ void k_set_bounds(const __grid_constant__ ModelC c_model, const StateP S, const double* __restrict__ lb, const double* __restrict__ ub,
                             const int per_instance) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S.n) return;
  double* T = tile_ptr(S, c_model, s);
  const int nb = c_model.nb, nv = c_model.nv;
  const size_t o = per_instance ? (size_t)s * nv : 0;
  for (int i = 1; i <= nb; ++i) {
    if (c_model.j[i].nvj > 1) {
      double* Pf = md_blk(T, c_model.off, c_model.j[i].mblk);
      for (int c = 0; c < c_model.j[i].nvj; ++c) { st(Pf, FR_LB + c, lb[o + c_model.j[i].idxv + c]); st(Pf, FR_UB + c, ub[o + c_model.j[i].idxv + c]); }
      continue;
    }
    if (!per_instance) continue;
    double* Pj = joint_blk(T, c_model.off, i - 1);
    st(Pj, JR_LB, lb[o + c_model.j[i].idxv]);
    st(Pj, JR_UB, ub[o + c_model.j[i].idxv]);
  }
}

// problem_.UpdateReferences with a v_ref of its own for every instance (v_refs [n][njoints][6]): Hv_i = H_ref_i v_ref_i
// (ik-id-description-optimized.hpp:113) into rows JR_HV of the joint blocks; H_ref_i from the reference table, or (H_refs)
// this instance's own weights
__global__ void k_set_vref(const __grid_constant__ ModelC c_model, const StateP S, const double* __restrict__ v_refs, const double* __restrict__ H_refs) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S.n) return;
  double* T = tile_ptr(S, c_model, s);
  const int nb = c_model.nb;
  for (int i = 1; i <= nb; ++i) {
    const HrefC& Hr = c_model.href[c_model.j[i].href];
    const double* vr = v_refs + ((size_t)s * (nb + 1) + i) * 6;
    double v[6], A[6], B[9], D[6];
    for (int c = 0; c < 6; ++c) v[c] = vr[c];
    double* Pj = joint_blk(T, c_model.off, i - 1);
    if (H_refs) {  // this instance's own weights [n][njoints][36], symmetric: the LL, LA, AA blocks go to rows JR_HREF
      const double* H = H_refs + ((size_t)s * (nb + 1) + i) * 36;
      for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) {
          if (a <= b) { A[si(a, b)] = H[6 * a + b]; D[si(a, b)] = H[6 * (3 + a) + 3 + b]; }
          B[3 * a + b] = H[6 * a + 3 + b];
        }
      for (int c = 0; c < 6; ++c) { st(Pj, JR_HREF + c, A[c]); st(Pj, JR_HREF + 15 + c, D[c]); }
      for (int c = 0; c < 9; ++c) st(Pj, JR_HREF + 6 + c, B[c]);
    } else {
      for (int c = 0; c < 6; ++c) { A[c] = Hr.A[c]; D[c] = Hr.D[c]; }
      for (int c = 0; c < 9; ++c) B[c] = Hr.B[c];
    }
    for (int a = 0; a < 3; ++a) {
      st(Pj, JR_HV + a, A[si(a, 0)] * v[0] + A[si(a, 1)] * v[1] + A[si(a, 2)] * v[2] + B[3 * a] * v[3] + B[3 * a + 1] * v[4] + B[3 * a + 2] * v[5]);
      st(Pj, JR_HV + 3 + a, B[a] * v[0] + B[3 + a] * v[1] + B[6 + a] * v[2] + D[si(a, 0)] * v[3] + D[si(a, 1)] * v[4] + D[si(a, 2)] * v[5]);
    }
  }
}

// batch-major gather of `nrows` rows: dst[s][k] = tile(s)[map[k]][lane(s)]  (map = absolute row indices).
// (`base` / `tile_rows`: the arena and its rows per tile -- the state arena, or the debug arena for the residual vectors)
__global__ void __launch_bounds__(256) k_gather(const __grid_constant__ ModelC c_model, const StateP S, const int nrows, const int* __restrict__ map,
                                                double* __restrict__ dst, const double* __restrict__ base, const int tile_rows) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)S.n * nrows) return;
  const int s = (int)(idx / nrows), k = (int)(idx % nrows);
  dst[idx] = map[k] < 0 ? 0.0 : base[((size_t)(s >> 5) * tile_rows + map[k]) * 32 + (s & 31)];
}
__global__ void k_gather_limi(const __grid_constant__ ModelC c_model, const StateP S, double* __restrict__ dst) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S.n) return;
  const double* T = tile_ptr(S, c_model, s);
  const int nb = c_model.nb;
  for (int i = 1; i <= nb; ++i) {
    double R[9], t[3];
    const double* Pj = joint_blk(const_cast<double*>(T), c_model.off, i - 1);
    if (c_model.j[i].nvj > 1) md_load_xf(md_blk(const_cast<double*>(T), c_model.off, c_model.j[i].mblk), R, t);
    else make_xf(c_model.j[i], ld(Pj, JR_JQ), ld(Pj, JR_JQ + 1), R, t);
    double* o = dst + ((size_t)s * nb + (i - 1)) * 12;
    for (int c = 0; c < 9; ++c) o[c] = R[c];
    for (int c = 0; c < 3; ++c) o[9 + c] = t[c];
  }
}
// which: 0 = iteration count, 1 = status flags (bit0 converged, bit1 primal infeasible, bit2 max_iter)
__global__ void k_gather_ctl(const __grid_constant__ ModelC c_model, const StateP S, const int which, int* __restrict__ dst) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S.n) return;
  const int2 c = ld_ctl(c_model, tile_ptr(S, c_model, s));
  const int stt = c.x;
  dst[s] = which == 0 ? c.y
                      : (((stt == ST_CONVERGED || stt == ST_CONVERGED_PINF) ? 1 : 0) | ((stt == ST_TAIL || stt == ST_INFEASIBLE_DONE || stt == ST_CONVERGED_PINF) ? 2 : 0) |
                         (stt == ST_MAXITER ? 4 : 0));
}
// out[0..2] = #converged, #infeasible, #maxiter; out[3] = sum iters
__global__ void k_stats(const __grid_constant__ ModelC c_model, const StateP S, unsigned long long* __restrict__ out) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  int stt = -1, it = 0;
  if (s < S.n) { const int2 c = ld_ctl(c_model, tile_ptr(S, c_model, s)); stt = c.x; it = c.y; }
  const unsigned c = __ballot_sync(0xffffffffu, stt == ST_CONVERGED || stt == ST_CONVERGED_PINF);  // (how the solve ended)
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
  int device = 0, batch = 0, ntiles = 0;
  int nj = 0, nb = 0, nc = 0, npend = 0, nv = 0, nq = 0;
  loik_params prm{};
  ModelC mc{};          // host copy of this solver's constant block
  bool problem_set = false;
  bool a_user_per = false;  // the caller's task matrices are per instance (loik_solve_init)
  bool debug = false;
  // backward->forward workspace (His, pis, UDinv, Dinv, r) of the home arena is the last backward pass of every
  // instance: true after in-place steps, false after a scheduled solve whose instances migrated without keep_ws
  bool ws_valid = true;
  double* arena = nullptr;  // tile records
  int* d_lists = nullptr;   // two compaction lists of `batch` ints
  double* scratch[2] = {nullptr, nullptr};  // packed arenas (allocated at the first solve)
  int* d_origin = nullptr;  // [2][batch] home slot of every packed slot
  bool debug_by_logging = false;  // debug mode was switched on by loik_set_logging (and goes off with it)
  double* d_hist = nullptr;  // solver log (loik_set_logging): [batch][hist_cap][kHistCols]
  int hist_cap = 0;
  int4* d_wide_tab = nullptr;  // step table of the wide sweeps of k_iterate_lane<4> (build_wide_table)
  int* d_counts = nullptr;  // [0],[1]: list lengths (ping-pong); [2]: n_active; [3]: work-queue head of the lane kernel; [5]: entries queued from the end of its list
  unsigned long long* d_stats = nullptr;
  int* h_counts = nullptr;  // pinned
  unsigned long long* h_stats = nullptr;
  int* d_map = nullptr;  // row maps of every loik_field (built once)
  int map_off[32] = {0}, map_len[32] = {0};
  StateP S{};
  // staging
  void* h_stage = nullptr; size_t h_stage_bytes = 0;
  void* d_stage = nullptr; size_t d_stage_bytes = 0;
  int64_t launches = 0;
  int64_t sweeps = 0;
  int dense_sweeps = 4;  // sweeps on the home arena before the first re-pack
  int sched_reps = 2;      // re-pack rounds per chunk size
  double sched_growth = 2.0;  // chunk growth factor
  int last_list = -1;  // index of the list holding the most recent compaction, -1 = none
  // CUDA-graph cache of (reset +) the launch schedule: one graph launch per solve instead of ~60 kernel launches
  cudaGraphExec_t g_exec = nullptr;
  ModelC g_mc{};
  int g_flags = -1, g_budget = -1, g_dense = -1, g_keep = -1;  // (loik_set_schedule drops the cached graph)
  int64_t g_launches = 0, g_sweeps = 0;
  bool use_graph = true;
  // the latency-bound tail rounds run on a high-priority stream so their few CTAs are dispatched ahead of the
  // bulk kernels of other solvers sharing the GPU
  cudaStream_t hi_stream = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  int small_after = 1 << 30, small_grid = 296;  // late rounds: at most small_grid CTAs, grid-stride
  int hi_after = 8;  // sweeps after which the schedule moves to the high-priority stream (<0: never)
  int seg_after = 32;  // sweeps after which a branching tree is swept by the segment-parallel kernel
  int seg_warps = 0;   // warps per tile of that kernel: 0 = as many as the tree has parallel chains (at most 4), 1 = never use it
  // lane-parallel shared-memory kernel (loik_lane.cuh): finishes every instance still active after `lane_after` sweeps
  // (0: the whole solve; < 0: never).  Geometry fixed at creation from the model's record size.
  int lane_after = -1;
  bool lane_ok = false;
  int lane_warps_req = 0;  // warps per CTA (0 = chosen from the record size)
  int lane_gpi_req = 0;    // groups of 8 lanes per instance: 1, 4, or 0 = default (4 for branching trees)
  double hard_ratio = 20.0;  // hand-over to the lane kernel: residual / tolerance beyond which an instance queues up first (0: no ordering)
  bool drop_ws = true;     // the tile kernels drop the consumed backward->forward workspace from L2 instead of writing it back
  int sms = 0, smem_optin = 0, smem_sm = 0;
  int sweeps_in_solve = 0;
  std::vector<int> model_parents, model_types;  // kept for loik_set_schedule (segment re-assignment)
};


static inline int grid_for(int n, int block = kBlock) { return (n + block - 1) / block; }

// One place that launches the fused iteration kernel.
// `seg`: use the segment-parallel kernel (several warps per tile) when the tree branches.  It shortens the critical
// path of a sweep (latency) but spends 2-3.5x the warp-slot time of the one-warp-per-tile kernel per iteration, so
// the schedule switches to it only for the late rounds, when few tiles are left and latency is all that matters.
static void launch_iterate(loik_solver* h, cudaStream_t st, const StateP& S_in, int iters, int fixed, int max_ctas = 0, bool seg = true) {
  StateP S = S_in;
  S.drop_ws = (h->drop_ws && !S.keep_ws && !h->debug) ? 1 : 0;
  if (S.drop_ws) h->ws_valid = false;  // the consumed workspace lines are dropped from L2: undefined until the next backward sweep
  int g = grid_for(h->batch);
  if (max_ctas > 0) g = std::min(g, max_ctas);
  const bool md = h->mc.nmd > 0 || h->mc.vref_per;  // multi-DoF joints or per-instance references: the general instantiations (span-level dispatch)
  if (seg && h->mc.nwarp > 1 && !h->debug) {  // segment-parallel: one CTA (nwarp warps) per tile
    int gt = h->ntiles;
    if (max_ctas > 0) gt = std::min(gt, max_ctas);
#define LOIK_LAUNCH_SEG(NW_)                                                                       \
  do {                                                                                             \
    if (md) k_iterate_seg<false, NW_, true><<<gt, 32 * NW_, 0, st>>>(h->mc, S, iters, fixed);      \
    else k_iterate_seg<false, NW_, false><<<gt, 32 * NW_, 0, st>>>(h->mc, S, iters, fixed);        \
  } while (0)
    switch (h->mc.nwarp) {
      case 2: LOIK_LAUNCH_SEG(2); break;
      case 3: LOIK_LAUNCH_SEG(3); break;
      default: LOIK_LAUNCH_SEG(4); break;
    }
#undef LOIK_LAUNCH_SEG
    h->launches++;
    return;
  }
#if defined(LOIK_PF_EXT) || defined(LOIK_L1_CARVEOUT)  // (LOIK_PF_EXT: set in loik_device.cuh unless LOIK_PF_BASIC)
  {  // all of the unified L1/shared array as L1 (k_iterate uses no shared memory)
    static bool done[64] = {false};
    if (h->device < 64 && !done[h->device]) {
      done[h->device] = true;
      cudaFuncSetAttribute(k_iterate<false, 4, false>, cudaFuncAttributePreferredSharedMemoryCarveout, 0);
      cudaFuncSetAttribute(k_iterate<false, 4, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 0);
    }
  }
#endif
#define LOIK_LAUNCH(DBG, MB, MD_) k_iterate<DBG, MB, MD_><<<g, kBlock, 0, st>>>(h->mc, S, iters, fixed)
  if (h->debug) { if (md) LOIK_LAUNCH(true, 4, true); else LOIK_LAUNCH(true, 4, false); }
  else if (md) { LOIK_LAUNCH(false, 4, true); }
  else { LOIK_LAUNCH(false, 4, false); }
#undef LOIK_LAUNCH
  h->launches++;
}

// Geometry of the lane-parallel kernel for the current problem constants: W warps per CTA = 4 W instance records + the
// per-CTA constants in shared memory.  Unless the caller fixed W (loik_schedule.lane_warps_per_cta): the smallest CTA that
// keeps >= 90 % of the warps an SM's shared memory can hold -- small CTAs let the few CTAs that end up with a straggler
// block less of an SM for the kernels of other solvers.
struct LaneGeom { int gpi = 1, warps = 0, ctas_per_sm = 0, grid = 0; size_t smem = 0; };
static bool lane_geometry(const loik_solver* h, LaneGeom& G) {
  // groups per instance: the four groups of a warp on different chains of one instance (wide sweeps, loik_lane.cuh) when the
  // tree branches (Talos: 12 steps per sweep instead of 32, 18.6 vs 31.4 us per iteration of a lone instance), else one
  // group per instance (on a serial chain the wide sweeps only add their bookkeeping: Panda 9.9 vs 6.9 us)
  const int gpi = h->lane_gpi_req > 0 ? h->lane_gpi_req : (h->mc.nwarp >= 2 ? 4 : 1);
  const LaneDims D = lane_dims(h->nb, h->nc, h->npend, h->mc.href_uniform, gpi, h->mc.a_per, h->mc.nsb + h->mc.nsf);
  auto ctas_of = [&](int W) -> int {
    const size_t bytes = lane_smem_bytes(D, W, gpi);
    if (bytes > (size_t)h->smem_optin) return 0;
    return std::min(32, (int)((size_t)h->smem_sm / (bytes + 1024)));
  };
  int W = h->lane_warps_req;
  if (W <= 0) {
    int best = 0;
    for (int w = 1; w <= 8; ++w) best = std::max(best, ctas_of(w) * w);
    for (int w = 1; w <= 8 && W <= 0; ++w)
      if (ctas_of(w) > 0 && 10 * ctas_of(w) * w >= 9 * best) W = w;
  }
  if (W <= 0 || W > 8 || ctas_of(W) == 0) return false;
  const int per_cta = (kLaneI / gpi) * W;
  G.gpi = gpi; G.warps = W; G.ctas_per_sm = ctas_of(W); G.smem = lane_smem_bytes(D, W, gpi);
  G.grid = std::min(h->sms * G.ctas_per_sm, (h->batch + per_cta - 1) / per_cta);
  return true;
}
// The lane-parallel kernel on the instances of `src` (all `batch` slots, or the `*n_list` slots named by `list`);
// `origin`: home slot of every slot of a packed arena.  Runs every instance to the end of its solve (or `iters`
// iterations each in fixed mode) and sends the results to the home arena.
static int launch_lane(loik_solver* h, cudaStream_t st, const double* src, const int* list, const int* n_list, const int* origin,
                       int fixed, int iters, const int* n_back = nullptr) {
  LaneGeom G;
  if (!lane_geometry(h, G)) return fail(LOIK_ERR_STATE, "lane-parallel kernel: the instance record does not fit shared memory");
  LaneP P{};
  P.src = src; P.list = list; P.n_list = n_list; P.n_back = n_back; P.cap = h->batch; P.n = h->batch; P.origin = origin; P.home = h->arena;
  P.queue = h->d_counts + 3; P.iters = iters; P.fixed = fixed; P.keep_ws = h->S.keep_ws; P.tab = h->d_wide_tab;
  CK(cudaMemsetAsync(h->d_counts + 3, 0, sizeof(int), st));
  if (G.gpi == 1) k_iterate_lane<1><<<G.grid, 32 * G.warps, G.smem, st>>>(h->mc, P);
  else k_iterate_lane<4><<<G.grid, 32 * G.warps, G.smem, st>>>(h->mc, P);
  h->launches++;
  return LOIK_OK;
}
static bool use_lane(const loik_solver* h) { return h->lane_ok && h->lane_after >= 0 && !h->debug && !h->mc.vref_per; }  // (the lane record has no rows for per-instance references)

// FwdPassInit stages q through shared memory (blockDim x nq doubles): the block shrinks for models whose q would not fit 48 KB
static void launch_set_q(loik_solver* h, cudaStream_t st, const double* dq) {
  int block = kBlock;
  while (block > 32 && (size_t)block * h->nq * sizeof(double) > 48 * 1024) block /= 2;
  k_set_q<<<grid_for(h->batch, block), block, (size_t)block * h->nq * sizeof(double), st>>>(h->mc, h->S, dq);
}

static int ensure_stage(loik_solver* h, size_t bytes, bool need_host) {
  if (need_host && bytes > h->h_stage_bytes) {
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
  if (loc == LOIK_HOST_PINNED) {  // caller's buffer is page-locked: DMA straight from it, no host-side copy, no sync
    CK(cudaMemcpyAsync((char*)h->d_stage + off, src, bytes, cudaMemcpyHostToDevice, st));
    *out = (char*)h->d_stage + off;
    return LOIK_OK;
  }
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

// Maximal register-carried chains of the tree = the segments different warps can sweep (k_iterate_seg); `max_warps`:
// 0 = as many warps per tile as the tree has parallel chains (at most 4), 1 = one segment, one warp (k_iterate only).
static void assign_segments(ModelC& M, int max_warps) {
  const int nj = M.nj;
  std::vector<int> seg_of(nj, -1), lo, hi;
  for (int i = 1; i < nj; ++i) {
    if (!M.j[i].carry) { lo.push_back(i); hi.push_back(i); }
    else hi.back() = i;
    seg_of[i] = (int)lo.size() - 1;
  }
  const int ns = (int)lo.size();
  M.nseg = 1; M.nwarp = 1; M.nblevel = 1; M.nflevel = 1;
  std::memset(M.seg, 0, sizeof(M.seg));
  M.seg[0] = SegC{1, (short)(nj - 1), 0, 0, 0, 0};
  if (ns < 2 || ns > kMaxSeg || max_warps == 1) return;
  std::vector<int> fl(ns, 0), bl(ns, 0), par(ns, -1);
  for (int g = 0; g < ns; ++g) {
    const int p = M.j[lo[g]].parent;
    if (p > 0) { par[g] = seg_of[p]; fl[g] = fl[par[g]] + 1; }
  }
  for (int g = ns - 1; g >= 0; --g)
    if (par[g] >= 0) bl[par[g]] = std::max(bl[par[g]], bl[g] + 1);
  int nbl = 0, nfl = 0, width = 1;
  for (int g = 0; g < ns; ++g) { nbl = std::max(nbl, bl[g] + 1); nfl = std::max(nfl, fl[g] + 1); }
  for (int lv = 0; lv < std::max(nbl, nfl); ++lv) {
    int cb = 0, cf = 0;
    for (int g = 0; g < ns; ++g) { cb += bl[g] == lv; cf += fl[g] == lv; }
    width = std::max(width, std::max(cb, cf));
  }
  int NW = std::min(4, width);
  if (max_warps > 1) NW = std::min(NW, max_warps);
  if (NW < 2) return;
  M.nseg = ns; M.nwarp = NW; M.nblevel = nbl; M.nflevel = nfl;
  // longest-processing-time-first assignment of the segments of every level to the warps
  auto assign = [&](const std::vector<int>& level, int nlev, bool backward) {
    for (int lv = 0; lv < nlev; ++lv) {
      std::vector<int> ids;
      for (int g = 0; g < ns; ++g) if (level[g] == lv) ids.push_back(g);
      std::sort(ids.begin(), ids.end(), [&](int a, int b) { return (hi[a] - lo[a]) > (hi[b] - lo[b]); });
      std::vector<int> load(NW, 0);
      for (int g : ids) {
        const int wmin = (int)(std::min_element(load.begin(), load.end()) - load.begin());
        load[wmin] += hi[g] - lo[g] + 1;
        if (backward) M.seg[g].bwarp = (short)wmin; else M.seg[g].fwarp = (short)wmin;
      }
    }
  };
  for (int g = 0; g < ns; ++g) { M.seg[g].lo = (short)lo[g]; M.seg[g].hi = (short)hi[g]; M.seg[g].blevel = (short)bl[g]; M.seg[g].flevel = (short)fl[g]; }
  assign(bl, nbl, true);
  assign(fl, nfl, false);
}

// Step table of the wide sweeps of k_iterate_lane<4> (loik_lane.cuh: the four 8-lane groups of a warp on different chains of
// one instance): per sweep direction a flat list of steps, four entries (WideStep) per step.  The chains are the segments of
// the tree (ModelC::seg); a chain may start once the chains it depends on have ended (towards the root: every chain hanging
// off it; on the way out: the chain of its parent joint) -- a step's stores are separated from the next step's loads by
// the __syncwarp() at the top of a step, so "ended in an earlier step" is all the ordering needed.  List scheduling on four
// groups, longest remaining path first (Talos: 10 steps per sweep, arms 8 -> torso 2, for 32 joints).  A group without work
// in a step gets an entry without WF_VALID on joint 1.  Also fixes the blocks of the joints in the shared-memory record
// (JointC::loff): joint i starts 4 x (group that sweeps it towards the root) doubles (mod 16) into its 108-double slot.
static void build_wide_table(ModelC& M, std::vector<int4>& tab) {
  tab.clear();
  // the chains: maximal runs of joints whose contribution travels in registers (JointC::carry), as in assign_segments but
  // without its cap on their number
  std::vector<int> clo, chi, seg_of(M.nj, 0);
  for (int i = 1; i < M.nj; ++i) {
    if (!M.j[i].carry) { clo.push_back(i); chi.push_back(i); }
    else chi.back() = i;
    seg_of[i] = (int)clo.size() - 1;
  }
  const int ns = (int)clo.size();
  std::vector<int> len(ns), par(ns, -1);
  for (int c = 0; c < ns; ++c) {
    len[c] = chi[c] - clo[c] + 1;
    const int p = M.j[clo[c]].parent;
    if (p > 0) par[c] = seg_of[p];
  }
  struct Slot { int start, group; };
  auto schedule = [&](const bool backward, std::vector<Slot>& slot) {
    // remaining path: backward = the chain and its ancestors, forward = the chain and its longest descendant line
    std::vector<int> prio(ns, 0);
    if (backward) { for (int c = 0; c < ns; ++c) prio[c] = len[c] + (par[c] >= 0 ? prio[par[c]] : 0); }  // (parents have smaller indices)
    else { for (int c = ns - 1; c >= 0; --c) { prio[c] += len[c]; if (par[c] >= 0) prio[par[c]] = std::max(prio[par[c]], prio[c]); } }
    slot.assign(ns, Slot{-1, -1});
    int free_at[4] = {0, 0, 0, 0}, done = 0, nsteps = 0;
    for (int t = 0; done < ns; ++t) {
      for (int g = 0; g < 4; ++g) {
        if (free_at[g] > t) continue;
        int best = -1;
        for (int c = 0; c < ns; ++c) {
          if (slot[c].start >= 0) continue;
          bool ready = true;
          if (backward) { for (int k = 0; k < ns; ++k) if (par[k] == c && (slot[k].start < 0 || slot[k].start + len[k] > t)) ready = false; }
          else if (par[c] >= 0) ready = slot[par[c]].start >= 0 && slot[par[c]].start + len[par[c]] <= t;
          if (ready && (best < 0 || prio[c] > prio[best])) best = c;
        }
        if (best < 0) continue;
        slot[best] = Slot{t, g};
        free_at[g] = t + len[best];
        nsteps = std::max(nsteps, free_at[g]);
        ++done;
      }
    }
    return nsteps;
  };
  std::vector<Slot> sb, sf;
  M.nsb = schedule(true, sb);
  M.nsf = schedule(false, sf);
  for (int c = 0; c < ns; ++c)
    for (int i = clo[c]; i <= chi[c]; ++i) M.j[i].loff = (short)(LJ_SLOT_WIDE * (i - 1) + 4 * ((sb[c].group - 3 * (i - 1)) & 3));
  auto emit = [&](const bool backward, const std::vector<Slot>& slot, const int nsteps) {
    const size_t base = tab.size();
    tab.resize(base + 4 * (size_t)nsteps, make_int4(1, 0, (int)M.j[1].loff, 0));
    for (int c = 0; c < ns; ++c)
      for (int t = 0; t < len[c]; ++t) {
        const int i = backward ? chi[c] - t : clo[c] + t;
        const JointC& J = M.j[i];
        int flags = WF_VALID | (t == 0 ? WF_FIRST : 0) | (J.parent == 0 ? WF_ROOT : 0) | (J.npin > 0 ? WF_PINS : 0);
        if (J.parent > 0 && !J.carry) flags |= WF_GIVE;
        tab[base + 4 * (size_t)(slot[c].start + t) + slot[c].group] =
            make_int4(i | (J.parent << 16), flags, (int)J.loff | ((int)J.pout << 16), ((int)J.sidx & 0xffff) | ((J.parent > 0 ? (int)M.j[J.parent].loff : 0) << 16));
      }
  };
  emit(true, sb, M.nsb);
  emit(false, sf, M.nsf);
}
static int upload_wide_table(loik_solver* h) {
  std::vector<int4> tab;
  build_wide_table(h->mc, tab);
  if (h->d_wide_tab) { cudaFree(h->d_wide_tab); h->d_wide_tab = nullptr; }
  CK(cudaMalloc(&h->d_wide_tab, tab.size() * sizeof(int4)));
  CK(cudaMemcpy(h->d_wide_tab, tab.data(), tab.size() * sizeof(int4), cudaMemcpyHostToDevice));
  return LOIK_OK;
}

extern "C" {

int32_t loik_abi_version(void) { return 2; }
const char* loik_last_error(void) { return g_err.c_str(); }

// `dry`: stop after the host-side part (validation, tree bookkeeping, tile-record layout) and hand back a handle that
// owns no CUDA resource -- what loik_model_layout reports, so that this logic is testable on a box without a GPU.
static int create_impl(const loik_model_desc* model, const loik_params* params, int32_t batch, int32_t device, loik_solver** out, bool dry) {
  if (!model || !params || !out) return fail(LOIK_ERR_INVALID, "loik_create: null argument");
  *out = nullptr;
  const int nj = model->njoints;
  // IkProblemFormulationOptimized ctor checks (ik-id-description-optimized.hpp:37-44)
  if (params->eq_c_dim != 6)
    return fail(LOIK_ERR_INVALID, "[IkProblemFormulation::IkProblemFormulation]: equality constraint dimension is not 6, problem formulation not supported !!!");
  if (nj < 2 || nj > LOIK_MAX_JOINTS) return fail(LOIK_ERR_INVALID, "loik_create: njoints out of range [2, LOIK_MAX_JOINTS]");
  static_assert(LOIK_MAX_JOINTS == kMaxJoints && LOIK_MAX_TASKS == kMaxTasksAll, "include/loik_b200.h and loik_device.cuh must agree");
  if (params->num_eq_c < 0 || params->num_eq_c > LOIK_MAX_TASKS) return fail(LOIK_ERR_INVALID, "loik_create: num_eq_c out of range [0, LOIK_MAX_TASKS]");
  if (batch < 1) return fail(LOIK_ERR_INVALID, "loik_create: batch must be >= 1");
  for (int i = 1; i < nj; ++i) {
    if (model->parents[i] < 0 || model->parents[i] >= i) return fail(LOIK_ERR_INVALID, "loik_create: parents[i] must be < i");
    if (model->joint_types[i] < 0 || model->joint_types[i] > LOIK_JOINT_SPHERICAL_ZYX)
      return fail(LOIK_ERR_UNSUPPORTED, "loik_create: unsupported joint type (1-DoF revolute/prismatic joints, free-flyer, spherical, SphericalZYX, translation and planar joints are supported)");
  }
  auto nv_of = [&](int i) { const int t = model->joint_types[i]; return t == LOIK_JOINT_FF ? 6 : ((t == LOIK_JOINT_SPHERICAL || t == LOIK_JOINT_TRANSLATION || t == LOIK_JOINT_PLANAR || t == LOIK_JOINT_SPHERICAL_ZYX) ? 3 : 1); };
  {
    int nmd = 0;
    for (int i = 1; i < nj; ++i) nmd += nv_of(i) > 1;
    if (nmd > kMaxMd) return fail(LOIK_ERR_UNSUPPORTED, "loik_create: more multi-DoF joints than kMaxMd");
  }
  loik_solver* h = new loik_solver();
  h->device = device; h->batch = batch; h->ntiles = (batch + 31) / 32;
  h->nj = nj; h->nb = nj - 1; h->nc = params->num_eq_c; h->prm = *params;
  auto unbounded = [&](int i) { return model->joint_types[i] >= LOIK_JOINT_RUBX && model->joint_types[i] <= LOIK_JOINT_RUBU; };
  auto nq_of = [&](int i) {
    const int t = model->joint_types[i];
    return t == LOIK_JOINT_FF ? 7 : ((t == LOIK_JOINT_SPHERICAL || t == LOIK_JOINT_PLANAR) ? 4 : ((t == LOIK_JOINT_TRANSLATION || t == LOIK_JOINT_SPHERICAL_ZYX) ? 3 : (unbounded(i) ? 2 : 1)));
  };
  h->nv = 0; h->nq = 0;
  for (int i = 1; i < nj; ++i) { h->nv += nv_of(i); h->nq += nq_of(i); }
  ModelC& M = h->mc;
  std::memset(&M, 0, sizeof(M));
  M.nj = nj; M.nb = nj - 1; M.nc = h->nc;
  M.nv = h->nv; M.nq = h->nq;
  M.max_iter = params->max_iter; M.rho = params->rho; M.mu0 = params->mu; M.mu_scale = params->mu_equality_scale_factor;
  M.tol_abs = params->tol_abs; M.tol_rel = params->tol_rel; M.tol_pinf = params->tol_primal_inf; M.tol_dinf = params->tol_dual_inf;
  M.tol_tail = params->tol_tail_solve;
  // tree bookkeeping.  A child's contribution to its parent travels in registers when the parent is the next joint
  // swept and has no other child; every other edge gets its own pending block (single writer, no read-modify-write).
  // Maximal register-carried chains are the segments that different warps can sweep (k_iterate_seg).
  std::vector<int> nchild(nj, 0);
  for (int i = 1; i < nj; ++i) nchild[model->parents[i]]++;
  int npend = 0, idxq = 0, idxv = 0, nmd = 0;
  for (int i = 1; i < nj; ++i) {
    JointC& J = M.j[i];
    J.parent = model->parents[i]; J.jtype = model->joint_types[i]; J.task = -1;
    for (int c = 0; c < 9; ++c) J.plR[c] = model->placement_R[9 * i + c];
    for (int c = 0; c < 3; ++c) { J.plp[c] = model->placement_p[3 * i + c]; J.axis[c] = model->joint_axes[3 * i + c]; }
    // (every edge into or out of a multi-DoF joint goes through a pending block: those joints have their own step)
    J.carry = (J.parent > 0 && J.parent == i - 1 && nchild[J.parent] == 1 && nv_of(J.parent) == 1 && nv_of(i) == 1) ? 1 : 0;
    J.pout = -1; J.npin = 0;
    J.nvj = nv_of(i); J.sel0 = model->joint_types[i] == LOIK_JOINT_FF ? 0x543210 : (model->joint_types[i] == LOIK_JOINT_SPHERICAL ? 0x543 : (model->joint_types[i] == LOIK_JOINT_PLANAR ? 0x510 : 0x210)); J.mblk = J.nvj > 1 ? nmd++ : -1;
    J.idxv = idxv; idxv += J.nvj;
    J.idxq = idxq; idxq += nq_of(i);
    J.qkind = unbounded(i) ? 1 : 0;
    // an unbounded revolute joint is its bounded twin everywhere but in how q enters (k_set_q, k_integrate, the q getter)
    if (unbounded(i)) J.jtype = model->joint_types[i] == LOIK_JOINT_RUBU ? LOIK_JOINT_RU : model->joint_types[i] - LOIK_JOINT_RUBX;
    J.sidx = (J.nvj == 1 && J.jtype <= LOIK_JOINT_PZ) ? (J.jtype <= LOIK_JOINT_RZ ? 3 + J.jtype : J.jtype - LOIK_JOINT_PX) : -1;
  }
  for (int i = 1; i < nj; ++i) {
    JointC& J = M.j[i];
    if (J.parent > 0 && !J.carry) {
      JointC& P = M.j[J.parent];
      if (P.npin >= kMaxPin) { delete h; return fail(LOIK_ERR_UNSUPPORTED, "loik_create: a joint has more branching children than kMaxPin"); }
      J.pout = npend;
      P.pin[P.npin++] = npend;
      ++npend;
    }
  }
  M.npend = npend; h->npend = npend; M.nmd = nmd;
  M.nspan = 0;
  for (int i = 1; i < nj; ++i) {
    if (M.j[i].nvj > 1) M.span[M.nspan++] = SpanC{(short)i, (short)i, (short)M.j[i].nvj, 0};
    else if (M.nspan > 0 && M.span[M.nspan - 1].md == 0) M.span[M.nspan - 1].hi = (short)i;
    else M.span[M.nspan++] = SpanC{(short)i, (short)i, 0, 0};
  }
  assign_segments(M, h->seg_warps);
  { std::vector<int4> tab; build_wide_table(M, tab); }  // (sets nsb / nsf; uploaded once the CUDA side exists)
  // tile record layout: [globals | joint blocks | task blocks | pending blocks | debug vectors]
  const int nb = h->nb, nc = std::max(h->nc, 1);
  Offs& O = M.off;
  int rows = 0;
  O.glob = rows; rows += GR_ROWS;
  O.joint0 = rows; rows += JR_ROWS * nb;
  O.task0 = rows; rows += TR_ROWS * nc;
  O.pend0 = rows; rows += PR_ROWS * std::max(npend, 1);
  O.ff0 = rows; rows += FR_ROWS * nmd;
  O.rows = rows;
  O.prv = 0; O.drv = 6 * nb + h->nv; O.drows = 2 * (6 * nb + h->nv);  // (a separate arena, allocated by loik_set_debug)
  if (dry) { *out = h; return LOIK_OK; }
  {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { delete h; return fail(LOIK_ERR_CUDA, "loik_create: no CUDA device (libloik_b200 has no CPU fallback)"); }
    const cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) { delete h; return fail(LOIK_ERR_CUDA, std::string("cudaSetDevice(device): ") + cudaGetErrorString(e)); }
  }
  // every allocation is checked: on failure the partially built handle is destroyed and the CUDA error reported
#define CKA(call)                                                                                        \
  do {                                                                                                   \
    const cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) {                                                                             \
      loik_destroy(h);                                                                                   \
      return fail(LOIK_ERR_CUDA, std::string("loik_create: " #call ": ") + cudaGetErrorString(e_));      \
    }                                                                                                    \
  } while (0)
  const size_t arena_doubles = (size_t)h->ntiles * rows * 32;
  CKA(cudaMalloc(&h->arena, arena_doubles * sizeof(double)));
  CKA(cudaMemset(h->arena, 0, arena_doubles * sizeof(double)));
  if (upload_wide_table(h) != LOIK_OK) { loik_destroy(h); return LOIK_ERR_CUDA; }
  CKA(cudaMalloc(&h->d_lists, 2 * (size_t)batch * sizeof(int)));
  CKA(cudaMalloc(&h->d_counts, 8 * sizeof(int)));
  CKA(cudaMemset(h->d_counts, 0, 8 * sizeof(int)));
  CKA(cudaMalloc(&h->d_stats, 4 * sizeof(unsigned long long)));
  {  // row maps of the gettable fields: field -> absolute rows of the tile record, in output order
    std::vector<int> all;
    const int ncq = h->nc;
    auto per_joint = [&](std::vector<int>& m, int jr, int width) { for (int j = 0; j < nb; ++j) for (int c = 0; c < width; ++c) m.push_back(O.joint0 + JR_ROWS * j + jr + c); };
    // one row per dof: a multi-DoF joint contributes nv rows of its own block (fr < 0: no such quantity -> zeros)
    auto per_dof = [&](std::vector<int>& m, int jr, int fr) {
      for (int j = 0; j < nb; ++j) {
        if (M.j[j + 1].nvj > 1) { for (int c = 0; c < M.j[j + 1].nvj; ++c) m.push_back(fr < 0 ? -1 : O.ff0 + FR_ROWS * M.j[j + 1].mblk + fr + c); }
        else m.push_back(O.joint0 + JR_ROWS * j + jr);
      }
    };
    auto per_joint_noff = [&](std::vector<int>& m, int jr, int width) {
      for (int j = 0; j < nb; ++j) for (int c = 0; c < width; ++c) m.push_back(M.j[j + 1].nvj > 1 ? -1 : O.joint0 + JR_ROWS * j + jr + c);
    };
    auto per_task = [&](std::vector<int>& m, int tr) { for (int k = 0; k < ncq; ++k) for (int c = 0; c < 6; ++c) m.push_back(O.task0 + TR_ROWS * k + tr + c); };
    auto span = [&](std::vector<int>& m, int r0, int n) { for (int c = 0; c < n; ++c) m.push_back(r0 + c); };
    for (int field = 0; field <= LOIK_F_Q; ++field) {
      std::vector<int> m;
      switch (field) {
        case LOIK_F_Z: per_dof(m, JR_Z, FR_Z); break;
        case LOIK_F_NU: per_dof(m, JR_NU, FR_NU); break;
        case LOIK_F_W: per_dof(m, JR_W, FR_W); break;
        case LOIK_F_Y: per_task(m, TR_Y); break;
        case LOIK_F_V: per_joint(m, JR_V, 6); break;
        case LOIK_F_F: per_joint(m, JR_F, 6); break;
        case LOIK_F_ATY: per_task(m, TR_ATY); break;
        case LOIK_F_FDPA: per_joint(m, JR_FD, 6); break;
        case LOIK_F_STF_PLUS_W: per_dof(m, JR_T, FR_T); break;
        case LOIK_F_P: per_joint(m, JR_P, 6); break;
        case LOIK_F_UDINV: per_joint_noff(m, JR_UD, 6); break;  // (zeros for multi-DoF joints: their 6 x K UDinv / K x K Dinv are not exposed per joint)
        case LOIK_F_DINV: per_joint_noff(m, JR_DINV, 1); break;
        case LOIK_F_R: per_dof(m, JR_R, FR_R); break;
        case LOIK_F_MU: span(m, O.glob + GR_MU, 1); break;
        case LOIK_F_RESIDUALS: span(m, O.glob + GR_RES, 4); break;
        case LOIK_F_NORMS: span(m, O.glob + GR_NORMS, LOIK_NUM_NORMS); break;
        case LOIK_F_PRIMAL_RES_VEC: span(m, O.prv, 6 * nb + h->nv); break;
        case LOIK_F_DUAL_RES_VEC: span(m, O.drv, 6 * nb + h->nv); break;
        case LOIK_F_Q:
          for (int j = 0; j < nb; ++j) {
            if (M.j[j + 1].nvj > 1) { for (int c = 0; c < nq_of(j + 1); ++c) m.push_back(O.ff0 + FR_ROWS * M.j[j + 1].mblk + FR_Q + c); }
            else if (M.j[j + 1].qkind == 1) { m.push_back(O.joint0 + JR_ROWS * j + JR_JQ + 1); m.push_back(O.joint0 + JR_ROWS * j + JR_JQ); }  // (cos, sin)
            else m.push_back(O.joint0 + JR_ROWS * j + JR_Q);
          }
          break;
        case LOIK_F_H:  // expand the 21 stored scalars of each joint to a full symmetric 6x6
          for (int j = 0; j < nb; ++j)
            for (int a = 0; a < 6; ++a)
              for (int c = 0; c < 6; ++c) {
                int r;
                if (a < 3 && c < 3) r = si(a, c);
                else if (a >= 3 && c >= 3) r = 15 + si(a - 3, c - 3);
                else if (a < 3) r = 6 + 3 * a + (c - 3);
                else r = 6 + 3 * c + (a - 3);
                m.push_back(O.joint0 + JR_ROWS * j + JR_H + r);
              }
          break;
        default: break;
      }
      h->map_off[field] = (int)all.size();
      h->map_len[field] = (int)m.size();
      all.insert(all.end(), m.begin(), m.end());
    }
    CKA(cudaMalloc(&h->d_map, std::max<size_t>(all.size(), 1) * sizeof(int)));
    CKA(cudaMemcpy(h->d_map, all.data(), all.size() * sizeof(int), cudaMemcpyHostToDevice));
  }
  {
    int lo = 0, hi = 0;
    CKA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CKA(cudaStreamCreateWithPriority(&h->hi_stream, cudaStreamNonBlocking, hi));
    CKA(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
    CKA(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
  }
  CKA(cudaMallocHost(&h->h_counts, 4 * sizeof(int)));
  CKA(cudaMallocHost(&h->h_stats, 4 * sizeof(unsigned long long)));
  h->S.arena = h->arena; h->S.n = batch; h->S.list = nullptr; h->S.n_list = nullptr; h->S.n_active = nullptr;
  CKA(cudaDeviceGetAttribute(&h->sms, cudaDevAttrMultiProcessorCount, device));
  CKA(cudaDeviceGetAttribute(&h->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
  CKA(cudaDeviceGetAttribute(&h->smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, device));
  {
    LaneGeom G;
    h->lane_ok = M.nmd == 0 && lane_geometry(h, G);
    if (h->lane_ok) {
      CKA(cudaFuncSetAttribute(k_iterate_lane<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem_optin));
      CKA(cudaFuncSetAttribute(k_iterate_lane<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem_optin));
    }
    // default switch point: short trees hand the instances still active after 32 sweeps to the lane-parallel kernel (one
    // Panda-65 536 solve 5.4 -> 2.8 ms, 57 -> 72 M solves/s with 10 solves in flight, -2 % with 32; a switch after 10 sweeps
    // gives 2.3 ms at 63 / 66 M); long / branching trees keep the tile kernels for throughput (Talos: the wide lane geometry
    // holds 5 instances per SM; lane_after = 8 gives one solve in 6.5 instead of 13.8 ms at 4.2 instead of 6.7 M solves/s
    // pipelined) unless the caller asks (loik_set_schedule)
    h->lane_after = (h->lane_ok && nb <= 12) ? 32 : -1;
  }
#undef CKA
  *out = h;
  if (params->logging && loik_set_logging(h, 1) != LOIK_OK) { loik_destroy(h); *out = nullptr; return LOIK_ERR_CUDA; }
  return LOIK_OK;
}

int loik_create(const loik_model_desc* model, const loik_params* params, int32_t batch, int32_t device, loik_solver** out) {
  return create_impl(model, params, batch, device, out, false);
}

int32_t loik_model_layout(const loik_model_desc* model, const loik_params* params, int32_t* out, int32_t cap) {
  if (!out || cap < 0) return fail(LOIK_ERR_INVALID, "loik_model_layout: null argument");
  loik_solver* h = nullptr;
  const int rc = create_impl(model, params, 32, 0, &h, true);
  if (rc) return rc;
  const ModelC& M = h->mc;
  std::vector<int32_t> v = {M.off.rows, M.npend, M.nseg, M.nwarp, M.nblevel, M.nflevel, M.nspan, M.nmd};
  for (int i = 1; i < M.nj; ++i) { v.push_back(M.j[i].carry); v.push_back(M.j[i].pout); v.push_back(M.j[i].npin); v.push_back(M.j[i].mblk); }
  for (int g = 0; g < M.nseg; ++g) for (short x : {M.seg[g].lo, M.seg[g].hi, M.seg[g].bwarp, M.seg[g].blevel, M.seg[g].fwarp, M.seg[g].flevel}) v.push_back(x);
  for (int g = 0; g < M.nspan; ++g) for (short x : {M.span[g].lo, M.span[g].hi, M.span[g].md}) v.push_back(x);
  delete h;  // a dry handle owns no CUDA resource
  if ((int)v.size() > cap) return fail(LOIK_ERR_INVALID, "loik_model_layout: output buffer too small");
  std::copy(v.begin(), v.end(), out);
  return (int32_t)v.size();
}

int32_t loik_wide_table(const loik_model_desc* model, const loik_params* params, int32_t* out, int32_t cap) {
  if (!out || cap < 0) return fail(LOIK_ERR_INVALID, "loik_wide_table: null argument");
  loik_solver* h = nullptr;
  const int rc = create_impl(model, params, 32, 0, &h, true);
  if (rc) return rc;
  std::vector<int4> tab;
  build_wide_table(h->mc, tab);
  std::vector<int32_t> v = {h->mc.nsb, h->mc.nsf};
  for (const int4& e : tab)
    for (int x : {e.x & 0xffff, e.x >> 16, e.y, e.z & 0xffff, e.z >> 16, (int)(short)(e.w & 0xffff), e.w >> 16}) v.push_back(x);
  delete h;
  if ((int)v.size() > cap) return fail(LOIK_ERR_INVALID, "loik_wide_table: output buffer too small");
  std::copy(v.begin(), v.end(), out);
  return (int32_t)v.size();
}

void loik_destroy(loik_solver* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  cudaFree(h->S.dbg); cudaFree(h->arena); cudaFree(h->scratch[0]); cudaFree(h->scratch[1]); cudaFree(h->d_origin); cudaFree(h->d_lists); cudaFree(h->d_counts); cudaFree(h->d_hist); cudaFree(h->d_wide_tab); cudaFree(h->d_stats); cudaFree(h->d_map);
  if (h->g_exec) cudaGraphExecDestroy(h->g_exec);
  if (h->hi_stream) cudaStreamDestroy(h->hi_stream);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  cudaFreeHost(h->h_counts); cudaFreeHost(h->h_stats);
  if (h->h_stage) cudaFreeHost(h->h_stage);
  if (h->d_stage) cudaFree(h->d_stage);
  delete h;
}

static int launch_reset(loik_solver* h, int flags, cudaStream_t st) {
  k_reset<<<grid_for(h->batch, 128), 128, 0, st>>>(h->mc, h->S, flags);
  h->launches++;
  h->last_list = -1;
  h->sweeps_in_solve = 0;
  if (flags & RST_SOLVER) h->ws_valid = true;  // every instance is active again: the next backward pass rewrites all of it
  CK(cudaGetLastError());
  return LOIK_OK;
}

// problem_.UpdateReference / UpdateIneqConstraints / UpdateEqConstraints: the batch-uniform part goes to the constant block.
// Everything is validated before the handle's block is touched (a failed call leaves the previous problem intact).
// A == nullptr: the task matrices are per instance (rows of the task blocks, k_set_b).
static int set_problem_consts(loik_solver* h, const double* H_ref, const double* v_ref, int n_ids, const int32_t* ids,
                              const double* A, const double* lb, const double* ub, bool bounds_shared) {
  if (n_ids != h->nc)
    return fail(LOIK_ERR_INVALID, "[IkProblemFormulation::UpdateEqConstraints]: number of equality constraints doesn't match initialization!!!");
  if (!is_symmetric(H_ref))
    return fail(LOIK_ERR_UNSUPPORTED, "loik_solve_init: H_ref must be symmetric (the optimized path's SE3actOn reads only the LL, LA, AA blocks)");
  for (int k = 0; k < n_ids; ++k) {
    if (ids[k] < 1 || ids[k] >= h->nj) return fail(LOIK_ERR_INVALID, "loik_solve_init: task joint id out of range [1, njoints-1]");
    for (int k2 = 0; k2 < k; ++k2)
      if (ids[k2] == ids[k])
        return fail(LOIK_ERR_UNSUPPORTED, "[IkProblemFormulation::UpdateEqConstraint]: multiple constraint specification for the same link id, not supported, terminating !!!");
  }
  ModelC& M = h->mc;
  double Hv[6];
  for (int i = 0; i < 6; ++i) { Hv[i] = 0; for (int j = 0; j < 6; ++j) Hv[i] += H_ref[6 * i + j] * v_ref[j]; }
  double hv_inf = 0; for (int i = 0; i < 6; ++i) hv_inf = std::max(hv_inf, std::fabs(Hv[i]));
  M.Hv_inf = hv_inf;  // = |Hv[0]|inf (ik-id-description-optimized.hpp:95)
  M.bounds_per_instance = bounds_shared ? 0 : 1;
  M.href_uniform = 1;
  M.vref_per = 0;
  M.href_per = 0;
  M.a_per = (A && n_ids <= kMaxTasks) ? 0 : 1;  // (more tasks than TaskC slots: the matrices live in the task rows too, k_set_b)
  sym_blocks(H_ref, M.href[0].A, M.href[0].B, M.href[0].D);  // UpdateReference: one reference broadcast to every joint
  for (int c = 0; c < 6; ++c) M.href[0].Hv[c] = Hv[c];
  for (int i = 1; i < h->nj; ++i) {
    JointC& J = M.j[i];
    J.href = 0;
    J.task = -1;
    if (bounds_shared) { J.lb = lb[J.idxv]; J.ub = ub[J.idxv]; }
  }
  std::memset(M.t, 0, sizeof(M.t));
  for (int k = 0; k < n_ids; ++k) {
    const int c = ids[k];
    M.j[c].task = (short)k;
    M.task_joint[k] = (short)c;
    if (M.a_per) continue;
    TaskC& T = M.t[k];
    double AtA[36];
    for (int i = 0; i < 36; ++i) T.A[i] = A[36 * k + i];
    for (int i = 0; i < 6; ++i)
      for (int j = 0; j < 6; ++j) { double s = 0; for (int r = 0; r < 6; ++r) s += T.A[6 * r + i] * T.A[6 * r + j]; AtA[6 * i + j] = s; }
    sym_blocks(AtA, T.AtA_A, T.AtA_B, T.AtA_D);
  }
  return LOIK_OK;
}

int loik_solve_init(loik_solver* h, const double* q, const double* H_ref, const double* v_ref, int32_t n_ids,
                    const int32_t* ids, const double* A, int32_t A_per_instance, const double* b, int32_t b_per_instance, const double* lb,
                    const double* ub, int32_t bounds_per_instance, int32_t loc, void* stream) {
  if (!h || !q || !H_ref || !v_ref || !lb || !ub || (n_ids > 0 && (!ids || !A || !b))) return fail(LOIK_ERR_INVALID, "loik_solve_init: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  CK(cudaSetDevice(h->device));
  const int B = h->batch, nc = h->nc;
  const size_t q_bytes = (size_t)B * h->nq * sizeof(double);
  const size_t b_bytes = (size_t)(b_per_instance ? B : 1) * nc * 6 * sizeof(double);
  const size_t bd_bytes = (size_t)(bounds_per_instance ? B : 1) * h->nv * sizeof(double);
  if (loc != LOIK_HOST && loc != LOIK_DEVICE && loc != LOIK_HOST_PINNED) return fail(LOIK_ERR_INVALID, "loik_solve_init: bad loc");
  // batch-shared bounds / task matrices are batch-uniform data like H_ref: always host pointers, they travel in the kernel parameter block
  int rc = set_problem_consts(h, H_ref, v_ref, n_ids, ids, A_per_instance ? nullptr : A, lb, ub, !bounds_per_instance);
  if (rc) return rc;
  h->a_user_per = A_per_instance != 0;
  const bool md_shared = !bounds_per_instance && h->mc.nmd > 0;  // shared bounds of multi-DoF joints: replicated into their rows (HOST pointers)
  const bool a_rows_shared = h->mc.a_per && !A_per_instance && nc > 0;  // more shared task matrices than TaskC slots: into the task rows (HOST pointer)
  const size_t A_bytes = A_per_instance ? (size_t)B * nc * 36 * sizeof(double) : (a_rows_shared ? (size_t)nc * 36 * sizeof(double) : 0);
  if (loc != LOIK_DEVICE || md_shared || a_rows_shared) {
    rc = ensure_stage(h, q_bytes + b_bytes + 2 * bd_bytes + A_bytes + 64, loc == LOIK_HOST || md_shared || a_rows_shared);
    if (rc) return rc;
  }
  const void *dq, *db = nullptr, *dlb = nullptr, *dub = nullptr, *dA = nullptr;
  size_t off = 0;
  rc = to_device(h, q, q_bytes, loc, off, st, &dq); if (rc) return rc; off += q_bytes;
  if (nc > 0) { rc = to_device(h, b, b_bytes, loc, off, st, &db); if (rc) return rc; off += b_bytes; }
  if (bounds_per_instance) {
    rc = to_device(h, lb, bd_bytes, loc, off, st, &dlb); if (rc) return rc; off += bd_bytes;
    rc = to_device(h, ub, bd_bytes, loc, off, st, &dub); if (rc) return rc; off += bd_bytes;
  } else if (md_shared) {
    off = q_bytes + b_bytes;  // (a fixed place in the staging buffer, whatever was staged before)
    rc = to_device(h, lb, bd_bytes, LOIK_HOST, off, st, &dlb); if (rc) return rc; off += bd_bytes;
    rc = to_device(h, ub, bd_bytes, LOIK_HOST, off, st, &dub); if (rc) return rc; off += bd_bytes;
  }
  if (A_per_instance && nc > 0) { rc = to_device(h, A, A_bytes, loc, off, st, &dA); if (rc) return rc; off += A_bytes; }
  else if (a_rows_shared) { off = q_bytes + b_bytes + 2 * bd_bytes; rc = to_device(h, A, A_bytes, LOIK_HOST, off, st, &dA); if (rc) return rc; }
  // ik_id_data_.Reset(warm_start) + ResetSolver() + FwdPassInit's y/Aty wipe (hpp:346-359, hxx:270-278)
  const int flags = RST_SOLVER | (h->prm.warm_start ? 0 : (RST_WZ | RST_NU | RST_VFF | RST_YATY));
  k_reset<<<grid_for(B, 128), 128, 0, st>>>(h->mc, h->S, flags);
  h->ws_valid = true;
  launch_set_q(h, st, (const double*)dq);
  h->launches += 2;
  h->last_list = -1;
  if (nc > 0) { k_set_b<<<grid_for(B, 128), 128, 0, st>>>(h->mc, h->S, (const double*)db, b_per_instance, -1, (const double*)dA, A_per_instance ? 1 : 0); h->launches++; }
  if (bounds_per_instance || md_shared) {
    k_set_bounds<<<grid_for(B, 128), 128, 0, st>>>(h->mc, h->S, (const double*)dlb, (const double*)dub, bounds_per_instance ? 1 : 0);
    h->launches++;
  }
  CK(cudaGetLastError());
  if (loc == LOIK_HOST || md_shared || a_rows_shared) CK(cudaStreamSynchronize(st));  // the staging buffer may be reused by the next call
  h->problem_set = true;
  return LOIK_OK;
}

int loik_update_references(loik_solver* h, const double* H_refs, const double* v_refs, void* stream) {
  (void)stream;
  if (!h || !H_refs || !v_refs) return fail(LOIK_ERR_INVALID, "loik_update_references: null argument");
  ModelC& M = h->mc;
  for (int i = 0; i < h->nj; ++i)
    if (!is_symmetric(H_refs + 36 * i)) return fail(LOIK_ERR_UNSUPPORTED, "loik_update_references: H_refs[i] must be symmetric");
  // distinct (H_ref, v_ref) pairs go to the reference table of the parameter block, joints point into it
  std::vector<int> entry(h->nj, 0), first;
  for (int i = 1; i < h->nj; ++i) {
    int e = -1;
    for (size_t k = 0; k < first.size() && e < 0; ++k)
      if (!std::memcmp(H_refs + 36 * i, H_refs + 36 * first[k], 36 * sizeof(double)) && !std::memcmp(v_refs + 6 * i, v_refs + 6 * first[k], 6 * sizeof(double))) e = (int)k;
    if (e < 0) { e = (int)first.size(); first.push_back(i); }
    entry[i] = e;
  }
  if ((int)first.size() > kMaxHref) return fail(LOIK_ERR_UNSUPPORTED, "loik_update_references: more distinct (H_ref, v_ref) pairs than the parameter block holds (kMaxHref)");
  for (int i = 0; i < h->nj; ++i) {
    double Hv[6], n = 0;
    for (int a = 0; a < 6; ++a) { Hv[a] = 0; for (int c = 0; c < 6; ++c) Hv[a] += H_refs[36 * i + 6 * a + c] * v_refs[6 * i + c]; n = std::max(n, std::fabs(Hv[a])); }
    if (n > M.Hv_inf) M.Hv_inf = n;  // only grows (ik-id-description-optimized.hpp:115-117)
    if (i >= 1) {
      HrefC& R = M.href[entry[i]];
      sym_blocks(H_refs + 36 * i, R.A, R.B, R.D);
      for (int a = 0; a < 6; ++a) R.Hv[a] = Hv[a];
      M.j[i].href = (short)entry[i];
    }
  }
  M.href_uniform = first.size() <= 1 ? 1 : 0;
  M.vref_per = 0; M.href_per = 0;  // (per-joint references shared by the batch replace per-instance ones; loik_update_references_batch sets them again)
  return LOIK_OK;
}

int loik_update_references_batch(loik_solver* h, const double* H_refs, int32_t H_per_instance, const double* v_refs, int32_t loc, void* stream) {
  if (!h || !H_refs || !v_refs) return fail(LOIK_ERR_INVALID, "loik_update_references_batch: null argument");
  if (!h->problem_set) return fail(LOIK_ERR_STATE, "loik_update_references_batch: call loik_solve_init first");
  int rc = LOIK_OK;
  if (!H_per_instance) {
    // the weights go through the shared path (reference table, symmetry check) with v_ref = 0: |Hv|inf does not grow there
    std::vector<double> zero(6 * (size_t)h->nj, 0.0);
    rc = loik_update_references(h, H_refs, zero.data(), stream);
    if (rc) return rc;
  } else if (loc != LOIK_DEVICE) {
    for (size_t i = 0; i < (size_t)h->batch * h->nj; ++i)
      if (!is_symmetric(H_refs + 36 * i)) return fail(LOIK_ERR_UNSUPPORTED, "loik_update_references_batch: every H_ref must be symmetric");
  }
  cudaStream_t st = (cudaStream_t)stream;
  CK(cudaSetDevice(h->device));
  const size_t v_bytes = (size_t)h->batch * h->nj * 6 * sizeof(double), H_bytes = H_per_instance ? 6 * v_bytes : 0;
  if (loc != LOIK_DEVICE) { rc = ensure_stage(h, v_bytes + H_bytes, loc == LOIK_HOST); if (rc) return rc; }
  const void *dv = nullptr, *dH = nullptr;
  rc = to_device(h, v_refs, v_bytes, loc, 0, st, &dv);
  if (rc) return rc;
  if (H_per_instance) { rc = to_device(h, H_refs, H_bytes, loc, v_bytes, st, &dH); if (rc) return rc; }
  h->mc.vref_per = 1;
  h->mc.href_per = H_per_instance ? 1 : 0;
  k_set_vref<<<grid_for(h->batch, 128), 128, 0, st>>>(h->mc, h->S, (const double*)dv, (const double*)dH);
  h->launches++;
  CK(cudaGetLastError());
  if (loc == LOIK_HOST) CK(cudaStreamSynchronize(st));  // the staging buffer is reused by the next call
  if (h->g_exec) { cudaGraphExecDestroy(h->g_exec); h->g_exec = nullptr; }  // (the schedule may change: no lane kernel with per-instance references)
  return LOIK_OK;
}

// Compact the still-active instances into the other list.  Everything stays on the stream.
static int compact(loik_solver* h, cudaStream_t st) {
  const int B = h->batch;
  const int in = h->last_list, outi = in < 0 ? 0 : 1 - in;
  int* list_out = h->d_lists + (size_t)outi * B;
  CK(cudaMemsetAsync(h->d_counts + outi, 0, sizeof(int), st));
  k_compact<<<grid_for(B, 256), 256, 0, st>>>(h->mc, h->S, in < 0 ? nullptr : h->d_lists + (size_t)in * B,
                                              in < 0 ? nullptr : h->d_counts + in, list_out, h->d_counts + outi);
  h->launches++;
  h->last_list = outi;
  return LOIK_OK;
}

// The main loop of Solve() (hpp:377-454) over the whole batch, enqueued as a fixed schedule of launches with NO
// host round trip.  Per-instance loop control lives on the device (finished instances are frozen), so the
// schedule only has to cover `budget` = max_iter sweeps.
//   1. a few dense sweeps on the home arena while (almost) every instance is active; the survivors claim the
//      slots of the next launch (claim_next);
//   2. then a sequence of MIGRATING launches of 1,1,2,2,4,4,... iterations: each reads its instances where the
//      previous launch left them and writes them, from its first iteration on, to the dense prefix of the other
//      scratch arena (full tiles, coalesced rows) -- the physical re-pack costs no pass of its own; an instance that
//      finishes copies its results to its home slot on the spot (retire_rows), survivors claim the next slots.
//   Launches past global convergence find a zero count and exit at once.
static int ensure_scratch(loik_solver* h) {
  if (h->scratch[0]) return LOIK_OK;
  const size_t bytes = (size_t)h->ntiles * h->mc.off.rows * 32 * sizeof(double);
  for (int i = 0; i < 2; ++i) CK(cudaMalloc(&h->scratch[i], bytes));
  CK(cudaMalloc(&h->d_origin, 2 * (size_t)h->batch * sizeof(int)));
  return LOIK_OK;
}

static int run_schedule(loik_solver* h, cudaStream_t st0, int budget) {
  const bool lane = use_lane(h);
  // sweeps done by the tile kernels before the lane-parallel kernel takes every instance that is still active
  const int pre = lane ? std::min(budget, h->lane_after) : budget;
  const int dense_sweeps = h->debug ? budget : h->dense_sweeps;  // debug mode: in place throughout (the debug arena is indexed by home slot)
  int rc = (pre > dense_sweeps) ? ensure_scratch(h) : LOIK_OK;
  if (rc) return rc;
  cudaStream_t st = st0;
  bool forked = false;
  const int B = h->batch;
  int done = 0;
  int li = 0;  // list / count that the NEXT launch reads
  const int dense = std::min(pre, dense_sweeps);
  // the tile launch right before the lane-parallel kernel orders its survivors: far-from-done first (StateP::next_back)
  const bool hand_over = lane && pre < budget && h->hard_ratio > 0.0;
  bool two_ended = false;
  StateP X = h->S;  // where the instances of the next launch live: the home arena first
  X.list = nullptr; X.n_list = nullptr;
  if (dense > 0) {
    CK(cudaMemsetAsync(h->d_counts + li, 0, sizeof(int), st));
    StateP P = h->S;
    P.next_list = h->d_lists + (size_t)li * B; P.next_count = h->d_counts + li;
    if (hand_over && dense == pre) { CK(cudaMemsetAsync(h->d_counts + 5, 0, sizeof(int), st)); P.next_back = h->d_counts + 5; P.next_cap = B; P.hard_ratio = h->hard_ratio; two_ended = true; }
    launch_iterate(h, st, P, dense, 0, 0, h->seg_after <= 0);
    h->sweeps += dense; done += dense;
    X.list = P.next_list; X.n_list = P.next_count;
  }
  int cur = -1;  // -1 = home arena, else scratch index
  int chunk = 1, reps = 0;
  while (done < pre) {
    if (!forked && h->hi_after >= 0 && done >= h->hi_after && h->hi_stream) {  // tail rounds: high-priority stream
      CK(cudaEventRecord(h->ev_fork, st0));
      CK(cudaStreamWaitEvent(h->hi_stream, h->ev_fork, 0));
      st = h->hi_stream;
      forked = true;
    }
    const int y = cur < 0 ? 0 : 1 - cur;
    const int c = std::min(chunk, pre - done);
    StateP P = X;
    P.dst = h->scratch[y];
    P.origin_src = cur < 0 ? nullptr : h->d_origin + (size_t)cur * B;
    P.origin_dst = h->d_origin + (size_t)y * B;
    P.home = h->S.arena;
    P.n_active = nullptr;
    int* nlist = nullptr; int* ncount = nullptr;
    if (done + c < budget) {
      CK(cudaMemsetAsync(h->d_counts + (1 - li), 0, sizeof(int), st));
      nlist = h->d_lists + (size_t)(1 - li) * B; ncount = h->d_counts + (1 - li);
    }
    P.next_list = nlist; P.next_count = ncount;
    if (hand_over && done + c == pre && nlist) { CK(cudaMemsetAsync(h->d_counts + 5, 0, sizeof(int), st)); P.next_back = h->d_counts + 5; P.next_cap = B; P.hard_ratio = h->hard_ratio; two_ended = true; }
    launch_iterate(h, st, P, c, 0, done >= h->small_after ? h->small_grid : 0, done >= h->seg_after);
    h->sweeps += c; done += c;
    X = h->S; X.arena = h->scratch[y]; X.list = nlist; X.n_list = ncount;
    cur = y; li = 1 - li;
    if (++reps == h->sched_reps) { reps = 0; if (chunk < 64) chunk = std::max(chunk + 1, (int)(chunk * h->sched_growth)); }
  }
  if (lane && done < budget) {  // everything still active runs to the end of its solve in the lane-parallel kernel
    rc = launch_lane(h, st, X.arena, X.list, X.n_list, cur < 0 ? nullptr : h->d_origin + (size_t)cur * B, 0, 0, two_ended ? h->d_counts + 5 : nullptr);
    if (rc) return rc;
    h->sweeps += budget - done;
  }
  CK(cudaMemsetAsync(h->d_counts + 2, 0, sizeof(int), st));  // nothing is active after a complete schedule
  if (forked) {
    CK(cudaEventRecord(h->ev_join, h->hi_stream));
    CK(cudaStreamWaitEvent(st0, h->ev_join, 0));
  }
  CK(cudaGetLastError());
  return LOIK_OK;
}

// (reset +) schedule, replayed from a CUDA graph when the stream can be captured (any stream but the legacy default
// one).  The graph is re-captured when the parameter block, the reset flags or the iteration budget change.
static int solve_scheduled_impl(loik_solver* h, cudaStream_t st, int reset_flags, int budget) {
  int rc = LOIK_OK;
  if ((use_lane(h) ? std::min(budget, h->lane_after) : budget) > h->dense_sweeps) { rc = ensure_scratch(h); if (rc) return rc; }  // (before any capture)
  const bool graphable = h->use_graph && st != nullptr && st != cudaStreamLegacy && !h->debug;
  if (!graphable) {
    if (reset_flags) { rc = launch_reset(h, reset_flags, st); if (rc) return rc; }
    h->last_list = -1; h->sweeps_in_solve = 0;
    return budget >= 1 ? run_schedule(h, st, budget) : LOIK_OK;
  }
  const bool valid = h->g_exec && h->g_flags == reset_flags && h->g_budget == budget && h->g_dense == h->dense_sweeps && h->g_keep == h->S.keep_ws &&
                     std::memcmp(&h->g_mc, &h->mc, sizeof(ModelC)) == 0;
  if (!valid) {
    if (h->g_exec) { cudaGraphExecDestroy(h->g_exec); h->g_exec = nullptr; }
    const int64_t l0 = h->launches, s0 = h->sweeps;
    cudaGraph_t graph = nullptr;
    CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed));
    rc = LOIK_OK;
    if (reset_flags) rc = launch_reset(h, reset_flags, st);
    if (!rc && budget >= 1) rc = run_schedule(h, st, budget);
    const cudaError_t ce = cudaStreamEndCapture(st, &graph);
    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (ce != cudaSuccess) return fail(LOIK_ERR_CUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(ce));
    const cudaError_t ie = cudaGraphInstantiate(&h->g_exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ie != cudaSuccess) { h->g_exec = nullptr; return fail(LOIK_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ie)); }
    h->g_mc = h->mc; h->g_flags = reset_flags; h->g_budget = budget; h->g_dense = h->dense_sweeps; h->g_keep = h->S.keep_ws;
    h->g_launches = h->launches - l0; h->g_sweeps = h->sweeps - s0;
    h->launches = l0; h->sweeps = s0;  // counted per replay below
  }
  CK(cudaGraphLaunch(h->g_exec, st));
  h->launches += h->g_launches; h->sweeps += h->g_sweeps;
  h->last_list = -1; h->sweeps_in_solve = 0;
  return LOIK_OK;
}

static int solve_scheduled(loik_solver* h, cudaStream_t st, int reset_flags, int budget) {
  const int rc = solve_scheduled_impl(h, st, reset_flags, budget);
  // the dense sweeps run in place; instances that finish in the migrating launches after them only bring their
  // workspace home with keep_ws
  if (!h->S.keep_ws && budget > 0 && (h->drop_ws || budget > h->dense_sweeps || (use_lane(h) && budget > h->lane_after))) h->ws_valid = false;
  return rc;
}

int loik_fwd_pass_init(loik_solver* h, const double* q, int32_t loc, void* stream) {
  if (!h || !q) return fail(LOIK_ERR_INVALID, "loik_fwd_pass_init: null argument");
  if (!h->problem_set) return fail(LOIK_ERR_STATE, "loik_fwd_pass_init: call loik_solve_init first");
  cudaStream_t st = (cudaStream_t)stream;
  CK(cudaSetDevice(h->device));
  const size_t q_bytes = (size_t)h->batch * h->nq * sizeof(double);
  int rc;
  if (loc != LOIK_DEVICE) { rc = ensure_stage(h, q_bytes + 64, loc == LOIK_HOST); if (rc) return rc; }
  const void* dq;
  rc = to_device(h, q, q_bytes, loc, 0, st, &dq); if (rc) return rc;
  launch_set_q(h, st, (const double*)dq);
  h->launches++;
  CK(cudaGetLastError());
  if (loc == LOIK_HOST) CK(cudaStreamSynchronize(st));
  return LOIK_OK;
}

int loik_reset_recursion(loik_solver* h, void* stream) {
  if (!h) return fail(LOIK_ERR_INVALID, "null handle");
  if (!h->problem_set) return fail(LOIK_ERR_STATE, "loik_reset_recursion: call loik_solve_init first");
  CK(cudaSetDevice(h->device));
  return launch_reset(h, RST_WZ | RST_VFF | RST_YATY | RST_SOLVER, (cudaStream_t)stream);
}

// ResetSolver() alone (hpp:168-186): iter_, the convergence / infeasibility flags, mu (-> mu_eq, mu_ineq) and the
// feasibility scalars; the primal and dual state is left as it is (a warm start keeps it).
int loik_reset_solver(loik_solver* h, void* stream) {
  if (!h) return fail(LOIK_ERR_INVALID, "null handle");
  if (!h->problem_set) return fail(LOIK_ERR_STATE, "loik_reset_solver: call loik_solve_init first");
  CK(cudaSetDevice(h->device));
  const bool ws = h->ws_valid;
  const int rc = launch_reset(h, RST_SOLVER, (cudaStream_t)stream);
  h->ws_valid = ws;  // (nothing of the workspace is touched)
  return rc;
}

static int check_strategy(loik_solver* h) {
  if (h->prm.mu_update_strat != LOIK_MU_DEFAULT)
    return fail(LOIK_ERR_UNSUPPORTED, "[FirstOrderLoikOptimizedTpl::UpdateMu]: mu update strategy not yet implemented");
  return LOIK_OK;
}

int loik_solve(loik_solver* h, void* stream) {
  if (!h) return fail(LOIK_ERR_INVALID, "null handle");
  if (!h->problem_set) return fail(LOIK_ERR_STATE, "loik_solve: call loik_solve_init first");
  int rc = check_strategy(h);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  CK(cudaSetDevice(h->device));
  // ResetRecursion + ResetSolver (hpp:370-374), then the main loop
  return solve_scheduled(h, st, RST_WZ | RST_VFF | RST_YATY | RST_SOLVER, h->prm.max_iter < 2 ? 0 : h->prm.max_iter);
}

int loik_solve_full(loik_solver* h, const double* q, const double* H_ref, const double* v_ref, int32_t n_ids,
                    const int32_t* ids, const double* A, int32_t A_per_instance, const double* b, int32_t b_per_instance, const double* lb,
                    const double* ub, int32_t bounds_per_instance, int32_t loc, void* stream) {
  if (h) { int rc = check_strategy(h); if (rc) return rc; }
  int rc = loik_solve_init(h, q, H_ref, v_ref, n_ids, ids, A, A_per_instance, b, b_per_instance, lb, ub, bounds_per_instance, loc, stream);
  if (rc) return rc;
  if (h->prm.max_iter < 2) return LOIK_OK;
  return solve_scheduled(h, (cudaStream_t)stream, 0, h->prm.max_iter);
}

int loik_solve_task(loik_solver* h, const double* q, int32_t c_id, const double* Ai, int32_t A_per_instance, const double* bi,
                    int32_t b_per_instance, int32_t loc, void* stream) {
  if (!h || !bi) return fail(LOIK_ERR_INVALID, "loik_solve_task: null argument");
  if (!h->problem_set) return fail(LOIK_ERR_STATE, "loik_solve_task: call loik_solve_init first");
  int rc = check_strategy(h);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  CK(cudaSetDevice(h->device));
  ModelC& M = h->mc;
  // problem_.UpdateEqConstraint(c_id, Ai, bi) (ik-id-description-optimized.hpp:178-218)
  int k = -1;
  for (int t = 0; t < h->nc; ++t) if (M.task_joint[t] == c_id) k = t;
  if (k < 0) return fail(LOIK_ERR_INVALID, "[IkProblemFormulation::UpdateEqConstraint]: constraint doesn't yet exist at link 'c_id' !!! ");
  // Ai == NULL: UpdateEqConstraint(c_id, bi) (:224-240): new target, the task keeps its matrix
  if (!Ai) A_per_instance = 0;
  if (Ai && (A_per_instance != 0) != h->a_user_per)
    return fail(LOIK_ERR_INVALID, "loik_solve_task: Ai must be per instance exactly when the task matrices of loik_solve_init were (all tasks of a handle "
                                  "keep their matrices in the same place)");
  const bool a_rows_shared = Ai && M.a_per && !A_per_instance;
  if (Ai && !M.a_per) {
    TaskC& T = M.t[k];
    double AtA[36];
    for (int i = 0; i < 36; ++i) T.A[i] = Ai[i];
    for (int i = 0; i < 6; ++i)
      for (int j = 0; j < 6; ++j) { double s = 0; for (int r = 0; r < 6; ++r) s += T.A[6 * r + i] * T.A[6 * r + j]; AtA[6 * i + j] = s; }
    sym_blocks(AtA, T.AtA_A, T.AtA_B, T.AtA_D);
  }
  const int B = h->batch;
  const size_t q_bytes = (size_t)B * h->nq * sizeof(double), b_bytes = (size_t)(b_per_instance ? B : 1) * 6 * sizeof(double);
  const size_t A_bytes = A_per_instance ? (size_t)B * 36 * sizeof(double) : (a_rows_shared ? 36 * sizeof(double) : 0);
  if (loc != LOIK_DEVICE || a_rows_shared) { rc = ensure_stage(h, q_bytes + b_bytes + A_bytes + 64, loc == LOIK_HOST || a_rows_shared); if (rc) return rc; }
  const void *dq = nullptr, *db, *dA = nullptr;
  if (q) { rc = to_device(h, q, q_bytes, loc, 0, st, &dq); if (rc) return rc; }
  rc = to_device(h, bi, b_bytes, loc, q_bytes, st, &db); if (rc) return rc;
  if (A_per_instance) { rc = to_device(h, Ai, A_bytes, loc, q_bytes + b_bytes, st, &dA); if (rc) return rc; }
  else if (a_rows_shared) { rc = to_device(h, Ai, A_bytes, LOIK_HOST, q_bytes + b_bytes, st, &dA); if (rc) return rc; }
  const int flags = RST_SOLVER | (h->prm.warm_start ? 0 : (RST_WZ | RST_NU | RST_VFF | RST_YATY));
  k_reset<<<grid_for(B, 128), 128, 0, st>>>(h->mc, h->S, flags);
  h->ws_valid = true;
  k_set_b<<<grid_for(B, 128), 128, 0, st>>>(h->mc, h->S, (const double*)db, b_per_instance, k, (const double*)dA, A_per_instance ? 1 : 0);
  if (q) launch_set_q(h, st, (const double*)dq);  // else: keep the device-resident q (loik_integrate)
  h->launches += 3;
  h->last_list = -1;
  CK(cudaGetLastError());
  if (loc == LOIK_HOST || a_rows_shared) CK(cudaStreamSynchronize(st));
  if (h->prm.max_iter < 2) return LOIK_OK;
  return solve_scheduled(h, st, 0, h->prm.max_iter);
}

int loik_integrate(loik_solver* h, double dt, void* stream) {
  if (!h) return fail(LOIK_ERR_INVALID, "null handle");
  if (!h->problem_set) return fail(LOIK_ERR_STATE, "loik_integrate: call loik_solve_init first");
  cudaStream_t st = (cudaStream_t)stream;
  CK(cudaSetDevice(h->device));
  k_integrate<<<grid_for(h->batch, 128), 128, 0, st>>>(h->mc, h->S, dt);
  h->launches++;
  CK(cudaGetLastError());
  return LOIK_OK;
}

int loik_iterate_fixed(loik_solver* h, int32_t iters, int32_t reset, void* stream) {
  if (!h) return fail(LOIK_ERR_INVALID, "null handle");
  if (!h->problem_set) return fail(LOIK_ERR_STATE, "loik_iterate_fixed: call loik_solve_init first");
  cudaStream_t st = (cudaStream_t)stream;
  CK(cudaSetDevice(h->device));
  if (reset) { int rc = launch_reset(h, RST_WZ | RST_VFF | RST_YATY | RST_SOLVER, st); if (rc) return rc; }
  if (use_lane(h) && h->lane_after == 0) {
    // the whole solve runs in the lane-parallel kernel: one launch, `iters` iterations of every instance
    if (iters > 0) { int rc = launch_lane(h, st, h->arena, nullptr, nullptr, nullptr, 1, iters); if (rc) return rc; }
    if (!h->S.keep_ws) h->ws_valid = false;
  } else {
    // one launch per iteration, dense: this is the quantity the roofline is quoted on
    for (int i = 0; i < iters; ++i) launch_iterate(h, st, h->S, 1, 1, 0, h->seg_after <= 0);  // the kernel the bulk of a solve runs
  }
  h->sweeps += iters;
  CK(cudaGetLastError());
  return LOIK_OK;
}

// ---- chunked solve for the batch-sharded multi-GPU driver: the caller interleaves chunks with the all-reduce of
// the active count (loik_active_count_device_ptr) and stops when the global count is zero.
int loik_solve_begin(loik_solver* h, void* stream) {
  if (!h) return fail(LOIK_ERR_INVALID, "null handle");
  if (!h->problem_set) return fail(LOIK_ERR_STATE, "loik_solve_begin: call loik_solve_init first");
  int rc = check_strategy(h);
  if (rc) return rc;
  CK(cudaSetDevice(h->device));
  return launch_reset(h, RST_WZ | RST_VFF | RST_YATY | RST_SOLVER, (cudaStream_t)stream);
}
int loik_solve_chunk(loik_solver* h, int32_t iters, void* stream) {
  if (!h) return fail(LOIK_ERR_INVALID, "null handle");
  if (!h->problem_set) return fail(LOIK_ERR_STATE, "loik_solve_chunk: call loik_solve_init and loik_solve_begin first");
  if (iters < 1) return fail(LOIK_ERR_INVALID, "loik_solve_chunk: iters must be >= 1");
  cudaStream_t st = (cudaStream_t)stream;
  CK(cudaSetDevice(h->device));
  int rc = LOIK_OK;
  const int B = h->batch;
  StateP S = h->S;
  if (h->last_list >= 0 || h->sweeps_in_solve >= 3) {  // dense for the first sweeps, compacted afterwards
    rc = compact(h, st);
    if (rc) return rc;
    S.list = h->d_lists + (size_t)h->last_list * B;
    S.n_list = h->d_counts + h->last_list;
  }
  CK(cudaMemsetAsync(h->d_counts + 2, 0, sizeof(int), st));
  S.n_active = h->d_counts + 2;
  launch_iterate(h, st, S, iters, 0);
  h->sweeps += iters; h->sweeps_in_solve += iters;
  CK(cudaGetLastError());
  return LOIK_OK;
}
int loik_active_count_device_ptr(loik_solver* h, void** dev_ptr) {
  if (!h || !dev_ptr) return fail(LOIK_ERR_INVALID, "null argument");
  *dev_ptr = h->d_counts + 2;
  return LOIK_OK;
}

int loik_step(loik_solver* h, int32_t step_id, void* stream) {
  if (!h) return fail(LOIK_ERR_INVALID, "null handle");
  if (!h->problem_set) return fail(LOIK_ERR_STATE, "loik_step: call loik_solve_init first");
  cudaStream_t st = (cudaStream_t)stream;
  CK(cudaSetDevice(h->device));
  const int g = grid_for(h->batch);
  switch (step_id) {
    case LOIK_STEP_BACKWARD: k_step_backward<<<g, kBlock, 0, st>>>(h->mc, h->S); break;
    case LOIK_STEP_FORWARD: k_step_forward<<<g, kBlock, 0, st>>>(h->mc, h->S); break;
    case LOIK_STEP_RESIDUAL: k_step_residual<<<g, kBlock, 0, st>>>(h->mc, h->S, 0); h->sweeps++; break;
    case LOIK_STEP_UPDATE_PREV: return LOIK_OK;  // the sweeps read the previous iterate before overwriting it
    case LOIK_STEP_RESET_INF_NORMS: case LOIK_STEP_FWD_PASS1: case LOIK_STEP_BWD_PASS: case LOIK_STEP_FWD_PASS2:
    case LOIK_STEP_BOX_PROJ: case LOIK_STEP_DUAL_UPDATE: case LOIK_STEP_COMPUTE_RESIDUALS: case LOIK_STEP_CHECK_CONVERGENCE:
    case LOIK_STEP_CHECK_FEASIBILITY: case LOIK_STEP_UPDATE_MU:
      if (!h->debug) return fail(LOIK_ERR_STATE, "loik_step: the per-method steps need loik_set_debug(h, 1)");
      if (h->mc.nmd > 0) return fail(LOIK_ERR_UNSUPPORTED, "loik_step: the per-method steps do not support multi-DoF joints (use the fused steps)");
      k_fine<<<g, kBlock, 0, st>>>(h->mc, h->S, step_id);
      break;
    default: return fail(LOIK_ERR_INVALID, "loik_step: unknown step id");
  }
  h->launches++;
  CK(cudaGetLastError());
  return LOIK_OK;
}

int loik_set_keep_workspace(loik_solver* h, int32_t on) {
  if (!h) return fail(LOIK_ERR_INVALID, "null handle");
  h->S.keep_ws = on != 0;
  return LOIK_OK;
}

int loik_set_debug(loik_solver* h, int32_t on) {
  if (!h) return fail(LOIK_ERR_INVALID, "null handle");
  if (on && !h->S.dbg) {  // the residual vectors of the debug kernels live in an arena of their own (production arenas do not carry them)
    CK(cudaSetDevice(h->device));
    const size_t bytes = (size_t)h->ntiles * h->mc.off.drows * 32 * sizeof(double);
    CK(cudaMalloc(&h->S.dbg, bytes));
    CK(cudaMemset(h->S.dbg, 0, bytes));
  }
  h->debug = on != 0;
  return LOIK_OK;
}

static_assert(kHistCols == LOIK_HISTORY_COLS, "include/loik_b200.h and loik_device.cuh must agree");
int loik_set_logging(loik_solver* h, int32_t on) {
  if (!h) return fail(LOIK_ERR_INVALID, "null handle");
  if (!on) {
    h->S.hist = nullptr;
    if (h->debug_by_logging) { h->debug = false; h->debug_by_logging = false; }
    return LOIK_OK;
  }
  if (!h->debug) h->debug_by_logging = true;
  int rc = loik_set_debug(h, 1);  // the log is written by the debug instantiation of the iteration kernel (in place, one launch)
  if (rc) return rc;
  CK(cudaSetDevice(h->device));
  const int cap = std::max(1, h->prm.max_iter);
  if (!h->d_hist || h->hist_cap < cap) {
    if (h->d_hist) cudaFree(h->d_hist);
    h->d_hist = nullptr;
    CK(cudaMalloc(&h->d_hist, (size_t)h->batch * cap * kHistCols * sizeof(double)));
    CK(cudaMemset(h->d_hist, 0, (size_t)h->batch * cap * kHistCols * sizeof(double)));
    h->hist_cap = cap;
  }
  h->S.hist = h->d_hist; h->S.hist_cap = h->hist_cap;
  return LOIK_OK;
}
int32_t loik_history_capacity(loik_solver* h) { return h ? h->hist_cap : 0; }
int loik_get_history(loik_solver* h, double* dst, int32_t loc, void* stream) {
  if (!h || !dst) return fail(LOIK_ERR_INVALID, "loik_get_history: null argument");
  if (!h->d_hist) return fail(LOIK_ERR_STATE, "loik_get_history: logging is off (loik_params.logging / loik_set_logging)");
  CK(cudaSetDevice(h->device));
  const size_t bytes = (size_t)h->batch * h->hist_cap * kHistCols * sizeof(double);
  CK(cudaMemcpyAsync(dst, h->d_hist, bytes, loc == LOIK_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  if (loc != LOIK_DEVICE) CK(cudaStreamSynchronize((cudaStream_t)stream));
  return LOIK_OK;
}

int loik_get(loik_solver* h, int32_t field, void* dst, int32_t loc, void* stream) {
  if (!h || !dst) return fail(LOIK_ERR_INVALID, "loik_get: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  CK(cudaSetDevice(h->device));
  int rc = LOIK_OK;
  const int B = h->batch, nb = h->nb, nc = h->nc;
  const Offs& O = h->mc.off;
  if (field < 0 || field > LOIK_F_Q) return fail(LOIK_ERR_INVALID, "loik_get: unknown field");
  const bool is_int = field == LOIK_F_ITER || field == LOIK_F_STATUS;
  const bool is_ws = field == LOIK_F_H || field == LOIK_F_P || field == LOIK_F_UDINV || field == LOIK_F_DINV || field == LOIK_F_R;
  if ((field == LOIK_F_PRIMAL_RES_VEC || field == LOIK_F_DUAL_RES_VEC) && !h->S.dbg)
    return fail(LOIK_ERR_STATE, "loik_get: the residual vectors are only kept in debug mode (loik_set_debug(h, 1) before stepping / solving)");
  if (is_ws && !h->ws_valid)
    return fail(LOIK_ERR_STATE, "loik_get: the backward-pass workspace (His, pis, UDinv, Dinv, r) of the last solve was not kept; "
                                "call loik_set_keep_workspace(h, 1) before solving");
  const int rows = is_int ? 1 : (field == LOIK_F_LIMI ? 12 * nb : h->map_len[field]);
  (void)nc;
  const size_t bytes = (size_t)B * rows * (is_int ? sizeof(int) : sizeof(double));
  void* ddst = dst;
  if (loc != LOIK_DEVICE) { rc = ensure_stage(h, bytes, loc == LOIK_HOST); if (rc) return rc; ddst = h->d_stage; }
  if (field == LOIK_F_LIMI) {
    k_gather_limi<<<grid_for(B, 128), 128, 0, st>>>(h->mc, h->S, (double*)ddst);
  } else if (is_int) {
    k_gather_ctl<<<grid_for(B, 128), 128, 0, st>>>(h->mc, h->S, field == LOIK_F_ITER ? 0 : 1, (int*)ddst);
  } else {
    const size_t total = (size_t)B * rows;
    const bool dbg_field = field == LOIK_F_PRIMAL_RES_VEC || field == LOIK_F_DUAL_RES_VEC;
    k_gather<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(h->mc, h->S, rows, h->d_map + h->map_off[field], (double*)ddst,
                                                              dbg_field ? h->S.dbg : h->arena, dbg_field ? O.drows : O.rows);
  }
  h->launches++;
  CK(cudaGetLastError());
  if (loc == LOIK_HOST) {
    CK(cudaMemcpyAsync(h->h_stage, ddst, bytes, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    std::memcpy(dst, h->h_stage, bytes);
  } else if (loc == LOIK_HOST_PINNED) {
    CK(cudaMemcpyAsync(dst, ddst, bytes, cudaMemcpyDeviceToHost, st));  // caller synchronizes the stream before reading
  }
  return LOIK_OK;
}

int loik_get_stats(loik_solver* h, int64_t out[5]) {
  if (!h || !out) return fail(LOIK_ERR_INVALID, "null argument");
  CK(cudaSetDevice(h->device));
  CK(cudaDeviceSynchronize());
  CK(cudaMemset(h->d_stats, 0, 4 * sizeof(unsigned long long)));
  k_stats<<<grid_for(h->ntiles * 32, 128), 128>>>(h->mc, h->S, h->d_stats);
  h->launches++;
  CK(cudaMemcpy(h->h_stats, h->d_stats, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  for (int i = 0; i < 4; ++i) out[i] = (int64_t)h->h_stats[i];
  out[4] = h->sweeps;
  return LOIK_OK;
}

int loik_reduce_stats(loik_solver* h, void* stream, void** dev_ptr) {
  if (!h || !dev_ptr) return fail(LOIK_ERR_INVALID, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  CK(cudaSetDevice(h->device));
  CK(cudaMemsetAsync(h->d_stats, 0, 4 * sizeof(unsigned long long), st));
  k_stats<<<grid_for(h->ntiles * 32, 128), 128, 0, st>>>(h->mc, h->S, h->d_stats);
  h->launches++;
  CK(cudaGetLastError());
  *dev_ptr = h->d_stats;
  return LOIK_OK;
}

int64_t loik_launch_count(loik_solver* h) { return h ? h->launches : 0; }

// setters / getters of the base class (task-solver-base.hpp:87-141, loik-loid-optimized.hpp:703).  The hyper-parameters
// travel to the kernels in the parameter block; a cached launch graph is re-captured when that block changes.
#define LOIK_SETTER(name, type, check, assign)                                                                      \
  int name(loik_solver* h, type v) {                                                                                \
    if (!h) return fail(LOIK_ERR_INVALID, #name ": null handle");                                                   \
    if (!(check)) return fail(LOIK_ERR_INVALID, #name ": value out of range");                                      \
    assign;                                                                                                         \
    return LOIK_OK;                                                                                                 \
  }
LOIK_SETTER(loik_set_max_iter, int32_t, v >= 0, (h->prm.max_iter = v, h->mc.max_iter = v))
LOIK_SETTER(loik_set_rho, double, v == v, (h->prm.rho = v, h->mc.rho = v))
LOIK_SETTER(loik_set_mu, double, v == v, (h->prm.mu = v, h->mc.mu0 = v))
LOIK_SETTER(loik_set_mu_equality_scale_factor, double, v == v, (h->prm.mu_equality_scale_factor = v, h->mc.mu_scale = v))
LOIK_SETTER(loik_set_tol_abs, double, v == v, (h->prm.tol_abs = v, h->mc.tol_abs = v))
LOIK_SETTER(loik_set_tol_rel, double, v == v, (h->prm.tol_rel = v, h->mc.tol_rel = v))
LOIK_SETTER(loik_set_tol_primal_inf, double, v == v, (h->prm.tol_primal_inf = v, h->mc.tol_pinf = v))
LOIK_SETTER(loik_set_tol_dual_inf, double, v == v, (h->prm.tol_dual_inf = v, h->mc.tol_dinf = v))
LOIK_SETTER(loik_set_tol_tail_solve, double, v == v, (h->prm.tol_tail_solve = v, h->mc.tol_tail = v))
LOIK_SETTER(loik_set_warm_start, int32_t, true, (h->prm.warm_start = v))
#undef LOIK_SETTER

int loik_get_params(loik_solver* h, loik_params* out) {
  if (!h || !out) return fail(LOIK_ERR_INVALID, "loik_get_params: null argument");
  *out = h->prm;
  return LOIK_OK;
}

int loik_get_schedule(loik_solver* h, loik_schedule* out) {
  if (!h || !out) return fail(LOIK_ERR_INVALID, "loik_get_schedule: null argument");
  out->dense_sweeps = h->dense_sweeps; out->repack_reps = h->sched_reps; out->repack_growth = h->sched_growth;
  out->hi_priority_after = h->hi_after; out->seg_after = h->seg_after; out->seg_warps = h->seg_warps;
  out->lane_after = h->lane_after; out->use_graph = h->use_graph ? 1 : 0;
  out->small_after = h->small_after; out->small_grid = h->small_grid; out->drop_workspace = h->drop_ws ? 1 : 0;
  out->lane_hard_first_ratio = h->hard_ratio;
  LaneGeom G;
  const bool ok = h->lane_ok && lane_geometry(h, G);
  out->lane_warps_per_cta = h->lane_warps_req; out->lane_groups_per_instance = h->lane_gpi_req; out->lane_groups_chosen = G.gpi;
  out->lane_available = ok ? 1 : 0; out->lane_warps_chosen = G.warps; out->lane_ctas = G.grid; out->lane_smem_bytes = (int32_t)G.smem;
  return LOIK_OK;
}

int loik_set_schedule(loik_solver* h, const loik_schedule* sc) {
  if (!h || !sc) return fail(LOIK_ERR_INVALID, "loik_set_schedule: null argument");
  if (sc->dense_sweeps < 0 || sc->repack_reps < 1 || !(sc->repack_growth >= 1.0) || sc->seg_warps < 0 || sc->seg_warps > 4 || sc->small_grid < 1 ||
      sc->lane_warps_per_cta < 0 || sc->lane_warps_per_cta > 8 || !(sc->lane_hard_first_ratio >= 0.0) || (sc->lane_groups_per_instance != 0 && sc->lane_groups_per_instance != 1 && sc->lane_groups_per_instance != 4))
    return fail(LOIK_ERR_INVALID, "loik_set_schedule: dense_sweeps >= 0, repack_reps >= 1, repack_growth >= 1, 0 <= seg_warps <= 4, small_grid >= 1, 0 <= lane_warps_per_cta <= 8, lane_groups_per_instance in {0, 1, 4}");
  h->dense_sweeps = sc->dense_sweeps; h->sched_reps = sc->repack_reps; h->sched_growth = sc->repack_growth;
  h->hi_after = sc->hi_priority_after; h->seg_after = sc->seg_after;
  h->lane_after = sc->lane_after; h->use_graph = sc->use_graph != 0; h->lane_warps_req = sc->lane_warps_per_cta; h->lane_gpi_req = sc->lane_groups_per_instance;
  h->small_after = sc->small_after; h->small_grid = sc->small_grid; h->drop_ws = sc->drop_workspace != 0;
  h->hard_ratio = sc->lane_hard_first_ratio;
  if (sc->seg_warps != h->seg_warps) {
    h->seg_warps = sc->seg_warps; assign_segments(h->mc, h->seg_warps);
    int rc = upload_wide_table(h);
    if (rc) return rc;
  }
  if (h->g_exec) { cudaGraphExecDestroy(h->g_exec); h->g_exec = nullptr; }  // the cached launch graph follows the schedule
  return LOIK_OK;
}

}  // extern "C"

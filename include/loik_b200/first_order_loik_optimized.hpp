// first_order_loik_optimized.hpp -- header-only C++ facade over the C ABI of libloik_b200.so with the shape of
// loik::FirstOrderLoikOptimizedTpl<double> (/root/reference/include/loik/loik-loid-optimized.hpp:22): same
// constructor argument order (:129-134), same SolveInit / Solve() / Solve(q,...) / Solve(q, c_id, Ai, bi) entry points
// (:335-338, :368, :475-478, :596-597), same getters (task-solver-base.hpp:87-141), same failure mode (C ABI error codes
// are re-thrown as std::runtime_error with the reference's messages).
//
// Differences, all forced by batching: vectors carry a leading batch dimension (row-major, one row per instance),
// the caller-owned IkIdData of the reference lives in HBM inside the handle (read it back with z(), nu(), w() ...),
// and the model is the flat table `loik_b200::Model` below.  The single-instance facade with the reference's exact
// argument types (const pinocchio::Model&, caller-owned IkIdData&, Eigen / aligned-vector arguments, results written
// back into the IkIdData members) is loik_b200/loik_pinocchio.hpp.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../loik_b200.h"

namespace loik_b200 {

#ifndef LOIK_B200_ADMM_STRAT_DEFINED
#define LOIK_B200_ADMM_STRAT_DEFINED
enum ADMMPenaltyUpdateStrat { DEFAULT = LOIK_MU_DEFAULT, OSQP = LOIK_MU_OSQP, MAXEIGENVALUE = LOIK_MU_MAXEIGENVALUE };
#endif

// What the hot path reads from pinocchio::Model (loik-loid-optimized.hxx:46-47,258-265).
struct Model {
  int njoints = 0;  // incl. universe
  int nv = 0;
  std::vector<int32_t> parents, joint_types;
  std::vector<double> joint_axes;   // [njoints][3]
  std::vector<double> placement_R;  // [njoints][9] row-major
  std::vector<double> placement_p;  // [njoints][3]
};

class FirstOrderLoikOptimized {
 public:
  // loik-loid-optimized.hpp:129-134 (+ batch, device; ik_id_data lives in the handle)
  FirstOrderLoikOptimized(int max_iter, double tol_abs, double tol_rel, double tol_primal_inf, double tol_dual_inf, double rho,
                          double mu, double mu_equality_scale_factor, ADMMPenaltyUpdateStrat mu_update_strat, int num_eq_c,
                          int eq_c_dim, const Model& model, int batch, bool warm_start, double tol_tail_solve, bool verbose,
                          bool logging, int device = 0, void* stream = nullptr)
      : model_(model), batch_(batch), nc_(num_eq_c), stream_(stream) {
    loik_model_desc md{model_.njoints, model_.parents.data(), model_.joint_types.data(), model_.joint_axes.data(),
                       model_.placement_R.data(), model_.placement_p.data()};
    loik_params p{max_iter, tol_abs, tol_rel, tol_primal_inf, tol_dual_inf, rho, mu, mu_equality_scale_factor,
                  static_cast<int32_t>(mu_update_strat), num_eq_c, eq_c_dim, warm_start ? 1 : 0, tol_tail_solve,
                  verbose ? 1 : 0, logging ? 1 : 0};
    check(loik_create(&md, &p, batch, device, &h_));
  }
  ~FirstOrderLoikOptimized() { loik_destroy(h_); }
  FirstOrderLoikOptimized(const FirstOrderLoikOptimized&) = delete;
  FirstOrderLoikOptimized& operator=(const FirstOrderLoikOptimized&) = delete;

  // SolveInit(q, H_ref, v_ref, ids, Ais, bis, lb, ub)  (hpp:335-338).  q [batch][nq]; Ais [nc][36] or [batch][nc][36];
  // bis [batch][nc][6] or [nc][6]; lb/ub [nv].
  void SolveInit(const std::vector<double>& q, const std::vector<double>& H_ref, const std::vector<double>& v_ref,
                 const std::vector<int32_t>& active_task_constraint_ids, const std::vector<double>& Ais,
                 const std::vector<double>& bis, const std::vector<double>& lb, const std::vector<double>& ub) {
    validate(active_task_constraint_ids, Ais, bis, lb, ub);
    check(loik_solve_init(h_, q.data(), H_ref.data(), v_ref.data(), (int32_t)active_task_constraint_ids.size(),
                          active_task_constraint_ids.data(), Ais.data(), a_per_instance(Ais, active_task_constraint_ids.size()), bis.data(), per_instance(bis), lb.data(), ub.data(), 0,
                          LOIK_HOST, stream_));
  }
  void Solve() { check(loik_solve(h_, stream_)); }  // hpp:368
  void Solve(const std::vector<double>& q, const std::vector<double>& H_ref, const std::vector<double>& v_ref,
             const std::vector<int32_t>& active_task_constraint_ids, const std::vector<double>& Ais,
             const std::vector<double>& bis, const std::vector<double>& lb, const std::vector<double>& ub) {  // hpp:475-478
    validate(active_task_constraint_ids, Ais, bis, lb, ub);
    check(loik_solve_full(h_, q.data(), H_ref.data(), v_ref.data(), (int32_t)active_task_constraint_ids.size(),
                          active_task_constraint_ids.data(), Ais.data(), a_per_instance(Ais, active_task_constraint_ids.size()), bis.data(), per_instance(bis), lb.data(), ub.data(), 0,
                          LOIK_HOST, stream_));
  }
  // Solve(q, c_id, Ai, bi)  (hpp:596-597): bi [batch][6] or [6]
  void Solve(const std::vector<double>& q, int c_id, const std::vector<double>& Ai, const std::vector<double>& bi) {
    check(loik_solve_task(h_, q.data(), c_id, Ai.data(), Ai.size() == (size_t)batch_ * 36 && batch_ > 1 ? 1 : 0, bi.data(),
                          bi.size() == (size_t)batch_ * 6 && batch_ > 1 ? 1 : 0, LOIK_HOST, stream_));
  }

  // results: ik_id_data.z / nu / w / yis / vis / fis of the reference, batch-major
  std::vector<double> z() const { return get(LOIK_F_Z, model_.nv); }
  std::vector<double> nu() const { return get(LOIK_F_NU, model_.nv); }
  std::vector<double> w() const { return get(LOIK_F_W, model_.nv); }
  std::vector<double> yis() const { return get(LOIK_F_Y, 6 * nc_); }
  std::vector<double> vis() const { return get(LOIK_F_V, 6 * (model_.njoints - 1)); }
  std::vector<double> fis() const { return get(LOIK_F_F, 6 * (model_.njoints - 1)); }
  // getters of IkIdSolverBaseTpl (task-solver-base.hpp:87-141), one value per instance
  std::vector<int32_t> get_iter() const { return geti(LOIK_F_ITER); }
  std::vector<double> get_mu() const { return get(LOIK_F_MU, 1); }
  std::vector<double> get_primal_residual() const { return column(0); }
  std::vector<double> get_dual_residual() const { return column(1); }
  std::vector<double> get_tol_primal() const { return column(2); }
  std::vector<double> get_tol_dual() const { return column(3); }
  std::vector<bool> get_convergence_status() const { return flag(1); }
  std::vector<bool> get_primal_infeasibility_status() const { return flag(2); }
  std::vector<bool> get_dual_infeasibility_status() const { return std::vector<bool>(batch_, false); }  // never evaluated by the optimized path
  // the rest of the public surface the reference's tests drive (loik-loid-optimized.hpp:168-264, 698-755;
  // tests/loik-loid.cpp:340-478): the per-step methods one by one and the feasibility scalars.  They need
  // set_debug(true) (the production path fuses the steps and keeps no running norms) and work on every instance.
  void set_debug(bool on) { check(loik_set_debug(h_, on ? 1 : 0)); }
  // logging_ / LoikSolverInfo (hpp:47-127): per-iteration log of the following solves, [batch][capacity][LOIK_HISTORY_COLS]
  void set_logging(bool on) { check(loik_set_logging(h_, on ? 1 : 0)); }
  int history_capacity() const { return loik_history_capacity(h_); }
  std::vector<double> history() const {
    std::vector<double> out((size_t)batch_ * history_capacity() * LOIK_HISTORY_COLS);
    check(loik_get_history(h_, out.data(), LOIK_HOST, stream_));
    return out;
  }
  // His() / pis() after Solve(), as the reference leaves them in ik_id_data (tests/loik-loid.cpp:597-615): opt-in, a
  // finished instance then brings its backward-pass workspace home too (off: those getters throw after a solve)
  void set_keep_workspace(bool on) { check(loik_set_keep_workspace(h_, on ? 1 : 0)); }
  // ResetSolver() (hpp:168-186): iteration counter, flags, mu, feasibility scalars -- the primal / dual state is kept
  void ResetSolver() { check(loik_reset_solver(h_, stream_)); }
  // ik_id_data_.ResetRecursion() + ResetSolver(): what Solve() does before its loop (hpp:370-374)
  void ResetRecursion() { check(loik_reset_recursion(h_, stream_)); }
  void FwdPassInit(const std::vector<double>& q) { check(loik_fwd_pass_init(h_, q.data(), LOIK_HOST, stream_)); }
  void UpdatePrev() { step(LOIK_STEP_UPDATE_PREV); }
  void ResetInfNorms() { step(LOIK_STEP_RESET_INF_NORMS); }
  void FwdPass1() { step(LOIK_STEP_FWD_PASS1); }
  void BwdPassOptimizedVisitor() { step(LOIK_STEP_BWD_PASS); }
  void FwdPass2OptimizedVisitor() { step(LOIK_STEP_FWD_PASS2); }
  void BoxProj() { step(LOIK_STEP_BOX_PROJ); }
  void DualUpdate() { step(LOIK_STEP_DUAL_UPDATE); }
  void ComputeResiduals() { step(LOIK_STEP_COMPUTE_RESIDUALS); }
  void CheckConvergence() { step(LOIK_STEP_CHECK_CONVERGENCE); }
  void CheckFeasibility() { step(LOIK_STEP_CHECK_FEASIBILITY); }
  void UpdateMu() { step(LOIK_STEP_UPDATE_MU); }
  // problem_.UpdateReferences(H_refs, v_refs)  (ik-id-description-optimized.hpp:103-121): [njoints][36], [njoints][6];
  // v_refs of size [batch][njoints][6] gives every instance its own reference velocities, H_refs of size [batch][njoints][36] its own weights
  void UpdateReferences(const std::vector<double>& H_refs, const std::vector<double>& v_refs) {
    const size_t nj = (size_t)model_.njoints;
    if (batch_ > 1 && v_refs.size() == (size_t)batch_ * nj * 6 && (H_refs.size() == nj * 36 || H_refs.size() == (size_t)batch_ * nj * 36)) {
      check(loik_update_references_batch(h_, H_refs.data(), H_refs.size() != nj * 36, v_refs.data(), LOIK_HOST, stream_));
      return;
    }
    if (H_refs.size() != nj * 36 || v_refs.size() != nj * 6)
      throw std::runtime_error("[IkProblemFormulation::UpdateReferences]: input arguments 'H_refs', 'v_refs' have wrong size!!");
    check(loik_update_references(h_, H_refs.data(), v_refs.data(), stream_));
  }
  // outer IK loop on the device: q <- integrate(q, dt z), then Solve(c_id, Ai, bi) on the device-resident q
  void Integrate(double dt) { check(loik_integrate(h_, dt, stream_)); }
  void Solve(int c_id, const std::vector<double>& Ai, const std::vector<double>& bi) {
    check(loik_solve_task(h_, nullptr, c_id, Ai.data(), Ai.size() == (size_t)batch_ * 36 && batch_ > 1 ? 1 : 0, bi.data(),
                          bi.size() == (size_t)batch_ * 6 && batch_ > 1 ? 1 : 0, LOIK_HOST, stream_));
  }
  std::vector<double> His() const { return get(LOIK_F_H, 36 * (model_.njoints - 1)); }   // after the per-step methods, or after Solve() with set_keep_workspace(true)
  std::vector<double> pis() const { return get(LOIK_F_P, 6 * (model_.njoints - 1)); }    // same
  std::vector<double> Aty() const { return get(LOIK_F_ATY, 6 * nc_); }
  std::vector<double> get_primal_residual_vec() const { return get(LOIK_F_PRIMAL_RES_VEC, 6 * (model_.njoints - 1) + model_.nv); }
  std::vector<double> get_dual_residual_vec() const { return get(LOIK_F_DUAL_RES_VEC, 6 * (model_.njoints - 1) + model_.nv); }
  std::vector<double> get_dual_residual_v() const { return norm(LOIK_N_DUAL_RES_V); }
  std::vector<double> get_dual_residual_nu() const { return norm(LOIK_N_DUAL_RES_NU); }
  std::vector<double> get_delta_x_qp_inf_norm() const { return norm(LOIK_N_DELTA_X_QP_INF); }
  std::vector<double> get_delta_z_qp_inf_norm() const { return norm(LOIK_N_DELTA_Z_INF); }
  std::vector<double> get_delta_y_qp_inf_norm() const { return norm(LOIK_N_DELTA_Y_QP_INF); }
  std::vector<double> get_A_qp_T_delta_y_qp_inf_norm() const { return norm(LOIK_N_AT_DELTA_Y_QP_INF); }
  std::vector<double> get_ub_qp_T_delta_y_qp_plus() const { return norm(LOIK_N_UB_T_DELTA_Y_PLUS); }
  std::vector<double> get_lb_qp_T_delta_y_qp_minus() const { return norm(LOIK_N_LB_T_DELTA_Y_MINUS); }
  std::vector<bool> get_primal_infeasibility_cond_1() const { return normflag(LOIK_N_PINF_COND_1); }
  std::vector<bool> get_primal_infeasibility_cond_2() const { return normflag(LOIK_N_PINF_COND_2); }
  void set_max_iter(int m) { check(loik_set_max_iter(h_, m)); }
  void set_rho(double r) { check(loik_set_rho(h_, r)); }
  void set_mu(double m) { check(loik_set_mu(h_, m)); }
  void set_tol_tail_solve(double t) { check(loik_set_tol_tail_solve(h_, t)); }
  void set_tol_abs(double t) { check(loik_set_tol_abs(h_, t)); }
  void set_tol_rel(double t) { check(loik_set_tol_rel(h_, t)); }
  void set_tol_primal_inf(double t) { check(loik_set_tol_primal_inf(h_, t)); }
  void set_tol_dual_inf(double t) { check(loik_set_tol_dual_inf(h_, t)); }
  void set_mu_equality_scale_factor(double f) { check(loik_set_mu_equality_scale_factor(h_, f)); }
  loik_params get_params() const { loik_params p; check(loik_get_params(h_, &p)); return p; }  // get_max_iter() ... get_tol_dual_inf()
  double get_rho() const { return get_params().rho; }
  loik_solver* handle() const { return h_; }

 private:
  static void check(int rc) {
    if (rc != LOIK_OK) throw std::runtime_error(loik_last_error());
  }
  int per_instance(const std::vector<double>& bis) const { return bis.size() == (size_t)batch_ * nc_ * 6 && batch_ > 1 ? 1 : 0; }
  // Ais [nc][36] shared by the batch, or [batch][nc][36]: every instance its own task matrices
  int a_per_instance(const std::vector<double>& Ais, size_t n_ids) const { return Ais.size() == (size_t)batch_ * n_ids * 36 && batch_ > 1 ? 1 : 0; }
  void validate(const std::vector<int32_t>& ids, const std::vector<double>& Ais, const std::vector<double>& bis,
                const std::vector<double>& lb, const std::vector<double>& ub) const {
    // ik-id-description-optimized.hpp:132-134, :328-335
    if ((Ais.size() != ids.size() * 36 && Ais.size() != (size_t)batch_ * ids.size() * 36) || (bis.size() != ids.size() * 6 && bis.size() != (size_t)batch_ * ids.size() * 6))
      throw std::runtime_error("[IkProblemFormulation::UpdateEqConstraints]: task_constraint_ids, Ais, and bis have different size !!!");
    if (lb.size() != ub.size())
      throw std::runtime_error("[IkProblemFormulation::UpdateIneqConstraints]: lower bound and upper bound have different dimensions!!!");
    if ((int)lb.size() != model_.nv)
      throw std::runtime_error("IkProblemFormulation::UpdateIneqConstraints]: inequality constraint dimension has changed, this is not supported currently!!!");
  }
  std::vector<double> get(int field, int width) const {
    std::vector<double> out((size_t)batch_ * width);
    check(loik_get(h_, field, out.data(), LOIK_HOST, stream_));
    return out;
  }
  std::vector<int32_t> geti(int field) const {
    std::vector<int32_t> out(batch_);
    check(loik_get(h_, field, out.data(), LOIK_HOST, stream_));
    return out;
  }
  std::vector<double> column(int c) const {
    auto r = get(LOIK_F_RESIDUALS, 4);
    std::vector<double> out(batch_);
    for (int i = 0; i < batch_; ++i) out[i] = r[4 * i + c];
    return out;
  }
  void step(int which) { check(loik_step(h_, which, stream_)); }
  std::vector<double> norm(int idx) const {
    auto r = get(LOIK_F_NORMS, LOIK_NUM_NORMS);
    std::vector<double> out(batch_);
    for (int i = 0; i < batch_; ++i) out[i] = r[(size_t)LOIK_NUM_NORMS * i + idx];
    return out;
  }
  std::vector<bool> normflag(int idx) const {
    auto r = norm(idx);
    std::vector<bool> out(batch_);
    for (int i = 0; i < batch_; ++i) out[i] = r[i] != 0.0;
    return out;
  }
  std::vector<bool> flag(int bit) const {
    auto s = geti(LOIK_F_STATUS);
    std::vector<bool> out(batch_);
    for (int i = 0; i < batch_; ++i) out[i] = (s[i] & bit) != 0;
    return out;
  }
  Model model_;
  int batch_, nc_;
  void* stream_;
  loik_solver* h_ = nullptr;
};

}  // namespace loik_b200

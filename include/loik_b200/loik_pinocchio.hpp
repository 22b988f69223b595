// loik_pinocchio.hpp -- the reference's class shape over the C ABI of libloik_b200.so.
//
// loik_b200::FirstOrderLoikOptimizedTpl<Scalar> has the constructor, the SolveInit / Solve() / Solve(q, ...8) /
// Solve(q, c_id, Ai, bi) entry points and the getters of loik::FirstOrderLoikOptimizedTpl<Scalar>
// (/root/reference/include/loik/loik-loid-optimized.hpp:129-134, 335-338, 368, 475-478, 596-597; getters
// task-solver-base.hpp:87-141), takes the same types -- `const pinocchio::Model&`, a caller-owned
// `loik::IkIdDataTypeOptimizedTpl&` borrowed by reference (hpp:763), Eigen vectors / matrices,
// PINOCCHIO_ALIGNED_STD_VECTOR arguments -- and, like the reference, leaves its results IN the caller's IkIdData:
// z (the answer, hpp:333), nu, w, yis[k], vis[i], fis[i], His[i], pis[i] (what tests/loik-loid.cpp:597-615 reads) plus
// Aty, fis_diff_plus_Aty, Stf_plus_w, r and liMi.  One solver object = one problem instance (batch 1 of the batched
// library); a caller that wants the batched throughput uses loik_b200/first_order_loik_optimized.hpp or the C ABI.
//
// Build: -DLOIK_B200_WITH_PINOCCHIO with Pinocchio 3 / Eigen 3.4 / the reference's loik-loid-data-optimized.hpp on the
// include path.  Pinocchio is not available in the offline build container, so CI compiles and RUNS this file against
// tests/cpp/stub/ (a stand-in with exactly the members used here) and compares the written-back fields with the
// oracle (tests/cpp/pinocchio_adapter_test.cpp, tests/test_cpp_facade.py).  The only line that differs between the two
// builds is how the axis of an *Unaligned joint is read (boost::get on the joint variant vs a plain member).
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include <pinocchio/multibody/model.hpp>
#include <loik/loik-loid-data-optimized.hpp>

#include "../loik_b200.h"

namespace loik_b200 {

#ifndef LOIK_B200_ADMM_STRAT_DEFINED
#define LOIK_B200_ADMM_STRAT_DEFINED
enum ADMMPenaltyUpdateStrat { DEFAULT = LOIK_MU_DEFAULT, OSQP = LOIK_MU_OSQP, MAXEIGENVALUE = LOIK_MU_MAXEIGENVALUE };
#endif

namespace detail {
// the flat description the C ABI takes (loik_model_desc), filled from a pinocchio::Model
struct FlatModel {
  int njoints = 0, nv = 0, nq = 0;
  std::vector<int32_t> parents, joint_types;
  std::vector<double> joint_axes, placement_R, placement_p;
};

template <typename JointModel>
inline void joint_axis(const JointModel& jm, const std::string& shortname, double (&ax)[3]) {
#ifdef LOIK_B200_PINOCCHIO_STUB
  (void)shortname;
  ax[0] = jm.axis[0]; ax[1] = jm.axis[1]; ax[2] = jm.axis[2];
#else
  if (shortname == "JointModelRevoluteUnaligned") {
    const auto& a = boost::get<pinocchio::JointModelRevoluteUnaligned>(jm.toVariant()).axis;
    ax[0] = a[0]; ax[1] = a[1]; ax[2] = a[2];
  } else if (shortname == "JointModelPrismaticUnaligned") {
    const auto& a = boost::get<pinocchio::JointModelPrismaticUnaligned>(jm.toVariant()).axis;
    ax[0] = a[0]; ax[1] = a[1]; ax[2] = a[2];
  } else {
    const auto& a = boost::get<pinocchio::JointModelRevoluteUnboundedUnaligned>(jm.toVariant()).axis;
    ax[0] = a[0]; ax[1] = a[1]; ax[2] = a[2];
  }
#endif
}

// pinocchio::Model -> loik_model_desc tables (what the hot path reads from the model: loik-loid-optimized.hxx:46-47,258-265)
template <typename Model>
inline FlatModel flatten(const Model& m) {
  FlatModel out;
  out.njoints = (int)m.njoints; out.nv = (int)m.nv; out.nq = (int)m.nq;
  out.parents.assign(out.njoints, 0); out.joint_types.assign(out.njoints, 0);
  out.joint_axes.assign(3 * out.njoints, 0.0); out.placement_R.assign(9 * out.njoints, 0.0); out.placement_p.assign(3 * out.njoints, 0.0);
  int idx_q = 0, idx_v = 0;  // cumulative, as pinocchio lays q and v out
  for (int i = 0; i < out.njoints; ++i) {
    out.parents[i] = (int32_t)m.parents[i];
    const auto& P = m.jointPlacements[i];
    for (int r = 0; r < 3; ++r) {
      out.placement_p[3 * i + r] = P.translation()[r];
      for (int c = 0; c < 3; ++c) out.placement_R[9 * i + 3 * r + c] = P.rotation()(r, c);  // (r, c): Eigen is column-major, the ABI row-major
    }
    if (i == 0) { out.joint_axes[2] = 1.0; continue; }
    const std::string s = m.joints[i].shortname();
    double ax[3] = {0, 0, 1};
    int code = -1;
    static const struct { const char* name; int code; int axis; } kAligned[] = {
        {"JointModelRX", LOIK_JOINT_RX, 0}, {"JointModelRY", LOIK_JOINT_RY, 1}, {"JointModelRZ", LOIK_JOINT_RZ, 2},
        {"JointModelPX", LOIK_JOINT_PX, 0}, {"JointModelPY", LOIK_JOINT_PY, 1}, {"JointModelPZ", LOIK_JOINT_PZ, 2},
        {"JointModelRUBX", LOIK_JOINT_RUBX, 0}, {"JointModelRUBY", LOIK_JOINT_RUBY, 1}, {"JointModelRUBZ", LOIK_JOINT_RUBZ, 2},
        {"JointModelFreeFlyer", LOIK_JOINT_FF, 2}, {"JointModelSpherical", LOIK_JOINT_SPHERICAL, 2},
        {"JointModelTranslation", LOIK_JOINT_TRANSLATION, 2}, {"JointModelPlanar", LOIK_JOINT_PLANAR, 2},
        {"JointModelSphericalZYX", LOIK_JOINT_SPHERICAL_ZYX, 2}};
    for (const auto& a : kAligned)
      if (s == a.name) { code = a.code; ax[0] = ax[1] = ax[2] = 0.0; ax[a.axis] = 1.0; }
    if (code < 0) {
      if (s == "JointModelRevoluteUnaligned") code = LOIK_JOINT_RU;
      else if (s == "JointModelPrismaticUnaligned") code = LOIK_JOINT_PU;
      else if (s == "JointModelRevoluteUnboundedUnaligned") code = LOIK_JOINT_RUBU;
      else throw std::runtime_error("loik_b200: unsupported joint type " + s + " (joints whose motion subspace depends on q or is stacked)");
      joint_axis(m.joints[i], s, ax);
    }
    if ((int)m.joints[i].idx_v() != idx_v || (int)m.joints[i].idx_q() != idx_q)
      throw std::runtime_error("loik_b200: unexpected idx_q / idx_v layout in pinocchio::Model");
    idx_q += (int)m.joints[i].nq();
    idx_v += (int)m.joints[i].nv();
    out.joint_types[i] = code;
    for (int c = 0; c < 3; ++c) out.joint_axes[3 * i + c] = ax[c];
  }
  return out;
}
}  // namespace detail

template <typename _Scalar>
class FirstOrderLoikOptimizedTpl {
  static_assert(sizeof(_Scalar) == sizeof(double), "libloik_b200 computes in IEEE double (as the reference does in practice: "
                                                   "ik-id-description-optimized.hpp:54-55 hard-codes Eigen::VectorXd)");

 public:
  typedef _Scalar Scalar;
  typedef pinocchio::ModelTpl<Scalar> Model;
  typedef loik::IkIdDataTypeOptimizedTpl<Scalar> IkIdData;
  typedef typename IkIdData::Motion Motion;
  typedef typename IkIdData::Force Force;
  typedef typename IkIdData::SE3 SE3;
  typedef typename IkIdData::DVec DVec;
  typedef typename IkIdData::Vec6 Vec6;
  typedef typename IkIdData::Mat6x6 Mat6x6;
  typedef typename IkIdData::Index Index;

  // loik-loid-optimized.hpp:129-134 (+ the CUDA device and stream at the end)
  FirstOrderLoikOptimizedTpl(const int max_iter, const Scalar& tol_abs, const Scalar& tol_rel, const Scalar& tol_primal_inf,
                             const Scalar& tol_dual_inf, const Scalar& rho, const Scalar& mu, const Scalar& mu_equality_scale_factor,
                             const ADMMPenaltyUpdateStrat& mu_update_strat, const int num_eq_c, const int eq_c_dim, const Model& model,
                             IkIdData& ik_id_data, const bool warm_start, const Scalar tol_tail_solve, const bool verbose,
                             const bool logging, const int device = 0, void* stream = nullptr)
      : flat_(detail::flatten(model)), ik_id_data_(ik_id_data), nj_(flat_.njoints), nb_(flat_.njoints - 1), nv_(flat_.nv),
        nc_(num_eq_c), stream_(stream) {
    loik_model_desc md{flat_.njoints, flat_.parents.data(), flat_.joint_types.data(), flat_.joint_axes.data(),
                       flat_.placement_R.data(), flat_.placement_p.data()};
    loik_params p{max_iter, tol_abs, tol_rel, tol_primal_inf, tol_dual_inf, rho, mu, mu_equality_scale_factor,
                  static_cast<int32_t>(mu_update_strat), num_eq_c, eq_c_dim, warm_start ? 1 : 0, tol_tail_solve,
                  verbose ? 1 : 0, logging ? 1 : 0};
    check(loik_create(&md, &p, 1, device, &h_));
    check(loik_set_keep_workspace(h_, 1));  // His, pis (and UDinv, Dinv, r) stay readable after Solve(), as in the caller's IkIdData
  }
  ~FirstOrderLoikOptimizedTpl() { loik_destroy(h_); }
  FirstOrderLoikOptimizedTpl(const FirstOrderLoikOptimizedTpl&) = delete;
  FirstOrderLoikOptimizedTpl& operator=(const FirstOrderLoikOptimizedTpl&) = delete;

  // hpp:335-338
  void SolveInit(const DVec& q, const Mat6x6& H_ref, const Motion& v_ref, const std::vector<Index>& active_task_constraint_ids,
                 const PINOCCHIO_ALIGNED_STD_VECTOR(Mat6x6)& Ais, const PINOCCHIO_ALIGNED_STD_VECTOR(Vec6)& bis, const DVec& lb,
                 const DVec& ub) {
    Problem pr(*this, q, H_ref, v_ref, active_task_constraint_ids, Ais, bis, lb, ub);
    check(loik_solve_init(h_, pr.q.data(), pr.H.data(), pr.v.data(), (int32_t)pr.ids.size(), pr.ids.data(), pr.A.data(), 0, pr.b.data(), 1,
                          pr.lb.data(), pr.ub.data(), 0, LOIK_HOST, stream_));
    write_back(false);
  }
  // hpp:368
  void Solve() {
    check(loik_solve(h_, stream_));
    write_back(true);
  }
  // hpp:475-478
  void Solve(const DVec& q, const Mat6x6& H_ref, const Motion& v_ref, const std::vector<Index>& active_task_constraint_ids,
             const PINOCCHIO_ALIGNED_STD_VECTOR(Mat6x6)& Ais, const PINOCCHIO_ALIGNED_STD_VECTOR(Vec6)& bis, const DVec& lb, const DVec& ub) {
    Problem pr(*this, q, H_ref, v_ref, active_task_constraint_ids, Ais, bis, lb, ub);
    check(loik_solve_full(h_, pr.q.data(), pr.H.data(), pr.v.data(), (int32_t)pr.ids.size(), pr.ids.data(), pr.A.data(), 0, pr.b.data(), 1,
                          pr.lb.data(), pr.ub.data(), 0, LOIK_HOST, stream_));
    write_back(true);
  }
  // hpp:596-597
  void Solve(const DVec& q, const Index c_id, const Mat6x6& Ai, const Vec6& bi) {
    std::vector<double> qv(q.size()), A(36), b(6);
    for (int k = 0; k < (int)q.size(); ++k) qv[k] = q[k];
    for (int r = 0; r < 6; ++r) { b[r] = bi[r]; for (int c = 0; c < 6; ++c) A[6 * r + c] = Ai(r, c); }
    check(loik_solve_task(h_, qv.data(), (int32_t)c_id, A.data(), 0, b.data(), 1, LOIK_HOST, stream_));
    write_back(true);
  }
  // hpp:168-186: iteration counter, flags, mu and the feasibility scalars; the primal / dual state is kept
  void ResetSolver() { check(loik_reset_solver(h_, stream_)); }

  // getters / setters of IkIdSolverBaseTpl (task-solver-base.hpp:87-141) and of the solver (hpp:698-755)
  int get_iter() const { int32_t v = 0; check(loik_get(h_, LOIK_F_ITER, &v, LOIK_HOST, stream_)); return v; }
  Scalar get_mu() const { double v = 0; check(loik_get(h_, LOIK_F_MU, &v, LOIK_HOST, stream_)); return v; }
  Scalar get_primal_residual() const { return residuals(0); }
  Scalar get_dual_residual() const { return residuals(1); }
  Scalar get_tol_primal() const { return residuals(2); }
  Scalar get_tol_dual() const { return residuals(3); }
  bool get_convergence_status() const { return (status() & 1) != 0; }
  bool get_primal_infeasibility_status() const { return (status() & 2) != 0; }
  bool get_dual_infeasibility_status() const { return false; }  // never evaluated by the optimized path (hxx:572-606)
  int get_max_iter() const { return params().max_iter; }
  Scalar get_tol_abs() const { return params().tol_abs; }
  Scalar get_tol_rel() const { return params().tol_rel; }
  Scalar get_tol_primal_inf() const { return params().tol_primal_inf; }
  Scalar get_tol_dual_inf() const { return params().tol_dual_inf; }
  Scalar get_rho() const { return params().rho; }
  // loik_solver_info_ (hpp:47-127; constructed with logging = true): rows [0, get_iter()) of [capacity][LOIK_HISTORY_COLS] =
  // primal_residual_task, primal_residual_slack, dual_residual_v, dual_residual_nu, mu, delta_x_qp_inf_norm, delta_z_inf_norm, tail flag
  std::vector<double> get_solver_log() const {
    std::vector<double> out((size_t)loik_history_capacity(h_) * LOIK_HISTORY_COLS);
    check(loik_get_history(h_, out.data(), LOIK_HOST, stream_));
    return out;
  }
  void set_max_iter(const int v) { check(loik_set_max_iter(h_, v)); }
  void set_tol_abs(const Scalar v) { check(loik_set_tol_abs(h_, v)); }
  void set_tol_rel(const Scalar v) { check(loik_set_tol_rel(h_, v)); }
  void set_tol_primal_inf(const Scalar v) { check(loik_set_tol_primal_inf(h_, v)); }
  void set_tol_dual_inf(const Scalar v) { check(loik_set_tol_dual_inf(h_, v)); }
  void set_rho(const Scalar v) { check(loik_set_rho(h_, v)); }
  void set_mu(const Scalar v) { check(loik_set_mu(h_, v)); }
  void set_mu_equality_scale_factor(const Scalar v) { check(loik_set_mu_equality_scale_factor(h_, v)); }
  void set_tol_tail_solve(const Scalar v) { check(loik_set_tol_tail_solve(h_, v)); }
  loik_solver* handle() const { return h_; }

 private:
  static void check(int rc) {
    if (rc != LOIK_OK) throw std::runtime_error(loik_last_error());
  }
  // Eigen / aligned-vector arguments -> the row-major arrays of the ABI, with the reference's argument checks
  // (ik-id-description-optimized.hpp:132-150, 328-335)
  struct Problem {
    std::vector<double> q, H, v, A, b, lb, ub;
    std::vector<int32_t> ids;
    Problem(const FirstOrderLoikOptimizedTpl& S, const DVec& q_, const Mat6x6& H_ref, const Motion& v_ref, const std::vector<Index>& ids_,
            const PINOCCHIO_ALIGNED_STD_VECTOR(Mat6x6)& Ais, const PINOCCHIO_ALIGNED_STD_VECTOR(Vec6)& bis, const DVec& lb_, const DVec& ub_)
        : q(q_.size()), H(36), v(6), A(36 * ids_.size()), b(6 * ids_.size()), lb(lb_.size()), ub(ub_.size()), ids(ids_.size()) {
      if (Ais.size() != ids_.size() || bis.size() != ids_.size())
        throw std::runtime_error("[IkProblemFormulation::UpdateEqConstraints]: task_constraint_ids, Ais, and bis have different size !!!");
      if (lb_.size() != ub_.size())
        throw std::runtime_error("[IkProblemFormulation::UpdateIneqConstraints]: lower bound and upper bound have different dimensions!!!");
      if ((int)lb_.size() != S.nv_)
        throw std::runtime_error("IkProblemFormulation::UpdateIneqConstraints]: inequality constraint dimension has changed, this is not supported currently!!!");
      if ((int)q_.size() != S.flat_.nq) throw std::runtime_error("loik_b200: q has the wrong dimension (model.nq)");
      for (int k = 0; k < (int)q_.size(); ++k) q[k] = q_[k];
      const auto& vr = v_ref.toVector();
      for (int r = 0; r < 6; ++r) { v[r] = vr[r]; for (int c = 0; c < 6; ++c) H[6 * r + c] = H_ref(r, c); }
      for (size_t k = 0; k < ids_.size(); ++k) {
        ids[k] = (int32_t)ids_[k];
        for (int r = 0; r < 6; ++r) { b[6 * k + r] = bis[k][r]; for (int c = 0; c < 6; ++c) A[36 * k + 6 * r + c] = Ais[k](r, c); }
      }
      for (int k = 0; k < (int)lb_.size(); ++k) { lb[k] = lb_[k]; ub[k] = ub_[k]; }
    }
  };
  std::vector<double> fetch(int field, int n) const {
    std::vector<double> out((size_t)n);
    check(loik_get(h_, field, out.data(), LOIK_HOST, stream_));
    return out;
  }
  // the caller's IkIdData after the call, as the reference leaves it (joint-indexed members: entries 1..njoints-1)
  void write_back(const bool solved) {
    IkIdData& d = ik_id_data_;
    const auto limi = fetch(LOIK_F_LIMI, 12 * nb_);
    for (int i = 1; i < nj_; ++i) {
      typename SE3::Matrix3 R;
      typename SE3::Vector3 p;
      for (int r = 0; r < 3; ++r) { p[r] = limi[12 * (i - 1) + 9 + r]; for (int c = 0; c < 3; ++c) R(r, c) = limi[12 * (i - 1) + 3 * r + c]; }
      d.liMi[i] = SE3(R, p);
    }
    const auto z = fetch(LOIK_F_Z, nv_), nu = fetch(LOIK_F_NU, nv_), w = fetch(LOIK_F_W, nv_), T = fetch(LOIK_F_STF_PLUS_W, nv_);
    for (int k = 0; k < nv_; ++k) { d.z[k] = z[k]; d.nu[k] = nu[k]; d.w[k] = w[k]; d.Stf_plus_w[k] = T[k]; }
    const auto y = fetch(LOIK_F_Y, 6 * nc_), aty = fetch(LOIK_F_ATY, 6 * nc_);
    for (int k = 0; k < nc_; ++k)
      for (int r = 0; r < 6; ++r) { d.yis[k][r] = y[6 * k + r]; d.Aty[k][r] = aty[6 * k + r]; }
    const auto v = fetch(LOIK_F_V, 6 * nb_), f = fetch(LOIK_F_F, 6 * nb_), F = fetch(LOIK_F_FDPA, 6 * nb_);
    for (int i = 1; i < nj_; ++i) {
      Vec6 a, b, c;
      for (int r = 0; r < 6; ++r) { a[r] = v[6 * (i - 1) + r]; b[r] = f[6 * (i - 1) + r]; c[r] = F[6 * (i - 1) + r]; }
      d.vis[i] = Motion(a); d.fis[i] = Force(b); d.fis_diff_plus_Aty[i] = Force(c);
    }
    if (!solved) return;  // (SolveInit: no backward pass has run yet)
    const auto H = fetch(LOIK_F_H, 36 * nb_), pp = fetch(LOIK_F_P, 6 * nb_), rr = fetch(LOIK_F_R, nv_);
    for (int i = 1; i < nj_; ++i) {
      Vec6 a;
      for (int r = 0; r < 6; ++r) { a[r] = pp[6 * (i - 1) + r]; for (int c = 0; c < 6; ++c) d.His[i](r, c) = H[36 * (i - 1) + 6 * r + c]; }
      d.pis[i] = Force(a);
    }
    for (int k = 0; k < nv_; ++k) d.r[k] = rr[k];
  }
  double residuals(int c) const { double v[4]; check(loik_get(h_, LOIK_F_RESIDUALS, v, LOIK_HOST, stream_)); return v[c]; }
  int status() const { int32_t s = 0; check(loik_get(h_, LOIK_F_STATUS, &s, LOIK_HOST, stream_)); return s; }
  loik_params params() const { loik_params p; check(loik_get_params(h_, &p)); return p; }

  detail::FlatModel flat_;  // (the reference copies the pinocchio::Model, hpp:762: only what the hot path reads is kept here)
  IkIdData& ik_id_data_;
  int nj_, nb_, nv_, nc_;
  void* stream_;
  loik_solver* h_ = nullptr;
};

typedef FirstOrderLoikOptimizedTpl<double> FirstOrderLoikOptimizedPin;

}  // namespace loik_b200

// dump_fixtures.cpp -- runs the UNMODIFIED reference (loik::FirstOrderLoikOptimizedTpl<double>) on its own test fixture
// and writes golden vectors for this repo's oracle / CUDA parity tests.  See README.md.  Not compiled in the offline
// build container (needs Pinocchio); written against the reference's public API as its tests use it
// (/root/reference/tests/loik-loid.cpp:87-165, 559-671).
#include <fstream>
#include <iomanip>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include <pinocchio/algorithm/joint-configuration.hpp>
#include <pinocchio/parsers/urdf.hpp>

#include "loik/loik-loid-optimized.hpp"

using Scalar = double;
using Model = pinocchio::ModelTpl<Scalar>;
using IkIdDataOptimized = loik::IkIdDataTypeOptimizedTpl<Scalar>;
using Motion = typename IkIdDataOptimized::Motion;
using Mat6x6 = typename IkIdDataOptimized::Mat6x6;
using DVec = typename IkIdDataOptimized::DVec;
using Vec6 = typename IkIdDataOptimized::Vec6;
using Index = typename IkIdDataOptimized::Index;
using Solver = loik::FirstOrderLoikOptimizedTpl<Scalar>;

namespace {
struct Json {
  std::ostringstream o;
  bool first = true;
  Json() { o << std::setprecision(17) << "{"; }
  void key(const std::string& k) { o << (first ? "" : ",") << "\n  \"" << k << "\": "; first = false; }
  template <typename T> void num(const std::string& k, const T& v) { key(k); o << v; }
  void str(const std::string& k, const std::string& v) { key(k); o << '"' << v << '"'; }
  template <typename It> void arr(const std::string& k, It b, It e) {
    key(k); o << "[";
    for (It i = b; i != e; ++i) o << (i == b ? "" : ", ") << *i;
    o << "]";
  }
  void vec(const std::string& k, const std::vector<double>& v) { arr(k, v.begin(), v.end()); }
  std::string done() { o << "\n}\n"; return o.str(); }
};

template <typename M> void push_rowmajor(std::vector<double>& out, const M& m) {
  for (Eigen::Index r = 0; r < m.rows(); ++r) for (Eigen::Index c = 0; c < m.cols(); ++c) out.push_back(m(r, c));
}

std::vector<double> joint_axis(const Model& model, Index i) {
  const std::string s = model.joints[i].shortname();
  Eigen::Vector3d a(0, 0, 1);
  if (s == "JointModelRevoluteUnaligned") a = boost::get<pinocchio::JointModelRevoluteUnaligned>(model.joints[i].toVariant()).axis;
  else if (s == "JointModelPrismaticUnaligned") a = boost::get<pinocchio::JointModelPrismaticUnaligned>(model.joints[i].toVariant()).axis;
  else if (s == "JointModelRevoluteUnboundedUnaligned") a = boost::get<pinocchio::JointModelRevoluteUnboundedUnaligned>(model.joints[i].toVariant()).axis;
  else if (s.back() == 'X') a = Eigen::Vector3d(1, 0, 0);
  else if (s.back() == 'Y') a = Eigen::Vector3d(0, 1, 0);
  return {a[0], a[1], a[2]};
}

void dump_case(const std::string& path, const std::string& urdf, const DVec* q_literal, const std::vector<int>& max_iters) {
  Model model;
  pinocchio::urdf::buildModel(urdf, model, false);
  DVec q = pinocchio::neutral(model);
  if (q_literal) q = *q_literal;
  // the fixture's problem and hyper-parameters (tests/loik-loid.cpp:91-131)
  const Scalar tol_abs = 1e-3, tol_rel = 1e-3, tol_primal_inf = 1e-2, tol_dual_inf = 1e-2, tol_tail_solve = 1e-1, rho = 1e-5, mu = 1e-2,
               mu_equality_scale_factor = 1e4, bound_magnitude = 4.0;
  const int num_eq_c = 1, eq_c_dim = 6;
  const Mat6x6 H_ref = Mat6x6::Identity();
  const Motion v_ref = Motion::Zero();
  const std::vector<Index> ids{static_cast<Index>(model.njoints - 1)};
  PINOCCHIO_ALIGNED_STD_VECTOR(Mat6x6) Ais{Mat6x6::Identity()};
  Vec6 bi = Vec6::Zero();
  bi[2] = 0.5;
  PINOCCHIO_ALIGNED_STD_VECTOR(Vec6) bis{bi};
  const DVec lb = -bound_magnitude * DVec::Ones(model.nv), ub = bound_magnitude * DVec::Ones(model.nv);

  Json J;
  J.str("urdf", urdf);
  J.num("njoints", model.njoints); J.num("nq", model.nq); J.num("nv", model.nv);
  {
    std::vector<double> par, ax, plR, plp, iq, iv;
    J.key("joint_shortnames"); J.o << "[";
    for (Index i = 0; i < (Index)model.njoints; ++i) J.o << (i ? ", " : "") << '"' << (i ? model.joints[i].shortname() : std::string("universe")) << '"';
    J.o << "]";
    for (Index i = 0; i < (Index)model.njoints; ++i) {
      par.push_back((double)model.parents[i]);
      const auto a = i ? joint_axis(model, i) : std::vector<double>{0, 0, 1};
      ax.insert(ax.end(), a.begin(), a.end());
      push_rowmajor(plR, model.jointPlacements[i].rotation());
      for (int r = 0; r < 3; ++r) plp.push_back(model.jointPlacements[i].translation()[r]);
      iq.push_back(i ? model.joints[i].idx_q() : 0); iv.push_back(i ? model.joints[i].idx_v() : 0);
    }
    J.vec("parents", par); J.vec("joint_axes", ax); J.vec("placement_R", plR); J.vec("placement_p", plp); J.vec("idx_q", iq); J.vec("idx_v", iv);
  }
  J.arr("q", q.data(), q.data() + q.size());
  J.arr("lb", lb.data(), lb.data() + lb.size()); J.arr("ub", ub.data(), ub.data() + ub.size());
  J.num("task_joint", ids[0]);
  J.arr("b", bi.data(), bi.data() + 6);
  J.num("tol_abs", tol_abs); J.num("tol_rel", tol_rel); J.num("tol_primal_inf", tol_primal_inf); J.num("tol_dual_inf", tol_dual_inf);
  J.num("tol_tail_solve", tol_tail_solve); J.num("rho", rho); J.num("mu", mu); J.num("mu_equality_scale_factor", mu_equality_scale_factor);
  J.key("max_iters"); J.o << "["; for (size_t k = 0; k < max_iters.size(); ++k) J.o << (k ? ", " : "") << max_iters[k]; J.o << "]";

  for (const int max_iter : max_iters) {
    IkIdDataOptimized data(model, num_eq_c);
    Solver solver{max_iter, tol_abs, tol_rel, tol_primal_inf, tol_dual_inf, rho, mu, mu_equality_scale_factor,
                  loik::ADMMPenaltyUpdateStrat::DEFAULT, num_eq_c, eq_c_dim, model, data, false, tol_tail_solve, false, false};
    solver.Solve(q, H_ref, v_ref, ids, Ais, bis, lb, ub);
    const std::string p = "m" + std::to_string(max_iter) + "_";
    J.num(p + "iter", solver.get_iter()); J.num(p + "mu", solver.get_mu());
    J.num(p + "primal_residual", solver.get_primal_residual()); J.num(p + "dual_residual", solver.get_dual_residual());
    J.num(p + "tol_primal", solver.get_tol_primal()); J.num(p + "tol_dual", solver.get_tol_dual());
    J.num(p + "converged", (int)solver.get_convergence_status()); J.num(p + "primal_infeasible", (int)solver.get_primal_infeasibility_status());
    J.arr(p + "z", data.z.data(), data.z.data() + data.z.size());
    J.arr(p + "nu", data.nu.data(), data.nu.data() + data.nu.size());
    J.arr(p + "w", data.w.data(), data.w.data() + data.w.size());
    std::vector<double> y, v, f, H, pp, R, t;
    for (int r = 0; r < 6; ++r) y.push_back(data.yis[0][r]);
    for (const auto& idx : data.joint_range) {
      const Vec6 vv = data.vis[idx].toVector(), ff = data.fis[idx].toVector(), pv = data.pis[idx].toVector();
      for (int r = 0; r < 6; ++r) { v.push_back(vv[r]); f.push_back(ff[r]); pp.push_back(pv[r]); }
      push_rowmajor(H, data.His[idx]);
      push_rowmajor(R, data.liMi[idx].rotation());
      for (int r = 0; r < 3; ++r) t.push_back(data.liMi[idx].translation()[r]);
    }
    J.vec(p + "yis", y); J.vec(p + "vis", v); J.vec(p + "fis", f); J.vec(p + "His", H); J.vec(p + "pis", pp);
    J.vec(p + "liMi_R", R); J.vec(p + "liMi_p", t);
  }
  std::ofstream(path) << J.done();
  std::cout << "wrote " << path << "\n";
}
}  // namespace

int main(int argc, char** argv) {
  const std::string out = argc > 1 ? argv[1] : ".";
  const std::string dir = EXAMPLE_ROBOT_DATA_MODEL_DIR;
  dump_case(out + "/ref_talos_fixture.json", dir + "/talos_data/robots/talos_full_v2.urdf", nullptr, {2, 8, 200});
  DVec q(9);
  q << -2.79684649, -0.55090374, 0.424806, -1.21112304, -0.89856966, 0.79726132, -0.07125267, 0.13154589, 0.13171856;
  dump_case(out + "/ref_panda_q.json", dir + "/panda_description/urdf/panda.urdf", &q, {2, 8, 200});
  return 0;
}

// Compiles and links the C++ facade against libloik_b200.so; with a GPU it solves the reference fixture shape
// (tests/loik-loid.cpp:87-165: neutral q, H_ref = I, v_ref = 0, one task A = I, b = (0,0,.5,0,0,0), bounds +-4)
// on a 3-joint chain and prints z.  Without a GPU the constructor must throw (no CPU fallback).
#include <cmath>
#include <cstdio>
#include <vector>

#include "loik_b200/first_order_loik_optimized.hpp"

int main() {
  loik_b200::Model m;
  m.njoints = 4; m.nv = 3;
  m.parents = {0, 0, 1, 2};
  m.joint_types = {0, LOIK_JOINT_RZ, LOIK_JOINT_RY, LOIK_JOINT_RX};
  m.joint_axes = {0, 0, 1, 0, 0, 1, 0, 1, 0, 1, 0, 0};
  m.placement_R.assign(9 * 4, 0.0);
  for (int i = 0; i < 4; ++i) for (int k = 0; k < 3; ++k) m.placement_R[9 * i + 4 * k] = 1.0;
  m.placement_p = {0, 0, 0, 0, 0, 0.3, 0.1, 0, 0.2, 0, 0.1, 0.3};
  const int B = 2;
  try {
    loik_b200::FirstOrderLoikOptimized solver(200, 1e-3, 1e-3, 1e-2, 1e-2, 1e-5, 1e-2, 1e4, loik_b200::DEFAULT, 1, 6, m, B, false, 1e-1,
                                              false, false);
    std::vector<double> q(B * 3, 0.1), H(36, 0.0), v(6, 0.0), A(36, 0.0), b(B * 6, 0.0), lb(3, -4.0), ub(3, 4.0);
    for (int k = 0; k < 6; ++k) H[7 * k] = A[7 * k] = 1.0;
    b[2] = 0.5; b[6 + 2] = 0.25;
    solver.Solve(q, H, v, {3}, A, b, lb, ub);
    auto z = solver.z();
    auto it = solver.get_iter();
    std::printf("facade ok: iter %d %d z0 %.6f %.6f %.6f\n", it[0], it[1], z[0], z[1], z[2]);
    {  // the reference's public per-step methods, driven the way tests/loik-loid.cpp:340-478 drives them: two iterations by
       // hand must land on the iterate Solve() reaches with max_iter = 3 (= 2 iterations).  Tolerances of 1e-12 (convergence and
       // primal infeasibility) keep both instances from stopping earlier, so the comparison cannot pass vacuously.
      loik_b200::FirstOrderLoikOptimized a(3, 1e-12, 1e-12, 1e-12, 1e-2, 1e-5, 1e-2, 1e4, loik_b200::DEFAULT, 1, 6, m, B, false, 1e-1, false, false);
      loik_b200::FirstOrderLoikOptimized s2(3, 1e-12, 1e-12, 1e-12, 1e-2, 1e-5, 1e-2, 1e4, loik_b200::DEFAULT, 1, 6, m, B, false, 1e-1, false, false);
      a.SolveInit(q, H, v, {3}, A, b, lb, ub);
      a.Solve();
      s2.set_debug(true);
      s2.SolveInit(q, H, v, {3}, A, b, lb, ub);
      s2.ResetRecursion();
      for (int itn = 1; itn <= 2; ++itn) {
        s2.UpdatePrev(); s2.ResetInfNorms(); s2.FwdPass1(); s2.BwdPassOptimizedVisitor(); s2.FwdPass2OptimizedVisitor(); s2.BoxProj();
        s2.DualUpdate(); s2.ComputeResiduals(); s2.CheckConvergence();
        if (itn > 1) s2.CheckFeasibility();
        s2.UpdateMu();
      }
      auto za = a.z(), zs = s2.z();
      auto ita = a.get_iter();
      auto mua = a.get_mu(), mus = s2.get_mu();
      double worst = 0.0;
      for (size_t k = 0; k < za.size(); ++k) worst = std::fmax(worst, std::fabs(za[k] - zs[k]));
      if (ita[0] != 2 || ita[1] != 2) { std::printf("the fixture must run exactly two iterations (ran %d, %d)\n", ita[0], ita[1]); return 1; }
      if (worst > 1e-12) { std::printf("per-step methods differ from Solve(): %.3e\n", worst); return 1; }
      if (mua[0] != mus[0] || mua[1] != mus[1]) { std::printf("per-step methods: mu differs from Solve(): %g %g vs %g %g\n", mua[0], mua[1], mus[0], mus[1]); return 1; }
      (void)s2.get_delta_x_qp_inf_norm(); (void)s2.get_primal_infeasibility_cond_1(); (void)s2.get_primal_residual_vec(); (void)s2.His();
      // ResetSolver() keeps the state (hpp:168-186), ResetRecursion() clears it
      a.ResetSolver();
      auto zk = a.z();
      if (a.get_iter()[0] != 0 || zk[0] != za[0]) { std::printf("ResetSolver must reset the counters and keep z\n"); return 1; }
      a.ResetRecursion();
      if (a.z()[0] != 0.0) { std::printf("ResetRecursion must clear z\n"); return 1; }
      std::printf("per-step ok (two iterations compared, max |dz| %.1e)\n", worst);
    }
    bool threw = false;
    try { solver.Solve(q, 1, A, std::vector<double>(6, 0.0)); } catch (const std::runtime_error& e) { threw = true; }
    if (!threw) { std::printf("missing error for unknown constraint id\n"); return 1; }
  } catch (const std::runtime_error& e) {
    std::printf("runtime_error: %s\n", e.what());
    return 2;
  }
  return 0;
}

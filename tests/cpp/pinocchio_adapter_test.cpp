// The reference-shaped facade (include/loik_b200/loik_pinocchio.hpp) compiled against the stand-in Pinocchio / Eigen /
// IkIdData of tests/cpp/stub/ and checked against the ORACLE (oracle/loik_oracle.c, linked in): after SolveInit + Solve(),
// Solve(q, ...8 arguments) and the tailored Solve(q, c_id, Ai, bi), every field the reference's tests read from the
// caller-owned IkIdData (tests/loik-loid.cpp:597-615: His, pis, vis, fis, nu, z, w, yis) plus Aty, fis_diff_plus_Aty,
// Stf_plus_w, r, liMi and the solver's scalars must equal the oracle's at 1e-10 abs-or-rel (the reference's comparator,
// tests/loik-loid.cpp:39-83).  Without a GPU the constructor must throw (no CPU fallback): exit code 2.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "loik_b200/loik_pinocchio.hpp"

extern "C" {
struct lo_solver;
lo_solver* lo_create(int nj, const int* parent, const int* jtype, const double* axis, const double* plR, const double* plp, int max_iter,
                     double tol_abs, double tol_rel, double tol_primal_inf, double tol_dual_inf, double rho, double mu,
                     double mu_equality_scale_factor, int mu_update_strat, int num_eq_c, int eq_c_dim, int warm_start, double tol_tail_solve);
void lo_destroy(lo_solver* s);
int lo_solve_init(lo_solver* s, const double* q, const double* H_ref, const double* v_ref, int n_ids, const int* ids, const double* Ais,
                  const double* bis, const double* lb, const double* ub);
int lo_solve(lo_solver* s);
int lo_solve_full(lo_solver* s, const double* q, const double* H_ref, const double* v_ref, int n_ids, const int* ids, const double* Ais,
                  const double* bis, const double* lb, const double* ub);
int lo_solve_task(lo_solver* s, const double* q, int c_id, const double* Ai, const double* bi);
double* lo_array(lo_solver* s, const char* field, int* n);
double lo_scalar(lo_solver* s, const char* field);
}

namespace {
struct JointRow { int parent; const char* name; int code; double axis[3]; double rpy[3]; double xyz[3]; };
// a small humanoid-like tree: two chains below the base, a torso carrying two arms; aligned, unaligned and prismatic joints
const JointRow kTree[] = {
    {0, "JointModelRZ", LOIK_JOINT_RZ, {0, 0, 1}, {0.1, -0.2, 0.3}, {0.0, 0.1, -0.2}},
    {1, "JointModelRX", LOIK_JOINT_RX, {1, 0, 0}, {0.4, 0.0, -0.3}, {0.02, 0.0, -0.3}},
    {2, "JointModelRY", LOIK_JOINT_RY, {0, 1, 0}, {0.0, 0.3, 0.2}, {0.0, 0.05, -0.35}},
    {0, "JointModelRevoluteUnaligned", LOIK_JOINT_RU, {0.6, 0.0, 0.8}, {-0.3, 0.2, 0.1}, {0.0, -0.1, -0.2}},
    {4, "JointModelPZ", LOIK_JOINT_PZ, {0, 0, 1}, {0.2, 0.2, 0.2}, {0.03, 0.0, -0.3}},
    {0, "JointModelRZ", LOIK_JOINT_RZ, {0, 0, 1}, {0.0, 0.1, 0.0}, {0.0, 0.0, 0.1}},
    {6, "JointModelRY", LOIK_JOINT_RY, {0, 1, 0}, {0.3, -0.1, 0.5}, {0.0, 0.0, 0.2}},
    {7, "JointModelRX", LOIK_JOINT_RX, {1, 0, 0}, {0.2, 0.4, -0.6}, {0.0, 0.2, 0.25}},
    {8, "JointModelPrismaticUnaligned", LOIK_JOINT_PU, {0.0, 0.6, 0.8}, {-0.5, 0.1, 0.2}, {0.25, 0.0, 0.0}},
    {7, "JointModelRZ", LOIK_JOINT_RZ, {0, 0, 1}, {0.1, 0.1, 0.7}, {0.0, -0.2, 0.25}},
    {10, "JointModelRY", LOIK_JOINT_RY, {0, 1, 0}, {-0.2, 0.3, 0.1}, {0.27, 0.0, 0.0}},
};
constexpr int kNb = sizeof(kTree) / sizeof(kTree[0]);

void rpy_to_R(const double* rpy, double (&R)[9]) {
  const double cr = std::cos(rpy[0]), sr = std::sin(rpy[0]), cp = std::cos(rpy[1]), sp = std::sin(rpy[1]), cy = std::cos(rpy[2]), sy = std::sin(rpy[2]);
  const double M[9] = {cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr, sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr, -sp, cp * sr, cp * cr};
  for (int i = 0; i < 9; ++i) R[i] = M[i];
}

int g_fail = 0;
void expect_close(const double* a, const double* b, int n, const std::string& what) {  // tests/loik-loid.cpp:39-83
  double d = 0, scale = 0;
  for (int i = 0; i < n; ++i) { d = std::fmax(d, std::fabs(a[i] - b[i])); scale = std::fmax(scale, std::fmax(std::fabs(a[i]), std::fabs(b[i]))); }
  if (!(d < 1e-10 || d < 1e-10 * scale)) { std::printf("MISMATCH %s: |a-b|inf = %.3e (scale %.3e)\n", what.c_str(), d, scale); ++g_fail; }
}
}  // namespace

int main() {
  using Solver = loik_b200::FirstOrderLoikOptimizedTpl<double>;
  using Data = Solver::IkIdData;
  const int nj = kNb + 1;
  // the stub pinocchio::Model and, independently, the oracle's flat tables, from the same joint rows
  pinocchio::Model model;
  std::vector<int> parent(nj, 0), jtype(nj, 0);
  std::vector<double> axis(3 * nj, 0.0), plR(9 * nj, 0.0), plp(3 * nj, 0.0);
  for (int k = 0; k < 3; ++k) plR[4 * k] = 1.0;
  axis[2] = 1.0;
  for (int i = 1; i < nj; ++i) {
    const JointRow& J = kTree[i - 1];
    double R[9];
    rpy_to_R(J.rpy, R);
    pinocchio::SE3::Matrix3 Rm;
    pinocchio::SE3::Vector3 pm;
    for (int r = 0; r < 3; ++r) { pm[r] = J.xyz[r]; for (int c = 0; c < 3; ++c) Rm(r, c) = R[3 * r + c]; }
    model.addJoint((pinocchio::JointIndex)J.parent, J.name, pinocchio::SE3(Rm, pm), 1, 1, J.axis[0], J.axis[1], J.axis[2]);
    parent[i] = J.parent; jtype[i] = J.code;
    for (int c = 0; c < 3; ++c) { axis[3 * i + c] = J.axis[c]; plp[3 * i + c] = J.xyz[c]; }
    for (int c = 0; c < 9; ++c) plR[9 * i + c] = R[c];
  }
  const int nv = model.nv, nc = 2;
  const int max_iter = 60;
  const double tol_abs = 1e-3, tol_rel = 1e-3, tol_pinf = 1e-2, tol_dinf = 1e-2, rho = 1e-5, mu = 1e-2, mu_scale = 1e4, tol_tail = 1e-1;
  // problem: non-symmetric task matrices (row- vs column-major mistakes would show), symmetric H_ref, two tasks
  Data::DVec q(nv), lb(nv), ub(nv);
  for (int k = 0; k < nv; ++k) { q[k] = 0.3 * std::sin(1.0 + k); lb[k] = -1.5 - 0.1 * k; ub[k] = 2.0 + 0.05 * k; }
  Data::Mat6x6 H_ref;
  for (int r = 0; r < 6; ++r) for (int c = 0; c < 6; ++c) H_ref(r, c) = (r == c ? 1.0 + 0.1 * r : 0.0) + 0.02 * (r + c);
  Data::Vec6 vr;
  for (int r = 0; r < 6; ++r) vr[r] = 0.05 * (r - 2);
  const pinocchio::Motion v_ref(vr);
  std::vector<Solver::Index> ids = {3, 11};
  PINOCCHIO_ALIGNED_STD_VECTOR(Data::Mat6x6) Ais(2);
  PINOCCHIO_ALIGNED_STD_VECTOR(Data::Vec6) bis(2);
  for (int k = 0; k < 2; ++k)
    for (int r = 0; r < 6; ++r) {
      bis[k][r] = 0.2 * std::cos(0.7 * r + k);
      for (int c = 0; c < 6; ++c) Ais[k](r, c) = (r == c ? 1.0 : 0.0) + 0.05 * std::sin(1.0 + 3 * r + c + 5 * k);
    }
  // row-major copies for the oracle
  std::vector<double> qo(nv), lbo(nv), ubo(nv), Ho(36), vo(6), Ao(72), bo(12);
  std::vector<int> ido = {3, 11};
  for (int k = 0; k < nv; ++k) { qo[k] = q[k]; lbo[k] = lb[k]; ubo[k] = ub[k]; }
  for (int r = 0; r < 6; ++r) { vo[r] = vr[r]; for (int c = 0; c < 6; ++c) Ho[6 * r + c] = H_ref(r, c); }
  for (int k = 0; k < 2; ++k) for (int r = 0; r < 6; ++r) { bo[6 * k + r] = bis[k][r]; for (int c = 0; c < 6; ++c) Ao[36 * k + 6 * r + c] = Ais[k](r, c); }

  lo_solver* O = lo_create(nj, parent.data(), jtype.data(), axis.data(), plR.data(), plp.data(), max_iter, tol_abs, tol_rel, tol_pinf, tol_dinf,
                           rho, mu, mu_scale, 0, nc, 6, 0, tol_tail);
  if (!O) { std::printf("oracle: lo_create failed\n"); return 1; }
  Data data(model, nc);
  try {
    Solver S(max_iter, tol_abs, tol_rel, tol_pinf, tol_dinf, rho, mu, mu_scale, loik_b200::DEFAULT, nc, 6, model, data, false, tol_tail, false, false);
    auto compare = [&](const std::string& tag, bool workspace) {
      int n = 0;
      auto arr = [&](const char* f) { return lo_array(O, f, &n); };
      expect_close(data.z.data(), arr("z"), nv, tag + " z");
      expect_close(data.nu.data(), arr("nu"), nv, tag + " nu");
      expect_close(data.w.data(), arr("w"), nv, tag + " w");
      expect_close(data.Stf_plus_w.data(), arr("Stf_plus_w"), nv, tag + " Stf_plus_w");
      for (int k = 0; k < nc; ++k) {
        expect_close(data.yis[k].data(), arr("yis") + 6 * k, 6, tag + " yis");
        expect_close(data.Aty[k].data(), arr("Aty") + 6 * k, 6, tag + " Aty");
      }
      for (int i = 1; i < nj; ++i) {
        expect_close(data.vis[i].toVector().data(), arr("vis") + 6 * i, 6, tag + " vis");
        expect_close(data.fis[i].toVector().data(), arr("fis") + 6 * i, 6, tag + " fis");
        double R[9], Hrow[36];
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) R[3 * r + c] = data.liMi[i].rotation()(r, c);
        expect_close(R, arr("liMi_R") + 9 * i, 9, tag + " liMi rotation");
        expect_close(data.liMi[i].translation().data(), arr("liMi_p") + 3 * i, 3, tag + " liMi translation");
        if (workspace) {
          for (int r = 0; r < 6; ++r) for (int c = 0; c < 6; ++c) Hrow[6 * r + c] = data.His[i](r, c);
          expect_close(Hrow, arr("His") + 36 * i, 36, tag + " His");
          expect_close(data.pis[i].toVector().data(), arr("pis") + 6 * i, 6, tag + " pis");
        }
      }
      if (workspace) {
        expect_close(data.r.data(), arr("r"), nv, tag + " r");
        const double it = S.get_iter(), muv = S.get_mu(), pr = S.get_primal_residual(), dr = S.get_dual_residual();
        const double oit = lo_scalar(O, "iter"), omu = lo_scalar(O, "mu"), opr = lo_scalar(O, "primal_residual"), odr = lo_scalar(O, "dual_residual");
        expect_close(&it, &oit, 1, tag + " iter"); expect_close(&muv, &omu, 1, tag + " mu");
        expect_close(&pr, &opr, 1, tag + " primal_residual");
        if (!(std::fabs(dr - odr) < 1e-9 * std::fmax(1.0, std::fabs(odr)))) { std::printf("MISMATCH %s dual_residual\n", tag.c_str()); ++g_fail; }
        if (S.get_convergence_status() != (lo_scalar(O, "converged") != 0.0)) { std::printf("MISMATCH %s converged flag\n", tag.c_str()); ++g_fail; }
      }
    };
    // SolveInit + Solve()
    S.SolveInit(q, H_ref, v_ref, ids, Ais, bis, lb, ub);
    lo_solve_init(O, qo.data(), Ho.data(), vo.data(), nc, ido.data(), Ao.data(), bo.data(), lbo.data(), ubo.data());
    compare("SolveInit", false);
    S.Solve();
    lo_solve(O);
    compare("Solve()", true);
    std::printf("Solve(): %d iterations, converged %d, z0 = %.9f\n", S.get_iter(), (int)S.get_convergence_status(), data.z[0]);
    if (S.get_iter() < 3) { std::printf("the fixture must run several iterations\n"); ++g_fail; }
    // Solve(q, H_ref, v_ref, ids, Ais, bis, lb, ub) with a different q
    for (int k = 0; k < nv; ++k) { q[k] = 0.2 * std::cos(2.0 + k); qo[k] = q[k]; }
    S.Solve(q, H_ref, v_ref, ids, Ais, bis, lb, ub);
    lo_solve_full(O, qo.data(), Ho.data(), vo.data(), nc, ido.data(), Ao.data(), bo.data(), lbo.data(), ubo.data());
    compare("Solve(8 args)", true);
    // tailored Solve(q, c_id, Ai, bi)
    Data::Mat6x6 A2 = Ais[1];
    Data::Vec6 b2 = bis[1];
    for (int r = 0; r < 6; ++r) { b2[r] += 0.03 * r; A2(r, (r + 1) % 6) += 0.02; }
    double A2o[36], b2o[6];
    for (int r = 0; r < 6; ++r) { b2o[r] = b2[r]; for (int c = 0; c < 6; ++c) A2o[6 * r + c] = A2(r, c); }
    S.Solve(q, 11, A2, b2);
    lo_solve_task(O, qo.data(), 11, A2o, b2o);
    compare("Solve(q, c_id, Ai, bi)", true);
    // ResetSolver keeps the state, resets the counters (hpp:168-186)
    const double z0 = data.z[0];
    S.ResetSolver();
    if (S.get_iter() != 0 || S.get_mu() != mu) { std::printf("ResetSolver: iter / mu not reset\n"); ++g_fail; }
    std::vector<double> zz(nv);
    loik_get(S.handle(), LOIK_F_Z, zz.data(), LOIK_HOST, nullptr);
    if (zz[0] != z0) { std::printf("ResetSolver: z was cleared\n"); ++g_fail; }
    // error paths keep the reference's messages
    bool threw = false;
    try { S.Solve(q, 5, A2, b2); } catch (const std::runtime_error& e) { threw = std::strstr(e.what(), "constraint doesn't yet exist") != nullptr; }
    if (!threw) { std::printf("missing error for an unknown constraint id\n"); ++g_fail; }
    threw = false;
    try { Data::DVec bad(nv + 1); S.SolveInit(q, H_ref, v_ref, ids, Ais, bis, bad, bad); } catch (const std::runtime_error&) { threw = true; }
    if (!threw) { std::printf("missing error for a bound vector of the wrong size\n"); ++g_fail; }
  } catch (const std::runtime_error& e) {
    std::printf("runtime_error: %s\n", e.what());
    lo_destroy(O);
    return 2;
  }
  lo_destroy(O);
  if (g_fail) { std::printf("%d mismatches\n", g_fail); return 1; }
  std::printf("pinocchio adapter ok\n");
  return 0;
}

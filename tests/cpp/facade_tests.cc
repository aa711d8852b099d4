// facade_tests.cc -- the reference's solver-level tests, written against the
// C++ facade (include/fbstab/*.h) of the B200 engine.
//
// Each TEST restates one googletest case of the reference
// (fbstab/test/fbstab_dense_unit_tests.cc:28-256,
//  fbstab/test/fbstab_mpc_unit_tests.cc:15-148) with the same data, options
// and assertions; gtest is not in this image, so a few macros stand in.
//
//   facade_tests          -> all cases (needs a GPU)
//   facade_tests --host   -> host-only cases: types, size validation and
//                            error behaviour (no CUDA call succeeds without a
//                            device, and the facade must say so loudly)
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <memory>
#include <string>
#include <vector>

#include "fbstab/closed_loop.h"
#include "fbstab/fbstab_dense.h"
#include "fbstab/fbstab_mpc.h"
#include "fbstab/fbstab_sparse.h"
#include "fbstab/ocp_generator.h"

namespace {

struct Case {
  const char* name;
  bool needs_gpu;
  std::function<void()> fn;
};
std::vector<Case>& Cases() {
  static std::vector<Case> c;
  return c;
}
struct Reg {
  Reg(const char* n, bool g, std::function<void()> f) { Cases().push_back({n, g, f}); }
};
int g_failures = 0;

#define TEST(suite, name, gpu)                                  \
  void suite##_##name();                                        \
  Reg reg_##suite##_##name(#suite "." #name, gpu, suite##_##name); \
  void suite##_##name()
#define FAIL_(msg)                                                      \
  do {                                                                  \
    printf("    %s:%d: %s\n", __FILE__, __LINE__, std::string(msg).c_str()); \
    g_failures++;                                                       \
  } while (0)
#define EXPECT_TRUE(c) \
  do {                 \
    if (!(c)) FAIL_("expected " #c); \
  } while (0)
#define ASSERT_EQ(a, b) EXPECT_TRUE((a) == (b))
#define ASSERT_LE(a, b) EXPECT_TRUE((a) <= (b))
#define EXPECT_NEAR(a, b, tol)                                            \
  do {                                                                    \
    if (!(std::fabs((a) - (b)) <= (tol)))                                 \
      FAIL_(std::string(#a " vs " #b ": ") + std::to_string((double)(a)) + \
            " vs " + std::to_string((double)(b)));                        \
  } while (0)
#define EXPECT_THROW(stmt, what_substr)                                   \
  do {                                                                    \
    bool thrown = false;                                                  \
    try {                                                                 \
      stmt;                                                               \
    } catch (const std::exception& e) {                                   \
      thrown = true;                                                      \
      if (std::string(e.what()).find(what_substr) == std::string::npos)   \
        FAIL_(std::string("wrong message: ") + e.what());                 \
    }                                                                     \
    if (!thrown) FAIL_("no exception from " #stmt);                       \
  } while (0)

}  // namespace

namespace fbstab {
namespace test {

using MatrixXd = Eigen::MatrixXd;
using VectorXd = Eigen::VectorXd;

// ---- fbstab_dense_unit_tests.cc:28-61 ----------------------------------------
TEST(FBstabDense, FeasibleQP, true) {
  int n = 2, m = 0, q = 2;
  FBstabDense::Variable x0(n, m, q);
  FBstabDense::ProblemData data(n, m, q);
  data.H << 3, 1, 1, 1;
  data.f << 10, 5;
  data.A << -1, 0, 0, 1;
  data.b << 0, 0;

  FBstabDense solver(n, m, q);
  FBstabDense::Options opts = FBstabDense::DefaultOptions();
  opts.abs_tol = 1e-8;
  opts.display_level = Display::OFF;
  solver.UpdateOptions(opts);

  SolverOut out = solver.Solve(data, &x0);
  ASSERT_EQ(out.eflag, ExitFlag::SUCCESS);

  VectorXd zopt(2), vopt(2);
  zopt << 0, -5;
  vopt << 5, 0;
  for (int i = 0; i < n; i++) EXPECT_NEAR(x0.z(i), zopt(i), 1e-8);
  for (int i = 0; i < q; i++) EXPECT_NEAR(x0.v(i), vopt(i), 1e-8);
}

// ---- fbstab_dense_unit_tests.cc:75-104 ---------------------------------------
TEST(FBstabDense, FeasibleQPwithEQ, true) {
  int n = 2, m = 1, q = 2;
  FBstabDense::Variable x0(n, m, q);
  FBstabDense::ProblemData data(n, m, q);
  data.H << 4, 1, 1, 2;
  data.f << 1, 1;
  data.G << 1, 1;
  data.h << 1;
  data.A << -1, 0, 0, -1;
  data.b << 0, 0;

  FBstabDense solver(n, m, q);
  FBstabDense::Options opts = FBstabDense::DefaultOptions();
  opts.abs_tol = 1e-8;
  opts.display_level = Display::OFF;
  solver.UpdateOptions(opts);

  SolverOut out = solver.Solve(data, &x0);
  ASSERT_EQ(out.eflag, ExitFlag::SUCCESS);
  VectorXd zopt(n);
  zopt << 2.5e-1, 7.5e-1;
  for (int i = 0; i < n; i++) EXPECT_NEAR(x0.z(i), zopt(i), 1e-8);
}

// ---- fbstab_dense_unit_tests.cc:121-177 (Ref types over raw memory) ----------
TEST(FBstabDense, DegenerateQP, true) {
  constexpr int n = 2, m = 0, q = 5;
  std::unique_ptr<double[]> zmem(new double[n]), lmem(new double[m]), vmem(new double[q]),
      ymem(new double[q]);
  Eigen::Map<VectorXd> z(zmem.get(), n), l(lmem.get(), m), v(vmem.get(), q), y(ymem.get(), q);
  FBstabDense::VariableRef x0(&z, &l, &v, &y);
  x0.fill(0.0);

  std::unique_ptr<double[]> Hmem(new double[n * n]), fmem(new double[n]),
      Gmem(new double[m * n]), hmem(new double[m]), Amem(new double[q * n]),
      bmem(new double[q]);
  Eigen::Map<MatrixXd> H(Hmem.get(), n, n), A(Amem.get(), q, n), G(Gmem.get(), m, n);
  Eigen::Map<VectorXd> f(fmem.get(), n), b(bmem.get(), q), h(hmem.get(), m);
  H << 1, 0, 0, 0;
  f << 1, 0;
  A << 0, 0, 1, 0, 0, 1, -1, 0, 0, -1;
  b << 0, 3, 3, -1, -1;

  FBstabDense::ProblemDataRef data(&H, &f, &G, &h, &A, &b);
  FBstabDense solver(n, m, q);
  FBstabDense::Options opts = FBstabDense::DefaultOptions();
  opts.abs_tol = 1e-8;
  opts.display_level = Display::OFF;
  solver.UpdateOptions(opts);

  SolverOut out = solver.Solve(data, &x0);
  ASSERT_EQ(out.eflag, ExitFlag::SUCCESS);
  EXPECT_NEAR(x0.z(0), 1, 1e-8);
  EXPECT_TRUE((x0.z(1) >= 1) && (x0.z(1) <= 3));

  // KKT: r1 = Hz + f + A'v, r2 = min(y, v)
  double r1sq = 0, r2sq = 0;
  for (int i = 0; i < n; i++) {
    double s = data.f(i);
    for (int j = 0; j < n; j++) s += data.H(i, j) * x0.z(j);
    for (int k = 0; k < q; k++) s += data.A(k, i) * x0.v(k);
    r1sq += s * s;
  }
  for (int k = 0; k < q; k++) {
    const double t = std::fmin(x0.y(k), x0.v(k));
    r2sq += t * t;
  }
  EXPECT_NEAR(std::sqrt(r1sq) + std::sqrt(r2sq), 0, 1e-6);
}

// ---- fbstab_dense_unit_tests.cc:195-217 --------------------------------------
TEST(FBstabDense, InfeasibleQP, true) {
  int n = 2, m = 0, q = 5;
  FBstabDense::ProblemData data(n, m, q);
  data.H << 1, 0, 0, 0;
  data.f << 1, -1;
  data.A << 1, 1, 1, 0, 0, 1, -1, 0, 0, -1;
  data.b << 0, 3, 3, -1, -1;
  FBstabDense::Variable x0(n, m, q);
  FBstabDense solver(n, m, q);
  FBstabDense::Options opts = FBstabDense::DefaultOptions();
  opts.abs_tol = 1e-8;
  opts.display_level = Display::OFF;
  solver.UpdateOptions(opts);
  SolverOut out = solver.Solve(data, &x0);
  ASSERT_EQ(out.eflag, ExitFlag::PRIMAL_INFEASIBLE);
}

// ---- fbstab_dense_unit_tests.cc:233-256 --------------------------------------
TEST(FBstabDense, UnboundedQP, true) {
  int n = 2, m = 0, q = 4;
  FBstabDense::ProblemData data(n, m, q);
  data.H << 1, 0, 0, 0;
  data.f << 1, -1;
  data.A << 0, 0, 1, 0, -1, 0, 0, -1;
  data.b << 0, 3, -1, -1;
  FBstabDense::Variable x0(n, m, q);
  FBstabDense solver(n, m, q);
  FBstabDense::Options opts = FBstabDense::DefaultOptions();
  opts.abs_tol = 1e-8;
  opts.display_level = Display::OFF;
  solver.UpdateOptions(opts);
  SolverOut out = solver.Solve(data, &x0);
  ASSERT_EQ(out.eflag, ExitFlag::DUAL_INFEASIBLE);
}

// ---- batched entry: a batch of the single-instance structs -------------------
TEST(FBstabDense, SolveBatchOfStructs, true) {
  int n = 2, m = 0, q = 2;
  std::vector<FBstabDense::QPData> qps;
  std::vector<FBstabDense::QPVariable> xs;
  for (int i = 0; i < 5; i++) {
    FBstabDense::QPData d(n, m, q);
    d.H << 3, 1, 1, 1;
    d.f << 10 + i, 5;
    d.A << -1, 0, 0, 1;
    d.b << 0, 0;
    qps.push_back(d);
    xs.emplace_back(n, m, q);
  }
  FBstabDense solver(n, m, q, /*max_batch=*/8);
  FBstabDense::Options opts = FBstabDense::DefaultOptions();
  opts.abs_tol = 1e-8;
  opts.display_level = Display::OFF;
  solver.UpdateOptions(opts);
  std::vector<SolverOut> outs = solver.SolveBatch(qps, &xs);
  ASSERT_EQ(outs.size(), (size_t)5);
  for (int i = 0; i < 5; i++) {
    ASSERT_EQ(outs[i].eflag, ExitFlag::SUCCESS);
    ASSERT_EQ(outs[i].status, 0);
    // same instance alone
    FBstabDense::Variable x1(n, m, q);
    SolverOut o1 = solver.Solve(qps[i], &x1);
    ASSERT_EQ(o1.newton_iters, outs[i].newton_iters);
    for (int k = 0; k < n; k++) EXPECT_NEAR(xs[i].z(k), x1.z(k), 0.0);
    EXPECT_NEAR(xs[i].z(0), 0.0, 1e-8);
    EXPECT_NEAR(xs[i].z(1), -5.0, 1e-8);
  }
}

// ---- fbstab_mpc_unit_tests.cc:15-60 -------------------------------------------
TEST(FBstabMpc, DoubleIntegrator, true) {
  OcpGenerator ocp;
  ocp.DoubleIntegrator(2);
  FBstabMpc::ProblemData data = ocp.GetFBstabInput();
  FBstabMpc::Variable x(ocp.ProblemSize());
  FBstabMpc solver(ocp.ProblemSize());
  FBstabMpc::Options opts = FBstabMpc::DefaultOptions();
  opts.abs_tol = 1e-8;
  opts.display_level = Display::OFF;
  solver.UpdateOptions(opts);
  SolverOut out = solver.Solve(data, &x);
  ASSERT_EQ(out.eflag, ExitFlag::SUCCESS);
  ASSERT_LE(out.residual, 1e-6);

  VectorXd zopt(ocp.nz()), lopt(ocp.nl()), vopt(ocp.nv());
  // "Computed using MATLAB's quadprog command", fbstab_mpc_unit_tests.cc:38-47
  zopt << -5.31028204670497e-14, 5.02854354118183e-13, 0.311688311338095,
      5.35637944798588e-13, 0.311688311339015, -0.0779220779990502, 0.311688311339667,
      0.233766233340057, -0.103896103779874;
  lopt << -5.24675324688535, -4.49350649223710, -3.55844155822323, -0.935064934014372,
      -1.48051948022526, 0.233766233996585;
  vopt << 1.06213597221667e-13, -1.41190425869539e-21, 0, 0, 0, 0, -1.50393600622818e-21,
      -8.75144622575045e-10, 0, 0, 0, 0, -8.75144611157041e-10, -6.56358459377444e-10, 0, 0, 0,
      0;
  for (int i = 0; i < ocp.nz(); i++) EXPECT_NEAR(x.z(i), zopt(i), 1e-8);
  for (int i = 0; i < ocp.nl(); i++) EXPECT_NEAR(x.l(i), lopt(i), 1e-8);
  for (int i = 0; i < ocp.nv(); i++) EXPECT_NEAR(x.v(i), vopt(i), 1e-8);
}

// ---- fbstab_mpc_unit_tests.cc:62-82 (ProblemDataRef) --------------------------
TEST(FBstabMpc, DoubleIntegratorLongHorizon, true) {
  OcpGenerator ocp;
  ocp.DoubleIntegrator(20);
  FBstabMpc::ProblemDataRef data = ocp.GetFBstabInputRef();
  FBstabMpc::Variable x(ocp.ProblemSize());
  FBstabMpc solver(ocp.ProblemSize());
  FBstabMpc::Options opts = FBstabMpc::DefaultOptions();
  opts.abs_tol = 1e-8;
  opts.display_level = Display::OFF;
  solver.UpdateOptions(opts);
  SolverOut out = solver.Solve(data, &x);
  ASSERT_EQ(out.eflag, ExitFlag::SUCCESS);
  ASSERT_LE(out.residual, 1e-6);
}

static void SolveOcp(OcpGenerator& ocp) {
  FBstabMpc::ProblemData data = ocp.GetFBstabInput();
  FBstabMpc::Variable x(ocp.ProblemSize());
  FBstabMpc solver(ocp.ProblemSize());
  FBstabMpc::Options opts = FBstabMpc::DefaultOptions();
  opts.abs_tol = 1e-8;
  opts.display_level = Display::OFF;
  solver.UpdateOptions(opts);
  SolverOut out = solver.Solve(data, &x);
  ASSERT_EQ(out.eflag, ExitFlag::SUCCESS);
  ASSERT_LE(out.residual, 1e-6);
}
// ---- fbstab_mpc_unit_tests.cc:84-104 ------------------------------------------
TEST(FBstabMpc, ServoMotor, true) {
  OcpGenerator ocp;
  ocp.ServoMotor(25);
  SolveOcp(ocp);
}
// ---- fbstab_mpc_unit_tests.cc:106-126 -----------------------------------------
TEST(FBstabMpc, SpacecraftRelativeMotion, true) {
  OcpGenerator ocp;
  ocp.SpacecraftRelativeMotion(40);
  SolveOcp(ocp);
}
// ---- fbstab_mpc_unit_tests.cc:128-148 -----------------------------------------
TEST(FBstabMpc, CopolymerizationReactor, true) {
  OcpGenerator ocp;
  ocp.CopolymerizationReactor(80);
  SolveOcp(ocp);
}

TEST(FBstabMpc, SolveBatchOfStructs, true) {
  OcpGenerator ocp;
  ocp.ServoMotor(10);
  std::vector<FBstabMpc::ProblemData> qps(3, ocp.GetFBstabInput());
  qps[1].x0(0) += 0.01;
  qps[2].x0(1) -= 0.01;
  std::vector<FBstabMpc::Variable> xs(3, FBstabMpc::Variable(ocp.ProblemSize()));
  FBstabMpc solver(ocp.N(), ocp.nx(), ocp.nu(), ocp.nc(), /*max_batch=*/4);
  FBstabMpc::Options opts = FBstabMpc::DefaultOptions();
  opts.display_level = Display::OFF;
  solver.UpdateOptions(opts);
  std::vector<SolverOut> outs = solver.SolveBatch(qps, &xs);
  for (int i = 0; i < 3; i++) {
    ASSERT_EQ(outs[i].eflag, ExitFlag::SUCCESS);
    FBstabMpc::Variable x1(ocp.ProblemSize());
    SolverOut o1 = solver.Solve(qps[i], &x1);
    ASSERT_EQ(o1.newton_iters, outs[i].newton_iters);
    for (int k = 0; k < ocp.nz(); k++) EXPECT_NEAR(xs[i].z(k), x1.z(k), 0.0);
  }
}

// One plant from many initial states: the shared-data and LTI entries return the bytes
// of the wire-format batch, which repeats the stage data per instance.
TEST(FBstabMpc, SharedAndLtiEqualTheWireFormat, true) {
  OcpGenerator ocp;
  ocp.DoubleIntegrator(8);
  const int B = 5, nx = ocp.nx(), nu = ocp.nu(), nc = ocp.nc();
  FBstabMpc::ProblemData one = ocp.GetFBstabInput();
  std::vector<FBstabMpc::ProblemData> qps(B, one);
  std::vector<double> x0(B * nx);
  for (int i = 0; i < B; i++) {
    for (int k = 0; k < nx; k++) {
      qps[i].x0(k) += 0.05 * (i + 1) * (k + 1);
      x0[i * nx + k] = qps[i].x0(k);
    }
  }
  std::vector<FBstabMpc::Variable> xs(B, FBstabMpc::Variable(ocp.ProblemSize()));
  FBstabMpc solver(ocp.N(), nx, nu, nc, /*max_batch=*/B);
  FBstabMpc::Options opts = FBstabMpc::DefaultOptions();
  opts.display_level = Display::OFF;
  solver.UpdateOptions(opts);
  std::vector<SolverOut> ref = solver.SolveBatch(qps, &xs);
  const int nz = ocp.nz(), nl = ocp.nl(), nv = ocp.nv();
  std::vector<double> z(B * nz, 0.0), l(B * nl, 0.0), v(B * nv, 0.0), y(B * nv, 0.0);
  std::vector<SolverOut> got =
      solver.SolveBatchShared(one, B, x0.data(), z.data(), l.data(), v.data(), y.data());
  for (int i = 0; i < B; i++) {
    ASSERT_EQ(got[i].eflag, ref[i].eflag);
    ASSERT_EQ(got[i].newton_iters, ref[i].newton_iters);
    for (int k = 0; k < nz; k++) EXPECT_NEAR(z[i * nz + k], xs[i].z(k), 0.0);
    for (int k = 0; k < nv; k++) EXPECT_NEAR(y[i * nv + k], xs[i].y(k), 0.0);
  }
  // one stage of each matrix: stage 1 carries E (the generator zeroes E(0))
  std::fill(z.begin(), z.end(), 0.0);
  std::fill(l.begin(), l.end(), 0.0);
  std::fill(v.begin(), v.end(), 0.0);
  got = solver.SolveBatchLti(B, one.Q(1).data(), one.R(1).data(), one.S(1).data(),
                             one.q(1).data(), one.r(1).data(), one.A(0).data(),
                             one.B(0).data(), one.c(0).data(), one.E(1).data(),
                             one.L(1).data(), one.d(1).data(), x0.data(), z.data(), l.data(),
                             v.data(), y.data());
  for (int i = 0; i < B; i++) {
    ASSERT_EQ(got[i].newton_iters, ref[i].newton_iters);
    for (int k = 0; k < nz; k++) EXPECT_NEAR(z[i * nz + k], xs[i].z(k), 0.0);
  }
}

// SolveBatch(..., devices): every visible GPU, host buffers, same bytes as one GPU.
TEST(FBstabDense, SolveBatchOnAllDevices, true) {
  const int nz = 6, nl = 2, nv = 9, B = 11;
  std::vector<double> H(B * nz * nz), f(B * nz), G(B * nl * nz), h(B * nl), A(B * nv * nz),
      b(B * nv);
  ASSERT_EQ(fbstab_random_dense_qp(21, 0, B, nz, nl, nv, 0, H.data(), f.data(), G.data(),
                                   h.data(), A.data(), b.data(), 2),
            FBSTAB_OK);
  FBstabDense solver(nz, nl, nv, B);
  std::vector<double> z(B * nz, 0.0), l(B * nl, 0.0), v(B * nv, 0.0), y(B * nv, 0.0);
  std::vector<SolverOut> one = solver.SolveBatch(B, H.data(), f.data(), G.data(), h.data(),
                                                 A.data(), b.data(), z.data(), l.data(),
                                                 v.data(), y.data());
  std::vector<int> devices;
  for (int d = 0; d < fbstab_device_count(); d++) devices.push_back(d);
  std::vector<double> z2(B * nz, 0.0), l2(B * nl, 0.0), v2(B * nv, 0.0), y2(B * nv, 0.0);
  std::vector<SolverOut> all = solver.SolveBatch(B, H.data(), f.data(), G.data(), h.data(),
                                                 A.data(), b.data(), z2.data(), l2.data(),
                                                 v2.data(), y2.data(), devices);
  for (int i = 0; i < B; i++) {
    ASSERT_EQ(all[i].eflag, one[i].eflag);
    ASSERT_EQ(all[i].newton_iters, one[i].newton_iters);
  }
  EXPECT_TRUE(z == z2 && l == l2 && v == v2 && y == y2);
  std::vector<int> dup(2, 0);
  EXPECT_THROW(solver.SolveBatch(B, H.data(), f.data(), G.data(), h.data(), A.data(), b.data(),
                                 z2.data(), l2.data(), v2.data(), y2.data(), dup),
               "duplicate device");
}

// OcpGenerator::GetSimulationInputs + the closed loop on the device: the servo motor is
// steered to its 30 degree target (ocp_generator.cc:286-288) within T = 40 steps, and
// warm starts cost fewer Newton iterations than cold starts.
TEST(ClosedLoop, ServoMotorReachesItsTarget, true) {
  OcpGenerator ocp;
  ocp.ServoMotor(20);
  OcpGenerator::SimulationInputs sim = ocp.GetSimulationInputs();
  ASSERT_EQ(sim.T, 40);
  ASSERT_EQ(sim.C.rows(), 2);
  const int B = 4, nx = ocp.nx();
  std::vector<double> x0(B * nx, 0.0);
  for (int i = 0; i < B; i++) x0[i * nx] = 0.01 * i;  // slightly different starting angles
  FBstabMpc::ProblemData qp = ocp.GetFBstabInput();
  ClosedLoopMpc loop(qp, B, x0.data(), sim.A.data(), sim.B.data(), sim.T);
  ClosedLoopMpc::Trajectory warm = loop.Run(sim.T, true);
  ClosedLoopMpc::Trajectory cold = loop.Run(sim.T, false);
  const double target = 30 * 3.1415926535897 / 180;
  long nw = 0, ncold = 0;
  for (int i = 0; i < B; i++) {
    EXPECT_NEAR(warm.x(i, 0)[0], 0.01 * i, 0.0);
    EXPECT_NEAR(warm.x(i, sim.T)[0], target, 2e-2);
    EXPECT_NEAR(cold.x(i, sim.T)[0], warm.x(i, sim.T)[0], 1e-6);
  }
  for (int t = 1; t < sim.T; t++)
    for (int i = 0; i < B; i++) {
      ASSERT_EQ(warm.out[(size_t)t * B + i].eflag, ExitFlag::SUCCESS);
      nw += warm.out[(size_t)t * B + i].newton_iters;
      ncold += cold.out[(size_t)t * B + i].newton_iters;
    }
  EXPECT_TRUE(nw < ncold);
}

// ---- host-only: types, validation, error behaviour ----------------------------
TEST(Host, OptionsDefaultsAndClamps, false) {
  // DefaultParameters, fbstab_algorithm-impl.h:33-59
  FBstabDense::Options o = FBstabDense::DefaultOptions();
  EXPECT_NEAR(o.sigma0, 1e-8, 0);
  EXPECT_NEAR(o.alpha, 0.95, 0);
  EXPECT_NEAR(o.beta, 0.75, 0);
  EXPECT_NEAR(o.inner_tol_max, 1e-2, 0);
  ASSERT_EQ(o.max_newton_iters, 200);
  ASSERT_EQ(o.max_prox_iters, 30);
  ASSERT_EQ(o.max_inner_iters, 50);
  EXPECT_TRUE(o.check_feasibility && o.nonmonotone_linesearch);
  EXPECT_TRUE(o.display_level == Display::FINAL);
  // ReliableParameters, impl:61-74
  FBstabMpc::Options r = FBstabMpc::ReliableOptions();
  EXPECT_NEAR(r.sigma0, 1e-4, 0);
  EXPECT_NEAR(r.beta, 0.9, 0);
  ASSERT_EQ(r.max_newton_iters, 500);
  EXPECT_TRUE(!r.nonmonotone_linesearch);
  // ValidateOptions clamps, impl:7-31
  o.alpha = 5.0;
  o.max_newton_iters = -3;
  o.sigma0 = 1.0;
  o.ValidateOptions();
  EXPECT_NEAR(o.alpha, 0.999, 0);
  ASSERT_EQ(o.max_newton_iters, 1);
  EXPECT_NEAR(o.sigma0, o.sigma_max, 0);
}

TEST(Host, ConstructorValidation, false) {
  // fbstab_dense.cc:18-27, fbstab_mpc.cc:61-66
  EXPECT_THROW(FBstabDense s(0, 0, 1), "FBstabDense");
  EXPECT_THROW(FBstabDense s(2, -1, 1), "FBstabDense");
  EXPECT_THROW(FBstabDense s(2, 0, 0), "FBstabDense");
  EXPECT_THROW(FBstabMpc s(0, 2, 1, 1), "FBstabMpc");
  EXPECT_THROW(FBstabMpc s(2, 2, 0, 1), "FBstabMpc");
}

TEST(Host, MatrixSequence, false) {
  // tools/matrix_sequence.h:30-51,81-121
  MatrixSequence s(3, 2, 2);
  ASSERT_EQ(s.size(), 12);
  s(1) << 1, 2, 3, 4;  // row by row into column-major storage
  EXPECT_NEAR(s.data()[4 + 0], 1, 0);
  EXPECT_NEAR(s.data()[4 + 1], 3, 0);
  EXPECT_NEAR(s.data()[4 + 2], 2, 0);
  EXPECT_NEAR(s(1)(1, 0), 3, 0);
  EXPECT_THROW(s(3), "Bad indexing");
  EXPECT_THROW(s(-1), "Bad indexing");
  EXPECT_THROW(MatrixSequence t(-1, 2, 2), "Negative length");
  EXPECT_THROW(MatrixSequence t(1, 0, 2), "Non-positive");
  MapMatrixSequence m(s);
  ASSERT_EQ(m.length(), 3);
  EXPECT_NEAR(m(1)(0, 1), 2, 0);
  EXPECT_THROW(MapMatrixSequence t(nullptr, 1, 1, 1), "nullptr");
  EXPECT_THROW(MapMatrixSequence t(s.data(), 0, 1, 1), "Non-positive length");
  MapMatrixSequence e;
  EXPECT_THROW(e(0), "Bad indexing");
}

TEST(Host, OcpGeneratorShapes, false) {
  OcpGenerator ocp;
  EXPECT_THROW(ocp.GetFBstabInput(), "problem creator");
  EXPECT_THROW(ocp.ServoMotor(0), "N <= 0");
  ocp.CopolymerizationReactor(5);
  ASSERT_EQ(ocp.nx(), 18);
  ASSERT_EQ(ocp.nu(), 5);
  ASSERT_EQ(ocp.nc(), 10);
  ASSERT_EQ(ocp.nz(), 6 * 23);
  FBstabMpc::ProblemData d = ocp.GetFBstabInput();
  ASSERT_EQ(d.A.length(), 5);
  ASSERT_EQ(d.E.length(), 6);
  // state constraints are dropped at stage 0: E(0) = 0, ocp_generator.cc:403-405
  ocp.ServoMotor(4);
  FBstabMpc::ProblemData s = ocp.GetFBstabInput();
  EXPECT_NEAR(s.E(0).norm(), 0, 0);
  EXPECT_TRUE(s.E(1).norm() > 0);
  EXPECT_TRUE(s.L(0).norm() > 0);
}

// With a GPU these exercise Solve's size validation; without one the
// constructor itself must fail loudly (no CPU fallback).
TEST(Host, NoSilentCpuFallback, false) {
  if (fbstab_device_count() > 0) return;
  EXPECT_THROW(FBstabDense s(2, 0, 2), "no CUDA device");
  EXPECT_THROW(FBstabMpc s(2, 2, 1, 1), "no CUDA device");
}

TEST(Host, SimulationInputs, false) {
  OcpGenerator ocp;
  EXPECT_THROW(ocp.GetSimulationInputs(), "problem creator");
  ocp.DoubleIntegrator(4);
  OcpGenerator::SimulationInputs s = ocp.GetSimulationInputs();
  ASSERT_EQ(s.T, 40);
  ASSERT_EQ(s.A.rows(), 2);
  ASSERT_EQ(s.B.cols(), 1);
  EXPECT_NEAR(s.A(0, 1), 1.0, 0.0);  // [1 1; 0 1], ocp_generator.cc:325-326
  EXPECT_NEAR(s.A(1, 0), 0.0, 0.0);
  EXPECT_NEAR(s.B(1, 0), 1.0, 0.0);
  EXPECT_NEAR(s.C(1, 1), 1.0, 0.0);
  EXPECT_NEAR(s.D.norm(), 0.0, 0.0);
  ocp.CopolymerizationReactor(3);
  s = ocp.GetSimulationInputs();
  ASSERT_EQ(s.T, 200);
  ASSERT_EQ(s.C.rows(), 4);
  ASSERT_EQ(s.C.cols(), 18);
  EXPECT_NEAR(s.C(3, 17), 1.8214, 0.0);
  ocp.SpacecraftRelativeMotion(3);
  s = ocp.GetSimulationInputs();
  ASSERT_EQ(s.T, 100);
  ASSERT_EQ(s.C.rows(), 6);
}

TEST(FBstabDense, SizeMismatchThrows, true) {
  FBstabDense solver(2, 0, 2);
  FBstabDense::ProblemData bad(3, 0, 2);
  FBstabDense::Variable x(2, 0, 2);
  EXPECT_THROW(solver.Solve(bad, &x), "mismatch between *this and data");
  {  // a short y is an error, not a heap overflow (the engine writes nv doubles into it)
    FBstabDense::ProblemData ok(2, 0, 2);
    FBstabDense::Variable xy(2, 0, 2);
    xy.y = VectorXd(1);
    EXPECT_THROW(solver.Solve(ok, &xy), "initial guess");
  }
  FBstabDense::ProblemData good(2, 0, 2);
  FBstabDense::Variable xb(2, 1, 2);
  EXPECT_THROW(solver.Solve(good, &xb), "initial guess");
  FBstabDense::ProblemData skew(2, 0, 2);
  skew.A = MatrixXd(2, 3);
  EXPECT_THROW(solver.Solve(skew, &x), "Az <= b");
}

TEST(FBstabMpc, SizeMismatchThrows, true) {
  OcpGenerator ocp;
  ocp.DoubleIntegrator(3);
  FBstabMpc solver(4, 2, 1, 6);
  FBstabMpc::ProblemData data = ocp.GetFBstabInput();
  FBstabMpc::Variable x(4, 2, 1, 6);
  EXPECT_THROW(solver.Solve(data, &x), "mismatch between *this and data");
  FBstabMpc solver3(3, 2, 1, 6);
  EXPECT_THROW(solver3.Solve(data, &x), "initial guess");
  data.A = MatrixSequence(2, 2, 2);
  FBstabMpc::Variable x3(3, 2, 1, 6);
  EXPECT_THROW(solver3.Solve(data, &x3), "Sequence length mismatch");
}

// The dense solver tests FeasibleQPwithEQ / InfeasibleQP (fbstab_dense_unit_tests.cc:75-104,
// 195-217) through the sparse solver: same data in compressed-column form, same assertions.
static SparsePattern CscOf(const MatrixXd& M, bool upper, VectorXd* vals) {
  SparsePattern s;
  s.rows = (int)M.rows();
  s.cols = (int)M.cols();
  s.p.push_back(0);
  std::vector<double> x;
  for (int c = 0; c < s.cols; c++) {
    for (int r = 0; r < s.rows; r++)
      if (M(r, c) != 0.0 && (!upper || r <= c)) {
        s.i.push_back(r);
        x.push_back(M(r, c));
      }
    s.p.push_back((int)s.i.size());
  }
  *vals = VectorXd((int)x.size());
  for (size_t k = 0; k < x.size(); k++) (*vals)(k) = x[k];
  return s;
}

TEST(FBstabSparse, FeasibleQPwithEQ, true) {
  MatrixXd H(2, 2), G(1, 2), A(2, 2);
  H << 4, 1, 1, 2;
  G << 1, 1;
  A << -1, 0, 0, -1;
  FBstabSparse::ProblemData qp;
  const SparsePattern pH = CscOf(H, true, &qp.Hx), pG = CscOf(G, false, &qp.Gx),
                      pA = CscOf(A, false, &qp.Ax);
  qp.f = VectorXd(2);
  qp.f << 1, 1;
  qp.h = VectorXd(1);
  qp.h << 1;
  qp.b = VectorXd(2);
  qp.b << 0, 0;
  FBstabSparse solver(pH, pG, pA);
  FBstabSparse::Options opts = FBstabSparse::DefaultOptions();
  opts.abs_tol = 1e-8;
  opts.display_level = Display::OFF;
  solver.UpdateOptions(opts);
  FBstabSparse::Variable x(2, 1, 2);
  SolverOut out = solver.Solve(qp, &x);
  ASSERT_EQ(out.eflag, ExitFlag::SUCCESS);
  EXPECT_NEAR(x.z(0), 0.25, 1e-8);
  EXPECT_NEAR(x.z(1), 0.75, 1e-8);
  EXPECT_TRUE(solver.FactorNonzeros() > 0);
  FBstabSparse::Variable bad(2, 0, 2);
  EXPECT_THROW(solver.Solve(qp, &bad), "initial guess");
}

// SolveBatch(..., devices) of the sparse facade: every visible GPU, host buffers, the same
// bytes as the single-handle batch solve.
TEST(FBstabSparse, SolveBatchOnAllDevices, true) {
  MatrixXd H(2, 2), G(1, 2), A(2, 2);
  H << 4, 1, 1, 2;
  G << 1, 1;
  A << -1, 0, 0, -1;
  VectorXd Hx0, Gx0, Ax0;
  const SparsePattern pH = CscOf(H, true, &Hx0), pG = CscOf(G, false, &Gx0),
                      pA = CscOf(A, false, &Ax0);
  const int B = 7, nH = (int)Hx0.size(), nG = (int)Gx0.size(), nA = (int)Ax0.size();
  std::vector<double> Hx(B * nH), Gx(B * nG), Ax(B * nA), f(B * 2), h(B), b(B * 2, 0.0);
  for (int i = 0; i < B; i++) {
    for (int k = 0; k < nH; k++) Hx[i * nH + k] = Hx0(k) * (1.0 + 0.1 * i);
    for (int k = 0; k < nG; k++) Gx[i * nG + k] = Gx0(k);
    for (int k = 0; k < nA; k++) Ax[i * nA + k] = Ax0(k);
    f[2 * i] = 1.0 - 0.2 * i;
    f[2 * i + 1] = 1.0 + 0.3 * i;
    h[i] = 1.0 + 0.05 * i;
  }
  FBstabSparse solver(pH, pG, pA, B);
  std::vector<double> z(B * 2, 0.0), l(B, 0.0), v(B * 2, 0.0), y(B * 2, 0.0);
  std::vector<SolverOut> one = solver.SolveBatch(B, Hx.data(), f.data(), Gx.data(), h.data(),
                                                 Ax.data(), b.data(), z.data(), l.data(),
                                                 v.data(), y.data());
  std::vector<int> devices;
  for (int d = 0; d < fbstab_device_count(); d++) devices.push_back(d);
  std::vector<double> z2(B * 2, 0.0), l2(B, 0.0), v2(B * 2, 0.0), y2(B * 2, 0.0);
  std::vector<SolverOut> all = solver.SolveBatch(B, Hx.data(), f.data(), Gx.data(), h.data(),
                                                 Ax.data(), b.data(), z2.data(), l2.data(),
                                                 v2.data(), y2.data(), devices);
  for (int i = 0; i < B; i++) {
    ASSERT_EQ(one[i].eflag, ExitFlag::SUCCESS);
    ASSERT_EQ(all[i].eflag, one[i].eflag);
    ASSERT_EQ(all[i].newton_iters, one[i].newton_iters);
  }
  EXPECT_TRUE(z == z2 && l == l2 && v == v2 && y == y2);
}

TEST(FBstabSparse, InfeasibleQP, true) {
  MatrixXd H(2, 2), G(0, 2), A(5, 2);
  H << 1, 0, 0, 0;
  A << 1, 1, 1, 0, 0, 1, -1, 0, 0, -1;
  FBstabSparse::ProblemData qp;
  const SparsePattern pH = CscOf(H, true, &qp.Hx), pG = CscOf(G, false, &qp.Gx),
                      pA = CscOf(A, false, &qp.Ax);
  qp.f = VectorXd(2);
  qp.f << 1, -1;
  qp.h = VectorXd(0);
  qp.b = VectorXd(5);
  qp.b << 0, 3, 3, -1, -1;
  FBstabSparse solver(pH, pG, pA);
  FBstabSparse::Options opts = FBstabSparse::DefaultOptions();
  opts.abs_tol = 1e-8;
  opts.display_level = Display::OFF;
  solver.UpdateOptions(opts);
  FBstabSparse::Variable x(2, 0, 5);
  SolverOut out = solver.Solve(qp, &x);
  ASSERT_EQ(out.eflag, ExitFlag::PRIMAL_INFEASIBLE);
}

TEST(FBstabSparse, BadPatternThrows, false) {
  SparsePattern H, G, A;
  H.rows = 2;
  H.cols = 3;  // not square
  H.p = {0, 0, 0, 0};
  A.rows = 1;
  A.cols = 3;
  A.p = {0, 0, 0, 0};
  EXPECT_THROW(FBstabSparse(H, G, A), "H must be square");
}

}  // namespace test
}  // namespace fbstab

int main(int argc, char** argv) {
  const bool host_only = argc > 1 && std::strcmp(argv[1], "--host") == 0;
  int ran = 0, failed = 0;
  for (const Case& c : Cases()) {
    if (host_only && c.needs_gpu) continue;
    const int before = g_failures;
    printf("[ RUN  ] %s\n", c.name);
    try {
      c.fn();
    } catch (const std::exception& e) {
      printf("    unexpected exception: %s\n", e.what());
      g_failures++;
    }
    ran++;
    if (g_failures != before) {
      failed++;
      printf("[ FAIL ] %s\n", c.name);
    } else {
      printf("[  OK  ] %s\n", c.name);
    }
  }
  printf("%d cases, %d failed\n", ran, failed);
  return failed ? 1 : 0;
}

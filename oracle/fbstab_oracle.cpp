// fbstab_oracle.cpp -- CPU oracle: a sequential restatement of the reference
// FBstab algorithm (dliaomcp/fbstab) in plain C++17 with no Eigen.
//
// TEST INFRASTRUCTURE ONLY (see fbstab_oracle.h).  It exists to check the CUDA
// engine; it is never linked into, imported by or called from fbstab_b200/.
//
// PINNING.  The real reference cannot be built in this environment (it needs
// Eigen 3.4.0, pinned at reference tools/eigen/repository.bzl:8-11, which is
// neither vendored nor installed, and there is no network).  The oracle is
// therefore pinned against every golden vector the reference's own tests hold
// for this path (tests/test_oracle_goldens.py lists them with file:line).
// Those tests pin exit flags, solutions and component values; they do NOT pin
// iteration counts.  The *trajectory* (newton / prox counts) is pinned by the
// reference's OWN algorithm sources, compiled where they lie under
// /root/reference against a stand-in for the Eigen API they use
// (oracle/_ref/libfbstab_ref.so: `make -C oracle _ref`, ref_wrapper.cpp,
// eigen_shim/Eigen/Dense): tests/test_reference_sources.py holds this restatement
// to the same exit flag and iteration counts on every instance of every benchmark
// family, and to the same BYTES on the MPC path.  What remains unpinned is only
// Eigen's own rounding (its blocked, packetised sums): two correctly rounded
// evaluations of the same formulas, like this file built with and without FMA
// contraction (tests/golden/trajectory_floor.json).
//
// Eigen arithmetic that is not in /root/reference is restated from its
// published algorithm (Eigen 3.4.0): LDLT = in-place unblocked LDL' of the
// lower triangle with symmetric diagonal pivoting on the largest |a_ii| of the
// not-yet-updated trailing diagonal; LLT = unblocked lower Cholesky (sizes
// < 32); triangular solves column-oriented for a lower factor and
// dot-product oriented for its transpose.  Summation order inside products is
// plain left-to-right (Eigen's packetised order is not reproducible by hand).
//
// Every function cites the reference file:line it follows.

#include "fbstab_oracle.h"

#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstring>
#include <limits>
#include <memory>
#include <thread>
#include <vector>

namespace {

using Vec = std::vector<double>;

struct SaturateError {};  // tools/utilities.h:21-25 throws runtime_error

// tools/utilities.h:19-28
double saturate(double x, double a, double b) {
  if (a > b) throw SaturateError{};
  return std::max(std::min(x, b), a);
}

double dot(const double* a, const double* b, int n) {
  double s = 0.0;
  for (int i = 0; i < n; i++) s += a[i] * b[i];
  return s;
}
double norm2(const double* a, int n) { return std::sqrt(dot(a, a, n)); }
double norminf(const double* a, int n) {
  double s = 0.0;
  for (int i = 0; i < n; i++) s = std::max(s, std::fabs(a[i]));
  return s;
}

// Grow-only per-thread scratch: the temporaries of the hot loops (Eigen evaluates the
// same products into preallocated or stack temporaries) cost no heap traffic, so that
// the port is not a pessimistic CPU baseline.  Slots are not shared between nested
// callers: 0 gemv_n_acc, 1 LDLT column update, 2 variant-2 row solve, 3 LLT column
// update, 4-6 MpcData stage temporaries.  Same arithmetic as with local vectors.
static double* Scratch(int slot, size_t n) {
  thread_local std::vector<double> buf[8];
  if (buf[slot].size() < n) buf[slot].resize(n);
  return buf[slot].data();
}

// y(rows) += a * M(rows x cols, column-major, leading dim ld) * x(cols)
void gemv_n_acc(const double* M, int rows, int cols, int ld, const double* x,
                double a, double* y) {
  if (rows <= 0 || cols <= 0) return;
  double* t = Scratch(0, rows);
  for (int i = 0; i < rows; i++) t[i] = 0.0;
  for (int j = 0; j < cols; j++) {
    const double xj = x[j];
    const double* c = M + (size_t)j * ld;
    for (int i = 0; i < rows; i++) t[i] += c[i] * xj;
  }
  for (int i = 0; i < rows; i++) y[i] += a * t[i];
}
// y(cols) += a * M' * x(rows)
void gemv_t_acc(const double* M, int rows, int cols, int ld, const double* x,
                double a, double* y) {
  if (rows <= 0 || cols <= 0) return;
  for (int j = 0; j < cols; j++) y[j] += a * dot(M + (size_t)j * ld, x, rows);
}

void scale_by_b(double b, double* y, int n) {
  if (b == 0.0) {
    for (int i = 0; i < n; i++) y[i] = 0.0;
  } else if (b != 1.0) {
    for (int i = 0; i < n; i++) y[i] *= b;
  }
}

// ---------------------------------------------------------------------------
// Data concept: reference fbstab/components/abstract_components.h:24-62.
// ---------------------------------------------------------------------------
struct Data {
  int nz = 0, nl = 0, nv = 0;
  double forcing_norm = 0.0;
  virtual ~Data() {}
  virtual void gemvH(const double* x, double a, double b, double* y) const = 0;
  virtual void gemvA(const double* x, double a, double b, double* y) const = 0;
  virtual void gemvAT(const double* x, double a, double b, double* y) const = 0;
  virtual void gemvG(const double* x, double a, double b, double* y) const = 0;
  virtual void gemvGT(const double* x, double a, double b, double* y) const = 0;
  virtual void axpyf(double a, double* y) const = 0;
  virtual void axpyh(double a, double* y) const = 0;
  virtual void axpyb(double a, double* y) const = 0;
};

// fbstab/components/dense_data.h:44-74, dense_data.cc:12-41
struct DenseData : Data {
  const double *H, *f, *G, *h, *A, *b;
  DenseData(int nz_, int nl_, int nv_, const double* H_, const double* f_,
            const double* G_, const double* h_, const double* A_,
            const double* b_)
      : H(H_), f(f_), G(G_), h(h_), A(A_), b(b_) {
    nz = nz_;
    nl = nl_;
    nv = nv_;
    // dense_data.h:72-73
    forcing_norm = std::sqrt(dot(b, b, nv) + dot(f, f, nz) + dot(h, h, nl));
  }
  // y = a*M*x + b*y : the product is evaluated first, then added to b*y.
  void gemvH(const double* x, double a, double bb, double* y) const override {
    scale_by_b(bb, y, nz);
    gemv_n_acc(H, nz, nz, nz, x, a, y);
  }
  void gemvA(const double* x, double a, double bb, double* y) const override {
    scale_by_b(bb, y, nv);
    gemv_n_acc(A, nv, nz, nv, x, a, y);
  }
  void gemvAT(const double* x, double a, double bb, double* y) const override {
    scale_by_b(bb, y, nz);
    gemv_t_acc(A, nv, nz, nv, x, a, y);
  }
  void gemvG(const double* x, double a, double bb, double* y) const override {
    scale_by_b(bb, y, nl);
    gemv_n_acc(G, nl, nz, nl, x, a, y);
  }
  void gemvGT(const double* x, double a, double bb, double* y) const override {
    scale_by_b(bb, y, nz);
    gemv_t_acc(G, nl, nz, nl, x, a, y);
  }
  void axpyf(double a, double* y) const override {
    for (int i = 0; i < nz; i++) y[i] += a * f[i];
  }
  void axpyh(double a, double* y) const override {
    for (int i = 0; i < nl; i++) y[i] += a * h[i];
  }
  void axpyb(double a, double* y) const override {
    for (int i = 0; i < nv; i++) y[i] += a * b[i];
  }
};

// fbstab/components/mpc_data.h:62-98, mpc_data.cc:17-289.
// Sequences: Q,R,S,q,r,E,L,d have N+1 entries; A,B,c have N entries.
struct MpcData : Data {
  int N, nx, nu, nc;
  const double *Q, *R, *S, *q, *r, *A, *B, *c, *E, *L, *d, *x0;
  const double* Qi(int i) const { return Q + (size_t)i * nx * nx; }
  const double* Ri(int i) const { return R + (size_t)i * nu * nu; }
  const double* Si(int i) const { return S + (size_t)i * nu * nx; }
  const double* qi(int i) const { return q + (size_t)i * nx; }
  const double* ri(int i) const { return r + (size_t)i * nu; }
  const double* Ai(int i) const { return A + (size_t)i * nx * nx; }
  const double* Bi(int i) const { return B + (size_t)i * nx * nu; }
  const double* ci(int i) const { return c + (size_t)i * nx; }
  const double* Ei(int i) const { return E + (size_t)i * nc * nx; }
  const double* Li(int i) const { return L + (size_t)i * nc * nu; }
  const double* di(int i) const { return d + (size_t)i * nc; }

  MpcData(int N_, int nx_, int nu_, int nc_, const double* Q_, const double* R_,
          const double* S_, const double* q_, const double* r_,
          const double* A_, const double* B_, const double* c_,
          const double* E_, const double* L_, const double* d_,
          const double* x0_)
      : N(N_), nx(nx_), nu(nu_), nc(nc_), Q(Q_), R(R_), S(S_), q(q_), r(r_),
        A(A_), B(B_), c(c_), E(E_), L(L_), d(d_), x0(x0_) {
    nz = (N + 1) * (nx + nu);
    nl = (N + 1) * nx;
    nv = (N + 1) * nc;
    // mpc_data.h:89-97
    double s = 0.0;
    for (int i = 0; i < N + 1; i++) {
      s += dot(qi(i), qi(i), nx);
      s += dot(ri(i), ri(i), nu);
      s += dot(di(i), di(i), nc);
      s += (i == 0 ? dot(x0, x0, nx) : dot(ci(i - 1), ci(i - 1), nx));
    }
    forcing_norm = std::sqrt(s);
  }

  // mpc_data.cc:17-64.  [yx;yu] += a*[Q S';S R][vx;vu] per stage.
  void gemvH(const double* x, double a, double bb, double* y) const override {
    scale_by_b(bb, y, nz);
    const int ns = nx + nu;
    double *tx = Scratch(4, nx), *tu = Scratch(5, nu);
    for (int i = 0; i < N + 1; i++) {
      const double* vx = x + (size_t)i * ns;
      const double* vu = vx + nx;
      double* yx = y + (size_t)i * ns;
      double* yu = yx + nx;
      std::fill(tx, tx + nx, 0.0);
      std::fill(tu, tu + nu, 0.0);
      gemv_n_acc(Qi(i), nx, nx, nx, vx, 1.0, tx);
      gemv_t_acc(Si(i), nu, nx, nu, vu, 1.0, tx);
      gemv_n_acc(Si(i), nu, nx, nu, vx, 1.0, tu);
      gemv_n_acc(Ri(i), nu, nu, nu, vu, 1.0, tu);
      for (int k = 0; k < nx; k++) yx[k] += a * tx[k];
      for (int k = 0; k < nu; k++) yu[k] += a * tu[k];
    }
  }
  // mpc_data.cc:66-105.  y(i) += a*(E(i)x(i) + L(i)u(i))
  void gemvA(const double* x, double a, double bb, double* y) const override {
    scale_by_b(bb, y, nv);
    const int ns = nx + nu;
    Vec t(nc);
    for (int i = 0; i < N + 1; i++) {
      const double* xi = x + (size_t)i * ns;
      const double* ui = xi + nx;
      std::fill(t.begin(), t.end(), 0.0);
      gemv_n_acc(Ei(i), nc, nx, nc, xi, 1.0, t.data());
      gemv_n_acc(Li(i), nc, nu, nc, ui, 1.0, t.data());
      double* yi = y + (size_t)i * nc;
      for (int k = 0; k < nc; k++) yi[k] += a * t[k];
    }
  }
  // mpc_data.cc:107-151.  y(0) += -a x(0); y(i) += a(A x(i-1)+B u(i-1) - x(i))
  void gemvG(const double* x, double a, double bb, double* y) const override {
    scale_by_b(bb, y, nl);
    const int ns = nx + nu;
    for (int k = 0; k < nx; k++) y[k] += -a * x[k];
    Vec t(nx);
    for (int i = 1; i < N + 1; i++) {
      const double* xm1 = x + (size_t)(i - 1) * ns;
      const double* um1 = xm1 + nx;
      const double* xi = x + (size_t)i * ns;
      double* yi = y + (size_t)i * nx;
      std::fill(t.begin(), t.end(), 0.0);
      gemv_n_acc(Ai(i - 1), nx, nx, nx, xm1, 1.0, t.data());
      gemv_n_acc(Bi(i - 1), nx, nu, nx, um1, 1.0, t.data());
      for (int k = 0; k < nx; k++) yi[k] += a * t[k];
      for (int k = 0; k < nx; k++) yi[k] -= a * xi[k];
    }
  }
  // mpc_data.cc:153-199.  (The reference drops the B' term when a is not +-1,
  // mpc_data.cc:192-194; every caller on the solve path passes a = 1, so the
  // oracle implements the mathematically intended operation.)
  void gemvGT(const double* x, double a, double bb, double* y) const override {
    scale_by_b(bb, y, nz);
    const int ns = nx + nu;
    for (int i = 0; i < N; i++) {
      const double* vi = x + (size_t)i * nx;
      const double* vp1 = x + (size_t)(i + 1) * nx;
      double* xi = y + (size_t)i * ns;
      double* ui = xi + nx;
      for (int k = 0; k < nx; k++) xi[k] += -a * vi[k];
      gemv_t_acc(Ai(i), nx, nx, nx, vp1, a, xi);
      gemv_t_acc(Bi(i), nx, nu, nx, vp1, a, ui);
    }
    double* xN = y + (size_t)N * ns;
    const double* vN = x + (size_t)N * nx;
    for (int k = 0; k < nx; k++) xN[k] += -a * vN[k];
  }
  // mpc_data.cc:201-240.  x(i) += a E(i)' v(i); u(i) += a L(i)' v(i)
  void gemvAT(const double* x, double a, double bb, double* y) const override {
    scale_by_b(bb, y, nz);
    const int ns = nx + nu;
    for (int i = 0; i < N + 1; i++) {
      const double* vi = x + (size_t)i * nc;
      double* xi = y + (size_t)i * ns;
      double* ui = xi + nx;
      gemv_t_acc(Ei(i), nc, nx, nc, vi, a, xi);
      gemv_t_acc(Li(i), nc, nu, nc, vi, a, ui);
    }
  }
  // mpc_data.cc:242-258   f = [q(i); r(i)]
  void axpyf(double a, double* y) const override {
    const int ns = nx + nu;
    for (int i = 0; i < N + 1; i++) {
      double* xi = y + (size_t)i * ns;
      for (int k = 0; k < nx; k++) xi[k] += a * qi(i)[k];
      for (int k = 0; k < nu; k++) xi[nx + k] += a * ri(i)[k];
    }
  }
  // mpc_data.cc:260-274   h = -[x0; c(0); ...; c(N-1)]
  void axpyh(double a, double* y) const override {
    for (int k = 0; k < nx; k++) y[k] += -a * x0[k];
    for (int i = 1; i < N + 1; i++)
      for (int k = 0; k < nx; k++) y[(size_t)i * nx + k] += -a * ci(i - 1)[k];
  }
  // mpc_data.cc:276-289   b = -d
  void axpyb(double a, double* y) const override {
    for (int i = 0; i < N + 1; i++)
      for (int k = 0; k < nc; k++) y[(size_t)i * nc + k] += -a * di(i)[k];
  }
};

// ---------------------------------------------------------------------------
// FullVariable: fbstab/components/full_variable.cc:40-83
// ---------------------------------------------------------------------------
struct Variable {
  const Data* data;
  Vec z, l, v, y;
  explicit Variable(const Data* d)
      : data(d), z(d->nz, 0.0), l(d->nl, 0.0), v(d->nv, 0.0), y(d->nv, 0.0) {}
  // full_variable.cc:47-53
  void InitializeConstraintMargin() {
    std::fill(y.begin(), y.end(), 0.0);
    data->axpyb(1.0, y.data());
    data->gemvA(z.data(), -1.0, 1.0, y.data());
  }
  // full_variable.cc:40-45
  void Fill(double a) {
    std::fill(z.begin(), z.end(), a);
    std::fill(l.begin(), l.end(), a);
    std::fill(v.begin(), v.end(), a);
    InitializeConstraintMargin();
  }
  // full_variable.cc:55-65 (y-aware)
  void axpy(double a, const Variable& x) {
    for (size_t i = 0; i < z.size(); i++) z[i] += a * x.z[i];
    for (size_t i = 0; i < l.size(); i++) l[i] += a * x.l[i];
    for (size_t i = 0; i < v.size(); i++) v[i] += a * x.v[i];
    for (size_t i = 0; i < y.size(); i++) y[i] += a * x.y[i];
    data->axpyb(-a, y.data());
  }
  void Copy(const Variable& x) {
    z = x.z;
    l = x.l;
    v = x.v;
    y = x.y;
  }
  // full_variable.cc:75
  void ProjectDuals() {
    for (auto& e : v) e = std::max(e, 0.0);
  }
  // full_variable.cc:77-83
  double Norm() const {
    const double t1 = norm2(z.data(), (int)z.size());
    const double t2 = norm2(l.data(), (int)l.size());
    const double t3 = norm2(v.data(), (int)v.size());
    return std::sqrt(t1 * t1 + t2 * t2 + t3 * t3);
  }
};

// ---------------------------------------------------------------------------
// FullResidual: fbstab/components/full_residual.cc:28-118
// ---------------------------------------------------------------------------
// full_residual.cc:115-118
double pfb(double a, double b, double alpha) {
  const double fb = a + b - std::sqrt(a * a + b * b);
  return alpha * fb + (1.0 - alpha) * std::max(0.0, a) * std::max(0.0, b);
}

struct Residual {
  const Data* data;
  double alpha = 0.95;
  Vec z, l, v;
  double znorm = 0.0, lnorm = 0.0, vnorm = 0.0;
  int* eval_counter = nullptr;
  explicit Residual(const Data* d)
      : data(d), z(d->nz, 0.0), l(d->nl, 0.0), v(d->nv, 0.0) {}
  void Fill(double a) {
    std::fill(z.begin(), z.end(), a);
    std::fill(l.begin(), l.end(), a);
    std::fill(v.begin(), v.end(), a);
  }
  void Negate() {
    for (auto& e : z) e *= -1;
    for (auto& e : l) e *= -1;
    for (auto& e : v) e *= -1;
  }
  // full_residual.cc:40-47 (norms are the cached ones)
  double Norm() const {
    return std::sqrt(znorm * znorm + lnorm * lnorm + vnorm * vnorm);
  }
  double Merit() const {
    const double t = Norm();
    return 0.5 * t * t;
  }
  void UpdateNorms() {
    znorm = norm2(z.data(), (int)z.size());
    lnorm = norm2(l.data(), (int)l.size());
    vnorm = norm2(v.data(), (int)v.size());
  }
  void Count() {
    if (eval_counter) (*eval_counter)++;
  }
  // full_residual.cc:49-74
  void InnerResidual(const Variable& x, const Variable& xbar, double sigma) {
    Count();
    std::fill(z.begin(), z.end(), 0.0);
    data->axpyf(1.0, z.data());
    data->gemvH(x.z.data(), 1.0, 1.0, z.data());
    data->gemvGT(x.l.data(), 1.0, 1.0, z.data());
    data->gemvAT(x.v.data(), 1.0, 1.0, z.data());
    for (size_t i = 0; i < z.size(); i++) z[i] += sigma * (x.z[i] - xbar.z[i]);

    std::fill(l.begin(), l.end(), 0.0);
    data->axpyh(1.0, l.data());
    data->gemvG(x.z.data(), -1.0, 1.0, l.data());
    for (size_t i = 0; i < l.size(); i++) l[i] += sigma * (x.l[i] - xbar.l[i]);

    for (size_t i = 0; i < v.size(); i++) {
      const double ys = x.y[i] + sigma * (x.v[i] - xbar.v[i]);
      v[i] = pfb(ys, x.v[i], alpha);
    }
    UpdateNorms();
  }
  // full_residual.cc:76-97
  void NaturalResidual(const Variable& x) {
    Count();
    std::fill(z.begin(), z.end(), 0.0);
    data->axpyf(1.0, z.data());
    data->gemvH(x.z.data(), 1.0, 1.0, z.data());
    data->gemvGT(x.l.data(), 1.0, 1.0, z.data());
    data->gemvAT(x.v.data(), 1.0, 1.0, z.data());

    std::fill(l.begin(), l.end(), 0.0);
    data->axpyh(1.0, l.data());
    data->gemvG(x.z.data(), -1.0, 1.0, l.data());

    for (size_t i = 0; i < v.size(); i++) v[i] = std::min(x.y[i], x.v[i]);
    UpdateNorms();
  }
  // full_residual.cc:99-109
  void PenalizedNaturalResidual(const Variable& x) {
    NaturalResidual(x);
    for (size_t i = 0; i < v.size(); i++) {
      v[i] = alpha * v[i] +
             (1 - alpha) * std::max(0.0, x.y[i]) * std::max(0.0, x.v[i]);
    }
    UpdateNorms();
  }
};

// ---------------------------------------------------------------------------
// FullFeasibility: fbstab/components/full_feasibility.cc:25-88
// returns 0 FEASIBLE, 1 PRIMAL_INFEASIBLE, 2 DUAL_INFEASIBLE, 3 BOTH
// ---------------------------------------------------------------------------
int CheckFeasibility(const Data* data, const double* xz, const double* xl,
                     const double* xv, double tol) {
  const int nz = data->nz, nl = data->nl, nv = data->nv;
  Vec tz(nz, 0.0), tl(nl, 0.0), tv(nv, 0.0);

  data->gemvA(xz, 1.0, 0.0, tv.data());
  double d1 = tv[0];
  for (int i = 1; i < nv; i++) d1 = std::max(d1, tv[i]);

  data->gemvG(xz, 1.0, 0.0, tl.data());
  const double d2 = norminf(tl.data(), nl);

  data->gemvH(xz, 1.0, 0.0, tz.data());
  const double d3 = norminf(tz.data(), nz);

  std::fill(tz.begin(), tz.end(), 0.0);
  data->axpyf(1.0, tz.data());
  const double d4 = dot(tz.data(), xz, nz);

  const double w = norminf(xz, nz);
  bool dual_feasible = true;
  if ((d1 <= w * tol) && (d2 <= tol * w) && (d3 <= tol * w) && (d4 < 0) &&
      (w > 1e-14)) {
    dual_feasible = false;
  }

  std::fill(tz.begin(), tz.end(), 0.0);
  data->gemvAT(xv, 1.0, 1.0, tz.data());
  data->gemvGT(xl, 1.0, 1.0, tz.data());
  const double p1 = norminf(tz.data(), nz);

  std::fill(tv.begin(), tv.end(), 0.0);
  data->axpyb(1.0, tv.data());
  std::fill(tl.begin(), tl.end(), 0.0);
  data->axpyh(1.0, tl.data());
  const double p2 = dot(tl.data(), xl, nl) + dot(tv.data(), xv, nv);

  const double u = std::max(norminf(xv, nv), norminf(xl, nl));
  bool primal_feasible = true;
  if ((p1 <= tol * u) && (p2 < 0)) primal_feasible = false;

  if (primal_feasible && dual_feasible) return 0;
  if (primal_feasible && !dual_feasible) return 2;
  if (!primal_feasible && dual_feasible) return 1;
  return 3;
}

// ---------------------------------------------------------------------------
// PFB gradient: dense_cholesky_solver.cc:129-148 (identical copy at
// riccati_linear_solver.cc:346-365).  zero_tolerance_ = 1e-13
// (dense_cholesky_solver.h / riccati_linear_solver.h).
// ---------------------------------------------------------------------------
void PFBGradient(double a, double b, double alpha, double* ga, double* gb) {
  const double r = std::sqrt(a * a + b * b);
  const double d = 1.0 / std::sqrt(2.0);
  if (r < 1e-13) {
    *ga = alpha * (1.0 - d);
    *gb = alpha * (1.0 - d);
  } else if ((a > 0) && (b > 0)) {
    *ga = alpha * (1.0 - a / r) + (1.0 - alpha) * b;
    *gb = alpha * (1.0 - b / r) + (1.0 - alpha) * a;
  } else {
    *ga = alpha * (1.0 - a / r);
    *gb = alpha * (1.0 - b / r);
  }
}

struct LinearSolver {
  double alpha = 0.95;
  Vec gamma, mus, Gamma;
  virtual ~LinearSolver() {}
  virtual bool Initialize(const Variable& x, const Variable& xbar,
                          double sigma) = 0;
  virtual bool Solve(const Residual& r, Variable* dx) = 0;
  // dense_cholesky_solver.cc:53-60 / riccati_linear_solver.cc:91-99
  void Barrier(const Variable& x, const Variable& xbar, double sigma) {
    const int nv = (int)x.v.size();
    gamma.resize(nv);
    mus.resize(nv);
    Gamma.resize(nv);
    for (int i = 0; i < nv; i++) {
      const double ys = x.y[i] + sigma * (x.v[i] - xbar.v[i]);
      double ga, gb;
      PFBGradient(ys, x.v[i], alpha, &ga, &gb);
      gamma[i] = ga;
      mus[i] = gb + sigma * ga;
      Gamma[i] = gamma[i] / mus[i];
    }
  }
};

// ---------------------------------------------------------------------------
// DenseCholeskySolver: fbstab/components/dense_cholesky_solver.cc:32-127.
// K = [E G'; G -sigma I] (lower triangle only), factored by Eigen::LDLT.
// ---------------------------------------------------------------------------
struct DenseSolver : LinearSolver {
  const DenseData* data;
  int variant;  // 0 Eigen-pivoted LDLT, 1 unpivoted LDLT, 2 Cholesky + Schur
  int nz, nl, nv, n;
  Vec K, E, B, r1, r2, temp;
  std::vector<int> transp;
  // variant 2 storage
  Vec LE, W, LS;
  // variant 3 storage (Gauss-Jordan of E, as the warp kernel does it)
  Vec Ygj, dgj, Egj;

  DenseSolver(const DenseData* d, int variant_)
      : data(d), variant(variant_), nz(d->nz), nl(d->nl), nv(d->nv),
        n(d->nz + d->nl), K((size_t)n * n, 0.0), E((size_t)nz * nz, 0.0),
        B((size_t)nv * nz, 0.0), r1(n, 0.0), r2(nv, 0.0), temp(n, 0.0),
        transp(n, 0) {}

  double& k(int i, int j) { return K[(size_t)j * n + i]; }

  // Eigen 3.4.0 LDLT.h ldlt_inplace<Lower>::unblocked, restated.
  bool LdltFactor(bool pivot) {
    bool found_zero_pivot = false;
    bool ret = true;
    if (n <= 1) {
      if (n == 1) transp[0] = 0;
      return true;
    }
    for (int kk = 0; kk < n; kk++) {
      int big = kk;
      if (pivot) {
        // Largest |diagonal| of the trailing part.  The algorithm is
        // left-looking, so trailing diagonal entries are the ORIGINAL ones.
        double best = std::fabs(k(kk, kk));
        for (int i = kk + 1; i < n; i++) {
          const double a = std::fabs(k(i, i));
          if (a > best) {
            best = a;
            big = i;
          }
        }
      }
      transp[kk] = big;
      if (kk != big) {
        const int s = n - big - 1;
        for (int j = 0; j < kk; j++) std::swap(k(kk, j), k(big, j));
        for (int i = 0; i < s; i++)
          std::swap(k(big + 1 + i, kk), k(big + 1 + i, big));
        std::swap(k(kk, kk), k(big, big));
        for (int i = kk + 1; i < big; i++) std::swap(k(i, kk), k(big, i));
      }
      const int rs = n - kk - 1;
      if (kk > 0) {
        // temp = D(0:k) .* A10' ; a_kk -= A10*temp ; A21 -= A20*temp
        for (int j = 0; j < kk; j++) temp[j] = k(j, j) * k(kk, j);
        double s = 0.0;
        for (int j = 0; j < kk; j++) s += k(kk, j) * temp[j];
        k(kk, kk) -= s;
        if (rs > 0) {
          // column-major gemv: accumulate column by column
          double* acc = Scratch(1, rs);
          for (int i = 0; i < rs; i++) acc[i] = 0.0;
          for (int j = 0; j < kk; j++) {
            const double tj = temp[j];
            for (int i = 0; i < rs; i++) acc[i] += k(kk + 1 + i, j) * tj;
          }
          for (int i = 0; i < rs; i++) k(kk + 1 + i, kk) -= acc[i];
        }
      }
      const double akk = k(kk, kk);
      const bool pivot_is_valid = std::fabs(akk) > 0.0;
      if (kk == 0 && !pivot_is_valid) {
        for (int j = 0; j < n; j++) {
          transp[j] = j;
          for (int i = j + 1; i < n; i++) ret = ret && (k(i, j) == 0.0);
        }
        return ret;
      }
      if (rs > 0 && pivot_is_valid) {
        for (int i = 0; i < rs; i++) k(kk + 1 + i, kk) /= akk;
      } else if (rs > 0) {
        for (int i = 0; i < rs; i++) ret = ret && (k(kk + 1 + i, kk) == 0.0);
      }
      if (found_zero_pivot && pivot_is_valid)
        ret = false;
      else if (!pivot_is_valid)
        found_zero_pivot = true;
    }
    return ret;
  }

  // Eigen 3.4.0 LDLT::_solve_impl_transposed, restated.
  void LdltSolve(double* x) {
    for (int i = 0; i < n; i++)
      if (transp[i] != i) std::swap(x[i], x[transp[i]]);
    // L^-1, unit lower, column oriented
    for (int j = 0; j < n; j++) {
      const double xj = x[j];
      for (int i = j + 1; i < n; i++) x[i] -= k(i, j) * xj;
    }
    const double tol = std::numeric_limits<double>::min();
    for (int i = 0; i < n; i++) {
      if (std::fabs(k(i, i)) > tol)
        x[i] /= k(i, i);
      else
        x[i] = 0.0;
    }
    // L^-T, dot-product oriented
    for (int i = n - 1; i >= 0; i--) {
      double s = 0.0;
      for (int j = i + 1; j < n; j++) s += k(j, i) * x[j];
      x[i] -= s;
    }
    for (int i = n - 1; i >= 0; i--)
      if (transp[i] != i) std::swap(x[i], x[transp[i]]);
  }

  // variant 2: E = LE LE', W = G LE^-T, S = sigma I + W W' = LS LS'.
  bool CholSchurFactor(double sigma) {
    LE.assign((size_t)nz * nz, 0.0);
    for (int j = 0; j < nz; j++)
      for (int i = j; i < nz; i++) LE[(size_t)j * nz + i] = E[(size_t)j * nz + i];
    if (!Chol(LE.data(), nz)) return false;
    W.assign((size_t)nl * nz, 0.0);
    for (int r = 0; r < nl; r++) {
      // row r of W: solve LE w' = g_r'
      double* w = Scratch(2, nz);
      for (int j = 0; j < nz; j++) w[j] = data->G[(size_t)j * nl + r];
      for (int j = 0; j < nz; j++) {
        w[j] /= LE[(size_t)j * nz + j];
        for (int i = j + 1; i < nz; i++) w[i] -= LE[(size_t)j * nz + i] * w[j];
      }
      for (int j = 0; j < nz; j++) W[(size_t)j * nl + r] = w[j];
    }
    LS.assign((size_t)nl * nl, 0.0);
    for (int j = 0; j < nl; j++)
      for (int i = j; i < nl; i++) {
        double s = (i == j) ? sigma : 0.0;
        for (int kk = 0; kk < nz; kk++)
          s += W[(size_t)kk * nl + i] * W[(size_t)kk * nl + j];
        LS[(size_t)j * nl + i] = s;
      }
    return nl == 0 || Chol(LS.data(), nl);
  }
  static bool Chol(double* M, int m) {
    for (int kk = 0; kk < m; kk++) {
      double x = M[(size_t)kk * m + kk];
      for (int j = 0; j < kk; j++) x -= M[(size_t)j * m + kk] * M[(size_t)j * m + kk];
      if (x <= 0.0) return false;
      x = std::sqrt(x);
      M[(size_t)kk * m + kk] = x;
      for (int i = kk + 1; i < m; i++) {
        double s = M[(size_t)kk * m + i];
        for (int j = 0; j < kk; j++) s -= M[(size_t)j * m + i] * M[(size_t)j * m + kk];
        M[(size_t)kk * m + i] = s / x;
      }
    }
    return true;
  }
  static void CholFwd(const double* M, int m, double* x) {
    for (int j = 0; j < m; j++) {
      x[j] /= M[(size_t)j * m + j];
      for (int i = j + 1; i < m; i++) x[i] -= M[(size_t)j * m + i] * x[j];
    }
  }
  static void CholBwd(const double* M, int m, double* x) {
    for (int i = m - 1; i >= 0; i--) {
      double s = x[i];
      for (int j = i + 1; j < m; j++) s -= M[(size_t)i * m + j] * x[j];
      x[i] = s / M[(size_t)i * m + i];
    }
  }
  // [E G';G -sigma I][dz;dl] = [a;c]:
  //   t = LE^-1 a ; S dl = W t - c ; dz = LE^-T (t - W' dl)
  void CholSchurSolve(double* x) {
    double* a = x;
    double* c = x + nz;
    CholFwd(LE.data(), nz, a);
    Vec s(nl);
    for (int i = 0; i < nl; i++) {
      double acc = 0.0;
      for (int kk = 0; kk < nz; kk++) acc += W[(size_t)kk * nl + i] * a[kk];
      s[i] = acc - c[i];
    }
    if (nl > 0) {
      CholFwd(LS.data(), nl, s.data());
      CholBwd(LS.data(), nl, s.data());
    }
    for (int kk = 0; kk < nz; kk++) {
      double acc = 0.0;
      for (int i = 0; i < nl; i++) acc += W[(size_t)kk * nl + i] * s[i];
      a[kk] -= acc;
    }
    CholBwd(LE.data(), nz, a);
    for (int i = 0; i < nl; i++) c[i] = s[i];
  }

  // variant 3: unpivoted Gauss-Jordan on the symmetric E with the columns of
  // G' and the rhs riding along, then the Schur complement in the equality
  // block.  [E G';G -sigma I][dz;dl] = [a;c].
  void GaussJordanSolve(double* x) {
    const int m = nl + 1;
    Vec M((size_t)nz * nz), Y((size_t)nz * m);
    for (int i = 0; i < nz; i++)
      for (int j = 0; j < nz; j++)
        M[(size_t)i * nz + j] = (j <= i) ? Egj[(size_t)j * nz + i] : Egj[(size_t)i * nz + j];
    for (int i = 0; i < nz; i++) {
      for (int r = 0; r < nl; r++) Y[(size_t)i * m + r] = data->G[(size_t)i * nl + r];
      Y[(size_t)i * m + nl] = x[i];
    }
    for (int k = 0; k < nz; k++) {
      const double rd = 1.0 / M[(size_t)k * nz + k];
      for (int i = 0; i < nz; i++) {
        if (i == k) continue;
        const double lik = M[(size_t)i * nz + k] * rd;
        for (int j = k + 1; j < nz; j++)
          M[(size_t)i * nz + j] -= lik * M[(size_t)k * nz + j];
        for (int r = 0; r < m; r++) Y[(size_t)i * m + r] -= lik * Y[(size_t)k * m + r];
      }
    }
    for (int i = 0; i < nz; i++) {
      const double rd = 1.0 / M[(size_t)i * nz + i];
      for (int r = 0; r < m; r++) Y[(size_t)i * m + r] *= rd;
    }
    // S = -sigma I - G Y ; rhs = c - G t
    Vec S((size_t)nl * nl), rhs(nl);
    for (int r = 0; r < nl; r++) {
      for (int q = 0; q < nl; q++) {
        double acc = 0.0;
        for (int i = 0; i < nz; i++) acc += data->G[(size_t)i * nl + r] * Y[(size_t)i * m + q];
        S[(size_t)q * nl + r] = -acc;
      }
      double acc = 0.0;
      for (int i = 0; i < nz; i++) acc += data->G[(size_t)i * nl + r] * Y[(size_t)i * m + nl];
      rhs[r] = x[nz + r] - acc;
    }
    for (int r = 0; r < nl; r++) S[(size_t)r * nl + r] -= sigma_gj;
    // unpivoted Gaussian elimination on the small negative definite S
    for (int k = 0; k < nl; k++)
      for (int i = k + 1; i < nl; i++) {
        const double l = S[(size_t)k * nl + i] / S[(size_t)k * nl + k];
        for (int j = k + 1; j < nl; j++) S[(size_t)j * nl + i] -= l * S[(size_t)j * nl + k];
        rhs[i] -= l * rhs[k];
      }
    for (int i = nl - 1; i >= 0; i--) {
      double acc = rhs[i];
      for (int j = i + 1; j < nl; j++) acc -= S[(size_t)j * nl + i] * rhs[j];
      rhs[i] = acc / S[(size_t)i * nl + i];
    }
    for (int i = 0; i < nz; i++) {
      double acc = Y[(size_t)i * m + nl];
      for (int r = 0; r < nl; r++) acc -= Y[(size_t)i * m + r] * rhs[r];
      x[i] = acc;
    }
    for (int r = 0; r < nl; r++) x[nz + r] = rhs[r];
  }
  double sigma_gj = 0.0;

  // dense_cholesky_solver.cc:32-79
  bool Initialize(const Variable& x, const Variable& xbar,
                  double sigma) override {
    const double* H = data->H;
    const double* G = data->G;
    const double* A = data->A;
    for (int j = 0; j < nz; j++)
      for (int i = 0; i < nz; i++)
        E[(size_t)j * nz + i] = H[(size_t)j * nz + i] + (i == j ? sigma : 0.0);
    Barrier(x, xbar, sigma);
    // B = diag(Gamma) A ; E += A' B
    for (int j = 0; j < nz; j++)
      for (int i = 0; i < nv; i++)
        B[(size_t)j * nv + i] = Gamma[i] * A[(size_t)j * nv + i];
    for (int j = 0; j < nz; j++)
      for (int i = 0; i < nz; i++)
        E[(size_t)j * nz + i] +=
            dot(A + (size_t)i * nv, B.data() + (size_t)j * nv, nv);
    if (variant == 2) return CholSchurFactor(sigma);
    if (variant == 3) {
      Egj = E;
      sigma_gj = sigma;
      return true;
    }
    // K lower blocks (upper-right block is never written nor read)
    for (int j = 0; j < nz; j++) {
      for (int i = 0; i < nz; i++) k(i, j) = E[(size_t)j * nz + i];
      for (int i = 0; i < nl; i++) k(nz + i, j) = G[(size_t)j * nl + i];
    }
    for (int j = 0; j < nl; j++)
      for (int i = 0; i < nl; i++) k(nz + i, nz + j) = (i == j) ? -sigma : 0.0;
    return LdltFactor(variant == 0);
  }

  // dense_cholesky_solver.cc:81-127
  bool Solve(const Residual& r, Variable* x) override {
    const double* A = data->A;
    const double* b = data->b;
    for (int i = 0; i < nv; i++) r2[i] = r.v[i] / mus[i];
    for (int j = 0; j < nz; j++)
      r1[j] = r.z[j] - dot(A + (size_t)j * nv, r2.data(), nv);
    for (int i = 0; i < nl; i++) r1[nz + i] = -r.l[i];

    if (variant == 2)
      CholSchurSolve(r1.data());
    else if (variant == 3)
      GaussJordanSolve(r1.data());
    else
      LdltSolve(r1.data());
    for (int i = 0; i < nz; i++) x->z[i] = r1[i];
    for (int i = 0; i < nl; i++) x->l[i] = r1[nz + i];

    std::fill(r2.begin(), r2.end(), 0.0);
    gemv_n_acc(A, nv, nz, nv, x->z.data(), 1.0, r2.data());
    for (int i = 0; i < nv; i++) r2[i] = gamma[i] * r2[i];
    for (int i = 0; i < nv; i++) r2[i] += r.v[i];
    for (int i = 0; i < nv; i++) x->v[i] = r2[i] / mus[i];

    Vec Az(nv, 0.0);
    gemv_n_acc(A, nv, nz, nv, x->z.data(), 1.0, Az.data());
    for (int i = 0; i < nv; i++) x->y[i] = b[i] - Az[i];
    return true;
  }
};

// ---------------------------------------------------------------------------
// Small dense helpers for the Riccati recursion (column-major, ld = rows).
// ---------------------------------------------------------------------------
// Eigen LLT unblocked (LLT.h llt_inplace<Lower>::unblocked), lower, in place.
bool LltLower(double* M, int m) {
  for (int k = 0; k < m; k++) {
    const int rs = m - k - 1;
    double x = M[(size_t)k * m + k];
    if (k > 0) {
      double s = 0.0;
      for (int j = 0; j < k; j++) s += M[(size_t)j * m + k] * M[(size_t)j * m + k];
      x -= s;
    }
    if (x <= 0.0) return false;
    x = std::sqrt(x);
    M[(size_t)k * m + k] = x;
    if (k > 0 && rs > 0) {
      double* acc = Scratch(3, rs);
      for (int i = 0; i < rs; i++) acc[i] = 0.0;
      for (int j = 0; j < k; j++) {
        const double a = M[(size_t)j * m + k];
        for (int i = 0; i < rs; i++) acc[i] += M[(size_t)j * m + k + 1 + i] * a;
      }
      for (int i = 0; i < rs; i++) M[(size_t)k * m + k + 1 + i] -= acc[i];
    }
    for (int i = 0; i < rs; i++) M[(size_t)k * m + k + 1 + i] /= x;
  }
  return true;
}
// x <- L^-1 x  (lower, non-unit), column oriented
void TrsvLower(const double* L, int m, double* x) {
  for (int j = 0; j < m; j++) {
    x[j] /= L[(size_t)j * m + j];
    const double xj = x[j];
    for (int i = j + 1; i < m; i++) x[i] -= L[(size_t)j * m + i] * xj;
  }
}
// x <- L^-T x, dot-product oriented
void TrsvLowerT(const double* L, int m, double* x) {
  for (int i = m - 1; i >= 0; i--) {
    double s = x[i];
    for (int j = i + 1; j < m; j++) s -= L[(size_t)i * m + j] * x[j];
    x[i] = s / L[(size_t)i * m + i];
  }
}
// X(rows x m) <- X * L^-T  (each row of X: forward substitution with L)
void TrsmRightLowerT(const double* L, int m, double* X, int rows) {
  for (int j = 0; j < m; j++) {
    for (int kk = 0; kk < j; kk++) {
      const double ljk = L[(size_t)kk * m + j];
      for (int r = 0; r < rows; r++)
        X[(size_t)j * rows + r] -= X[(size_t)kk * rows + r] * ljk;
    }
    const double d = L[(size_t)j * m + j];
    for (int r = 0; r < rows; r++) X[(size_t)j * rows + r] /= d;
  }
}

// ---------------------------------------------------------------------------
// RiccatiLinearSolver: fbstab/components/riccati_linear_solver.cc:77-344
// ---------------------------------------------------------------------------
struct RiccatiSolver : LinearSolver {
  const MpcData* data;
  int N, nx, nu, nc, nz, nl, nv;
  std::vector<Vec> Q, S, R, P, SG, M, L, SM, AM, h, th;
  Vec Etemp, Ltemp, Linv, tx, tu, tl, r1, r2, r3;

  explicit RiccatiSolver(const MpcData* d)
      : data(d), N(d->N), nx(d->nx), nu(d->nu), nc(d->nc), nz(d->nz),
        nl(d->nl), nv(d->nv) {
    auto mk = [&](std::vector<Vec>& v, int r, int c) {
      v.assign(N + 1, Vec((size_t)r * c, 0.0));
    };
    mk(Q, nx, nx);
    mk(S, nu, nx);
    mk(R, nu, nu);
    mk(P, nx, nu);
    mk(SG, nu, nu);
    mk(M, nx, nx);
    mk(L, nx, nx);
    mk(SM, nu, nx);
    mk(AM, nx, nx);
    mk(h, nx, 1);
    mk(th, nx, 1);
    Etemp.assign((size_t)nc * nx, 0.0);
    Ltemp.assign((size_t)nc * nu, 0.0);
    Linv.assign((size_t)nx * nx, 0.0);
    tx.assign(nx, 0.0);
    tu.assign(nu, 0.0);
    tl.assign(nx, 0.0);
    r1.assign(nz, 0.0);
    r2.assign(nl, 0.0);
    r3.assign(nv, 0.0);
  }

  // Linv = inv(L L') as a full matrix: riccati_linear_solver.cc:142-144
  void InvLLt(const Vec& Lm) {
    std::fill(Linv.begin(), Linv.end(), 0.0);
    for (int j = 0; j < nx; j++) Linv[(size_t)j * nx + j] = 1.0;
    for (int j = 0; j < nx; j++) {
      TrsvLower(Lm.data(), nx, Linv.data() + (size_t)j * nx);
      TrsvLowerT(Lm.data(), nx, Linv.data() + (size_t)j * nx);
    }
  }
  // M(lower) = Q(lower) + Linv ; chol in place
  bool FactorM(int i) {
    for (int j = 0; j < nx; j++)
      for (int r = j; r < nx; r++)
        M[i][(size_t)j * nx + r] = Q[i][(size_t)j * nx + r] + Linv[(size_t)j * nx + r];
    return LltLower(M[i].data(), nx);
  }
  // SG = chol(R - SM SM')  (lower)
  bool FactorSG(int i) {
    for (int j = 0; j < nu; j++)
      for (int r = j; r < nu; r++) {
        double s = 0.0;
        for (int kk = 0; kk < nx; kk++)
          s += SM[i][(size_t)kk * nu + r] * SM[i][(size_t)kk * nu + j];
        SG[i][(size_t)j * nu + r] = R[i][(size_t)j * nu + r] - s;
      }
    return LltLower(SG[i].data(), nu);
  }

  // riccati_linear_solver.cc:77-210
  bool Initialize(const Variable& x, const Variable& xbar,
                  double sigma) override {
    Barrier(x, xbar, sigma);
    // Barrier-augmented stage Hessians, :102-123
    for (int i = 0; i < N + 1; i++) {
      const double* Ei = data->Ei(i);
      const double* Li = data->Li(i);
      const double* Gam = Gamma.data() + (size_t)i * nc;
      for (int j = 0; j < nx; j++)
        for (int r = j; r < nx; r++)
          Q[i][(size_t)j * nx + r] =
              data->Qi(i)[(size_t)j * nx + r] + (r == j ? sigma : 0.0);
      for (int j = 0; j < nu; j++)
        for (int r = j; r < nu; r++)
          R[i][(size_t)j * nu + r] =
              data->Ri(i)[(size_t)j * nu + r] + (r == j ? sigma : 0.0);
      for (int e = 0; e < nu * nx; e++) S[i][e] = data->Si(i)[e];

      for (int j = 0; j < nx; j++)
        for (int r = 0; r < nc; r++)
          Etemp[(size_t)j * nc + r] = Gam[r] * Ei[(size_t)j * nc + r];
      for (int j = 0; j < nx; j++)
        for (int r = j; r < nx; r++)
          Q[i][(size_t)j * nx + r] +=
              dot(Ei + (size_t)r * nc, Etemp.data() + (size_t)j * nc, nc);
      for (int j = 0; j < nu; j++)
        for (int r = 0; r < nc; r++)
          Ltemp[(size_t)j * nc + r] = Gam[r] * Li[(size_t)j * nc + r];
      for (int j = 0; j < nu; j++)
        for (int r = j; r < nu; r++)
          R[i][(size_t)j * nu + r] +=
              dot(Li + (size_t)r * nc, Ltemp.data() + (size_t)j * nc, nc);
      for (int j = 0; j < nx; j++)
        for (int r = 0; r < nu; r++)
          S[i][(size_t)j * nu + r] +=
              dot(Li + (size_t)r * nc, Etemp.data() + (size_t)j * nc, nc);
    }
    // L(0) = sqrt(sigma) I, :127
    std::fill(L[0].begin(), L[0].end(), 0.0);
    for (int j = 0; j < nx; j++) L[0][(size_t)j * nx + j] = std::sqrt(sigma);

    for (int i = 0; i < N; i++) {
      InvLLt(L[i]);
      if (!FactorM(i)) return false;
      // AM = A M^-T ; SM = S M^-T   :149-161
      for (int e = 0; e < nx * nx; e++) AM[i][e] = data->Ai(i)[e];
      TrsmRightLowerT(M[i].data(), nx, AM[i].data(), nx);
      SM[i] = S[i];
      TrsmRightLowerT(M[i].data(), nx, SM[i].data(), nu);
      if (!FactorSG(i)) return false;
      // P = (AM SM' - B) SG^-T   :170-175
      for (int j = 0; j < nu; j++)
        for (int r = 0; r < nx; r++) {
          double s = 0.0;
          for (int kk = 0; kk < nx; kk++)
            s += AM[i][(size_t)kk * nx + r] * SM[i][(size_t)kk * nu + j];
          P[i][(size_t)j * nx + r] = s;
        }
      for (int e = 0; e < nx * nu; e++) P[i][e] -= data->Bi(i)[e];
      TrsmRightLowerT(SG[i].data(), nu, P[i].data(), nx);
      // L(i+1) = chol(sigma I + P P' + AM AM')   :179-183
      Vec& Ln = L[i + 1];
      std::fill(Ln.begin(), Ln.end(), 0.0);
      for (int j = 0; j < nx; j++) Ln[(size_t)j * nx + j] = sigma;
      for (int j = 0; j < nx; j++)
        for (int r = 0; r < nx; r++) {
          double s = 0.0;
          for (int kk = 0; kk < nu; kk++)
            s += P[i][(size_t)kk * nx + r] * P[i][(size_t)kk * nx + j];
          Ln[(size_t)j * nx + r] += s;
        }
      for (int j = 0; j < nx; j++)
        for (int r = 0; r < nx; r++) {
          double s = 0.0;
          for (int kk = 0; kk < nx; kk++)
            s += AM[i][(size_t)kk * nx + r] * AM[i][(size_t)kk * nx + j];
          Ln[(size_t)j * nx + r] += s;
        }
      if (!LltLower(Ln.data(), nx)) return false;
    }
    // i = N step, :187-206
    InvLLt(L[N]);
    if (!FactorM(N)) return false;
    SM[N] = S[N];
    TrsmRightLowerT(M[N].data(), nx, SM[N].data(), nu);
    if (!FactorSG(N)) return false;
    return true;
  }

  // riccati_linear_solver.cc:212-344
  bool Solve(const Residual& r, Variable* dx) override {
    const int ns = nx + nu;
    r1 = r.z;
    for (int i = 0; i < nv; i++) r3[i] = r.v[i] / mus[i];
    data->gemvAT(r3.data(), -1.0, 1.0, r1.data());
    for (int i = 0; i < nl; i++) r2[i] = -r.l[i];
    auto r1x = [&](int i) { return r1.data() + (size_t)i * ns; };
    auto r1u = [&](int i) { return r1.data() + (size_t)i * ns + nx; };
    auto r2c = [&](int i) { return r2.data() + (size_t)i * nx; };

    // base case :232-236
    for (int k = 0; k < nx; k++) th[0][k] = r2c(0)[k];
    h[0] = th[0];
    TrsvLower(L[0].data(), nx, h[0].data());
    TrsvLowerT(L[0].data(), nx, h[0].data());
    for (int k = 0; k < nx; k++) h[0][k] -= r1x(0)[k];

    // forward :239-262
    for (int i = 0; i < N; i++) {
      tx = h[i];
      TrsvLower(M[i].data(), nx, tx.data());
      std::fill(tu.begin(), tu.end(), 0.0);
      gemv_n_acc(SM[i].data(), nu, nx, nu, tx.data(), 1.0, tu.data());
      for (int k = 0; k < nu; k++) tu[k] += r1u(i)[k];
      TrsvLower(SG[i].data(), nu, tu.data());

      std::fill(th[i + 1].begin(), th[i + 1].end(), 0.0);
      gemv_n_acc(P[i].data(), nx, nu, nx, tu.data(), 1.0, th[i + 1].data());
      gemv_n_acc(AM[i].data(), nx, nx, nx, tx.data(), 1.0, th[i + 1].data());
      for (int k = 0; k < nx; k++) th[i + 1][k] += r2c(i + 1)[k];

      h[i + 1] = th[i + 1];
      TrsvLower(L[i + 1].data(), nx, h[i + 1].data());
      TrsvLowerT(L[i + 1].data(), nx, h[i + 1].data());
      for (int k = 0; k < nx; k++) h[i + 1][k] -= r1x(i + 1)[k];
    }

    // terminal :267-285
    tx = h[N];
    TrsvLower(M[N].data(), nx, tx.data());
    std::fill(tu.begin(), tu.end(), 0.0);
    gemv_n_acc(SM[N].data(), nu, nx, nu, tx.data(), 1.0, tu.data());
    for (int k = 0; k < nu; k++) tu[k] += r1u(N)[k];
    TrsvLower(SG[N].data(), nu, tu.data());
    TrsvLowerT(SG[N].data(), nu, tu.data());

    tx = h[N];
    TrsvLower(M[N].data(), nx, tx.data());
    gemv_t_acc(SM[N].data(), nu, nx, nu, tu.data(), 1.0, tx.data());
    TrsvLowerT(M[N].data(), nx, tx.data());
    for (int k = 0; k < nx; k++) tx[k] *= -1.0;

    for (int k = 0; k < nx; k++) tl[k] = tx[k] + th[N][k];
    TrsvLower(L[N].data(), nx, tl.data());
    TrsvLowerT(L[N].data(), nx, tl.data());
    for (int k = 0; k < nx; k++) tl[k] *= -1.0;

    double* dz = dx->z.data();
    double* dl = dx->l.data();
    for (int k = 0; k < nx; k++) dz[(size_t)N * ns + k] = tx[k];
    for (int k = 0; k < nu; k++) dz[(size_t)N * ns + nx + k] = tu[k];
    for (int k = 0; k < nx; k++) dl[(size_t)N * nx + k] = tl[k];

    // backward :297-327
    for (int i = N - 1; i >= 0; i--) {
      tx = h[i];
      TrsvLower(M[i].data(), nx, tx.data());
      double* ui = dz + (size_t)i * ns + nx;
      double* xi = dz + (size_t)i * ns;
      double* li = dl + (size_t)i * nx;
      const double* lp1 = dl + (size_t)(i + 1) * nx;

      for (int k = 0; k < nu; k++) ui[k] = 0.0;
      gemv_n_acc(SM[i].data(), nu, nx, nu, tx.data(), 1.0, ui);
      for (int k = 0; k < nu; k++) ui[k] += r1u(i)[k];
      TrsvLower(SG[i].data(), nu, ui);
      gemv_t_acc(P[i].data(), nx, nu, nx, lp1, 1.0, ui);
      TrsvLowerT(SG[i].data(), nu, ui);

      for (int k = 0; k < nx; k++) xi[k] = h[i][k];
      TrsvLower(M[i].data(), nx, xi);
      gemv_t_acc(SM[i].data(), nu, nx, nu, ui, 1.0, xi);
      gemv_t_acc(AM[i].data(), nx, nx, nx, lp1, 1.0, xi);
      TrsvLowerT(M[i].data(), nx, xi);
      for (int k = 0; k < nx; k++) xi[k] *= -1.0;

      for (int k = 0; k < nx; k++) li[k] = th[i][k] + xi[k];
      TrsvLower(L[i].data(), nx, li);
      TrsvLowerT(L[i].data(), nx, li);
      for (int k = 0; k < nx; k++) li[k] *= -1.0;
    }

    // dv, dy  :331-341
    data->gemvA(dx->z.data(), 1.0, 0.0, r3.data());
    for (int i = 0; i < nv; i++)
      dx->v[i] = (r.v[i] + gamma[i] * r3[i]) / mus[i];
    data->gemvA(dx->z.data(), -1.0, 0.0, dx->y.data());
    data->axpyb(1.0, dx->y.data());
    return true;
  }
};

// ---------------------------------------------------------------------------
// Options: fbstab/fbstab_algorithm-impl.h:7-74
// ---------------------------------------------------------------------------
void DefaultParameters(oracle_options* o) {
  o->sigma0 = 1e-8;
  o->sigma_max = 1e-6;
  o->sigma_min = 1e-12;
  o->alpha = 0.95;
  o->beta = 0.75;
  o->eta = 1e-8;
  o->delta = 0.2;
  o->gamma = 0.1;
  o->abs_tol = 1e-6;
  o->rel_tol = 1e-12;
  o->stall_tol = 1e-10;
  o->infeas_tol = 1e-8;
  o->inner_tol_max = 1e-2;
  o->inner_tol_min = 1e-12;
  o->max_newton_iters = 200;
  o->max_prox_iters = 30;
  o->max_inner_iters = 50;
  o->max_linesearch_iters = 20;
  o->check_feasibility = 1;
  o->nonmonotone_linesearch = 1;
  o->display_level = 1;
  o->refine_steps = 0;
  o->regularize_retries = 0;
}

struct TrajSink {
  double* buf;
  int cap;
  int len;
  void push(double k, double a, double b, double c, double d, double e,
            double f, double g) {
    if (buf && len < cap) {
      double* p = buf + (size_t)len * ORACLE_TRAJ_STRIDE;
      p[0] = k; p[1] = a; p[2] = b; p[3] = c; p[4] = d; p[5] = e; p[6] = f; p[7] = g;
    }
    len++;
  }
};

// ---------------------------------------------------------------------------
// FBstabAlgorithm: fbstab/fbstab_algorithm-impl.h:113-304, 385-409
// ---------------------------------------------------------------------------
struct Algorithm {
  const Data* data;
  LinearSolver* ls;
  oracle_options opts;
  Variable xk, xi, xp, dx;
  Residual rk, ri;
  std::array<double, 5> merit_buffer{};  // fbstab_algorithm.h:175-181
  int newton_iters = 0, prox_iters = 0;
  int ls_backtracks = 0, residual_evals = 0;
  int status = 0;
  TrajSink traj{nullptr, 0, 0};

  Algorithm(const Data* d, LinearSolver* s, const oracle_options& o)
      : data(d), ls(s), opts(o), xk(d), xi(d), xp(d), dx(d), rk(d), ri(d) {
    rk.eval_counter = &residual_evals;
    ri.eval_counter = &residual_evals;
  }

  void InsertMerit(double x) {  // impl:402-409
    for (size_t i = merit_buffer.size() - 1; i > 0; i--)
      merit_buffer[i] = merit_buffer[i - 1];
    merit_buffer[0] = x;
  }
  double MaxMerit() const {
    return *std::max_element(merit_buffer.begin(), merit_buffer.end());
  }

  // impl:385-400 (status -> ExitFlag)
  int CheckForInfeasibility(const Variable& x) {
    const int feas = CheckFeasibility(data, x.z.data(), x.l.data(),
                                      x.v.data(), opts.infeas_tol);
    if (feas == 0) return 0;  // SUCCESS
    if (feas == 1) return 3;  // PRIMAL_INFEASIBLE
    if (feas == 2) return 4;  // DUAL_INFEASIBLE
    return 5;                 // PRIMAL_DUAL_INFEASIBLE
  }

  // One step of iterative refinement of the Newton system just solved: ri holds the
  // right-hand side, dx the solution, ls the factors.  A dz is read off dx.y = b - A dz.
  double sigma_ls = 0.0;
  bool Refine(double sigma) {
    const int nz = data->nz, nl = data->nl, nv = data->nv;
    Residual rr(data);
    Variable ddx(data);
    // (V dx)_z = H dz + G' dl + A' dv + sigma dz
    Vec t(nz, 0.0);
    data->gemvH(dx.z.data(), 1.0, 0.0, t.data());
    data->gemvGT(dx.l.data(), 1.0, 1.0, t.data());
    data->gemvAT(dx.v.data(), 1.0, 1.0, t.data());
    for (int i = 0; i < nz; i++) rr.z[i] = ri.z[i] - (t[i] + sigma * dx.z[i]);
    // (V dx)_l = -G dz + sigma dl
    Vec g(nl, 0.0);
    data->gemvG(dx.z.data(), 1.0, 0.0, g.data());
    for (int i = 0; i < nl; i++) rr.l[i] = ri.l[i] - (sigma * dx.l[i] - g[i]);
    // (V dx)_v = -gamma .* (A dz) + mu .* dv,  A dz = b - dx.y
    Vec adz(nv, 0.0);
    data->axpyb(1.0, adz.data());
    for (int i = 0; i < nv; i++) {
      adz[i] -= dx.y[i];
      rr.v[i] = ri.v[i] - (ls->mus[i] * dx.v[i] - ls->gamma[i] * adz[i]);
    }
    if (!ls->Solve(rr, &ddx)) return false;
    for (int i = 0; i < nz; i++) dx.z[i] += ddx.z[i];
    for (int i = 0; i < nl; i++) dx.l[i] += ddx.l[i];
    for (int i = 0; i < nv; i++) dx.v[i] += ddx.v[i];
    // dy = b - A (dz + ddz) = dx.y + ddx.y - b
    Vec bb(nv, 0.0);
    data->axpyb(1.0, bb.data());
    for (int i = 0; i < nv; i++) dx.y[i] = (dx.y[i] + ddx.y[i]) - bb[i];
    return true;
  }

  // impl:229-304
  double SolveProximalSubproblem(Variable* x, Variable* xbar, double tol,
                                 double sigma, double current_outer_residual) {
    merit_buffer.fill(0.0);
    double Eo = 0;
    double t = 1.0;
    for (int i = 0; i < opts.max_inner_iters; i++) {
      ri.InnerResidual(*x, *xbar, sigma);
      const double Ei = ri.Norm();
      rk.PenalizedNaturalResidual(*x);
      Eo = rk.Norm();
      if ((Ei <= tol && Eo < current_outer_residual) ||
          (Ei <= opts.inner_tol_min)) {
        break;
      }
      if (newton_iters >= opts.max_newton_iters) break;

      {
        // "TODO: regularize and retry" (riccati_linear_solver.cc:129-130): with
        // regularize_retries > 0 a failed factorisation is repeated with sigma x 100
        // per attempt in the linear solver only (default 0: throw like the reference)
        bool ok = ls->Initialize(*x, *xbar, sigma);
        double sig_r = sigma;
        for (int j = 0; !ok && j < opts.regularize_retries; j++) {
          sig_r *= 100.0;
          ok = ls->Initialize(*x, *xbar, sig_r);
        }
        if (!ok) {
          status = 1;
          throw status;
        }
        sigma_ls = sig_r;
      }
      ri.Negate();
      if (!ls->Solve(ri, &dx)) {
        status = 3;
        throw status;
      }
      // "TODO: implement iterative refinement" (abstract_components.h:335-337):
      // rho = rhs - V dx with V = [H + sigma I, G', A'; -G, sigma I, 0; -gamma A, 0, mu]
      // (dense_cholesky_solver.h:49-62: "the matrix V(x,xbar,sigma)"), one more solve with the same factors, dx += ddx.
      for (int it = 0; it < std::min(opts.refine_steps, 1); it++) {
        if (!Refine(sigma_ls)) {
          status = 3;
          throw status;
        }
      }
      newton_iters++;

      const double current_merit = ri.Merit();
      InsertMerit(current_merit);
      const double m0 =
          opts.nonmonotone_linesearch ? MaxMerit() : current_merit;
      t = 1.0;
      int rejected = 0;
      for (int j = 0; j < opts.max_linesearch_iters; j++) {
        xp.Copy(*x);
        xp.axpy(t, dx);
        ri.InnerResidual(xp, *xbar, sigma);
        const double mp = ri.Merit();
        if (mp <= m0 - 2.0 * t * opts.eta * current_merit) {
          break;
        } else {
          t *= opts.beta;
          rejected++;
        }
      }
      ls_backtracks += rejected;
      x->axpy(t, dx);
      traj.push(1, prox_iters, newton_iters, Ei, Eo, t, rejected, i);
    }
    x->ProjectDuals();
    return Eo;
  }

  // impl:113-224.  Returns the exit flag; the solution is left in *res.
  int Solve(double* z0, double* l0, double* v0, double* y0, double* residual,
            double* initial_residual) {
    rk.alpha = opts.alpha;
    ri.alpha = opts.alpha;
    ls->alpha = opts.alpha;
    const int nz = data->nz, nl = data->nl, nv = data->nv;

    const double sigma = opts.sigma0;
    const double combo_tol =
        opts.abs_tol + opts.rel_tol * (1.0 + data->forcing_norm);

    auto write = [&](const Variable& x) {  // impl:349-360
      std::copy(x.z.begin(), x.z.end(), z0);
      std::copy(x.l.begin(), x.l.end(), l0);
      std::copy(x.v.begin(), x.v.end(), v0);
      std::copy(x.y.begin(), x.y.end(), y0);
    };

    // impl:334-347
    std::copy(z0, z0 + nz, xk.z.begin());
    std::copy(l0, l0 + nl, xk.l.begin());
    std::copy(v0, v0 + nv, xk.v.begin());
    xk.InitializeConstraintMargin();
    xi.Copy(xk);
    dx.Fill(1.0);

    rk.PenalizedNaturalResidual(xk);
    ri.Fill(0.0);
    const double E0 = rk.Norm();
    double Ek = E0;
    *initial_residual = E0;
    double inner_tol = saturate(E0, opts.inner_tol_min, opts.inner_tol_max);

    newton_iters = 0;
    prox_iters = 0;

    for (int k = 0; k < opts.max_prox_iters; k++) {
      rk.PenalizedNaturalResidual(xk);
      Ek = rk.Norm();
      traj.push(0, prox_iters, newton_iters, Ek, inner_tol, 0, 0, 0);
      if (Ek <= combo_tol || dx.Norm() <= opts.stall_tol) {
        *residual = rk.Norm();
        write(xk);
        return 0;  // SUCCESS
      }
      inner_tol = saturate(inner_tol * opts.delta, opts.inner_tol_min, Ek);

      xi.Copy(xk);
      const double Eo =
          SolveProximalSubproblem(&xi, &xk, inner_tol, sigma, Ek);

      if (newton_iters >= opts.max_newton_iters) {
        if (Eo < Ek) {
          write(xi);
          rk.PenalizedNaturalResidual(xi);
        } else {
          write(xk);
          rk.PenalizedNaturalResidual(xk);
        }
        *residual = rk.Norm();
        return 2;  // MAXITERATIONS
      }

      dx.Copy(xi);
      dx.axpy(-1.0, xk);
      if (opts.check_feasibility) {
        const int eflag = CheckForInfeasibility(dx);
        if (eflag != 0) {
          *residual = rk.Norm();
          write(dx);
          return eflag;
        }
      }
      xk.Copy(xi);
      prox_iters++;
    }
    *residual = rk.Norm();
    write(xk);
    return 2;  // MAXITERATIONS
  }
};

}  // namespace

// ===========================================================================
// Sparse QPs (FBstabSparse).  The reference has no sparse solver yet: it plans
// "general sparse matrix components" (ROADMAP.md:10, README.md:47) on the LDL'
// wrapper tools/qdldl/qdldl_wrapper.h:19-84, whose upstream source (QDLDL,
// github.com/oxfordcontrol/qdldl, v0.1.x; tools/qdldl/BUILD.bazel:15,19 expects
// it under tools/qdldl/qdldl/) is NOT in the reference tree.  QDLDL_etree,
// QDLDL_factor and QDLDL_solve are therefore restated here from the published
// algorithm and pinned by the wrapper's own test
// (tools/qdldl/test/qdldl_test.cc:33-60: the upstream example, ||Ax - b|| <= 1e-12).
// The data concept is abstract_components.h:24-62 over compressed-column
// matrices; the Newton system is the quasi-definite
//   [ H + sigma I   G'       (Gamma^1/2 A)' ]
//   [ G            -sigma I   0             ]   (order [z; l; w], w = Gamma^1/2 A dz)
//   [ Gamma^1/2 A   0        -I             ]
// -- the reduction of dense_cholesky_solver.cc:32-127 with A' Gamma A unformed.
// ===========================================================================
static const int QDLDL_UNKNOWN = -1;

// Elimination tree and column counts of L for an upper-triangular CSC matrix.
// Returns nnz(L), -1 if an entry lies below the diagonal or a column is empty.
int QdldlEtree(int n, const int* Ap, const int* Ai, int* work, int* Lnz, int* etree) {
  for (int i = 0; i < n; i++) {
    work[i] = 0;
    Lnz[i] = 0;
    etree[i] = QDLDL_UNKNOWN;
    if (Ap[i] == Ap[i + 1]) return -1;
  }
  for (int j = 0; j < n; j++) {
    work[j] = j;
    for (int p = Ap[j]; p < Ap[j + 1]; p++) {
      int i = Ai[p];
      if (i > j) return -1;
      while (work[i] != j) {
        if (etree[i] == QDLDL_UNKNOWN) etree[i] = j;
        Lnz[i]++;
        work[i] = j;
        i = etree[i];
      }
    }
  }
  int sum = 0;
  for (int i = 0; i < n; i++) sum += Lnz[i];
  return sum;
}

// Up-looking LDL' (L unit lower triangular, strictly lower part stored by columns).
// Returns the number of positive entries of D, -1 on a zero pivot.
int QdldlFactor(int n, const int* Ap, const int* Ai, const double* Ax, int* Lp, int* Li,
                double* Lx, double* D, double* Dinv, const int* Lnz, const int* etree,
                unsigned char* bwork, int* iwork, double* fwork) {
  int positive = 0;
  unsigned char* yMarkers = bwork;
  int* yIdx = iwork;
  int* elimBuffer = iwork + n;
  int* LNextSpaceInCol = iwork + 2 * n;
  double* yVals = fwork;
  Lp[0] = 0;
  for (int i = 0; i < n; i++) {
    Lp[i + 1] = Lp[i] + Lnz[i];
    yMarkers[i] = 0;
    yVals[i] = 0.0;
    D[i] = 0.0;
    LNextSpaceInCol[i] = Lp[i];
  }
  D[0] = Ax[0];
  if (D[0] == 0.0) return -1;
  if (D[0] > 0.0) positive++;
  Dinv[0] = 1 / D[0];
  for (int k = 1; k < n; k++) {
    int nnzY = 0;
    for (int i = Ap[k]; i < Ap[k + 1]; i++) {
      const int bidx = Ai[i];
      if (bidx == k) {
        D[k] = Ax[i];
        continue;
      }
      yVals[bidx] = Ax[i];
      int nextIdx = bidx;
      if (yMarkers[nextIdx] == 0) {
        yMarkers[nextIdx] = 1;
        elimBuffer[0] = nextIdx;
        int nnzE = 1;
        nextIdx = etree[bidx];
        while (nextIdx != QDLDL_UNKNOWN && nextIdx < k) {
          if (yMarkers[nextIdx] == 1) break;
          yMarkers[nextIdx] = 1;
          elimBuffer[nnzE] = nextIdx;
          nnzE++;
          nextIdx = etree[nextIdx];
        }
        while (nnzE) yIdx[nnzY++] = elimBuffer[--nnzE];
      }
    }
    for (int i = nnzY - 1; i >= 0; i--) {
      const int cidx = yIdx[i];
      const int tmpIdx = LNextSpaceInCol[cidx];
      const double yVals_cidx = yVals[cidx];
      for (int j = Lp[cidx]; j < tmpIdx; j++) yVals[Li[j]] -= Lx[j] * yVals_cidx;
      Li[tmpIdx] = k;
      Lx[tmpIdx] = yVals_cidx * Dinv[cidx];
      D[k] -= yVals_cidx * Lx[tmpIdx];
      LNextSpaceInCol[cidx]++;
      yVals[cidx] = 0.0;
      yMarkers[cidx] = 0;
    }
    if (D[k] == 0.0) return -1;
    if (D[k] > 0.0) positive++;
    Dinv[k] = 1 / D[k];
  }
  return positive;
}

// x <- (L D L')^-1 x
void QdldlSolve(int n, const int* Lp, const int* Li, const double* Lx, const double* Dinv,
                double* x) {
  for (int i = 0; i < n; i++) {
    const double val = x[i];
    for (int j = Lp[i]; j < Lp[i + 1]; j++) x[Li[j]] -= Lx[j] * val;
  }
  for (int i = 0; i < n; i++) x[i] *= Dinv[i];
  for (int i = n - 1; i >= 0; i--) {
    double val = x[i];
    for (int j = Lp[i]; j < Lp[i + 1]; j++) val -= Lx[j] * x[Li[j]];
    x[i] = val;
  }
}

// tools/qdldl/qdldl_wrapper.h:19-84
struct QdldlWrapper {
  int n, nnz = 0;
  std::vector<int> etree, Lnz, Lp, Li, iwork;
  Vec Lx, D, Dinv, fwork;
  std::vector<unsigned char> bwork;
  bool ok = false;
  QdldlWrapper(int n_, const int* Ap, const int* Ai)
      : n(n_), etree(n_), Lnz(n_), Lp(n_ + 1), iwork(3 * n_), D(n_), Dinv(n_), fwork(n_),
        bwork(n_) {
    nnz = QdldlEtree(n, Ap, Ai, iwork.data(), Lnz.data(), etree.data());
    if (nnz < 0) return;
    Li.resize(nnz);
    Lx.resize(nnz);
    ok = true;
  }
  bool Factor(const int* Ap, const int* Ai, const double* Ax) {
    return ok && QdldlFactor(n, Ap, Ai, Ax, Lp.data(), Li.data(), Lx.data(), D.data(),
                             Dinv.data(), Lnz.data(), etree.data(), bwork.data(), iwork.data(),
                             fwork.data()) >= 0;
  }
  void Solve(double* x) const { QdldlSolve(n, Lp.data(), Li.data(), Lx.data(), Dinv.data(), x); }
};

// Data concept over compressed-column matrices: H by its upper triangle, G, A.
// Products gather along rows (entries of a row in increasing column order), the
// transposed ones along the stored columns.
struct SparseData : Data {
  const int *Hp, *Hi, *Gp, *Gi, *Ap, *Ai;
  const double *Hx, *f, *Gx, *h, *Ax, *b;
  struct Rows {
    std::vector<int> ptr, col, val;
  };
  Rows Hr, Gr, Ar;
  static Rows MakeRows(int rows, int cols, const int* p, const int* i, bool symmetric) {
    std::vector<std::vector<std::pair<int, int>>> r(rows);
    for (int c = 0; c < cols; c++)
      for (int e = p[c]; e < p[c + 1]; e++) {
        r[i[e]].push_back({c, e});
        if (symmetric && i[e] != c) r[c].push_back({i[e], e});
      }
    Rows out;
    out.ptr.assign(rows + 1, 0);
    for (int k = 0; k < rows; k++) {
      std::sort(r[k].begin(), r[k].end());
      out.ptr[k + 1] = out.ptr[k] + (int)r[k].size();
      for (auto& pr : r[k]) {
        out.col.push_back(pr.first);
        out.val.push_back(pr.second);
      }
    }
    return out;
  }
  SparseData(int nz_, int nl_, int nv_, const int* Hp_, const int* Hi_, const double* Hx_,
             const double* f_, const int* Gp_, const int* Gi_, const double* Gx_,
             const double* h_, const int* Ap_, const int* Ai_, const double* Ax_,
             const double* b_)
      : Hp(Hp_), Hi(Hi_), Gp(Gp_), Gi(Gi_), Ap(Ap_), Ai(Ai_), Hx(Hx_), f(f_), Gx(Gx_), h(h_),
        Ax(Ax_), b(b_) {
    nz = nz_;
    nl = nl_;
    nv = nv_;
    Hr = MakeRows(nz, nz, Hp, Hi, true);
    if (nl > 0) Gr = MakeRows(nl, nz, Gp, Gi, false);
    else Gr.ptr.assign(1, 0);
    Ar = MakeRows(nv, nz, Ap, Ai, false);
    forcing_norm = std::sqrt(dot(b, b, nv) + dot(f, f, nz) + dot(h, h, nl));
  }
  static void RowsGemv(const Rows& R, const double* X, int rows, const double* x, double a,
                       double bb, double* y) {
    scale_by_b(bb, y, rows);
    for (int r = 0; r < rows; r++) {
      double s = 0.0;
      for (int q = R.ptr[r]; q < R.ptr[r + 1]; q++) s += X[R.val[q]] * x[R.col[q]];
      y[r] += a * s;
    }
  }
  static void ColsGemvT(int cols, const int* p, const int* i, const double* X, const double* x,
                        double a, double bb, double* y) {
    scale_by_b(bb, y, cols);
    for (int c = 0; c < cols; c++) {
      double s = 0.0;
      for (int e = p[c]; e < p[c + 1]; e++) s += X[e] * x[i[e]];
      y[c] += a * s;
    }
  }
  void gemvH(const double* x, double a, double bb, double* y) const override {
    RowsGemv(Hr, Hx, nz, x, a, bb, y);
  }
  void gemvA(const double* x, double a, double bb, double* y) const override {
    RowsGemv(Ar, Ax, nv, x, a, bb, y);
  }
  void gemvAT(const double* x, double a, double bb, double* y) const override {
    ColsGemvT(nz, Ap, Ai, Ax, x, a, bb, y);
  }
  void gemvG(const double* x, double a, double bb, double* y) const override {
    RowsGemv(Gr, Gx, nl, x, a, bb, y);
  }
  void gemvGT(const double* x, double a, double bb, double* y) const override {
    if (nl > 0) ColsGemvT(nz, Gp, Gi, Gx, x, a, bb, y);
    else scale_by_b(bb, y, nz);
  }
  void axpyf(double a, double* y) const override {
    for (int i = 0; i < nz; i++) y[i] += a * f[i];
  }
  void axpyh(double a, double* y) const override {
    for (int i = 0; i < nl; i++) y[i] += a * h[i];
  }
  void axpyb(double a, double* y) const override {
    for (int i = 0; i < nv; i++) y[i] += a * b[i];
  }
};

// LinearSolver over the QDLDL factorisation of the permuted K.
struct SparseSolver : LinearSolver {
  const SparseData* data;
  int nz, nl, nv, n;
  std::vector<int> perm, iperm;  // perm[new] = old
  std::vector<int> Kp, Ki;       // permuted upper-triangular CSC pattern
  struct Src {
    int kind, idx, row;  // 0 Hx[idx], 1 Hx[idx] + sigma, 2 sigma, 3 Gx[idx], 4 -sigma,
  };                     // 5 sqrt(Gamma[row]) * Ax[idx], 6 -1
  std::vector<Src> Ksrc;
  Vec Kx, x, r3, adz;
  std::unique_ptr<QdldlWrapper> ldl;

  SparseSolver(const SparseData* d, const int* perm_)
      : data(d), nz(d->nz), nl(d->nl), nv(d->nv), n(d->nz + d->nl + d->nv), perm(n), iperm(n),
        x(n, 0.0), r3(d->nv, 0.0), adz(d->nv, 0.0) {
    // default order: the w block first, then z, then l -- the reduction of
    // dense_cholesky_solver.cc:32-127 (E = H + sigma I + A' Gamma A, then the Schur
    // complement on l)
    for (int k = 0; k < n; k++)
      perm[k] = perm_ ? perm_[k] : (k < nv ? nz + nl + k : k - nv);
    for (int k = 0; k < n; k++) iperm[perm[k]] = k;
    struct T {
      int row, col;
      Src s;
    };
    std::vector<T> t;
    auto add = [&](int r, int c, Src s) {
      const int a = iperm[r], b = iperm[c];
      t.push_back({std::min(a, b), std::max(a, b), s});
    };
    for (int c = 0; c < nz; c++) {
      bool diag = false;
      for (int e = d->Hp[c]; e < d->Hp[c + 1]; e++) {
        if (d->Hi[e] == c) {
          add(c, c, {1, e, 0});
          diag = true;
        } else {
          add(d->Hi[e], c, {0, e, 0});
        }
      }
      if (!diag) add(c, c, {2, 0, 0});
    }
    for (int c = 0; c < nz; c++)
      for (int e = (nl > 0 ? d->Gp[c] : 0); e < (nl > 0 ? d->Gp[c + 1] : 0); e++)
        add(c, nz + d->Gi[e], {3, e, 0});
    for (int r = 0; r < nl; r++) add(nz + r, nz + r, {4, 0, 0});
    for (int c = 0; c < nz; c++)
      for (int e = d->Ap[c]; e < d->Ap[c + 1]; e++)
        add(c, nz + nl + d->Ai[e], {5, e, d->Ai[e]});
    for (int k = 0; k < nv; k++) add(nz + nl + k, nz + nl + k, {6, 0, 0});
    std::sort(t.begin(), t.end(), [](const T& a, const T& b) {
      return a.col != b.col ? a.col < b.col : a.row < b.row;
    });
    Kp.assign(n + 1, 0);
    for (const T& e : t) {
      Kp[e.col + 1]++;
      Ki.push_back(e.row);
      Ksrc.push_back(e.s);
    }
    for (int c = 0; c < n; c++) Kp[c + 1] += Kp[c];
    Kx.assign(Ki.size(), 0.0);
    ldl.reset(new QdldlWrapper(n, Kp.data(), Ki.data()));
  }

  bool Initialize(const Variable& xv, const Variable& xbar, double sigma) override {
    Barrier(xv, xbar, sigma);
    for (size_t e = 0; e < Ksrc.size(); e++) {
      const Src& s = Ksrc[e];
      double v;
      switch (s.kind) {
        case 0: v = data->Hx[s.idx]; break;
        case 1: v = data->Hx[s.idx] + sigma; break;
        case 2: v = sigma; break;
        case 3: v = data->Gx[s.idx]; break;
        case 4: v = -sigma; break;
        case 5: v = std::sqrt(Gamma[s.row]) * data->Ax[s.idx]; break;
        default: v = -1.0;
      }
      Kx[e] = v;
    }
    return ldl->Factor(Kp.data(), Ki.data(), Kx.data());
  }

  // r is already negated by the caller (impl:270); same reduction as
  // dense_cholesky_solver.cc:81-127
  bool Solve(const Residual& r, Variable* dx) override {
    Vec r1 = r.z;
    for (int i = 0; i < nv; i++) r3[i] = r.v[i] / mus[i];
    data->gemvAT(r3.data(), -1.0, 1.0, r1.data());
    for (int i = 0; i < nz; i++) x[iperm[i]] = r1[i];
    for (int i = 0; i < nl; i++) x[iperm[nz + i]] = -r.l[i];
    for (int i = 0; i < nv; i++) x[iperm[nz + nl + i]] = 0.0;
    ldl->Solve(x.data());
    for (int i = 0; i < nz; i++) dx->z[i] = x[iperm[i]];
    for (int i = 0; i < nl; i++) dx->l[i] = x[iperm[nz + i]];
    data->gemvA(dx->z.data(), 1.0, 0.0, adz.data());
    for (int i = 0; i < nv; i++) {
      dx->v[i] = (r.v[i] + gamma[i] * adz[i]) / mus[i];
      dx->y[i] = data->b[i] - adz[i];
    }
    return true;
  }
};

// ===========================================================================
// C interface
// ===========================================================================
struct oracle_problem {
  Data* data = nullptr;
  DenseData* dense = nullptr;
  MpcData* mpc = nullptr;
  SparseData* sparse = nullptr;
  std::vector<int> perm;  // sparse: elimination order of K (empty = natural)
  ~oracle_problem() { delete data; }
};

static LinearSolver* MakeSolver(const oracle_problem* p, int variant) {
  if (p->dense) return new DenseSolver(p->dense, variant);
  if (p->sparse) return new SparseSolver(p->sparse, p->perm.empty() ? nullptr : p->perm.data());
  return new RiccatiSolver(p->mpc);
}

extern "C" {

void oracle_default_options(oracle_options* o) { DefaultParameters(o); }

void oracle_reliable_options(oracle_options* o) {
  DefaultParameters(o);
  o->sigma0 = 1e-4;
  o->sigma_max = 1e-2;
  o->sigma_min = 1e-10;
  o->beta = 0.9;
  o->abs_tol = 1e-4;
  o->rel_tol = 1e-6;
  o->max_linesearch_iters = 40;
  o->max_newton_iters = 500;
  o->max_prox_iters = 100;
  o->nonmonotone_linesearch = 0;
}

int oracle_validate_options(oracle_options* o) {
  try {
    o->sigma0 = std::max(o->sigma0, 1e-10);
    o->sigma_max = saturate(o->sigma_max, 1e-6, 1e2);
    o->sigma_min = saturate(o->sigma_min, 1e-13, 1e-8);
    o->sigma0 = saturate(o->sigma0, o->sigma_min, o->sigma_max);
    o->alpha = saturate(o->alpha, 0.001, 0.999);
    o->beta = saturate(o->beta, 0.1, 0.99);
    o->eta = saturate(o->eta, 1e-12, 0.499);
    o->delta = saturate(o->delta, 0.0001, 0.99);
    o->gamma = saturate(o->gamma, 0.001, 0.9);
    o->abs_tol = std::max(o->abs_tol, 1e-14);
    o->rel_tol = std::max(o->rel_tol, 0.0);
    o->stall_tol = std::max(o->stall_tol, 1e-14);
    o->infeas_tol = std::max(o->infeas_tol, 1e-14);
    o->inner_tol_max = saturate(o->inner_tol_max, 1e-8, 1e2);
    o->inner_tol_min = saturate(o->inner_tol_min, 1e-14, 1e-2);
    o->max_newton_iters = std::max(o->max_newton_iters, 1);
    o->max_prox_iters = std::max(o->max_prox_iters, 1);
    o->max_inner_iters = std::max(o->max_inner_iters, 1);
    o->max_linesearch_iters = std::max(o->max_linesearch_iters, 1);
    o->refine_steps = std::min(std::max(o->refine_steps, 0), 1);
    o->regularize_retries = std::min(std::max(o->regularize_retries, 0), 8);
  } catch (SaturateError&) {
    return 2;
  }
  return 0;
}

oracle_problem* oracle_dense_create(int nz, int nl, int nv, const double* H,
                                    const double* f, const double* G,
                                    const double* h, const double* A,
                                    const double* b) {
  if (nz <= 0 || nv <= 0 || nl < 0) return nullptr;  // fbstab_dense.cc:19-23
  auto* p = new oracle_problem;
  p->dense = new DenseData(nz, nl, nv, H, f, G, h, A, b);
  p->data = p->dense;
  return p;
}

oracle_problem* oracle_mpc_create(int N, int nx, int nu, int nc,
                                  const double* Q, const double* R,
                                  const double* S, const double* q,
                                  const double* r, const double* A,
                                  const double* B, const double* c,
                                  const double* E, const double* L,
                                  const double* d, const double* x0) {
  if (N < 1 || nx < 1 || nu < 1 || nc < 1) return nullptr;  // fbstab_mpc.cc:62-65
  auto* p = new oracle_problem;
  p->mpc = new MpcData(N, nx, nu, nc, Q, R, S, q, r, A, B, c, E, L, d, x0);
  p->data = p->mpc;
  return p;
}

void oracle_destroy(oracle_problem* p) { delete p; }

void oracle_sizes(const oracle_problem* p, int* nz, int* nl, int* nv) {
  *nz = p->data->nz;
  *nl = p->data->nl;
  *nv = p->data->nv;
}

double oracle_forcing_norm(const oracle_problem* p) {
  return p->data->forcing_norm;
}

int oracle_gemv(const oracle_problem* p, int op, const double* x, double a,
                double b, double* y) {
  switch (op) {
    case 0: p->data->gemvH(x, a, b, y); return 0;
    case 1: p->data->gemvA(x, a, b, y); return 0;
    case 2: p->data->gemvAT(x, a, b, y); return 0;
    case 3: p->data->gemvG(x, a, b, y); return 0;
    case 4: p->data->gemvGT(x, a, b, y); return 0;
  }
  return -1;
}

int oracle_axpy(const oracle_problem* p, int which, double a, double* y) {
  switch (which) {
    case 0: p->data->axpyf(a, y); return 0;
    case 1: p->data->axpyh(a, y); return 0;
    case 2: p->data->axpyb(a, y); return 0;
  }
  return -1;
}

void oracle_margin(const oracle_problem* p, const double* z, double* y) {
  Variable x(p->data);
  std::copy(z, z + p->data->nz, x.z.begin());
  x.InitializeConstraintMargin();
  std::copy(x.y.begin(), x.y.end(), y);
}

static void LoadVar(Variable* x, const double* z, const double* l,
                    const double* v, const double* y) {
  if (z) std::copy(z, z + x->z.size(), x->z.begin());
  if (l) std::copy(l, l + x->l.size(), x->l.begin());
  if (v) std::copy(v, v + x->v.size(), x->v.begin());
  if (y) std::copy(y, y + x->y.size(), x->y.begin());
}

void oracle_variable_axpy(const oracle_problem* p, double a, const double* dz,
                          const double* dl, const double* dv, const double* dy,
                          double* z, double* l, double* v, double* y) {
  Variable x(p->data), d(p->data);
  LoadVar(&x, z, l, v, y);
  LoadVar(&d, dz, dl, dv, dy);
  x.axpy(a, d);
  std::copy(x.z.begin(), x.z.end(), z);
  std::copy(x.l.begin(), x.l.end(), l);
  std::copy(x.v.begin(), x.v.end(), v);
  std::copy(x.y.begin(), x.y.end(), y);
}

void oracle_residual(const oracle_problem* p, int kind, double alpha,
                     double sigma, const double* z, const double* l,
                     const double* v, const double* y, const double* zbar,
                     const double* lbar, const double* vbar, double* rz,
                     double* rl, double* rv, double* norms) {
  Variable x(p->data), xbar(p->data);
  LoadVar(&x, z, l, v, y);
  LoadVar(&xbar, zbar, lbar, vbar, nullptr);
  Residual r(p->data);
  r.alpha = alpha;
  if (kind == 0)
    r.InnerResidual(x, xbar, sigma);
  else if (kind == 1)
    r.NaturalResidual(x);
  else
    r.PenalizedNaturalResidual(x);
  std::copy(r.z.begin(), r.z.end(), rz);
  std::copy(r.l.begin(), r.l.end(), rl);
  std::copy(r.v.begin(), r.v.end(), rv);
  if (norms) {
    norms[0] = r.znorm;
    norms[1] = r.lnorm;
    norms[2] = r.vnorm;
  }
}

int oracle_linear_solve(const oracle_problem* p, int variant, double alpha,
                        double sigma, const double* z, const double* l,
                        const double* v, const double* y, const double* zbar,
                        const double* lbar, const double* vbar,
                        const double* rz, const double* rl, const double* rv,
                        double* dz, double* dl, double* dv, double* dy,
                        double* gamma, double* mus) {
  Variable x(p->data), xbar(p->data), dx(p->data);
  LoadVar(&x, z, l, v, y);
  LoadVar(&xbar, zbar, lbar, vbar, nullptr);
  Residual r(p->data);
  std::copy(rz, rz + r.z.size(), r.z.begin());
  std::copy(rl, rl + r.l.size(), r.l.begin());
  std::copy(rv, rv + r.v.size(), r.v.begin());
  LinearSolver* ls = MakeSolver(p, variant);
  ls->alpha = alpha;
  int rc = 0;
  if (!ls->Initialize(x, xbar, sigma)) {
    rc = 1;
  } else {
    ls->Solve(r, &dx);
    std::copy(dx.z.begin(), dx.z.end(), dz);
    std::copy(dx.l.begin(), dx.l.end(), dl);
    std::copy(dx.v.begin(), dx.v.end(), dv);
    std::copy(dx.y.begin(), dx.y.end(), dy);
    if (gamma) std::copy(ls->gamma.begin(), ls->gamma.end(), gamma);
    if (mus) std::copy(ls->mus.begin(), ls->mus.end(), mus);
  }
  delete ls;
  return rc;
}

int oracle_feasibility(const oracle_problem* p, const double* dz,
                       const double* dl, const double* dv, double tol) {
  return CheckFeasibility(p->data, dz, dl, dv, tol);
}

int oracle_solve(const oracle_problem* p, int variant, const oracle_options* o,
                 double* z, double* l, double* v, double* y, oracle_out* out,
                 double* traj, int traj_cap, int* traj_len) {
  const auto t0 = std::chrono::high_resolution_clock::now();
  LinearSolver* ls = MakeSolver(p, variant);
  Algorithm alg(p->data, ls, *o);
  alg.traj = TrajSink{traj, traj_cap, 0};
  out->eflag = 2;
  out->residual = 0.0;
  out->initial_residual = 0.0;
  out->status = 0;
  try {
    out->eflag = alg.Solve(z, l, v, y, &out->residual, &out->initial_residual);
  } catch (SaturateError&) {
    out->status = 2;
  } catch (int s) {
    out->status = s;
  }
  out->newton_iters = alg.newton_iters;
  out->prox_iters = alg.prox_iters;
  out->ls_backtracks = alg.ls_backtracks;
  out->residual_evals = alg.residual_evals;
  if (traj_len) *traj_len = alg.traj.len;
  delete ls;
  const auto t1 = std::chrono::high_resolution_clock::now();
  out->solve_time = std::chrono::duration<double>(t1 - t0).count();
  return out->status;
}

oracle_problem* oracle_sparse_create(int nz, int nl, int nv, const int* Hp, const int* Hi,
                                     const double* Hx, const double* f, const int* Gp,
                                     const int* Gi, const double* Gx, const double* h,
                                     const int* Ap, const int* Ai, const double* Ax,
                                     const double* b, const int* perm) {
  if (nz <= 0 || nv <= 0 || nl < 0) return nullptr;
  auto* p = new oracle_problem;
  p->sparse = new SparseData(nz, nl, nv, Hp, Hi, Hx, f, Gp, Gi, Gx, h, Ap, Ai, Ax, b);
  p->data = p->sparse;
  if (perm) p->perm.assign(perm, perm + nz + nl + nv);
  return p;
}

int oracle_qdldl_solve(int n, const int* Ap, const int* Ai, const double* Ax, double* x) {
  QdldlWrapper w(n, Ap, Ai);
  if (!w.Factor(Ap, Ai, Ax)) return -1;
  w.Solve(x);
  return 0;
}

int oracle_sparse_solve_batch(int nz, int nl, int nv, int batch, const int* Hp, const int* Hi,
                              const double* Hx, const double* f, const int* Gp, const int* Gi,
                              const double* Gx, const double* h, const int* Ap, const int* Ai,
                              const double* Ax, const double* b, const int* perm, double* z,
                              double* l, double* v, double* y, const oracle_options* o,
                              oracle_out* out, int nthreads) {
  if (nthreads < 1) nthreads = 1;
  const size_t nH = Hp[nz], nG = nl > 0 ? Gp[nz] : 0, nA = Ap[nz];
  auto work = [&](int lo, int hi) {
    for (int i = lo; i < hi; i++) {
      oracle_problem* p = oracle_sparse_create(
          nz, nl, nv, Hp, Hi, Hx + (size_t)i * nH, f + (size_t)i * nz, Gp, Gi,
          Gx + (size_t)i * nG, h + (size_t)i * nl, Ap, Ai, Ax + (size_t)i * nA,
          b + (size_t)i * nv, perm);
      oracle_solve(p, 0, o, z + (size_t)i * nz, l + (size_t)i * nl, v + (size_t)i * nv,
                   y + (size_t)i * nv, out + i, nullptr, 0, nullptr);
      oracle_destroy(p);
    }
  };
  std::vector<std::thread> th;
  const int per = (batch + nthreads - 1) / nthreads;
  for (int t = 0; t < nthreads; t++) {
    const int lo = t * per, hi = std::min(batch, lo + per);
    if (lo < hi) th.emplace_back(work, lo, hi);
  }
  for (auto& t : th) t.join();
  return 0;
}

int oracle_dense_solve_batch(int nz, int nl, int nv, int batch, const double* H,
                             const double* f, const double* G, const double* h,
                             const double* A, const double* b, double* z,
                             double* l, double* v, double* y,
                             const oracle_options* o, oracle_out* out,
                             int variant, int nthreads) {
  if (nthreads < 1) nthreads = 1;
  auto work = [&](int lo, int hi) {
    for (int i = lo; i < hi; i++) {
      oracle_problem* p = oracle_dense_create(
          nz, nl, nv, H + (size_t)i * nz * nz, f + (size_t)i * nz,
          G + (size_t)i * nl * nz, h + (size_t)i * nl, A + (size_t)i * nv * nz,
          b + (size_t)i * nv);
      oracle_solve(p, variant, o, z + (size_t)i * nz, l + (size_t)i * nl,
                   v + (size_t)i * nv, y + (size_t)i * nv, out + i, nullptr, 0,
                   nullptr);
      oracle_destroy(p);
    }
  };
  std::vector<std::thread> th;
  const int per = (batch + nthreads - 1) / nthreads;
  for (int t = 0; t < nthreads; t++) {
    const int lo = t * per, hi = std::min(batch, lo + per);
    if (lo < hi) th.emplace_back(work, lo, hi);
  }
  for (auto& t : th) t.join();
  return 0;
}

int oracle_mpc_solve_batch(int N, int nx, int nu, int nc, int batch,
                           const double* Q, const double* R, const double* S,
                           const double* q, const double* r, const double* A,
                           const double* B, const double* c, const double* E,
                           const double* L, const double* d, const double* x0,
                           double* z, double* l, double* v, double* y,
                           const oracle_options* o, oracle_out* out,
                           int nthreads) {
  if (nthreads < 1) nthreads = 1;
  const size_t nz = (size_t)(N + 1) * (nx + nu), nl = (size_t)(N + 1) * nx,
               nv = (size_t)(N + 1) * nc;
  const size_t sQ = (size_t)(N + 1) * nx * nx, sR = (size_t)(N + 1) * nu * nu,
               sS = (size_t)(N + 1) * nu * nx, sq = (size_t)(N + 1) * nx,
               sr = (size_t)(N + 1) * nu, sA = (size_t)N * nx * nx,
               sB = (size_t)N * nx * nu, sc = (size_t)N * nx,
               sE = (size_t)(N + 1) * nc * nx, sL = (size_t)(N + 1) * nc * nu,
               sd = (size_t)(N + 1) * nc;
  auto work = [&](int lo, int hi) {
    for (int i = lo; i < hi; i++) {
      oracle_problem* p = oracle_mpc_create(
          N, nx, nu, nc, Q + i * sQ, R + i * sR, S + i * sS, q + i * sq,
          r + i * sr, A + i * sA, B + i * sB, c + i * sc, E + i * sE,
          L + i * sL, d + i * sd, x0 + (size_t)i * nx);
      oracle_solve(p, 0, o, z + i * nz, l + i * nl, v + i * nv, y + i * nv,
                   out + i, nullptr, 0, nullptr);
      oracle_destroy(p);
    }
  };
  std::vector<std::thread> th;
  const int per = (batch + nthreads - 1) / nthreads;
  for (int t = 0; t < nthreads; t++) {
    const int lo = t * per, hi = std::min(batch, lo + per);
    if (lo < hi) th.emplace_back(work, lo, hi);
  }
  for (auto& t : th) t.join();
  return 0;
}

}  // extern "C"

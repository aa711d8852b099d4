// problems.cpp -- synthetic problem generators (host code, no CUDA).
//
// * The four optimal-control problems of the reference's test fixture,
//   restated as data: reference fbstab/test/ocp_generator.cc:73-169
//   (CopolymerizationReactor), :171-244 (SpacecraftRelativeMotion), :245-315
//   (ServoMotor), :319-363 (DoubleIntegrator), replicated over the horizon in
//   the reference's time-varying wire format with E(0)=0
//   (CopyOverHorizon, ocp_generator.cc:365-421).
// * The seeded random dense QP family of SURVEY.md section 8(d) (the reference
//   has no random generator: ocp_generator.h:131-132 is a TODO).
//
// All matrices are column-major, sequences are `len x rows x cols` contiguous
// (reference tools/matrix_sequence.h:81-83).

#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "fbstab_b200.h"

namespace {

// ----------------------------------------------------------------------------
// Counter-based RNG: splitmix64 seeding -> xoshiro256** stream, Box-Muller.
// ----------------------------------------------------------------------------
struct Rng {
  uint64_t s[4];
  bool have_spare = false;
  double spare = 0.0;
  static uint64_t splitmix(uint64_t& x) {
    uint64_t z = (x += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  Rng(uint64_t config, uint64_t instance) {
    uint64_t x = 0xFB57ABull;
    x = splitmix(x) ^ (config * 0xD1342543DE82EF95ull);
    x = splitmix(x) ^ (instance * 0xA0761D6478BD642Full);
    for (int i = 0; i < 4; i++) s[i] = splitmix(x);
  }
  static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
  uint64_t next() {
    const uint64_t result = rotl(s[1] * 5, 7) * 9;
    const uint64_t t = s[1] << 17;
    s[2] ^= s[0];
    s[3] ^= s[1];
    s[1] ^= s[2];
    s[0] ^= s[3];
    s[2] ^= t;
    s[3] = rotl(s[3], 45);
    return result;
  }
  double uniform() { return (double)(next() >> 11) * 0x1.0p-53; }          // [0,1)
  double uniform_pos() { return (double)((next() >> 11) + 1) * 0x1.0p-53; }  // (0,1]
  double normal() {
    if (have_spare) {
      have_spare = false;
      return spare;
    }
    const double u1 = uniform_pos();
    const double u2 = uniform();
    const double r = std::sqrt(-2.0 * std::log(u1));
    const double a = 6.283185307179586476925286766559 * u2;
    spare = r * std::sin(a);
    have_spare = true;
    return r * std::cos(a);
  }
};

// kind: 0 feasible, 1 primal infeasible, 2 unbounded (dual infeasible)
void RandomDenseQp(int config, long instance, int nz, int nl, int nv, int kind,
                   double* H, double* f, double* G, double* h, double* A,
                   double* b) {
  Rng rng((uint64_t)config, (uint64_t)instance);
  std::vector<double> M((size_t)nz * nz), zs(nz);
  for (auto& e : M) e = rng.normal();
  for (size_t e = 0; e < (size_t)nv * nz; e++) A[e] = rng.normal();
  for (size_t e = 0; e < (size_t)nl * nz; e++) G[e] = rng.normal();
  for (auto& e : zs) e = rng.normal();
  std::vector<double> s(nv);
  for (auto& e : s) e = rng.uniform();
  for (int i = 0; i < nz; i++) f[i] = rng.normal();

  if (kind == 2) {
    // Direction e_0 becomes a ray of unbounded descent: H e_0 = 0, G e_0 = 0,
    // A e_0 <= 0, f_0 < 0 (modelled on the reference's UnboundedQP,
    // fbstab_dense_unit_tests.cc:233-256).
    for (int j = 0; j < nz; j++) M[(size_t)j * nz + 0] = 0.0;
    for (int i = 0; i < nl; i++) G[i] = 0.0;
    for (int i = 0; i < nv; i++) A[i] = -std::fabs(A[i]);
    f[0] = -1.0;
  }
  // H = M M'/nz + 1e-2 I   (no regularisation for the unbounded kind)
  const double reg = (kind == 2) ? 0.0 : 1e-2;
  for (int j = 0; j < nz; j++)
    for (int i = j; i < nz; i++) {
      double acc = 0.0;
      for (int k = 0; k < nz; k++)
        acc += M[(size_t)k * nz + i] * M[(size_t)k * nz + j];
      acc /= nz;
      if (i == j) acc += reg;
      H[(size_t)j * nz + i] = acc;
      H[(size_t)i * nz + j] = acc;
    }
  if (kind == 2)
    for (int j = 1; j < nz; j++) H[(size_t)j * nz + j] += 1e-2;
  for (int i = 0; i < nl; i++) {
    double acc = 0.0;
    for (int k = 0; k < nz; k++) acc += G[(size_t)k * nl + i] * zs[k];
    h[i] = acc;
  }
  for (int i = 0; i < nv; i++) {
    double acc = 0.0;
    for (int k = 0; k < nz; k++) acc += A[(size_t)k * nv + i] * zs[k];
    b[i] = acc + s[i];
  }
  if (kind == 1 && nv >= 2) {
    // Row 1 = -Row 0 and b_1 = -b_0 - 1: a_0 z <= b_0 and a_0 z >= b_0 + 1
    // (modelled on the reference's InfeasibleQP, fbstab_dense_unit_tests.cc:195-217).
    for (int k = 0; k < nz; k++) A[(size_t)k * nv + 1] = -A[(size_t)k * nv + 0];
    b[1] = -b[0] - 1.0;
  }
}

// ----------------------------------------------------------------------------
// OCP models.  Small dense matrices in column-major storage.
// ----------------------------------------------------------------------------
struct Model {
  int nx = 0, nu = 0, nc = 0;
  std::vector<double> Q, R, S, q, r, A, B, c, E, L, d, x0;
  // simulation side (OcpGenerator::GetSimulationInputs, ocp_generator.cc:56-71):
  // output map y = C x (ny x nx, column-major; D = 0) and the number of steps T
  int ny = 0, T = 0;
  std::vector<double> C;
  void Output(int ny_, int T_) {
    ny = ny_;
    T = T_;
    C.assign((size_t)ny * nx, 0.0);
  }
  double& cm(int i, int j) { return C[(size_t)j * ny + i]; }
  void Alloc(int nx_, int nu_, int nc_) {
    nx = nx_;
    nu = nu_;
    nc = nc_;
    Q.assign((size_t)nx * nx, 0.0);
    R.assign((size_t)nu * nu, 0.0);
    S.assign((size_t)nu * nx, 0.0);
    q.assign(nx, 0.0);
    r.assign(nu, 0.0);
    A.assign((size_t)nx * nx, 0.0);
    B.assign((size_t)nx * nu, 0.0);
    c.assign(nx, 0.0);
    E.assign((size_t)nc * nx, 0.0);
    L.assign((size_t)nc * nu, 0.0);
    d.assign(nc, 0.0);
    x0.assign(nx, 0.0);
  }
  double& a(int i, int j) { return A[(size_t)j * nx + i]; }
  double& bm(int i, int j) { return B[(size_t)j * nx + i]; }
  double& qm(int i, int j) { return Q[(size_t)j * nx + i]; }
  double& rm(int i, int j) { return R[(size_t)j * nu + i]; }
  double& sm(int i, int j) { return S[(size_t)j * nu + i]; }
  double& e(int i, int j) { return E[(size_t)j * nc + i]; }
  double& l(int i, int j) { return L[(size_t)j * nc + i]; }
};

// ocp_generator.cc:319-363
void DoubleIntegrator(Model* m) {
  m->Alloc(2, 1, 6);
  m->qm(0, 0) = 2;
  m->qm(1, 1) = 1;
  m->sm(0, 0) = 1;
  m->rm(0, 0) = 3;
  m->q[0] = -2;
  m->a(0, 0) = 1;
  m->a(0, 1) = 1;
  m->a(1, 1) = 1;
  m->bm(1, 0) = 1;
  // E rows: [-1 0],[0 -1],[1 0],[0 1],[0 0],[0 0]
  m->e(0, 0) = -1;
  m->e(1, 1) = -1;
  m->e(2, 0) = 1;
  m->e(3, 1) = 1;
  m->l(4, 0) = -1;
  m->l(5, 0) = 1;
  const double dd[6] = {0, 0, -2, -2, -1, -1};
  for (int i = 0; i < 6; i++) m->d[i] = dd[i];
  m->Output(2, 40);  // ocp_generator.cc:357-362
  m->cm(0, 0) = 1;
  m->cm(1, 1) = 1;
}

// ocp_generator.cc:245-315
void ServoMotor(Model* m) {
  m->Alloc(4, 1, 4);
  const double kt = 10.0, bl = 25.0, Jm = 0.5, bm = 0.1, ktheta = 1280.2,
               RR = 20.0, rho = 20.0, Jl = 20 * Jm;
  const double umax = 220.0, ymax = 78.5358;
  double Ac[4][4] = {{0, 1, 0, 0},
                     {-ktheta / Jl, -bl / Jl, ktheta / (rho * Jl), 0},
                     {0, 0, 0, 1},
                     {ktheta / (rho * Jm), 0, -ktheta / (rho * rho * Jm),
                      -(bm + kt * kt / RR) / Jm}};
  double Bc[4] = {0, 0, 0, kt / (RR * Jm)};
  const double C1[4] = {ktheta, 0, -ktheta / rho, 0};  // second output row
  const double ts = 0.05;
  for (int i = 0; i < 4; i++) {
    for (int j = 0; j < 4; j++) m->a(i, j) = (i == j ? 1.0 : 0.0) + ts * Ac[i][j];
    m->bm(i, 0) = ts * Bc[i];
  }
  m->qm(0, 0) = 1000;
  m->rm(0, 0) = 1e-4;
  constexpr double pi = 3.1415926535897;  // as in the fixture
  const double xtrg[4] = {30 * pi / 180, 0, 0, 0};
  // q = -Q*xtrg, r = -R*utrg (utrg = 0)
  for (int i = 0; i < 4; i++) {
    double acc = 0.0;
    for (int j = 0; j < 4; j++) acc += m->qm(i, j) * xtrg[j];
    m->q[i] = -acc;
  }
  m->r[0] = -(m->rm(0, 0) * 0.0);
  for (int j = 0; j < 4; j++) {
    m->e(0, j) = C1[j];
    m->e(1, j) = -C1[j];
  }
  m->l(2, 0) = 1;
  m->l(3, 0) = -1;
  m->d[0] = -ymax;
  m->d[1] = -ymax;
  m->d[2] = -umax;
  m->d[3] = -umax;
  m->Output(2, 40);  // ocp_generator.cc:271-272,309-314
  m->cm(0, 0) = 1;
  for (int j = 0; j < 4; j++) m->cm(1, j) = C1[j];
}

// ocp_generator.cc:171-244
void SpacecraftRelativeMotion(Model* m) {
  m->Alloc(6, 3, 12);
  const double mu = 398600.4418, Re = 6371, alt = 650;
  const double n = std::sqrt(mu / std::pow(Re + alt, 3));
  double Ac[6][6] = {};
  for (int i = 0; i < 3; i++) Ac[i][3 + i] = 1.0;
  Ac[3][0] = 2 * n * n;  // A21 diag(2n^2, 0, -n^2)
  Ac[5][2] = -n * n;
  Ac[3][4] = 2 * n;  // A22 = [0 2n 0; -2n 0 0; 0 0 0]
  Ac[4][3] = -2 * n;
  const double ts = 30.0;
  double Ad[6][6], Bd[6][3] = {};
  for (int i = 0; i < 6; i++)
    for (int j = 0; j < 6; j++) Ad[i][j] = (i == j ? 1.0 : 0.0) + ts * Ac[i][j];
  for (int i = 0; i < 3; i++) Bd[3 + i][i] = ts * 1.0;
  // B = A*B (Delta-v input)
  for (int i = 0; i < 6; i++)
    for (int j = 0; j < 3; j++) {
      double acc = 0.0;
      for (int k = 0; k < 6; k++) acc += Ad[i][k] * Bd[k][j];
      m->bm(i, j) = acc;
    }
  for (int i = 0; i < 6; i++)
    for (int j = 0; j < 6; j++) m->a(i, j) = Ad[i][j];
  const double x0[6] = {-2.8, -0.01, -1, 0, 0, 0};
  for (int i = 0; i < 6; i++) m->x0[i] = x0[i];
  for (int i = 0; i < 3; i++) {
    m->qm(i, i) = 1.0;
    m->qm(3 + i, 3 + i) = 1e-3;
    m->rm(i, i) = 1.0;
  }
  const double umax = 1e-3, vmax = 1e-3;
  // E = [0(6x6); 0(3x3) I; 0(3x3) -I],  L = [I; -I; 0(6x3)]
  for (int i = 0; i < 3; i++) {
    m->e(6 + i, 3 + i) = 1.0;
    m->e(9 + i, 3 + i) = -1.0;
    m->l(i, i) = 1.0;
    m->l(3 + i, i) = -1.0;
  }
  for (int i = 0; i < 6; i++) {
    m->d[i] = -umax;
    m->d[6 + i] = -vmax;
  }
}

// ocp_generator.cc:73-169
void CopolymerizationReactor(Model* m) {
  m->Alloc(18, 5, 10);
  const int ai[26] = {1, 2, 3, 4, 5, 6, 7, 8, 7, 8, 9, 10, 11, 12, 13, 12, 13,
                      14, 15, 16, 15, 16, 17, 18, 17, 18};
  const int aj[26] = {1, 2, 3, 4, 5, 6, 7, 7, 8, 8, 9, 10, 11, 12, 12, 13, 13,
                      14, 15, 15, 16, 16, 17, 17, 18, 18};
  const double av[26] = {0.55531, 0.81264, 0.82131, 0.30408, 0.71811, 0.72276,
                         0.97319, 0.12353, -0.16471, 0.98966, 0.70834, 0.69048,
                         0.83152, -0.016569, 0.07277, -0.040608, 0.17835,
                         0.53526, -0.015422, 0.04805, -0.093847, 0.2924,
                         -0.22577, 0.43126, -0.38505, 0.2517};
  for (int k = 0; k < 26; k++) m->a(ai[k] - 1, aj[k] - 1) = av[k];
  const int bj[18] = {1, 1, 1, 2, 2, 2, 3, 3, 3, 3, 4, 5, 5, 5, 5, 5, 5, 5};
  const double bv[18] = {0.18899, 0.22577, 0.11347, 0.14614, 0.21282, 0.21347,
                         0.24707, 0.015512, 0.21145, 0.41785, 0.11415, 0.14554,
                         2.9448, 0.1859, 0.04805, 0.36229, 0.21563, 0.41905};
  for (int k = 0; k < 18; k++) m->bm(k, bj[k] - 1) = bv[k];
  double C[4][18] = {
      {0.8, 0, 0, 1, 0, 0, 0.0416666666666667, 0.333333333333333, 0, 0, 0,
       25.9553571428571, 1.80245535714286, 0, 0, 0, 0, 0},
      {0, -0.340248962655602, 0, 0, 0.874172185430464, 0, 0, 0,
       -0.413793103448276, 0, 0, 0, 0, -0.930000000000000, 0, 0, 0, 0},
      {0, 0, 0.47244, 0, 0, 0.63636, 0, 0, 0, -0.52593, -0.2952, 0, 0, 0, 0,
       -9.1992, 0, 0},
      {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1.6757, 1.8214}};
  for (int k = 0; k < 18; k++) m->x0[k] = 0.2 * std::sin((double)(k + 1));
  // Q = C'C
  for (int i = 0; i < 18; i++)
    for (int j = 0; j < 18; j++) {
      double acc = 0.0;
      for (int k = 0; k < 4; k++) acc += C[k][i] * C[k][j];
      m->qm(i, j) = acc;
    }
  for (int i = 0; i < 5; i++) m->rm(i, i) = 0.1;
  const double umax = 5.0 / 100.0;
  for (int i = 0; i < 5; i++) {
    m->l(i, i) = 1.0;
    m->l(5 + i, i) = -1.0;
  }
  for (int i = 0; i < 10; i++) m->d[i] = -umax;
  m->Output(4, 200);  // ocp_generator.cc:130-138,163-168
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 18; j++) m->cm(i, j) = C[i][j];
}

bool BuildModel(int kind, Model* m) {
  switch (kind) {
    case FBSTAB_OCP_DOUBLE_INTEGRATOR: DoubleIntegrator(m); return true;
    case FBSTAB_OCP_SERVO_MOTOR: ServoMotor(m); return true;
    case FBSTAB_OCP_SPACECRAFT: SpacecraftRelativeMotion(m); return true;
    case FBSTAB_OCP_COPOLYMERIZATION: CopolymerizationReactor(m); return true;
  }
  return false;
}

void Rep(double* dst, const std::vector<double>& src, int count) {
  for (int i = 0; i < count; i++)
    std::memcpy(dst + (size_t)i * src.size(), src.data(),
                src.size() * sizeof(double));
}

// ocp_generator.cc:365-421
void CopyOverHorizon(const Model& m, int N, double* Q, double* R, double* S,
                     double* q, double* r, double* A, double* B, double* c,
                     double* E, double* L, double* d, double* x0) {
  Rep(Q, m.Q, N + 1);
  Rep(R, m.R, N + 1);
  Rep(S, m.S, N + 1);
  Rep(q, m.q, N + 1);
  Rep(r, m.r, N + 1);
  Rep(A, m.A, N);
  Rep(B, m.B, N);
  Rep(c, m.c, N);
  Rep(E, m.E, N + 1);
  std::memset(E, 0, m.E.size() * sizeof(double));  // no constraint on x(0)
  Rep(L, m.L, N + 1);
  Rep(d, m.d, N + 1);
  std::memcpy(x0, m.x0.data(), m.x0.size() * sizeof(double));
}

}  // namespace

extern "C" {

int fbstab_ocp_dims(int kind, int* nx, int* nu, int* nc) {
  Model m;
  if (!BuildModel(kind, &m)) return FBSTAB_ERR_INVALID;
  *nx = m.nx;
  *nu = m.nu;
  *nc = m.nc;
  return FBSTAB_OK;
}

// OcpGenerator::GetSimulationInputs (ocp_generator.cc:56-71): the plant x+ = A x + B u,
// y = C x (+ D u, D = 0 in all four fixtures) and the number of steps T.  A, B, C are
// written column-major when non-null (nx*nx, nx*nu, ny*nx doubles); x0 likewise.
int fbstab_ocp_simulation(int kind, double* A, double* B, double* C, double* x0,
                          int* ny, int* T) {
  Model m;
  if (!BuildModel(kind, &m)) return FBSTAB_ERR_INVALID;
  if (m.ny == 0) {
    // SpacecraftRelativeMotion: C = I(6), T = 100 (ocp_generator.cc:195,238-243)
    m.Output(m.nx, 100);
    for (int i = 0; i < m.nx; i++) m.cm(i, i) = 1.0;
  }
  if (A) std::memcpy(A, m.A.data(), m.A.size() * sizeof(double));
  if (B) std::memcpy(B, m.B.data(), m.B.size() * sizeof(double));
  if (C) std::memcpy(C, m.C.data(), m.C.size() * sizeof(double));
  if (x0) std::memcpy(x0, m.x0.data(), m.x0.size() * sizeof(double));
  if (ny) *ny = m.ny;
  if (T) *T = m.T;
  return FBSTAB_OK;
}

int fbstab_ocp_generate(int kind, int N, double* Q, double* R, double* S,
                        double* q, double* r, double* A, double* B, double* c,
                        double* E, double* L, double* d, double* x0) {
  Model m;
  if (N < 1 || !BuildModel(kind, &m)) return FBSTAB_ERR_INVALID;
  CopyOverHorizon(m, N, Q, R, S, q, r, A, B, c, E, L, d, x0);
  return FBSTAB_OK;
}

int fbstab_ocp_generate_batch(int kind, int N, int config, long first,
                              int count, double rho, double* Q, double* R,
                              double* S, double* q, double* r, double* A,
                              double* B, double* c, double* E, double* L,
                              double* d, double* x0) {
  Model m;
  if (N < 1 || count < 0 || !BuildModel(kind, &m)) return FBSTAB_ERR_INVALID;
  const size_t nx = m.nx, nu = m.nu, nc = m.nc, K = N + 1;
  for (int i = 0; i < count; i++) {
    double* x0i = x0 + i * nx;
    CopyOverHorizon(m, N, Q + i * K * nx * nx, R + i * K * nu * nu,
                    S + i * K * nu * nx, q + i * K * nx, r + i * K * nu,
                    A + (size_t)i * N * nx * nx, B + (size_t)i * N * nx * nu,
                    c + (size_t)i * N * nx, E + i * K * nc * nx,
                    L + i * K * nc * nu, d + i * K * nc, x0i);
    const long inst = first + i;
    if (inst != 0 && rho != 0.0) {
      Rng rng((uint64_t)config, (uint64_t)inst);
      // rho < 0: one-sided perturbation |rho| U(0,1), for OCPs whose nominal
      // x0 sits on a constraint boundary (double integrator: x >= 0)
      for (size_t k = 0; k < nx; k++)
        x0i[k] += rho > 0.0 ? rho * (2.0 * rng.uniform() - 1.0) : -rho * rng.uniform();
    }
  }
  return FBSTAB_OK;
}

int fbstab_random_dense_qp(int config, long first, int count, int nz, int nl,
                           int nv, int kind, double* H, double* f, double* G,
                           double* h, double* A, double* b, int nthreads) {
  if (nz <= 0 || nl < 0 || nv <= 0 || count < 0 || kind < 0 || kind > 2)
    return FBSTAB_ERR_INVALID;
  if (nthreads < 1) nthreads = 1;
  auto work = [&](int lo, int hi) {
    for (int i = lo; i < hi; i++)
      RandomDenseQp(config, first + i, nz, nl, nv, kind,
                    H + (size_t)i * nz * nz, f + (size_t)i * nz,
                    G + (size_t)i * nl * nz, h + (size_t)i * nl,
                    A + (size_t)i * nv * nz, b + (size_t)i * nv);
  };
  std::vector<std::thread> th;
  const int per = (count + nthreads - 1) / nthreads;
  for (int t = 0; t < nthreads; t++) {
    const int lo = t * per, hi = std::min(count, lo + per);
    if (lo < hi) th.emplace_back(work, lo, hi);
  }
  for (auto& t : th) t.join();
  return FBSTAB_OK;
}

}  // extern "C"

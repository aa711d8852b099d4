// fbstab_mpc.h -- FBstabMpc: the reference's MPC-structured QP solver
// interface over the batched B200 engine.
//
//     min.  sum_{i=0}^N 1/2 [x(i)]' [Q(i) S(i)'] [x(i)] + [q(i)]'[x(i)]
//                           [u(i)]  [S(i) R(i) ] [u(i)]   [r(i)] [u(i)]
//     s.t.  x(i+1) = A(i)x(i) + B(i)u(i) + c(i), i = 0 ... N-1
//           x(0) = x0
//           E(i)x(i) + L(i)u(i) + d(i) <= 0,     i = 0 ... N
//
// Public surface = the reference's (fbstab/fbstab_mpc.h:56-243,
// fbstab/fbstab_mpc.cc:18-114): FBstabMpc(N,nx,nu,nc), ProblemData(Ref),
// Variable(Ref), Options, Solve(qp,&x[,os]), UpdateOptions, DefaultOptions,
// ReliableOptions -- same argument meaning and std::runtime_error behaviour
// (sizes validated like MpcData::ValidateInputs, components/mpc_data.cc:291-363)
// -- plus the batched entry SolveBatch and the QPData / QPVariable aliases.
#pragma once

#include <memory>
#include <new>
#include <stdexcept>
#include <vector>

#include "fbstab/fbstab_algorithm.h"
#include "fbstab/linalg.h"
#include "fbstab/matrix_sequence.h"
#include "fbstab_b200.h"

namespace fbstab {

class FBstabMpc {
 public:
  FBstabMpc(const FBstabMpc&) = delete;
  FBstabMpc& operator=(const FBstabMpc&) = delete;

  /** Owning problem data (reference fbstab_mpc.h:67-83). */
  struct ProblemData {
    ProblemData() = default;
    MatrixSequence Q;    ///< N + 1 sequence of nx x nx matrices
    MatrixSequence R;    ///< N + 1 sequence of nu x nu matrices
    MatrixSequence S;    ///< N + 1 sequence of nu x nx matrices
    MatrixSequence q;    ///< N + 1 sequence of nx x 1  matrices
    MatrixSequence r;    ///< N + 1 sequence of nu x 1  matrices
    MatrixSequence A;    ///< N     sequence of nx x nx matrices
    MatrixSequence B;    ///< N     sequence of nx x nu matrices
    MatrixSequence c;    ///< N     sequence of nx x 1  matrices
    MatrixSequence E;    ///< N + 1 sequence of nc x nx matrices
    MatrixSequence L;    ///< N + 1 sequence of nc x nu matrices
    MatrixSequence d;    ///< N + 1 sequence of nc x 1  matrices
    Eigen::VectorXd x0;  ///< nx x 1 vector
  };

  /** Non-owning problem data (reference fbstab_mpc.h:92-120). */
  struct ProblemDataRef {
    ProblemDataRef() : x0(nullptr, 0) {}

    template <class Vector>
    void SetX0(const Vector& x0_) {
      new (&x0) Eigen::Map<const Eigen::VectorXd>(x0_.data(), x0_.size());
    }

    ProblemDataRef(const MatrixSequence* Q_, const MatrixSequence* R_, const MatrixSequence* S_,
                   const MatrixSequence* q_, const MatrixSequence* r_, const MatrixSequence* A_,
                   const MatrixSequence* B_, const MatrixSequence* c_, const MatrixSequence* E_,
                   const MatrixSequence* L_, const MatrixSequence* d_,
                   const Eigen::VectorXd* x0_)
        : Q(*Q_), R(*R_), S(*S_), q(*q_), r(*r_), A(*A_), B(*B_), c(*c_), E(*E_), L(*L_),
          d(*d_), x0(x0_->data(), x0_->size()) {}

    MapMatrixSequence Q, R, S, q, r, A, B, c, E, L, d;
    Eigen::Map<const Eigen::VectorXd> x0;
  };

  /** Initial guess in, solution out (reference fbstab_mpc.h:126-137). */
  struct Variable {
    Variable(int N, int nx, int nu, int nc)
        : z((N + 1) * (nx + nu)), l((N + 1) * nx), v((N + 1) * nc), y((N + 1) * nc) {
      z.setZero();
      l.setZero();
      v.setZero();
      y.setZero();
    }
    /** s = (N, nx, nu, nc) */
    explicit Variable(const Eigen::Vector4d& s)
        : Variable((int)s(0), (int)s(1), (int)s(2), (int)s(3)) {}
    Eigen::VectorXd z;  ///< decision variables (x0,u0,...,xN,uN)
    Eigen::VectorXd l;  ///< co-states
    Eigen::VectorXd v;  ///< inequality duals
    Eigen::VectorXd y;  ///< constraint margin b - Az
  };

  /** Variable over existing memory (reference fbstab_mpc.h:140-150). */
  struct VariableRef {
    VariableRef(Eigen::Map<Eigen::VectorXd> z_, Eigen::Map<Eigen::VectorXd> l_,
                Eigen::Map<Eigen::VectorXd> v_, Eigen::Map<Eigen::VectorXd> y_)
        : z(z_), l(l_), v(v_), y(y_) {}
    void fill(double a) {
      z.fill(a);
      l.fill(a);
      v.fill(a);
      y.fill(a);
    }
    Eigen::Map<Eigen::VectorXd> z, l, v, y;
  };

  using QPData = ProblemData;
  using QPVariable = Variable;

  struct Options : public AlgorithmParameters {};

  /**
   * Allocates the device workspaces.  Throws std::runtime_error if any size
   * is non-positive (reference fbstab_mpc.cc:61-66) or no CUDA device is usable.
   */
  explicit FBstabMpc(int N, int nx, int nu, int nc, int max_batch = 1, int device = 0)
      : N_(N), nx_(nx), nu_(nu), nc_(nc) {
    if (N < 1 || nx < 1 || nu < 1 || nc < 1)
      throw std::runtime_error("In FBstabMpc::FBstabMpc: problem sizes must be positive.");
    nz_ = (N + 1) * (nx + nu);
    nl_ = (N + 1) * nx;
    nv_ = (N + 1) * nc;
    fbstab_mpc_batch* h = nullptr;
    detail::Check(fbstab_mpc_batch_create(N, nx, nu, nc, max_batch, device, &h),
                  "FBstabMpc::FBstabMpc");
    handle_.reset(h);
    opts_.DefaultParameters();
  }

  /** s = (N, nx, nu, nc), reference fbstab_mpc.h:168. */
  explicit FBstabMpc(const Eigen::Vector4d& s)
      : FBstabMpc((int)s(0), (int)s(1), (int)s(2), (int)s(3)) {}

  /** Solves one instance (reference fbstab_mpc.h:181-195). */
  template <class InputData, class InputVariable, class OutStream>
  SolverOut Solve(const InputData& qp, InputVariable* x, const OutStream& os) {
    ValidateData(qp);
    if (x->z.size() != nz_ || x->l.size() != nl_ || x->v.size() != nv_ || x->y.size() != nv_)
      throw std::runtime_error(
          "In FBstabMpc::Solve: mismatch between *this and initial guess dimensions.");
    fbstab_out out;
    detail::Check(fbstab_mpc_batch_solve(handle_.get(), 1, qp.Q.data(), qp.R.data(),
                                         qp.S.data(), qp.q.data(), qp.r.data(), qp.A.data(),
                                         qp.B.data(), qp.c.data(), qp.E.data(), qp.L.data(),
                                         qp.d.data(), qp.x0.data(), x->z.data(), x->l.data(),
                                         x->v.data(), x->y.data(), &out, nullptr),
                  "FBstabMpc::Solve");
    SolverOut s = detail::FromC(out);
    detail::ThrowOnStatus(s);
    detail::PrintFinal(opts_, s, os);
    return s;
  }

  template <class InputData, class InputVariable>
  SolverOut Solve(const InputData& qp, InputVariable* x) {
    StandardOutput os;
    return Solve(qp, x, os);
  }

  /**
   * Solves `batch` independent OCPs in one call.  Every array is
   * instance-major and contiguous; within an instance each sequence is laid
   * out like MatrixSequence.  Pointers may be host or device memory.
   */
  std::vector<SolverOut> SolveBatch(int batch, const double* Q, const double* R,
                                    const double* S, const double* q, const double* r,
                                    const double* A, const double* B, const double* c,
                                    const double* E, const double* L, const double* d,
                                    const double* x0, double* z, double* l, double* v,
                                    double* y, void* stream = nullptr) {
    std::vector<fbstab_out> out((size_t)(batch > 0 ? batch : 0));
    detail::Check(fbstab_mpc_batch_solve(handle_.get(), batch, Q, R, S, q, r, A, B, c, E, L, d,
                                         x0, z, l, v, y, out.data(), stream),
                  "FBstabMpc::SolveBatch");
    std::vector<SolverOut> res;
    res.reserve(out.size());
    for (const fbstab_out& o : out) res.push_back(detail::FromC(o));
    return res;
  }

  /**
   * One plant, `batch` initial states: `qp` is ONE problem (its x0 is ignored),
   * x0 holds batch * nx doubles, instance-major.  The stage data crosses the
   * boundary once instead of `batch` times (fbstab_mpc_batch_solve_shared) and
   * the results equal SolveBatch on the replicated data bit for bit.
   */
  template <class InputData>
  std::vector<SolverOut> SolveBatchShared(const InputData& qp, int batch, const double* x0,
                                          double* z, double* l, double* v, double* y,
                                          void* stream = nullptr) {
    ValidateData(qp);
    std::vector<fbstab_out> out((size_t)(batch > 0 ? batch : 0));
    detail::Check(
        fbstab_mpc_batch_solve_shared(handle_.get(), batch, qp.Q.data(), qp.R.data(), qp.S.data(),
                                      qp.q.data(), qp.r.data(), qp.A.data(), qp.B.data(),
                                      qp.c.data(), qp.E.data(), qp.L.data(), qp.d.data(), x0, z,
                                      l, v, y, out.data(), stream),
        "FBstabMpc::SolveBatchShared");
    std::vector<SolverOut> res;
    res.reserve(out.size());
    for (const fbstab_out& o : out) res.push_back(detail::FromC(o));
    return res;
  }

  /**
   * Time-invariant problem from ONE stage of each matrix (column-major: Q nx*nx,
   * R nu*nu, S nu*nx, q nx, r nu, A nx*nx, B nx*nu, c nx, E nc*nx, L nc*nu, d nc),
   * replicated over the horizon like the reference's OcpGenerator::CopyOverHorizon
   * (E(0) = 0), for `batch` initial states (fbstab_mpc_batch_solve_lti).
   */
  std::vector<SolverOut> SolveBatchLti(int batch, const double* Q, const double* R,
                                       const double* S, const double* q, const double* r,
                                       const double* A, const double* B, const double* c,
                                       const double* E, const double* L, const double* d,
                                       const double* x0, double* z, double* l, double* v,
                                       double* y, void* stream = nullptr) {
    std::vector<fbstab_out> out((size_t)(batch > 0 ? batch : 0));
    detail::Check(fbstab_mpc_batch_solve_lti(handle_.get(), batch, Q, R, S, q, r, A, B, c, E, L,
                                             d, x0, z, l, v, y, out.data(), stream),
                  "FBstabMpc::SolveBatchLti");
    std::vector<SolverOut> res;
    res.reserve(out.size());
    for (const fbstab_out& o : out) res.push_back(detail::FromC(o));
    return res;
  }

  /** The wire-format batch on several GPUs (HOST pointers; see FBstabDense). */
  std::vector<SolverOut> SolveBatch(int batch, const double* Q, const double* R,
                                    const double* S, const double* q, const double* r,
                                    const double* A, const double* B, const double* c,
                                    const double* E, const double* L, const double* d,
                                    const double* x0, double* z, double* l, double* v,
                                    double* y, const std::vector<int>& devices) {
    fbstab_mpc_multi_gpu* m = nullptr;
    detail::Check(fbstab_mpc_multi_gpu_create((int)devices.size(), devices.data(), N_, nx_, nu_,
                                              nc_, batch > 0 ? batch : 1, &m),
                  "FBstabMpc::SolveBatch");
    std::unique_ptr<fbstab_mpc_multi_gpu, DestroyMulti> guard(m);
    fbstab_options o = opts_.ToC();
    detail::Check(fbstab_mpc_multi_gpu_set_options(m, &o), "FBstabMpc::SolveBatch");
    std::vector<fbstab_out> out((size_t)(batch > 0 ? batch : 0));
    detail::Check(fbstab_mpc_multi_gpu_solve(m, batch, Q, R, S, q, r, A, B, c, E, L, d, x0, z, l,
                                             v, y, out.data()),
                  "FBstabMpc::SolveBatch");
    std::vector<SolverOut> res;
    res.reserve(out.size());
    for (const fbstab_out& oo : out) res.push_back(detail::FromC(oo));
    return res;
  }

  /** Batched solve over arrays of the single-instance structs. */
  template <class InputData, class InputVariable>
  std::vector<SolverOut> SolveBatch(const std::vector<InputData>& qps,
                                    std::vector<InputVariable>* xs) {
    const size_t B = qps.size();
    if (xs->size() != B)
      throw std::runtime_error("In FBstabMpc::SolveBatch: qps and xs differ in length.");
    const size_t K = N_ + 1, N = N_, nx = nx_, nu = nu_, nc = nc_;
    const size_t sz[12] = {K * nx * nx, K * nu * nu, K * nu * nx, K * nx,      K * nu, N * nx * nx,
                           N * nx * nu, N * nx,      K * nc * nx, K * nc * nu, K * nc, nx};
    std::vector<double> buf[12];
    for (int k = 0; k < 12; k++) buf[k].resize(B * sz[k]);
    std::vector<double> z(B * nz_), l(B * nl_), v(B * nv_), y(B * nv_);
    for (size_t i = 0; i < B; i++) {
      const InputData& p = qps[i];
      ValidateData(p);
      const double* src[12] = {p.Q.data(), p.R.data(), p.S.data(), p.q.data(),
                               p.r.data(), p.A.data(), p.B.data(), p.c.data(),
                               p.E.data(), p.L.data(), p.d.data(), p.x0.data()};
      for (int k = 0; k < 12; k++) Copy(src[k], buf[k].data() + i * sz[k], sz[k]);
      const InputVariable& x = (*xs)[i];
      if (x.z.size() != nz_ || x.l.size() != nl_ || x.v.size() != nv_ || x.y.size() != nv_)
        throw std::runtime_error(
            "In FBstabMpc::Solve: mismatch between *this and initial guess dimensions.");
      Copy(x.z.data(), &z[i * nz_], nz_);
      Copy(x.l.data(), &l[i * nl_], nl_);
      Copy(x.v.data(), &v[i * nv_], nv_);
    }
    std::vector<SolverOut> res = SolveBatch(
        (int)B, buf[0].data(), buf[1].data(), buf[2].data(), buf[3].data(), buf[4].data(),
        buf[5].data(), buf[6].data(), buf[7].data(), buf[8].data(), buf[9].data(),
        buf[10].data(), buf[11].data(), z.data(), l.data(), v.data(), y.data());
    for (size_t i = 0; i < B; i++) {
      InputVariable& x = (*xs)[i];
      Copy(&z[i * nz_], x.z.data(), nz_);
      Copy(&l[i * nl_], x.l.data(), nl_);
      Copy(&v[i * nv_], x.v.data(), nv_);
      Copy(&y[i * nv_], x.y.data(), nv_);
    }
    return res;
  }

  void UpdateOptions(const Options& options) {
    fbstab_options o = options.ToC();
    detail::Check(fbstab_mpc_batch_set_options(handle_.get(), &o), "FBstabMpc::UpdateOptions");
    detail::Check(fbstab_mpc_batch_get_options(handle_.get(), &o), "FBstabMpc::UpdateOptions");
    opts_.FromC(o);
  }

  static Options DefaultOptions() {
    Options o;
    o.DefaultParameters();
    return o;
  }
  static Options ReliableOptions() {
    Options o;
    o.ReliableParameters();
    return o;
  }

  const Options& options() const { return opts_; }
  const char* Path() const { return fbstab_mpc_batch_path(handle_.get()); }

 private:
  struct Destroy {
    void operator()(fbstab_mpc_batch* h) const { fbstab_mpc_batch_destroy(h); }
  };
  struct DestroyMulti {
    void operator()(fbstab_mpc_multi_gpu* h) const { fbstab_mpc_multi_gpu_destroy(h); }
  };
  static void Copy(const double* src, double* dst, size_t n) {
    for (size_t i = 0; i < n; i++) dst[i] = src[i];
  }

  // MpcData::ValidateInputs (components/mpc_data.cc:291-363) followed by
  // FBstabMpc::ValidateInputSizes (fbstab_mpc.h:229-242).
  template <class InputData>
  void ValidateData(const InputData& p) const {
    const int K = p.Q.length();
    if (K <= 0) throw std::runtime_error("Horizon length must be at least 1.");
    bool ok = K == p.R.length() && K == p.S.length() && K == p.q.length() &&
              K == p.r.length() && (K - 1) == p.A.length() && (K - 1) == p.B.length() &&
              (K - 1) == p.c.length() && K == p.E.length() && K == p.L.length() &&
              K == p.d.length();
    if (!ok) throw std::runtime_error("Sequence length mismatch in input data to MpcData.");
    const int nx = p.Q.rows();
    auto bad = [](const char* what) {
      throw std::runtime_error(std::string("Size mismatch in ") + what +
                               " input to MpcData.");
    };
    if (p.x0.size() != nx) bad("x0");
    if (p.Q.cols() != nx) bad("Q");
    if (p.S.cols() != nx) bad("S");
    if (p.q.rows() != nx) bad("q");
    if (p.E.cols() != nx) bad("E");
    if (p.A.rows() != nx || p.A.cols() != nx) bad("A");
    if (p.B.rows() != nx) bad("B");
    if (p.c.rows() != nx) bad("c");
    const int nu = p.R.rows();
    if (p.R.cols() != nu) bad("R");
    if (p.S.rows() != nu) bad("S");
    if (p.r.rows() != nu) bad("r");
    if (p.L.cols() != nu) bad("L");
    if (p.B.cols() != nu) bad("B");
    const int nc = p.E.rows();
    if (p.L.rows() != nc) bad("L");
    if (p.d.rows() != nc) bad("d");
    if (K - 1 != N_ || nx != nx_ || nu != nu_ || nc != nc_)
      throw std::runtime_error(
          "In FBstabMpc::Solve: mismatch between *this and data dimensions.");
  }

  int N_ = 0, nx_ = 0, nu_ = 0, nc_ = 0;
  int nz_ = 0, nl_ = 0, nv_ = 0;
  Options opts_;
  std::unique_ptr<fbstab_mpc_batch, Destroy> handle_;
};

}  // namespace fbstab

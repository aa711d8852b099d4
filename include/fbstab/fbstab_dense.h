// fbstab_dense.h -- FBstabDense: the reference's dense QP solver interface
// over the batched B200 engine.
//
//     min.  1/2 z'Hz + f'z   s.t.  Gz = h,  Az <= b
//
// Public surface = the reference's (fbstab/fbstab_dense.h:50-194,
// fbstab/fbstab_dense.cc:18-110): FBstabDense(nz,nl,nv), ProblemData(Ref),
// Variable(Ref), Options, Solve(qp,&x[,os]), UpdateOptions, DefaultOptions,
// ReliableOptions -- same argument meaning, same std::runtime_error on bad
// sizes -- plus the batched entry SolveBatch and the QPData / QPVariable
// aliases of the older API generation.  Everything below this header is the
// C-ABI of fbstab_b200.h; there is no CPU solver behind it.
#pragma once

#include <memory>
#include <stdexcept>
#include <vector>

#include "fbstab/fbstab_algorithm.h"
#include "fbstab/linalg.h"
#include "fbstab_b200.h"

namespace fbstab {

class FBstabDense {
 public:
  FBstabDense(const FBstabDense&) = delete;
  FBstabDense& operator=(const FBstabDense&) = delete;

  /** Owning problem data (reference fbstab_dense.h:55-64). */
  struct ProblemData {
    ProblemData() = default;
    ProblemData(int nz, int nl, int nv) : H(nz, nz), G(nl, nz), A(nv, nz), f(nz), h(nl), b(nv) {}
    Eigen::MatrixXd H;  ///< nz x nz positive semidefinite Hessian
    Eigen::MatrixXd G;  ///< nl x nz equality Jacobian
    Eigen::MatrixXd A;  ///< nv x nz inequality Jacobian
    Eigen::VectorXd f;  ///< nz linear cost
    Eigen::VectorXd h;  ///< nl equality rhs
    Eigen::VectorXd b;  ///< nv inequality rhs
  };

  /** Problem data over preallocated memory (reference fbstab_dense.h:67-82). */
  struct ProblemDataRef {
    ProblemDataRef() = delete;
    ProblemDataRef(const Eigen::Map<Eigen::MatrixXd>* H_, const Eigen::Map<Eigen::VectorXd>* f_,
                   const Eigen::Map<Eigen::MatrixXd>* G_, const Eigen::Map<Eigen::VectorXd>* h_,
                   const Eigen::Map<Eigen::MatrixXd>* A_, const Eigen::Map<Eigen::VectorXd>* b_)
        : H(H_->data(), H_->rows(), H_->cols()),
          G(G_->data(), G_->rows(), G_->cols()),
          A(A_->data(), A_->rows(), A_->cols()),
          f(f_->data(), f_->size()),
          h(h_->data(), h_->size()),
          b(b_->data(), b_->size()) {}
    Eigen::Map<const Eigen::MatrixXd> H;
    Eigen::Map<const Eigen::MatrixXd> G;
    Eigen::Map<const Eigen::MatrixXd> A;
    Eigen::Map<const Eigen::VectorXd> f;
    Eigen::Map<const Eigen::VectorXd> h;
    Eigen::Map<const Eigen::VectorXd> b;
  };

  /** Initial guess in, solution out (reference fbstab_dense.h:85-92). */
  struct Variable {
    Variable(int nz, int nl, int nv) : z(nz), l(nl), v(nv), y(nv) {
      z.setZero();
      l.setZero();
      v.setZero();
      y.setZero();
    }
    Eigen::VectorXd z;  ///< decision variables
    Eigen::VectorXd l;  ///< equality duals
    Eigen::VectorXd v;  ///< inequality duals
    Eigen::VectorXd y;  ///< constraint margin b - Az
  };

  /** Variable over preallocated memory (reference fbstab_dense.h:95-107). */
  struct VariableRef {
    VariableRef() = delete;
    VariableRef(Eigen::Map<Eigen::VectorXd>* z_, Eigen::Map<Eigen::VectorXd>* l_,
                Eigen::Map<Eigen::VectorXd>* v_, Eigen::Map<Eigen::VectorXd>* y_)
        : z(z_->data(), z_->size()),
          l(l_->data(), l_->size()),
          v(v_->data(), v_->size()),
          y(y_->data(), y_->size()) {}
    void fill(double a) {
      z.fill(a);
      l.fill(a);
      v.fill(a);
      y.fill(a);
    }
    Eigen::Map<Eigen::VectorXd> z;
    Eigen::Map<Eigen::VectorXd> l;
    Eigen::Map<Eigen::VectorXd> v;
    Eigen::Map<Eigen::VectorXd> y;
  };

  // Names of the older API generation (BASELINE north_star).
  using QPData = ProblemData;
  using QPVariable = Variable;

  struct Options : public AlgorithmParameters {};

  /**
   * Allocates the device workspaces for problems of size (nz, nl, nv).
   * Throws std::runtime_error unless nz > 0, nv > 0, nl >= 0
   * (reference fbstab_dense.cc:18-27), or if no CUDA device is usable.
   *
   * @param max_batch  largest batch SolveBatch will be called with
   * @param device     CUDA device ordinal
   */
  explicit FBstabDense(int nz, int nl, int nv, int max_batch = 1, int device = 0)
      : nz_(nz), nl_(nl), nv_(nv), max_batch_(max_batch) {
    if (nz <= 0 || nl < 0 || nv <= 0)
      throw std::runtime_error("In FBstabDense::FBstabDense: Inputs must be positive.");
    fbstab_dense_batch* h = nullptr;
    detail::Check(fbstab_dense_batch_create(nz, nl, nv, max_batch, device, &h),
                  "FBstabDense::FBstabDense");
    handle_.reset(h);
    opts_.DefaultParameters();
  }

  /**
   * Solves one instance.  x is the initial guess and is overwritten with the
   * solution (or with the infeasibility certificate, as in the reference).
   * InputData: ProblemData or ProblemDataRef; InputVariable: Variable or
   * VariableRef (reference fbstab_dense.h:136-149).
   */
  template <class InputData, class InputVariable, class OutStream>
  SolverOut Solve(const InputData& qp, InputVariable* x, const OutStream& os) {
    ValidateData(qp.H.rows(), qp.H.cols(), qp.f.size(), qp.G.rows(), qp.G.cols(),
                 qp.h.size(), qp.A.rows(), qp.A.cols(), qp.b.size());
    // (y is out-only, but the engine writes nv doubles through x->y.data(): a
    // short y must be an error here, as in FBstabMpc, not a heap overflow)
    if (nz_ != x->z.size() || x->l.size() != nl_ || nv_ != x->v.size() || nv_ != x->y.size())
      throw std::runtime_error(
          "In FBstabDense::Solve: mismatch between *this and initial guess dimensions.");
    fbstab_out out;
    detail::Check(
        fbstab_dense_batch_solve(handle_.get(), 1, qp.H.data(), qp.f.data(), qp.G.data(),
                                 qp.h.data(), qp.A.data(), qp.b.data(), x->z.data(),
                                 x->l.data(), x->v.data(), x->y.data(), &out, nullptr),
        "FBstabDense::Solve");
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
   * Solves `batch` independent instances in one call.  All arrays are
   * instance-major and contiguous (instance i's H starts at H + i*nz*nz, ...)
   * with each matrix column-major; pointers may be host or device memory
   * (see fbstab_b200.h).  z, l, v: initial guesses in, solutions out; y out.
   * Per-instance failures are reported in SolverOut::status, not thrown.
   */
  std::vector<SolverOut> SolveBatch(int batch, const double* H, const double* f,
                                    const double* G, const double* h, const double* A,
                                    const double* b, double* z, double* l, double* v,
                                    double* y, void* stream = nullptr) {
    std::vector<fbstab_out> out((size_t)(batch > 0 ? batch : 0));
    detail::Check(fbstab_dense_batch_solve(handle_.get(), batch, H, f, G, h, A, b, z, l, v, y,
                                           out.data(), stream),
                  "FBstabDense::SolveBatch");
    std::vector<SolverOut> res;
    res.reserve(out.size());
    for (const fbstab_out& o : out) res.push_back(detail::FromC(o));
    return res;
  }

  /**
   * The same batch on several GPUs of this box: the instances are cut into
   * contiguous ranges, one per device, and solved concurrently (one host thread
   * and one engine handle per device: fbstab_dense_multi_gpu_solve).  HOST
   * pointers only; results land in the caller's arrays, identical byte for byte
   * to the single-GPU call.  `devices`: CUDA ordinals, no duplicates.
   */
  std::vector<SolverOut> SolveBatch(int batch, const double* H, const double* f,
                                    const double* G, const double* h, const double* A,
                                    const double* b, double* z, double* l, double* v,
                                    double* y, const std::vector<int>& devices) {
    if (!multi_ || multi_devices_ != devices || multi_cap_ < batch) {
      fbstab_dense_multi_gpu* m = nullptr;
      detail::Check(fbstab_dense_multi_gpu_create((int)devices.size(), devices.data(), nz_, nl_,
                                                  nv_, batch > max_batch_ ? batch : max_batch_,
                                                  &m),
                    "FBstabDense::SolveBatch");
      multi_.reset(m);
      multi_devices_ = devices;
      multi_cap_ = batch > max_batch_ ? batch : max_batch_;
      fbstab_options o = opts_.ToC();
      detail::Check(fbstab_dense_multi_gpu_set_options(m, &o), "FBstabDense::SolveBatch");
    }
    std::vector<fbstab_out> out((size_t)(batch > 0 ? batch : 0));
    detail::Check(fbstab_dense_multi_gpu_solve(multi_.get(), batch, H, f, G, h, A, b, z, l, v,
                                               y, out.data()),
                  "FBstabDense::SolveBatch");
    std::vector<SolverOut> res;
    res.reserve(out.size());
    for (const fbstab_out& o : out) res.push_back(detail::FromC(o));
    return res;
  }

  /** Batched solve over arrays of the single-instance structs. */
  template <class InputData, class InputVariable>
  std::vector<SolverOut> SolveBatch(const std::vector<InputData>& qps,
                                    std::vector<InputVariable>* xs) {
    const size_t B = qps.size();
    if (xs->size() != B)
      throw std::runtime_error("In FBstabDense::SolveBatch: qps and xs differ in length.");
    const size_t nz = nz_, nl = nl_, nv = nv_;
    std::vector<double> H(B * nz * nz), f(B * nz), G(B * nl * nz), h(B * nl), A(B * nv * nz),
        b(B * nv), z(B * nz), l(B * nl), v(B * nv), y(B * nv);
    for (size_t i = 0; i < B; i++) {
      const InputData& q = qps[i];
      ValidateData(q.H.rows(), q.H.cols(), q.f.size(), q.G.rows(), q.G.cols(), q.h.size(),
                   q.A.rows(), q.A.cols(), q.b.size());
      const InputVariable& x = (*xs)[i];
      if (nz_ != x.z.size() || x.l.size() != nl_ || nv_ != x.v.size() || nv_ != x.y.size())
        throw std::runtime_error(
            "In FBstabDense::Solve: mismatch between *this and initial guess dimensions.");
      Copy(q.H.data(), &H[i * nz * nz], nz * nz);
      Copy(q.f.data(), &f[i * nz], nz);
      Copy(q.G.data(), G.data() + i * nl * nz, nl * nz);
      Copy(q.h.data(), h.data() + i * nl, nl);
      Copy(q.A.data(), &A[i * nv * nz], nv * nz);
      Copy(q.b.data(), &b[i * nv], nv);
      Copy(x.z.data(), &z[i * nz], nz);
      Copy(x.l.data(), l.data() + i * nl, nl);
      Copy(x.v.data(), &v[i * nv], nv);
    }
    std::vector<SolverOut> res = SolveBatch((int)B, H.data(), f.data(), G.data(), h.data(),
                                            A.data(), b.data(), z.data(), l.data(), v.data(),
                                            y.data());
    for (size_t i = 0; i < B; i++) {
      InputVariable& x = (*xs)[i];
      Copy(&z[i * nz], x.z.data(), nz);
      Copy(l.data() + i * nl, x.l.data(), nl);
      Copy(&v[i * nv], x.v.data(), nv);
      Copy(&y[i * nv], x.y.data(), nv);
    }
    return res;
  }

  /** Sets solver options; fields are clamped like the reference's
   *  ValidateOptions (fbstab_dense.cc:44-48 -> fbstab_algorithm-impl.h:308-332). */
  void UpdateOptions(const Options& options) {
    fbstab_options o = options.ToC();
    detail::Check(fbstab_dense_batch_set_options(handle_.get(), &o),
                  "FBstabDense::UpdateOptions");
    detail::Check(fbstab_dense_batch_get_options(handle_.get(), &o),
                  "FBstabDense::UpdateOptions");
    opts_.FromC(o);
    if (multi_)
      detail::Check(fbstab_dense_multi_gpu_set_options(multi_.get(), &o),
                    "FBstabDense::UpdateOptions");
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
  /** Name of the device code path selected for this problem size. */
  const char* Path() const { return fbstab_dense_batch_path(handle_.get()); }

 private:
  struct Destroy {
    void operator()(fbstab_dense_batch* h) const { fbstab_dense_batch_destroy(h); }
  };
  struct DestroyMulti {
    void operator()(fbstab_dense_multi_gpu* h) const { fbstab_dense_multi_gpu_destroy(h); }
  };
  static void Copy(const double* src, double* dst, size_t n) {
    for (size_t i = 0; i < n; i++) dst[i] = src[i];
  }
  // The reference's DenseData constructor checks (components/dense_data.h:53-66)
  // followed by FBstabDense::ValidateInputs (fbstab_dense.h:169-174).
  void ValidateData(long Hr, long Hc, long fn, long Gr, long Gc, long hn, long Ar, long Ac,
                    long bn) const {
    if (Hr != Hc || Hr != fn)
      throw std::runtime_error(
          "In DenseData::DenseData: H must be square and the same size as f");
    if (Ac != Hr || Ar != bn)
      throw std::runtime_error(
          "In DenseData::DenseData: Sizing of data defining Az <= b is inconsistent.");
    if (Gc != Hr || Gr != hn)
      throw std::runtime_error("In DenseData::DenseData: Sizing of Gz = h is inconsistent.");
    if (nz_ != fn || nv_ != bn || nl_ != hn)
      throw std::runtime_error(
          "In FBstabDense::Solve: mismatch between *this and data dimensions.");
  }

  int nz_ = 0, nl_ = 0, nv_ = 0, max_batch_ = 1;
  Options opts_;
  std::unique_ptr<fbstab_dense_batch, Destroy> handle_;
  // SolveBatch(..., devices): created on first use
  std::unique_ptr<fbstab_dense_multi_gpu, DestroyMulti> multi_;
  std::vector<int> multi_devices_;
  int multi_cap_ = 0;
};

}  // namespace fbstab

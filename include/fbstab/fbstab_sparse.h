// fbstab_sparse.h -- FBstabSparse: QPs with SPARSE H, G, A over the batched B200 engine.
//
//     min.  1/2 z'Hz + f'z   s.t.  Gz = h,  Az <= b
//
// The reference does not ship this solver yet: it plans "general sparse matrix
// components" (ROADMAP.md:10, README.md:47) over the LDL' wrapper of
// tools/qdldl/qdldl_wrapper.h:19-84.  The class follows FBstabDense
// (fbstab/fbstab_dense.h:50-194) member for member -- ProblemData, Variable, Options,
// Solve(qp,&x[,os]), UpdateOptions, DefaultOptions, ReliableOptions, the same
// std::runtime_error on bad sizes -- with the matrices in compressed-column form, the
// storage qdldl_wrapper.h:12-14 names (H by its upper triangle), and with the wrapper's
// split between analysis (constructor: ordering, elimination tree, pattern of L) and
// numeric work (every Newton step, on the device).  Every instance a solver object
// sees -- one after the other through Solve, or together through SolveBatch -- has the
// pattern given to the constructor.  There is no CPU solver behind it.
#pragma once

#include <memory>
#include <stdexcept>
#include <vector>

#include "fbstab/fbstab_algorithm.h"
#include "fbstab/linalg.h"
#include "fbstab_b200.h"

namespace fbstab {

/** Compressed-column pattern (0-based, row indices increasing within a column). */
struct SparsePattern {
  int rows = 0, cols = 0;
  std::vector<int> p;  ///< cols + 1 column pointers
  std::vector<int> i;  ///< row indices
  int nnz() const { return p.empty() ? 0 : p.back(); }
};

class FBstabSparse {
 public:
  FBstabSparse(const FBstabSparse&) = delete;
  FBstabSparse& operator=(const FBstabSparse&) = delete;

  /** Values of one instance on the solver's pattern (cf. fbstab_dense.h:55-64). */
  struct ProblemData {
    ProblemData() = default;
    ProblemData(int nnzH, int nnzG, int nnzA, int nz, int nl, int nv)
        : Hx(nnzH), Gx(nnzG), Ax(nnzA), f(nz), h(nl), b(nv) {}
    Eigen::VectorXd Hx;  ///< values of the upper triangle of H, column by column
    Eigen::VectorXd Gx;  ///< values of G
    Eigen::VectorXd Ax;  ///< values of A
    Eigen::VectorXd f;   ///< nz linear cost
    Eigen::VectorXd h;   ///< nl equality rhs
    Eigen::VectorXd b;   ///< nv inequality rhs
  };

  /** Initial guess in, solution out (fbstab_dense.h:85-92). */
  struct Variable {
    Variable(int nz, int nl, int nv) : z(nz), l(nl), v(nv), y(nv) {
      z.setZero();
      l.setZero();
      v.setZero();
      y.setZero();
    }
    Eigen::VectorXd z, l, v, y;
  };
  using QPData = ProblemData;
  using QPVariable = Variable;
  struct Options : public AlgorithmParameters {};

  /**
   * Symbolic analysis + device workspaces for QPs with these patterns: H (nz x nz, upper
   * triangle only), G (nl x nz), A (nv x nz).  Throws std::runtime_error unless nz > 0,
   * nv > 0, nl >= 0 (as fbstab_dense.cc:18-27), on an inconsistent pattern, or if no
   * CUDA device is usable.
   */
  FBstabSparse(const SparsePattern& H, const SparsePattern& G, const SparsePattern& A,
               int max_batch = 1, int device = 0)
      : nz_(H.cols), nl_(G.rows), nv_(A.rows), nnzH_(H.nnz()), nnzG_(G.nnz()), nnzA_(A.nnz()),
        H_(H), G_(G), A_(A) {
    if (nz_ <= 0 || nl_ < 0 || nv_ <= 0)
      throw std::runtime_error("In FBstabSparse::FBstabSparse: Inputs must be positive.");
    if (H.rows != nz_ || (int)H.p.size() != nz_ + 1 || (int)H.i.size() != nnzH_)
      throw std::runtime_error("In FBstabSparse::FBstabSparse: H must be square, nz x nz.");
    if (A.cols != nz_ || (int)A.p.size() != nz_ + 1 || (int)A.i.size() != nnzA_)
      throw std::runtime_error(
          "In FBstabSparse::FBstabSparse: Sizing of data defining Az <= b is inconsistent.");
    if (nl_ > 0 && (G.cols != nz_ || (int)G.p.size() != nz_ + 1 || (int)G.i.size() != nnzG_))
      throw std::runtime_error(
          "In FBstabSparse::FBstabSparse: Sizing of Gz = h is inconsistent.");
    fbstab_sparse_batch* h = nullptr;
    detail::Check(fbstab_sparse_batch_create(nz_, nl_, nv_, H.p.data(), H.i.data(),
                                             nl_ > 0 ? G.p.data() : nullptr,
                                             nl_ > 0 ? G.i.data() : nullptr, A.p.data(),
                                             A.i.data(), nullptr, max_batch, device, &h),
                  "FBstabSparse::FBstabSparse");
    handle_.reset(h);
    opts_.DefaultParameters();
  }

  /** Solves one instance; x is the initial guess and is overwritten (fbstab_dense.h:136-149). */
  template <class OutStream>
  SolverOut Solve(const ProblemData& qp, Variable* x, const OutStream& os) {
    Validate(qp, *x);
    fbstab_out out;
    detail::Check(fbstab_sparse_batch_solve(handle_.get(), 1, qp.Hx.data(), qp.f.data(),
                                            qp.Gx.data(), qp.h.data(), qp.Ax.data(),
                                            qp.b.data(), x->z.data(), x->l.data(),
                                            x->v.data(), x->y.data(), &out, nullptr),
                  "FBstabSparse::Solve");
    SolverOut s = detail::FromC(out);
    detail::ThrowOnStatus(s);
    detail::PrintFinal(opts_, s, os);
    return s;
  }
  SolverOut Solve(const ProblemData& qp, Variable* x) {
    StandardOutput os;
    return Solve(qp, x, os);
  }

  /**
   * Solves `batch` instances in one call: instance-major value arrays (instance i's Hx at
   * Hx + i*nnz(H), ...), host or device memory; z, l, v: initial guesses in, solutions
   * out; y out.  Per-instance failures are reported in SolverOut::status, not thrown.
   */
  std::vector<SolverOut> SolveBatch(int batch, const double* Hx, const double* f,
                                    const double* Gx, const double* h, const double* Ax,
                                    const double* b, double* z, double* l, double* v,
                                    double* y, void* stream = nullptr) {
    std::vector<fbstab_out> out((size_t)(batch > 0 ? batch : 0));
    detail::Check(fbstab_sparse_batch_solve(handle_.get(), batch, Hx, f, Gx, h, Ax, b, z, l, v,
                                            y, out.data(), stream),
                  "FBstabSparse::SolveBatch");
    std::vector<SolverOut> res;
    res.reserve(out.size());
    for (const fbstab_out& o : out) res.push_back(detail::FromC(o));
    return res;
  }

  /**
   * The same batch on several GPUs of this box (cf. FBstabDense::SolveBatch(..., devices)):
   * contiguous instance ranges, one host thread and one engine handle per device, every
   * device with the same elimination order (fbstab_sparse_multi_gpu_solve).  HOST pointers
   * only; `devices`: CUDA ordinals, no duplicates.
   */
  std::vector<SolverOut> SolveBatch(int batch, const double* Hx, const double* f,
                                    const double* Gx, const double* h, const double* Ax,
                                    const double* b, double* z, double* l, double* v,
                                    double* y, const std::vector<int>& devices) {
    fbstab_sparse_multi_gpu* m = nullptr;
    detail::Check(fbstab_sparse_multi_gpu_create(
                      (int)devices.size(), devices.data(), nz_, nl_, nv_, H_.p.data(),
                      H_.i.data(), nl_ > 0 ? G_.p.data() : nullptr,
                      nl_ > 0 ? G_.i.data() : nullptr, A_.p.data(), A_.i.data(), nullptr,
                      batch > 0 ? batch : 1, &m),
                  "FBstabSparse::SolveBatch");
    std::unique_ptr<fbstab_sparse_multi_gpu, DestroyMulti> guard(m);
    fbstab_options o = opts_.ToC();
    detail::Check(fbstab_sparse_multi_gpu_set_options(m, &o), "FBstabSparse::SolveBatch");
    std::vector<fbstab_out> out((size_t)(batch > 0 ? batch : 0));
    detail::Check(fbstab_sparse_multi_gpu_solve(m, batch, Hx, f, Gx, h, Ax, b, z, l, v, y,
                                                out.data()),
                  "FBstabSparse::SolveBatch");
    std::vector<SolverOut> res;
    res.reserve(out.size());
    for (const fbstab_out& r : out) res.push_back(detail::FromC(r));
    return res;
  }

  void UpdateOptions(const Options& options) {
    fbstab_options o = options.ToC();
    detail::Check(fbstab_sparse_batch_set_options(handle_.get(), &o),
                  "FBstabSparse::UpdateOptions");
    detail::Check(fbstab_sparse_batch_get_options(handle_.get(), &o),
                  "FBstabSparse::UpdateOptions");
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
  const char* Path() const { return fbstab_sparse_batch_path(handle_.get()); }
  /** Entries of the factor L of the Newton matrix (the symbolic analysis' result). */
  int FactorNonzeros() const {
    int nnzL = 0;
    fbstab_sparse_batch_analysis(handle_.get(), nullptr, nullptr, &nnzL, nullptr);
    return nnzL;
  }

 private:
  struct Destroy {
    void operator()(fbstab_sparse_batch* h) const { fbstab_sparse_batch_destroy(h); }
  };
  struct DestroyMulti {
    void operator()(fbstab_sparse_multi_gpu* h) const { fbstab_sparse_multi_gpu_destroy(h); }
  };
  void Validate(const ProblemData& qp, const Variable& x) const {
    if (qp.Hx.size() != nnzH_ || qp.Gx.size() != nnzG_ || qp.Ax.size() != nnzA_ ||
        qp.f.size() != nz_ || qp.h.size() != nl_ || qp.b.size() != nv_)
      throw std::runtime_error(
          "In FBstabSparse::Solve: mismatch between *this and data dimensions.");
    if (nz_ != x.z.size() || x.l.size() != nl_ || nv_ != x.v.size() || nv_ != x.y.size())
      throw std::runtime_error(
          "In FBstabSparse::Solve: mismatch between *this and initial guess dimensions.");
  }
  int nz_ = 0, nl_ = 0, nv_ = 0, nnzH_ = 0, nnzG_ = 0, nnzA_ = 0;
  SparsePattern H_, G_, A_;  // kept for the per-device handles of SolveBatch(..., devices)
  Options opts_;
  std::unique_ptr<fbstab_sparse_batch, Destroy> handle_;
};

}  // namespace fbstab

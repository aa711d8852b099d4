// linalg.h -- the matrix/vector types of the facade's public structs.
//
// The reference's public structs are Eigen-typed (fbstab/fbstab_dense.h:55-107,
// fbstab/fbstab_mpc.h:67-150).  Its data classes only ever ask their inputs
// for data()/rows()/cols()/size() (fbstab/components/dense_data.h:44-52,
// mpc_data.h:62-78), so the facade is written against that duck type:
//   * with Eigen on the include path (FBSTAB_USE_EIGEN, or auto-detected) the
//     real Eigen types are used and the facade is source compatible;
//   * without Eigen (this image ships none) a minimal stand-in is provided in
//     namespace Eigen with the members user code of the reference's tests
//     touches: sized construction, comma initialisation (row by row, as in
//     Eigen), operator()(i[,j]), data/rows/cols/size, fill/setZero/setConstant,
//     norm, and Map<> views over raw memory.  Storage is column-major.
#pragma once

#if !defined(FBSTAB_NO_EIGEN) && defined(__has_include)
#if __has_include(<Eigen/Dense>)
#define FBSTAB_HAVE_EIGEN 1
#endif
#endif

#ifdef FBSTAB_HAVE_EIGEN
#include <Eigen/Dense>
#else

#include <cmath>
#include <cstddef>
#include <stdexcept>
#include <type_traits>
#include <vector>

namespace Eigen {

class MatrixXd;
class VectorXd;

namespace shim {

// Row-by-row comma initialiser: `M << 1, 2, 3, 4;`
template <class Derived>
class CommaInit {
 public:
  CommaInit(Derived* m, double first) : m_(m) { put(first); }
  CommaInit& operator,(double v) {
    put(v);
    return *this;
  }

 private:
  void put(double v) {
    const long r = m_->rows(), c = m_->cols();
    if (k_ >= r * c) throw std::out_of_range("too many coefficients");
    m_->coeffRef(k_ / c, k_ % c) = v;
    k_++;
  }
  Derived* m_;
  long k_ = 0;
};

// Shared element access / reductions over (ptr, rows, cols), column-major.
template <class Derived, class Scalar>
class Base {
 public:
  long rows() const { return self().rows_(); }
  long cols() const { return self().cols_(); }
  long size() const { return rows() * cols(); }
  Scalar* data() const { return self().data_(); }
  Scalar& coeffRef(long i, long j) const { return data()[i + j * rows()]; }
  Scalar& operator()(long i, long j) const { return coeffRef(i, j); }
  Scalar& operator()(long i) const { return data()[i]; }
  Scalar& operator[](long i) const { return data()[i]; }
  double norm() const {
    double s = 0.0;
    for (long i = 0; i < size(); i++) s += data()[i] * data()[i];
    return std::sqrt(s);
  }
  // What the reference's tests compute from a solution (the KKT check of
  // fbstab_dense_unit_tests.cc:173-174: H z + f + A' v and min(y, v)); defined below.
  MatrixXd transpose() const;
  template <class D2, class S2>
  VectorXd cwiseMin(const Base<D2, S2>& o) const;
  template <class S = Scalar,
            class = typename std::enable_if<!std::is_const<S>::value>::type>
  void fill(double a) const {
    for (long i = 0; i < size(); i++) data()[i] = a;
  }
  template <class S = Scalar,
            class = typename std::enable_if<!std::is_const<S>::value>::type>
  void setConstant(double a) const {
    fill(a);
  }
  template <class S = Scalar,
            class = typename std::enable_if<!std::is_const<S>::value>::type>
  void setZero() const {
    fill(0.0);
  }

 private:
  const Derived& self() const { return *static_cast<const Derived*>(this); }
};

}  // namespace shim

class MatrixXd : public shim::Base<MatrixXd, double> {
 public:
  MatrixXd() = default;
  MatrixXd(long r, long c) : r_(r), c_(c), v_((size_t)(r * c), 0.0) {}
  void resize(long r, long c) {
    r_ = r;
    c_ = c;
    v_.assign((size_t)(r * c), 0.0);
  }
  static MatrixXd Zero(long r, long c) { return MatrixXd(r, c); }
  shim::CommaInit<MatrixXd> operator<<(double first) {
    return shim::CommaInit<MatrixXd>(this, first);
  }
  long rows_() const { return r_; }
  long cols_() const { return c_; }
  double* data_() const { return const_cast<double*>(v_.data()); }

 private:
  long r_ = 0, c_ = 0;
  std::vector<double> v_;
};

class VectorXd : public shim::Base<VectorXd, double> {
 public:
  VectorXd() = default;
  explicit VectorXd(long n) : v_((size_t)n, 0.0) {}
  void resize(long n) { v_.assign((size_t)n, 0.0); }
  static VectorXd Zero(long n) { return VectorXd(n); }
  shim::CommaInit<VectorXd> operator<<(double first) {
    return shim::CommaInit<VectorXd>(this, first);
  }
  long rows_() const { return (long)v_.size(); }
  long cols_() const { return 1; }
  double* data_() const { return const_cast<double*>(v_.data()); }

 private:
  std::vector<double> v_;
};

// Fixed-size 4-vector used for the (N, nx, nu, nc) size summary
// (fbstab/fbstab_mpc.h:131,168).
class Vector4d : public shim::Base<Vector4d, double> {
 public:
  Vector4d() = default;
  Vector4d(double a, double b, double c, double d) : v_{a, b, c, d} {}
  shim::CommaInit<Vector4d> operator<<(double first) {
    return shim::CommaInit<Vector4d>(this, first);
  }
  long rows_() const { return 4; }
  long cols_() const { return 1; }
  double* data_() const { return const_cast<double*>(v_); }

 private:
  double v_[4] = {0, 0, 0, 0};
};

template <class T>
class Map;

template <>
class Map<MatrixXd> : public shim::Base<Map<MatrixXd>, double> {
 public:
  Map(double* p, long r, long c) : p_(p), r_(r), c_(c) {}
  shim::CommaInit<Map> operator<<(double first) {
    return shim::CommaInit<Map>(this, first);
  }
  long rows_() const { return r_; }
  long cols_() const { return c_; }
  double* data_() const { return p_; }

 private:
  double* p_;
  long r_, c_;
};

template <>
class Map<const MatrixXd> : public shim::Base<Map<const MatrixXd>, const double> {
 public:
  Map(const double* p, long r, long c) : p_(p), r_(r), c_(c) {}
  long rows_() const { return r_; }
  long cols_() const { return c_; }
  const double* data_() const { return p_; }

 private:
  const double* p_;
  long r_, c_;
};

template <>
class Map<VectorXd> : public shim::Base<Map<VectorXd>, double> {
 public:
  Map(double* p, long n) : p_(p), n_(n) {}
  shim::CommaInit<Map> operator<<(double first) {
    return shim::CommaInit<Map>(this, first);
  }
  long rows_() const { return n_; }
  long cols_() const { return 1; }
  double* data_() const { return p_; }

 private:
  double* p_;
  long n_;
};

template <>
class Map<const VectorXd> : public shim::Base<Map<const VectorXd>, const double> {
 public:
  Map(const double* p, long n) : p_(p), n_(n) {}
  long rows_() const { return n_; }
  long cols_() const { return 1; }
  const double* data_() const { return p_; }

 private:
  const double* p_;
  long n_;
};

namespace shim {
template <class D, class S>
MatrixXd Base<D, S>::transpose() const {
  MatrixXd t(cols(), rows());
  for (long j = 0; j < cols(); j++)
    for (long i = 0; i < rows(); i++) t(j, i) = coeffRef(i, j);
  return t;
}
template <class D, class S>
template <class D2, class S2>
VectorXd Base<D, S>::cwiseMin(const Base<D2, S2>& o) const {
  VectorXd m(size());
  for (long i = 0; i < size(); i++) m(i) = data()[i] < o.data()[i] ? data()[i] : o.data()[i];
  return m;
}
// matrix * vector and vector + vector (column vectors: the right operand has one column)
template <class D1, class S1, class D2, class S2>
VectorXd operator*(const Base<D1, S1>& a, const Base<D2, S2>& x) {
  if (a.cols() != x.rows() || x.cols() != 1) throw std::invalid_argument("size mismatch in product");
  VectorXd y(a.rows());
  for (long j = 0; j < a.cols(); j++)
    for (long i = 0; i < a.rows(); i++) y(i) += a.coeffRef(i, j) * x.data()[j];
  return y;
}
template <class D1, class S1, class D2, class S2>
VectorXd operator+(const Base<D1, S1>& a, const Base<D2, S2>& b) {
  if (a.size() != b.size()) throw std::invalid_argument("size mismatch in sum");
  VectorXd y(a.size());
  for (long i = 0; i < a.size(); i++) y(i) = a.data()[i] + b.data()[i];
  return y;
}
}  // namespace shim

}  // namespace Eigen

#endif  // FBSTAB_HAVE_EIGEN

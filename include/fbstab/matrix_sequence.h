// matrix_sequence.h -- sequences of equally sized matrices, the input wire
// format of FBstabMpc.
//
// Same interface and error behaviour as the reference
// (tools/matrix_sequence.h:18-164): element (i,j) of matrix k lives at
// data[k*rows*cols + j*rows + i]; indexing out of range throws
// std::out_of_range, bad sizes / null data throw std::runtime_error.
// MatrixSequence owns its storage (deep copies); MapMatrixSequence is a
// non-owning view (shallow copies).  The contiguous layout is exactly the
// per-instance layout fbstab_mpc_batch_solve consumes.
#pragma once

#include <stdexcept>
#include <vector>

#include "fbstab/linalg.h"

namespace fbstab {

class MatrixSequence {
 public:
  MatrixSequence() = default;

  // Allocates (but does not initialise) len matrices of nrows x ncols.
  MatrixSequence(int len, int nrows, int ncols = 1) {
    if (len < 0) throw std::runtime_error("Negative length input in MatrixSequence");
    if (nrows <= 0 || ncols <= 0)
      throw std::runtime_error("Non-positive row or column count in MatrixSequence");
    len_ = len;
    rows_ = nrows;
    cols_ = ncols;
    storage_.resize((size_t)len * nrows * ncols);
  }

  Eigen::Map<Eigen::MatrixXd> operator()(int k) {
    CheckIndex(k);
    return Eigen::Map<Eigen::MatrixXd>(data() + (size_t)k * rows_ * cols_, rows_, cols_);
  }
  Eigen::Map<const Eigen::MatrixXd> operator()(int k) const {
    CheckIndex(k);
    return Eigen::Map<const Eigen::MatrixXd>(data() + (size_t)k * rows_ * cols_, rows_,
                                             cols_);
  }

  int rows() const { return rows_; }
  int cols() const { return cols_; }
  int length() const { return len_; }
  int size() const { return len_ * rows_ * cols_; }
  double* data() { return storage_.data(); }
  const double* data() const { return storage_.data(); }

 private:
  void CheckIndex(int k) const {
    if (k < 0 || k >= len_) throw std::out_of_range("Bad indexing in MatrixSequence");
  }
  int len_ = 0;
  int rows_ = 1;
  int cols_ = 1;
  std::vector<double> storage_;
};

class MapMatrixSequence {
 public:
  MapMatrixSequence() = default;

  // Views len matrices of nrows x ncols stored contiguously at data.
  MapMatrixSequence(const double* data, int len, int nrows, int ncols) : ptr_(data) {
    if (len <= 0) throw std::runtime_error("Non-positive length input in MapMatrixSequence");
    if (nrows <= 0 || ncols <= 0)
      throw std::runtime_error("Non-positive row or column count in MapMatrixSequence");
    if (data == nullptr)
      throw std::runtime_error("Cannot initialize MapMatrixSequence will a nullptr");
    len_ = len;
    rows_ = nrows;
    cols_ = ncols;
  }

  // Views an owning sequence; the caller keeps it alive.
  MapMatrixSequence(const MatrixSequence& A)
      : ptr_(A.data()), len_(A.length()), rows_(A.rows()), cols_(A.cols()) {}

  Eigen::Map<const Eigen::MatrixXd> operator()(int k) const {
    if (k < 0 || k >= len_) throw std::out_of_range("Bad indexing in MapMatrixSequence");
    if (ptr_ == nullptr)
      throw std::runtime_error("In MapMatrixSequence, cannot index into null data.");
    return Eigen::Map<const Eigen::MatrixXd>(ptr_ + (size_t)k * rows_ * cols_, rows_, cols_);
  }

  int rows() const { return rows_; }
  int cols() const { return cols_; }
  int length() const { return len_; }
  int size() const { return len_ * rows_ * cols_; }
  const double* data() const { return ptr_; }

 private:
  const double* ptr_ = nullptr;
  int len_ = 0;
  int rows_ = 1;
  int cols_ = 1;
};

}  // namespace fbstab

// ocp_generator.h -- the four benchmark optimal control problems in the
// format FBstabMpc accepts.
//
// Mirrors the reference's test fixture (fbstab/test/ocp_generator.h:19-160):
// call one of the problem creators, then GetFBstabInput() /
// GetFBstabInputRef().  The stage data comes from fbstab_ocp_generate (host
// code in the engine library, see fbstab_b200.h), which restates
// ocp_generator.cc:73-421 (time-varying wire format, E(0) = 0).
#pragma once

#include <stdexcept>

#include "fbstab/fbstab_mpc.h"
#include "fbstab_b200.h"

namespace fbstab {
namespace test {

class OcpGenerator {
 public:
  OcpGenerator() = default;

  void DoubleIntegrator(int N = 10) { Generate(FBSTAB_OCP_DOUBLE_INTEGRATOR, N); }
  void ServoMotor(int N = 20) { Generate(FBSTAB_OCP_SERVO_MOTOR, N); }
  void SpacecraftRelativeMotion(int N = 40) { Generate(FBSTAB_OCP_SPACECRAFT, N); }
  void CopolymerizationReactor(int N = 70) { Generate(FBSTAB_OCP_COPOLYMERIZATION, N); }

  /** A deep copy of the problem data. */
  FBstabMpc::ProblemData GetFBstabInput() const {
    RequireInitialized();
    return data_;
  }
  /** A view of the generator's own storage; keep the generator alive. */
  FBstabMpc::ProblemDataRef GetFBstabInputRef() const {
    RequireInitialized();
    return FBstabMpc::ProblemDataRef(&data_.Q, &data_.R, &data_.S, &data_.q, &data_.r, &data_.A,
                                     &data_.B, &data_.c, &data_.E, &data_.L, &data_.d,
                                     &data_.x0);
  }

  /**
   * Inputs for a closed-loop simulation (reference ocp_generator.h:31-38,69):
   *     x(i+1) = A x(i) + B u(i),   y(i) = C x(i) + D u(i)
   * for T steps from x(0) = x0.  Feed them to fbstab::ClosedLoopMpc
   * (include/fbstab/closed_loop.h) to run the loop on the GPU.
   */
  struct SimulationInputs {
    Eigen::VectorXd x0;
    Eigen::MatrixXd A;
    Eigen::MatrixXd B;
    Eigen::MatrixXd C;
    Eigen::MatrixXd D;
    int T = 0;
  };
  SimulationInputs GetSimulationInputs() const {
    if (!initialized_)
      throw std::runtime_error(
          "In OcpGenerator::GetSimulationInputs: Call a problem creator method first.");
    SimulationInputs out;
    int ny = 0;
    if (fbstab_ocp_simulation(kind_, nullptr, nullptr, nullptr, nullptr, &ny, &out.T) !=
        FBSTAB_OK)
      throw std::runtime_error(fbstab_last_error());
    out.x0 = Eigen::VectorXd(nx_);
    out.A = Eigen::MatrixXd(nx_, nx_);
    out.B = Eigen::MatrixXd(nx_, nu_);
    out.C = Eigen::MatrixXd(ny, nx_);
    out.D = Eigen::MatrixXd(ny, nu_);
    for (int i = 0; i < ny * nu_; i++) out.D.data()[i] = 0.0;
    if (fbstab_ocp_simulation(kind_, out.A.data(), out.B.data(), out.C.data(), out.x0.data(),
                              nullptr, nullptr) != FBSTAB_OK)
      throw std::runtime_error(fbstab_last_error());
    return out;
  }

  /** (N, nx, nu, nc) */
  Eigen::Vector4d ProblemSize() const {
    Eigen::Vector4d s;
    s << (double)N_, (double)nx_, (double)nu_, (double)nc_;
    return s;
  }
  int nz() const { return (N_ + 1) * (nx_ + nu_); }
  int nl() const { return (N_ + 1) * nx_; }
  int nv() const { return (N_ + 1) * nc_; }
  int N() const { return N_; }
  int nx() const { return nx_; }
  int nu() const { return nu_; }
  int nc() const { return nc_; }

 private:
  void RequireInitialized() const {
    if (!initialized_)
      throw std::runtime_error(
          "In OcpGenerator: call a problem creator method before requesting data.");
  }
  void Generate(int kind, int N) {
    if (N <= 0) throw std::runtime_error("In OcpGenerator: N <= 0.");
    if (fbstab_ocp_dims(kind, &nx_, &nu_, &nc_) != FBSTAB_OK)
      throw std::runtime_error(fbstab_last_error());
    N_ = N;
    kind_ = kind;
    data_.Q = MatrixSequence(N + 1, nx_, nx_);
    data_.R = MatrixSequence(N + 1, nu_, nu_);
    data_.S = MatrixSequence(N + 1, nu_, nx_);
    data_.q = MatrixSequence(N + 1, nx_, 1);
    data_.r = MatrixSequence(N + 1, nu_, 1);
    data_.A = MatrixSequence(N, nx_, nx_);
    data_.B = MatrixSequence(N, nx_, nu_);
    data_.c = MatrixSequence(N, nx_, 1);
    data_.E = MatrixSequence(N + 1, nc_, nx_);
    data_.L = MatrixSequence(N + 1, nc_, nu_);
    data_.d = MatrixSequence(N + 1, nc_, 1);
    data_.x0 = Eigen::VectorXd(nx_);
    if (fbstab_ocp_generate(kind, N, data_.Q.data(), data_.R.data(), data_.S.data(),
                            data_.q.data(), data_.r.data(), data_.A.data(), data_.B.data(),
                            data_.c.data(), data_.E.data(), data_.L.data(), data_.d.data(),
                            data_.x0.data()) != FBSTAB_OK)
      throw std::runtime_error(fbstab_last_error());
    initialized_ = true;
  }

  FBstabMpc::ProblemData data_;
  int N_ = 0, nx_ = 0, nu_ = 0, nc_ = 0, kind_ = 0;
  bool initialized_ = false;
};

}  // namespace test
}  // namespace fbstab

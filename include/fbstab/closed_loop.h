// closed_loop.h -- batched receding-horizon (closed-loop) MPC on the GPU.
//
// The reference ships the INPUTS of such a simulation -- OcpGenerator::
// GetSimulationInputs (fbstab/test/ocp_generator.h:31-38,69; ocp_generator.cc:
// 56-71: x0, A, B, C, D, T) -- and leaves the loop to the user: solve the OCP
// from the measured state, apply the first input, shift the previous solution
// as the next warm start ("can be easily warmstarted", README.md:20).  This
// class is that loop for `batch` plants at once, running entirely on the
// device behind fbstab_mpc_closed_loop_* (include/fbstab_b200.h).
#pragma once

#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "fbstab/fbstab_algorithm.h"
#include "fbstab/fbstab_mpc.h"
#include "fbstab_b200.h"

namespace fbstab {

class ClosedLoopMpc {
 public:
  struct Trajectory {
    int batch = 0, steps = 0, nx = 0, nu = 0;
    std::vector<double> X;  // batch x (steps + 1) x nx
    std::vector<double> U;  // batch x steps x nu
    std::vector<SolverOut> out;  // steps x batch
    const double* x(int plant, int step) const {
      return X.data() + ((size_t)plant * (steps + 1) + step) * nx;
    }
    const double* u(int plant, int step) const {
      return U.data() + ((size_t)plant * steps + step) * nu;
    }
  };

  /**
   * One plant model, `batch` initial states (x_init: batch * nx doubles).
   * qp: ONE FBstabMpc::ProblemData / ProblemDataRef (its x0 is ignored).
   * Asim, Bsim: the plant x+ = Asim x + Bsim u, column-major, e.g. from
   * OcpGenerator::GetSimulationInputs(); nullptr: stage 0 of the OCP itself,
   * x+ = A(0) x + B(0) u + c(0).
   */
  template <class InputData>
  ClosedLoopMpc(const InputData& qp, int batch, const double* x_init, const double* Asim,
                const double* Bsim, int max_steps, int device = 0)
      : batch_(batch), max_steps_(max_steps) {
    N_ = qp.Q.length() - 1;
    nx_ = qp.Q.rows();
    nu_ = qp.R.rows();
    nc_ = qp.E.rows();
    fbstab_mpc_closed_loop* h = nullptr;
    detail::Check(fbstab_mpc_closed_loop_create(
                      N_, nx_, nu_, nc_, batch, device, /*shared_data=*/1, qp.Q.data(),
                      qp.R.data(), qp.S.data(), qp.q.data(), qp.r.data(), qp.A.data(), qp.B.data(),
                      qp.c.data(), qp.E.data(), qp.L.data(), qp.d.data(), x_init, Asim, Bsim,
                      max_steps, &h),
                  "ClosedLoopMpc::ClosedLoopMpc");
    handle_.reset(h);
  }

  void UpdateOptions(const FBstabMpc::Options& options) {
    fbstab_options o = options.ToC();
    detail::Check(fbstab_mpc_closed_loop_set_options(handle_.get(), &o),
                  "ClosedLoopMpc::UpdateOptions");
  }

  /** Simulates `steps` control steps from x_init and returns the logged trajectory. */
  Trajectory Run(int steps, bool warm_start = true) {
    Trajectory t;
    t.batch = batch_;
    t.steps = steps;
    t.nx = nx_;
    t.nu = nu_;
    t.X.resize((size_t)batch_ * (steps + 1) * nx_);
    t.U.resize((size_t)batch_ * steps * nu_);
    std::vector<fbstab_out> out((size_t)steps * batch_);
    detail::Check(fbstab_mpc_closed_loop_run(handle_.get(), steps, warm_start ? 1 : 0,
                                             t.X.data(), t.U.data(), out.data(), nullptr),
                  "ClosedLoopMpc::Run");
    t.out.reserve(out.size());
    for (const fbstab_out& o : out) t.out.push_back(detail::FromC(o));
    return t;
  }

  const char* Path() const { return fbstab_mpc_closed_loop_path(handle_.get()); }

 private:
  struct Destroy {
    void operator()(fbstab_mpc_closed_loop* h) const { fbstab_mpc_closed_loop_destroy(h); }
  };
  int N_ = 0, nx_ = 0, nu_ = 0, nc_ = 0, batch_ = 0, max_steps_ = 0;
  std::unique_ptr<fbstab_mpc_closed_loop, Destroy> handle_;
};

}  // namespace fbstab

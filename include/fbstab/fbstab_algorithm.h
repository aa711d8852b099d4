// fbstab_algorithm.h -- exit flags, solver output, display levels, algorithm
// parameters and the output-stream concept of the C++ facade.
//
// Same names and meaning as the reference (fbstab/fbstab_algorithm.h:17-82,
// tools/output_stream.h:16-38).  The algorithm itself runs on the GPU behind
// the C-ABI of fbstab_b200.h; this header only carries the host-side types.
#pragma once

#include <cstdio>
#include <stdexcept>
#include <string>

#include "fbstab_b200.h"

namespace fbstab {

// Return codes for the solver (reference fbstab_algorithm.h:17-24).
enum class ExitFlag {
  SUCCESS = 0,
  DIVERGENCE = 1,
  MAXITERATIONS = 2,
  PRIMAL_INFEASIBLE = 3,
  DUAL_INFEASIBLE = 4,
  PRIMAL_DUAL_INFEASIBLE = 5
};

// Output data (reference fbstab_algorithm.h:30-37).  A negative solve_time
// means that no timing data is available.  For batched calls solve_time is
// the wall time of the whole batched call.
struct SolverOut {
  ExitFlag eflag = ExitFlag::MAXITERATIONS;
  double residual = 0.0;
  int newton_iters = 0;
  int prox_iters = 0;
  double solve_time = 0.0;
  double initial_residual = 0.0;
  // Batched-engine extras (not in the reference): where the reference throws
  // for one instance (factorisation failure, saturate misuse) a batched call
  // reports FBSTAB_STATUS_* here instead; trajectory counters.
  int status = 0;
  int ls_backtracks = 0;
  int residual_evals = 0;
};

// Display settings (reference fbstab_algorithm.h:40-45).
enum class Display { OFF = 0, FINAL = 1, ITER = 2, ITER_DETAILED = 3 };

// Algorithm parameters (reference fbstab_algorithm.h:48-82).  The in-class
// initialisers are the reference's; solvers start from DefaultParameters()
// exactly like the reference does (fbstab_algorithm-impl.h:107).
struct AlgorithmParameters {
  double sigma0 = 1e-8;
  double sigma_max = 1e-8;
  double sigma_min = 1e-10;
  double alpha = 0.95;
  double beta = 0.7;
  double eta = 1e-8;
  double delta = 1.0 / 5.0;
  double gamma = 1.0 / 10.0;

  double abs_tol = 1e-6;
  double rel_tol = 1e-12;
  double stall_tol = 1e-10;
  double infeas_tol = 1e-8;

  double inner_tol_max = 1e-1;
  double inner_tol_min = 1e-12;

  int max_newton_iters = 500;
  int max_prox_iters = 100;
  int max_inner_iters = 100;
  int max_linesearch_iters = 20;

  bool check_feasibility = true;
  bool nonmonotone_linesearch = true;
  Display display_level = Display::FINAL;

  // Engine extensions, both off by default (= the reference's behaviour): one
  // step of iterative refinement per Newton system and regularise-and-retry
  // attempts after a failed factorisation -- the two TODOs the reference leaves
  // in its linear solvers (components/abstract_components.h:335-337,
  // components/riccati_linear_solver.cc:129-130); see fbstab_options.
  int refine_steps = 0;
  int regularize_retries = 0;

  // Checks validity of fields and overwrites if necessary
  // (fbstab_algorithm-impl.h:7-31); throws std::runtime_error where the
  // reference's saturate() throws.
  void ValidateOptions() {
    fbstab_options o = ToC();
    if (fbstab_validate_options(&o) != FBSTAB_OK)
      throw std::runtime_error(fbstab_last_error());
    FromC(o);
  }
  // Overwrites with defaults (fbstab_algorithm-impl.h:33-59).
  void DefaultParameters() {
    fbstab_options o;
    fbstab_default_options(&o);
    FromC(o);
  }
  // Overwrites with parameters for hard problems (impl:61-74).
  void ReliableParameters() {
    fbstab_options o;
    fbstab_reliable_options(&o);
    FromC(o);
  }

  fbstab_options ToC() const {
    fbstab_options o;
    o.sigma0 = sigma0;
    o.sigma_max = sigma_max;
    o.sigma_min = sigma_min;
    o.alpha = alpha;
    o.beta = beta;
    o.eta = eta;
    o.delta = delta;
    o.gamma = gamma;
    o.abs_tol = abs_tol;
    o.rel_tol = rel_tol;
    o.stall_tol = stall_tol;
    o.infeas_tol = infeas_tol;
    o.inner_tol_max = inner_tol_max;
    o.inner_tol_min = inner_tol_min;
    o.max_newton_iters = max_newton_iters;
    o.max_prox_iters = max_prox_iters;
    o.max_inner_iters = max_inner_iters;
    o.max_linesearch_iters = max_linesearch_iters;
    o.check_feasibility = check_feasibility ? 1 : 0;
    o.nonmonotone_linesearch = nonmonotone_linesearch ? 1 : 0;
    o.display_level = static_cast<int>(display_level);
    o.refine_steps = refine_steps;
    o.regularize_retries = regularize_retries;
    return o;
  }
  void FromC(const fbstab_options& o) {
    sigma0 = o.sigma0;
    sigma_max = o.sigma_max;
    sigma_min = o.sigma_min;
    alpha = o.alpha;
    beta = o.beta;
    eta = o.eta;
    delta = o.delta;
    gamma = o.gamma;
    abs_tol = o.abs_tol;
    rel_tol = o.rel_tol;
    stall_tol = o.stall_tol;
    infeas_tol = o.infeas_tol;
    inner_tol_max = o.inner_tol_max;
    inner_tol_min = o.inner_tol_min;
    max_newton_iters = o.max_newton_iters;
    max_prox_iters = o.max_prox_iters;
    max_inner_iters = o.max_inner_iters;
    max_linesearch_iters = o.max_linesearch_iters;
    check_feasibility = o.check_feasibility != 0;
    nonmonotone_linesearch = o.nonmonotone_linesearch != 0;
    display_level = static_cast<Display>(o.display_level);
    refine_steps = o.refine_steps;
    regularize_retries = o.regularize_retries;
  }
};

// Printing interface (reference tools/output_stream.h:16-38): CRTP base whose
// Print forwards to T::PrintImplementation.
template <class T>
class OutputStream {
 public:
  void Print(const char* message) const {
    static_cast<const T*>(this)->PrintImplementation(message);
  }
};

class StandardOutput : public OutputStream<StandardOutput> {
 public:
  StandardOutput() = default;

 protected:
  void PrintImplementation(const char* message) const { printf("%s", message); }
  friend class OutputStream<StandardOutput>;
};

namespace detail {

inline void Check(int rc, const char* where) {
  if (rc != FBSTAB_OK)
    throw std::runtime_error(std::string("In ") + where + ": " + fbstab_last_error());
}

inline SolverOut FromC(const fbstab_out& o) {
  SolverOut s;
  s.eflag = static_cast<ExitFlag>(o.eflag);
  s.residual = o.residual;
  s.newton_iters = o.newton_iters;
  s.prox_iters = o.prox_iters;
  s.solve_time = o.solve_time;
  s.initial_residual = o.initial_residual;
  s.status = o.status;
  s.ls_backtracks = o.ls_backtracks;
  s.residual_evals = o.residual_evals;
  return s;
}

// Where the reference throws for a single instance
// (fbstab_algorithm-impl.h:263-274, tools/utilities.h:21-25) so does the
// single-instance facade.
inline void ThrowOnStatus(const SolverOut& s) {
  if (s.status == FBSTAB_STATUS_FACTOR_FAILED)
    throw std::runtime_error("In FBstabAlgorithm::Solve: Linear solver failed.");
  if (s.status == FBSTAB_STATUS_SATURATE)
    throw std::runtime_error("In saturate: upper bound must be larger than the lower bound");
}

// Summary printed at Display::FINAL and above, in the reference's format
// (fbstab_algorithm-impl.h:488-541).  Per-iteration lines (ITER,
// ITER_DETAILED) have no batched counterpart: the iterations happen on the
// device; the per-instance counters are in SolverOut instead.
template <class OutStream>
void PrintFinal(const AlgorithmParameters& opts, const SolverOut& s,
                const OutStream& os) {
  if (opts.display_level < Display::FINAL) return;
  static const char* names[6] = {" Success\n",
                                 " Divergence\n",
                                 " Iteration limit exceeded\n",
                                 " Primal Infeasibility\n",
                                 " Dual Infeasibility\n",
                                 " Primal-Dual Infeasibility\n"};
  char buff[100];
  os.Print("\nOptimization completed!  Exit code:");
  os.Print(names[static_cast<int>(s.eflag)]);
  snprintf(buff, 100, "Time elapsed: %f ms (-1.0 indicates timing disabled)\n",
           1000.0 * s.solve_time);
  os.Print(buff);
  snprintf(buff, 100, "Proximal iterations: %d out of %d\n", s.prox_iters,
           opts.max_prox_iters);
  os.Print(buff);
  snprintf(buff, 100, "Newton iterations: %d out of %d\n", s.newton_iters,
           opts.max_newton_iters);
  os.Print(buff);
  snprintf(buff, 100, "%10s  %10s\n", "|r|", "|r0|");
  os.Print(buff);
  snprintf(buff, 100, "%10.4e  %10.4e\n\n", s.residual, s.initial_residual);
  os.Print(buff);
}

}  // namespace detail

}  // namespace fbstab

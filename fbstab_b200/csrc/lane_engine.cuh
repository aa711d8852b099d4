// lane_engine.cuh -- FBstabAlgorithm::Solve as a PER-LANE phase machine.
//
// Device restatement of FBstabAlgorithm::Solve / ::SolveProximalSubproblem
// (reference fbstab/fbstab_algorithm-impl.h:113-304) for kernels in which every
// LANE owns one QP instance (mpc_lane.cu: small-stage OCPs; sparse_lane.cu:
// sparse QPs with a common sparsity pattern).  The warp runs rounds of
//   evaluate -> decide -> (commit) -> Newton step -> end of subproblem -> output
// and a lane takes part in the sweeps its instance needs; sweeps that feed a
// shared-memory ring or that are pattern-uniform are executed by the WHOLE warp
// with `p.on` marking the lanes they are for (stores are predicated on it).
// Lanes pull instance indices from the global atomic counter on their own, so a
// lane whose instance has finished starts the next one in the following round.
//
// The policy LN supplies (semantics as in engine.cuh, per lane):
//   static constexpr int O_XK, O_XI, O_DX        ids of the iterate blocks
//   int nz, nl, nv; bool on, enabled            (enabled: the lane may own instances)
//   void   bind(args, inst)
//   double init(z0, l0, v0)                       -> forcing norm; xk = xi = x0, y = b - A z0
//   EvalOut evaluate(base, trial, t, self_bar, sigma, alpha)   fused residual evaluation
//   void   commit(t)                              xi <- xi + t dx
//   bool   factor(sigma, alpha, with_commit, t, any_commit)    LinearSolver::Initialize
//   void   solve()                                LinearSolver::Solve on -ri -> dx
//   double prox_end(do_diff, with_commit, t, any_commit, tol, check, &feas)
//   void   write_result(from, z, l, v, y)
#pragma once

#include "common.cuh"
#include "engine.cuh"
#include "fbstab_b200.h"

namespace fbs {

enum { PH_TOP = 0, PH_TRIAL = 1, PH_REEVAL = 2, PH_FINAL = 3 };

template <class LN, class Args>
__device__ __forceinline__ void lane_solve_loop(LN& p, const Args& a) {
  const fbstab_options& o = a.opts;
  const double sigma = o.sigma0, alpha = o.alpha;
  // per-lane solver state (fbstab_algorithm-impl.h:113-304 as a phase machine)
  bool active = false, exhausted = !p.enabled;  // a disabled lane never owns an instance
  int inst = 0, phase = PH_TOP;
  int eflag = FBSTAB_MAXITERATIONS, status = FBSTAB_STATUS_OK;
  int newton = 0, prox = 0, backtracks = 0, evals = 0;
  int k = 0, inner_i = 0, ls_j = 0;
  double E0 = 0, Ek = 0, last_rk = 0, inner_tol = 0, combo_tol = 0, dx_norm = 0;
  double merit[5] = {0, 0, 0, 0, 0};
  double Eo = 0, Ei_c = 0, Eo_c = 0, tstep = 1.0, m0 = 0, current_merit = 0;
  bool need_eval = false, pick_xi = false;

  for (;;) {
    // ---- idle lanes pull the next instance --------------------------------
    if (!active && !exhausted) {
      inst = atomicAdd(a.counter, 1);
      if (inst >= a.batch) {
        exhausted = true;
      } else {
        p.bind(a, inst);
        const double fn = p.init(a.z + (size_t)inst * p.nz, a.l + (size_t)inst * p.nl,
                                 a.v + (size_t)inst * p.nv);
        combo_tol = o.abs_tol + o.rel_tol * (1.0 + fn);
        dx_norm = sqrt((double)p.nz + (double)p.nl + (double)p.nv);  // dx_.Fill(1.0), impl:142
        eflag = FBSTAB_MAXITERATIONS;
        status = FBSTAB_STATUS_OK;
        newton = prox = backtracks = evals = 0;
        k = inner_i = ls_j = 0;
        E0 = Ek = last_rk = inner_tol = 0.0;
        phase = PH_TOP;
        need_eval = true;
        active = true;
      }
    }
    if (!__any_sync(0xffffffffu, active)) break;

    // ---- evaluate -----------------------------------------------------------
    // (the ring sweeps are executed by the whole warp -- lane 0 feeds the ring --
    // and `on` marks the lanes they are for)
    EvalOut e;
    e.Ei = e.Eo = 0.0;
    const bool ev = active && need_eval;
    if (__any_sync(0xffffffffu, ev)) {
      const bool self_bar = (phase == PH_TOP) || (phase == PH_FINAL);
      const int base = (phase == PH_FINAL && !pick_xi) ? LN::O_XK : LN::O_XI;
      p.on = ev;
      const EvalOut e2 = p.evaluate(base, phase == PH_TRIAL, tstep, self_bar, sigma, alpha);
      p.on = true;
      if (ev) {
        e = e2;
        evals++;
      }
    }
    // ---- decide ---------------------------------------------------------------
    bool do_commit = false, do_newton = false, prox_end = false, finish = false;
    bool to_inner_top = false;
    int which = 0;  // 0: xk, 1: xi, 2: dx is the result
    if (active) {
      need_eval = false;
      if (phase == PH_TOP) {  // impl:158-185
        Ek = e.Eo;
        last_rk = Ek;
        bool bad = false;
        if (k == 0) {
          E0 = Ek;
          inner_tol = saturate(E0, o.inner_tol_min, o.inner_tol_max, &bad);
        }
        if (!bad && (Ek <= combo_tol || dx_norm <= o.stall_tol)) {
          eflag = FBSTAB_SUCCESS;
          finish = true;
        } else {
          if (!bad) inner_tol = saturate(inner_tol * o.delta, o.inner_tol_min, Ek, &bad);
          if (bad) {
            status = FBSTAB_STATUS_SATURATE;
            finish = true;
          } else {
#pragma unroll
            for (int m = 0; m < 5; m++) merit[m] = 0.0;
            Ei_c = e.Ei;
            Eo_c = e.Eo;
            inner_i = 0;
            to_inner_top = true;
          }
        }
      } else if (phase == PH_TRIAL) {  // Armijo test, impl:286-296
        const double mp = 0.5 * e.Ei * e.Ei;
        if (mp <= m0 - 2.0 * tstep * o.eta * current_merit) {
          do_commit = true;
          Ei_c = e.Ei;
          Eo_c = e.Eo;
          inner_i++;
          to_inner_top = true;
        } else {
          tstep *= o.beta;
          backtracks++;
          ls_j++;
          if (ls_j < o.max_linesearch_iters) {
            need_eval = true;  // next trial
          } else {
            // every trial failed: the step is still taken (impl:295-298)
            do_commit = true;
            inner_i++;
            if (inner_i < o.max_inner_iters) {
              phase = PH_REEVAL;
              need_eval = true;
            } else {
              prox_end = true;
            }
          }
        }
      } else if (phase == PH_REEVAL) {
        Ei_c = e.Ei;
        Eo_c = e.Eo;
        to_inner_top = true;
      } else {  // PH_FINAL
        last_rk = e.Eo;
        eflag = FBSTAB_MAXITERATIONS;
        which = pick_xi ? 1 : 0;
        finish = true;
      }
      if (to_inner_top) {  // top of an inner iteration, impl:237-260
        bool inner_done = (inner_i >= o.max_inner_iters);
        if (!inner_done) {
          Eo = Eo_c;
          last_rk = Eo;
          if ((Ei_c <= inner_tol && Eo < Ek) || (Ei_c <= o.inner_tol_min)) inner_done = true;
          if (newton >= o.max_newton_iters) inner_done = true;
        }
        if (inner_done)
          prox_end = true;
        else
          do_newton = true;
      }
    }
    // ---- commit the accepted (or forced) step: xi <- xi + t dx ---------------
    // (lanes that go on to a Newton step commit inside the factor sweep)
    // (and lanes whose subproblem ends commit inside the prox_end sweep)
    if (do_commit && !do_newton && !prox_end) p.commit(tstep);
    // ---- Newton step ------------------------------------------------------------
    if (__any_sync(0xffffffffu, do_newton)) {
      const bool wc = do_newton && do_commit;
      p.on = do_newton;
      const bool fok = p.factor(sigma, alpha, wc, tstep, __any_sync(0xffffffffu, wc));
      p.solve();
      p.on = true;
      if (do_newton) {
        if (!fok) {  // impl:263-267
          status = FBSTAB_STATUS_FACTOR_FAILED;
          finish = true;
          which = 0;
        } else {
          newton++;
          current_merit = 0.5 * Ei_c * Ei_c;
#pragma unroll
          for (int m = 4; m > 0; m--) merit[m] = merit[m - 1];
          merit[0] = current_merit;
          m0 = current_merit;
          if (o.nonmonotone_linesearch) {
#pragma unroll
            for (int m = 1; m < 5; m++) m0 = fmax(m0, merit[m]);
          }
          tstep = 1.0;
          ls_j = 0;
          phase = PH_TRIAL;
          need_eval = true;
        }
      }
    }
    // ---- end of the subproblem, impl:300-216 -------------------------------------
    if (__any_sync(0xffffffffu, prox_end)) {
      const bool at_cap = newton >= o.max_newton_iters;  // impl:188-199
      int feas = 0;
      const bool pc = prox_end && do_commit;
      p.on = prox_end;
      const double dn = p.prox_end(prox_end && !at_cap, pc, tstep, __any_sync(0xffffffffu, pc),
                                   o.infeas_tol, o.check_feasibility != 0, &feas);
      p.on = true;
      if (prox_end) {
        if (at_cap) {
          pick_xi = Eo < Ek;
          phase = PH_FINAL;
          need_eval = true;
        } else {
          dx_norm = dn;
          if (feas != 0) {
            eflag = (feas == 1)   ? FBSTAB_PRIMAL_INFEASIBLE
                    : (feas == 2) ? FBSTAB_DUAL_INFEASIBLE
                                  : FBSTAB_PRIMAL_DUAL_INFEASIBLE;
            which = 2;
            finish = true;
          } else {
            prox++;
            k++;
            if (k >= o.max_prox_iters) {
              eflag = FBSTAB_MAXITERATIONS;
              which = 0;
              finish = true;
            } else {
              phase = PH_TOP;
              need_eval = true;
            }
          }
        }
      }
    }
    // ---- WriteVariable + PrepareOutput, impl:349-383 ---------------------------
    if (finish) {
      const int from = which == 0 ? LN::O_XK : which == 1 ? LN::O_XI : LN::O_DX;
      p.write_result(from, a.z + (size_t)inst * p.nz, a.l + (size_t)inst * p.nl,
                     a.v + (size_t)inst * p.nv, a.y + (size_t)inst * p.nv);
      fbstab_out* out = a.out + inst;
      out->eflag = eflag;
      out->newton_iters = newton;
      out->prox_iters = prox;
      out->status = status;
      out->residual = last_rk;
      out->initial_residual = E0;
      out->solve_time = -1.0;
      out->ls_backtracks = backtracks;
      out->residual_evals = evals;
      active = false;
    }
  }
}

}  // namespace fbs

// engine.cuh -- the per-instance FBstab state machine, executed by one team.
//
// Device restatement of FBstabAlgorithm::Solve and ::SolveProximalSubproblem
// (reference fbstab/fbstab_algorithm-impl.h:113-304).  The Problem policy
// supplies the structure-exploiting pieces that the reference obtains from its
// (Data, LinearSolver, Feasibility) components
// (fbstab/components/abstract_components.h:24-338):
//
//   int nz, nl, nv
//   double forcing_norm(team)                       Data::ForcingNorm
//   double b(i)                                     entry i of the rhs of Az<=b
//   void margin(team, z, y)                         y = b - A z
//   void kkt(team, x, tz, tl)                       tz = f+Hz+G'l+A'v, tl = h-Gz
//   bool factor(team, x, xbar, sigma, alpha)        LinearSolver::Initialize
//   void solve(team, rz, rl, rv, dx)                LinearSolver::Solve on -r
//   int  feasibility(team, dx, tol)                 CheckFeasibility
//   double fvec(i), hvec(i); double* gamma, mus     f, h and the barrier terms of the
//                                                   last factor() (iterative refinement)
//
// Work the reference does twice is done once (results are bit-identical):
//  * InnerResidual and PenalizedNaturalResidual share f+Hz+G'l+A'v and h-Gz
//    (full_residual.cc:49-109) -> one fused evaluation yields both norms;
//  * an accepted Armijo trial point IS the next iterate, so its residual is
//    carried to the next top-of-loop instead of being recomputed (impl:239 vs
//    impl:286-289);
//  * at the start of a subproblem x == xbar, so R(x,xbar,sigma) equals the
//    natural-residual evaluation just done at the top of the proximal loop
//    (impl:162 vs impl:239-243), and E0 equals the k=0 evaluation (impl:144).
#pragma once

#include "common.cuh"

namespace fbs {

struct Resid {
  double* z;
  double* l;
  double* v;
};

struct EvalOut {
  double Ei;  // ||R(x,xbar,sigma)||      (FullResidual::Norm of the inner residual)
  double Eo;  // ||pi_pen(x)||            (penalised natural residual norm)
};

// x <- x + a*dx with the y-aware rule of FullVariable::axpy
// (full_variable.cc:55-65): y += a*dx.y ; y += (-a)*b.
template <class P>
__device__ __forceinline__ void vars_axpy(const Team& t, const P& p,
                                          const Vars& src, double a,
                                          const Vars& dx, const Vars& dst) {
  for (int i = t.rank(); i < p.nz; i += t.size()) dst.z[i] = src.z[i] + a * dx.z[i];
  for (int i = t.rank(); i < p.nl; i += t.size()) dst.l[i] = src.l[i] + a * dx.l[i];
  for (int i = t.rank(); i < p.nv; i += t.size()) {
    dst.v[i] = src.v[i] + a * dx.v[i];
    double y = src.y[i] + a * dx.y[i];
    dst.y[i] = y + (-a) * p.b(i);
  }
  t.sync();
}

template <class P>
__device__ __forceinline__ void vars_copy(const Team& t, const P& p,
                                          const Vars& src, const Vars& dst) {
  for (int i = t.rank(); i < p.nz; i += t.size()) dst.z[i] = src.z[i];
  for (int i = t.rank(); i < p.nl; i += t.size()) dst.l[i] = src.l[i];
  for (int i = t.rank(); i < p.nv; i += t.size()) {
    dst.v[i] = src.v[i];
    dst.y[i] = src.y[i];
  }
  t.sync();
}

// Fused residual evaluation (see header comment).  ri receives the inner
// residual vectors; both norms are returned.
template <class P>
__device__ __forceinline__ EvalOut evaluate(const Team& t, P& p, const Vars& x,
                                            const Vars& xbar, double sigma,
                                            double alpha, const Resid& ri) {
  p.kkt(t, x, ri.z, ri.l);
  double s[6] = {0, 0, 0, 0, 0, 0};
  for (int i = t.rank(); i < p.nz; i += t.size()) {
    const double tz = ri.z[i];
    s[3] += tz * tz;
    const double r = tz + sigma * (x.z[i] - xbar.z[i]);
    ri.z[i] = r;
    s[0] += r * r;
  }
  for (int i = t.rank(); i < p.nl; i += t.size()) {
    const double tl = ri.l[i];
    s[4] += tl * tl;
    const double r = tl + sigma * (x.l[i] - xbar.l[i]);
    ri.l[i] = r;
    s[1] += r * r;
  }
  for (int i = t.rank(); i < p.nv; i += t.size()) {
    const double y = x.y[i], v = x.v[i];
    const double ys = y + sigma * (v - xbar.v[i]);
    const double r = pfb(ys, v, alpha);
    ri.v[i] = r;
    s[2] += r * r;
    const double n = pnr(y, v, alpha);
    s[5] += n * n;
  }
  team_sum(t, s);
  t.sync();
  EvalOut e;
  // Norm() = sqrt(znorm^2 + lnorm^2 + vnorm^2) with cached component norms
  // (full_residual.cc:40-42,71-73).
  const double zn = sqrt(s[0]), ln = sqrt(s[1]), vn = sqrt(s[2]);
  e.Ei = sqrt(zn * zn + ln * ln + vn * vn);
  const double zo = sqrt(s[3]), lo = sqrt(s[4]), vo = sqrt(s[5]);
  e.Eo = sqrt(zo * zo + lo * lo + vo * vo);
  return e;
}

// tools::saturate (tools/utilities.h:19-28); *bad is set where it would throw.
__device__ __forceinline__ double saturate(double x, double a, double b,
                                           bool* bad) {
  if (a > b) *bad = true;
  return fmax(fmin(x, b), a);
}

struct Buffers {
  Vars xk, xi, xp, dx;
  Resid ri;
};

// One step of iterative refinement of the Newton system V dx = -r just solved
// (abstract_components.h:335-337 lists it as a TODO; off by default).  On entry
// `keep` (the trial-point buffer, free at this time) holds r in z, l, v -- saved
// before the solve, which may overwrite its argument -- and dx the solution with
// dx.y = b - A dz.  rho = -r - V dx with
//   V = [H + sigma I, G', A'; -G, sigma I, 0; -gamma A, 0, mu]
// (reference components/dense_cholesky_solver.h:49-62); the same factors solve
// V ddx = rho and dx += ddx.  The MPC policy parks scratch in the dx buffer while
// it solves, so the correction is solved INTO dx after dx has moved to `keep`.
template <class P>
__device__ __forceinline__ void refine_step(const Team& t, P& p, double sigma, const Resid& ri,
                                            const Vars& keep, const Vars& dx) {
  // ri <- (f + H dz + G' dl + A' dv, h - G dz)
  p.kkt(t, dx, ri.z, ri.l);
  // ri <- r + V dx: solve() returns the solution for MINUS its argument, i.e. ddx
  for (int i = t.rank(); i < p.nz; i += t.size())
    ri.z[i] = keep.z[i] + ((ri.z[i] - p.fvec(i)) + sigma * dx.z[i]);
  for (int i = t.rank(); i < p.nl; i += t.size())
    ri.l[i] = keep.l[i] + (sigma * dx.l[i] + (ri.l[i] - p.hvec(i)));
  for (int i = t.rank(); i < p.nv; i += t.size()) {
    const double adz = p.b(i) - dx.y[i];
    ri.v[i] = keep.v[i] + (p.mus[i] * dx.v[i] - p.gamma[i] * adz);
  }
  t.sync();
  vars_copy(t, p, dx, keep);           // r is used up: keep <- dx
  p.solve(t, ri.z, ri.l, ri.v, dx);    // dx <- ddx
  for (int i = t.rank(); i < p.nz; i += t.size()) dx.z[i] = keep.z[i] + dx.z[i];
  for (int i = t.rank(); i < p.nl; i += t.size()) dx.l[i] = keep.l[i] + dx.l[i];
  for (int i = t.rank(); i < p.nv; i += t.size()) {
    dx.v[i] = keep.v[i] + dx.v[i];
    dx.y[i] = (keep.y[i] + dx.y[i]) - p.b(i);  // b - A (dz + ddz)
  }
  t.sync();
}

// Solves one instance.  (z0,l0,v0): warm start in global memory, overwritten
// with the result together with y0.  Returns through *out.
template <class P>
__device__ void solve_instance(const Team& t, P& p, const fbstab_options& o,
                               Buffers w, double* z0, double* l0, double* v0,
                               double* y0, fbstab_out* out) {
  Vars xk = w.xk, xi = w.xi, xp = w.xp, dx = w.dx;
  const Resid ri = w.ri;
  const double sigma = o.sigma0;  // constant for the whole solve, impl:136
  const double alpha = o.alpha;
  const double combo_tol = o.abs_tol + o.rel_tol * (1.0 + p.forcing_norm(t));

  // CopyIntoVariable, impl:334-347: the incoming y is ignored.
  for (int i = t.rank(); i < p.nz; i += t.size()) xk.z[i] = z0[i];
  for (int i = t.rank(); i < p.nl; i += t.size()) xk.l[i] = l0[i];
  for (int i = t.rank(); i < p.nv; i += t.size()) xk.v[i] = v0[i];
  t.sync();
  p.margin(t, xk.z, xk.y);

  // dx_.Fill(1.0), impl:142 -> ||dx|| = sqrt(nz+nl+nv): the first stall test
  // cannot fire.
  double dx_norm = sqrt((double)p.nz + (double)p.nl + (double)p.nv);

  int eflag = FBSTAB_MAXITERATIONS;
  int status = FBSTAB_STATUS_OK;
  int newton = 0, prox = 0, backtracks = 0, evals = 0;
  double E0 = 0.0, Ek = 0.0, last_rk = 0.0, inner_tol = 0.0;
  const Vars* result = &xk;
  bool done = false;

  for (int k = 0; k < o.max_prox_iters && !done; k++) {
    // rk_->PenalizedNaturalResidual(xk), impl:144 / impl:162 (and, because
    // xi == xk here, also the first InnerResidual of the subproblem).
    EvalOut e = evaluate(t, p, xk, xk, sigma, alpha, ri);
    evals++;
    Ek = e.Eo;
    last_rk = Ek;
    if (k == 0) {
      E0 = Ek;
      bool bad = false;
      inner_tol = saturate(E0, o.inner_tol_min, o.inner_tol_max, &bad);
      if (bad) {
        status = FBSTAB_STATUS_SATURATE;
        break;
      }
    }
    if (Ek <= combo_tol || dx_norm <= o.stall_tol) {  // impl:164
      eflag = FBSTAB_SUCCESS;
      result = &xk;
      done = true;
      break;
    }
    {
      bool bad = false;
      inner_tol = saturate(inner_tol * o.delta, o.inner_tol_min, Ek, &bad);
      if (bad) {  // impl:179-180 would throw
        status = FBSTAB_STATUS_SATURATE;
        break;
      }
    }

    // ---- SolveProximalSubproblem(xi, xk, inner_tol, sigma, Ek), impl:229-304
    vars_copy(t, p, xk, xi);
    double merit[5] = {0, 0, 0, 0, 0};  // impl:233
    double Eo = 0.0;
    bool have = true;  // (Ei_c, Eo_c, ri) valid for the current xi
    double Ei_c = e.Ei, Eo_c = e.Eo;
    for (int i = 0; i < o.max_inner_iters; i++) {
      if (!have) {
        EvalOut ee = evaluate(t, p, xi, xk, sigma, alpha, ri);
        evals++;
        Ei_c = ee.Ei;
        Eo_c = ee.Eo;
        have = true;
      }
      const double Ei = Ei_c;
      Eo = Eo_c;
      last_rk = Eo;
      if ((Ei <= inner_tol && Eo < Ek) || (Ei <= o.inner_tol_min)) break;  // impl:250
      if (newton >= o.max_newton_iters) break;                              // impl:258

      double sigma_ls = sigma;
      {
        bool ok = p.factor(t, xi, xk, sigma, alpha);  // impl:263-267
        // regularise and retry (riccati_linear_solver.cc:129-130 lists it as a TODO):
        // sigma x 100 per attempt, in the linear solver only; default 0 attempts
        for (int j = 0; !ok && j < o.regularize_retries; j++) {
          sigma_ls *= 100.0;
          ok = p.factor(t, xi, xk, sigma_ls, alpha);
        }
        if (!ok) {
          status = FBSTAB_STATUS_FACTOR_FAILED;
          done = true;
          break;
        }
      }
      if (o.refine_steps > 0) {
        // the policies may overwrite r while solving: keep it (xp is free until the
        // line search builds the first trial point)
        for (int i = t.rank(); i < p.nz; i += t.size()) xp.z[i] = ri.z[i];
        for (int i = t.rank(); i < p.nl; i += t.size()) xp.l[i] = ri.l[i];
        for (int i = t.rank(); i < p.nv; i += t.size()) xp.v[i] = ri.v[i];
        t.sync();
      }
      p.solve(t, ri.z, ri.l, ri.v, dx);  // solves V dx = -r, impl:268-274
      if (o.refine_steps > 0) refine_step(t, p, sigma_ls, ri, xp, dx);
      newton++;

      const double current_merit = 0.5 * Ei * Ei;  // impl:278
#pragma unroll
      for (int m = 4; m > 0; m--) merit[m] = merit[m - 1];  // impl:402-409
      merit[0] = current_merit;
      double m0 = current_merit;
      if (o.nonmonotone_linesearch) {
#pragma unroll
        for (int m = 1; m < 5; m++) m0 = fmax(m0, merit[m]);
      }
      double tstep = 1.0;
      bool accepted = false;
      EvalOut et;
      for (int j = 0; j < o.max_linesearch_iters; j++) {  // impl:283-297
        vars_axpy(t, p, xi, tstep, dx, xp);
        et = evaluate(t, p, xp, xk, sigma, alpha, ri);
        evals++;
        const double mp = 0.5 * et.Ei * et.Ei;
        if (mp <= m0 - 2.0 * tstep * o.eta * current_merit) {
          accepted = true;
          break;
        }
        tstep *= o.beta;
        backtracks++;
      }
      if (accepted) {
        // x + t*dx is bit-identical to the accepted trial point.
        const Vars tmp = xi;
        xi = xp;
        xp = tmp;
        Ei_c = et.Ei;
        Eo_c = et.Eo;
      } else {
        // every trial failed: the step is still taken with the shrunken t
        // (impl:295-298); its residual must be evaluated afresh.
        vars_axpy(t, p, xi, tstep, dx, xi);
        have = false;
      }
    }
    if (done) break;
    // ProjectDuals, impl:301 / full_variable.cc:75
    for (int i = t.rank(); i < p.nv; i += t.size()) xi.v[i] = fmax(xi.v[i], 0.0);
    t.sync();

    // Newton iteration cap, impl:188-199
    if (newton >= o.max_newton_iters) {
      const Vars& pick = (Eo < Ek) ? xi : xk;
      EvalOut ef = evaluate(t, p, pick, pick, sigma, alpha, ri);
      evals++;
      last_rk = ef.Eo;
      eflag = FBSTAB_MAXITERATIONS;
      result = (Eo < Ek) ? &xi : &xk;
      done = true;
      break;
    }

    // dx = xi - xk (y-aware), impl:202-203
    {
      double s[3] = {0, 0, 0};
      for (int i = t.rank(); i < p.nz; i += t.size()) {
        const double d = xi.z[i] + (-1.0) * xk.z[i];
        dx.z[i] = d;
        s[0] += d * d;
      }
      for (int i = t.rank(); i < p.nl; i += t.size()) {
        const double d = xi.l[i] + (-1.0) * xk.l[i];
        dx.l[i] = d;
        s[1] += d * d;
      }
      for (int i = t.rank(); i < p.nv; i += t.size()) {
        const double d = xi.v[i] + (-1.0) * xk.v[i];
        dx.v[i] = d;
        s[2] += d * d;
        const double y = xi.y[i] + (-1.0) * xk.y[i];
        dx.y[i] = y + p.b(i);
      }
      team_sum(t, s);
      t.sync();
      const double a = sqrt(s[0]), b = sqrt(s[1]), c = sqrt(s[2]);
      dx_norm = sqrt(a * a + b * b + c * c);  // full_variable.cc:77-83
    }
    if (o.check_feasibility) {  // impl:204-212, 385-400
      const int feas = p.feasibility(t, dx, o.infeas_tol);
      if (feas != 0) {
        eflag = (feas == 1)   ? FBSTAB_PRIMAL_INFEASIBLE
                : (feas == 2) ? FBSTAB_DUAL_INFEASIBLE
                              : FBSTAB_PRIMAL_DUAL_INFEASIBLE;
        result = &dx;  // the certificate is what the caller receives, impl:209
        done = true;
        break;
      }
    }
    // xk <- xi, impl:215-216
    {
      const Vars tmp = xk;
      xk = xi;
      xi = tmp;
      result = &xk;
    }
    prox++;
  }

  // WriteVariable, impl:349-360
  const Vars r = *result;
  for (int i = t.rank(); i < p.nz; i += t.size()) z0[i] = r.z[i];
  for (int i = t.rank(); i < p.nl; i += t.size()) l0[i] = r.l[i];
  for (int i = t.rank(); i < p.nv; i += t.size()) {
    v0[i] = r.v[i];
    y0[i] = r.y[i];
  }
  if (t.rank() == 0) {  // PrepareOutput, impl:362-383
    out->eflag = eflag;
    out->newton_iters = newton;
    out->prox_iters = prox;
    out->status = status;
    out->residual = last_rk;
    out->initial_residual = E0;
    out->solve_time = -1.0;
    out->ls_backtracks = backtracks;
    out->residual_evals = evals;
  }
  t.sync();
}

}  // namespace fbs

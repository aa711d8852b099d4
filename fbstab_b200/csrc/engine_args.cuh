// engine_args.cuh -- kernel argument block shared by the CTA-per-instance
// kernels (api.cu, mpc_riccati.cu) and the component-stage runner used by the
// per-kernel parity tests.
#pragma once

#include "common.cuh"
#include "engine.cuh"
#include "fbstab_b200.h"

namespace fbs {

struct CommonArgs {
  int batch;
  double *z, *l, *v, *y;
  fbstab_out* out;
  double* ws;        // per-CTA workspace base
  size_t ws_stride;  // doubles per CTA
  int* counter;
  int vec_in_smem;
  fbstab_options opts;
  // component mode
  int comp;
  fbstab_component_io io;
};

// One component stage on caller-supplied iterates (per-kernel parity tests).
template <class P>
__device__ void RunComponent(const fbs::Team& t, P& p, const CommonArgs& c,
                             int inst, fbs::Buffers& w) {
  const fbstab_component_io& io = c.io;
  const size_t oz = (size_t)inst * p.nz, ol = (size_t)inst * p.nl,
               ov = (size_t)inst * p.nv;
  const double alpha = c.opts.alpha;
  if (c.comp == FBSTAB_COMP_MARGIN) {
    p.margin(t, io.z + oz, io.dy + ov);
    return;
  }
  // load x (and xbar) into the work buffers
  for (int i = t.rank(); i < p.nz; i += t.size()) {
    w.xi.z[i] = io.z[oz + i];
    w.xk.z[i] = io.zbar ? io.zbar[oz + i] : io.z[oz + i];
  }
  for (int i = t.rank(); i < p.nl; i += t.size()) {
    w.xi.l[i] = io.l[ol + i];
    w.xk.l[i] = io.lbar ? io.lbar[ol + i] : io.l[ol + i];
  }
  for (int i = t.rank(); i < p.nv; i += t.size()) {
    w.xi.v[i] = io.v[ov + i];
    w.xk.v[i] = io.vbar ? io.vbar[ov + i] : io.v[ov + i];
    w.xi.y[i] = io.y ? io.y[ov + i] : 0.0;
  }
  t.sync();
  if (c.comp == FBSTAB_COMP_RESIDUAL) {
    fbs::EvalOut e = fbs::evaluate(t, p, w.xi, w.xk, io.sigma, alpha, w.ri);
    (void)e;
    double s[6] = {0, 0, 0, 0, 0, 0};
    // recompute component norms for reporting (evaluate() returns the totals)
    for (int i = t.rank(); i < p.nz; i += t.size()) {
      const double r = w.ri.z[i];
      io.rz[oz + i] = r;
      s[0] += r * r;
      const double n = r - io.sigma * (w.xi.z[i] - w.xk.z[i]);
      s[3] += n * n;
    }
    for (int i = t.rank(); i < p.nl; i += t.size()) {
      const double r = w.ri.l[i];
      io.rl[ol + i] = r;
      s[1] += r * r;
      const double n = r - io.sigma * (w.xi.l[i] - w.xk.l[i]);
      s[4] += n * n;
    }
    for (int i = t.rank(); i < p.nv; i += t.size()) {
      const double r = w.ri.v[i];
      io.rv[ov + i] = r;
      s[2] += r * r;
      const double n = fbs::pnr(w.xi.y[i], w.xi.v[i], alpha);
      s[5] += n * n;
    }
    fbs::team_sum(t, s);
    if (t.rank() == 0 && io.norms) {
      for (int k = 0; k < 6; k++) io.norms[(size_t)inst * 8 + k] = sqrt(s[k]);
      io.norms[(size_t)inst * 8 + 6] = e.Ei;
      io.norms[(size_t)inst * 8 + 7] = e.Eo;
    }
  } else if (c.comp == FBSTAB_COMP_NEWTON) {
    const bool ok = p.factor(t, w.xi, w.xk, io.sigma, alpha);
    // engine convention: solve() receives the residual and solves for -r; the
    // component API takes the right-hand side r itself, so negate on load.
    for (int i = t.rank(); i < p.nz; i += t.size()) w.ri.z[i] = -io.rz[oz + i];
    for (int i = t.rank(); i < p.nl; i += t.size()) w.ri.l[i] = -io.rl[ol + i];
    for (int i = t.rank(); i < p.nv; i += t.size()) w.ri.v[i] = -io.rv[ov + i];
    t.sync();
    p.solve(t, w.ri.z, w.ri.l, w.ri.v, w.dx);
    for (int i = t.rank(); i < p.nz; i += t.size()) io.dz[oz + i] = w.dx.z[i];
    for (int i = t.rank(); i < p.nl; i += t.size()) io.dl[ol + i] = w.dx.l[i];
    for (int i = t.rank(); i < p.nv; i += t.size()) {
      io.dv[ov + i] = w.dx.v[i];
      io.dy[ov + i] = w.dx.y[i];
      if (io.gamma) io.gamma[ov + i] = p.gamma[i];
      if (io.mus) io.mus[ov + i] = p.mus[i];
    }
    if (t.rank() == 0 && io.status)
      io.status[inst] = ok ? FBSTAB_STATUS_OK : FBSTAB_STATUS_FACTOR_FAILED;
  } else if (c.comp == FBSTAB_COMP_FEAS) {
    const int feas = p.feasibility(t, w.xi, io.tol);
    if (t.rank() == 0 && io.status) io.status[inst] = feas;
  }
  t.sync();
}


}  // namespace fbs

// ref_wrapper.cpp -- C entry points around the REFERENCE's own classes (TEST INFRASTRUCTURE).
//
// Compiled by `make -C oracle _ref` TOGETHER WITH the reference's unmodified sources, read
// where they lie under /root/reference (fbstab/fbstab_dense.cc, fbstab/fbstab_mpc.cc,
// fbstab/components/*.cc, the templated fbstab/fbstab_algorithm-impl.h), against
// oracle/eigen_shim (a stand-in for the Eigen API subset they use; real Eigen is not in this
// image) into oracle/_ref/libfbstab_ref.so.  Nothing of the reference is copied into this
// repository.  This file only marshals flat batch arrays -- the layout of
// oracle_dense_solve_batch / oracle_mpc_solve_batch (fbstab_oracle.h) -- into
// FBstabDense::ProblemDataRef / FBstabMpc::ProblemDataRef and calls Solve
// (fbstab/fbstab_dense.h:136-149, fbstab/fbstab_mpc.h:181-195), one solver object per thread
// (the reference's objects are not thread safe, tools/copyable_macros.h:16-20).
#include <chrono>
#include <exception>
#include <memory>
#include <new>
#include <thread>
#include <vector>

#include "fbstab/fbstab_dense.h"
#include "fbstab/fbstab_mpc.h"
#include "fbstab/test/ocp_generator.h"
#include "fbstab_oracle.h"
#include "tools/matrix_sequence.h"

namespace {

template <class Options>
void SetOptions(const oracle_options* o, Options* p) {
  p->sigma0 = o->sigma0;
  p->sigma_max = o->sigma_max;
  p->sigma_min = o->sigma_min;
  p->alpha = o->alpha;
  p->beta = o->beta;
  p->eta = o->eta;
  p->delta = o->delta;
  p->gamma = o->gamma;
  p->abs_tol = o->abs_tol;
  p->rel_tol = o->rel_tol;
  p->stall_tol = o->stall_tol;
  p->infeas_tol = o->infeas_tol;
  p->inner_tol_max = o->inner_tol_max;
  p->inner_tol_min = o->inner_tol_min;
  p->max_newton_iters = o->max_newton_iters;
  p->max_prox_iters = o->max_prox_iters;
  p->max_inner_iters = o->max_inner_iters;
  p->max_linesearch_iters = o->max_linesearch_iters;
  p->check_feasibility = o->check_feasibility != 0;
  p->nonmonotone_linesearch = o->nonmonotone_linesearch != 0;
  p->display_level = fbstab::Display::OFF;
}

void Record(const fbstab::SolverOut& s, oracle_out* out) {
  out->eflag = static_cast<int>(s.eflag);
  out->newton_iters = s.newton_iters;
  out->prox_iters = s.prox_iters;
  out->status = 0;
  out->residual = s.residual;
  out->initial_residual = s.initial_residual;
  out->solve_time = s.solve_time;
  out->ls_backtracks = -1;   // the reference does not count them
  out->residual_evals = -1;
}

template <class F>
void ForRanges(int batch, int nthreads, F f) {
  nthreads = std::max(1, std::min(nthreads, batch));
  const int per = (batch + nthreads - 1) / nthreads;
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; t++) {
    const int lo = t * per, hi = std::min(batch, lo + per);
    if (lo >= hi) break;
    th.emplace_back([=] { f(lo, hi); });
  }
  for (auto& t : th) t.join();
}

}  // namespace

extern "C" {

// 1 if the library was built from the reference's sources (always; the symbol is the probe).
int ref_available(void) { return 1; }

int ref_dense_solve_batch(int nz, int nl, int nv, int batch, const double* H, const double* f,
                          const double* G, const double* h, const double* A, const double* b,
                          double* z, double* l, double* v, double* y, const oracle_options* o,
                          oracle_out* out, int nthreads) {
  ForRanges(batch, nthreads, [=](int lo, int hi) {
    try {
      fbstab::FBstabDense solver(nz, nl, nv);
      if (o) {
        fbstab::FBstabDense::Options opts = fbstab::FBstabDense::DefaultOptions();
        SetOptions(o, &opts);
        solver.UpdateOptions(opts);
      }
      for (int i = lo; i < hi; i++) {
        const size_t k = (size_t)i;
        Eigen::Map<Eigen::MatrixXd> Hm(H + k * nz * nz, nz, nz), Gm(G + k * nl * nz, nl, nz),
            Am(A + k * nv * nz, nv, nz);
        Eigen::Map<Eigen::VectorXd> fm(f + k * nz, nz), hm(h + k * nl, nl), bm(b + k * nv, nv);
        Eigen::Map<Eigen::VectorXd> zm(z + k * nz, nz), lm(l + k * nl, nl), vm(v + k * nv, nv),
            ym(y + k * nv, nv);
        fbstab::FBstabDense::ProblemDataRef qp(&Hm, &fm, &Gm, &hm, &Am, &bm);
        fbstab::FBstabDense::VariableRef x(&zm, &lm, &vm, &ym);
        try {
          Record(solver.Solve(qp, &x), out + i);
        } catch (const std::exception&) {
          out[i] = oracle_out();
          out[i].eflag = 2;
          out[i].status = 1;  // the reference threw from inside Solve
        }
      }
    } catch (const std::exception&) {
      for (int i = lo; i < hi; i++) {
        out[i] = oracle_out();
        out[i].status = 3;
      }
    }
  });
  return 0;
}

int ref_mpc_solve_batch(int N, int nx, int nu, int nc, int batch, const double* Q,
                        const double* R, const double* S, const double* q, const double* r,
                        const double* A, const double* B, const double* c, const double* E,
                        const double* L, const double* d, const double* x0, double* z, double* l,
                        double* v, double* y, const oracle_options* o, oracle_out* out,
                        int nthreads) {
  const size_t nz = (size_t)(N + 1) * (nx + nu), nl = (size_t)(N + 1) * nx,
               nv = (size_t)(N + 1) * nc;
  ForRanges(batch, nthreads, [=](int lo, int hi) {
    try {
      fbstab::FBstabMpc solver(N, nx, nu, nc);
      if (o) {
        fbstab::FBstabMpc::Options opts = fbstab::FBstabMpc::DefaultOptions();
        SetOptions(o, &opts);
        solver.UpdateOptions(opts);
      }
      for (int i = lo; i < hi; i++) {
        const size_t k = (size_t)i, M = (size_t)N + 1;
        fbstab::FBstabMpc::ProblemDataRef qp;
        qp.Q = fbstab::MapMatrixSequence(Q + k * M * nx * nx, N + 1, nx, nx);
        qp.R = fbstab::MapMatrixSequence(R + k * M * nu * nu, N + 1, nu, nu);
        qp.S = fbstab::MapMatrixSequence(S + k * M * nu * nx, N + 1, nu, nx);
        qp.q = fbstab::MapMatrixSequence(q + k * M * nx, N + 1, nx, 1);
        qp.r = fbstab::MapMatrixSequence(r + k * M * nu, N + 1, nu, 1);
        qp.A = fbstab::MapMatrixSequence(A + k * N * nx * nx, N, nx, nx);
        qp.B = fbstab::MapMatrixSequence(B + k * N * nx * nu, N, nx, nu);
        qp.c = fbstab::MapMatrixSequence(c + k * N * nx, N, nx, 1);
        qp.E = fbstab::MapMatrixSequence(E + k * M * nc * nx, N + 1, nc, nx);
        qp.L = fbstab::MapMatrixSequence(L + k * M * nc * nu, N + 1, nc, nu);
        qp.d = fbstab::MapMatrixSequence(d + k * M * nc, N + 1, nc, 1);
        Eigen::Map<const Eigen::VectorXd> x0m(x0 + k * nx, nx);
        qp.SetX0(x0m);
        Eigen::Map<Eigen::VectorXd> zm(z + k * nz, (int)nz), lm(l + k * nl, (int)nl),
            vm(v + k * nv, (int)nv), ym(y + k * nv, (int)nv);
        fbstab::FBstabMpc::VariableRef x(zm, lm, vm, ym);
        try {
          Record(solver.Solve(qp, &x), out + i);
        } catch (const std::exception&) {
          out[i] = oracle_out();
          out[i].eflag = 2;
          out[i].status = 1;
        }
      }
    } catch (const std::exception&) {
      for (int i = lo; i < hi; i++) {
        out[i] = oracle_out();
        out[i].status = 3;
      }
    }
  });
  return 0;
}

// The reference's own OCP generator (fbstab/test/ocp_generator.cc:73-421): the eleven
// sequences and x0 of kind 0 DoubleIntegrator, 1 ServoMotor, 2 SpacecraftRelativeMotion,
// 3 CopolymerizationReactor at horizon N, in the wire format -- the cross-check of the
// engine's restated generator (fbstab_b200/csrc/problems.cpp, fbstab_ocp_generate).
// sizes (may be NULL): nx, nu, nc.  Any output pointer may be NULL.
int ref_ocp_generate(int kind, int N, int* sizes, double* Q, double* R, double* S, double* q,
                     double* r, double* A, double* B, double* c, double* E, double* L, double* d,
                     double* x0) {
  try {
    fbstab::test::OcpGenerator g;
    if (kind == 0) g.DoubleIntegrator(N);
    else if (kind == 1) g.ServoMotor(N);
    else if (kind == 2) g.SpacecraftRelativeMotion(N);
    else if (kind == 3) g.CopolymerizationReactor(N);
    else return 1;
    const fbstab::FBstabMpc::ProblemDataRef p = g.GetFBstabInputRef();
    if (sizes) {
      sizes[0] = p.Q.rows();
      sizes[1] = p.R.rows();
      sizes[2] = p.E.rows();
    }
    const fbstab::MapMatrixSequence* seq[11] = {&p.Q, &p.R, &p.S, &p.q, &p.r, &p.A,
                                                &p.B, &p.c, &p.E, &p.L, &p.d};
    double* dst[11] = {Q, R, S, q, r, A, B, c, E, L, d};
    for (int k = 0; k < 11; k++)
      if (dst[k]) std::copy(seq[k]->data(), seq[k]->data() + seq[k]->size(), dst[k]);
    if (x0) std::copy(p.x0.data(), p.x0.data() + p.x0.size(), x0);
    return 0;
  } catch (const std::exception&) {
    return 2;
  }
}

}  // extern "C"

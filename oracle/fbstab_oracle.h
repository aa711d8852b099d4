/*
 * fbstab_oracle.h -- C interface of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  The oracle is a sequential CPU restatement of the
 * reference algorithm (dliaomcp/fbstab) used to check the CUDA engine.  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load it.  Nothing under fbstab_b200/ links, imports or
 * calls it.
 *
 * Pinned (a) by every golden vector of the reference's own tests
 * (tests/test_oracle_goldens.py) and (b) by the reference's OWN algorithm sources
 * compiled here against a stand-in for Eigen (oracle/_ref, `make -C oracle _ref`,
 * ref_wrapper.cpp, eigen_shim/): same exit flags and iteration counts on every
 * instance of every benchmark family, the same bytes on the MPC path
 * (tests/test_reference_sources.py).
 *
 * Layout conventions are the reference's: dense matrices column-major
 * (Eigen default), MPC sequences `len x rows x cols` contiguous with each
 * matrix column-major (reference tools/matrix_sequence.h:81-83).
 */
#ifndef FBSTAB_ORACLE_H_
#define FBSTAB_ORACLE_H_

#ifdef __cplusplus
extern "C" {
#endif

/* Mirrors AlgorithmParameters, reference fbstab/fbstab_algorithm.h:48-82. */
typedef struct oracle_options {
  double sigma0, sigma_max, sigma_min;
  double alpha, beta, eta, delta, gamma;
  double abs_tol, rel_tol, stall_tol, infeas_tol;
  double inner_tol_max, inner_tol_min;
  int max_newton_iters, max_prox_iters, max_inner_iters, max_linesearch_iters;
  int check_feasibility, nonmonotone_linesearch, display_level;
  /* Linear-solver extensions the reference lists as missing (both default 0 = the
   * reference's behaviour): steps of iterative refinement of every Newton system
   * (abstract_components.h:335-337) and regularise-and-retry attempts when a
   * factorisation fails (riccati_linear_solver.cc:129-130). */
  int refine_steps, regularize_retries;
} oracle_options;

/* Mirrors SolverOut (fbstab_algorithm.h:30-37) + trajectory counters. */
typedef struct oracle_out {
  int eflag;            /* ExitFlag 0..5, fbstab_algorithm.h:17-24 */
  int newton_iters;
  int prox_iters;
  int status;           /* 0 ok, 1 factor failed, 2 saturate(lower>upper), 3 solve failed */
  double residual;
  double initial_residual;
  double solve_time;    /* seconds, wall clock */
  int ls_backtracks;    /* rejected Armijo trials, total */
  int residual_evals;   /* InnerResidual + (Penalized)NaturalResidual calls */
} oracle_out;

/* One trajectory record = 8 doubles:
 *  kind 0 (prox top):   {0, prox, newton, Ek, inner_tol, 0, 0, 0}
 *  kind 1 (newton step):{1, prox, newton_after, Ei, Eo, t, backtracks, inner_i} */
#define ORACLE_TRAJ_STRIDE 8

void oracle_default_options(oracle_options* o);   /* fbstab_algorithm-impl.h:33-59 */
void oracle_reliable_options(oracle_options* o);  /* fbstab_algorithm-impl.h:61-74 */
/* fbstab_algorithm-impl.h:7-31; returns 0, or 2 if a saturate() would throw. */
int oracle_validate_options(oracle_options* o);

typedef struct oracle_problem oracle_problem;

/* The problem object BORROWS every pointer (like DenseData / MpcData). */
oracle_problem* oracle_dense_create(int nz, int nl, int nv, const double* H,
                                    const double* f, const double* G,
                                    const double* h, const double* A,
                                    const double* b);
oracle_problem* oracle_mpc_create(int N, int nx, int nu, int nc,
                                  const double* Q, const double* R,
                                  const double* S, const double* q,
                                  const double* r, const double* A,
                                  const double* B, const double* c,
                                  const double* E, const double* L,
                                  const double* d, const double* x0);
void oracle_destroy(oracle_problem* p);
void oracle_sizes(const oracle_problem* p, int* nz, int* nl, int* nv);
double oracle_forcing_norm(const oracle_problem* p);

/* Data ops.  op: 0 gemvH, 1 gemvA, 2 gemvAT, 3 gemvG, 4 gemvGT (y <- a*M*x + b*y)
 *            which: 0 axpyf, 1 axpyh, 2 axpyb (y <- a*w + y). */
int oracle_gemv(const oracle_problem* p, int op, const double* x, double a,
                double b, double* y);
int oracle_axpy(const oracle_problem* p, int which, double a, double* y);

/* y = b - A z  (FullVariable::InitializeConstraintMargin). */
void oracle_margin(const oracle_problem* p, const double* z, double* y);
/* x <- x + a*dx including the y-aware rule (FullVariable::axpy). */
void oracle_variable_axpy(const oracle_problem* p, double a, const double* dz,
                          const double* dl, const double* dv, const double* dy,
                          double* z, double* l, double* v, double* y);

/* kind: 0 InnerResidual(x,xbar,sigma), 1 NaturalResidual(x),
 *       2 PenalizedNaturalResidual(x).  y / ybar are inputs.  norms[3]. */
void oracle_residual(const oracle_problem* p, int kind, double alpha,
                     double sigma, const double* z, const double* l,
                     const double* v, const double* y, const double* zbar,
                     const double* lbar, const double* vbar, double* rz,
                     double* rl, double* rv, double* norms);

/* LinearSolver::Initialize(x,xbar,sigma) then ::Solve(r,&dx).
 * variant (dense only): 0 = Eigen-style diagonally pivoted LDLT on K (reference),
 *                       1 = same LDLT without pivoting,
 *                       2 = Cholesky of E + Cholesky of the Schur complement,
 *                       3 = unpivoted Gauss-Jordan of E + Schur complement
 *                           (the elimination order of the warp kernel).
 * returns 0 ok, 1 factor failed. gamma/mus may be NULL. */
int oracle_linear_solve(const oracle_problem* p, int variant, double alpha,
                        double sigma, const double* z, const double* l,
                        const double* v, const double* y, const double* zbar,
                        const double* lbar, const double* vbar,
                        const double* rz, const double* rl, const double* rv,
                        double* dz, double* dl, double* dv, double* dy,
                        double* gamma, double* mus);

/* FullFeasibility::CheckFeasibility: 0 FEASIBLE, 1 PRIMAL_INFEASIBLE(status),
 * 2 DUAL_INFEASIBLE, 3 BOTH.  (enum order of full_feasibility.h) */
int oracle_feasibility(const oracle_problem* p, const double* dz,
                       const double* dl, const double* dv, double tol);

/* FBstabAlgorithm::Solve.  z,l,v are the warm start in / solution out, y out.
 * traj may be NULL; at most traj_cap records are written, *traj_len gets the
 * number produced.  Returns out->status. */
int oracle_solve(const oracle_problem* p, int variant, const oracle_options* o,
                 double* z, double* l, double* v, double* y, oracle_out* out,
                 double* traj, int traj_cap, int* traj_len);

/* Sparse QPs (FBstabSparse; the reference plans them, ROADMAP.md:10): H by its upper
 * triangle, G and A, all compressed-column; perm (nz+nl+nv, perm[new] = old over
 * [z; l; w]) is the elimination order of the Newton matrix, NULL = natural. */
oracle_problem* oracle_sparse_create(int nz, int nl, int nv, const int* Hp, const int* Hi,
                                     const double* Hx, const double* f, const int* Gp,
                                     const int* Gi, const double* Gx, const double* h,
                                     const int* Ap, const int* Ai, const double* Ax,
                                     const double* b, const int* perm);
/* QdldlWrapper (tools/qdldl/qdldl_wrapper.h:19-84): factor the upper-triangular CSC
 * matrix and solve in place.  Returns 0, -1 on a zero pivot / bad pattern. */
int oracle_qdldl_solve(int n, const int* Ap, const int* Ai, const double* Ax, double* x);
int oracle_sparse_solve_batch(int nz, int nl, int nv, int batch, const int* Hp, const int* Hi,
                              const double* Hx, const double* f, const int* Gp, const int* Gi,
                              const double* Gx, const double* h, const int* Ap, const int* Ai,
                              const double* Ax, const double* b, const int* perm, double* z,
                              double* l, double* v, double* y, const oracle_options* o,
                              oracle_out* out, int nthreads);

/* Batched convenience for timing: instance-major arrays, nthreads host threads,
 * one solver per thread (static contiguous partition). Returns 0. */
int oracle_dense_solve_batch(int nz, int nl, int nv, int batch, const double* H,
                             const double* f, const double* G, const double* h,
                             const double* A, const double* b, double* z,
                             double* l, double* v, double* y,
                             const oracle_options* o, oracle_out* out,
                             int variant, int nthreads);
int oracle_mpc_solve_batch(int N, int nx, int nu, int nc, int batch,
                           const double* Q, const double* R, const double* S,
                           const double* q, const double* r, const double* A,
                           const double* B, const double* c, const double* E,
                           const double* L, const double* d, const double* x0,
                           double* z, double* l, double* v, double* y,
                           const oracle_options* o, oracle_out* out,
                           int nthreads);

#ifdef __cplusplus
}
#endif
#endif  /* FBSTAB_ORACLE_H_ */

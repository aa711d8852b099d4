/*
 * fbstab_b200.h -- C-ABI of the B200-native batched FBstab engine.
 *
 * This is the drop-in boundary: plain C types only (pointers, sizes, PODs), no
 * torch / Eigen types.  The C++ facade in include/fbstab/ (FBstabDense,
 * FBstabMpc -- same names and signatures as the reference) is a thin host
 * layer over these entry points, and INTEGRATION.md shows the binding a
 * maintainer of the reference would add.
 * (include/fbstab/fbstab_sparse.h, FBstabSparse, is the solver the reference
 * plans: ROADMAP.md:10.)
 *
 * What each entry point replaces in the reference (paths relative to the
 * reference tree):
 *   fbstab_dense_batch_create   <- FBstabDense::FBstabDense(nz,nl,nv)
 *                                  fbstab/fbstab_dense.cc:18-42
 *   fbstab_dense_batch_solve    <- FBstabDense::Solve(qp,&x)
 *                                  fbstab/fbstab_dense.h:136-149 ->
 *                                  FBstabAlgorithm::Solve
 *                                  fbstab/fbstab_algorithm-impl.h:113-224
 *                                  (one call = `batch` independent Solve calls)
 *   fbstab_mpc_batch_create     <- FBstabMpc::FBstabMpc(N,nx,nu,nc)
 *                                  fbstab/fbstab_mpc.cc:61-89
 *   fbstab_mpc_batch_solve      <- FBstabMpc::Solve(qp,&x)
 *                                  fbstab/fbstab_mpc.h:181-195
 *   fbstab_mpc_batch_solve_shared / _lti
 *                               <- the same with ONE copy of the stage data /
 *                                  one stage, as CopyOverHorizon replicates it
 *                                  fbstab/test/ocp_generator.cc:397-418
 *   fbstab_mpc_closed_loop_*    <- the receding-horizon loop around Solve with
 *                                  OcpGenerator::GetSimulationInputs
 *                                  fbstab/test/ocp_generator.h:31-38,69
 *   fbstab_sparse_batch_create  <- QdldlWrapper::QdldlWrapper(n,Ap,Ai): the
 *                                  symbolic analysis of the planned sparse
 *                                  solver, tools/qdldl/qdldl_wrapper.h:24-44
 *   fbstab_sparse_batch_solve   <- FBstabAlgorithm::Solve over sparse data
 *                                  with QdldlWrapper::Factor / ::Solve
 *                                  (qdldl_wrapper.h:46-60; ROADMAP.md:10)
 *   fbstab_sparse_analyze       <- the same symbolic analysis on the host only
 *                                  (no handle, no device)
 *   fbstab_multi_gpu_*, fbstab_*_multi_gpu_solve
 *                               <- (none: the reference is single-threaded) the
 *                                  batch sharded by instance over the GPUs of a
 *                                  box, results gathered with NCCL
 *   fbstab_*_batch_set_options  <- UpdateOptions -> UpdateParameters +
 *                                  ValidateOptions
 *                                  fbstab/fbstab_algorithm-impl.h:7-31,308-332
 *   fbstab_default_options      <- AlgorithmParameters::DefaultParameters
 *                                  fbstab/fbstab_algorithm-impl.h:33-59
 *   fbstab_reliable_options     <- AlgorithmParameters::ReliableParameters
 *                                  fbstab/fbstab_algorithm-impl.h:61-74
 *   fbstab_*_batch_component    <- the component methods the algorithm is
 *                                  built from (fbstab/components/
 *                                  abstract_components.h:24-338); exposed so
 *                                  each kernel stage can be parity-tested
 *
 * Data layout (identical to the reference's, so a facade call is a memcpy):
 *   dense: per instance H (nz x nz), G (nl x nz), A (nv x nz) column-major
 *          (Eigen default, fbstab/components/dense_data.h:44-52), vectors
 *          f(nz) h(nl) b(nv); a batch is instance-major and contiguous:
 *          H[i] starts at H + i*nz*nz, and so on.
 *   mpc:   per instance the 11 sequences of fbstab/fbstab_mpc.h:67-83, each
 *          `len x rows x cols` contiguous, column-major per matrix
 *          (tools/matrix_sequence.h:81-83); Q,R,S,q,r,E,L,d have N+1 entries,
 *          A,B,c have N; x0(nx).  A batch is instance-major and contiguous.
 *   sparse: ONE pattern per handle (H by its upper triangle, G, A in
 *          compressed-column form, as tools/qdldl/qdldl_wrapper.h:12-14), the
 *          values Hx, Gx, Ax and f, h, b per instance, instance-major.
 *   iterates: z(nz) l(nl) v(nv) y(nv) per instance, instance-major.
 *
 * Every data/iterate pointer may be a HOST pointer or a DEVICE pointer on the
 * handle's device (checked per pointer: a pointer that belongs to another
 * device is FBSTAB_ERR_INVALID).  Host buffers are staged through the handle's
 * device buffers inside the call; device buffers are used in place.
 * If every pointer is a device pointer the call only enqueues work on `stream`
 * (asynchronous); otherwise it returns after the results are in host memory.
 *
 * Streams and threads: a handle owns ONE instance counter and ONE set of
 * workspaces, so its launches are serialised -- each call makes its stream wait
 * (cudaStreamWaitEvent) for the handle's previous launch, whatever stream that
 * ran on.  Calls on one handle must not be issued from two host threads at the
 * same time (like the reference's solver objects, fbstab/components/
 * dense_cholesky_solver.h:27); distinct handles are independent.
 *
 * There is NO CPU fallback: with no usable CUDA device every create/solve call
 * fails with FBSTAB_ERR_NOGPU / FBSTAB_ERR_CUDA.
 */
#ifndef FBSTAB_B200_H_
#define FBSTAB_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- error codes (return value of every int function) ------------------- */
#define FBSTAB_OK 0
#define FBSTAB_ERR_INVALID 1 /* bad size / null pointer / batch > max_batch   */
#define FBSTAB_ERR_CUDA 2    /* a CUDA runtime call failed                    */
#define FBSTAB_ERR_NOGPU 3   /* no CUDA device: the engine has no CPU path    */
#define FBSTAB_ERR_ALLOC 4   /* device or host allocation failed              */
#define FBSTAB_ERR_NCCL 5    /* NCCL missing or a NCCL call failed            */

/* ---- ExitFlag, reference fbstab/fbstab_algorithm.h:17-24 ---------------- */
#define FBSTAB_SUCCESS 0
#define FBSTAB_DIVERGENCE 1
#define FBSTAB_MAXITERATIONS 2
#define FBSTAB_PRIMAL_INFEASIBLE 3
#define FBSTAB_DUAL_INFEASIBLE 4
#define FBSTAB_PRIMAL_DUAL_INFEASIBLE 5

/* ---- per-instance status (where the reference throws) ------------------- */
#define FBSTAB_STATUS_OK 0
#define FBSTAB_STATUS_FACTOR_FAILED 1 /* LinearSolver::Initialize returned false,
                                         fbstab_algorithm-impl.h:263-267       */
#define FBSTAB_STATUS_SATURATE 2      /* tools::saturate lower > upper,
                                         tools/utilities.h:21-25               */

/* AlgorithmParameters, reference fbstab/fbstab_algorithm.h:48-82. */
typedef struct fbstab_options {
  double sigma0, sigma_max, sigma_min;
  double alpha, beta, eta, delta, gamma;
  double abs_tol, rel_tol, stall_tol, infeas_tol;
  double inner_tol_max, inner_tol_min;
  int32_t max_newton_iters, max_prox_iters, max_inner_iters,
      max_linesearch_iters;
  int32_t check_feasibility, nonmonotone_linesearch, display_level;
  /* Linear-solver extensions the reference lists as missing; both default to 0,
   * which is the reference's behaviour (fbstab_default_options):
   *  refine_steps        0 or 1: one step of iterative refinement of every Newton
   *                      system V dx = -R: rho = -R - V dx, V ddx = rho with the
   *                      same factors, dx += ddx ("TODO: implement iterative
   *                      refinement", components/abstract_components.h:335-337)
   *  regularize_retries  a failed factorisation is repeated up to this many times
   *                      with sigma x 100 per attempt inside the linear solver
   *                      ("TODO: regularize and retry",
   *                      components/riccati_linear_solver.cc:129-130) before the
   *                      instance is given FBSTAB_STATUS_FACTOR_FAILED.
   * Either option selects the team-per-instance kernels (the warp / lane kernels
   * keep no factors to re-use). */
  int32_t refine_steps, regularize_retries;
} fbstab_options;

/* SolverOut, reference fbstab/fbstab_algorithm.h:30-37, plus the trajectory
 * counters the work model needs.  48 bytes. */
typedef struct fbstab_out {
  int32_t eflag;
  int32_t newton_iters;
  int32_t prox_iters;
  int32_t status;
  double residual;
  double initial_residual;
  double solve_time; /* seconds for the whole batch; negative = not timed */
  int32_t ls_backtracks;
  int32_t residual_evals; /* residual evaluations the engine actually ran */
} fbstab_out;

void fbstab_default_options(fbstab_options* o);
void fbstab_reliable_options(fbstab_options* o);
/* Clamps exactly like ValidateOptions; returns FBSTAB_ERR_INVALID where the
 * reference's saturate() would throw. */
int fbstab_validate_options(fbstab_options* o);

/* Thread-local description of the last error on this thread. */
const char* fbstab_last_error(void);
/* Number of CUDA devices visible (0 if none / driver missing). */
int fbstab_device_count(void);

/* Measured FP64 peaks of `device` in TFLOP/s: a register-resident DFMA loop
 * and an FP64 mma.sync (DMMA) loop over all SMs (roofline denominators; the
 * driver's MEASURED_PEAKS.json has no FP64 figure). */
int fbstab_fp64_peak(int device, double* dfma_tflops, double* dmma_tflops);
/* Combined TFLOP/s when half the warps of every SM run the DFMA loop and the
 * other half the DMMA loop at the same time (do the two FP64 paths overlap?). */
int fbstab_fp64_peak_concurrent(int device, double* total_tflops);

/* ---- dense QPs ----------------------------------------------------------- */
typedef struct fbstab_dense_batch fbstab_dense_batch;

int fbstab_dense_batch_create(int nz, int nl, int nv, int max_batch, int device,
                              fbstab_dense_batch** handle);
int fbstab_dense_batch_destroy(fbstab_dense_batch* handle);
int fbstab_dense_batch_set_options(fbstab_dense_batch* handle,
                                   const fbstab_options* o);
int fbstab_dense_batch_get_options(const fbstab_dense_batch* handle,
                                   fbstab_options* o);
/* z,l,v: warm start in, solution out (or the infeasibility certificate, as in
 * fbstab_algorithm-impl.h:209).  y: out only (input ignored, impl:342).
 * stream: a cudaStream_t (NULL = default stream).
 * H must be symmetric, as the QP requires (README.md:2-18 of the reference):
 * the small-problem path (nz <= 32) keeps only its lower triangle on chip. */
int fbstab_dense_batch_solve(fbstab_dense_batch* handle, int batch,
                             const double* H, const double* f, const double* G,
                             const double* h, const double* A, const double* b,
                             double* z, double* l, double* v, double* y,
                             fbstab_out* out, void* stream);
/* Number of kernels the last solve launched on this handle. */
int fbstab_dense_batch_last_launches(const fbstab_dense_batch* handle);
/* Engine path the handle selected: a short static string for logs. */
const char* fbstab_dense_batch_path(const fbstab_dense_batch* handle);

/* ---- MPC-structured QPs -------------------------------------------------- */
typedef struct fbstab_mpc_batch fbstab_mpc_batch;

int fbstab_mpc_batch_create(int N, int nx, int nu, int nc, int max_batch,
                            int device, fbstab_mpc_batch** handle);
int fbstab_mpc_batch_destroy(fbstab_mpc_batch* handle);
int fbstab_mpc_batch_set_options(fbstab_mpc_batch* handle,
                                 const fbstab_options* o);
int fbstab_mpc_batch_get_options(const fbstab_mpc_batch* handle,
                                 fbstab_options* o);
int fbstab_mpc_batch_solve(fbstab_mpc_batch* handle, int batch, const double* Q,
                           const double* R, const double* S, const double* q,
                           const double* r, const double* A, const double* B,
                           const double* c, const double* E, const double* L,
                           const double* d, const double* x0, double* z,
                           double* l, double* v, double* y, fbstab_out* out,
                           void* stream);
int fbstab_mpc_batch_last_launches(const fbstab_mpc_batch* handle);
const char* fbstab_mpc_batch_path(const fbstab_mpc_batch* handle);

/* Shared stage data: ONE copy of the 11 sequences for the whole batch, x0 (and
 * the iterates) per instance -- one plant solved from `batch` initial states,
 * which is the MPC use case and exactly what the reference's wire format cannot
 * say: FBstabMpc::ProblemDataRef (fbstab/fbstab_mpc.h:90-120) points at one
 * MapMatrixSequence per field (tools/matrix_sequence.h:88-164), so a batch in
 * that format repeats the same matrices `batch` times.  Ships 1/batch of the
 * bytes and skips the device-side common-data detection of
 * fbstab_mpc_batch_solve; results are bit-identical to passing the replicated
 * data. */
int fbstab_mpc_batch_solve_shared(fbstab_mpc_batch* handle, int batch,
                                  const double* Q, const double* R, const double* S,
                                  const double* q, const double* r, const double* A,
                                  const double* B, const double* c, const double* E,
                                  const double* L, const double* d, const double* x0,
                                  double* z, double* l, double* v, double* y,
                                  fbstab_out* out, void* stream);
/* Time-invariant (LTI) problem: ONE STAGE of each sequence (Q nx*nx, R nu*nu,
 * S nu*nx, q nx, r nu, A nx*nx, B nx*nu, c nx, E nc*nx, L nc*nu, d nc),
 * replicated over the horizon as OcpGenerator::CopyOverHorizon does
 * (fbstab/test/ocp_generator.cc:397-418): stages 0..N (A, B, c: 0..N-1), with
 * E(0) = 0 -- no constraint on the measured state. */
int fbstab_mpc_batch_solve_lti(fbstab_mpc_batch* handle, int batch, const double* Q,
                               const double* R, const double* S, const double* q,
                               const double* r, const double* A, const double* B,
                               const double* c, const double* E, const double* L,
                               const double* d, const double* x0, double* z,
                               double* l, double* v, double* y, fbstab_out* out,
                               void* stream);

/* ---- sparse QPs with a common sparsity pattern (FBstabSparse) --------------
 * The reference plans "general sparse matrix components" (ROADMAP.md:10,
 * README.md:47) on an LDL' of the quasi-definite Newton matrix behind the
 * interface its tools/qdldl/qdldl_wrapper.h:19-84 sketches:
 *   fbstab_sparse_batch_create  <- QdldlWrapper::QdldlWrapper(n, Ap, Ai): the
 *                                  symbolic analysis (ordering, elimination tree,
 *                                  pattern of L), once per pattern, on the host;
 *   fbstab_sparse_batch_solve   <- FBstabAlgorithm::Solve over sparse data, with
 *                                  QdldlWrapper::Factor / ::Solve per Newton step
 *                                  on the device (one lane per instance).
 * Every instance of a batch has the SAME pattern: H (nz x nz) as its upper
 * triangle in compressed-column form (the storage qdldl_wrapper.h:12-14 names),
 * G (nl x nz) and A (nv x nz) in compressed-column form, row indices strictly
 * increasing within a column, 0-based int32.  Values are instance-major:
 * Hx[batch][nnz(H)], Gx[batch][nnz(G)], Ax[batch][nnz(A)]; f, h, b, z, l, v, y
 * as for the dense entry.  perm (nz+nl+nv entries, perm[new] = old over the
 * variables [z; l; w]) overrides the built-in minimum-degree order; NULL = built-in. */
typedef struct fbstab_sparse_batch fbstab_sparse_batch;

int fbstab_sparse_batch_create(int nz, int nl, int nv, const int* Hp, const int* Hi,
                               const int* Gp, const int* Gi, const int* Ap, const int* Ai,
                               const int* perm, int max_batch, int device,
                               fbstab_sparse_batch** handle);
int fbstab_sparse_batch_destroy(fbstab_sparse_batch* handle);
int fbstab_sparse_batch_set_options(fbstab_sparse_batch* handle, const fbstab_options* o);
int fbstab_sparse_batch_get_options(const fbstab_sparse_batch* handle, fbstab_options* o);
int fbstab_sparse_batch_solve(fbstab_sparse_batch* handle, int batch, const double* Hx,
                              const double* f, const double* Gx, const double* h,
                              const double* Ax, const double* b, double* z, double* l,
                              double* v, double* y, fbstab_out* out, void* stream);
int fbstab_sparse_batch_last_launches(const fbstab_sparse_batch* handle);
const char* fbstab_sparse_batch_path(const fbstab_sparse_batch* handle);
/* Result of the symbolic analysis: size of K, its stored entries, entries of L, and
 * (perm != NULL) the elimination order in use.  Any pointer may be NULL. */
int fbstab_sparse_batch_analysis(const fbstab_sparse_batch* handle, int* n, int* nnzK,
                                 int* nnzL, int* perm);
/* The same analysis without a handle and without a device (host only): what
 * QdldlWrapper's constructor computes from (n, Ap, Ai) before any numeric work
 * (tools/qdldl/qdldl_wrapper.h:33-40).  user_perm as in fbstab_sparse_batch_create
 * (NULL: the built-in minimum-degree order); any output pointer may be NULL. */
int fbstab_sparse_analyze(int nz, int nl, int nv, const int* Hp, const int* Hi,
                          const int* Gp, const int* Gi, const int* Ap, const int* Ai,
                          const int* user_perm, int* n, int* nnzK, int* nnzL, int* perm);
/* The pattern of the factor L of the permuted Newton matrix (strictly lower triangle,
 * compressed columns: Lp n+1 entries, Li nnzL entries) -- what QdldlWrapper keeps in
 * Lp_ / Li_ (tools/qdldl/qdldl_wrapper.h:70-73).  Either pointer may be NULL. */
int fbstab_sparse_batch_factor_pattern(const fbstab_sparse_batch* handle, int* Lp, int* Li);

/* ---- receding-horizon (closed-loop) MPC ----------------------------------
 * What OcpGenerator::GetSimulationInputs exists for (fbstab/test/
 * ocp_generator.h:31-38,69; ocp_generator.cc:56-71) and what the reference's
 * README calls "can be easily warmstarted" (README.md:20): `batch` plants are
 * simulated for T control steps.  One step, for every plant at once and entirely
 * on the device: shift the previous solution by one stage (warm start), solve
 * the OCP from the current state, apply the first input to the plant
 * x+ = Asim x + Bsim u (Asim = NULL: stage 0 of the plant's own OCP,
 * x+ = A(0) x + B(0) u + c(0)), log x and u.  The handle owns the data, the
 * iterates and the logs; a step only enqueues kernels on `stream`.
 * shared_data != 0: the 11 sequences are one copy for all plants.  A plant
 * whose solve ends infeasible / failed holds its previous input and restarts
 * cold.  run(): T steps from x_init, then X (batch x (T+1) x nx), U (batch x T
 * x nu) and out (T x batch) are copied to the caller (host or device, any may
 * be NULL). */
typedef struct fbstab_mpc_closed_loop fbstab_mpc_closed_loop;
int fbstab_mpc_closed_loop_create(int N, int nx, int nu, int nc, int batch, int device,
                                  int shared_data, const double* Q, const double* R,
                                  const double* S, const double* q, const double* r,
                                  const double* A, const double* B, const double* c,
                                  const double* E, const double* L, const double* d,
                                  const double* x_init, const double* Asim,
                                  const double* Bsim, int max_steps,
                                  fbstab_mpc_closed_loop** handle);
int fbstab_mpc_closed_loop_destroy(fbstab_mpc_closed_loop* handle);
int fbstab_mpc_closed_loop_set_options(fbstab_mpc_closed_loop* handle,
                                       const fbstab_options* o);
int fbstab_mpc_closed_loop_reset(fbstab_mpc_closed_loop* handle, void* stream);
int fbstab_mpc_closed_loop_step(fbstab_mpc_closed_loop* handle, int warm_start,
                                void* stream);
int fbstab_mpc_closed_loop_run(fbstab_mpc_closed_loop* handle, int steps,
                               int warm_start, double* X, double* U, fbstab_out* out,
                               void* stream);
const char* fbstab_mpc_closed_loop_path(const fbstab_mpc_closed_loop* handle);

/* ---- component stages (for per-kernel parity tests) ----------------------
 * One CTA per instance runs ONE stage of the engine on caller-supplied
 * iterates.  All pointers host or device as above; unused ones may be NULL.
 *
 *  FBSTAB_COMP_MARGIN     y = b - A z
 *                         (FullVariable::InitializeConstraintMargin,
 *                          fbstab/components/full_variable.cc:47-53)
 *  FBSTAB_COMP_RESIDUAL   inner residual R(x,xbar,sigma) -> (rz,rl,rv) and
 *                         norms[0..2]; penalised natural residual norms ->
 *                         norms[3..5] (FullResidual::InnerResidual /
 *                         PenalizedNaturalResidual, full_residual.cc:49-109).
 *                         y of x is an INPUT here.
 *  FBSTAB_COMP_NEWTON     LinearSolver::Initialize(x,xbar,sigma) then
 *                         ::Solve(r,&dx) with r = (rz,rl,rv) -> (dz,dl,dv,dy),
 *                         gamma, mus; status[i] = factor status
 *                         (dense_cholesky_solver.cc:32-127 /
 *                          riccati_linear_solver.cc:77-344)
 *  FBSTAB_COMP_FEAS       FullFeasibility::CheckFeasibility on (dz,dl,dv)
 *                         passed in (z,l,v); status[i] = 0 feasible,
 *                         1 primal infeasible, 2 dual infeasible, 3 both
 *                         (full_feasibility.cc:25-88)
 */
#define FBSTAB_COMP_MARGIN 0
#define FBSTAB_COMP_RESIDUAL 1
#define FBSTAB_COMP_NEWTON 2
#define FBSTAB_COMP_FEAS 3

typedef struct fbstab_component_io {
  /* iterate x and proximal centre xbar */
  const double *z, *l, *v, *y;
  const double *zbar, *lbar, *vbar;
  /* residual in (NEWTON) / out (RESIDUAL) */
  double *rz, *rl, *rv;
  /* step out (NEWTON), margin out (MARGIN uses dy) */
  double *dz, *dl, *dv, *dy;
  double *gamma, *mus; /* nv each, out (NEWTON) */
  double* norms;       /* 8 per instance, out (RESIDUAL): inner |rz|,|rl|,|rv|,
                          natural |rz|,|rl|,|rv|, then Ei and Eo */
  int32_t* status;     /* 1 per instance, out */
  double sigma;
  double tol; /* FEAS */
} fbstab_component_io;

int fbstab_dense_batch_component(fbstab_dense_batch* handle, int comp,
                                 int batch, const double* H, const double* f,
                                 const double* G, const double* h,
                                 const double* A, const double* b,
                                 const fbstab_component_io* io, void* stream);
int fbstab_mpc_batch_component(fbstab_mpc_batch* handle, int comp, int batch,
                               const double* Q, const double* R,
                               const double* S, const double* q,
                               const double* r, const double* A,
                               const double* B, const double* c,
                               const double* E, const double* L,
                               const double* d, const double* x0,
                               const fbstab_component_io* io, void* stream);

/* ---- multi-GPU: the batch shards by instance -----------------------------
 * QP instances are independent (one reference Solve call each,
 * fbstab/fbstab_dense.h:137-142), so a batch is cut into contiguous instance
 * ranges, one per GPU, and the solve itself needs no collective.
 *
 * (1) One process per GPU (torchrun, MPI): rank 0 calls
 * fbstab_multi_gpu_unique_id, the launcher's own channel broadcasts the 128
 * bytes, every rank calls fbstab_multi_gpu_create and solves its shard
 * [fbstab_multi_gpu_shard] with its own fbstab_*_batch handle on DEVICE result
 * buffers.  fbstab_multi_gpu_gather is the only exchange: one grouped NCCL
 * send/recv per result array moves each rank's rows straight from its result
 * buffers into the root's global arrays (z: nz, l: nl, v and y: nv doubles and
 * one fbstab_out per instance; Z..OUT are read on the root only), on `stream`.
 * NCCL is resolved at run time (libnccl.so.2).                                */
typedef struct fbstab_multi_gpu fbstab_multi_gpu;
int fbstab_multi_gpu_unique_id(char id[128]);
int fbstab_multi_gpu_create(int rank, int nranks, const char id[128], int device,
                            fbstab_multi_gpu** handle);
int fbstab_multi_gpu_destroy(fbstab_multi_gpu* handle);
/* Contiguous range [first, first + count) of `global_batch` instances owned by
 * `rank`: the first (global_batch % nranks) ranks take one instance more. */
int fbstab_multi_gpu_shard(int nranks, int rank, long global_batch, long* first,
                           long* count);
int fbstab_multi_gpu_gather(fbstab_multi_gpu* handle, int root, long global_batch,
                            int nz, int nl, int nv, const double* z,
                            const double* l, const double* v, const double* y,
                            const fbstab_out* out, double* Z, double* L, double* V,
                            double* Y, fbstab_out* OUT, void* stream);

/* (2) One process driving several GPUs with HOST buffers (what the C++ facade's
 * SolveBatch(..., devices) calls): one batch handle and one host thread per
 * device, each running the pipelined host-buffer solve on its shard; results
 * land in the caller's arrays at the shard's offset.  Same argument meaning as
 * fbstab_dense_batch_solve / fbstab_mpc_batch_solve. */
typedef struct fbstab_dense_multi_gpu fbstab_dense_multi_gpu;
typedef struct fbstab_mpc_multi_gpu fbstab_mpc_multi_gpu;
int fbstab_dense_multi_gpu_create(int ndev, const int* devices, int nz, int nl,
                                  int nv, long max_batch,
                                  fbstab_dense_multi_gpu** handle);
int fbstab_dense_multi_gpu_destroy(fbstab_dense_multi_gpu* handle);
int fbstab_dense_multi_gpu_set_options(fbstab_dense_multi_gpu* handle,
                                       const fbstab_options* o);
int fbstab_dense_multi_gpu_solve(fbstab_dense_multi_gpu* handle, long batch,
                                 const double* H, const double* f, const double* G,
                                 const double* h, const double* A, const double* b,
                                 double* z, double* l, double* v, double* y,
                                 fbstab_out* out);
int fbstab_mpc_multi_gpu_create(int ndev, const int* devices, int N, int nx, int nu,
                                int nc, long max_batch,
                                fbstab_mpc_multi_gpu** handle);
int fbstab_mpc_multi_gpu_destroy(fbstab_mpc_multi_gpu* handle);
int fbstab_mpc_multi_gpu_set_options(fbstab_mpc_multi_gpu* handle,
                                     const fbstab_options* o);
int fbstab_mpc_multi_gpu_solve(fbstab_mpc_multi_gpu* handle, long batch,
                               const double* Q, const double* R, const double* S,
                               const double* q, const double* r, const double* A,
                               const double* B, const double* c, const double* E,
                               const double* L, const double* d, const double* x0,
                               double* z, double* l, double* v, double* y,
                               fbstab_out* out);

/* The same for sparse QPs with a common pattern (fbstab_sparse_batch_*): every device runs
 * the same symbolic analysis (same elimination order), values and results are
 * instance-major HOST arrays sharded by contiguous index range. */
typedef struct fbstab_sparse_multi_gpu fbstab_sparse_multi_gpu;
int fbstab_sparse_multi_gpu_create(int ndev, const int* devices, int nz, int nl, int nv,
                                   const int* Hp, const int* Hi, const int* Gp,
                                   const int* Gi, const int* Ap, const int* Ai,
                                   const int* perm, long max_batch,
                                   fbstab_sparse_multi_gpu** handle);
int fbstab_sparse_multi_gpu_destroy(fbstab_sparse_multi_gpu* handle);
int fbstab_sparse_multi_gpu_set_options(fbstab_sparse_multi_gpu* handle,
                                        const fbstab_options* o);
int fbstab_sparse_multi_gpu_solve(fbstab_sparse_multi_gpu* handle, long batch,
                                  const double* Hx, const double* f, const double* Gx,
                                  const double* h, const double* Ax, const double* b,
                                  double* z, double* l, double* v, double* y,
                                  fbstab_out* out);

/* ---- synthetic problems (host code; mirrors fbstab/test/ocp_generator.h) - */
#define FBSTAB_OCP_DOUBLE_INTEGRATOR 0 /* ocp_generator.cc:319-363 nx2 nu1 nc6  */
#define FBSTAB_OCP_SERVO_MOTOR 1       /* ocp_generator.cc:245-315 nx4 nu1 nc4  */
#define FBSTAB_OCP_SPACECRAFT 2        /* ocp_generator.cc:171-244 nx6 nu3 nc12 */
#define FBSTAB_OCP_COPOLYMERIZATION 3  /* ocp_generator.cc:73-169 nx18 nu5 nc10 */

int fbstab_ocp_dims(int kind, int* nx, int* nu, int* nc);
/* One instance in the wire format (time-varying, E(0)=0; ocp_generator.cc:365-421). */
int fbstab_ocp_generate(int kind, int N, double* Q, double* R, double* S,
                        double* q, double* r, double* A, double* B, double* c,
                        double* E, double* L, double* d, double* x0);
/* OcpGenerator::GetSimulationInputs (ocp_generator.cc:56-71): the plant
 * x+ = A x + B u, y = C x (D = 0 in all four fixtures), initial state and number
 * of steps T.  A (nx*nx), B (nx*nu), C (ny*nx), x0 (nx): column-major, written
 * where non-NULL; ny <= nx. */
int fbstab_ocp_simulation(int kind, double* A, double* B, double* C, double* x0,
                          int* ny, int* T);
/* `count` instances first..first+count-1 of benchmark config `config`:
 * identical stage data, x0 = nominal + rho*U(-1,1)^nx (instance 0 unperturbed);
 * rho < 0 perturbs one-sidedly, x0 = nominal + |rho|*U(0,1)^nx. */
int fbstab_ocp_generate_batch(int kind, int N, int config, long first, int count,
                              double rho, double* Q, double* R, double* S,
                              double* q, double* r, double* A, double* B,
                              double* c, double* E, double* L, double* d,
                              double* x0);
/* Seeded random strictly convex, strictly feasible dense QPs (kind 0), or with
 * a planted infeasibility (kind 1) / unbounded ray (kind 2). */
int fbstab_random_dense_qp(int config, long first, int count, int nz, int nl,
                           int nv, int kind, double* H, double* f, double* G,
                           double* h, double* A, double* b, int nthreads);

#ifdef __cplusplus
}
#endif
#endif /* FBSTAB_B200_H_ */

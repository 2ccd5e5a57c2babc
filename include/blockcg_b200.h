/* blockcg_b200 -- C-ABI of the B200-native block-CG hot path.
 *
 * This is the ONLY surface through which host code reaches the CUDA kernels:
 * plain pointers, sizes and opaque handles, no C++/torch types.  It replaces,
 * for the hot path only, what the reference does inside its header templates
 * (citations relative to the lkeegan/blockCG tree):
 *
 *   bcg_solve_bcg     <- BCG<N>      inc/block_solvers.hpp:10-45
 *   bcg_solve_bcgrq   <- BCGrQ<N>    inc/block_solvers.hpp:50-86
 *   bcg_solve_sbcgrq  <- SBCGrQ<N>   inc/block_solvers.hpp:91-185
 *   bcg_op            <- dirac_op::op<N>                 inc/dirac_op.hpp:36-43
 *   bcg_set_links     <- dirac_op ctor / private U       inc/dirac_op.hpp:9-11,24-32
 *   bcg_gram          <- block_fermion_field::hermitian_dot   inc/fields.hpp:103-122
 *   bcg_add           <- block_fermion_field::add             inc/fields.hpp:70-77
 *   bcg_rescale_add   <- block_fermion_field::rescale_add     inc/fields.hpp:79-90
 *   bcg_trsm          <- multiply_upper_triangular_inverse_RHS inc/fields.hpp:125-136
 *   bcg_thinqr        <- block_fermion_field::thinQR          inc/fields.hpp:140-146
 *
 * Memory layouts are the reference's own, as interleaved (re,im) doubles:
 *   field  : [V][N][3] complex128 (site-major, 3xN column-major per site)
 *   links  : [V][3][3] complex128 (column-major 3x3 per site)
 *   matrix : N x N complex128, column-major
 *
 * Every function returns 0 on success or a bcg_status code; no exception ever
 * crosses this boundary.  bcg_last_error() gives a human-readable message.
 * A context is bound to one CUDA device and one host thread at a time.
 * There is no CPU fallback: without a CUDA device bcg_ctx_create fails.
 */
#ifndef BLOCKCG_B200_H
#define BLOCKCG_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bcg_ctx bcg_ctx;

typedef enum {
  BCG_OK = 0,
  BCG_ERR_INVALID = 1,       /* bad argument (null pointer, unsupported N, ...) */
  BCG_ERR_CUDA = 2,          /* a CUDA runtime call failed                       */
  BCG_ERR_NOT_PD = 3,        /* Gram matrix not positive definite (LLT pivot <= 0;
                                the reference leaves this unchecked, LLT.h:449)  */
  BCG_ERR_NCCL = 4,          /* a NCCL call failed                               */
  BCG_ERR_NO_COMM = 5,       /* multi-rank context used before bcg_comm_init     */
  BCG_ERR_NAN = 6            /* residual became NaN (the reference exits its loop
                                silently in that case)                           */
} bcg_status;

#define BCG_MAX_SHIFTS 32
#define BCG_UNIQUE_ID_BYTES 128
#define BCG_IPC_HANDLE_BYTES 64

/* ---- library / build info -------------------------------------------------------- */
const char* bcg_version(void);
/* 1 if kernels for this N_rhs are compiled in (the reference fixes N at compile
 * time, inc/fields.hpp:19-25; here it is a run-time argument). */
int bcg_supports_nrhs(int n_rhs);

/* ---- context ---------------------------------------------------------------------- */
/* v_local   : number of sites owned by this rank (contiguous slab of the global
 *             site index; for one rank = the reference's V).
 * n_rhs     : N_rhs.     max_shifts : largest shift count SBCGrQ will be asked for.
 * device    : CUDA device ordinal.
 * rank/nranks : position in the slab decomposition (0/1 for a single GPU). */
int bcg_ctx_create(bcg_ctx** ctx, int64_t v_local, int n_rhs, int max_shifts, int device, int rank,
                   int nranks);
/* 4-D extension (NOT in the reference, whose operator is a 1-D chain, inc/dirac_op.hpp:13-21):
 * local lattice L0 x L1 x L2 x L3, x = x0 + L0 (x1 + L1 (x2 + L2 x3)), slab-decomposed along
 * x3 (dims_local[3] = this rank's thickness), four links per site.  Every other entry point
 * works unchanged on such a context; v_local = L0*L1*L2*L3. */
int bcg_ctx_create_4d(bcg_ctx** ctx, const int64_t* dims_local /* [4] */, int n_rhs, int max_shifts, int device,
                      int rank, int nranks);
int bcg_ctx_destroy(bcg_ctx* ctx);
const char* bcg_last_error(const bcg_ctx* ctx);

/* multi-GPU plumbing: rank 0 obtains an id, the host distributes it (e.g. with
 * torch.distributed / MPI), every rank calls bcg_comm_init. */
int bcg_comm_get_unique_id(void* id_out /* BCG_UNIQUE_ID_BYTES */);
int bcg_comm_init(bcg_ctx* ctx, const void* id /* BCG_UNIQUE_ID_BYTES */);
/* Peer-memory exchange over NVLink (optional, one process per GPU on one node): every rank
 * obtains the CUDA-IPC handle of its communication buffer, the host gathers the handles of
 * all ranks (rank order) and hands the array to every rank, then synchronises the ranks
 * once.  With it the iteration loop contains no NCCL call: the Gram kernels store their
 * N x N block into every peer's buffer themselves, the halo sites travel as P2P stores. */
int bcg_comm_ipc_handle(bcg_ctx* ctx, void* handle_out /* BCG_IPC_HANDLE_BYTES */);
int bcg_comm_ipc_open(bcg_ctx* ctx, const void* handles /* nranks * BCG_IPC_HANDLE_BYTES */);
/* Back to the NCCL exchange (e.g. when some rank could not map its peers). */
int bcg_comm_ipc_disable(bcg_ctx* ctx);

/* ---- operator --------------------------------------------------------------------- */
/* links_host: this rank's [v_local][3][3] links; halos (2 sites each side) are
 * filled by periodic wrap (1 rank) or neighbour exchange (nranks > 1). */
int bcg_set_links(bcg_ctx* ctx, const double* links_host, double mass);
/* 4-D context: links_host = [v_local][4][3][3] (direction-major per site, column-major 3x3);
 * D v[x] = 1/2 sum_mu ( U_mu[x] v[x+mu] - U_mu[x-mu]^dag v[x-mu] ), operator m^2 - D^2. */
int bcg_set_links_4d(bcg_ctx* ctx, const double* links_host, double mass);

/* Links drawn in place on the device, uniform in [-1, 1) per real/imaginary part -- the
 * distribution of the reference's dirac_op constructor (inc/dirac_op.hpp:24-32, Eigen setRandom),
 * from a counter-based generator instead of libc rand(): double number k of the GLOBAL
 * [V][3][3] (or [V][4][3][3]) array is a fixed function of (seed, k), whatever the number of ranks.
 * For volumes too large to stage through the host. */
int bcg_set_links_random(bcg_ctx* ctx, uint64_t seed, double mass);

/* ---- device-resident fields --------------------------------------------------------- */
int bcg_field_alloc(bcg_ctx* ctx, int* handle_out);
int bcg_field_free(bcg_ctx* ctx, int handle);
int bcg_field_upload(bcg_ctx* ctx, int handle, const double* host);   /* [v_local][N][3] */
int bcg_field_download(bcg_ctx* ctx, int handle, double* host);
/* The same generator for a field (reference: block_fermion_field setRandom in benchmark.cpp:60-62);
 * a different stream from the links, so equal seeds do not correlate them. */
int bcg_field_random(bcg_ctx* ctx, int handle, uint64_t seed);
int bcg_field_zero(bcg_ctx* ctx, int handle);
int bcg_field_copy(bcg_ctx* ctx, int dst, int src);

/* ---- primitives (unit tests, micro-benchmarks, verification) -------------------------- */
/* out = (m^2 - D^2) in + sigma * in      (sigma = 0 for the bare operator)
 * gram_host (optional, may be NULL): receives in^dag out, the Gram fused into
 * the stencil epilogue. */
int bcg_op(bcg_ctx* ctx, int out, int in, double sigma, double* gram_host);
/* R = a^dag b: lower triangle + diagonal accumulated, upper = conj mirror. */
int bcg_gram(bcg_ctx* ctx, int a, int b, double* r_host);
/* dst += src * M   (M: N x N, host) */
int bcg_add(bcg_ctx* ctx, int dst, int src, const double* m_host);
/* dst += src * s   (real scalar; block_solvers.hpp:136) */
int bcg_add_scalar(bcg_ctx* ctx, int dst, int src, double s);
/* dst = dst * L + src * r */
int bcg_rescale_add(bcg_ctx* ctx, int dst, const double* l_host, int src, double r);
/* q <- q R^-1, R upper triangular (back substitution, column order of the reference) */
int bcg_trsm(bcg_ctx* ctx, int q, const double* r_host);
/* CholQR: R = chol(q^dag q)^dag, q <- q R^-1; r_host receives R (zero below diag). */
int bcg_thinqr(bcg_ctx* ctx, int q, double* r_host);
/* sqrt(diag((A+sigma)x - b)^dag(...) / diag(b^dag b)) per rhs: the reference's only
 * acceptance criterion (benchmark.cpp:93-103, test/solvers.cpp:99-118). */
int bcg_true_residual(bcg_ctx* ctx, int x, int b, double sigma, double* res_host /* [N] */);

/* ---- solvers on device-resident fields (benchmark path: no PCIe traffic in the loop) --- */
typedef struct {
  int iterations;          /* number of operator applications (the reference's return value) */
  double residual;         /* solver's own residual estimate at exit                       */
  int n_unconverged;       /* SBCGrQ: shifts still being updated at exit                    */
  double solve_ms;         /* device time of the iteration loop (CUDA events)               */
  double setup_ms;         /* device time of the setup (thinQR of B, copies)                */
  int64_t kernel_launches; /* kernels launched by this solve                                */
} bcg_solve_info;

int bcg_solve_bcg_dev(bcg_ctx* ctx, int x, int b, double eps, int max_iterations, bcg_solve_info* info);
int bcg_solve_bcgrq_dev(bcg_ctx* ctx, int x, int b, double eps, int max_iterations, bcg_solve_info* info);
/* x_handles[n_shifts]; sigma ascending, sigma[0] >= 0 (block_solvers.hpp:97-101) */
int bcg_solve_sbcgrq_dev(bcg_ctx* ctx, const int* x_handles, int b, const double* sigma, int n_shifts,
                         double eps, double eps_shifts, int max_iterations, bcg_solve_info* info);

/* ---- solvers on host buffers (drop-in for the reference's templates) -------------------- */
int bcg_solve_bcg(bcg_ctx* ctx, double* x_host, const double* b_host, double eps, int max_iterations,
                  bcg_solve_info* info);
int bcg_solve_bcgrq(bcg_ctx* ctx, double* x_host, const double* b_host, double eps, int max_iterations,
                    bcg_solve_info* info);
/* x_host: n_shifts pointers, each to a [v_local][N][3] buffer */
int bcg_solve_sbcgrq(bcg_ctx* ctx, double* const* x_host, const double* b_host, const double* sigma,
                     int n_shifts, double eps, double eps_shifts, int max_iterations, bcg_solve_info* info);

/* ---- CG / SCG: the reference's solvers for ONE right-hand side (context with n_rhs = 1) ------
 *   bcg_solve_cg   <- CG    src/standard_solvers.cpp:3-32    (inc/standard_solvers.hpp:10-13)
 *   bcg_solve_scg  <- SCG   src/standard_solvers.cpp:34-95   (inc/standard_solvers.hpp:15-20)
 * Scalar recurrences (alpha, beta, zeta_s, theta_s) evaluated on the device exactly as the reference
 * writes them; stopping rule |r| > eps |b| on the lowest shift, a shifted system is dropped once
 * |r| zeta_s < eps_shifts (NOT normalised by |b|, as in the reference, :89-92). */
int bcg_solve_cg_dev(bcg_ctx* ctx, int x, int b, double eps, int max_iterations, bcg_solve_info* info);
int bcg_solve_scg_dev(bcg_ctx* ctx, const int* x_handles, int b, const double* sigma, int n_shifts, double eps,
                      double eps_shifts, int max_iterations, bcg_solve_info* info);
int bcg_solve_cg(bcg_ctx* ctx, double* x_host, const double* b_host, double eps, int max_iterations,
                 bcg_solve_info* info);
int bcg_solve_scg(bcg_ctx* ctx, double* const* x_host, const double* b_host, const double* sigma, int n_shifts,
                  double eps, double eps_shifts, int max_iterations, bcg_solve_info* info);

/* ---- statistics of the last solve on this context (measurement; nothing the reference has) ---- */
typedef struct {
  int iterations;
  int n_shifts;
  int paired;                                    /* schedule of the multishift update (bcg_shift_schedule /
                                                    bcg_stag_schedule): 0 every system every iteration, 1 / 2 the
                                                    shifted systems every second iteration, 3 every depth-th */
  int depth;                                     /* deferral depth of that schedule (1 for schedule 0) */
  uint32_t active_hist[BCG_MAX_SHIFTS + 1];      /* [a] = iterations in which a systems were still updated
                                                    (block_solvers.hpp:161,179-181: shifts retire at eps_shifts) */
  uint64_t shift_update_field_passes;            /* field-sized (48 N V bytes) reads + writes the multishift
                                                    update kernels moved through HBM over the whole solve */
  double resid_shift[BCG_MAX_SHIFTS];            /* last residual estimate per system ([0]: the stopping residual) */
} bcg_solve_stats;
int bcg_last_solve_stats(bcg_ctx* ctx, bcg_solve_stats* out);

/* The schedule of the multishift update, as the kernels evaluate it on the device from the loop's control
 * block (host-side copy of the same function; needs no device): schedule 0 = every active system in every
 * iteration (the reference's order, block_solvers.hpp:161-182), 1 = shifted systems served every second
 * iteration with both pending updates, 2 = the same, odd-numbered systems in odd and even-numbered systems in
 * even iterations.  A system's updates are always applied in iteration order with the coefficients and the Q
 * of their own iteration, so all three produce identical bits.  kinds[i]: 0 Q <- Q rho^-1, 1 the same and kept
 * as the previous Q, 2 previous Q read, 3 / 4 / 5: system systems[i] gets this / the previous / both
 * iterations' updates.  Returns the number of items (<= BCG_MAX_SHIFTS + 2). */
int bcg_shift_schedule(int schedule, int iteration, int stop, int n_active, int n_active_prev, int* kinds, int* systems,
                       int* field_passes);

/* Schedule 3 (BCG_PAIR=3, deferral depth BCG_DEPTH = 2..4): system s >= 1 is served in the iterations i with
 * i % depth == s % depth and then receives its (up to) `depth` pending updates in iteration order, each with the
 * coefficients and the Q of its own iteration (Q lives in a ring of `ring` >= depth fields), so the bits are again
 * those of the plain loop (block_solvers.hpp:161-182).  n_active_ring[j % ring] = systems active in iteration j, for
 * the depth-1 iterations before `iteration`.  part = 0: the whole launch; with BCG_OVERLAP=1 the launch is split into
 * part 1 (Q <- Q rho^-1 and system 0, on the loop's stream) and part 2 (the shifted systems, ring = depth + 1, on a
 * second stream beside the next iterations' kernels).  Per item i: systems[i] (-1 = the Q item), first_back[i] (its
 * first pending update is that of iteration `iteration - first_back[i]`), n_updates[i].  Returns the number of items
 * (-1: bad argument). */
int bcg_stag_schedule(int depth, int ring, int part, int iteration, int stop, int n_active, const int* n_active_ring,
                      int* systems, int* first_back, int* n_updates, int* field_passes);

/* In-loop profile: n_iterations (<= 4096) of the NEXT solve on this context, starting once at least
 * after_iterations have run (so the GPU is at its sustained clocks), are submitted kernel by kernel with a
 * CUDA event after each stage instead of as graph batches; the solve is otherwise unchanged.  ms[] = mean
 * device milliseconds per iteration over the window: [0] stencil + fused Gram, [1] coefficient A-step,
 * [2] Q update + fused Gram, [3] coefficient B-step, [4] multishift update of an ODD iteration, [5] of an
 * EVEN iteration (paired update: shift_pair.cuh), [6] halo refresh, [7] whole iteration. */
typedef struct {
  int iterations;       /* iterations in the window (0: no profile taken) */
  int first_iteration;  /* iterations completed before the window          */
  int active_systems;   /* systems still being updated at the end of it    */
  double ms[8];
} bcg_loop_profile;
int bcg_set_loop_profile(bcg_ctx* ctx, int n_iterations, int after_iterations);
int bcg_get_loop_profile(bcg_ctx* ctx, bcg_loop_profile* out);

/* ---- unit-test entry points of the device N x N routines (isolated parity with Eigen) ----------
 * bcg_small_inverse : out = a^-1 with the Gauss-Jordan inverse the (S)BCGrQ loops use in place of
 *                     fullPivLu().solve(I) (block_solvers.hpp:142,166); pivot = 1: row pivoting (beta_s),
 *                     0: pivot-free (Hermitian positive definite P^dag T).  *info_out = -1, or the column
 *                     whose pivot vanished (no rank truncation: documented divergence from FullPivLU.h:317-341).
 * bcg_small_lu_solve: x = a^-1 b with the Eigen-faithful full-pivoting LU of the BCG loop
 *                     (block_solvers.hpp:31,36; FullPivLU.h:487-590,745-790, rank threshold included).
 * Matrices N x N complex128 column-major on the host. */
int bcg_small_inverse(bcg_ctx* ctx, const double* a_host, double* out_host, int pivot, int* info_out);
int bcg_small_lu_solve(bcg_ctx* ctx, const double* a_host, const double* b_host, double* x_host);

/* ---- micro-benchmark hooks (timed on the context's stream with CUDA events) -------------- */
/* Runs `reps` back-to-back launches of one kernel and returns the mean device
 * time per launch in *ms_out.  which: 0 dirac apply (+fused Gram), 1 dirac apply only,
 * 2 Gram, 3 Q -= T*alpha with fused Gram, 4 multishift update over n_shifts shifts (7: its
 * first-generation register-direct variant),
 * 5 X += P*M, 6 P = P*L + Q, 9 / 10: the first-generation stencil with / without the fused
 * Gram, 11 / 12: the first-generation Q += T*M with / without it (kept for comparison).  Fields are the context's own scratch, filled by the caller
 * through handles f0..f3 where needed. */
int bcg_bench_kernel(bcg_ctx* ctx, int which, int reps, int n_shifts, const int* handles, int n_handles,
                     double* ms_out, int64_t* launches_out);

#ifdef __cplusplus
}
#endif
#endif /* BLOCKCG_B200_H */

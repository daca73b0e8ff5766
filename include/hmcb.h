/* hmcb.h -- C ABI of the B200-native batched HMC engine (libhmcb.so).
 *
 * This is the drop-in boundary for the hot path of hmclab's HMC sampler.  The
 * reference has no FFI seam on this path: the seam is its Python object protocol
 * (SURVEY.md section 8b).  Each entry point below names the reference interface it
 * replaces (paths relative to the hmclab repository).  The only FFI precedent in the
 * reference is hmclab/Helpers/InterfaceMKL.py:26-33,87-121 (ctypes, raw pointers,
 * int32 CSR), whose style this header follows: plain pointers and sizes, no C++ or
 * torch types, every function returns an int status (0 = ok, <0 = error, message via
 * hmcb_last_error()), no exceptions cross the boundary, no ownership transfer.
 *
 * Memory:
 *   - "HOST" pointers are read during the call and copied; the engine owns only its
 *     private device copies of model constants (freed by hmcb_destroy).
 *   - "DEVICE" pointers are CUDA device pointers owned by the caller (PyTorch tensors:
 *     tensor.data_ptr()); calls enqueue work on `stream` (a cudaStream_t passed as
 *     void*, NULL = legacy default stream) and return without synchronising.
 *   - One engine per device and per host thread (one process per GPU); not thread-safe.
 *
 * Layout: every batch is [chains x dims] row-major (chain-major, dims contiguous),
 * IEEE float64.  Engines with a coupled dense/CSR likelihood keep a private
 * transposed [dims x chains] working copy; that is invisible here.
 */
#ifndef HMCB_H
#define HMCB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HMCB_ABI_VERSION 2

typedef struct hmcb_engine hmcb_engine;

/* Samplers.py:1728-1738 integrator registry {"lf","3s","4s"} */
enum { HMCB_INTEGRATOR_LF = 0, HMCB_INTEGRATOR_3S = 1, HMCB_INTEGRATOR_4S = 2 };
/* elementwise priors: base.py:539-574 (Normal, diagonal), base.py:689-710 (Laplace) */
enum { HMCB_PRIOR_NORMAL = 0, HMCB_PRIOR_LAPLACE = 1 };
/* execution path chosen by hmcb_finalize (reported by hmcb_path) */
enum { HMCB_PATH_FUSED_PRIORS = 0, HMCB_PATH_FUSED_SRCLOC = 1, HMCB_PATH_STAGED = 2,
       HMCB_PATH_FUSED_DENSE = 3 /* staged workspaces + whole-proposal kernel for dims <= 128 */ };

int hmcb_abi_version(void);

/* Host-side self check (no GPU): builds the strip tables of the shared-memory staged CSR SpMM for
 * the given thread mapping (warps x rw rows, cpl chains per lane) and strip limits exactly as
 * hmcb_finalize does, verifies their format invariants and evaluates Y = A B through them the way
 * the kernel walks them.  B [cols x chains], Y [rows x chains] row-major; info[0..3] = strips per
 * chunk, number of (chunk, strip) groups, 1 if the compact 8-byte nonzeros are used, table bytes. */
int hmcb_debug_spmm_tables(int64_t rows, int64_t cols, int64_t nnz, const int32_t *indptr,
                           const int32_t *indices, const double *data, int warps, int rw, int cpl,
                           int kb, int emax, int allow_compact, int64_t chains, const double *B,
                           double *Y, int64_t *info);
/* Same for the row-blocked tensor-core tables (rows regrouped into groups of 8 rows with similar
 * column sets = one DMMA M-tile, `groups_per_warp` groups per consumer warp, k-tiles of 4 columns with
 * the nonzeros of their 8 x 4 A fragments, slabs of 16 `chain_boxes` chains):
 * clusters, builds, checks and walks them like csr_spmm_block_kernel.  `cap16` = size of a pipeline
 * stage in 16-byte units (B rows of a strip + k-tiles of one group).  info[0..7] = strips per chunk,
 * (chunk, strip) groups, column blocks (distinct columns per row group), table bytes, nonzeros after
 * summing duplicates, k-tiles, k-tiles that carry 4 columns, 1 if the matrix values are stored in
 * fp32. */
int hmcb_debug_spmm_block_tables(int64_t rows, int64_t cols, int64_t nnz, const int32_t *indptr,
                                 const int32_t *indices, const double *data, int warps,
                                 int groups_per_warp, int chain_boxes, int cap16, int allow_compact,
                                 int64_t chains, const double *B, double *Y, int64_t *info);
const char *hmcb_last_error(void);

/* lifetime ------------------------------------------------------------------------- */
int hmcb_create(int device, int64_t chains, int64_t dims, hmcb_engine **out);
int hmcb_destroy(hmcb_engine *e);

/* tuning: HMC.sample(amount_of_steps=, integrator=) -- Samplers.py:1351-1358,1379-1384 */
int hmcb_set_integrator(hmcb_engine *e, int integrator, int amount_of_steps);

/* Arithmetic of the priors-only whole-proposal kernel (Samplers.py:1539-1584: `p -= eps * g`,
 * `q += eps * dK/dp`).  on = 0 (default): each of the two updates is one fused multiply-add -- results
 * agree with numpy's to the last few bits (far inside the 1e-10 parity bar; accept / reject decisions
 * are unchanged).  on = 1: multiply and add are rounded separately, in the reference's operation
 * order, and a trajectory on a separable target is bit-identical to numpy's.  Every other kernel
 * always works in the reference's operation order.  May be changed at any time.  The environment
 * variable HMCB_EXACT=1 makes exact the default. */
int hmcb_set_exact_arithmetic(hmcb_engine *e, int on);

/* MassMatrices.Unit (MassMatrices.py:82-156) */
int hmcb_set_mass_unit(hmcb_engine *e);
/* MassMatrices.Diagonal (MassMatrices.py:159-238); HOST arrays of `dims` doubles.
 * inverse_diagonal is passed (not recomputed) because the reference multiplies by its
 * precomputed rounded reciprocal (:180,218). */
int hmcb_set_mass_diagonal(hmcb_engine *e, const double *diagonal,
                           const double *inverse_diagonal);

/* MassMatrices.Full (MassMatrices.py:241-327): momentum = cholesky @ normal, kinetic energy
 * 0.5 p . M^-1 p, dK/dp = M^-1 p (the reference calls scipy's cho_solve; here M^-1 is applied as a
 * dense product over the chain batch on the fp64 tensor cores).  HOST arrays [dims x dims] row-major:
 * the lower Cholesky factor of the mass matrix and its inverse.  Runs on the staged path (also for
 * priors-only targets); not available together with a SourceLocation likelihood. */
int hmcb_set_mass_full(hmcb_engine *e, const double *cholesky_lower, const double *inverse);

/* Exact int8 slice products on the tcgen05 tensor cores (building block of the Ozaki-sliced dense
 * products, csrc/ozaki.cuh): A [SA][M x K], B [SB][N x K] int8 DEVICE arrays (K contiguous; M, K, N
 * multiples of 128), C [orders][M x N] int32 DEVICE: C[o] = sum over s + t = o of A_s B_t^T. */
int hmcb_debug_i8_gemm(int device, int64_t M, int64_t N, int64_t K, int SA, int SB, int orders,
                       const signed char *A, const signed char *B, int32_t *C, void *stream);
/* The modular variant of the sliced products ("Ozaki II", csrc/ozaki.cuh): Y = A X for a HOST matrix A [M x K]
 * (row-major) and a DEVICE chain batch X [K x N] (chains contiguous; M, K, N multiples of 128): operands scaled to
 * 44-bit integers, 13 exact int8 products modulo 13 coprime moduli on tcgen05, Chinese-remainder reconstruction
 * -> DEVICE Y [M x N] fp64 (2^-44 of |A|_row-max |X|_chain-max per term). */
int hmcb_debug_crt_product(int device, int64_t M, int64_t N, int64_t K, const double *A_host,
                           const double *X_dev, double *Y_dev, void *stream);
/* Gathered (block-sparse) slice products (csrc/ozaki_sparse.cuh): the building block of the sparse
 * LinearMatrix products on tcgen05.  A [SA][128 x Ktot]: the dense int8 tiles of `n_bundles` row bundles end to
 * end on the K axis; bundles [n_bundles] = {offset on that axis, k-blocks of 128}; list [Ktot]: the row of B every
 * list entry multiplies; B [SB][rows_b x N] digit planes with the N chains contiguous (N % 128 == 0);
 * C [orders][n_bundles * 128 x N] int32.  All DEVICE arrays. */
int hmcb_debug_i8_gather_gemm(int device, int64_t n_bundles, int64_t N, int64_t Ktot, int64_t rows_b, int SA,
                              int SB, int orders, const signed char *A, const int32_t *bundles,
                              const int32_t *list, const signed char *B, int32_t *C, void *stream);
/* Host side of the slicing (no GPU needed): balanced radix-256 digits of a HOST matrix [rows x cols],
 * a[i][k] = 2^ea[i] * sum_s slices[s][i][k] 256^-(s+1) + rounding; slices [S][rows x cols] int8 (may be
 * NULL), ea [rows].  Returns max over rows of sum_k |rounding| / sum_k |a[i][k]|, or -1 on bad input. */
double hmcb_debug_oz_slice_rows(const double *A, int64_t rows, int64_t cols, int S, signed char *slices,
                                int32_t *ea);

/* target distribution ----------------------------------------------------------------
 * Built from a distribution object tree (BayesRule / Composite / priors / likelihood) by
 * hmclab_b200/_lowering.py.  All HOST arrays. */
int hmcb_clear_target(hmcb_engine *e);
/* Normal: a = means, b = inverse variances; Laplace: a = means, b = inverse dispersions;
 * `constant` = normalization_constant.  Acts on coordinates [offset, offset+len)
 * (CompositeDistribution, base.py:867-898). */
int hmcb_add_prior(hmcb_engine *e, int kind, int64_t offset, int64_t len,
                   const double *a, const double *b, double constant);
/* misfit_bounds of one distribution object (base.py:361-374): misfit += inf when any
 * coordinate of the range is outside [lb, ub]; if in_gradient, the gradient on the range
 * gets +inf too (priors and containers do that, likelihoods do not).  lb / ub: `len`
 * doubles each, either may be NULL. */
int hmcb_add_bound_check(hmcb_engine *e, int64_t offset, int64_t len, const double *lb,
                         const double *ub, int in_gradient);
/* Bounds the trajectory reflects on (corrector: base.py:239-270,913-978,1111-1142);
 * `dims` doubles each (+-inf where unbounded), either may be NULL. */
int hmcb_set_reflection(hmcb_engine *e, const double *lb, const double *ub);

/* LinearMatrix, dense G, premultiplied form (LinearMatrix.py:164-182,185-191,204-206):
 * GtG [dims x dims] row-major, Gtd0 [dims], dtd scalar. */
int hmcb_set_likelihood_dense_premult(hmcb_engine *e, const double *GtG,
                                      const double *Gtd0, double dtd);
/* LinearMatrix, dense G, direct form (LinearMatrix.py:192-202,207-208): G [N x dims]
 * row-major (the dtype-rounded matrix), Gt [dims x N] row-major or NULL when it equals
 * G^T (the reference keeps the caller's un-rounded G^T, :182), d/var/sigma [N]. */
int hmcb_set_likelihood_dense_direct(hmcb_engine *e, int64_t N, const double *G,
                                     const double *Gt, const double *d, const double *var,
                                     const double *sigma);
/* LinearMatrix, dense G, dense (N x N) data covariance, direct form (LinearMatrix.py:257-288):
 * gradient Gt @ invcov @ (G m - d), misfit 0.5 |U (G m - d)|^2 with U the upper Cholesky factor of
 * the inverse covariance.  G [N x dims], GtCinv = Gt @ invcov [dims x N] (the product the reference
 * re-forms in its dtype at every call), d [N], UG = U @ G [N x dims], Ud = U @ d [N]; row-major. */
int hmcb_set_likelihood_dense_direct_cov(hmcb_engine *e, int64_t N, const double *G,
                                         const double *GtCinv, const double *d, const double *UG,
                                         const double *Ud);
/* LinearMatrix, sparse G, direct form (LinearMatrix.py:406-426; replaces the MKL
 * mkl_cspblas_dcsrgemv binding, InterfaceMKL.py:87-121): CSR of G [N x dims] and CSR of
 * G^T [dims x N], int32 indices, float64 values.  Both hold nnz entries; rows need not be
 * sorted by column and an entry may be split over several slots (they are summed).  Values that
 * are all exact in float32 (the reference rounds G to numpy.single by default) are stored in a
 * compact 8-byte form on the device; that changes no result. */
int hmcb_set_likelihood_csr_direct(hmcb_engine *e, int64_t N, int64_t nnz,
                                   const int32_t *indptr, const int32_t *indices,
                                   const double *data, const int32_t *t_indptr,
                                   const int32_t *t_indices, const double *t_data,
                                   const double *d, const double *var, const double *sigma);
/* LinearMatrix, sparse G, premultiplied form (LinearMatrix.py:341-357,390-396,417-419):
 * CSR of the sparse GtG [dims x dims]. */
int hmcb_set_likelihood_csr_premult(hmcb_engine *e, int64_t nnz, const int32_t *indptr,
                                    const int32_t *indices, const double *data,
                                    const double *Gtd0, double dtd);
/* SourceLocation3D (SourceLocation.py:482-540,697-713): stations rx/ry/rz [S],
 * observed travel times and standard deviations [E x S] row-major (NaN = missing pick);
 * dims must be 4*E (+1 if infer_velocity). */
int hmcb_set_likelihood_srcloc3d(hmcb_engine *e, int64_t events, int64_t stations,
                                 const double *rx, const double *ry, const double *rz,
                                 const double *tobs, const double *std, int infer_velocity,
                                 double velocity);

/* SourceLocation2D (SourceLocation.py:100-158): parameters (x, z, T) per event (+ velocity);
 * stations rx / rz [S]; dims must be 3*E (+1).  Missing picks (NaN) are skipped in misfit and
 * gradient alike (the reference's 2-D gradient uses a plain sum and would return NaN). */
int hmcb_set_likelihood_srcloc2d(hmcb_engine *e, int64_t events, int64_t stations,
                                 const double *rx, const double *rz, const double *tobs,
                                 const double *std, int infer_velocity, double velocity);

/* Validate the configuration, upload constants, choose the execution path and allocate
 * workspaces.  Must be called after the setters and before any evaluation. */
int hmcb_finalize(hmcb_engine *e);
int hmcb_path(const hmcb_engine *e);
/* 0, or the number of int8 slice products per gradient evaluation (both products together) when the
 * dense direct products G q and G^T r run as int8 slice products on the tcgen05 tensor cores (Ozaki
 * scheme, csrc/ozaki.cuh; chosen by hmcb_finalize for large problems, HMCB_OZAKI=0 / 1 forces it off /
 * on where valid, HMCB_OZAKI_ORDERS=4..7 sets the orders kept, default 6) */
int hmcb_dense_products_on_tcgen05(const hmcb_engine *e);
/* gradient evaluations per proposal: amount_of_steps x {1,3,4} */
int64_t hmcb_grads_per_proposal(const hmcb_engine *e);
/* number of kernels launched by this engine since creation (bench bookkeeping) */
int64_t hmcb_launch_count(const hmcb_engine *e);
/* Device time of the dominant kernels (bench bookkeeping, measurement rule of the roofline record):
 * between _begin and _end every gradient pass of the likelihood (the DMMA GEMM / strip SpMM
 * launches of one gradient evaluation; for the whole-block fused kernels the one launch of
 * hmcb_run_block) is bracketed by a CUDA event pair on the launching stream.  _end synchronises
 * and returns total_ms[2] / passes[2]: class 0 = gradient passes (or fused blocks), class 1 = the
 * likelihood-misfit passes of the accept/reject step. */
int hmcb_kernel_timing_begin(hmcb_engine *e);
int hmcb_kernel_timing_end(hmcb_engine *e, double *total_ms, int64_t *passes);
/* fp64 roofline denominators measured on `device` (no engine needed): kind 0 = DFMA loop of the
 * SIMT fp64 pipe, 1 = DMMA (mma.sync.m8n8k4.f64) loop of the fp64 tensor path, operands in
 * registers, 2 = fp32 -> fp64 conversions (F2F) next to one DADD each (flops_per_launch then
 * counts conversions); best of `launches` launches of `iters` iterations -> best_ms,
 * flops_per_launch. */
int hmcb_debug_fp64_peak(int device, int kind, int iters, int launches, double *best_ms,
                         double *flops_per_launch);

/* Distributions contract on a batch (DEVICE pointers) ---------------------------------
 * misfit(m) -> float, gradient(m) -> (d,1)  (base.py:71,141), one row per chain. */
int hmcb_misfit(hmcb_engine *e, const double *q, double *x, void *stream);
int hmcb_gradient(hmcb_engine *e, const double *q, double *g, void *stream);
/* corrector(q, p) in place (base.py:239-270) */
int hmcb_reflect(hmcb_engine *e, double *q, double *p, void *stream);
/* mass-matrix protocol (MassMatrices.py:42-69): p = sqrt(M) z ; K(p) ; dK/dp */
int hmcb_scale_momentum(hmcb_engine *e, const double *z, double *p, void *stream);
int hmcb_kinetic_energy(hmcb_engine *e, const double *p, double *k, void *stream);
int hmcb_kinetic_gradient(hmcb_engine *e, const double *p, double *dk, void *stream);

/* A block of HMC proposals for all chains (replaces the body of
 * _AbstractSampler._sample_loop, Samplers.py:579-587,675-678, i.e. HMC._propose :1463 +
 * HMC._evaluate_acceptance :1471 for `proposals` consecutive proposals).  DEVICE pointers. */
typedef struct hmcb_block {
  int64_t proposals;        /* B > 0 */
  int64_t thinning;         /* online_thinning; global proposal k is stored iff k % thinning == 0 */
  int64_t proposal_offset;  /* global index of the first proposal of this block */
  int64_t chain_offset;     /* global id of local chain 0 (keys the on-device RNG) */
  uint64_t seed;            /* Philox key for the on-device RNG */
  double stepsize;          /* > 0 */
  int32_t randomize_stepsize; /* eps = U(0.5,1.5) * stepsize per chain and proposal */
  int32_t reserved;
  double *q;                /* [C x d] in: current models; out: models after the block */
  double *x;                /* [C]     in: current misfits; out: misfits after the block */
  /* injected draws (parity mode); NULL = draw on the device (Philox4x32-10) */
  const double *z_in;       /* [B x C x d] standard normals (momentum = sqrt(M) z) */
  const double *u_step_in;  /* [B x C] uniforms in [0.5,1.5) */
  const double *u_accept_in;/* [B x C] uniforms in [0,1) */
  /* outputs; any may be NULL */
  double *out_samples;      /* [ceil-stored x C x (d+1)] rows [model, misfit], post-decision */
  uint8_t *out_accept;      /* [B x C] 1 = accepted */
  double *out_h0;           /* [B x C] H(current) */
  double *out_h1;           /* [B x C] H(proposed) */
  int32_t *accepted_total;  /* [C] += accepted proposals */
  /* debugging / parity outputs; any may be NULL */
  double *out_q_prop;       /* [B x C x d] end-of-trajectory positions */
  double *out_p_prop;       /* [B x C x d] end-of-trajectory momenta */
  double *trace_q;          /* [B x G x C x d] position at every gradient evaluation */
  double *trace_g;          /* [B x G x C x d] gradient at every gradient evaluation */
  /* per-chain step sizes and their autotuning (HMC.autotune, Samplers.py:1494-1522):
   * eps_c -= (k+1)^-learning_rate * (target_acceptance_rate - min(exp(H0-H1), 1)) after every
   * proposal k (NaN rate counts as 0; a non-positive result is floored at 1e-18). */
  double *stepsize_chain;   /* [C] in/out; NULL = the scalar `stepsize` for every chain */
  int32_t autotune;         /* 1: update stepsize_chain (which must then be non-NULL) */
  int32_t reserved2;
  double target_acceptance_rate;
  double learning_rate;
  double *out_stepsize;     /* [B x C] step size of each proposal before randomisation, or NULL */
} hmcb_block;

int hmcb_run_block(hmcb_engine *e, const hmcb_block *block, void *stream);

/* A block of Random Walk Metropolis-Hastings proposals for all chains (RWMH._propose /
 * _evaluate_acceptance, Samplers.py:1060-1086): proposed = current + (stepsize * step_vector) *
 * normal, accepted iff exp(x - x') > u.  Uses the same hmcb_block (randomize_stepsize, u_step_in,
 * out_p_prop and the traces are ignored; out_h0 / out_h1 receive the current / proposed misfit;
 * autotuning as in Samplers.py:1029-1058).  step_vector: DEVICE [dims] per-coordinate factors
 * (the reference's ndarray stepsize / _stepsize_non_scalar_part) or NULL. */
int hmcb_run_block_rwmh(hmcb_engine *e, const hmcb_block *block, const double *step_vector,
                        void *stream);

/* Whole sampling call with HOST buffers: q0_host [C x d] initial models (pinned or
 * pageable), samples_host [(proposals/thinning) x C x (d+1)] output.  Copies run on
 * private streams and overlap with the kernels of the next block; returns after the
 * last sample has landed.  accept_host [C] (int32 accepted counts) may be NULL. */
int hmcb_sample_host(hmcb_engine *e, const double *q0_host, int64_t proposals,
                     int64_t thinning, int64_t block_proposals, double stepsize,
                     int randomize_stepsize, uint64_t seed, int64_t chain_offset,
                     double *samples_host, int32_t *accept_host, double *final_q_host,
                     double *final_x_host);

#ifdef __cplusplus
}
#endif
#endif /* HMCB_H */

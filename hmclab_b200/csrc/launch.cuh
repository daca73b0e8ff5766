// launch.cuh -- host-callable launchers, one translation unit per kernel family so the
// template instantiations compile in parallel (see hmclab_b200/_build.py).
#pragma once
#include "common.cuh"
#include "fused.cuh"
#include "srcloc.cuh"
#include "staged.cuh"
#include "spmm_types.cuh"

namespace hmcb {

// ---- launch_fused.cu -------------------------------------------------------------------
// Thread mapping of the priors-only fused kernel for `dims` coordinates.
void fused_priors_shape(int dims, int* tpc, int* ppt);
bool fused_priors_supported(int dims);
cudaError_t launch_fused_priors(const FusedArgs& A, cudaStream_t s);
cudaError_t launch_prior_misfit(const DevTarget& T, int chains, const double* q, double* x,
                                const double* lik_misfit, cudaStream_t s);
cudaError_t launch_prior_gradient(const DevTarget& T, int chains, const double* q, double* g,
                                  int accumulate, cudaStream_t s);
cudaError_t launch_reflect(const DevTarget& T, int chains, double* q, double* p, cudaStream_t s);
cudaError_t launch_mass_elementwise(const DevTarget& T, int chains, int mode, const double* in,
                                    double* out, cudaStream_t s);
cudaError_t launch_kinetic_energy(const DevTarget& T, int chains, const double* p, double* k,
                                  cudaStream_t s);

// random walk Metropolis-Hastings (rwmh.cuh, compiled in launch_fused.cu)
struct RwmhDecide;
cudaError_t launch_rwmh_propose(int chains, int dims, const double* q, double* qp, const double* step_vec,
                                const double* step_chain, double stepsize, const double* z_in,
                                unsigned long long seed, long long chain_offset, long long kglob,
                                cudaStream_t s);
cudaError_t launch_rwmh_decide(const RwmhDecide& D, cudaStream_t s);

// ---- launch_srcloc.cu ------------------------------------------------------------------
bool srcloc_supported(int events, int stations);
size_t srcloc_smem_bytes(const SrcLocDev& L);
cudaError_t launch_fused_srcloc(const FusedArgs& A, const SrcLocDev& L, cudaStream_t s);
cudaError_t launch_srcloc_eval(const DevTarget& T, const SrcLocDev& L, int chains, int mode,
                               const double* q, double* out, cudaStream_t s);

// ---- launch_fused_dense.cu -------------------------------------------------------------
// premultiplied dense likelihood with dims <= 128: GtG is the zero padded [128 x 128] matrix
cudaError_t launch_fused_dense(const FusedArgs& A, const double* GtG, const double* Gtd0, double dtd,
                               cudaStream_t s);

// ---- launch_staged.cu ------------------------------------------------------------------
cudaError_t staged_init();  // opt-in shared memory sizes
cudaError_t launch_gemm_update(const double* A, int lda, int M, const double* B, int ldb, int K,
                               const UpdateEpi& epi, cudaStream_t s);
cudaError_t launch_gemm_residual(const double* A, int lda, int M, const double* B, int ldb, int K,
                                 const ResidualEpi& epi, cudaStream_t s);
cudaError_t launch_gemm_misfit(const double* A, int lda, int M, const double* B, int ldb, int K,
                               const MisfitEpi& epi, cudaStream_t s);
cudaError_t launch_gemm_store(const double* A, int lda, int M, const double* B, int ldb, int K,
                              const StoreEpi& epi, cudaStream_t s);
cudaError_t launch_st_kpos(const StagedCommon& S, double a_mult, const double* q_in, double* q_out, double* p,
                           const double* eps, double* k0part, unsigned* flags_out, cudaStream_t s);
cudaError_t launch_st_colsum(const double* part, int tiles, int ld, int C, double scale, double* out,
                             cudaStream_t s);
struct CsrDev {
  const int* indptr;
  const int* indices;
  const double* data;
  int rows;
  int rows_per_chunk;
  int chunks;
};
cudaError_t launch_spmm_update(const CsrDev& M, const double* B, int ldb, const UpdateEpi& epi,
                               cudaStream_t s);
cudaError_t launch_spmm_residual(const CsrDev& M, const double* B, int ldb, const ResidualEpi& epi,
                                 cudaStream_t s);
cudaError_t launch_spmm_misfit(const CsrDev& M, const double* B, int ldb, const MisfitEpi& epi,
                               cudaStream_t s);
// ---- launch_spmm.cu: shared-memory staged SpMM (spmm_strip.cuh) -------------------------------
// thread mappings (consumer warps, rows per warp, chains per lane) the library is built with
struct SpmmShape { int warps, rw, cpl; };
int spmm_strip_shapes(const SpmmShape** out);
bool spmm_block_shape_supported(int warps, int gw, int nb);   // row-blocked kernel (csr_spmm_block_kernel)
cudaError_t spmm_strip_init(const StripDev& M);  // opt-in shared memory size of the mapping
// `bmap`: tensor map of the operand B for M (spmm_strip_tensor_map)
cudaError_t spmm_strip_tensor_map(const StripDev& M, const double* B, int ldb, long long rows_allocated,
                                  CUtensorMap* out);
cudaError_t launch_spmm_strip_update(const StripDev& M, const CUtensorMap& bmap, const double* B, int ldb,
                                     const UpdateEpi& epi, cudaStream_t s);
cudaError_t launch_spmm_strip_residual(const StripDev& M, const CUtensorMap& bmap, const double* B, int ldb,
                                       const ResidualEpi& epi, cudaStream_t s);
cudaError_t launch_spmm_strip_misfit(const StripDev& M, const CUtensorMap& bmap, const double* B, int ldb,
                                     const MisfitEpi& epi, cudaStream_t s);
cudaError_t launch_st_begin(const StagedCommon& S, long long kglob, double a_mult, const double* q_cur,
                            double* q_w, double* p, const double* z_in, const double* u_step_in,
                            const double* u_acc_in, double* eps_out, double* uacc_out, double* k0part,
                            unsigned* flags_out, double* stepsize_out, cudaStream_t s);
cudaError_t launch_st_position(const StagedCommon& S, double a_mult, double* q_w, double* p,
                               const double* eps, unsigned* flags_out, cudaStream_t s);
cudaError_t launch_st_update(const StagedCommon& S, const UpdateEpi& epi, cudaStream_t s);
cudaError_t launch_st_energy(const StagedCommon& S, const double* q, const double* p, double* k1part,
                             double* upart, unsigned* flags_out, cudaStream_t s);
cudaError_t launch_st_decide(const DecideArgs& D, cudaStream_t s);
cudaError_t launch_st_commit(int C, int d, int ld, const unsigned char* acc, const double* q_w,
                             const double* p, double* q_cur, double* sample_rows, double* q_prop,
                             double* p_prop, cudaStream_t s);
cudaError_t launch_st_transpose(const double* in, int R, int Cc, int ldin, double* out, int ldout,
                                cudaStream_t s);

// ---- launch_ozaki.cu: int8-sliced dense products on tcgen05 (ozaki.cuh) ---------------------------
cudaError_t ozaki_init();
cudaError_t ozaki_slice_map(const signed char* base, long long K, long long rows, int slices, int box_rows,
                            CUtensorMap* out);
// digits of the chain batch / orders kept by default (HMCB_OZAKI_ORDERS=7: one more of each); the model
// matrix gets the digits its rows need, at most OZ_SLICES_MAX
constexpr int OZ_SLICES_B = 6, OZ_NUM_ORDERS = 6, OZ_SLICES_MAX = 7, OZ_SLICE_BITS = 8;
double oz_slice_rows_host(const double* A, long long rows, long long cols, long long rows_pad, long long cols_pad,
                          int S, signed char* slices, int* ea);
cudaError_t launch_oz_colmax(const double* X, int rows, int ld, unsigned long long* maxbits, cudaStream_t s);
cudaError_t launch_oz_slice_chains(const double* X, int K, int ld, int SB, const unsigned long long* maxbits,
                                   signed char* out, cudaStream_t s);
cudaError_t launch_oz_combine_residual(const int* C, long long plane_stride, int rows, int ld, int orders,
                                       const int* ea, const unsigned long long* maxbits_in, const ResidualEpi& epi,
                                       unsigned long long* maxbits_out, cudaStream_t s);
cudaError_t launch_oz_combine_update(const int* C, long long plane_stride, int rows, int ld, int orders, const int* ea,
                                     const unsigned long long* maxbits_in, const UpdateEpi& epi, cudaStream_t s);
cudaError_t launch_oz_combine_misfit(const int* C, long long plane_stride, int rows, int ld, int orders, const int* ea,
                                     const unsigned long long* maxbits_in, const MisfitEpi& epi, cudaStream_t s);
cudaError_t ozaki_plane_map(const int* base, long long ldc, long long M, int orders, long long plane_stride,
                            CUtensorMap* out);
cudaError_t launch_i8_gemm_orders(const CUtensorMap& mapA, const CUtensorMap& mapB, const CUtensorMap& mapBh,
                                  const CUtensorMap& mapC, long long M, long long N, long long K, int SA, int SB,
                                  int orders, int ldc, cudaStream_t s);

struct OzBundle;
cudaError_t launch_i8_gemm_moduli(const CUtensorMap& mapA, const CUtensorMap& mapB, const CUtensorMap& mapBh,
                                  long long M, long long N, long long K, int ldc, signed char* res, cudaStream_t s);
void oz_residue_rows_host(const double* A, long long rows, long long cols, long long rows_pad, long long cols_pad,
                          signed char* res, int* ea);
cudaError_t launch_oz_residue_chains(const double* X, int K, int ld, const unsigned long long* maxbits, signed char* out,
                                     cudaStream_t s);
cudaError_t launch_oz_crt_update(const signed char* res, long long plane, int rows, int ld, const int* ea,
                                 const unsigned long long* maxbits_in, const UpdateEpi& epi, cudaStream_t s);
constexpr int OZ_NUM_MODULI = 13;
struct OzPlan;
bool oz_build_plan(int SA, int SB, int orders, long long M, long long N, OzPlan* plan);
cudaError_t ozaki_sparse_init();
cudaError_t launch_i8_gather_gemm(const CUtensorMap& mapA, const CUtensorMap& mapC, int n_bundles, int SA, int SB,
                                  int orders, const OzBundle* bundles, const int* list, const signed char* Bg,
                                  long long bg_plane, int ld, cudaStream_t s);
cudaError_t launch_oz_slice_plain(const double* X, int K, int ld, int SB, const unsigned long long* maxbits,
                                  signed char* out, long long plane, cudaStream_t s);

}  // namespace hmcb

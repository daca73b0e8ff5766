// launch_staged.cu -- launchers of the staged (matrix-likelihood) pipeline: DMMA GEMM and
// CSR SpMM with fused epilogues, and the elementwise trajectory kernels.
#define HMCB_STAGED_KERNELS
#include "launch.cuh"

namespace hmcb {

cudaError_t staged_init() {
  cudaError_t e;
  e = cudaFuncSetAttribute(dmma_gemm_kernel<UpdateEpi>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)GEMM_SMEM_BYTES);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(dmma_gemm_kernel<ResidualEpi>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)GEMM_SMEM_BYTES);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(dmma_gemm_kernel<MisfitEpi>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)GEMM_SMEM_BYTES);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(dmma_gemm_kernel<StoreEpi>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)GEMM_SMEM_BYTES);
}

// M: padded row count of A (multiple of 128); ldb: padded chain count (multiple of 128);
// K: padded inner dimension (multiple of 16).
template <class Epi>
static cudaError_t launch_gemm(const double* A, int lda, int M, const double* B, int ldb, int K,
                               const Epi& epi, cudaStream_t s) {
  if (M % GEMM_BM || ldb % GEMM_BN || K % GEMM_BK) return cudaErrorInvalidValue;
  const dim3 grid(ldb / GEMM_BN, M / GEMM_BM);
  dmma_gemm_kernel<Epi><<<grid, GEMM_LAUNCH_THREADS, GEMM_SMEM_BYTES, s>>>(A, lda, B, ldb, K, epi);
  return cudaGetLastError();
}

cudaError_t launch_gemm_update(const double* A, int lda, int M, const double* B, int ldb, int K,
                               const UpdateEpi& epi, cudaStream_t s) {
  return launch_gemm(A, lda, M, B, ldb, K, epi, s);
}
cudaError_t launch_gemm_residual(const double* A, int lda, int M, const double* B, int ldb, int K,
                                 const ResidualEpi& epi, cudaStream_t s) {
  return launch_gemm(A, lda, M, B, ldb, K, epi, s);
}
cudaError_t launch_gemm_misfit(const double* A, int lda, int M, const double* B, int ldb, int K,
                               const MisfitEpi& epi, cudaStream_t s) {
  return launch_gemm(A, lda, M, B, ldb, K, epi, s);
}

template <class Epi>
static cudaError_t launch_spmm(const CsrDev& M, const double* B, int ldb, const Epi& epi,
                               cudaStream_t s) {
  const dim3 grid(M.chunks, ldb / SPMM_SLAB);  // chunk index fastest: L2-resident chain slabs
  csr_spmm_kernel<Epi><<<grid, SPMM_THREADS, 0, s>>>(M.indptr, M.indices, M.data, M.rows,
                                                     M.rows_per_chunk, B, ldb, epi);
  return cudaGetLastError();
}

cudaError_t launch_spmm_update(const CsrDev& M, const double* B, int ldb, const UpdateEpi& epi,
                               cudaStream_t s) {
  return launch_spmm(M, B, ldb, epi, s);
}
cudaError_t launch_spmm_residual(const CsrDev& M, const double* B, int ldb, const ResidualEpi& epi,
                                 cudaStream_t s) {
  return launch_spmm(M, B, ldb, epi, s);
}
cudaError_t launch_spmm_misfit(const CsrDev& M, const double* B, int ldb, const MisfitEpi& epi,
                               cudaStream_t s) {
  return launch_spmm(M, B, ldb, epi, s);
}

static dim3 st_grid(const StagedCommon& S) { return dim3(S.ld / ST_THREADS, S.jtiles); }

cudaError_t launch_st_begin(const StagedCommon& S, long long kglob, double a_mult, const double* q_cur,
                            double* q_w, double* p, const double* z_in, const double* u_step_in,
                            const double* u_acc_in, double* eps_out, double* uacc_out, double* k0part,
                            unsigned* flags_out, double* stepsize_out, cudaStream_t s) {
  st_begin_kernel<<<st_grid(S), ST_THREADS, 0, s>>>(S, kglob, a_mult, q_cur, q_w, p, z_in, u_step_in,
                                                    u_acc_in, eps_out, uacc_out, k0part, flags_out,
                                                    stepsize_out);
  return cudaGetLastError();
}

cudaError_t launch_st_position(const StagedCommon& S, double a_mult, double* q_w, double* p,
                               const double* eps, unsigned* flags_out, cudaStream_t s) {
  st_position_kernel<<<st_grid(S), ST_THREADS, 0, s>>>(S, a_mult, q_w, p, eps, flags_out);
  return cudaGetLastError();
}

cudaError_t launch_gemm_store(const double* A, int lda, int M, const double* B, int ldb, int K,
                              const StoreEpi& epi, cudaStream_t s) {
  return launch_gemm(A, lda, M, B, ldb, K, epi, s);
}

cudaError_t launch_st_kpos(const StagedCommon& S, double a_mult, const double* q_in, double* q_out, double* p,
                           const double* eps, double* k0part, unsigned* flags_out, cudaStream_t s) {
  st_kpos_kernel<<<st_grid(S), ST_THREADS, 0, s>>>(S, a_mult, q_in, q_out, p, eps, k0part, flags_out);
  return cudaGetLastError();
}

cudaError_t launch_st_colsum(const double* part, int tiles, int ld, int C, double scale, double* out,
                             cudaStream_t s) {
  st_colsum_kernel<<<(C + ST_THREADS - 1) / ST_THREADS, ST_THREADS, 0, s>>>(part, tiles, ld, C, scale, out);
  return cudaGetLastError();
}

cudaError_t launch_st_update(const StagedCommon& S, const UpdateEpi& epi, cudaStream_t s) {
  st_update_kernel<<<st_grid(S), ST_THREADS, 0, s>>>(S, epi);
  return cudaGetLastError();
}

cudaError_t launch_st_energy(const StagedCommon& S, const double* q, const double* p, double* k1part,
                             double* upart, unsigned* flags_out, cudaStream_t s) {
  st_energy_kernel<<<st_grid(S), ST_THREADS, 0, s>>>(S, q, p, k1part, upart, flags_out);
  return cudaGetLastError();
}

cudaError_t launch_st_decide(const DecideArgs& D, cudaStream_t s) {
  st_decide_kernel<<<D.ld / ST_THREADS, ST_THREADS, 0, s>>>(D);
  return cudaGetLastError();
}

cudaError_t launch_st_commit(int C, int d, int ld, const unsigned char* acc, const double* q_w,
                             const double* p, double* q_cur, double* sample_rows, double* q_prop,
                             double* p_prop, cudaStream_t s) {
  const dim3 grid((C + 31) / 32, (d + 31) / 32), block(32, 8);
  st_commit_kernel<<<grid, block, 0, s>>>(C, d, ld, acc, q_w, p, q_cur, sample_rows, q_prop, p_prop);
  return cudaGetLastError();
}

cudaError_t launch_st_transpose(const double* in, int R, int Cc, int ldin, double* out, int ldout,
                                cudaStream_t s) {
  const dim3 grid((Cc + 31) / 32, (R + 31) / 32), block(32, 8);
  st_transpose_kernel<<<grid, block, 0, s>>>(in, R, Cc, ldin, out, ldout);
  return cudaGetLastError();
}

}  // namespace hmcb

// launch_spmm.cu -- launchers of the shared-memory staged CSR SpMM (spmm_strip.cuh).
#include "launch.cuh"
#include "spmm_strip.cuh"

namespace hmcb {

// (consumer warps, rows per warp, chains per lane); index 0 is the default mapping
static const SpmmShape kShapes[] = {{20, 12, 2}, {16, 16, 2}, {28, 16, 1}, {16, 32, 1}};

int spmm_strip_shapes(const SpmmShape** out) {
  *out = kShapes;
  return (int)(sizeof(kShapes) / sizeof(kShapes[0]));
}

#define HMCB_SPMM_DISPATCH(M, CALL)                                                    \
  do {                                                                                 \
    if ((M).warps == 16 && (M).rw == 16 && (M).cpl == 2) { CALL(16, 16, 2); }          \
    else if ((M).warps == 20 && (M).rw == 12 && (M).cpl == 2) { CALL(20, 12, 2); }     \
    else if ((M).warps == 28 && (M).rw == 16 && (M).cpl == 1) { CALL(28, 16, 1); }     \
    else if ((M).warps == 16 && (M).rw == 32 && (M).cpl == 1) { CALL(16, 32, 1); }     \
    else return cudaErrorInvalidValue;                                                 \
  } while (0)

template <class Epi>
static cudaError_t init_one(const StripDev& M) {
  const int bytes = M.stages * M.stage_bytes;
#define HMCB_CALL(W, R, P)                                                                       \
  return cudaFuncSetAttribute(csr_spmm_strip_kernel<Epi, W, R, P>,                               \
                              cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)
  HMCB_SPMM_DISPATCH(M, HMCB_CALL);
#undef HMCB_CALL
  return cudaSuccess;
}

cudaError_t spmm_strip_init(const StripDev& M) {
  cudaError_t e;
  if ((e = init_one<UpdateEpi>(M)) != cudaSuccess) return e;
  if ((e = init_one<ResidualEpi>(M)) != cudaSuccess) return e;
  return init_one<MisfitEpi>(M);
}

template <class Epi>
static cudaError_t launch_strip(const StripDev& M, const double* B, int ldb, const Epi& epi, cudaStream_t s) {
  const int S = 32 * M.cpl;
  if (ldb % S) return cudaErrorInvalidValue;
  const dim3 grid(M.chunks, ldb / S);   // chunk index fastest: blocks in flight share a chain slab in L2
  const size_t smem = (size_t)M.stages * M.stage_bytes;
#define HMCB_CALL(W, R, P) \
  csr_spmm_strip_kernel<Epi, W, R, P><<<grid, (W + SPMM_PRODUCERS) * 32, smem, s>>>(M, B, ldb, epi)
  HMCB_SPMM_DISPATCH(M, HMCB_CALL);
#undef HMCB_CALL
  return cudaGetLastError();
}

cudaError_t launch_spmm_strip_update(const StripDev& M, const double* B, int ldb, const UpdateEpi& epi,
                                     cudaStream_t s) {
  return launch_strip(M, B, ldb, epi, s);
}
cudaError_t launch_spmm_strip_residual(const StripDev& M, const double* B, int ldb, const ResidualEpi& epi,
                                       cudaStream_t s) {
  return launch_strip(M, B, ldb, epi, s);
}
cudaError_t launch_spmm_strip_misfit(const StripDev& M, const double* B, int ldb, const MisfitEpi& epi,
                                     cudaStream_t s) {
  return launch_strip(M, B, ldb, epi, s);
}

}  // namespace hmcb

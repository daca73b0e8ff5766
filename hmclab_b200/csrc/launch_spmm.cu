// launch_spmm.cu -- launchers of the shared-memory staged CSR SpMM (spmm_strip.cuh).
#include <cudaTypedefs.h>

#include "launch.cuh"
#include "spmm_strip.cuh"

namespace hmcb {

// (consumer warps, rows per warp, chains per lane); index 0 is the default mapping
static const SpmmShape kShapes[] = {{31, 8, 2}, {27, 8, 2}, {23, 12, 2}, {20, 12, 2}, {19, 16, 2}};

int spmm_strip_shapes(const SpmmShape** out) {
  *out = kShapes;
  return (int)(sizeof(kShapes) / sizeof(kShapes[0]));
}

#define HMCB_SPMM_DISPATCH(M, CALL)                                                    \
  do {                                                                                 \
    if ((M).warps == 20 && (M).rw == 12 && (M).cpl == 2) { CALL(20, 12, 2); }          \
    else if ((M).warps == 19 && (M).rw == 16 && (M).cpl == 2) { CALL(19, 16, 2); }     \
    else if ((M).warps == 23 && (M).rw == 12 && (M).cpl == 2) { CALL(23, 12, 2); }     \
    else if ((M).warps == 27 && (M).rw == 8 && (M).cpl == 2) { CALL(27, 8, 2); }       \
    else if ((M).warps == 31 && (M).rw == 8 && (M).cpl == 2) { CALL(31, 8, 2); }       \
    else return cudaErrorInvalidValue;                                                 \
  } while (0)

// (consumer warps, groups per warp, 16-chain boxes per slab) of the row-blocked tensor-core kernel
#define HMCB_SPMM_BLOCK_DISPATCH(M, CALL)                                              \
  do {                                                                                 \
    if ((M).warps == 31 && (M).gw == 4 && (M).nb == 1) { CALL(31, 4, 1); }             \
    else if ((M).warps == 31 && (M).gw == 2 && (M).nb == 1) { CALL(31, 2, 1); }        \
    else if ((M).warps == 31 && (M).gw == 2 && (M).nb == 2) { CALL(31, 2, 2); }        \
    else if ((M).warps == 23 && (M).gw == 4 && (M).nb == 2) { CALL(23, 4, 2); }        \
    else if ((M).warps == 19 && (M).gw == 4 && (M).nb == 2) { CALL(19, 4, 2); }        \
    else if ((M).warps == 15 && (M).gw == 4 && (M).nb == 2) { CALL(15, 4, 2); }        \
    else if ((M).warps == 15 && (M).gw == 8 && (M).nb == 1) { CALL(15, 8, 1); }        \
    else if ((M).warps == 3 && (M).gw == 2 && (M).nb == 1) { CALL(3, 2, 1); }          \
    else return cudaErrorInvalidValue;                                                 \
  } while (0)

bool spmm_block_shape_supported(int warps, int gw, int nb) {
  return (warps == 31 && gw == 4 && nb == 1) || (warps == 31 && gw == 2 && nb == 1) ||
         (warps == 31 && gw == 2 && nb == 2) || (warps == 15 && gw == 4 && nb == 2) ||
         (warps == 23 && gw == 4 && nb == 2) || (warps == 19 && gw == 4 && nb == 2) ||
         (warps == 15 && gw == 8 && nb == 1) || (warps == 3 && gw == 2 && nb == 1);
}

template <class Epi>
static cudaError_t init_one(const StripDev& M) {
  const int bytes = M.stages * M.stage_bytes;
  if (M.blocked) {
    const int bbytes = bytes + 1024;   // the kernel rounds its window up to a 1024-byte swizzle atom
#define HMCB_CALL(W, G, N)                                                                              \
  return M.compact ? cudaFuncSetAttribute(csr_spmm_block_kernel<Epi, W, G, N, true>,                    \
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, bbytes)          \
                   : cudaFuncSetAttribute(csr_spmm_block_kernel<Epi, W, G, N, false>,                   \
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, bbytes)
    HMCB_SPMM_BLOCK_DISPATCH(M, HMCB_CALL);
#undef HMCB_CALL
  }
#define HMCB_CALL(W, R, P)                                                                       \
  return M.compact ? cudaFuncSetAttribute(csr_spmm_strip_kernel<Epi, W, R, P, true>,             \
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)    \
                   : cudaFuncSetAttribute(csr_spmm_strip_kernel<Epi, W, R, P, false>,            \
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

// B [rows x ldb] (chains contiguous) as the 3-D tensor {ldb chains, T strips, ceil(rows / T) rows per
// strip}: row c of B is (strip c % T, local row c / T).  The buffer must hold T rows of slack
// behind `rows_allocated - T` (hmcb_finalize allocates them) so that every (strip, local row)
// inside the tensor is backed by memory.
cudaError_t spmm_strip_tensor_map(const StripDev& M, const double* B, int ldb, long long rows_allocated,
                                  CUtensorMap* out) {
  static PFN_cuTensorMapEncodeTiled_v12000 encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    if (e != cudaSuccess) return e;
    if (q != cudaDriverEntryPointSuccess || !fn) return cudaErrorNotSupported;
    encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  }
  const unsigned long long T = (unsigned long long)M.cstride;
  const unsigned long long J = (unsigned long long)(rows_allocated / (long long)T);
  if (J < (unsigned long long)M.kb_box) return cudaErrorInvalidValue;
  const cuuint64_t dims[3] = {(cuuint64_t)ldb, T, J};
  const cuuint64_t strides[2] = {(cuuint64_t)ldb * 8ull, T * (cuuint64_t)ldb * 8ull};
  // row-blocked kernel: boxes of 16 chains (128 bytes) x box_rows rows, 128-byte swizzle (rows of a box
  // past the end of the tensor are zero filled); plain kernel: one un-swizzled box per strip
  const cuuint32_t box[3] = {(cuuint32_t)(M.blocked ? 16 : 32 * M.cpl), 1u,
                             (cuuint32_t)(M.blocked ? M.box_rows : M.kb_box)};
  const cuuint32_t estr[3] = {1u, 1u, 1u};
  const CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double*>(B), dims, strides, box,
                            estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            M.blocked ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

template <class Epi>
static cudaError_t launch_strip(const StripDev& M, const CUtensorMap& bmap, const double* B, int ldb,
                                const Epi& epi, cudaStream_t s) {
  const size_t smem = (size_t)M.stages * M.stage_bytes;
  if (M.blocked) {
    // slabs of 16 nb chains; the slabs of 128 chains are interleaved per chunk (x), chunk index next
    if (ldb % SPMM_SLAB_CHAINS_MINOR) return cudaErrorInvalidValue;
    const int sl = SPMM_SLAB_CHAINS_MINOR / (16 * M.nb);
    const dim3 bgrid(M.chunks * sl, ldb / SPMM_SLAB_CHAINS_MINOR);
#define HMCB_CALL(W, G, N)                                                                                       \
  if (M.compact)                                                                                                \
    csr_spmm_block_kernel<Epi, W, G, N, true><<<bgrid, (W + SPMM_PRODUCERS) * 32, smem + 1024, s>>>(M, bmap, epi); \
  else                                                                                                          \
    csr_spmm_block_kernel<Epi, W, G, N, false><<<bgrid, (W + SPMM_PRODUCERS) * 32, smem + 1024, s>>>(M, bmap, epi)
    HMCB_SPMM_BLOCK_DISPATCH(M, HMCB_CALL);
#undef HMCB_CALL
    return cudaGetLastError();
  }
  const int S = 32 * M.cpl;
  if (ldb % S) return cudaErrorInvalidValue;
  const dim3 grid(M.chunks, ldb / S);   // chunk index fastest: blocks in flight share a chain slab in L2
#define HMCB_CALL(W, R, P)                                                                              \
  if (M.compact)                                                                                       \
    csr_spmm_strip_kernel<Epi, W, R, P, true><<<grid, (W + SPMM_PRODUCERS) * 32, smem, s>>>(M, bmap, epi); \
  else                                                                                                 \
    csr_spmm_strip_kernel<Epi, W, R, P, false><<<grid, (W + SPMM_PRODUCERS) * 32, smem, s>>>(M, bmap, epi)
  HMCB_SPMM_DISPATCH(M, HMCB_CALL);
#undef HMCB_CALL
  return cudaGetLastError();
}

cudaError_t launch_spmm_strip_update(const StripDev& M, const CUtensorMap& bmap, const double* B, int ldb,
                                     const UpdateEpi& epi, cudaStream_t s) {
  return launch_strip(M, bmap, B, ldb, epi, s);
}
cudaError_t launch_spmm_strip_residual(const StripDev& M, const CUtensorMap& bmap, const double* B, int ldb,
                                       const ResidualEpi& epi, cudaStream_t s) {
  return launch_strip(M, bmap, B, ldb, epi, s);
}
cudaError_t launch_spmm_strip_misfit(const StripDev& M, const CUtensorMap& bmap, const double* B, int ldb,
                                     const MisfitEpi& epi, cudaStream_t s) {
  return launch_strip(M, bmap, B, ldb, epi, s);
}

}  // namespace hmcb

// launch_ozaki.cu -- launchers of the int8-sliced (Ozaki) dense products on tcgen05 (ozaki.cuh).
#include <cudaTypedefs.h>

#include "launch.cuh"
#include "ozaki.cuh"

namespace hmcb {

static PFN_cuTensorMapEncodeTiled_v12000 tensor_map_encoder() {
  static PFN_cuTensorMapEncodeTiled_v12000 encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  }
  return encode;
}

// slices [S][rows x K] int8, K contiguous -> 3-D tensor {K, rows, S}, box {128, box_rows, 1}, 128-byte swizzle
cudaError_t ozaki_slice_map(const signed char* base, long long K, long long rows, int slices, int box_rows,
                            CUtensorMap* out) {
  PFN_cuTensorMapEncodeTiled_v12000 encode = tensor_map_encoder();
  if (!encode) return cudaErrorNotSupported;
  const cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)slices};
  const cuuint64_t strides[2] = {(cuuint64_t)K, (cuuint64_t)K * (cuuint64_t)rows};
  const cuuint32_t box[3] = {(cuuint32_t)OZ_BK, (cuuint32_t)box_rows, 1u};
  const cuuint32_t estr[3] = {1u, 1u, 1u};
  const CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<signed char*>(base), dims, strides, box,
                            estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

cudaError_t ozaki_init() {
  return cudaFuncSetAttribute(i8_gemm_orders_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)OZ_SMEM_BYTES);
}

// C[o] (o < orders) = sum_{s+t=o} A_s B_t^T;  M % 128 == 0, K % 128 == 0, ldc = padded N (% 128 == 0)
cudaError_t launch_i8_gemm_orders(const CUtensorMap& mapA, const CUtensorMap& mapB, long long M, long long N,
                                  long long K, int SA, int SB, int orders, int* C, long long plane_stride, int ldc,
                                  cudaStream_t s) {
  if (M % OZ_BM || K % OZ_BK || N % 128 || orders < 1 || orders > SA + SB - 1) return cudaErrorInvalidValue;
  const dim3 grid((unsigned)((N + OZ_BN - 1) / OZ_BN), (unsigned)(M / OZ_BM), (unsigned)orders);
  i8_gemm_orders_kernel<<<grid, OZ_THREADS, OZ_SMEM_BYTES, s>>>(mapA, mapB, (int)(K / OZ_BK), SA, SB, C, plane_stride,
                                                               ldc);
  return cudaGetLastError();
}

cudaError_t launch_oz_colmax(const double* X, int rows, int ld, unsigned long long* maxbits, cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(maxbits, 0, (size_t)ld * sizeof(unsigned long long), s);
  if (e != cudaSuccess) return e;
  oz_colmax_kernel<<<dim3(ld / 128, (rows + 63) / 64), 128, 0, s>>>(X, rows, ld, maxbits);
  return cudaGetLastError();
}

cudaError_t launch_oz_slice_chains(const double* X, int K, int ld, const unsigned long long* maxbits,
                                   signed char* out, cudaStream_t s) {
  if (K % 128 || ld % 32) return cudaErrorInvalidValue;
  oz_slice_chains_kernel<<<dim3(ld / 32, K / 128), 256, 0, s>>>(X, K, ld, maxbits, out);
  return cudaGetLastError();
}

cudaError_t launch_oz_combine_residual(const int* C, long long plane_stride, int rows, int ld, const int* ea,
                                       const unsigned long long* maxbits_in, const ResidualEpi& epi,
                                       unsigned long long* maxbits_out, cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(maxbits_out, 0, (size_t)ld * sizeof(unsigned long long), s);
  if (e != cudaSuccess) return e;
  oz_combine_kernel<ResidualEpi, true><<<dim3(ld / 128, (rows + 15) / 16), 128, 0, s>>>(C, plane_stride, rows, ld, ea,
                                                                                      maxbits_in, epi, maxbits_out);
  return cudaGetLastError();
}

cudaError_t launch_oz_combine_update(const int* C, long long plane_stride, int rows, int ld, const int* ea,
                                     const unsigned long long* maxbits_in, const UpdateEpi& epi, cudaStream_t s) {
  oz_combine_kernel<UpdateEpi, false><<<dim3(ld / 128, (rows + 15) / 16), 128, 0, s>>>(C, plane_stride, rows, ld, ea,
                                                                                     maxbits_in, epi, nullptr);
  return cudaGetLastError();
}

}  // namespace hmcb

// Test / measurement entry: exact int8 slice products on the tcgen05 tensor cores.
extern "C" int hmcb_debug_i8_gemm(int device, int64_t M, int64_t N, int64_t K, int SA, int SB, int orders,
                                  const signed char* A, const signed char* B, int32_t* C, void* stream) {
  using namespace hmcb;
  if (!A || !B || !C || M <= 0 || N <= 0 || K <= 0 || SA < 1 || SB < 1) return -1;
  if (cudaSetDevice(device) != cudaSuccess) return -1;
  if (ozaki_init() != cudaSuccess) return -2;
  CUtensorMap mapA, mapB;
  if (ozaki_slice_map(A, K, M, SA, OZ_BM, &mapA) != cudaSuccess) return -3;
  if (ozaki_slice_map(B, K, N, SB, OZ_BN, &mapB) != cudaSuccess) return -3;
  if (launch_i8_gemm_orders(mapA, mapB, M, N, K, SA, SB, orders, C, M * N, (int)N, static_cast<cudaStream_t>(stream)) !=
      cudaSuccess)
    return -4;
  return 0;
}

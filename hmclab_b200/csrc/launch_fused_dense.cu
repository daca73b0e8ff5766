// launch_fused_dense.cu -- launcher of the fused small-dense (premultiplied GtG, dims <= 128) kernel.
#include "fused_dense.cuh"
#include "launch.cuh"

namespace hmcb {

cudaError_t launch_fused_dense(const FusedArgs& A, const double* GtG, const double* Gtd0, double dtd,
                               cudaStream_t s) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(hmc_fused_dense_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)FD_SMEM_BYTES);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  FusedDenseArgs D{A, GtG, Gtd0, dtd};
  const int grid = (A.chains + FD_BN - 1) / FD_BN;
  hmc_fused_dense_kernel<<<grid, FD_THREADS, FD_SMEM_BYTES, s>>>(D);
  return cudaGetLastError();
}

}  // namespace hmcb

// launch_fused_ppt.cu -- instantiations of the priors-only fused HMC kernel for one value
// of PPT (coordinate pairs per thread); compiled once per -DHMCB_PPT={1,2,4}.
#include "launch.cuh"

#ifndef HMCB_PPT
#error "compile with -DHMCB_PPT=<pairs per thread>"
#endif
#define HMCB_CAT2(a, b) a##b
#define HMCB_CAT(a, b) HMCB_CAT2(a, b)

namespace hmcb {

template <int TPC, int PPT>
static cudaError_t launch_fp(const FusedArgs& A, cudaStream_t s) {
  constexpr int BLOCK = TPC < 256 ? 256 : TPC;
  constexpr int CPB = BLOCK / TPC;
  const int grid = (A.chains + CPB - 1) / CPB;
  // AUX: reflection bounds and/or a diagonal mass matrix live in registers too
  const bool aux = A.T.invm || A.T.refl_lb || A.T.refl_ub;
  if (A.exact) {
    if (aux) hmc_fused_priors_kernel<TPC, PPT, true, false><<<grid, BLOCK, 0, s>>>(A);
    else hmc_fused_priors_kernel<TPC, PPT, false, false><<<grid, BLOCK, 0, s>>>(A);
  } else {
    if (aux) hmc_fused_priors_kernel<TPC, PPT, true, true><<<grid, BLOCK, 0, s>>>(A);
    else hmc_fused_priors_kernel<TPC, PPT, false, true><<<grid, BLOCK, 0, s>>>(A);
  }
  return cudaGetLastError();
}

cudaError_t HMCB_CAT(launch_fused_priors_ppt, HMCB_PPT)(const FusedArgs& A, int tpc, cudaStream_t s) {
  switch (tpc) {
#if HMCB_PPT == 1
    case 2: return launch_fp<2, 1>(A, s);
    case 8: return launch_fp<8, 1>(A, s);
    case 32: return launch_fp<32, 1>(A, s);
    case 64: return launch_fp<64, 1>(A, s);
    case 128: return launch_fp<128, 1>(A, s);
#elif HMCB_PPT == 2
    case 128: return launch_fp<128, 2>(A, s);
    case 256: return launch_fp<256, 2>(A, s);
#elif HMCB_PPT == 4
    case 128: return launch_fp<128, 4>(A, s);
    case 256: return launch_fp<256, 4>(A, s);
#endif
  }
  return cudaErrorInvalidConfiguration;
}

}  // namespace hmcb

// launch_srcloc_lpe.cu -- instantiations of the SourceLocation3D kernels for one value of
// LPE (lanes per event); compiled once per -DHMCB_LPE={1,2,4} x -DHMCB_NP={3,4}.
#include "launch.cuh"

#if !defined(HMCB_LPE) || !defined(HMCB_NP)
#error "compile with -DHMCB_LPE=<lanes per event> -DHMCB_NP=<parameters per event: 3 or 4>"
#endif
#define HMCB_CAT2(a, b) a##b
#define HMCB_CAT(a, b) HMCB_CAT2(a, b)
#define HMCB_CAT4(a, b, c, d) HMCB_CAT(HMCB_CAT(a, b), HMCB_CAT(c, d))

namespace hmcb {

template <int TPC, int LPE>
static cudaError_t launch_fs(const FusedArgs& A, const SrcLocDev& L, cudaStream_t s) {
  constexpr int BLOCK = TPC <= 32 ? SRCLOC_SMALL_BLOCK : TPC;
  constexpr int CPB = BLOCK / TPC;
  const size_t smem = srcloc_smem_bytes(L);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(hmc_fused_srcloc_kernel<TPC, LPE, HMCB_NP>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  hmc_fused_srcloc_kernel<TPC, LPE, HMCB_NP><<<(A.chains + CPB - 1) / CPB, BLOCK, smem, s>>>(A, L);
  return cudaGetLastError();
}

template <int TPC, int LPE>
static cudaError_t launch_ev(const DevTarget& T, const SrcLocDev& L, int chains, int mode,
                             const double* q, double* out, cudaStream_t s) {
  constexpr int BLOCK = TPC <= 32 ? SRCLOC_SMALL_BLOCK : TPC;
  constexpr int CPB = BLOCK / TPC;
  const size_t smem = srcloc_smem_bytes(L);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(srcloc_eval_kernel<TPC, LPE, HMCB_NP>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  srcloc_eval_kernel<TPC, LPE, HMCB_NP><<<(chains + CPB - 1) / CPB, BLOCK, smem, s>>>(T, L, chains, mode, q, out);
  return cudaGetLastError();
}

#define HMCB_EPAD_CASES(CALL)                         \
  switch (epad) {                                     \
    case 1: return CALL(1 * HMCB_LPE);                \
    case 2: return CALL(2 * HMCB_LPE);                \
    case 4: return CALL(4 * HMCB_LPE);                \
    case 8: return CALL(8 * HMCB_LPE);                \
    case 16: return CALL(16 * HMCB_LPE);              \
    case 32: return CALL(32 * HMCB_LPE);              \
    case 64: return CALL(64 * HMCB_LPE);              \
    case 128: return CALL(128 * HMCB_LPE);            \
    case 256: return CALL(256 * HMCB_LPE);            \
  }                                                   \
  return cudaErrorInvalidConfiguration;

cudaError_t HMCB_CAT4(launch_fused_srcloc_lpe, HMCB_LPE, _np, HMCB_NP)(const FusedArgs& A, const SrcLocDev& L,
                                                         int epad, cudaStream_t s) {
#define HMCB_CALL(TPC_) launch_fs<TPC_, HMCB_LPE>(A, L, s)
  HMCB_EPAD_CASES(HMCB_CALL)
#undef HMCB_CALL
}

cudaError_t HMCB_CAT4(launch_srcloc_eval_lpe, HMCB_LPE, _np, HMCB_NP)(const DevTarget& T, const SrcLocDev& L,
                                                        int chains, int mode, const double* q,
                                                        double* out, int epad, cudaStream_t s) {
#define HMCB_CALL(TPC_) launch_ev<TPC_, HMCB_LPE>(T, L, chains, mode, q, out, s)
  HMCB_EPAD_CASES(HMCB_CALL)
#undef HMCB_CALL
}

}  // namespace hmcb

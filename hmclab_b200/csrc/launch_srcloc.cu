// launch_srcloc.cu -- dispatch of the SourceLocation3D kernels (instantiated per LPE in
// launch_srcloc_lpe.cu).
#include "launch.cuh"

#include <cstdlib>

namespace hmcb {

static int pow2_ceil(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

// Lanes per event: each lane walks stations sub, sub+LPE, ...
static int srcloc_lpe(int stations) {
  // measured on config 5 (30 stations): 1 lane per event 5.5e8, 2 lanes 4.6e8, 4 lanes 2.6e8 evals/s
  int lpe = stations <= 48 ? 1 : (stations <= 128 ? 2 : 4);
  if (const char* env = std::getenv("HMCB_SRCLOC_LPE")) {
    const int v = std::atoi(env);
    if (v == 1 || v == 2 || v == 4) lpe = v;
  }
  return lpe;
}

size_t srcloc_smem_bytes(const SrcLocDev& L) {
  return sizeof(double) * srcloc_smem_doubles(L.events, L.stations);
}

bool srcloc_supported(int events, int stations) {
  if (events < 1 || stations < 1) return false;
  if (pow2_ceil(events) * srcloc_lpe(stations) > 1024) return false;
  return sizeof(double) * srcloc_smem_doubles(events, stations) <= 200 * 1024;
}

#define HMCB_DECLARE(L_, N_)                                                                              \
  cudaError_t launch_fused_srcloc_lpe##L_##_np##N_(const FusedArgs&, const SrcLocDev&, int, cudaStream_t); \
  cudaError_t launch_srcloc_eval_lpe##L_##_np##N_(const DevTarget&, const SrcLocDev&, int, int,            \
                                                  const double*, double*, int, cudaStream_t);
HMCB_DECLARE(1, 3) HMCB_DECLARE(2, 3) HMCB_DECLARE(4, 3)
HMCB_DECLARE(1, 4) HMCB_DECLARE(2, 4) HMCB_DECLARE(4, 4)
#undef HMCB_DECLARE

cudaError_t launch_fused_srcloc(const FusedArgs& A, const SrcLocDev& L, cudaStream_t s) {
  const int epad = pow2_ceil(L.events);
  const int key = srcloc_lpe(L.stations) * 10 + L.np;
  switch (key) {
    case 13: return launch_fused_srcloc_lpe1_np3(A, L, epad, s);
    case 23: return launch_fused_srcloc_lpe2_np3(A, L, epad, s);
    case 43: return launch_fused_srcloc_lpe4_np3(A, L, epad, s);
    case 14: return launch_fused_srcloc_lpe1_np4(A, L, epad, s);
    case 24: return launch_fused_srcloc_lpe2_np4(A, L, epad, s);
    case 44: return launch_fused_srcloc_lpe4_np4(A, L, epad, s);
  }
  return cudaErrorInvalidConfiguration;
}

cudaError_t launch_srcloc_eval(const DevTarget& T, const SrcLocDev& L, int chains, int mode,
                               const double* q, double* out, cudaStream_t s) {
  const int epad = pow2_ceil(L.events);
  const int key = srcloc_lpe(L.stations) * 10 + L.np;
  switch (key) {
    case 13: return launch_srcloc_eval_lpe1_np3(T, L, chains, mode, q, out, epad, s);
    case 23: return launch_srcloc_eval_lpe2_np3(T, L, chains, mode, q, out, epad, s);
    case 43: return launch_srcloc_eval_lpe4_np3(T, L, chains, mode, q, out, epad, s);
    case 14: return launch_srcloc_eval_lpe1_np4(T, L, chains, mode, q, out, epad, s);
    case 24: return launch_srcloc_eval_lpe2_np4(T, L, chains, mode, q, out, epad, s);
    case 44: return launch_srcloc_eval_lpe4_np4(T, L, chains, mode, q, out, epad, s);
  }
  return cudaErrorInvalidConfiguration;
}

}  // namespace hmcb

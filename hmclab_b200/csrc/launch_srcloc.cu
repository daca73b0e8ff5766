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

cudaError_t launch_fused_srcloc_lpe1(const FusedArgs&, const SrcLocDev&, int, cudaStream_t);
cudaError_t launch_fused_srcloc_lpe2(const FusedArgs&, const SrcLocDev&, int, cudaStream_t);
cudaError_t launch_fused_srcloc_lpe4(const FusedArgs&, const SrcLocDev&, int, cudaStream_t);
cudaError_t launch_srcloc_eval_lpe1(const DevTarget&, const SrcLocDev&, int, int, const double*, double*, int, cudaStream_t);
cudaError_t launch_srcloc_eval_lpe2(const DevTarget&, const SrcLocDev&, int, int, const double*, double*, int, cudaStream_t);
cudaError_t launch_srcloc_eval_lpe4(const DevTarget&, const SrcLocDev&, int, int, const double*, double*, int, cudaStream_t);

cudaError_t launch_fused_srcloc(const FusedArgs& A, const SrcLocDev& L, cudaStream_t s) {
  const int epad = pow2_ceil(L.events);
  switch (srcloc_lpe(L.stations)) {
    case 1: return launch_fused_srcloc_lpe1(A, L, epad, s);
    case 2: return launch_fused_srcloc_lpe2(A, L, epad, s);
    case 4: return launch_fused_srcloc_lpe4(A, L, epad, s);
  }
  return cudaErrorInvalidConfiguration;
}

cudaError_t launch_srcloc_eval(const DevTarget& T, const SrcLocDev& L, int chains, int mode,
                               const double* q, double* out, cudaStream_t s) {
  const int epad = pow2_ceil(L.events);
  switch (srcloc_lpe(L.stations)) {
    case 1: return launch_srcloc_eval_lpe1(T, L, chains, mode, q, out, epad, s);
    case 2: return launch_srcloc_eval_lpe2(T, L, chains, mode, q, out, epad, s);
    case 4: return launch_srcloc_eval_lpe4(T, L, chains, mode, q, out, epad, s);
  }
  return cudaErrorInvalidConfiguration;
}

}  // namespace hmcb

// launch_fused.cu -- instantiations and dispatch of the priors-only fused HMC kernel and of
// the standalone prior / mass-matrix kernels.
#define HMCB_FUSED_AUX_KERNELS
#include "launch.cuh"
#include "rwmh.cuh"

#include <cstdio>
#include <cstdlib>

namespace hmcb {

// Instantiated (TPC, PPT) shapes of the priors-only fused kernel: pairs = ceil(dims / 2)
// coordinate pairs are spread over TPC threads x PPT pairs each.
static const int kShapes[][2] = {{2, 1},   {8, 1},   {32, 1},  {64, 1}, {128, 1},
                                 {128, 2}, {256, 2}, {128, 4}, {256, 4}};

static bool shape_instantiated(int tpc, int ppt) {
  for (const auto& sh : kShapes)
    if (sh[0] == tpc && sh[1] == ppt) return true;
  return false;
}

void fused_priors_shape(int dims, int* tpc, int* ppt) {
  const int pairs = (dims + 1) / 2;
  int T = 0, P = 0;
  if (pairs <= 2) { T = 2; P = 1; }
  else if (pairs <= 8) { T = 8; P = 1; }
  else if (pairs <= 32) { T = 32; P = 1; }
  else if (pairs <= 64) { T = 64; P = 1; }
  else if (pairs <= 128) { T = 128; P = 1; }
  else if (pairs <= 256) { T = 128; P = 2; }
  else if (pairs <= 512) { T = 128; P = 4; }   // 1000 dims: measured best of the instantiated shapes
  else if (pairs <= 1024) { T = 256; P = 4; }  // beyond 2048 dims the staged path takes over
  if (const char* env = std::getenv("HMCB_FUSED_SHAPE")) {  // tuning override "TPC,PPT"
    int t = 0, p = 0;
    if (std::sscanf(env, "%d,%d", &t, &p) == 2 && shape_instantiated(t, p) && t * p >= pairs) { T = t; P = p; }
  }
  *tpc = T;
  *ppt = P;
}

bool fused_priors_supported(int dims) {
  int t, p;
  fused_priors_shape(dims, &t, &p);
  return t > 0;
}

// One translation unit per PPT (launch_fused_ppt.cu compiled with -DHMCB_PPT=n).
cudaError_t launch_fused_priors_ppt1(const FusedArgs& A, int tpc, cudaStream_t s);
cudaError_t launch_fused_priors_ppt2(const FusedArgs& A, int tpc, cudaStream_t s);
cudaError_t launch_fused_priors_ppt4(const FusedArgs& A, int tpc, cudaStream_t s);

cudaError_t launch_fused_priors(const FusedArgs& A, cudaStream_t s) {
  int tpc, ppt;
  fused_priors_shape(A.T.dims, &tpc, &ppt);
  if (A.T.n_terms > 1) return cudaErrorInvalidConfiguration;
  switch (ppt) {
    case 1: return launch_fused_priors_ppt1(A, tpc, s);
    case 2: return launch_fused_priors_ppt2(A, tpc, s);
    case 4: return launch_fused_priors_ppt4(A, tpc, s);
  }
  return cudaErrorInvalidConfiguration;
}

// The standalone kernels loop over coordinates, so two thread shapes are enough.
#define HMCB_TPC_DISPATCH(dims, CALL32, CALL256) \
  if ((dims) <= 256) { CALL32; } else { CALL256; }

cudaError_t launch_prior_misfit(const DevTarget& T, int chains, const double* q, double* x,
                                const double* lik_misfit, cudaStream_t s) {
  HMCB_TPC_DISPATCH(T.dims,
                    (prior_misfit_kernel<32><<<(chains + 7) / 8, 256, 0, s>>>(T, chains, q, x, lik_misfit)),
                    (prior_misfit_kernel<256><<<chains, 256, 0, s>>>(T, chains, q, x, lik_misfit)))
  return cudaGetLastError();
}

cudaError_t launch_prior_gradient(const DevTarget& T, int chains, const double* q, double* g,
                                  int accumulate, cudaStream_t s) {
  HMCB_TPC_DISPATCH(T.dims,
                    (prior_gradient_kernel<32><<<(chains + 7) / 8, 256, 0, s>>>(T, chains, q, g, accumulate)),
                    (prior_gradient_kernel<256><<<chains, 256, 0, s>>>(T, chains, q, g, accumulate)))
  return cudaGetLastError();
}

cudaError_t launch_kinetic_energy(const DevTarget& T, int chains, const double* p, double* k,
                                  cudaStream_t s) {
  HMCB_TPC_DISPATCH(T.dims,
                    (kinetic_energy_kernel<32><<<(chains + 7) / 8, 256, 0, s>>>(T, chains, p, k)),
                    (kinetic_energy_kernel<256><<<chains, 256, 0, s>>>(T, chains, p, k)))
  return cudaGetLastError();
}

cudaError_t launch_reflect(const DevTarget& T, int chains, double* q, double* p, cudaStream_t s) {
  const size_t total = (size_t)chains * T.dims;
  reflect_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(T, total, q, p);
  return cudaGetLastError();
}

cudaError_t launch_mass_elementwise(const DevTarget& T, int chains, int mode, const double* in,
                                    double* out, cudaStream_t s) {
  const size_t total = (size_t)chains * T.dims;
  mass_elementwise_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(T, total, mode, in, out);
  return cudaGetLastError();
}

cudaError_t launch_rwmh_propose(int chains, int dims, const double* q, double* qp, const double* step_vec,
                                const double* step_chain, double stepsize, const double* z_in,
                                unsigned long long seed, long long chain_offset, long long kglob,
                                cudaStream_t s) {
  const long long work = (long long)chains * ((dims + 1) / 2);
  rwmh_propose_kernel<<<(unsigned)((work + 255) / 256), 256, 0, s>>>(chains, dims, q, qp, step_vec, step_chain,
                                                                    stepsize, z_in, seed, chain_offset, kglob);
  return cudaGetLastError();
}

cudaError_t launch_rwmh_decide(const RwmhDecide& D, cudaStream_t s) {
  rwmh_decide_kernel<<<D.chains, 128, 0, s>>>(D);
  return cudaGetLastError();
}

}  // namespace hmcb

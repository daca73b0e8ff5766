// rwmh.cuh -- Random Walk Metropolis-Hastings on the batched engine
// (hmclab.Samplers.RWMH._propose / _evaluate_acceptance, Samplers.py:1060-1086).
// The proposal and the decision are elementwise / per-chain kernels around the engine's
// batched misfit evaluation (hmcb_misfit), so every target the HMC path supports works here.
#pragma once
#include "common.cuh"

namespace hmcb {

#ifdef HMCB_FUSED_AUX_KERNELS   // kernels are compiled by launch_fused.cu only
// proposed = current + (stepsize * non_scalar_part) * normal       (Samplers.py:1064-1069)
__global__ void __launch_bounds__(256)
rwmh_propose_kernel(int chains, int dims, const double* __restrict__ q, double* __restrict__ qp,
                    const double* __restrict__ step_vec /* [dims] or null */,
                    const double* __restrict__ step_chain /* [chains] or null */, double stepsize,
                    const double* __restrict__ z_in /* [chains x dims] of this proposal or null */,
                    unsigned long long seed, long long chain_offset, long long kglob) {
  const int pairs = (dims + 1) / 2;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)chains * pairs) return;
  const int c = (int)(idx / pairs), m = (int)(idx % pairs);
  const size_t row = (size_t)c * dims;
  double z[2];
  if (z_in) {
    z[0] = z_in[row + 2 * m];
    z[1] = (2 * m + 1 < dims) ? z_in[row + 2 * m + 1] : 0.0;
  } else {
    normal_pair(seed, (uint32_t)(chain_offset + c), (uint32_t)kglob, (uint32_t)m, z[0], z[1]);
  }
  const double s = step_chain ? step_chain[c] : stepsize;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int j = 2 * m + h;
    if (j < dims) {
      const double sj = step_vec ? __dmul_rn(s, __ldg(step_vec + j)) : s;
      qp[row + j] = __dadd_rn(q[row + j], __dmul_rn(sj, z[h]));
    }
  }
}

#endif  // HMCB_FUSED_AUX_KERNELS

struct RwmhDecide {
  int chains, dims;
  double* q;            // [C x d] current models (updated on accept)
  const double* qp;     // [C x d] proposals
  double* x;            // [C] current misfits
  const double* x1;     // [C] proposed misfits
  const double* u_in;   // [C] acceptance uniforms of this proposal or null
  unsigned long long seed;
  long long chain_offset, kglob;
  double* sample_rows;  // [C x (d+1)] row block of this stored proposal or null
  unsigned char* out_accept;  // [C] slices or null
  double *out_h0, *out_h1;
  int* accepted_total;
  double* stepsize_chain;
  double* out_stepsize;
  AutotuneArgs tune;
};

#ifdef HMCB_FUSED_AUX_KERNELS
// exp(x - x') > u  (Samplers.py:1075-1086); one block per chain
__global__ void __launch_bounds__(128)
rwmh_decide_kernel(const RwmhDecide D) {
  __shared__ int acc_s;
  const int c = blockIdx.x;
  if (threadIdx.x == 0) {
    double u_step, u_acc;
    uniform_pair(D.seed, (uint32_t)(D.chain_offset + c), (uint32_t)D.kglob, u_step, u_acc);
    if (D.u_in) u_acc = D.u_in[c];
    const double x0 = D.x[c], x1 = D.x1[c];
    const bool acc = metropolis_accept(x0, x1, u_acc);
    if (D.out_stepsize) D.out_stepsize[c] = D.stepsize_chain ? D.stepsize_chain[c] : 0.0;
    if (D.tune.enabled) D.stepsize_chain[c] = autotune_stepsize(D.tune, D.stepsize_chain[c], x0, x1, D.kglob);
    if (acc) {
      D.x[c] = x1;
      if (D.accepted_total) D.accepted_total[c] += 1;
    }
    if (D.out_accept) D.out_accept[c] = acc ? 1 : 0;
    if (D.out_h0) D.out_h0[c] = x0;
    if (D.out_h1) D.out_h1[c] = x1;
    if (D.sample_rows) D.sample_rows[(size_t)c * (D.dims + 1) + D.dims] = acc ? x1 : x0;
    acc_s = acc ? 1 : 0;
  }
  __syncthreads();
  const bool acc = acc_s != 0;
  const size_t row = (size_t)c * D.dims;
  for (int j = threadIdx.x; j < D.dims; j += blockDim.x) {
    double v = D.q[row + j];
    if (acc) {
      v = D.qp[row + j];
      D.q[row + j] = v;
    }
    if (D.sample_rows) D.sample_rows[(size_t)c * (D.dims + 1) + j] = v;
  }
}

#endif  // HMCB_FUSED_AUX_KERNELS

}  // namespace hmcb

// racecheck_ring.cu -- the smallest producer/consumer ring of the kind the DMMA GEMM and the strip SpMM
// use, to check what compute-sanitizer's racecheck reports on it:
//   mode 0: the producer fills a stage with ONE bulk asynchronous copy (cp.async.bulk ... mbarrier::
//           complete_tx::bytes, the TMA engine) after arrive.expect_tx on the stage's "full" barrier;
//   mode 1: the producer warp fills the stage with plain st.shared and then arrives on the same barrier.
// In both modes the consumers wait on "full" (mbarrier.try_wait.parity), read the stage, and release it
// through an "empty" barrier the producer waits on before refilling.  The result is checked on the host.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -o racecheck_ring racecheck_ring.cu
// Run:   compute-sanitizer --tool racecheck ./racecheck_ring 0   (and 1)
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int STAGES = 2, STAGE_DOUBLES = 512, CONSUMERS = 4, TURNS = 64;

__device__ __forceinline__ unsigned saddr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@!p bra W;\n}\n" ::
                   "r"(saddr(bar)), "r"(parity) : "memory");
}

__global__ void ring(const double* __restrict__ src, double* __restrict__ out, int mode) {
  __shared__ __align__(128) double stage[STAGES][STAGE_DOUBLES];
  __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(saddr(&full_bar[s])), "r"(mode == 0 ? 1 : 32));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(saddr(&empty_bar[s])), "r"(CONSUMERS));
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  }
  __syncthreads();
  if (warp == CONSUMERS) {   // producer warp
    for (int t = 0; t < TURNS; ++t) {
      const int s = t % STAGES;
      if (t >= STAGES) mbar_wait(&empty_bar[s], (unsigned)((t / STAGES - 1) & 1));
      const double* g = src + (size_t)t * STAGE_DOUBLES;
      if (mode == 0) {
        if (lane == 0) {
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(saddr(&full_bar[s])),
                       "r"((unsigned)(STAGE_DOUBLES * 8)) : "memory");
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::
                           "r"(saddr(stage[s])), "l"(g), "r"((unsigned)(STAGE_DOUBLES * 8)), "r"(saddr(&full_bar[s]))
                       : "memory");
        }
      } else {
        for (int i = lane; i < STAGE_DOUBLES; i += 32) stage[s][i] = g[i];
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(saddr(&full_bar[s])) : "memory");
      }
    }
    return;
  }
  double acc = 0.0;
  for (int t = 0; t < TURNS; ++t) {
    const int s = t % STAGES;
    mbar_wait(&full_bar[s], (unsigned)((t / STAGES) & 1));
    for (int i = lane; i < STAGE_DOUBLES; i += 32) acc += stage[s][i] * (double)(warp + 1);
    __syncwarp();
    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(saddr(&empty_bar[s])) : "memory");
  }
  out[blockIdx.x * CONSUMERS * 32 + warp * 32 + lane] = acc;
}

int main(int argc, char** argv) {
  const int mode = argc > 1 ? atoi(argv[1]) : 0, blocks = 8;
  const size_t n = (size_t)TURNS * STAGE_DOUBLES;
  double *src, *out, *h = (double*)malloc(n * 8), *ho = (double*)malloc(blocks * CONSUMERS * 32 * 8);
  for (size_t i = 0; i < n; ++i) h[i] = (double)(i % 97) * 0.25;
  cudaMalloc(&src, n * 8); cudaMalloc(&out, blocks * CONSUMERS * 32 * 8);
  cudaMemcpy(src, h, n * 8, cudaMemcpyHostToDevice);
  ring<<<blocks, (CONSUMERS + 1) * 32>>>(src, out, mode);
  cudaError_t e = cudaDeviceSynchronize();
  cudaMemcpy(ho, out, blocks * CONSUMERS * 32 * 8, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int w = 0; w < CONSUMERS; ++w)
    for (int l = 0; l < 32; ++l) {
      double ref = 0.0;
      for (int t = 0; t < TURNS; ++t)
        for (int i = l; i < STAGE_DOUBLES; i += 32) ref += h[(size_t)t * STAGE_DOUBLES + i] * (double)(w + 1);
      for (int b = 0; b < blocks; ++b) bad += ho[b * CONSUMERS * 32 + w * 32 + l] != ref;
    }
  printf("mode %d (%s): %s, %d wrong sums, cuda: %s\n", mode, mode == 0 ? "bulk copy + complete_tx" : "st.shared + arrive",
         bad ? "FAIL" : "ok", bad, cudaGetErrorString(e));
  return bad != 0;
}

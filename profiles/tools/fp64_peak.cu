// fp64_peak.cu -- measures the two fp64 roofs this pool's B200 offers, for the roofline
// denominators that MEASURED_PEAKS.json does not hold:
//   * DFMA issue peak of the SIMT fp64 pipe (register-resident dependent-chain FMAs, enough
//     independent chains per thread and warps per SM to saturate the pipe),
//   * DMMA (mma.sync.m8n8k4.f64) peak of the tensor path, operands in registers.
// cuBLAS DGEMM is measured next to it by profiles/tools/measure_fp64_peak.py (torch.matmul).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_kernel(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6,
         x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

__global__ void dmma_kernel(double* out, int iters, double a, double b) {
  double c[8][2];
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = threadIdx.x + i;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c[j][0]), "+d"(c[j][1])
                   : "d"(a), "d"(b));
  }
  double s = 0;
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class K>
static double time_ms(K launch, int reps) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  launch();
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(e0);
    launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount, threads = 256, blocks = sms * 8, iters = 20000;
  double* out;
  cudaMalloc(&out, sizeof(double) * blocks * threads);
  const double ms_fma = time_ms([&] { dfma_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, 5);
  const double fma_flops = 2.0 * 8 * iters * (double)blocks * threads;
  const double ms_mma = time_ms([&] { dmma_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, 5);
  const double mma_flops = 2.0 * 8 * 8 * 4 * 8.0 * iters * (double)blocks * (threads / 32);
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"dfma_tflops\": %.3f, \"dfma_ginst_per_s\": %.1f, "
         "\"dmma_tflops\": %.3f, \"dfma_ms\": %.3f, \"dmma_ms\": %.3f}\n",
         prop.name, sms, fma_flops / ms_fma / 1e9, fma_flops / 2 / ms_fma / 1e6, mma_flops / ms_mma / 1e9,
         ms_fma, ms_mma);
  return 0;
}

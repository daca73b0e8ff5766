// debug_peak.cu -- fp64 roofline denominators measured in the run that reports against them
// (MEASURED_PEAKS.json holds HBM and bf16 numbers only): the DFMA issue peak of the SIMT fp64
// pipe and the DMMA (mma.sync.m8n8k4.f64) peak of the fp64 tensor path, operands in registers.
// bench.py calls hmcb_debug_fp64_peak next to a cuBLAS DGEMM burst and records the clocks.
#include "../../include/hmcb.h"

#include <cuda_runtime.h>

namespace {

__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters, double a, double b) {
  double x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = fma(x[i], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) dmma_peak_kernel(double* out, int iters, double a, double b) {
  double c[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c[j][0]), "+d"(c[j][1])
                   : "d"(a), "d"(b));
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// kind 2: one F2F.F64.F32 conversion + one DADD per step (how fast does the fp32 -> fp64
// conversion issue next to fp64 arithmetic? decides the nonzero format of the SpMM tables)
__global__ void __launch_bounds__(256) f2f_peak_kernel(double* out, int iters, double a, double b) {
  float f[8];
  double x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { f[i] = (float)(threadIdx.x + i); x[i] = 0.0; }
  const float fa = (float)a, fb = (float)b;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      f[i] = fmaf(f[i], fa, fb);
      x[i] += (double)f[i];
    }
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace

extern "C" int hmcb_debug_fp64_peak(int device, int kind, int iters, int launches, double* best_ms,
                                    double* flops_per_launch) {
  if (!best_ms || !flops_per_launch || iters <= 0 || launches <= 0 || kind < 0 || kind > 2) return -1;
  if (cudaSetDevice(device) != cudaSuccess) return -1;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return -1;
  const int threads = 256, blocks = prop.multiProcessorCount * 8;
  double* out = nullptr;
  if (cudaMalloc(&out, sizeof(double) * (size_t)blocks * threads) != cudaSuccess) return -1;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int r = 0; r <= launches; ++r) {   // launch 0 is the warm-up
    cudaEventRecord(e0);
    if (kind == 0) dfma_peak_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
    else if (kind == 1) dmma_peak_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
    else f2f_peak_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(out); return -1; }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (r > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(out);
  *best_ms = best;
  // kind 2 reports conversions (one per step and thread) instead of flops
  *flops_per_launch = kind == 1 ? 2.0 * 8 * 8 * 4 * 8.0 * iters * (double)blocks * (threads / 32)
                                : (kind == 0 ? 2.0 : 1.0) * 8 * iters * (double)blocks * threads;
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

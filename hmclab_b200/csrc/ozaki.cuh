// ozaki.cuh -- the dense LinearMatrix products on Blackwell's 5th-generation tensor cores.
//
// tcgen05.mma has no fp64 kind, but it multiplies int8 exactly into int32 accumulators in tensor
// memory.  An fp64 product  Y = A B  is therefore split (Ozaki scheme):
//
//     A[i][k] = 2^ea[i] * sum_s A_s[i][k] 2^(-7 (s+1)),   B[k][j] = 2^eb[j] * sum_t B_t[k][j] 2^(-7 (t+1))
//
// with int8 slices A_s, B_t in [-127, 127] (7 bits + sign, truncation toward zero; row scales ea for
// the model matrix, per-chain scales eb for the chain batch).  Every slice product  A_s B_t  is an
// exact int8 GEMM; products of equal order o = s + t share one int32 accumulator (no overflow as long
// as K * pairs * 127^2 < 2^31, checked by the host), and
//
//     Y[i][j] = 2^(ea[i] + eb[j]) * sum_o 2^(-7 (o + 2)) C_o[i][j]
//
// is recombined in fp64.  G is float32 by the reference's own rounding (LinearMatrix.py:148-153): 5
// slices (35 bits below the row maximum) hold all but the mantissa tails of its tiniest entries; the
// chain batch gets 7 slices (49 bits below the chain maximum); orders 0..6 are kept (25 slice pairs).
// What is dropped is below 2^-48 of |A|_row-max |B|_chain-max per term: a relative error of a few 1e-14
// on a gradient, the same level as the summation-order differences between BLAS and the DMMA GEMM,
// and far inside the 1e-10 parity bar.
//
// This file: (1) i8_gemm_orders_kernel -- TMA-fed (128-byte swizzle), one elected thread issuing
// tcgen05.mma.kind::i8 (SASS UTCIMMA) into a TMEM accumulator, epilogue warps reading it back with
// tcgen05.ld (SASS LDTM); blockIdx.z = order, the K loop runs over the slice pairs of that order;
// (2) the slicing kernels; (3) the fp64 recombination with the fused HMC epilogues.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "gemm.cuh"
#include "staged.cuh"

namespace hmcb {

constexpr int OZ_BM = 128, OZ_BN = 256, OZ_BK = 128;   // CTA tile; BK int8 = one 128-byte swizzle row
constexpr int OZ_STAGES = 4;
constexpr int OZ_A_BYTES = OZ_BM * OZ_BK, OZ_B_BYTES = OZ_BN * OZ_BK;
constexpr int OZ_STAGE_BYTES = OZ_A_BYTES + OZ_B_BYTES;                 // 48 KB
constexpr size_t OZ_SMEM_BYTES = (size_t)OZ_STAGES * OZ_STAGE_BYTES + 1024;   // + alignment slack
constexpr int OZ_THREADS = 256;    // warp 0: TMA, warp 1: MMA, warp 2: TMEM allocation, warps 4-7: epilogue
constexpr int OZ_BITS = 7;         // magnitude bits per slice: slices in [-127, 127]

// shared-memory matrix descriptor of a K-major operand tile laid out by a 128-byte-swizzled TMA box:
// rows of 128 bytes, 8-row swizzle atoms 1024 bytes apart (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t umma_desc_k_major_sw128(unsigned smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4)        // start address
         | ((uint64_t)1 << 16)                           // leading byte offset (unused for swizzled K-major)
         | ((uint64_t)(1024 >> 4) << 32)                 // stride byte offset: next 8-row atom
         | ((uint64_t)1 << 46)                           // descriptor version (sm_100)
         | ((uint64_t)2 << 61);                          // SWIZZLE_128B
}

// instruction descriptor (cute::UMMA::InstrDescriptor): D = s32, A = B = signed int8, both K-major
__host__ __device__ constexpr unsigned umma_idesc_i8(int M, int N) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(N >> 3) << 17) | ((unsigned)(M >> 4) << 24);
}

__device__ __forceinline__ void tma_load_3d_raw(unsigned dst, const CUtensorMap* map, int c0, int c1, int c2,
                                                unsigned bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n" ::
          "r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}

// C[o][m][n] = sum over slice pairs (s, t = o - s) of A_s[m][:] . B_t[n][:]   (int8 x int8 -> int32, exact)
//   mapA: {K, M, SA} int8, box {128, 128, 1};  mapB: {K, N, SB} int8, box {128, 256, 1}; 128-byte swizzle
__global__ void __launch_bounds__(OZ_THREADS, 1)
i8_gemm_orders_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                      int kblocks, int SA, int SB, int* __restrict__ C, long long plane_stride, int ldc) {
  extern __shared__ __align__(1024) unsigned char oz_smem[];
  __shared__ uint64_t full_bar[OZ_STAGES], empty_bar[OZ_STAGES], tmem_full_bar;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int order = blockIdx.z, m0 = blockIdx.y * OZ_BM, n0 = blockIdx.x * OZ_BN;
  const int s_lo = max(0, order - (SB - 1)), s_hi = min(SA - 1, order);
  const int total = (s_hi - s_lo + 1) * kblocks;          // k-blocks of this order
  const unsigned smem0 = ((unsigned)__cvta_generic_to_shared(oz_smem) + 1023u) & ~1023u;

  if (threadIdx.x == 0) {
    for (int s = 0; s < OZ_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&tmem_full_bar, 1);
    fence_async_proxy();
  }
  if (warp == 2) {   // one warp allocates the accumulator columns of tensor memory
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::
                     "r"((unsigned)__cvta_generic_to_shared(&tmem_base_s)), "r"((unsigned)OZ_BN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = tmem_base_s;

  if (warp == 0) {
    if (lane == 0) {   // ---- TMA producer
      for (int it = 0; it < total; ++it) {
        const int stage = it % OZ_STAGES;
        if (it >= OZ_STAGES) mbar_wait(&empty_bar[stage], (unsigned)((it / OZ_STAGES - 1) & 1));
        const int s = s_lo + it / kblocks, kb = it % kblocks, t = order - s;
        const unsigned bar = (unsigned)__cvta_generic_to_shared(&full_bar[stage]);
        const unsigned dst = smem0 + (unsigned)stage * OZ_STAGE_BYTES;
        mbar_expect_tx(&full_bar[stage], (unsigned)OZ_STAGE_BYTES);
        tma_load_3d_raw(dst, &mapA, kb * OZ_BK, m0, s, bar);
        tma_load_3d_raw(dst + OZ_A_BYTES, &mapB, kb * OZ_BK, n0, t, bar);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {   // ---- MMA issuer: one thread drives the tensor core
      constexpr unsigned idesc = umma_idesc_i8(OZ_BM, OZ_BN);
      for (int it = 0; it < total; ++it) {
        const int stage = it % OZ_STAGES;
        mbar_wait(&full_bar[stage], (unsigned)((it / OZ_STAGES) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        const unsigned a_addr = smem0 + (unsigned)stage * OZ_STAGE_BYTES;
        const uint64_t a_desc = umma_desc_k_major_sw128(a_addr), b_desc = umma_desc_k_major_sw128(a_addr + OZ_A_BYTES);
#pragma unroll
        for (int k = 0; k < OZ_BK / 32; ++k) {     // K = 32 int8 per instruction: 32 bytes further along the row
          const unsigned accumulate = (it > 0 || k > 0) ? 1u : 0u;
          asm volatile(
              "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
              "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n}\n" ::
                  "r"(tmem), "l"(a_desc + (uint64_t)(2 * k)), "l"(b_desc + (uint64_t)(2 * k)), "r"(idesc),
              "r"(accumulate), "r"(0u) : "memory");
        }
        // commit: the barrier is arrived on when the MMAs issued so far have finished reading the stage
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::
                         "r"((unsigned)__cvta_generic_to_shared(&empty_bar[stage])) : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::
                       "r"((unsigned)__cvta_generic_to_shared(&tmem_full_bar)) : "memory");
    }
  } else if (warp >= 4) {
    // ---- epilogue: warp w reads TMEM lanes 32 (w % 4) .. +31 = rows m0 + 32 (w - 4) + lane
    mbar_wait(&tmem_full_bar, 0u);
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const int row = m0 + (warp - 4) * 32 + lane;
    int* crow = C + (size_t)order * plane_stride + (size_t)row * ldc + n0;
    const uint32_t taddr = tmem + ((uint32_t)((warp - 4) * 32) << 16);
#pragma unroll 1
    for (int c = 0; c < OZ_BN; c += 32) {
      uint32_t r[32];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
            "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
            "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
            "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
          : "r"(taddr + (uint32_t)c));
      asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
      if (n0 + c < ldc) {
#pragma unroll
        for (int v = 0; v < 8; ++v)
          *reinterpret_cast<int4*>(crow + c + 4 * v) =
              make_int4((int)r[4 * v], (int)r[4 * v + 1], (int)r[4 * v + 2], (int)r[4 * v + 3]);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 2)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"((unsigned)OZ_BN));
}


// ---- slicing and recombination --------------------------------------------------------------------

constexpr int OZ_SA = 5, OZ_SB = 7, OZ_ORDERS = 7;   // = OZ_SLICES_A / _B / OZ_NUM_ORDERS of launch.cuh

// exponent eb with max |x| < 2^eb from the bits of max |x| (0 for an all-zero chain); INT_MIN marks a
// chain that holds an inf / NaN (or a value next to the overflow threshold): its products are NaN
__device__ __forceinline__ int oz_exponent(unsigned long long maxbits) {
  const int e = (int)(maxbits >> 52);
  if (e >= 2046) return INT_MIN;
  if (e == 0) return 0;          // zero (or subnormal: treated as zero)
  return e - 1022;
}
__device__ __forceinline__ double oz_pow2(int e) {   // 2^e for e in [-1022, 1023]
  return __longlong_as_double((long long)(e + 1023) << 52);
}

// per-chain max |X[r][c]| over the rows of a plane [rows x ld] (as the bit pattern: NaN > inf > finite)
__global__ void __launch_bounds__(128)
oz_colmax_kernel(const double* __restrict__ X, int rows, int ld, unsigned long long* __restrict__ maxbits) {
  const int c = blockIdx.x * 128 + threadIdx.x;
  if (c >= ld) return;
  const int r1 = min(rows, ((int)blockIdx.y + 1) * 64);
  unsigned long long m = 0ull;
  for (int r = blockIdx.y * 64; r < r1; ++r) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(fabs(X[(size_t)r * ld + c]));
    m = b > m ? b : m;
  }
  if (m) atomicMax(maxbits + c, m);
}

// X [K x ld] (chains contiguous) -> OZ_SB int8 slices, chain-major [t][ld][K] (K contiguous): the B
// operand of the tensor-core product.  Block = 128 k x 32 chains, transposed through shared memory.
__global__ void __launch_bounds__(256)
oz_slice_chains_kernel(const double* __restrict__ X, int K, int ld, const unsigned long long* __restrict__ maxbits,
                       signed char* __restrict__ out) {
  __shared__ __align__(16) signed char sl[OZ_SB][32][132];
  const int k0 = blockIdx.y * 128, c0 = blockIdx.x * 32;
  const int tc = threadIdx.x & 31, tr = threadIdx.x >> 5;
  const int eb = oz_exponent(maxbits[c0 + tc]);
  const double scale = (eb == INT_MIN) ? 0.0 : oz_pow2(-eb);   // |x| * scale < 1
#pragma unroll 4
  for (int r = tr; r < 128; r += 8) {
    double x = (k0 + r < K) ? X[(size_t)(k0 + r) * ld + c0 + tc] * scale : 0.0;
#pragma unroll
    for (int t = 0; t < OZ_SB; ++t) {
      x *= 128.0;                       // exact
      const double v = trunc(x);        // |v| <= 127
      x -= v;                           // exact, same sign, |x| < 1
      sl[t][tc][r] = (signed char)(int)v;
    }
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int row = warp; row < OZ_SB * 32; row += 8) {
    const int t = row >> 5, ch = row & 31;
    const unsigned v = *reinterpret_cast<const unsigned*>(&sl[t][ch][lane * 4]);
    *reinterpret_cast<unsigned*>(out + ((size_t)t * ld + c0 + ch) * K + k0 + lane * 4) = v;
  }
}

// Y[i][c] = 2^(ea[i] + eb[c]) sum_o 2^(-7 (o + 2)) C_o[i][c]  -> epilogue functor; optionally the per-chain
// max |R| of what a ResidualEpi stored (the scale of the next slicing pass)
template <class Epi, bool TRACK_MAX>
__global__ void __launch_bounds__(128)
oz_combine_kernel(const int* __restrict__ C, long long plane_stride, int rows, int ld, const int* __restrict__ ea,
                  const unsigned long long* __restrict__ maxbits_in, Epi epi,
                  unsigned long long* __restrict__ maxbits_out) {
  const int c = blockIdx.x * 128 + threadIdx.x;
  if (c >= ld) return;
  const int eb = oz_exponent(maxbits_in[c]);
  unsigned long long m = 0ull;
  const int r1 = min(rows, ((int)blockIdx.y + 1) * 16);
  for (int i = blockIdx.y * 16; i < r1; ++i) {
    const int* p = C + (size_t)i * ld + c;
    double acc = 0.0;
#pragma unroll
    for (int o = OZ_ORDERS - 1; o >= 0; --o) acc = fma(acc, 0.0078125, (double)p[(size_t)o * plane_stride]);
    double y;
    if (eb == INT_MIN) y = CUDART_NAN;
    else {
      // two exact power-of-two factors (their product can leave the normal range although y does not)
      const int e = ea[i] + eb - 2 * OZ_BITS, h = e / 2;
      y = acc * oz_pow2(h) * oz_pow2(e - h);
    }
    if constexpr (TRACK_MAX) {
      if (i < epi.N && c < epi.C) {
        const double r = __ddiv_rn(__dsub_rn(y, __ldg(epi.dvec + i)), __ldg(epi.var + i));
        epi.R[(size_t)i * epi.ld + c] = r;
        const unsigned long long b = (unsigned long long)__double_as_longlong(fabs(r));
        m = b > m ? b : m;
      }
    } else {
      epi.row(i, c, y);
    }
  }
  if constexpr (TRACK_MAX) {
    if (m) atomicMax(maxbits_out + c, m);
  }
}

}  // namespace hmcb

// spmm_strip.cuh -- CSR SpMM over the chain batch with the gathered operand staged in shared
// memory (replaces the SciPy csr_matvec / MKL mkl_cspblas_dcsrgemv calls of
// LinearMatrix.py:389-426 and Helpers/InterfaceMKL.py:87-121 for a whole batch of chains).
//
//   Y[i][c] = sum_k val[k] * B[col[k]][c]        B: [cols x chains], chains contiguous
//
// The L2-gather kernel (csr_spmm_kernel, gemm.cuh) moves nnz x chains x 8 B from L2 into the SMs
// and sits on the L2 fabric limit.  Here a block owns RB = WARPS x RW rows and one slab of
// S = 32 x CPL chains and walks the columns strip by strip:
//   * hmcb_finalize deals the columns round-robin into T strips of at most KB columns (column c
//     belongs to strip c mod T: every row spreads its nonzeros evenly over the strips, whatever
//     its geometry, so the warps of a block stay balanced) and stores, per (chunk, strip), the nonzeros of the chunk's rows that
//     fall into the strip as one flat stream per consumer warp ({value, local column offset,
//     row within the warp} as 16 bytes, closed by a sentinel);
//   * a producer warp stages strip after strip into a ring of stages: the B rows of the strip
//     with ONE tensor-map TMA load (B viewed as the 3-D tensor {chains, T, rows / T}; SASS
//     UTMALDG), the nonzero streams with one bulk copy (UBLKCP), both completing on the
//     stage's mbarrier; consumer warps release a stage through an
//     "empty" mbarrier, there is no block-wide barrier in the loop;
//   * a consumer warp keeps the accumulators of its RW rows x CPL chains per lane in registers
//     and walks its stream with the next nonzero always in flight; every load of the inner loop
//     is a shared-memory load: one 16-byte broadcast for the nonzero, one conflict-free
//     S x 8 B row segment for the gather (CPL = 2: one gather feeds two FMAs).
// L2 -> SM traffic drops from nnz x chains x 8 B to (rows / RB) x cols x chains x 8 B.
#pragma once
#include "common.cuh"
#include "gemm.cuh"
#include "spmm_types.cuh"

namespace hmcb {

constexpr int SPMM_PRODUCERS = 1;   // one warp drives the TMA ring

// 3-D tiled TMA load (SASS UTMALDG) completing on an mbarrier: box {S chains, 1 strip, kb rows}
__device__ __forceinline__ void tma_load_3d(void* smem, const CUtensorMap* map, int c0, int c1, int c2,
                                            uint64_t* bar) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem);
  const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n" ::
          "r"(d), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(b) : "memory");
}

// shared-memory loads through 32-bit shared-window addresses: with generic pointers the compiler
// re-derives the window base (S2UR / ULEA / UMOV) inside every unrolled row of the inner loop
__device__ __forceinline__ unsigned lds_u32(unsigned a) {
  unsigned v;
  asm volatile("ld.shared.u32 %0, [%1];\n" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ uint2 lds_u32x2(unsigned a) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];\n" : "=r"(v.x), "=r"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ int4 lds_s32x4(unsigned a) {
  int4 v;
  asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];\n" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ double lds_f64(unsigned a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];\n" : "=d"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ double2 lds_f64x2(unsigned a) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"(v.x), "=d"(v.y) : "r"(a));
  return v;
}

template <class Epilogue, int WARPS, int RW, int CPL, bool COMPACT>
__global__ void __launch_bounds__((WARPS + SPMM_PRODUCERS) * 32, 1)
csr_spmm_strip_kernel(const StripDev M, const __grid_constant__ CUtensorMap bmap, Epilogue epi) {
  constexpr int S = 32 * CPL, RB = WARPS * RW;
  extern __shared__ __align__(128) unsigned char strip_smem[];
  __shared__ uint64_t full_bar[SPMM_MAX_STAGES], empty_bar[SPMM_MAX_STAGES];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int chunk = blockIdx.x, slab0 = blockIdx.y * S;
  const int s_begin = __ldg(M.strip_ptr + chunk);
  const int nst = __ldg(M.strip_ptr + chunk + 1) - s_begin;
  const int nstages = M.stages;

  if (tid == 0) {
    for (int st = 0; st < nstages; ++st) {
      mbar_init(&full_bar[st], 1);   // the producer's expect_tx arrival; the copies complete the bytes
      mbar_init(&empty_bar[st], WARPS);
    }
    fence_async_proxy();
  }
  __syncthreads();

  if (warp >= WARPS) {   // ---- producer warp: one lane drives the ring
    // B is seen as a 3-D tensor {chains, T strips, rows of a strip} (row c of B = strip c mod T,
    // local row c / T): one tensor-map load lands the whole strip; rows past the end of B are
    // zero filled by the engine.  The nonzero streams of the (chunk, strip) group: one bulk copy.
    if (lane != 0) return;
    int stage = 0;
    unsigned phase = 1;  // parity of the previous use of the stage (first pass: nothing to wait for)
    for (int t = 0; t < nst; ++t) {
      if (t >= nstages) mbar_wait(&empty_bar[stage], phase);
      const int4 raw = __ldg(reinterpret_cast<const int4*>(M.strips + s_begin + t));
      unsigned char* base = strip_smem + (size_t)stage * M.stage_bytes;
      mbar_expect_tx(&full_bar[stage], (unsigned)M.kb_box * (S * 8u) + (unsigned)raw.w * 16u);
      tma_load_3d(base, &bmap, slab0, raw.x, 0, &full_bar[stage]);
      bulk_copy_g2s(base + M.b_bytes, reinterpret_cast<const int4*>(M.ent) + raw.z, (unsigned)raw.w * 16u,
                    &full_bar[stage]);
      if (++stage == nstages) { stage = 0; phase ^= 1u; }
    }
    return;
  }

  // ---- consumer warps: rows [row0, row0 + RW), chains slab0 + lane * CPL + {0 .. CPL-1}
  double acc[RW][CPL];
#pragma unroll
  for (int r = 0; r < RW; ++r)
#pragma unroll
    for (int h = 0; h < CPL; ++h) acc[r][h] = 0.0;

  {
    const unsigned smem0 = (unsigned)__cvta_generic_to_shared(strip_smem);
    int stage = 0;
    unsigned phase = 0;
    for (int t = 0; t < nst; ++t) {
      mbar_wait(&full_bar[stage], phase);
      const unsigned base = smem0 + (unsigned)stage * (unsigned)M.stage_bytes;
      const unsigned bs = base + lane * (CPL * 8);
      unsigned es = base + (unsigned)M.b_bytes;
      if constexpr (COMPACT) {
        // 8-byte nonzeros {fp32 value, row << 24 | byte offset}: one LDS.64 broadcast each
        es += 8u * lds_u32(es + 4u * warp);   // header: first slot of every warp
        uint2 e = lds_u32x2(es);
#pragma unroll
        for (int r = 0; r < RW; ++r) {
#pragma unroll 1
          // rows come in ascending order: "row == r" is one unsigned compare, and the row bits
          // leave the address by a compile-time constant
          while (e.y < ((unsigned)(r + 1) << 24)) {
            const double v = (double)__uint_as_float(e.x);
            const unsigned at = bs + e.y - ((unsigned)r << 24);
            es += 8u;
            e = lds_u32x2(es);
            if constexpr (CPL == 1) {
              acc[r][0] = fma(v, lds_f64(at), acc[r][0]);
            } else {
              const double2 b = lds_f64x2(at);
              acc[r][0] = fma(v, b.x, acc[r][0]);
              acc[r][1] = fma(v, b.y, acc[r][1]);
            }
          }
        }
      } else {
        es += 16u * lds_u32(es + 4u * warp);
        int4 e = lds_s32x4(es);
#pragma unroll
        for (int r = 0; r < RW; ++r) {
#pragma unroll 1
          while (e.w == r) {
            const double v = __hiloint2double(e.y, e.x);
            const unsigned at = bs + (unsigned)e.z;
            es += 16u;
            e = lds_s32x4(es);
            if constexpr (CPL == 1) {
              acc[r][0] = fma(v, lds_f64(at), acc[r][0]);
            } else {
              const double2 b = lds_f64x2(at);
              acc[r][0] = fma(v, b.x, acc[r][0]);
              acc[r][1] = fma(v, b.y, acc[r][1]);
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[stage]);
      if (++stage == nstages) { stage = 0; phase ^= 1u; }
    }
  }

  // ---- epilogue: one functor copy per chain of the lane (MisfitEpi carries a running sum)
  const int row0 = chunk * RB + warp * RW;
  const int c0 = slab0 + lane * CPL;
  Epilogue ep[CPL];
#pragma unroll
  for (int h = 0; h < CPL; ++h) {
    ep[h] = epi;
    ep[h].tile_begin(row0, slab0);
  }
#pragma unroll
  for (int r = 0; r < RW; ++r) {
    const int i = row0 + r;
    if (i < M.rows) {
#pragma unroll
      for (int h = 0; h < CPL; ++h) ep[h].row(i, c0 + h, acc[r][h]);
    }
  }
#pragma unroll
  for (int h = 0; h < CPL; ++h) ep[h].chunk_end(chunk * WARPS + warp, c0 + h);
}

}  // namespace hmcb

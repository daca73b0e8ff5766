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


// ---- row-blocked tensor-core variant -----------------------------------------------------------
// The kernel above gathers one 8-byte B element from shared memory per FMA: the shared-memory data
// pipe (5 wavefronts per nonzero and 64 chains) and the issue slots (18 instructions per nonzero)
// bind, the fp64 pipe idles at 17 %.  Broadcasting the matrix values costs a wavefront per 8 bytes
// as well (a SIMT kernel with 8-row register blocking measured slower than the plain one), so the
// operand reuse has to happen inside an instruction: the fp64 tensor core.
//
// hmcb_finalize regroups the rows into groups of 8 rows with similar column sets (cluster_rows in
// hmcb.cu; rays that run side by side cross the same cells).  A group is one DMMA M-tile; a k-tile
// is 4 columns of the group with the nonzeros of its 8 x 4 A fragment: one per-lane load delivers
// the lane's matrix value, one 16-byte per-lane load the B fragments of two N-tiles (4 gathered
// rows x 16 chains), two mma.sync.m8n8k4.f64 do 512 FMAs.
//   * slab = 16 NB chains, staged as 128-byte-swizzled tensor-map TMA boxes {16 chains, 1 strip,
//     box_rows rows}; a box feeds two N-tiles, the even and the odd chains: lane (k, n) reads chains
//     2n, 2n + 1 of row k with one 16-byte load.  A quarter warp (k = 0..3, two n) reads the same
//     32-byte column of 4 rows; the swizzle moves row r to bank group (r >> 1) & 3 of that quarter, and
//     the host picks the 4 columns of a k-tile from 4 different classes whenever it can: no conflict;
//   * a consumer warp owns GW groups (8 GW rows): 2 NB GW accumulator pairs per lane, and with the
//     narrow slab a block covers WARPS x GW x 8 rows -- 4x the rows of the plain kernel, so 4x less
//     L2 -> shared-memory staging of B per useful flop;
//   * the grid interleaves the slabs of 128 chains per chunk so that a chunk's tables are re-read from L2.
constexpr int SPMM_SLAB_CHAINS_MINOR = 128;   // chains covered by the interleaved slabs of one chunk

__device__ __forceinline__ float lds_f32(unsigned a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];\n" : "=f"(v) : "r"(a));
  return v;
}

template <class Epilogue, int WARPS, int GW, int NB, bool COMPACT>
__global__ void __launch_bounds__((WARPS + SPMM_PRODUCERS) * 32, 1)
csr_spmm_block_kernel(const StripDev M, const __grid_constant__ CUtensorMap bmap, Epilogue epi) {
  constexpr int S = 16 * NB, NT = 2 * NB, R = SPMM_BLOCK_R, GPC = WARPS * GW;
  constexpr unsigned HDR = (unsigned)(((GPC + 1) * 4 + 15) / 16 * 16);   // header bytes
  constexpr unsigned REC = 32u, VSZ = COMPACT ? 4u : 8u;                  // record header / value bytes
  extern __shared__ __align__(1024) unsigned char strip_smem[];
  __shared__ uint64_t full_bar[SPMM_MAX_STAGES], empty_bar[SPMM_MAX_STAGES];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int SL = SPMM_SLAB_CHAINS_MINOR / S;   // slabs interleaved per chunk (ld is a multiple of 128)
  const int chunk = blockIdx.x / SL;
  const int slab0 = (blockIdx.y * SL + blockIdx.x % SL) * S;
  const int s_begin = __ldg(M.strip_ptr + chunk);
  const int nst = __ldg(M.strip_ptr + chunk + 1) - s_begin;
  const int nstages = M.stages;
  const unsigned smem0 = ((unsigned)__cvta_generic_to_shared(strip_smem) + 1023u) & ~1023u;

  if (tid == 0) {
    for (int st = 0; st < nstages; ++st) {
      mbar_init(&full_bar[st], 1);
      mbar_init(&empty_bar[st], WARPS);
    }
    fence_async_proxy();
  }
  __syncthreads();

  if (warp >= WARPS) {   // ---- producer warp: one lane drives the ring
    if (lane != 0) return;
    int stage = 0;
    unsigned phase = 1;
    const unsigned box_bytes = (unsigned)M.box_rows * 128u;
    for (int t = 0; t < nst; ++t) {
      if (t >= nstages) mbar_wait(&empty_bar[stage], phase);
      const int4 raw = __ldg(reinterpret_cast<const int4*>(M.strips + s_begin + t));
      const unsigned base = smem0 + (unsigned)stage * (unsigned)M.stage_bytes;
      const unsigned bar = (unsigned)__cvta_generic_to_shared(&full_bar[stage]);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar),
                   "r"((unsigned)M.b_bytes + (unsigned)raw.w * 16u) : "memory");
      for (int rb = 0; rb < M.row_boxes; ++rb)
#pragma unroll
        for (int cb = 0; cb < NB; ++cb)
          asm volatile(
              "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n" ::
                  "r"(base + (unsigned)(rb * NB + cb) * box_bytes), "l"(&bmap), "r"(slab0 + 16 * cb), "r"(raw.x),
                  "r"(rb * M.box_rows), "r"(bar) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::
                       "r"(base + (unsigned)M.b_bytes), "l"(reinterpret_cast<const int4*>(M.ent) + raw.z),
                   "r"((unsigned)raw.w * 16u), "r"(bar) : "memory");
      if (++stage == nstages) { stage = 0; phase ^= 1u; }
    }
    return;
  }

  // ---- consumer warps: groups (chunk * WARPS + warp) * GW + g.  Lane holds rows lane / 4 of its groups
  // and, per 16-chain box, chains 4 (lane % 4) + {0, 2} (even N-tile) and + {1, 3} (odd N-tile)
  double acc[GW][NT][2];
#pragma unroll
  for (int g = 0; g < GW; ++g)
#pragma unroll
    for (int j = 0; j < NT; ++j) acc[g][j][0] = acc[g][j][1] = 0.0;

  {
    const unsigned rowbit = 1u << (lane >> 2), below = rowbit - 1u;
    const unsigned lconst = (unsigned)(lane >> 2) << 4;   // the lane's 16-byte granule (chains 2n, 2n + 1)
    const unsigned box_bytes = (unsigned)M.box_rows * 128u;
    int stage = 0;
    unsigned phase = 0;
    for (int t = 0; t < nst; ++t) {
      mbar_wait(&full_bar[stage], phase);
      const unsigned base = smem0 + (unsigned)stage * (unsigned)M.stage_bytes;
      const unsigned es = base + (unsigned)M.b_bytes;
#pragma unroll
      for (int g = 0; g < GW; ++g) {
        const unsigned mine = lds_u32(es + 4u * (warp * GW + g));
        int n = (int)(mine & 1023u);
        unsigned rp = es + HDR + 4u * (mine >> 10) + 8u * (lane & 3);   // the lane's column word of the record
#pragma unroll 2
        for (; n > 0; --n) {
          const uint2 h = lds_u32x2(rp);
          const unsigned colmask = h.x >> 13;      // bits 0-7: rows with a value in the lane's column
          const unsigned vp = rp - 8u * (lane & 3) + REC + VSZ * ((h.x >> 21) + (unsigned)__popc(colmask & below));
          double a = 0.0;
          if (colmask & rowbit) {
            if constexpr (COMPACT) a = (double)lds_f32(vp);
            else a = lds_f64(vp);
          }
          // the host stores row * 128 + (row & 7) * 16: XOR with the lane's granule = its swizzled address
          const unsigned x = base + (((h.x & 0x1FFFu) << 4) ^ lconst);
          rp += REC + ((VSZ * h.y + 7u) & ~7u);   // records are 8-byte aligned
#pragma unroll
          for (int cb = 0; cb < NB; ++cb) {
            const double2 b = lds_f64x2(x + cb * box_bytes);
            dmma_8x8x4(acc[g][2 * cb][0], acc[g][2 * cb][1], a, b.x);
            dmma_8x8x4(acc[g][2 * cb + 1][0], acc[g][2 * cb + 1][1], a, b.y);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[stage]);
      if (++stage == nstages) { stage = 0; phase ^= 1u; }
    }
  }

  // ---- epilogue: block row -> original row through the permutation; accumulator (N-tile j, h) of box
  // cb = j / 2 is chain 16 cb + 4 (lane % 4) + 2 h + j % 2
  const int brow0 = (chunk * WARPS + warp) * GW * R + (lane >> 2);
  const int c0 = slab0 + 4 * (lane & 3);
  if constexpr (Epilogue::kPerChainSum) {
    // per-chain sums over the warp's rows: the rows of a chain sit in the 8 lanes with equal lane % 4
    int rows[GW];
#pragma unroll
    for (int g = 0; g < GW; ++g) rows[g] = __ldg(M.perm + brow0 + g * R);
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c = c0 + 16 * (j >> 1) + 2 * h + (j & 1);
        double v = 0.0;
#pragma unroll
        for (int g = 0; g < GW; ++g)
          if (rows[g] >= 0) v = __dadd_rn(v, epi.term(rows[g], c, acc[g][j][h]));
        v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, 4));
        v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, 8));
        v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, 16));
        if (lane < 4 && c < epi.ld) epi.part[(size_t)(chunk * WARPS + warp) * epi.ld + c] = v;
      }
  } else {
#pragma unroll
    for (int g = 0; g < GW; ++g) {
      const int i = __ldg(M.perm + brow0 + g * R);
      if (i < 0) continue;
#pragma unroll
      for (int cb = 0; cb < NB; ++cb) {
        epi.row(i, c0 + 16 * cb, acc[g][2 * cb][0]);
        epi.row(i, c0 + 16 * cb + 1, acc[g][2 * cb + 1][0]);
        epi.row(i, c0 + 16 * cb + 2, acc[g][2 * cb][1]);
        epi.row(i, c0 + 16 * cb + 3, acc[g][2 * cb + 1][1]);
      }
    }
  }
}

}  // namespace hmcb

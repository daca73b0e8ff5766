// ozaki.cuh -- the dense LinearMatrix products on Blackwell's 5th-generation tensor cores.
//
// tcgen05.mma has no fp64 kind, but it multiplies int8 exactly into int32 accumulators in tensor
// memory.  An fp64 product  Y = A B  is therefore split (Ozaki scheme) into radix-256 digits:
//
//     A[i][k] = 2^ea[i] * sum_s A_s[i][k] 256^-(s+1),     B[k][j] = 2^eb[j] * sum_t B_t[k][j] 256^-(t+1)
//
// with BALANCED int8 digits A_s, B_t in [-128, 127] (the value scaled to |y| < 0.494, rounded to an
// integer of 8 S bits and decomposed from the least significant digit upward with carries: 8 bits
// per slice where sign + magnitude truncation would give 7; row scales ea for the model matrix,
// per-chain scales eb for the chain batch).  Every slice product  A_s B_t  is an exact int8 GEMM;
// products of equal order o = s + t share one int32 accumulator (no overflow as long as
// K * pairs * 128^2 < 2^31, checked by the host), and
//
//     Y[i][j] = 2^(ea[i] + eb[j]) * sum_o 256^-(o + 2) C_o[i][j]
//
// is recombined in fp64.  The chain batch gets 6 digits (47 bits below the chain maximum), the model
// matrix as many as it needs for a row-wise representation error below 2^-45 of the row's absolute
// sum (5 for a matrix that is float32 by the reference's own rounding, LinearMatrix.py:148-153), and
// orders 0..5 are kept: 20 slice pairs.  What is dropped is of order 2^-48 |A|_row-max |B|_chain-max
// per term: a relative error near 1e-13 on a gradient, the class of the summation-order differences
// between BLAS and the DMMA GEMM at these sizes, far inside the 1e-10 parity bar
// (HMCB_OZAKI_ORDERS=7 adds a digit and an order: 2^-8 of that).
//
// This file: (1) i8_gemm_groups_kernel -- the slice products.  A CTA owns a 128 x 256 tile of TWO
// consecutive orders (two 256-column accumulators = all 512 columns of tensor memory), so that every
// operand tile it fetches feeds two products: the slice pairs (s, o - s), (s, o + 1 - s) share A_s, and
// B_t meets A_(o-t), A_(o+1-t).  Operand tiles travel through two rings of single tiles (A: 16 KB,
// B: 32 KB, TMA boxes with the 128-byte swizzle) following a small dataflow program the host builds
// per order group (which tile to load next; which (A, B, accumulator) to multiply; which ring slot
// the finished MMAs release) -- 1.1 tile loads per slice product instead of 2, which is what the
// L2 -> shared-memory path could not deliver.  One elected thread issues tcgen05.mma.kind::i8 (SASS
// UTCIMMA), four epilogue warps read the accumulators back with tcgen05.ld (SASS LDTM);
// (2) the slicing kernels; (3) the fp64 recombination with the fused HMC epilogues.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "gemm.cuh"
#include "staged.cuh"

namespace hmcb {

constexpr int OZ_BM = 128, OZ_BN = 256, OZ_BK = 128;   // CTA tile; BK int8 = one 128-byte swizzle row
constexpr int OZ_NA = 6, OZ_NB = 4;                    // slots of the A / B tile rings
constexpr int OZ_A_BYTES = OZ_BM * OZ_BK, OZ_B_BYTES = OZ_BN * OZ_BK;   // 16 KB, 32 KB
constexpr size_t OZ_SMEM_BYTES = (size_t)OZ_NA * OZ_A_BYTES + (size_t)OZ_NB * OZ_B_BYTES + 1024;   // + alignment slack
constexpr int OZ_THREADS = 256;    // warp 0: TMA (A), warp 1: MMA, warp 2: TMEM allocation, warp 3: TMA (B), warps 4-7: epilogue
constexpr int OZ_BITS = 8;         // bits per slice (balanced digits in [-128, 127])
constexpr int OZ_MAX_SLICES = 7, OZ_MAX_ORDERS = 7;
constexpr int OZ_MAX_ACC = 2, OZ_MAX_OPS = 16, OZ_MAX_GROUPS = 8, OZ_PANEL = 8;

// Modular variant ("Ozaki II") for products with a long contraction: the operands are scaled to 44-bit
// integers, multiplied modulo 13 pairwise coprime moduli <= 256 (one exact int8 product each, residues in
// [-128, 127]) and the integer product is rebuilt from its residues (Chinese remainder theorem): 13 slice
// products instead of 21 for 2^-44 instead of 2^-48 of the row / chain maximum.  P = prod p = 2^102.5 must
// exceed 2 K 2^86: K < 46 000.
constexpr int OZ_NMOD = 13, OZ_CRT_BITS = 44;
__constant__ int c_oz_mod[OZ_NMOD] = {256, 255, 253, 251, 247, 241, 239, 233, 229, 227, 223, 217, 211};
// symmetric residue of an integer-valued double |x| < 2^52 modulo p (any representative in [-128, 127])
__device__ __forceinline__ int oz_residue(double x, int p) {
  const double q = rint(x * (1.0 / (double)p));
  int r = (int)fma(-q, (double)p, x);         // |r| <= (p + 1) / 2: the quotient may be off by one
  return r > 127 ? r - p : r;
}

// The dataflow program of one order group, the same for every k-block (built by oz_build_plan)
struct OzProgram {
  int n_acc, order[OZ_MAX_ACC];        // accumulators of the group and the order each one holds
  int n_loads, nA, nB;                 // operand tiles per k-block, in order of first use
  unsigned char load_is_b[OZ_MAX_OPS], load_slice[OZ_MAX_OPS];
  int n_mma;                           // slice products per k-block
  unsigned char mma_a[OZ_MAX_OPS], mma_b[OZ_MAX_OPS];   // index among the A / B loads of the k-block
  unsigned char mma_acc[OZ_MAX_OPS], mma_flags[OZ_MAX_OPS];   // 1: last use of A, 2: last use of B, 4: first product of its accumulator
};
struct OzPlan {
  int n_groups, tiles_m, tiles_n;
  OzProgram g[OZ_MAX_GROUPS];          // heaviest group first
};

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
template <bool PAIR>
__device__ __forceinline__ void umma_commit(uint64_t* bar) {   // arrives when the MMAs issued so far have completed
  if constexpr (PAIR)   // ... on the barrier at this address in both CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::
                     "r"((unsigned)__cvta_generic_to_shared(bar)), "h"((unsigned short)3) : "memory");
  else
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::
                     "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}

// C[o][m][n] = sum over slice pairs (s, t = o - s) of A_s[m][:] . B_t[n][:]   (int8 x int8 -> int32, exact)
//   mapA: {K, M, SA} int8, box {128, 128, 1};  mapB: {K, N, SB} int8, box {128, 256, 1} (PAIR: {128, 128, 1});
//   128-byte swizzle
// grid.x = groups x tiles: heaviest group first; inside a group panels of OZ_PANEL column tiles, row tiles
// fastest inside a panel, so a wave of CTAs covers a near-square region and shares its operand tiles in L2.
//
// PAIR = true: the CTA PAIR version (cluster of two CTAs on one TPC, tcgen05 cta_group::2).  The pair owns a
// 256 x 256 tile: CTA r holds rows m0 + 128 r of A and rows n0 + 128 r of B (HALF of the B tile), one thread of
// the leader CTA issues 256 x 256 x 32 MMAs that read both CTAs' shared memory and write each CTA's half of the
// accumulator into its own tensor memory.  Per CTA and slice product that is 16 KB + 16 KB of operand instead of
// 16 KB + 32 KB: a third less L2 -> SM traffic (the path that caps the single-CTA kernel) and half the
// shared-memory reads of B.  Both CTAs load (their TMA completes on the LEADER's full barrier, which expects the
// bytes of both), the leader's commits arrive on the empty barriers of both.
// RESIDUES = true: the accumulators hold products modulo c_oz_mod[P.order[a]]; the epilogue reduces them and
// stores int8 residue planes res[modulus][m_rows x ldc] directly (mapC unused).
template <bool PAIR, bool RESIDUES>
__global__ void __launch_bounds__(OZ_THREADS, 1)
i8_gemm_groups_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                      const __grid_constant__ CUtensorMap mapC, const __grid_constant__ OzPlan plan, int kblocks,
                      int ldc, int m_rows, signed char* __restrict__ res, long long res_plane) {
  constexpr int NB = PAIR ? 2 * OZ_NB : OZ_NB, B_BYTES = PAIR ? OZ_B_BYTES / 2 : OZ_B_BYTES;
  extern __shared__ __align__(1024) unsigned char oz_smem[];
  __shared__ uint64_t fullA[OZ_NA], emptyA[OZ_NA], fullB[NB], emptyB[NB], tmem_full_bar;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned rank = 0;
  if constexpr (PAIR) asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(rank));
  const int rows_m = PAIR ? (plan.tiles_m + 1) / 2 : plan.tiles_m;      // row tiles (pairs of row tiles)
  const int tiles = rows_m * plan.tiles_n;
  const int unit = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int gi = unit / tiles, tile = unit % tiles;
  const int panel = tile / (OZ_PANEL * rows_m), within = tile % (OZ_PANEL * rows_m);
  const int pw = min(OZ_PANEL, plan.tiles_n - panel * OZ_PANEL);
  const int m0 = ((within / pw) * (PAIR ? 2 : 1) + (int)rank) * OZ_BM, n0 = (panel * OZ_PANEL + within % pw) * OZ_BN;
  const OzProgram& P = plan.g[gi];
  const unsigned smemA = ((unsigned)__cvta_generic_to_shared(oz_smem) + 1023u) & ~1023u;
  const unsigned smemB = smemA + OZ_NA * OZ_A_BYTES;

  if (threadIdx.x == 0) {
    for (int s = 0; s < OZ_NA; ++s) { mbar_init(&fullA[s], 1); mbar_init(&emptyA[s], 1); }
    for (int s = 0; s < NB; ++s) { mbar_init(&fullB[s], 1); mbar_init(&emptyB[s], 1); }
    mbar_init(&tmem_full_bar, 1);
    fence_async_proxy();
    if constexpr (PAIR) asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 2) {   // one warp allocates tensor memory: all 512 columns (two accumulators; one CTA per SM)
    if constexpr (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::
                       "r"((unsigned)__cvta_generic_to_shared(&tmem_base_s)), "r"((unsigned)(OZ_MAX_ACC * OZ_BN)));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::
                       "r"((unsigned)__cvta_generic_to_shared(&tmem_base_s)), "r"((unsigned)(OZ_MAX_ACC * OZ_BN)));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if constexpr (PAIR) {   // the peer's barriers and tensor memory exist before anything crosses the pair
    asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = tmem_base_s;

  if (warp == 0 || warp == 3) {
    if (lane == 0) {   // ---- TMA producers: warp 0 walks the A loads of the group's list, warp 3 the B loads
      const bool is_b = warp == 3;
      const CUtensorMap* map = is_b ? &mapB : &mapA;
      uint64_t* full = is_b ? fullB : fullA;
      uint64_t* empty = is_b ? emptyB : emptyA;
      const int slots = is_b ? NB : OZ_NA, bytes = is_b ? B_BYTES : OZ_A_BYTES;
      const int row0 = is_b ? n0 + (PAIR ? (int)rank * (OZ_BN / 2) : 0) : m0;
      const unsigned base = is_b ? smemB : smemA;
      int seq = 0;
      for (int kb = 0; kb < kblocks; ++kb)
        for (int l = 0; l < P.n_loads; ++l) {
          if ((P.load_is_b[l] != 0) != is_b) continue;
          const int slot = seq % slots;
          if (seq >= slots) mbar_wait(&empty[slot], (unsigned)((seq / slots - 1) & 1));
          const unsigned bar = (unsigned)__cvta_generic_to_shared(&full[slot]);
          if constexpr (PAIR) {
            // the leader's barrier counts the bytes of both CTAs; the peer bit of the address is cleared so
            // that the peer's copy completes there too
            if (rank == 0) mbar_expect_tx(&full[slot], (unsigned)(2 * bytes));
            asm volatile(
                "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes "
                "[%0], [%1, {%2, %3, %4}], [%5];\n" ::"r"(base + (unsigned)slot * bytes), "l"(map), "r"(kb * OZ_BK),
                "r"(row0), "r"((int)P.load_slice[l]), "r"(bar & 0xFEFFFFFFu) : "memory");
          } else {
            mbar_expect_tx(&full[slot], (unsigned)bytes);
            tma_load_3d_raw(base + (unsigned)slot * bytes, map, kb * OZ_BK, row0, P.load_slice[l], bar);
          }
          ++seq;
        }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {   // ---- MMA issuer: one thread (of the leader CTA) drives the tensor core
      constexpr unsigned idesc = umma_idesc_i8(PAIR ? 2 * OZ_BM : OZ_BM, OZ_BN);
      for (int kb = 0; kb < kblocks; ++kb)
        for (int m = 0; m < P.n_mma; ++m) {
          const int a_seq = kb * P.nA + P.mma_a[m], b_seq = kb * P.nB + P.mma_b[m];
          const int sa = a_seq % OZ_NA, sb = b_seq % NB;
          const unsigned flags = P.mma_flags[m];
          mbar_wait(&fullA[sa], (unsigned)((a_seq / OZ_NA) & 1));
          mbar_wait(&fullB[sb], (unsigned)((b_seq / NB) & 1));
          asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
          const uint64_t a_desc = umma_desc_k_major_sw128(smemA + (unsigned)sa * OZ_A_BYTES);
          const uint64_t b_desc = umma_desc_k_major_sw128(smemB + (unsigned)sb * B_BYTES);
          const unsigned d_tmem = tmem + (unsigned)P.mma_acc[m] * OZ_BN;
          const bool fresh = kb == 0 && (flags & 4u);
#pragma unroll
          for (int k = 0; k < OZ_BK / 32; ++k) {     // K = 32 int8 per instruction: 32 bytes further along the row
            const unsigned accumulate = (fresh && k == 0) ? 0u : 1u;
            if constexpr (PAIR)
              asm volatile(
                  "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                  "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n}\n" ::
                      "r"(d_tmem), "l"(a_desc + (uint64_t)(2 * k)), "l"(b_desc + (uint64_t)(2 * k)), "r"(idesc),
                  "r"(accumulate), "r"(0u) : "memory");
            else
              asm volatile(
                  "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                  "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n}\n" ::
                      "r"(d_tmem), "l"(a_desc + (uint64_t)(2 * k)), "l"(b_desc + (uint64_t)(2 * k)), "r"(idesc),
                  "r"(accumulate), "r"(0u) : "memory");
          }
          if (flags & 1u) umma_commit<PAIR>(&emptyA[sa]);
          if (flags & 2u) umma_commit<PAIR>(&emptyB[sb]);
        }
      umma_commit<PAIR>(&tmem_full_bar);
    }
  } else if (warp >= 4) {
    // ---- epilogue: warp w reads TMEM lanes 32 (w % 4) .. +31 = rows m0 + 32 (w - 4) + lane, 32 columns at a
    // time, lays the 32 x 32 block out in shared memory the way a 128-byte-swizzled TMA box expects it (lane =
    // row: its eight 16-byte pieces land in eight different bank groups) and one lane hands it to the TMA
    // engine, which writes whole 128-byte lines of the order plane.  The operand rings are free by now (every
    // MMA has completed): two 4 KB buffers per warp out of the A ring.
    mbar_wait(&tmem_full_bar, 0u);
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const int row0 = m0 + (warp - 4) * 32;
    const unsigned stage0 = smemA + (unsigned)(warp - 4) * 8192u;
    int it = 0;
    for (int a = 0; a < P.n_acc && row0 < m_rows; ++a) {
      const uint32_t taddr = tmem + ((uint32_t)((warp - 4) * 32) << 16) + (uint32_t)(a * OZ_BN);
#pragma unroll 1
      for (int c = 0; c < OZ_BN && n0 + c < ldc; c += 32, ++it) {
        uint32_t r[32];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
              "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
              "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr + (uint32_t)c));
        if constexpr (RESIDUES) {
          // lane = row: its 32 columns reduced modulo the accumulator's modulus, packed into 32 bytes
          asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
          const int pm = c_oz_mod[P.order[a]];
          unsigned w[8];
#pragma unroll
          for (int v = 0; v < 8; ++v) {
            unsigned word = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j)
              word |= ((unsigned)oz_residue((double)(int)r[4 * v + j], pm) & 255u) << (8 * j);
            w[v] = word;
          }
          if (row0 + lane < m_rows) {
            uint4* dst = reinterpret_cast<uint4*>(res + (size_t)P.order[a] * res_plane + (size_t)(row0 + lane) * ldc + n0 + c);
            dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
            dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
          }
          continue;
        }
        if (it >= 2) {   // the buffer written two blocks ago must have been read by its store
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;\n" ::: "memory");
          __syncwarp();
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
        const unsigned buf = stage0 + (unsigned)(it & 1) * 4096u, line = buf + (unsigned)lane * 128u;
#pragma unroll
        for (int v = 0; v < 8; ++v)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(line + (unsigned)((v ^ (lane & 7)) << 4)),
                       "r"(r[4 * v]), "r"(r[4 * v + 1]), "r"(r[4 * v + 2]), "r"(r[4 * v + 3]) : "memory");
        fence_async_proxy();
        __syncwarp();
        if (lane == 0) {
          asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];\n" ::"l"(&mapC),
                       "r"(n0 + c), "r"(row0), "r"(P.order[a]), "r"(buf) : "memory");
          asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
        }
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");   // before the buffers go away
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if constexpr (PAIR) {   // neither CTA leaves (or frees tensor memory) while the other may still touch it
    asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
  }
  if (warp == 2) {
    if constexpr (PAIR)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"((unsigned)(OZ_MAX_ACC * OZ_BN)));
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"((unsigned)(OZ_MAX_ACC * OZ_BN)));
  }
}


// ---- slicing and recombination --------------------------------------------------------------------

// exponent e with |x| 2^-e < 0.494 for every |x| <= max (so that the leading balanced digit stays inside
// int8 after the carries), from the bits of max |x|; 0 for an all-zero (or subnormal) maximum; INT_MIN
// marks a maximum that is inf / NaN or next to the overflow threshold: the products are NaN then
__host__ __device__ __forceinline__ int oz_exponent(unsigned long long maxbits) {
  const int e = (int)(maxbits >> 52);
  if (e >= 2040) return INT_MIN;
  if (e == 0) return 0;
  const int top6 = (int)((maxbits >> 46) & 63ull);   // max = 1.m 2^(e-1023) < 2^(e-1022); 1.m / 4 < 0.494 unless m is near 1
  return e - 1022 + 1 + (top6 >= 62 ? 1 : 0);
}
__host__ __device__ __forceinline__ double oz_pow2(int e) {   // 2^e for e in [-1022, 1023]
  union { unsigned long long u; double d; } v;
  v.u = (unsigned long long)(e + 1023) << 52;
  return v.d;
}
// x 2^-e (exact; e may lie outside the range of a single power-of-two factor)
__host__ __device__ __forceinline__ double oz_scale_down(double x, int e) {
  const int h = e / 2;
  return x * oz_pow2(-h) * oz_pow2(-(e - h));
}
// balanced radix-256 digits of X (|X| < 0.496 * 256^S): digit t = S-1 is the least significant
__host__ __device__ __forceinline__ void oz_digits(long long X, int S, signed char (&digit)[OZ_MAX_SLICES]) {
#pragma unroll
  for (int t = OZ_MAX_SLICES - 1; t >= 0; --t)
    if (t < S) {
      const int d = (int)((X + 128) & 255) - 128;
      X = (X - d) >> 8;
      digit[t] = (signed char)d;
    }
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

// all balanced digits of X at once: with u_t = d_t + 128 in [0, 255], sum_t u_t 256^t = X + 128 sum_t 256^t, so the
// bytes of X + 0x80...80 are the u_t and flipping their top bits gives the two's-complement d_t (byte 0 = the
// least significant digit = digit S-1).  Same digits as oz_digits: the balanced representation is unique.
__device__ __forceinline__ unsigned long long oz_digit_bytes(long long X, int S) {
  const unsigned long long bias = 0x8080808080808080ull >> (8 * (8 - S));
  return ((unsigned long long)X + bias) ^ bias;
}

// X [K x ld] (chains contiguous) -> SB int8 digit planes, chain-major [t][ld][K] (K contiguous): the B
// operand of the tensor-core product.  Block = 128 k x 32 chains, transposed through shared memory: a thread
// slices four consecutive k of its chain and packs each plane's four digits into one 32-bit word.
__global__ void __launch_bounds__(256)
oz_slice_chains_kernel(const double* __restrict__ X, int K, int ld, int SB, const unsigned long long* __restrict__ maxbits,
                       signed char* __restrict__ out) {
  __shared__ __align__(16) signed char sl[OZ_MAX_SLICES][32][132];
  const int k0 = blockIdx.y * 128, c0 = blockIdx.x * 32;
  const int tc = threadIdx.x & 31, tr = threadIdx.x >> 5;
  const int eb = oz_exponent(maxbits[c0 + tc]);
  const double up = oz_pow2(OZ_BITS * SB);
  double x[4][4];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = (tr + 8 * j) * 4 + i;
      x[j][i] = (k0 + r < K) ? __ldcs(X + (size_t)(k0 + r) * ld + c0 + tc) : 0.0;
    }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    unsigned lo[4], hi[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const long long Xi = eb != INT_MIN ? __double2ll_rn(oz_scale_down(x[j][i], eb) * up) : 0ll;
      const unsigned long long y = oz_digit_bytes(Xi, SB);
      lo[i] = (unsigned)y;
      hi[i] = (unsigned)(y >> 32);
    }
    const int r = (tr + 8 * j) * 4;
#pragma unroll
    for (int t = 0; t < OZ_MAX_SLICES; ++t)
      if (t < SB) {
        const int byte = SB - 1 - t;                       // digit t lives in byte S-1-t
        const unsigned* w = byte < 4 ? lo : hi;
        const unsigned sel = 0x4040u + (unsigned)(byte & 3) * 0x1111u;   // result bytes 0, 1 <- byte b of x, byte b of y
        const unsigned p01 = __byte_perm(w[0], w[1], sel), p23 = __byte_perm(w[2], w[3], sel);
        *reinterpret_cast<unsigned*>(&sl[t][tc][r]) = __byte_perm(p01, p23, 0x5410u);
      }
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int row = warp; row < SB * 32; row += 8) {
    const int t = row >> 5, ch = row & 31;
    const unsigned v = *reinterpret_cast<const unsigned*>(&sl[t][ch][lane * 4]);
    *reinterpret_cast<unsigned*>(out + ((size_t)t * ld + c0 + ch) * K + k0 + lane * 4) = v;
  }
}

// ---- modular variant: residues of the chain batch, reconstruction from the residue planes -------------

// X [K x ld] (chains contiguous) -> OZ_NMOD residue planes, chain-major [m][ld][K] (K contiguous): every value
// scaled to a 44-bit integer (per-chain exponent as above) and reduced modulo each modulus.  Block = 128 k x 16
// chains, transposed through shared memory (a thread reduces four consecutive k of its chain and packs each
// plane's four residues into one word).
__global__ void __launch_bounds__(256)
oz_residue_chains_kernel(const double* __restrict__ X, int K, int ld, const unsigned long long* __restrict__ maxbits,
                         signed char* __restrict__ out) {
  __shared__ __align__(16) signed char sl[OZ_NMOD][16][132];
  const int k0 = blockIdx.y * 128, c0 = blockIdx.x * 16;
  const int tc = threadIdx.x & 15, tr = threadIdx.x >> 4;
  const int eb = oz_exponent(maxbits[c0 + tc]);
  const double up = oz_pow2(OZ_CRT_BITS);
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int r = (tr + 16 * h) * 4;
    double xi[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const double x = (k0 + r + i < K) ? __ldcs(X + (size_t)(k0 + r + i) * ld + c0 + tc) : 0.0;
      xi[i] = eb != INT_MIN ? rint(oz_scale_down(x, eb) * up) : 0.0;      // |xi| < 2^43, an integer
    }
#pragma unroll
    for (int m = 0; m < OZ_NMOD; ++m) {
      unsigned word = 0;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int res = m == 0 ? (int)(signed char)(__double2ll_rn(xi[i]) & 255ll) : oz_residue(xi[i], c_oz_mod[m]);
        word |= ((unsigned)res & 255u) << (8 * i);
      }
      *reinterpret_cast<unsigned*>(&sl[m][tc][r]) = word;
    }
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int row = warp; row < OZ_NMOD * 16; row += 8) {
    const int m = row >> 4, ch = row & 15;
    const unsigned v = *reinterpret_cast<const unsigned*>(&sl[m][ch][lane * 4]);
    *reinterpret_cast<unsigned*>(out + ((size_t)m * ld + c0 + ch) * K + k0 + lane * 4) = v;
  }
}

// Chinese remainder reconstruction.  x = sum_m c_m W_m mod P with W_m / P = y_m / p_m (y_m the inverse of P / p_m
// modulo p_m): x / P = frac(sum_m c_m f_m), f_m = y_m / p_m held as three 40-bit chunks, so that the three sums
// of at most 13 products |c| 2^40 are exact in fp64 and the fractional part is good to 2^-109.
struct OzCrt { double F[OZ_NMOD][3]; double P; };

// Y[i][c] = 2^(ea[i] + eb[c] - 88) x[i][c], x rebuilt from the residue planes res[m][rows x ld] -> epilogue
// functor (row interface as oz_combine_kernel); a thread handles four neighbouring chains (one word per plane)
template <class Epi, int ROWS>
__global__ void __launch_bounds__(128)
oz_crt_combine_kernel(const signed char* __restrict__ res, long long plane, int rows, int ld,
                      const int* __restrict__ ea, const unsigned long long* __restrict__ maxbits_in,
                      const __grid_constant__ OzCrt crt, Epi epi) {
  const int c = (blockIdx.x * 128 + threadIdx.x) * 4;
  if (c >= ld) return;
  int eb[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) eb[j] = oz_exponent(maxbits_in[c + j]);
  const int r1 = min(rows, ((int)blockIdx.y + 1) * ROWS);
  Epi fn = epi;
  for (int i = blockIdx.y * ROWS; i < r1; ++i) {
    unsigned w[OZ_NMOD];
#pragma unroll
    for (int m = 0; m < OZ_NMOD; ++m)
      w[m] = __ldcs(reinterpret_cast<const unsigned*>(res + (size_t)m * plane + (size_t)i * ld + c));
    const int ei = ea[i];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      double s0 = 0.0, s1 = 0.0, s2 = 0.0;
#pragma unroll
      for (int m = 0; m < OZ_NMOD; ++m) {
        const double cm = (double)(int)(signed char)((w[m] >> (8 * j)) & 255u);
        s0 = fma(cm, crt.F[m][0], s0); s1 = fma(cm, crt.F[m][1], s1); s2 = fma(cm, crt.F[m][2], s2);
      }
      const double u = s0 * 0x1p-40;                       // exact; its fractional part is a multiple of 2^-40
      const double frac = (u - rint(u)) + (s1 * 0x1p-80 + s2 * 0x1p-120);
      double y;
      if (eb[j] == INT_MIN) y = CUDART_NAN;
      else {
        const int e = ei + eb[j] - 2 * OZ_CRT_BITS, h = e / 2;
        y = (frac - rint(frac)) * crt.P * oz_pow2(h) * oz_pow2(e - h);
      }
      fn.row(i, c + j, y);
    }
  }
}

// Y[i][c] = 2^(ea[i] + eb[c]) sum_o 256^-(o + 2) C_o[i][c]  -> epilogue functor (SpMM interface: tile_begin /
// row / chunk_end; a block = 128 chains x ROWS rows); optionally the per-chain max |R| of what a ResidualEpi
// stored (the scale of the next slicing pass)
template <class Epi, bool TRACK_MAX, int ROWS, int ORD>
__global__ void __launch_bounds__(128)
oz_combine_kernel(const int* __restrict__ C, long long plane_stride, int rows, int ld, int orders,
                  const int* __restrict__ ea, const unsigned long long* __restrict__ maxbits_in, Epi epi,
                  unsigned long long* __restrict__ maxbits_out) {
  // ORD > 0: the number of order planes at compile time (all loads of two rows are issued before the
  // first use: the kernel streams 4 ORD bytes per element and lives on loads in flight); ORD = 0: `orders`
  constexpr int NV = ORD > 0 ? ORD : OZ_MAX_ORDERS;
  const int n_ord = ORD > 0 ? ORD : orders;
  const int c = blockIdx.x * 128 + threadIdx.x;
  if (c >= ld) return;
  const int eb = oz_exponent(maxbits_in[c]);
  unsigned long long m = 0ull;
  const int r1 = min(rows, ((int)blockIdx.y + 1) * ROWS);
  Epi fn = epi;
  fn.tile_begin(blockIdx.y, c);
  auto finish = [&](int i, const int (&v)[NV]) {
    double acc = 0.0;
#pragma unroll
    for (int o = NV - 1; o >= 0; --o)
      if (o < n_ord) acc = fma(acc, 0.00390625, (double)v[o]);
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
      fn.row(i, c, y);
    }
  };
  for (int i = blockIdx.y * ROWS; i < r1; i += 2) {
    const int* p = C + (size_t)i * ld + c;
    const bool two = i + 1 < r1;
    int v0[NV], v1[NV];
#pragma unroll
    for (int o = 0; o < NV; ++o) {
      v0[o] = o < n_ord ? __ldcs(p + (size_t)o * plane_stride) : 0;
      v1[o] = (o < n_ord && two) ? __ldcs(p + (size_t)o * plane_stride + ld) : 0;
    }
    finish(i, v0);
    if (two) finish(i + 1, v1);
  }
  fn.chunk_end(blockIdx.y, c);
  if constexpr (TRACK_MAX) {
    if (m) atomicMax(maxbits_out + c, m);
  }
}

}  // namespace hmcb

// ozaki_sparse.cuh -- the SPARSE LinearMatrix products (CSR G, straight-ray tomography) on tcgen05.
//
// A sparse matrix whose rows cluster is block-sparse in the right basis: a BUNDLE of 128 similar rows
// (rays that run side by side) touches only a short list of columns (the cells of a band), and on that
// list the bundle is a dense 128 x K_b tile -- mostly zeros (11 % fill for 50 000 rays on a 100 x 100
// grid), but the int8 tensor pipe turns over 80 x more multiply-adds per second than the fp64 pipes,
// so 20 exact slice products over that tile still cost a third of what the row-blocked DMMA kernel
// needs for the nonzeros alone.  Per bundle the host stores the column list (padded to whole k-blocks
// of 128) and the balanced int8 digits of the tile ([digit][128 rows][all bundles' lists end to end],
// loaded by the same tensor map as a dense model matrix); the chain batch is sliced into digit planes
// [digit][row of B][chain] with the CHAINS contiguous, so that the B row of any listed column is one
// contiguous 256-byte piece per tile.
//
// Kernel = ozaki.cuh's slice-product kernel (dataflow program per order group, tile rings, one MMA
// thread, TMEM accumulators, TMA-stored order planes) with one change: the B tiles are GATHERED.  Eight
// producer warps replace the B TMA lane: warp w owns chains 32 w .. 32 w + 31 of the tile, lane g the
// four list entries 4 g .. 4 g + 3 of the k-block; a lane loads 32 bytes (32 chains) of each of its
// four B rows, transposes 4 x 4 byte blocks with byte permutes and stores, per chain, one 32-bit word
// (four consecutive k) into the K-major 128-byte-swizzled tile the MMA descriptor expects -- the 32
// lanes of a store hit 32 different banks.  ~1 warp instruction per clock and SM, 28 % of the issue
// slots, next to a tensor pipe that needs 128 clocks per instruction.
#pragma once
#include "ozaki.cuh"

namespace hmcb {

constexpr int OZS_THREADS = 512;       // warps 0-7 as in ozaki.cuh (warp 3 idle), warps 8-15: B gather
constexpr int OZS_GATHER_WARPS = 8;

struct OzBundle { int koff, kblocks; };   // offset of the bundle's list on the concatenated K axis; k-blocks

// transpose of a 4 x 4 byte block: in w[c] = bytes (chain 0..3) of list entry c -> out v[j] = bytes
// (entry 0..3) of chain j
__device__ __forceinline__ void oz_transpose4(const unsigned (&w)[4], unsigned (&v)[4]) {
  const unsigned t0 = __byte_perm(w[0], w[1], 0x5140u), t1 = __byte_perm(w[2], w[3], 0x5140u);
  const unsigned t2 = __byte_perm(w[0], w[1], 0x7362u), t3 = __byte_perm(w[2], w[3], 0x7362u);
  v[0] = __byte_perm(t0, t1, 0x5410u); v[1] = __byte_perm(t0, t1, 0x7632u);
  v[2] = __byte_perm(t2, t3, 0x5410u); v[3] = __byte_perm(t2, t3, 0x7632u);
}

// C[o][b * 128 + r][n] = sum over slice pairs (s, o - s) of sum_j A_s[r][koff_b + j] * Bg_(o-s)[list[koff_b + j]][n]
//   mapA: {Ktot, 128, SA} int8, box {128, 128, 1}, 128-byte swizzle;  Bg: [SB][rows_b x ld] int8, chains contiguous
//   mapC: order planes [orders][n_bundles * 128 x ldc]
// grid.x = groups x (bundles x column tiles): heaviest group first; panels of OZ_PANEL column tiles, bundles
// (sorted by length) fastest inside a panel: a wave works on few column tiles (their B rows stay in L2) and
// on neighbouring bundles
__global__ void __launch_bounds__(OZS_THREADS, 1)
i8_gather_gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapC,
                      const __grid_constant__ OzPlan plan, const OzBundle* __restrict__ bundles,
                      const int* __restrict__ list, const signed char* __restrict__ Bg, long long bg_plane, int ld,
                      int ldc) {
  extern __shared__ __align__(1024) unsigned char oz_smem[];
  __shared__ uint64_t fullA[OZ_NA], emptyA[OZ_NA], fullB[OZ_NB], emptyB[OZ_NB], tmem_full_bar;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_bundles = plan.tiles_m;
  const int tiles = n_bundles * plan.tiles_n;
  const int gi = (int)blockIdx.x / tiles, tile = (int)blockIdx.x % tiles;
  const int panel = tile / (OZ_PANEL * n_bundles), within = tile % (OZ_PANEL * n_bundles);
  const int pw = min(OZ_PANEL, plan.tiles_n - panel * OZ_PANEL);
  const int b = within / pw, n0 = (panel * OZ_PANEL + within % pw) * OZ_BN;
  const int m0 = b * OZ_BM;
  const OzBundle bun = bundles[b];
  const int kblocks = bun.kblocks;
  const OzProgram& P = plan.g[gi];
  const unsigned smemA = ((unsigned)__cvta_generic_to_shared(oz_smem) + 1023u) & ~1023u;
  const unsigned smemB = smemA + OZ_NA * OZ_A_BYTES;

  if (threadIdx.x == 0) {
    for (int s = 0; s < OZ_NA; ++s) { mbar_init(&fullA[s], 1); mbar_init(&emptyA[s], 1); }
    for (int s = 0; s < OZ_NB; ++s) { mbar_init(&fullB[s], OZS_GATHER_WARPS); mbar_init(&emptyB[s], 1); }
    mbar_init(&tmem_full_bar, 1);
    fence_async_proxy();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::
                     "r"((unsigned)__cvta_generic_to_shared(&tmem_base_s)), "r"((unsigned)(OZ_MAX_ACC * OZ_BN)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = tmem_base_s;

  if (warp == 0) {
    if (lane == 0) {   // ---- TMA producer of the model digits
      int seq = 0;
      for (int kb = 0; kb < kblocks; ++kb)
        for (int l = 0; l < P.n_loads; ++l) {
          if (P.load_is_b[l]) continue;
          const int slot = seq % OZ_NA;
          if (seq >= OZ_NA) mbar_wait(&emptyA[slot], (unsigned)((seq / OZ_NA - 1) & 1));
          mbar_expect_tx(&fullA[slot], (unsigned)OZ_A_BYTES);
          tma_load_3d_raw(smemA + (unsigned)slot * OZ_A_BYTES, &mapA, bun.koff + kb * OZ_BK, 0, P.load_slice[l],
                          (unsigned)__cvta_generic_to_shared(&fullA[slot]));
          ++seq;
        }
    }
  } else if (warp >= 8) {
    // ---- gather producers of the chain digits: warp w -> chains n0 + 32 w .., lane g -> list entries 4 g ..
    const int w = warp - 8, chain0 = n0 + 32 * w;
    const bool live = chain0 < ld;
    int seq = 0;
    for (int kb = 0; kb < kblocks; ++kb) {
      const int4 rows4 = __ldg(reinterpret_cast<const int4*>(list + bun.koff + kb * OZ_BK) + lane);
      const int rows[4] = {rows4.x, rows4.y, rows4.z, rows4.w};
      for (int l = 0; l < P.n_loads; ++l) {
        if (!P.load_is_b[l]) continue;
        const int slot = seq % OZ_NB;
        if (seq >= OZ_NB) mbar_wait(&emptyB[slot], (unsigned)((seq / OZ_NB - 1) & 1));
        const signed char* plane = Bg + (size_t)P.load_slice[l] * bg_plane + chain0;
        int4 v[4][2];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (live) {
            const int4* p = reinterpret_cast<const int4*>(plane + (size_t)rows[c] * ld);
            v[c][0] = __ldg(p); v[c][1] = __ldg(p + 1);
          } else {
            v[c][0] = make_int4(0, 0, 0, 0); v[c][1] = make_int4(0, 0, 0, 0);
          }
        }
        const unsigned tileB = smemB + (unsigned)slot * OZ_B_BYTES;
#pragma unroll
        for (int a = 0; a < 8; ++a) {   // chains 4 a .. 4 a + 3 of the warp's 32
          unsigned in[4], out[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int4 q = v[c][a >> 2];
            in[c] = (unsigned)((a & 3) == 0 ? q.x : (a & 3) == 1 ? q.y : (a & 3) == 2 ? q.z : q.w);
          }
          oz_transpose4(in, out);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const unsigned row = (unsigned)(32 * w + 4 * a + j);      // chain row of the tile
            const unsigned addr = tileB + (row >> 3) * 1024u + (row & 7u) * 128u +
                                  ((((unsigned)lane >> 2) ^ (row & 7u)) << 4) + ((unsigned)lane & 3u) * 4u;
            asm volatile("st.shared.b32 [%0], %1;\n" ::"r"(addr), "r"(out[j]) : "memory");
          }
        }
        fence_async_proxy();
        __syncwarp();
        if (lane == 0) mbar_arrive(&fullB[slot]);
        ++seq;
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {   // ---- MMA issuer
      constexpr unsigned idesc = umma_idesc_i8(OZ_BM, OZ_BN);
      for (int kb = 0; kb < kblocks; ++kb)
        for (int m = 0; m < P.n_mma; ++m) {
          const int a_seq = kb * P.nA + P.mma_a[m], b_seq = kb * P.nB + P.mma_b[m];
          const int sa = a_seq % OZ_NA, sb = b_seq % OZ_NB;
          const unsigned flags = P.mma_flags[m];
          mbar_wait(&fullA[sa], (unsigned)((a_seq / OZ_NA) & 1));
          mbar_wait(&fullB[sb], (unsigned)((b_seq / OZ_NB) & 1));
          asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
          const uint64_t a_desc = umma_desc_k_major_sw128(smemA + (unsigned)sa * OZ_A_BYTES);
          const uint64_t b_desc = umma_desc_k_major_sw128(smemB + (unsigned)sb * OZ_B_BYTES);
          const unsigned d_tmem = tmem + (unsigned)P.mma_acc[m] * OZ_BN;
          const bool fresh = kb == 0 && (flags & 4u);
#pragma unroll
          for (int k = 0; k < OZ_BK / 32; ++k) {
            const unsigned accumulate = (fresh && k == 0) ? 0u : 1u;
            asm volatile(
                "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n}\n" ::
                    "r"(d_tmem), "l"(a_desc + (uint64_t)(2 * k)), "l"(b_desc + (uint64_t)(2 * k)), "r"(idesc),
                "r"(accumulate), "r"(0u) : "memory");
          }
          if (flags & 1u) umma_commit<false>(&emptyA[sa]);
          if (flags & 2u) umma_commit<false>(&emptyB[sb]);
        }
      umma_commit<false>(&tmem_full_bar);
    }
  } else if (warp >= 4 && warp < 8) {
    // ---- epilogue: as in ozaki.cuh (TMEM -> registers -> swizzled staging -> TMA store of the order planes)
    mbar_wait(&tmem_full_bar, 0u);
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const int row0 = m0 + (warp - 4) * 32;
    const unsigned stage0 = smemA + (unsigned)(warp - 4) * 8192u;
    int it = 0;
    for (int a = 0; a < P.n_acc; ++a) {
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
        if (it >= 2) {
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
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 2)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"((unsigned)(OZ_MAX_ACC * OZ_BN)));
}

// X [K x ld] (chains contiguous) -> SB digit planes [t][K x ld], chains contiguous (the layout the gather
// producers read); a thread slices four consecutive chains of a row and writes one word per plane
__global__ void __launch_bounds__(256)
oz_slice_plain_kernel(const double* __restrict__ X, int K, int ld, int SB, const unsigned long long* __restrict__ maxbits,
                      signed char* __restrict__ out, long long plane) {
  const int c = (blockIdx.x * 256 + threadIdx.x) * 4;
  if (c >= ld) return;
  int eb[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) eb[j] = oz_exponent(maxbits[c + j]);
  const double up = oz_pow2(OZ_BITS * SB);
  const int r1 = min(K, ((int)blockIdx.y + 1) * 8);
  for (int r = blockIdx.y * 8; r < r1; ++r) {
    const double4 x = *reinterpret_cast<const double4*>(X + (size_t)r * ld + c);
    const double xs[4] = {x.x, x.y, x.z, x.w};
    unsigned lo[4], hi[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long Xi = eb[j] != INT_MIN ? __double2ll_rn(oz_scale_down(xs[j], eb[j]) * up) : 0ll;
      const unsigned long long y = oz_digit_bytes(Xi, SB);
      lo[j] = (unsigned)y;
      hi[j] = (unsigned)(y >> 32);
    }
#pragma unroll
    for (int t = 0; t < OZ_MAX_SLICES; ++t)
      if (t < SB) {
        const int byte = SB - 1 - t;
        const unsigned* w = byte < 4 ? lo : hi;
        const unsigned sel = 0x4040u + (unsigned)(byte & 3) * 0x1111u;
        const unsigned p01 = __byte_perm(w[0], w[1], sel), p23 = __byte_perm(w[2], w[3], sel);
        *reinterpret_cast<unsigned*>(out + (size_t)t * plane + (size_t)r * ld + c) = __byte_perm(p01, p23, 0x5410u);
      }
  }
}

}  // namespace hmcb

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
// producer warps replace the B TMA lane: they load the listed B rows (coalesced: 128 contiguous bytes of
// a row per 8 lanes), transpose 4 x 4 byte blocks with byte permutes and store, per chain, one 32-bit
// word (four consecutive k) into the K-major 128-byte-swizzled tile the MMA descriptor expects (details
// at the producer code).
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
// grid.x = groups x (bundles x column tiles): heaviest group first; panels of `panel_tiles` column tiles, bundles
// (sorted by length) fastest inside a panel: a wave works on few column tiles (their B rows stay in L2) and
// on neighbouring bundles
__global__ void __launch_bounds__(OZS_THREADS, 1)
i8_gather_gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapC,
                      const __grid_constant__ OzPlan plan, const OzBundle* __restrict__ bundles,
                      const int* __restrict__ list, const signed char* __restrict__ Bg, long long bg_plane, int ld,
                      int ldc, int panel_tiles) {
  extern __shared__ __align__(1024) unsigned char oz_smem[];
  __shared__ uint64_t fullA[OZ_NA], emptyA[OZ_NA], fullB[OZ_NB], emptyB[OZ_NB], tmem_full_bar;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_bundles = plan.tiles_m;
  // the order groups of a (bundle, column tile) run next to each other: they read the same operands
  const int gi = (int)blockIdx.x % plan.n_groups, tile = (int)blockIdx.x / plan.n_groups;
  const int panel = tile / (panel_tiles * n_bundles), within = tile % (panel_tiles * n_bundles);
  const int pw = min(panel_tiles, plan.tiles_n - panel * panel_tiles);
  const int b = within / pw, n0 = (panel * panel_tiles + within % pw) * OZ_BN;
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
    // ---- gather producers of the chain digits.  Warp w owns the list entries 16 w .. 16 w + 15 of every
    // k-block (four k-quads), lane (q, h) = (lane >> 3, lane & 7) the k-quad 4 w + q and 16 bytes of each
    // 128-chain half of its four B rows: one warp-level load touches 4 rows x 128 contiguous bytes = four
    // lines (a lane-per-row mapping costs the L1 tag stage a lookup per 16 bytes and runs 4x slower).  The
    // digit planes hold the chains of a 128-chain block PERMUTED -- byte 16 h + m <-> chain h + 8 m -- so
    // that the four bytes of a loaded word are chains 8 apart: after the 4 x 4 byte transpose the 32 lanes
    // of a store write rows with 8 different swizzle phases x 4 different words = 32 different banks.
    // The loads of the NEXT tile are in flight while the current one is transposed and stored.
    const int w = warp - 8, q = lane >> 3, h = lane & 7;
    unsigned long long bsl = 0ull;   // slice of the i-th B load of a k-block, 4 bits each
    {
      int i = 0;
      for (int l = 0; l < P.n_loads; ++l)
        if (P.load_is_b[l]) bsl |= (unsigned long long)P.load_slice[l] << (4 * i++);
    }
    const int nB = P.nB;
    const bool live0 = n0 < ld, live1 = n0 + 128 < ld;
    const int4* lst = reinterpret_cast<const int4*>(list + bun.koff) + 4 * w + q;   // + 32 per k-block
    auto load_tile = [&](int4 (&v)[4][2], int slice, const int4& r4) {
      const signed char* plane = Bg + (size_t)slice * bg_plane + n0 + 16 * h;
      const int rows[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const signed char* p = plane + (size_t)rows[c] * ld;
        v[c][0] = live0 ? __ldg(reinterpret_cast<const int4*>(p)) : make_int4(0, 0, 0, 0);
        v[c][1] = live1 ? __ldg(reinterpret_cast<const int4*>(p + 128)) : make_int4(0, 0, 0, 0);
      }
    };
    int4 rows_cur = __ldg(lst), rows_nxt = kblocks > 1 ? __ldg(lst + 32) : rows_cur;
    int4 v[4][2];
    load_tile(v, (int)(bsl & 15ull), rows_cur);
    // byte offset of (row h of a swizzle atom, k-quad 4 w + q): the row adds 1024 (row >> 3) per atom
    const unsigned lane_off = (unsigned)h * 128u + ((unsigned)(w ^ h) << 4) + (unsigned)q * 4u;
    int seq = 0;
    for (int kb = 0; kb < kblocks; ++kb) {
      const int4 rows_n2 = kb + 2 < kblocks ? __ldg(lst + 32 * (kb + 2)) : rows_nxt;
      for (int i = 0; i < nB; ++i) {
        int4 vn[4][2];
        const bool last_of_kb = i + 1 == nB;
        if (!(last_of_kb && kb + 1 == kblocks))
          load_tile(vn, (int)((bsl >> (4 * (last_of_kb ? 0 : i + 1))) & 15ull), last_of_kb ? rows_nxt : rows_cur);
        const int slot = seq % OZ_NB;
        if (seq >= OZ_NB) mbar_wait(&emptyB[slot], (unsigned)((seq / OZ_NB - 1) & 1));
        const unsigned tileB = smemB + (unsigned)slot * OZ_B_BYTES + lane_off;
#pragma unroll
        for (int half = 0; half < 2; ++half)
#pragma unroll
          for (int u = 0; u < 4; ++u) {   // word u of the 16 bytes: chains h + 8 (4 u + j), j = 0..3, of the half
            unsigned in[4], out[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const int4 x = v[c][half];
              in[c] = (unsigned)(u == 0 ? x.x : u == 1 ? x.y : u == 2 ? x.z : x.w);
            }
            oz_transpose4(in, out);
#pragma unroll
            for (int j = 0; j < 4; ++j)   // row 128 half + h + 8 (4 u + j): atom 16 half + 4 u + j
              asm volatile("st.shared.b32 [%0], %1;\n" ::"r"(tileB + (unsigned)(16 * half + 4 * u + j) * 1024u),
                           "r"(out[j]) : "memory");
          }
        fence_async_proxy();
        __syncwarp();
        if (lane == 0) mbar_arrive(&fullB[slot]);
        ++seq;
#pragma unroll
        for (int c = 0; c < 4; ++c) { v[c][0] = vn[c][0]; v[c][1] = vn[c][1]; }
      }
      rows_cur = rows_nxt;
      rows_nxt = rows_n2;
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

// position of chain n inside its digit-plane row: the chains of a 128-chain block are permuted, byte
// 16 h + m <-> chain h + 8 m (see the gather producers)
__host__ __device__ __forceinline__ int oz_plane_pos(int n) { return (n & ~127) + 16 * (n & 7) + ((n & 127) >> 3); }

// X [K x ld] (chains contiguous) -> SB digit planes [t][K x ld] in the gather layout (oz_plane_pos); a thread
// slices the four chains of one plane word (8 apart) of a row and writes one word per plane.  ld % 128 == 0.
__global__ void __launch_bounds__(256)
oz_slice_plain_kernel(const double* __restrict__ X, int K, int ld, int SB, const unsigned long long* __restrict__ maxbits,
                      signed char* __restrict__ out, long long plane) {
  const int word = blockIdx.x * 256 + threadIdx.x;       // word of a plane row: bytes 4 word .. 4 word + 3
  if (4 * word >= ld) return;
  const int blk = (4 * word) & ~127, p = (4 * word) & 127, h = p >> 4, m0 = p & 15;   // chains h + 8 (m0 + j)
  int eb[4], col[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) { col[j] = blk + h + 8 * (m0 + j); eb[j] = oz_exponent(maxbits[col[j]]); }
  const double up = oz_pow2(OZ_BITS * SB);
  const int r1 = min(K, ((int)blockIdx.y + 1) * 8);
  for (int r = blockIdx.y * 8; r < r1; ++r) {
    unsigned lo[4], hi[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const double x = __ldcs(X + (size_t)r * ld + col[j]);
      const long long Xi = eb[j] != INT_MIN ? __double2ll_rn(oz_scale_down(x, eb[j]) * up) : 0ll;
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
        *reinterpret_cast<unsigned*>(out + (size_t)t * plane + (size_t)r * ld + 4 * word) = __byte_perm(p01, p23, 0x5410u);
      }
  }
}

}  // namespace hmcb

// gemm.cuh -- fp64 tensor-core GEMM (DMMA via mma.sync.m8n8k4.f64) and CSR SpMM over the
// chain batch, both with fused epilogues.
//
//   Y[M x N] = A[M x K] * B[K x N]
//   A : model matrix (GtG, G or G^T), engine-private, zero padded to M % 128 == 0 and
//       K % 16 == 0, stored tile-major: [M/128][K/16] tiles of 128 rows x 20 doubles (16 values
//       + the 4 padding doubles of the shared-memory layout), see pack in hmcb.cu
//   B : chain batch in the transposed working layout [K x chains], chains contiguous,
//       leading dimension ldb % 128 == 0, padded rows/columns are zero
//   Y is never stored as such: an epilogue functor consumes the accumulator fragments
//       (momentum/position update, residual scaling, or misfit partial sums).
//
// tcgen05 has no f64 kind; the fp64 tensor path on sm_100a is the warp-level DMMA.
// Block tile 128x128x16, 8 warps (2 along M x 4 along N), warp tile 64x32, multi-stage
// pipeline fed by TMA bulk copies completing on mbarriers, padded shared tiles (A: 20
// doubles/row, B: 132 doubles/row) so both fragment loads are bank-conflict free.
#pragma once
#include "common.cuh"

namespace hmcb {

constexpr int GEMM_BM = 128, GEMM_BN = 128, GEMM_BK = 16;
constexpr int GEMM_LDA_S = GEMM_BK + 4;   // 20 doubles
constexpr int GEMM_LDB_S = GEMM_BN + 4;   // 132 doubles
constexpr int GEMM_STAGES = 4;
constexpr int GEMM_GROUP_M = 16;   // row tiles per group of the tile order (L2 reuse)
constexpr int GEMM_THREADS = 256;                      // 8 consumer warps
constexpr int GEMM_LAUNCH_THREADS = GEMM_THREADS + 32;  // + 1 producer warp
constexpr int GEMM_A_STAGE = GEMM_BM * GEMM_LDA_S;  // doubles
constexpr int GEMM_B_STAGE = GEMM_BK * GEMM_LDB_S;
constexpr size_t GEMM_SMEM_BYTES = (size_t)GEMM_STAGES * (GEMM_A_STAGE + GEMM_B_STAGE) * sizeof(double);

// ---- bulk asynchronous copies (the TMA engine, SASS UBLKCP) completing on an mbarrier -----
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(a), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(a), "r"(bytes) : "memory");
}
// (a suspend-time hint on try_wait was measured: no change in run time, and compute-sanitizer's
// racecheck no longer recognised the wait as a synchronisation)
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@!p bra WAIT_LOOP;\n"
      "}\n" ::"r"(a), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* smem, const void* gmem, unsigned bytes, uint64_t* bar) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem);
  const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(d),
               "l"(gmem), "r"(bytes), "r"(b) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(a) : "memory");
}
// barrier among the 256 consumer threads of the GEMM (the producer warp has left by then)
__device__ __forceinline__ void consumer_barrier() { asm volatile("bar.sync 1, 256;\n" ::: "memory"); }
__device__ __forceinline__ void fence_async_proxy() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

__device__ __forceinline__ void dmma_8x8x4(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// Epilogue concept (GEMM side):
//   __device__ void tile(int m0, int n0, int base_m, int base_n, const double (&acc)[8][4][2],
//                        double* smem);
// called once per thread with its 64 accumulators: acc[i][j][h] = Y[base_m + 8 i][base_n + 8 j + h];
// `smem` is the (now idle) pipeline buffer for block-wide reductions.
template <class Epilogue>
// 288 threads are budgeted like 384 by the launch check (168 registers per thread): asking for
// more with __maxnreg__ fails at launch with 'too many resources requested'
__global__ void __launch_bounds__(GEMM_LAUNCH_THREADS, 1)
dmma_gemm_kernel(const double* __restrict__ A, int lda, const double* __restrict__ B, int ldb,
                 int K, Epilogue epi) {
  extern __shared__ __align__(16) double gemm_smem[];
  double* As = gemm_smem;
  double* Bs = gemm_smem + GEMM_STAGES * GEMM_A_STAGE;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 2, wn = warp & 3;          // 2 x 4 warps
  // Tile order: the blocks of a launch are numbered along x first; walking the tiles in that order
  // row by row re-reads the whole chain batch B once per wave (5 GB of DRAM reads for a product whose
  // operands are 0.8 GB).  Grouped order instead: GEMM_GROUP_M row tiles are swept column by column, so
  // a wave of 148 blocks works on ~16 row tiles x ~9 column tiles whose operands stay in L2.
  const int tiles_n = gridDim.x, tiles_m = gridDim.y;
  const int lin = blockIdx.y * tiles_n + blockIdx.x;
  const int group = lin / (GEMM_GROUP_M * tiles_n), first_m = group * GEMM_GROUP_M;
  const int gsz = min(tiles_m - first_m, GEMM_GROUP_M), in_group = lin - group * GEMM_GROUP_M * tiles_n;
  const int tile_m = first_m + in_group % gsz, tile_n = in_group / gsz;
  const int m0 = tile_m * GEMM_BM, n0 = tile_n * GEMM_BN;
  const int ktiles = K / GEMM_BK;

  double acc[8][4][2];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  const int arow = lane >> 2, acol = lane & 3;       // A frag: row = lane/4, k = lane%4
  const int brow = lane & 3, bcol = lane >> 2;       // B frag: k = lane%4, col = lane/4
  auto compute_stage = [&](int stage) {
    const double* as = As + stage * GEMM_A_STAGE + (wm * 64 + arow) * GEMM_LDA_S + acol;
    const double* bs = Bs + stage * GEMM_B_STAGE + brow * GEMM_LDB_S + wn * 32 + bcol;
#pragma unroll
    for (int kk = 0; kk < GEMM_BK; kk += 4) {
      double a[8], b[4];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = as[i * 8 * GEMM_LDA_S + kk];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = bs[kk * GEMM_LDB_S + j * 8];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma_8x8x4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
  };

  // Operand tiles are staged by the bulk asynchronous copy engine (TMA, SASS UBLKCP).  The
  // model matrix A is engine-private, so hmcb_finalize stores it tile-major with the shared-
  // memory padding baked in ([M/128][K/16] tiles of 128 x 20 doubles): one 20 KB copy lands a
  // whole A tile in its conflict-free layout.  B (the chain batch, chains contiguous) takes one
  // 1 KB copy per k-row.  Every stage completes on its own mbarrier (expect_tx = 36 KB).
  // Warp 8 is the producer: it waits for a stage to be released by the 8 consumer warps
  // (empty barrier, one arrival per warp) and refills it; the consumers never meet at a
  // block-wide barrier inside the main loop.
  __shared__ uint64_t full_bar[GEMM_STAGES], empty_bar[GEMM_STAGES];
  if (tid == 0) {
#pragma unroll
    for (int st = 0; st < GEMM_STAGES; ++st) {
      mbar_init(&full_bar[st], 1);
      mbar_init(&empty_bar[st], GEMM_THREADS / 32);
    }
    fence_async_proxy();
  }
  __syncthreads();
  constexpr unsigned kStageBytes = (GEMM_A_STAGE + GEMM_BK * GEMM_BN) * sizeof(double);
  if (warp == GEMM_THREADS / 32) {
    const double* a_tiles = A + (size_t)tile_m * ktiles * GEMM_A_STAGE;
    for (int kt = 0; kt < ktiles; ++kt) {
      const int stage = kt % GEMM_STAGES, use = kt / GEMM_STAGES;
      if (use > 0) mbar_wait(&empty_bar[stage], (unsigned)((use - 1) & 1));
      double* as = As + stage * GEMM_A_STAGE;
      double* bs = Bs + stage * GEMM_B_STAGE;
      if (lane == 0) {
        mbar_expect_tx(&full_bar[stage], kStageBytes);
        bulk_copy_g2s(as, a_tiles + (size_t)kt * GEMM_A_STAGE, GEMM_A_STAGE * sizeof(double), &full_bar[stage]);
      }
      __syncwarp();
      if (lane < GEMM_BK)
        bulk_copy_g2s(bs + lane * GEMM_LDB_S, B + (size_t)(kt * GEMM_BK + lane) * ldb + n0,
                      GEMM_BN * sizeof(double), &full_bar[stage]);
    }
    return;
  }
  for (int kt = 0; kt < ktiles; ++kt) {
    const int stage = kt % GEMM_STAGES;
    mbar_wait(&full_bar[stage], (unsigned)((kt / GEMM_STAGES) & 1));
    compute_stage(stage);
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[stage]);
  }
  consumer_barrier();   // the pipeline buffers are free for the epilogue

  // C frag of sub-tile (i, j): row = base_m + 8 i, columns base_n + 8 j + {0, 1}
  epi.tile(m0, n0, m0 + wm * 64 + (lane >> 2), n0 + wn * 32 + 2 * (lane & 3), acc, gemm_smem);
}

// ------------------------------------------------------------------------ CSR SpMM ---
// Y[i][c] = sum_k data[k] * B[indices[k]][c] for rows i of a CSR matrix.
// lane = chain inside a 32-chain slab, the 4 warps of a block interleave the rows of one row
// chunk.  (col, val) loads are warp-uniform; the gather of a B row is one coalesced 256 B
// segment per warp.  The gathers are the traffic that matters (nnz x chains x 8 B per product,
// 383 GB at config 4) and nothing on chip can hold a slab of B, so the grid is ordered to keep
// them in L2: the chunk index is the fast one and the host sizes the chunks so that the few
// slabs in flight at any time (rows(B) x 256 B each) fit in L2 together.
constexpr int SPMM_THREADS = 128;
constexpr int SPMM_WARPS = SPMM_THREADS / 32;
constexpr int SPMM_SLAB = 32;

template <class Epilogue>
__global__ void __launch_bounds__(SPMM_THREADS)
csr_spmm_kernel(const int* __restrict__ indptr, const int* __restrict__ indices,
                const double* __restrict__ data, int rows, int rows_per_chunk,
                const double* __restrict__ B, int ldb, Epilogue epi) {
  const int chunk = blockIdx.x, warp = threadIdx.x >> 5;
  const int c = blockIdx.y * SPMM_SLAB + (threadIdx.x & 31);
  const int r_begin = chunk * rows_per_chunk;
  const int r_end = min(rows, r_begin + rows_per_chunk);
  const double* Bc = B + c;
  epi.tile_begin(r_begin, blockIdx.y * SPMM_SLAB);
  for (int i = r_begin + warp; i < r_end; i += SPMM_WARPS) {
    const int k0 = __ldg(indptr + i), k1 = __ldg(indptr + i + 1);
    double acc0 = 0.0, acc1 = 0.0;
    int k = k0;
    for (; k + 4 <= k1; k += 4) {
      const int j0 = __ldg(indices + k), j1 = __ldg(indices + k + 1), j2 = __ldg(indices + k + 2),
                j3 = __ldg(indices + k + 3);
      const double v0 = __ldg(data + k), v1 = __ldg(data + k + 1), v2 = __ldg(data + k + 2),
                   v3 = __ldg(data + k + 3);
      const double b0 = Bc[(size_t)j0 * ldb], b1 = Bc[(size_t)j1 * ldb], b2 = Bc[(size_t)j2 * ldb],
                   b3 = Bc[(size_t)j3 * ldb];
      acc0 = fma(v0, b0, acc0); acc1 = fma(v1, b1, acc1); acc0 = fma(v2, b2, acc0); acc1 = fma(v3, b3, acc1);
    }
    for (; k < k1; ++k) acc0 = fma(__ldg(data + k), Bc[(size_t)__ldg(indices + k) * ldb], acc0);
    epi.row(i, c, acc0 + acc1);
  }
  epi.chunk_end(chunk * SPMM_WARPS + warp, c);
}

}  // namespace hmcb

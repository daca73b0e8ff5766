// gemm.cuh -- fp64 tensor-core GEMM (DMMA via mma.sync.m8n8k4.f64) and CSR SpMM over the
// chain batch, both with fused epilogues.
//
//   Y[M x N] = A[M x K] * B[K x N]
//   A : model matrix (GtG, G or G^T), row-major, K contiguous, engine-private, zero padded
//       to M % 128 == 0 and K % 16 == 0
//   B : chain batch in the transposed working layout [K x chains], chains contiguous,
//       leading dimension ldb % 128 == 0, padded rows/columns are zero
//   Y is never stored as such: an epilogue functor consumes the accumulator fragments
//       (momentum/position update, residual scaling, or misfit partial sums).
//
// tcgen05 has no f64 kind; the fp64 tensor path on sm_100a is the warp-level DMMA.
// Block tile 128x128x16, 8 warps (2 along M x 4 along N), warp tile 64x32, 3-stage
// cp.async pipeline, padded shared tiles (A: 20 doubles/row, B: 132 doubles/row) so both
// fragment loads are bank-conflict free.
#pragma once
#include "common.cuh"

namespace hmcb {

constexpr int GEMM_BM = 128, GEMM_BN = 128, GEMM_BK = 16;
constexpr int GEMM_LDA_S = GEMM_BK + 4;   // 20 doubles
constexpr int GEMM_LDB_S = GEMM_BN + 4;   // 132 doubles
constexpr int GEMM_STAGES = 3;
constexpr int GEMM_THREADS = 256;
constexpr int GEMM_A_STAGE = GEMM_BM * GEMM_LDA_S;  // doubles
constexpr int GEMM_B_STAGE = GEMM_BK * GEMM_LDB_S;
constexpr size_t GEMM_SMEM_BYTES = (size_t)GEMM_STAGES * (GEMM_A_STAGE + GEMM_B_STAGE) * sizeof(double);

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

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
__global__ void __launch_bounds__(GEMM_THREADS, 1)
dmma_gemm_kernel(const double* __restrict__ A, int lda, const double* __restrict__ B, int ldb,
                 int K, Epilogue epi) {
  extern __shared__ __align__(16) double gemm_smem[];
  double* As = gemm_smem;
  double* Bs = gemm_smem + GEMM_STAGES * GEMM_A_STAGE;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 2, wn = warp & 3;          // 2 x 4 warps
  const int m0 = blockIdx.y * GEMM_BM, n0 = blockIdx.x * GEMM_BN;
  const int ktiles = K / GEMM_BK;

  auto load_stage = [&](int stage, int kt) {
    const int k0 = kt * GEMM_BK;
    double* as = As + stage * GEMM_A_STAGE;
    double* bs = Bs + stage * GEMM_B_STAGE;
#pragma unroll
    for (int i = 0; i < 4; ++i) {                    // A: 128 rows x 8 chunks of 2 doubles
      const int chunk = tid + i * GEMM_THREADS;
      const int r = chunk >> 3, cc = (chunk & 7) * 2;
      cp_async16(as + r * GEMM_LDA_S + cc, A + (size_t)(m0 + r) * lda + k0 + cc);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {                    // B: 16 rows x 64 chunks
      const int chunk = tid + i * GEMM_THREADS;
      const int r = chunk >> 6, cc = (chunk & 63) * 2;
      cp_async16(bs + r * GEMM_LDB_S + cc, B + (size_t)(k0 + r) * ldb + n0 + cc);
    }
  };

  double acc[8][4][2];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
  for (int s = 0; s < GEMM_STAGES - 1; ++s) {
    if (s < ktiles) load_stage(s, s);
    cp_async_commit();
  }

  const int arow = lane >> 2, acol = lane & 3;       // A frag: row = lane/4, k = lane%4
  const int brow = lane & 3, bcol = lane >> 2;       // B frag: k = lane%4, col = lane/4
  for (int kt = 0; kt < ktiles; ++kt) {
    cp_async_wait<GEMM_STAGES - 2>();
    __syncthreads();
    {
      const int nk = kt + GEMM_STAGES - 1;
      if (nk < ktiles) load_stage(nk % GEMM_STAGES, nk);
      cp_async_commit();
    }
    const double* as = As + (kt % GEMM_STAGES) * GEMM_A_STAGE + (wm * 64 + arow) * GEMM_LDA_S + acol;
    const double* bs = Bs + (kt % GEMM_STAGES) * GEMM_B_STAGE + brow * GEMM_LDB_S + wn * 32 + bcol;
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
  }
  cp_async_wait<0>();
  __syncthreads();

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

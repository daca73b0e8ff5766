// launch_ozaki.cu -- launchers of the int8-sliced (Ozaki) dense products on tcgen05 (ozaki.cuh).
#include <cudaTypedefs.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "launch.cuh"
#include "ozaki.cuh"
#include "ozaki_sparse.cuh"

namespace hmcb {

static PFN_cuTensorMapEncodeTiled_v12000 tensor_map_encoder() {
  static PFN_cuTensorMapEncodeTiled_v12000 encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  }
  return encode;
}

// slices [S][rows x K] int8, K contiguous -> 3-D tensor {K, rows, S}, box {128, box_rows, 1}, 128-byte swizzle
cudaError_t ozaki_slice_map(const signed char* base, long long K, long long rows, int slices, int box_rows,
                            CUtensorMap* out) {
  PFN_cuTensorMapEncodeTiled_v12000 encode = tensor_map_encoder();
  if (!encode) return cudaErrorNotSupported;
  const cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)slices};
  const cuuint64_t strides[2] = {(cuuint64_t)K, (cuuint64_t)K * (cuuint64_t)rows};
  const cuuint32_t box[3] = {(cuuint32_t)OZ_BK, (cuuint32_t)box_rows, 1u};
  const cuuint32_t estr[3] = {1u, 1u, 1u};
  const CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<signed char*>(base), dims, strides, box,
                            estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

// order planes [orders][M x ldc] int32 -> 3-D tensor {ldc, M, orders}, box {32, 32, 1}, 128-byte swizzle: what one
// epilogue warp stores at a time
cudaError_t ozaki_plane_map(const int* base, long long ldc, long long M, int orders, long long plane_stride,
                            CUtensorMap* out) {
  PFN_cuTensorMapEncodeTiled_v12000 encode = tensor_map_encoder();
  if (!encode) return cudaErrorNotSupported;
  const cuuint64_t dims[3] = {(cuuint64_t)ldc, (cuuint64_t)M, (cuuint64_t)orders};
  const cuuint64_t strides[2] = {(cuuint64_t)ldc * 4u, (cuuint64_t)plane_stride * 4u};
  const cuuint32_t box[3] = {32u, 32u, 1u};
  const cuuint32_t estr[3] = {1u, 1u, 1u};
  const CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_INT32, 3, const_cast<int*>(base), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

cudaError_t ozaki_init() {
  cudaError_t e = cudaFuncSetAttribute(i8_gemm_groups_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)OZ_SMEM_BYTES);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(i8_gemm_groups_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)OZ_SMEM_BYTES);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(i8_gemm_groups_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)OZ_SMEM_BYTES);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(i8_gemm_groups_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)OZ_SMEM_BYTES);
  return e;
}

// CTA pairs (tcgen05 cta_group::2) unless HMCB_OZAKI_PAIR=0
bool ozaki_pair_mode() {
  const char* v = std::getenv("HMCB_OZAKI_PAIR");
  return v ? std::atoi(v) != 0 : true;
}

// The dataflow program of the order group {orders[0], orders[1]}: walking the B digits t upward, B_t meets
// A_(o - t) for each order o of the group; a tile is loaded right before its first product and its ring
// slot is released by the last product that reads it.
static bool build_program(int SA, int SB, const int* orders, int n_acc, OzProgram* P) {
  std::memset(P, 0, sizeof(*P));
  P->n_acc = n_acc;
  int a_idx[OZ_MAX_SLICES], last_a[OZ_MAX_OPS], last_b[OZ_MAX_OPS];
  bool acc_seen[OZ_MAX_ACC] = {false, false};
  for (int s = 0; s < OZ_MAX_SLICES; ++s) a_idx[s] = -1;
  for (int a = 0; a < n_acc; ++a) P->order[a] = orders[a];
  for (int t = 0; t < SB; ++t) {
    int b_idx = -1;
    for (int a = 0; a < n_acc; ++a) {
      const int s = orders[a] - t;
      if (s < 0 || s >= SA) continue;
      if (P->n_loads + 2 > OZ_MAX_OPS || P->n_mma >= OZ_MAX_OPS) return false;
      if (b_idx < 0) {
        b_idx = P->nB++;
        P->load_is_b[P->n_loads] = 1; P->load_slice[P->n_loads++] = (unsigned char)t;
      }
      if (a_idx[s] < 0) {
        a_idx[s] = P->nA++;
        P->load_is_b[P->n_loads] = 0; P->load_slice[P->n_loads++] = (unsigned char)s;
      }
      const int m = P->n_mma++;
      P->mma_a[m] = (unsigned char)a_idx[s]; P->mma_b[m] = (unsigned char)b_idx; P->mma_acc[m] = (unsigned char)a;
      P->mma_flags[m] = acc_seen[a] ? 0 : 4;
      acc_seen[a] = true;
      last_a[a_idx[s]] = m; last_b[b_idx] = m;
    }
  }
  for (int i = 0; i < P->nA; ++i) P->mma_flags[last_a[i]] |= 1;
  for (int i = 0; i < P->nB; ++i) P->mma_flags[last_b[i]] |= 2;
  // the A ring must hold the tiles in flight between a load and its release: every product may only wait
  // for tiles that precede, in load order, the tiles the producer can be blocked on
  return P->n_mma > 0 && P->nA > 0 && P->nB > 0;
}

// orders 0 .. orders-1 in groups of two consecutive orders (the last one alone when `orders` is odd),
// heaviest group first
bool oz_build_plan(int SA, int SB, int orders, long long M, long long N, OzPlan* plan) {
  std::memset(plan, 0, sizeof(*plan));
  if (SA < 1 || SB < 1 || SA > OZ_MAX_SLICES || SB > OZ_MAX_SLICES || orders < 1 || orders > OZ_MAX_ORDERS ||
      orders > SA + SB - 1)
    return false;
  plan->tiles_m = (int)(M / OZ_BM);
  plan->tiles_n = (int)((N + OZ_BN - 1) / OZ_BN);
  for (int o = orders - 1; o >= 0; o -= 2) {
    if (plan->n_groups >= OZ_MAX_GROUPS) return false;
    int ord[2] = {o > 0 ? o - 1 : 0, o};
    const int n_acc = o > 0 ? 2 : 1;
    if (!build_program(SA, SB, n_acc == 2 ? ord : &ord[1], n_acc, &plan->g[plan->n_groups++])) return false;
  }
  std::stable_sort(plan->g, plan->g + plan->n_groups,
                   [](const OzProgram& a, const OzProgram& b) { return a.n_mma > b.n_mma; });
  return true;
}

template <bool RESIDUES>
static cudaError_t launch_i8_plan(const OzPlan& plan, const CUtensorMap& mapA, const CUtensorMap& mapB,
                                  const CUtensorMap& mapBh, const CUtensorMap& mapC, long long M, long long K, int ldc,
                                  signed char* res, long long res_plane, cudaStream_t s) {
  const bool pair = ozaki_pair_mode();
  const long long rows_m = pair ? (plan.tiles_m + 1) / 2 : plan.tiles_m;
  const long long ctas = (long long)plan.n_groups * rows_m * plan.tiles_n * (pair ? 2 : 1);
  if (ctas <= 0 || ctas > 0x7fffffffll) return cudaErrorInvalidValue;
  if (!pair) {
    i8_gemm_groups_kernel<false, RESIDUES><<<(unsigned)ctas, OZ_THREADS, OZ_SMEM_BYTES, s>>>(
        mapA, mapB, mapC, plan, (int)(K / OZ_BK), ldc, (int)M, res, res_plane);
    return cudaGetLastError();
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)ctas);
  cfg.blockDim = dim3(OZ_THREADS);
  cfg.dynamicSmemBytes = OZ_SMEM_BYTES;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, i8_gemm_groups_kernel<true, RESIDUES>, mapA, mapBh, mapC, plan, (int)(K / OZ_BK), ldc,
                            (int)M, res, res_plane);
}

// C[o] (o < orders) = sum_{s+t=o} A_s B_t^T;  M % 128 == 0, K % 128 == 0, ldc = padded N (% 128 == 0);
// mapB: box of 256 rows (one CTA per tile), mapBh: box of 128 rows (CTA pairs: each CTA loads half a B tile);
// mapC = ozaki_plane_map of the order planes
cudaError_t launch_i8_gemm_orders(const CUtensorMap& mapA, const CUtensorMap& mapB, const CUtensorMap& mapBh,
                                  const CUtensorMap& mapC, long long M, long long N, long long K, int SA, int SB,
                                  int orders, int ldc, cudaStream_t s) {
  if (M % OZ_BM || K % OZ_BK || N % 128) return cudaErrorInvalidValue;
  OzPlan plan;
  if (!oz_build_plan(SA, SB, orders, M, N, &plan)) return cudaErrorInvalidValue;
  return launch_i8_plan<false>(plan, mapA, mapB, mapBh, mapC, M, K, ldc, nullptr, 0, s);
}

// Modular products: res[m] = (A_m B_m^T) mod p_m for the OZ_NMOD residue planes of A (mapA: {K, M, OZ_NMOD}) and
// B; two moduli per CTA (two accumulators; no tile is shared between moduli).  res: [OZ_NMOD][M x ldc] int8.
cudaError_t launch_i8_gemm_moduli(const CUtensorMap& mapA, const CUtensorMap& mapB, const CUtensorMap& mapBh,
                                  long long M, long long N, long long K, int ldc, signed char* res, cudaStream_t s) {
  if (M % OZ_BM || K % OZ_BK || N % 128 || K * 128 * 128 >= (1ll << 31)) return cudaErrorInvalidValue;
  OzPlan plan;
  std::memset(&plan, 0, sizeof(plan));
  plan.tiles_m = (int)(M / OZ_BM);
  plan.tiles_n = (int)((N + OZ_BN - 1) / OZ_BN);
  for (int m = 0; m < OZ_NMOD; m += 2) {
    OzProgram& P = plan.g[plan.n_groups++];
    P.n_acc = m + 1 < OZ_NMOD ? 2 : 1;
    for (int a = 0; a < P.n_acc; ++a) {
      P.order[a] = m + a;                                   // the plane / modulus of accumulator a
      P.load_is_b[P.n_loads] = 1; P.load_slice[P.n_loads++] = (unsigned char)(m + a);
      P.load_is_b[P.n_loads] = 0; P.load_slice[P.n_loads++] = (unsigned char)(m + a);
      P.mma_a[P.n_mma] = (unsigned char)a; P.mma_b[P.n_mma] = (unsigned char)a; P.mma_acc[P.n_mma] = (unsigned char)a;
      P.mma_flags[P.n_mma++] = 1 | 2 | 4;                   // last use of both tiles, first product of its accumulator
    }
    P.nA = P.nB = P.n_acc;
  }
  CUtensorMap unused;
  std::memset(&unused, 0, sizeof(unused));
  return launch_i8_plan<true>(plan, mapA, mapB, mapBh, unused, M, K, ldc, res, M * (long long)ldc, s);
}

cudaError_t launch_oz_colmax(const double* X, int rows, int ld, unsigned long long* maxbits, cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(maxbits, 0, (size_t)ld * sizeof(unsigned long long), s);
  if (e != cudaSuccess) return e;
  oz_colmax_kernel<<<dim3(ld / 128, (rows + 63) / 64), 128, 0, s>>>(X, rows, ld, maxbits);
  return cudaGetLastError();
}

cudaError_t launch_oz_slice_chains(const double* X, int K, int ld, int SB, const unsigned long long* maxbits,
                                   signed char* out, cudaStream_t s) {
  if (K % 128 || ld % 32 || SB < 1 || SB > OZ_MAX_SLICES) return cudaErrorInvalidValue;
  oz_slice_chains_kernel<<<dim3(ld / 32, K / 128), 256, 0, s>>>(X, K, ld, SB, maxbits, out);
  return cudaGetLastError();
}

cudaError_t launch_oz_combine_residual(const int* C, long long plane_stride, int rows, int ld, int orders,
                                       const int* ea, const unsigned long long* maxbits_in, const ResidualEpi& epi,
                                       unsigned long long* maxbits_out, cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(maxbits_out, 0, (size_t)ld * sizeof(unsigned long long), s);
  if (e != cudaSuccess) return e;
  const dim3 grid(ld / 128, (rows + 15) / 16);
  if (orders == 6)
    oz_combine_kernel<ResidualEpi, true, 16, 6><<<grid, 128, 0, s>>>(C, plane_stride, rows, ld, orders, ea, maxbits_in,
                                                                      epi, maxbits_out);
  else
    oz_combine_kernel<ResidualEpi, true, 16, 0><<<grid, 128, 0, s>>>(C, plane_stride, rows, ld, orders, ea, maxbits_in,
                                                                      epi, maxbits_out);
  return cudaGetLastError();
}

cudaError_t launch_oz_combine_update(const int* C, long long plane_stride, int rows, int ld, int orders, const int* ea,
                                     const unsigned long long* maxbits_in, const UpdateEpi& epi, cudaStream_t s) {
  const dim3 grid(ld / 128, (rows + 15) / 16);
  if (orders == 6)
    oz_combine_kernel<UpdateEpi, false, 16, 6><<<grid, 128, 0, s>>>(C, plane_stride, rows, ld, orders, ea, maxbits_in,
                                                                     epi, nullptr);
  else
    oz_combine_kernel<UpdateEpi, false, 16, 0><<<grid, 128, 0, s>>>(C, plane_stride, rows, ld, orders, ea, maxbits_in,
                                                                     epi, nullptr);
  return cudaGetLastError();
}

// per-chain misfit partial sums: one partial per 128-row chunk, like the tiles of the DMMA GEMM (epi.part
// holds [rows / 128 x ld])
cudaError_t launch_oz_combine_misfit(const int* C, long long plane_stride, int rows, int ld, int orders, const int* ea,
                                     const unsigned long long* maxbits_in, const MisfitEpi& epi, cudaStream_t s) {
  const dim3 grid(ld / 128, (rows + 127) / 128);
  if (orders == 6)
    oz_combine_kernel<MisfitEpi, false, 128, 6><<<grid, 128, 0, s>>>(C, plane_stride, rows, ld, orders, ea, maxbits_in,
                                                                      epi, nullptr);
  else
    oz_combine_kernel<MisfitEpi, false, 128, 0><<<grid, 128, 0, s>>>(C, plane_stride, rows, ld, orders, ea, maxbits_in,
                                                                      epi, nullptr);
  return cudaGetLastError();
}

// Host side of the slicing: balanced radix-256 digits of a row-major matrix [rows x cols], zero padded to
// [rows_pad x cols_pad], S digit planes; a[i][k] = 2^ea[i] * sum_s slice_s[i][k] 256^-(s+1) + rounding.
// Returns max over the rows of (sum_k |rounding|) / (sum_k |a[i][k]|).
double oz_slice_rows_host(const double* A, long long rows, long long cols, long long rows_pad, long long cols_pad,
                          int S, signed char* slices, int* ea) {
  double worst = 0.0;
  if (slices) std::memset(slices, 0, (size_t)S * rows_pad * cols_pad);
  std::memset(ea, 0, sizeof(int) * (size_t)rows_pad);
  const double up = std::ldexp(1.0, OZ_BITS * S);
  for (long long i = 0; i < rows; ++i) {
    const double* a = A + (size_t)i * cols;
    double m = 0.0, sum = 0.0, err = 0.0;
    for (long long k = 0; k < cols; ++k) { m = std::max(m, std::fabs(a[k])); sum += std::fabs(a[k]); }
    unsigned long long bits;
    std::memcpy(&bits, &m, sizeof(bits));
    const int e = oz_exponent(bits);
    if (e == INT_MIN || !(m > 0.0)) continue;   // non-finite rows are refused by the caller
    ea[i] = e;
    for (long long k = 0; k < cols; ++k) {
      const double y = oz_scale_down(a[k], e) * up;   // exact
      const double r = std::nearbyint(y);
      err += std::fabs(y - r);
      if (slices) {
        signed char digit[OZ_MAX_SLICES];
        oz_digits((long long)r, S, digit);
        for (int s = 0; s < S; ++s) slices[((size_t)s * rows_pad + i) * cols_pad + k] = digit[s];
      }
    }
    if (sum > 0.0) worst = std::max(worst, std::ldexp(err, e - OZ_BITS * S) / sum);
  }
  return worst;
}

}  // namespace hmcb

// Test / measurement entry: exact int8 slice products on the tcgen05 tensor cores.
extern "C" int hmcb_debug_i8_gemm(int device, int64_t M, int64_t N, int64_t K, int SA, int SB, int orders,
                                  const signed char* A, const signed char* B, int32_t* C, void* stream) {
  using namespace hmcb;
  if (!A || !B || !C || M <= 0 || N <= 0 || K <= 0 || SA < 1 || SB < 1) return -1;
  if (cudaSetDevice(device) != cudaSuccess) return -1;
  if (ozaki_init() != cudaSuccess) return -2;
  CUtensorMap mapA, mapB, mapBh, mapC;
  if (ozaki_slice_map(A, K, M, SA, OZ_BM, &mapA) != cudaSuccess) return -3;
  if (ozaki_slice_map(B, K, N, SB, OZ_BN, &mapB) != cudaSuccess) return -3;
  if (ozaki_slice_map(B, K, N, SB, OZ_BN / 2, &mapBh) != cudaSuccess) return -3;
  if (ozaki_plane_map(C, N, M, orders, M * N, &mapC) != cudaSuccess) return -3;
  if (launch_i8_gemm_orders(mapA, mapB, mapBh, mapC, M, N, K, SA, SB, orders, (int)N,
                            static_cast<cudaStream_t>(stream)) != cudaSuccess)
    return -4;
  return 0;
}

// Test entry (host only): the digits oz_slice_rows_host gives a matrix, and its representation error.
extern "C" double hmcb_debug_oz_slice_rows(const double* A, int64_t rows, int64_t cols, int S, signed char* slices,
                                           int32_t* ea) {
  if (!A || !ea || rows <= 0 || cols <= 0 || S < 1 || S > hmcb::OZ_MAX_SLICES) return -1.0;
  return hmcb::oz_slice_rows_host(A, rows, cols, rows, cols, S, slices, ea);
}

// ---- modular variant: host side -----------------------------------------------------------------------
namespace hmcb {

static const int kOzMod[OZ_NMOD] = {256, 255, 253, 251, 247, 241, 239, 233, 229, 227, 223, 217, 211};

// the reconstruction constants: f_m = y_m / p_m in three 40-bit chunks, y_m = (P / p_m)^-1 mod p_m; P as a double
OzCrt oz_crt_constants() {
  OzCrt c;
  c.P = 1.0;
  for (int m = 0; m < OZ_NMOD; ++m) c.P *= (double)kOzMod[m];
  for (int m = 0; m < OZ_NMOD; ++m) {
    const int p = kOzMod[m];
    int Mm = 1;
    for (int j = 0; j < OZ_NMOD; ++j)
      if (j != m) Mm = (int)(((long long)Mm * (kOzMod[j] % p)) % p);
    int y = 1;
    while ((Mm * y) % p != 1) ++y;
    long long rem = y;
    for (int j = 0; j < 3; ++j) {
      const long long num = rem << 40;          // rem < 256
      c.F[m][j] = (double)(num / p);
      rem = num % p;
    }
  }
  return c;
}

static int sym_mod(long long X, int p) {
  int r = (int)(X % p);
  if (r < 0) r += p;
  return r > 127 ? r - p : r;
}

// Residue planes of a row-major HOST matrix [rows x cols], zero padded to [rows_pad x cols_pad]: every row scaled
// to 44-bit integers (exponent ea[i] as in oz_slice_rows_host), res[m][i][k] = that integer modulo p_m in [-128, 127].
void oz_residue_rows_host(const double* A, long long rows, long long cols, long long rows_pad, long long cols_pad,
                          signed char* res, int* ea) {
  std::memset(res, 0, (size_t)OZ_NMOD * rows_pad * cols_pad);
  std::memset(ea, 0, sizeof(int) * (size_t)rows_pad);
  const double up = std::ldexp(1.0, OZ_CRT_BITS);
  for (long long i = 0; i < rows; ++i) {
    const double* a = A + (size_t)i * cols;
    double mx = 0.0;
    for (long long k = 0; k < cols; ++k) mx = std::max(mx, std::fabs(a[k]));
    unsigned long long bits;
    std::memcpy(&bits, &mx, sizeof(bits));
    const int e = oz_exponent(bits);
    if (e == INT_MIN || !(mx > 0.0)) continue;
    ea[i] = e;
    for (long long k = 0; k < cols; ++k) {
      const long long X = (long long)std::nearbyint(oz_scale_down(a[k], e) * up);
      for (int m = 0; m < OZ_NMOD; ++m) res[((size_t)m * rows_pad + i) * cols_pad + k] = (signed char)sym_mod(X, kOzMod[m]);
    }
  }
}

cudaError_t launch_oz_residue_chains(const double* X, int K, int ld, const unsigned long long* maxbits, signed char* out,
                                     cudaStream_t s) {
  if (K % 128 || ld % 16) return cudaErrorInvalidValue;
  oz_residue_chains_kernel<<<dim3(ld / 16, K / 128), 256, 0, s>>>(X, K, ld, maxbits, out);
  return cudaGetLastError();
}

cudaError_t launch_oz_crt_update(const signed char* res, long long plane, int rows, int ld, const int* ea,
                                 const unsigned long long* maxbits_in, const UpdateEpi& epi, cudaStream_t s) {
  static const OzCrt crt = oz_crt_constants();
  oz_crt_combine_kernel<UpdateEpi, 16><<<dim3((ld / 4 + 127) / 128, (rows + 15) / 16), 128, 0, s>>>(
      res, plane, rows, ld, ea, maxbits_in, crt, epi);
  return cudaGetLastError();
}

struct RawStoreEpi {     // Y -> out[i][c] (tests)
  double* out; int ld;
  __device__ __forceinline__ void row(int i, int c, double y) const { out[(size_t)i * ld + c] = y; }
};

}  // namespace hmcb

// Test entry of the modular variant: Y = A X for HOST A [M x K] (row-major) and DEVICE chain batch X [K x N]
// (chains contiguous): residues of A on the host, of X on the device, 13 modular int8 products on tcgen05,
// Chinese-remainder reconstruction -> DEVICE Y [M x N] fp64.  M, K % 128 == 0, N % 128 == 0.
extern "C" int hmcb_debug_crt_product(int device, int64_t M, int64_t N, int64_t K, const double* A_host,
                                      const double* X_dev, double* Y_dev, void* stream) {
  using namespace hmcb;
  if (!A_host || !X_dev || !Y_dev || M <= 0 || N <= 0 || K <= 0 || M % 128 || K % 128 || N % 128) return -1;
  if (cudaSetDevice(device) != cudaSuccess) return -1;
  if (ozaki_init() != cudaSuccess) return -2;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  std::vector<signed char> resA((size_t)OZ_NMOD * M * K);
  std::vector<int> ea((size_t)M);
  oz_residue_rows_host(A_host, M, K, M, K, resA.data(), ea.data());
  signed char *dA = nullptr, *dB = nullptr, *dC = nullptr;
  int* dea = nullptr;
  unsigned long long* dmax = nullptr;
  int rc = 0;
  if (cudaMalloc(&dA, resA.size()) != cudaSuccess || cudaMalloc(&dB, (size_t)OZ_NMOD * N * K) != cudaSuccess ||
      cudaMalloc(&dC, (size_t)OZ_NMOD * M * N) != cudaSuccess || cudaMalloc(&dea, sizeof(int) * M) != cudaSuccess ||
      cudaMalloc(&dmax, sizeof(unsigned long long) * N) != cudaSuccess)
    rc = -3;
  CUtensorMap mapA, mapB, mapBh;
  if (!rc && (cudaMemcpyAsync(dA, resA.data(), resA.size(), cudaMemcpyHostToDevice, s) != cudaSuccess ||
              cudaMemcpyAsync(dea, ea.data(), sizeof(int) * M, cudaMemcpyHostToDevice, s) != cudaSuccess))
    rc = -3;
  if (!rc && (ozaki_slice_map(dA, K, M, OZ_NMOD, OZ_BM, &mapA) != cudaSuccess ||
              ozaki_slice_map(dB, K, N, OZ_NMOD, OZ_BN, &mapB) != cudaSuccess ||
              ozaki_slice_map(dB, K, N, OZ_NMOD, OZ_BN / 2, &mapBh) != cudaSuccess))
    rc = -3;
  if (!rc && (launch_oz_colmax(X_dev, (int)K, (int)N, dmax, s) != cudaSuccess ||
              launch_oz_residue_chains(X_dev, (int)K, (int)N, dmax, dB, s) != cudaSuccess ||
              launch_i8_gemm_moduli(mapA, mapB, mapBh, M, N, K, (int)N, dC, s) != cudaSuccess))
    rc = -4;
  if (!rc) {
    static const OzCrt crt = oz_crt_constants();
    RawStoreEpi st{Y_dev, (int)N};
    oz_crt_combine_kernel<RawStoreEpi, 16><<<dim3((unsigned)((N / 4 + 127) / 128), (unsigned)((M + 15) / 16)), 128, 0, s>>>(
        dC, M * N, (int)M, (int)N, dea, dmax, crt, st);
    if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(s) != cudaSuccess) rc = -5;
  }
  cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(dea); cudaFree(dmax);
  return rc;
}

// ---- gathered (block-sparse) slice products (ozaki_sparse.cuh) ------------------------------------------
namespace hmcb {

cudaError_t ozaki_sparse_init() {
  return cudaFuncSetAttribute(i8_gather_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)OZ_SMEM_BYTES);
}

// C[o] (o < orders), rows = bundles x 128: the slice products of every bundle's dense tile (mapA: the digits of
// all bundles end to end on the K axis, 128 rows) with the B rows its list names, gathered from the digit
// planes Bg [SB][rows_b x ld] (chains contiguous).  ld % 128 == 0.
cudaError_t launch_i8_gather_gemm(const CUtensorMap& mapA, const CUtensorMap& mapC, int n_bundles, int SA, int SB,
                                  int orders, const OzBundle* bundles, const int* list, const signed char* Bg,
                                  long long bg_plane, int ld, cudaStream_t s) {
  if (ld % 128 || n_bundles < 1) return cudaErrorInvalidValue;
  OzPlan plan;
  if (!oz_build_plan(SA, SB, orders, (long long)n_bundles * OZ_BM, ld, &plan)) return cudaErrorInvalidValue;
  const long long ctas = (long long)plan.n_groups * plan.tiles_m * plan.tiles_n;
  if (ctas <= 0 || ctas > 0x7fffffffll) return cudaErrorInvalidValue;
  // column tiles per panel: the digit planes of a panel's chains (all B rows x 256 chains x SB digits per tile)
  // should stay in L2 next to the model tiles and the order planes streaming through
  const char* pv = std::getenv("HMCB_OZAKI_SPARSE_PANEL");
  const int panel = std::max(1, pv ? std::atoi(pv) : 3);
  i8_gather_gemm_kernel<<<(unsigned)ctas, OZS_THREADS, OZ_SMEM_BYTES, s>>>(mapA, mapC, plan, bundles, list, Bg, bg_plane,
                                                                            ld, ld, panel);
  return cudaGetLastError();
}

cudaError_t launch_oz_slice_plain(const double* X, int K, int ld, int SB, const unsigned long long* maxbits,
                                  signed char* out, long long plane, cudaStream_t s) {
  if (ld % 128 || SB < 1 || SB > OZ_MAX_SLICES) return cudaErrorInvalidValue;
  oz_slice_plain_kernel<<<dim3((ld / 4 + 255) / 256, (K + 7) / 8), 256, 0, s>>>(X, K, ld, SB, maxbits, out, plane);
  return cudaGetLastError();
}

}  // namespace hmcb

// Test / measurement entry: gathered int8 slice products.  A: [SA][128][Ktot] digits of the bundles' tiles,
// bundles: n_bundles x {koff, kblocks}, list: [Ktot] B-row index per list entry, B: [SB][rows_b][N] digit
// planes (chains contiguous), C: [orders][n_bundles * 128][N].
extern "C" int hmcb_debug_i8_gather_gemm(int device, int64_t n_bundles, int64_t N, int64_t Ktot, int64_t rows_b, int SA,
                                         int SB, int orders, const signed char* A, const int32_t* bundles,
                                         const int32_t* list, const signed char* B, int32_t* C, void* stream) {
  using namespace hmcb;
  if (!A || !B || !C || !bundles || !list || n_bundles <= 0 || N <= 0 || Ktot <= 0 || Ktot % OZ_BK) return -1;
  if (cudaSetDevice(device) != cudaSuccess) return -1;
  if (ozaki_sparse_init() != cudaSuccess) return -2;
  CUtensorMap mapA, mapC;
  if (ozaki_slice_map(A, Ktot, OZ_BM, SA, OZ_BM, &mapA) != cudaSuccess) return -3;
  if (ozaki_plane_map(C, N, n_bundles * OZ_BM, orders, n_bundles * OZ_BM * N, &mapC) != cudaSuccess) return -3;
  static_assert(sizeof(OzBundle) == 2 * sizeof(int32_t), "OzBundle layout");
  if (launch_i8_gather_gemm(mapA, mapC, (int)n_bundles, SA, SB, orders, reinterpret_cast<const OzBundle*>(bundles), list, B,
                            rows_b * N, (int)N, static_cast<cudaStream_t>(stream)) != cudaSuccess)
    return -4;
  return 0;
}

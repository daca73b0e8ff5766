// staged.cuh -- HMC proposal pipeline for targets with a coupled (matrix) likelihood:
// dense LinearMatrix (premultiplied GtG or direct G / G^T form) and CSR LinearMatrix.
//
// Batching chains turns every gradient into a matrix product over the chain batch, so the
// trajectory is a sequence of launches; the momentum + position update, the bounds
// reflection and the prior gradient are fused into the GEMM / SpMM epilogue.
//
// Working layout: transposed [dims x chains] (chains contiguous, leading dimension ld a
// multiple of 128, rows padded to dpad); the chain-major API tensors are transposed once
// per block of proposals.  Per-chain reductions are two-pass with fixed-order partial
// sums (deterministic).
#pragma once
#include "common.cuh"
#include "gemm.cuh"

namespace hmcb {

constexpr int ST_THREADS = 128;  // chains per block of the elementwise kernels
constexpr int ST_DT = 32;        // coordinates per block of the elementwise kernels

enum { LIK_NONE = 0, LIK_PREMULT = 1, LIK_DIRECT = 2 };

// ------------------------------------------------------------------- epilogues ---

// Likelihood gradient Y -> total gradient -> p -= b*eps*g ; q += a*eps*dK/dp ; reflect.
struct UpdateEpi {
  static constexpr bool kPerChainSum = false;
  DevTarget T;
  int C, ld;
  const double* sub;       // [d] subtracted from Y (Gtd0 for the premultiplied form) or null
  const double* q_in;
  double* q_out;
  double* p;
  const double* eps;       // [ld]
  double b_mult, a_mult;
  const unsigned* flags_in;  // bound violations of q_in per chain, or null
  unsigned* flags_out;       // bound violations of q_out per chain, or null
  double* trace_q;           // chain-major [C x d] slices for this gradient evaluation, or null
  double* trace_g;
  int grad_only;             // 1: store the total gradient into p and stop
  int momentum_only;         // 1: p -= b*eps*g only (Full mass matrix: the position update needs M^-1 p,
                             //    a product over all coordinates, and follows as its own launches)

  __device__ __forceinline__ void apply(int j, int c, double y) const {
    if (j >= T.dims || c >= C) return;
    const size_t o = (size_t)j * ld + c;
    const double lik = sub ? __dsub_rn(y, __ldg(sub + j)) : y;
    double q = q_in[o];
    const unsigned oob = flags_in ? flags_in[c] : 0u;
    const double g = __dadd_rn(prior_gradient(T, j, q, oob), lik);
    if (grad_only) { p[o] = g; return; }
    if (trace_q) {
      trace_q[(size_t)c * T.dims + j] = q;
      trace_g[(size_t)c * T.dims + j] = g;
    }
    const double e = eps[c];
    double pp = p[o];
    momentum_update(__dmul_rn(b_mult, e), g, pp);
    if (momentum_only) { p[o] = pp; return; }
    position_update(T, j, __dmul_rn(a_mult, e), q, pp);
    p[o] = pp;
    q_out[o] = q;
    if (flags_out) {
      const unsigned m = bound_violations(T, j, q);
      if (m) atomicOr(flags_out + c, m);
    }
  }
  // GEMM interface: the thread's fragment is 8 rows x 4 column pairs.  Everything that
  // depends only on the coordinate (prior slot, Gtd0, mass, reflection bounds) is loaded once
  // per row, everything that depends only on the chain (eps, bound flags) once per column
  // pair, and q / p move as 16-byte pairs.
  __device__ __forceinline__ void tile(int, int, int base_m, int base_n, const double (&acc)[8][4][2],
                                       double*) const {
    if (T.n_terms > 1 || grad_only || trace_q || momentum_only) {   // generic element-wise path
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          apply(base_m + 8 * i, base_n + 8 * j, acc[i][j][0]);
          apply(base_m + 8 * i, base_n + 8 * j + 1, acc[i][j][1]);
        }
      return;
    }
    double2 e2[4];
    uint2 fl[4];
    bool col_ok[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = base_n + 8 * j;
      col_ok[j] = c < C;
      e2[j] = col_ok[j] ? *reinterpret_cast<const double2*>(eps + c) : make_double2(0.0, 0.0);
      fl[j] = (flags_in && col_ok[j]) ? *reinterpret_cast<const uint2*>(flags_in + c) : make_uint2(0u, 0u);
    }
    const bool has_refl = T.refl_lb != nullptr || T.refl_ub != nullptr;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int m = base_m + 8 * i;
      if (m >= T.dims) continue;
      const int kind = T.n_terms ? (int)__ldg(T.t_kind + m) : TERM_NONE;
      const double ta = T.n_terms ? __ldg(T.t_a + m) : 0.0, tb = T.n_terms ? __ldg(T.t_b + m) : 0.0;
      const double sub_m = sub ? __ldg(sub + m) : 0.0;
      const double im = T.invm ? __ldg(T.invm + m) : 1.0;
      const double lb = T.refl_lb ? __ldg(T.refl_lb + m) : -CUDART_INF;
      const double ub = T.refl_ub ? __ldg(T.refl_ub + m) : CUDART_INF;
      const unsigned cover = (T.grad_check_mask && T.n_checks) ? (unsigned)__ldg(T.c_cover + m) & T.grad_check_mask : 0u;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (!col_ok[j]) continue;
        const size_t o = (size_t)m * ld + base_n + 8 * j;
        double2 q2 = *reinterpret_cast<const double2*>(q_in + o);
        double2 p2 = *reinterpret_cast<double2*>(p + o);
        auto one = [&](double y, double& q, double& pp, double e, unsigned flag, int c) {
          const double lik = sub ? __dsub_rn(y, sub_m) : y;
          double g = kind ? __dadd_rn(0.0, term_gradient(kind, ta, tb, q)) : 0.0;
          if (flag & cover) g = __dadd_rn(g, CUDART_INF);
          g = __dadd_rn(g, lik);
          momentum_update(__dmul_rn(b_mult, e), g, pp);
          q = __dadd_rn(q, __dmul_rn(__dmul_rn(a_mult, e), T.invm ? __dmul_rn(im, pp) : pp));
          if (has_refl) reflect_on(lb, ub, q, pp);
          if (flags_out) {
            const unsigned v = bound_violations(T, m, q);
            if (v) atomicOr(flags_out + c, v);
          }
        };
        one(acc[i][j][0], q2.x, p2.x, e2[j].x, fl[j].x, base_n + 8 * j);
        one(acc[i][j][1], q2.y, p2.y, e2[j].y, fl[j].y, base_n + 8 * j + 1);
        *reinterpret_cast<double2*>(p + o) = p2;
        *reinterpret_cast<double2*>(q_out + o) = q2;
      }
    }
  }
  // SpMM interface
  __device__ __forceinline__ void tile_begin(int, int) {}
  __device__ __forceinline__ void row(int i, int c, double y) const { apply(i, c, y); }
  __device__ __forceinline__ void chunk_end(int, int) {}
};

// R[i][c] = (Y - d_i) / var_i     (LinearMatrix.py:207-208, 425-426)
struct ResidualEpi {
  static constexpr bool kPerChainSum = false;
  int N, C, ld;
  const double* dvec;
  const double* var;
  double* R;
  __device__ __forceinline__ void apply(int i, int c, double y) const {
    if (i >= N || c >= C) return;
    R[(size_t)i * ld + c] = __ddiv_rn(__dsub_rn(y, __ldg(dvec + i)), __ldg(var + i));
  }
  __device__ __forceinline__ void tile(int, int, int base_m, int base_n, const double (&acc)[8][4][2],
                                       double*) const {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int m = base_m + 8 * i;
      if (m >= N) continue;
      const double dm = __ldg(dvec + m), vm = __ldg(var + m);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = base_n + 8 * j;
        if (c >= C) continue;   // c + 1 may be a padding column: harmless, never read back
        *reinterpret_cast<double2*>(R + (size_t)m * ld + c) =
            make_double2(__ddiv_rn(__dsub_rn(acc[i][j][0], dm), vm), __ddiv_rn(__dsub_rn(acc[i][j][1], dm), vm));
      }
    }
  }
  __device__ __forceinline__ void tile_begin(int, int) {}
  __device__ __forceinline__ void row(int i, int c, double y) const { apply(i, c, y); }
  __device__ __forceinline__ void chunk_end(int, int) {}
};

// out[i][c] = Y   (Full mass matrix: p = L z and dK/dp = M^-1 p over the chain batch,
// MassMatrices.py:241-327)
struct StoreEpi {
  static constexpr bool kPerChainSum = false;
  int rows, C, ld;
  double* out;
  __device__ __forceinline__ void tile(int, int, int base_m, int base_n, const double (&acc)[8][4][2],
                                       double*) const {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int m = base_m + 8 * i;
      if (m >= rows) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = base_n + 8 * j;
        if (c >= C) continue;
        *reinterpret_cast<double2*>(out + (size_t)m * ld + c) = make_double2(acc[i][j][0], acc[i][j][1]);
      }
    }
  }
};

// Per-chain partial sums of the likelihood misfit over the rows of one tile / chunk:
//   premultiplied: q_j * (Y_j - 2*Gtd0_j)            (LinearMatrix.py:185-191)
//   direct       : ((Y_i - d_i) / sigma_i)^2         (LinearMatrix.py:192-202)
struct MisfitEpi {
  static constexpr bool kPerChainSum = true;   // the row-blocked SpMM reduces term() over rows itself
  int mode;  // LIK_PREMULT or LIK_DIRECT
  int rows, C, ld;
  const double* vec;    // Gtd0 [d] or d [N]
  const double* sigma;  // [N] (direct)
  const double* q;      // working positions [dpad x ld] (premult)
  double* part;         // [tiles x ld]
  double acc;           // SpMM: this thread's (= chain's) sum

  __device__ __forceinline__ double term(int i, int c, double y) const {
    if (i >= rows || c >= C) return 0.0;
    if (mode == LIK_PREMULT) {
      const double v = __dsub_rn(y, __dmul_rn(2.0, __ldg(vec + i)));
      return __dmul_rn(q[(size_t)i * ld + c], v);
    }
    const double r = __ddiv_rn(__dsub_rn(y, __ldg(vec + i)), __ldg(sigma + i));
    return __dmul_rn(r, r);
  }
  __device__ __forceinline__ void tile(int m0, int n0, int base_m, int base_n, const double (&a)[8][4][2],
                                       double* smem) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wm = warp >> 2, wn = warp & 3;
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        double v = 0.0;
#pragma unroll
        for (int i = 0; i < 8; ++i) v = __dadd_rn(v, term(base_m + 8 * i, base_n + 8 * j + h, a[i][j][h]));
        v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, 4));
        v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, 8));
        v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, 16));
        if ((lane >> 2) == 0) smem[wm * GEMM_BN + wn * 32 + j * 8 + 2 * (lane & 3) + h] = v;
      }
    consumer_barrier();
    if (threadIdx.x < GEMM_BN) {
      const double v = __dadd_rn(smem[threadIdx.x], smem[GEMM_BN + threadIdx.x]);
      part[(size_t)(m0 / GEMM_BM) * ld + n0 + threadIdx.x] = v;
    }
  }
  __device__ __forceinline__ void tile_begin(int, int) { acc = 0.0; }
  __device__ __forceinline__ void row(int i, int c, double y) { acc = __dadd_rn(acc, term(i, c, y)); }
  __device__ __forceinline__ void chunk_end(int chunk, int c) { part[(size_t)chunk * ld + c] = acc; }
};

// ------------------------------------------------------------ elementwise kernels ---
// (compiled only by launch_staged.cu, which defines HMCB_STAGED_KERNELS)

struct StagedCommon {
  DevTarget T;
  int C, ld, jtiles;
  long long chain_offset;
  unsigned long long seed;
  double stepsize;
  int randomize;
  const double* stepsize_chain;  // [C] per-chain step sizes or null
  const double* v;               // Full mass matrix: dK/dp = M^-1 p [dpad x ld] (null otherwise)
  int draw_only;                 // st_begin: store the standard normals into p and stop (Full mass)
};

#ifdef HMCB_STAGED_KERNELS
// Momentum draw, kinetic energy partials, first (lone) position update.
__global__ void __launch_bounds__(ST_THREADS)
st_begin_kernel(const StagedCommon S, long long kglob, double a_mult,
                const double* __restrict__ q_cur, double* __restrict__ q_w, double* __restrict__ p,
                const double* __restrict__ z_in /* [C x d] of this proposal or null */,
                const double* __restrict__ u_step_in, const double* __restrict__ u_acc_in,
                double* __restrict__ eps_out, double* __restrict__ uacc_out,
                double* __restrict__ k0part, unsigned* __restrict__ flags_out,
                double* __restrict__ stepsize_out /* [C] slice of out_stepsize or null */) {
  const int c = blockIdx.x * ST_THREADS + threadIdx.x;
  if (c >= S.C) return;
  const DevTarget& T = S.T;
  const int jt = blockIdx.y, d = T.dims;
  const uint32_t cg = (uint32_t)(S.chain_offset + c), kg = (uint32_t)kglob;
  double u_step, u_acc;
  uniform_pair(S.seed, cg, kg, u_step, u_acc);
  if (u_step_in) u_step = u_step_in[c];
  if (u_acc_in) u_acc = u_acc_in[c];
  const double eps0 = S.stepsize_chain ? S.stepsize_chain[c] : S.stepsize;
  const double eps = S.randomize ? __dmul_rn(u_step, eps0) : eps0;
  if (jt == 0) {
    eps_out[c] = eps; uacc_out[c] = u_acc;
    if (stepsize_out) stepsize_out[c] = eps0;
  }
  const double ca = __dmul_rn(a_mult, eps);
  double k0 = 0.0;
  unsigned mask = 0;
  const int j_end = min(d, (jt + 1) * ST_DT);
  for (int j = jt * ST_DT; j < j_end; j += 2) {
    double z[2];
    if (z_in) {
      z[0] = z_in[(size_t)c * d + j];
      z[1] = (j + 1 < d) ? z_in[(size_t)c * d + j + 1] : 0.0;
    } else {
      normal_pair(S.seed, cg, kg, (uint32_t)(j >> 1), z[0], z[1]);
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int jj = j + h;
      if (jj < j_end) {
        const size_t o = (size_t)jj * S.ld + c;
        if (S.draw_only) { p[o] = z[h]; continue; }
        double pp = T.sqrtm ? __dmul_rn(__ldg(T.sqrtm + jj), z[h]) : z[h];
        k0 = __dadd_rn(k0, kinetic_term(T, jj, pp));
        double q = q_cur[o];
        position_update(T, jj, ca, q, pp);
        p[o] = pp;
        q_w[o] = q;
        mask |= bound_violations(T, jj, q);
      }
    }
  }
  if (S.draw_only) return;
  k0part[(size_t)jt * S.ld + c] = k0;
  if (flags_out && mask) atomicOr(flags_out + c, mask);
}

// Full mass matrix: position update q_out = q_in + a*eps*v with v = M^-1 p (S.v), reflection, bound
// flags, and (k0part != null) the kinetic-energy partial sums p . v of the momentum it was given.
__global__ void __launch_bounds__(ST_THREADS)
st_kpos_kernel(const StagedCommon S, double a_mult, const double* __restrict__ q_in,
               double* __restrict__ q_out, double* __restrict__ p, const double* __restrict__ eps,
               double* __restrict__ k0part, unsigned* __restrict__ flags_out) {
  const int c = blockIdx.x * ST_THREADS + threadIdx.x;
  if (c >= S.C) return;
  const DevTarget& T = S.T;
  const double ca = __dmul_rn(a_mult, eps[c]);
  double k0 = 0.0;
  unsigned mask = 0;
  const int j_end = min(T.dims, ((int)blockIdx.y + 1) * ST_DT);
  for (int j = blockIdx.y * ST_DT; j < j_end; ++j) {
    const size_t o = (size_t)j * S.ld + c;
    double q = q_in[o], pp = p[o];
    const double vv = S.v[o];
    k0 = __dadd_rn(k0, __dmul_rn(pp, vv));
    q = __dadd_rn(q, __dmul_rn(ca, vv));
    const double before = pp;
    reflect(T, j, q, pp);
    q_out[o] = q;
    if (pp != before) p[o] = pp;
    mask |= bound_violations(T, j, q);
  }
  if (k0part) k0part[(size_t)blockIdx.y * S.ld + c] = k0;
  if (flags_out && mask) atomicOr(flags_out + c, mask);
}

// [C] = scale * column sums of part [tiles x ld]
__global__ void __launch_bounds__(ST_THREADS)
st_colsum_kernel(const double* __restrict__ part, int tiles, int ld, int C, double scale,
                 double* __restrict__ out) {
  const int c = blockIdx.x * ST_THREADS + threadIdx.x;
  if (c >= C) return;
  double s = 0.0;
  for (int t = 0; t < tiles; ++t) s = __dadd_rn(s, part[(size_t)t * ld + c]);
  out[c] = __dmul_rn(scale, s);
}

// Lone position update (the leading a1 sub-step of every 3s/4s step).
__global__ void __launch_bounds__(ST_THREADS)
st_position_kernel(const StagedCommon S, double a_mult, double* __restrict__ q_w,
                   double* __restrict__ p, const double* __restrict__ eps,
                   unsigned* __restrict__ flags_out) {
  const int c = blockIdx.x * ST_THREADS + threadIdx.x;
  if (c >= S.C) return;
  const DevTarget& T = S.T;
  const double ca = __dmul_rn(a_mult, eps[c]);
  unsigned mask = 0;
  const int j_end = min(T.dims, ((int)blockIdx.y + 1) * ST_DT);
  for (int j = blockIdx.y * ST_DT; j < j_end; ++j) {
    const size_t o = (size_t)j * S.ld + c;
    double q = q_w[o], pp = p[o];
    position_update(T, j, ca, q, pp);
    q_w[o] = q; p[o] = pp;
    mask |= bound_violations(T, j, q);
  }
  if (flags_out && mask) atomicOr(flags_out + c, mask);
}

// Gradient + momentum + position update for targets without a matrix likelihood (the
// epilogue applied to a zero likelihood gradient); used when dims exceeds the fused kernel.
__global__ void __launch_bounds__(ST_THREADS)
st_update_kernel(const StagedCommon S, const UpdateEpi epi) {
  const int c = blockIdx.x * ST_THREADS + threadIdx.x;
  if (c >= S.C) return;
  const int j_end = min(S.T.dims, ((int)blockIdx.y + 1) * ST_DT);
  for (int j = blockIdx.y * ST_DT; j < j_end; ++j) epi.apply(j, c, 0.0);
}

// Partial sums of kinetic energy and prior misfit (+ optional bound violations) of (q, p).
__global__ void __launch_bounds__(ST_THREADS)
st_energy_kernel(const StagedCommon S, const double* __restrict__ q, const double* __restrict__ p,
                 double* __restrict__ k1part, double* __restrict__ upart,
                 unsigned* __restrict__ flags_out) {
  const int c = blockIdx.x * ST_THREADS + threadIdx.x;
  if (c >= S.C) return;
  const DevTarget& T = S.T;
  double k1 = 0.0, u1 = 0.0;
  unsigned mask = 0;
  const int j_end = min(T.dims, ((int)blockIdx.y + 1) * ST_DT);
  for (int j = blockIdx.y * ST_DT; j < j_end; ++j) {
    const size_t o = (size_t)j * S.ld + c;
    const double qq = q[o];
    if (p) k1 = __dadd_rn(k1, S.v ? __dmul_rn(p[o], S.v[o]) : kinetic_term(T, j, p[o]));
    u1 = __dadd_rn(u1, prior_misfit(T, j, qq));
    if (flags_out) mask |= bound_violations(T, j, qq);
  }
  const size_t po = (size_t)blockIdx.y * S.ld + c;
  if (k1part) k1part[po] = k1;
  upart[po] = u1;
  if (flags_out && mask) atomicOr(flags_out + c, mask);
}

#endif  // HMCB_STAGED_KERNELS

struct DecideArgs {
  int C, ld, jtiles, ltiles, lik_mode;
  double dtd, const_sum;
  const double *k0part, *k1part, *upart, *lpart, *uacc;
  const unsigned* flags;
  double* x;              // [C] in/out (decide) or out (misfit only)
  unsigned char* acc;     // [ld]
  unsigned char* out_accept;  // [C] slice of this proposal or null
  double *out_h0, *out_h1;    // [C] slices or null
  int* accepted_total;
  double* sample_misfit;  // out_samples + row*C*(d+1) + d, stride (d+1), or null
  int sample_stride;
  int misfit_only;
  double* stepsize_chain;  // [C] in/out when tune.enabled
  AutotuneArgs tune;
  long long kglob;
};

#ifdef HMCB_STAGED_KERNELS
__device__ __forceinline__ double column_sum(const double* part, int tiles, int ld, int c) {
  double s = 0.0;
  for (int t = 0; t < tiles; ++t) s = __dadd_rn(s, part[(size_t)t * ld + c]);
  return s;
}

__global__ void __launch_bounds__(ST_THREADS)
st_decide_kernel(const DecideArgs D) {
  const int c = blockIdx.x * ST_THREADS + threadIdx.x;
  if (c >= D.C) return;
  const double u1 = column_sum(D.upart, D.jtiles, D.ld, c);
  double lik = 0.0;
  if (D.lik_mode != LIK_NONE) {
    const double L = column_sum(D.lpart, D.ltiles, D.ld, c);
    if (D.lik_mode == LIK_PREMULT) {
      lik = __dmul_rn(0.5, __dadd_rn(L, D.dtd));
    } else {
      const double nrm = sqrt(L);  // 0.5 * numpy.linalg.norm(r)**2
      lik = __dmul_rn(0.5, __dmul_rn(nrm, nrm));
    }
  }
  double x1 = __dadd_rn(__dadd_rn(u1, D.const_sum), lik);
  if (D.flags && D.flags[c]) x1 = __dadd_rn(x1, CUDART_INF);
  if (D.misfit_only) { D.x[c] = x1; return; }
  const double k0 = column_sum(D.k0part, D.jtiles, D.ld, c);
  const double k1 = column_sum(D.k1part, D.jtiles, D.ld, c);
  const double x0 = D.x[c];
  const double h0 = __dadd_rn(x0, __dmul_rn(0.5, k0));
  const double h1 = __dadd_rn(x1, __dmul_rn(0.5, k1));
  const bool acc = metropolis_accept(h0, h1, D.uacc[c]);
  if (D.tune.enabled) D.stepsize_chain[c] = autotune_stepsize(D.tune, D.stepsize_chain[c], h0, h1, D.kglob);
  D.acc[c] = acc ? 1 : 0;
  const double xn = acc ? x1 : x0;
  if (acc) {
    D.x[c] = x1;
    if (D.accepted_total) D.accepted_total[c] += 1;
  }
  if (D.out_accept) D.out_accept[c] = acc ? 1 : 0;
  if (D.out_h0) D.out_h0[c] = h0;
  if (D.out_h1) D.out_h1[c] = h1;
  if (D.sample_misfit) D.sample_misfit[(size_t)c * D.sample_stride] = xn;
}

// Accepted chains take the proposal; optional chain-major outputs go through a shared
// memory transpose so both sides stay coalesced.  block (32, 8), tile 32 chains x 32 dims.
__global__ void __launch_bounds__(256)
st_commit_kernel(int C, int d, int ld, const unsigned char* __restrict__ acc,
                 const double* __restrict__ q_w, const double* __restrict__ p,
                 double* __restrict__ q_cur, double* __restrict__ sample_rows /* [C x (d+1)] or null */,
                 double* __restrict__ q_prop /* [C x d] or null */, double* __restrict__ p_prop) {
  __shared__ double tq[32][33], tw[32][33], tp[32][33];
  const int c0 = blockIdx.x * 32, j0 = blockIdx.y * 32;
  const int c = c0 + threadIdx.x;
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int j = j0 + r;
    if (c < C && j < d) {
      const size_t o = (size_t)j * ld + c;
      const double w = q_w[o];
      double cur = q_cur[o];
      if (acc[c]) { cur = w; q_cur[o] = w; }
      tq[r][threadIdx.x] = cur;
      tw[r][threadIdx.x] = w;
      tp[r][threadIdx.x] = p ? p[o] : 0.0;
    }
  }
  __syncthreads();
  const int j = j0 + threadIdx.x;
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int cc = c0 + r;
    if (cc < C && j < d) {
      if (sample_rows) sample_rows[(size_t)cc * (d + 1) + j] = tq[threadIdx.x][r];
      if (q_prop) {
        q_prop[(size_t)cc * d + j] = tw[threadIdx.x][r];
        p_prop[(size_t)cc * d + j] = tp[threadIdx.x][r];
      }
    }
  }
}

// out[c_][r_] = in[r_][c_]  for in [R x ldin] (only Cc valid columns) -> out [Cc x ldout]
__global__ void __launch_bounds__(256)
st_transpose_kernel(const double* __restrict__ in, int R, int Cc, int ldin, double* __restrict__ out,
                    int ldout) {
  __shared__ double tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += 8)
    if (r0 + r < R && c0 + threadIdx.x < Cc)
      tile[r][threadIdx.x] = in[(size_t)(r0 + r) * ldin + c0 + threadIdx.x];
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += 8)
    if (c0 + r < Cc && r0 + threadIdx.x < R)
      out[(size_t)(c0 + r) * ldout + r0 + threadIdx.x] = tile[threadIdx.x][r];
}

#endif  // HMCB_STAGED_KERNELS

}  // namespace hmcb

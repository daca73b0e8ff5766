// fused.cuh -- whole-proposal fused HMC kernels for targets whose gradient does not couple
// chains through a matrix product: (a) elementwise priors only, (b) SourceLocation3D.
//
// One launch advances every chain by B proposals: momentum draw, kinetic energy, the full
// leapfrog / 3-stage / 4-stage trajectory with bounds reflection, misfit, Hamiltonians,
// Metropolis decision, state update and sample-row write.  The chain state lives in
// registers for the whole block of proposals; HBM sees q once in, once out, and one
// (d+1)-row per stored sample.
//
// Replaces Samplers.py:1463-1492 (propose + evaluate_acceptance), :1524-1726 (integrators),
// MassMatrices.py:100-142,185-227, base.py:239-270 (corrector) for these targets.
#pragma once
#include "common.cuh"

#ifndef HMCB_FUSED_MINBLOCKS
#define HMCB_FUSED_MINBLOCKS 3  // 80 registers: 3 blocks of 256 threads per SM measured 8% faster than 2
#endif

namespace hmcb {

struct FusedArgs {
  DevTarget T;
  Schedule S;
  int chains;
  int proposals;          // B
  long long thinning;
  long long proposal_offset;
  long long chain_offset;
  unsigned long long seed;
  double stepsize;
  int randomize;
  double* q;              // [C x d]
  double* x;              // [C]
  const double* z_in;     // [B x C x d] or null
  const double* u_step_in;
  const double* u_accept_in;
  double* out_samples;
  unsigned char* out_accept;
  double* out_h0;
  double* out_h1;
  int* accepted_total;
  double* out_q_prop;
  double* out_p_prop;
  double* trace_q;
  double* trace_g;
  double* stepsize_chain;  // [C] in/out or null
  double* out_stepsize;    // [B x C] or null
  AutotuneArgs tune;
  int exact;               // 1: no FMA contraction in the trajectory (bit-identical to numpy on separable targets)
};

// ------------------------------------------------------------------- priors only ---
// Thread t of a chain owns coordinate pairs m = t + i*TPC (i < PPT): coordinates 2m, 2m+1.
// Everything that depends only on the coordinate (prior parameters, and with AUX the
// reflection bounds and the mass matrix) is loaded into registers once; the trajectory
// loop is pure fp64 arithmetic.  Padding coordinates (>= dims) carry q = p = 0 and a
// "none" prior, so they add exact zeros to every reduction and need no predicates.
// Supports at most one prior term per coordinate (T.n_terms <= 1); AUX = false requires
// a unit mass matrix and no reflection bounds.
// FMA = true contracts the two updates of a leapfrog sub-step (p -= cb * g, q += ca * dK/dp) into one
// fused multiply-add each: 4 instead of 6 fp64 instructions per coordinate and gradient evaluation and a
// dependent chain of 3 instead of 5.  Results then differ from numpy's in the last bits (well inside the
// 1e-10 parity bar, accept/reject decisions unchanged on every golden); FMA = false
// (hmcb_set_exact_arithmetic) keeps the reference's operation order bit for bit.
template <int TPC, int PPT, bool AUX, bool FMA>
__global__ void __launch_bounds__((TPC < 256 ? 256 : TPC), (TPC > 256 ? 1 : (PPT < 4 ? HMCB_FUSED_MINBLOCKS : 2)))
hmc_fused_priors_kernel(const FusedArgs A) {
  constexpr int BLOCK = TPC < 256 ? 256 : TPC;
  constexpr int CPB = BLOCK / TPC;  // chains per block
  constexpr int E = 2 * PPT;        // coordinates per thread
  constexpr int EA = AUX ? E : 1;
  __shared__ double scratch[ChainReduce<TPC>::scratch_doubles(BLOCK)];
  const ChainReduce<TPC> red{scratch};
  const DevTarget& T = A.T;
  const int d = T.dims;

  const int t = threadIdx.x % TPC;
  int c = blockIdx.x * CPB + threadIdx.x / TPC;
  const bool live = c < A.chains;
  if (!live) c = A.chains - 1;  // keep the lanes converged; all writes are predicated
  const size_t row = (size_t)c * d;
  const size_t C = (size_t)A.chains;

  int jj[E];
  unsigned okmask = 0, kinds = 0;
  double qc[E], q[E], p[E] = {}, ta[E], tb[E];
  double rlb[EA] = {}, rub[EA] = {}, im[EA] = {}, sm[EA] = {};
#pragma unroll
  for (int i = 0; i < PPT; ++i) {
    const int m = t + i * TPC;
    jj[2 * i] = 2 * m; jj[2 * i + 1] = 2 * m + 1;
  }
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const bool ok = jj[e] < d;
    if (ok) okmask |= 1u << e;
    qc[e] = ok ? A.q[row + jj[e]] : 0.0;
    ta[e] = 0.0; tb[e] = 0.0;
    if (ok && T.n_terms > 0) {
      kinds |= (unsigned)T.t_kind[jj[e]] << (2 * e);
      ta[e] = T.t_a[jj[e]]; tb[e] = T.t_b[jj[e]];
    }
    if constexpr (AUX) {
      rlb[e] = (ok && T.refl_lb) ? T.refl_lb[jj[e]] : -CUDART_INF;
      rub[e] = (ok && T.refl_ub) ? T.refl_ub[jj[e]] : CUDART_INF;
      im[e] = (ok && T.invm) ? T.invm[jj[e]] : 1.0;
      sm[e] = (ok && T.sqrtm) ? T.sqrtm[jj[e]] : 1.0;
    }
  }
  const bool has_mass = AUX && T.invm != nullptr;
  const bool has_refl = AUX && (T.refl_lb != nullptr || T.refl_ub != nullptr);
  auto ok = [&](int e) { return (okmask >> e) & 1u; };
  auto kind_of = [&](int e) { return (int)((kinds >> (2 * e)) & 3u); };
  auto dkdp = [&](int e) {
    if constexpr (AUX) return has_mass ? __dmul_rn(im[e], p[e]) : p[e];
    else return p[e];
  };
  auto violations = [&]() {
    unsigned m = 0;
#pragma unroll
    for (int e = 0; e < E; ++e)
      if (ok(e)) m |= bound_violations(T, jj[e], q[e]);
    return red.any_bits(m);
  };

  double x = A.x[c];
  int accepted = 0;
  const bool grad_checks = T.grad_check_mask != 0u;
  // Fast path: every coordinate carries exactly one prior term of the same kind, no bound
  // check is visible to the gradient and nothing is traced -> branch-free element loops.
  const int fast_kind = (!grad_checks && !A.trace_q) ? T.uniform_kind : TERM_NONE;

  // Sub-steps of the trajectory (see run_schedule): momentum update with the gradient of
  // the priors, position update with reflection.
  int gi = 0, kb = 0;
  auto mom_normal = [&](double cb) {
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const double g = term_gradient(TERM_NORMAL, ta[e], tb[e], q[e]);
      if constexpr (FMA) p[e] = fma(-cb, g, p[e]);
      else momentum_update(cb, g, p[e]);
    }
  };
  auto mom_laplace = [&](double cb) {
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const double g = term_gradient(TERM_LAPLACE, ta[e], tb[e], q[e]);
      if constexpr (FMA) p[e] = fma(-cb, g, p[e]);
      else momentum_update(cb, g, p[e]);
    }
  };
  auto mom_generic = [&](double cb) {
    const unsigned oob = grad_checks ? (violations() & T.grad_check_mask) : 0u;
    const size_t C_ = (size_t)A.chains;
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int kind = kind_of(e);
      double g = kind ? term_gradient(kind, ta[e], tb[e], q[e]) : 0.0;
      if (oob && ok(e) && (oob & T.c_cover[jj[e]])) g = __dadd_rn(g, CUDART_INF);
      if (A.trace_q && live && ok(e)) {
        const size_t o = (((size_t)kb * A.S.grads_per_proposal + gi) * C_ + c) * d + jj[e];
        A.trace_q[o] = q[e];
        A.trace_g[o] = g;
      }
      if constexpr (FMA) p[e] = fma(-cb, g, p[e]);
      else momentum_update(cb, g, p[e]);
    }
    ++gi;
  };
  auto pos = [&](double ca) {
#pragma unroll
    for (int e = 0; e < E; ++e) {
      if constexpr (FMA) q[e] = fma(ca, dkdp(e), q[e]);
      else q[e] = __dadd_rn(q[e], __dmul_rn(ca, dkdp(e)));
      if constexpr (AUX) { if (has_refl) reflect_on(rlb[e], rub[e], q[e], p[e]); }
    }
  };

  UniformPairCache<TPC> ucache;
  ucache.us = ucache.ua = 0.0;
  double eps0 = A.stepsize_chain ? A.stepsize_chain[c] : A.stepsize;  // this chain's step size
  // stored proposals: global index k with k % thinning == 0 (no division inside the loop)
  long long next_store = (A.proposal_offset + A.thinning - 1) / A.thinning * A.thinning;
  size_t store_row = 0;
  for (kb = 0; kb < A.proposals; ++kb) {
    const long long kglob = A.proposal_offset + kb;
    const size_t kc = (size_t)kb * C + c;

    // ---- draws -------------------------------------------------------------------
    double u_step, u_acc;
    if (A.u_step_in && A.u_accept_in) {
      u_step = A.u_step_in[kc]; u_acc = A.u_accept_in[kc];
    } else {
      constexpr int W = UniformPairCache<TPC>::W;
      if ((kb % W) == 0) ucache.fill(A.seed, (uint32_t)(A.chain_offset + c), kglob);
      ucache.get(kb % W, u_step, u_acc);
      if (A.u_step_in) u_step = A.u_step_in[kc];
      if (A.u_accept_in) u_acc = A.u_accept_in[kc];
    }
    const double eps = A.randomize ? __dmul_rn(u_step, eps0) : eps0;
    if (A.out_stepsize && live && t == 0) A.out_stepsize[kc] = eps0;

#pragma unroll
    for (int i = 0; i < PPT; ++i) {
      double z0, z1;
      if (A.z_in) {
        const double* zr = A.z_in + kc * d;
        z0 = ok(2 * i) ? zr[jj[2 * i]] : 0.0;
        z1 = ok(2 * i + 1) ? zr[jj[2 * i + 1]] : 0.0;
      } else {
        normal_pair(A.seed, (uint32_t)(A.chain_offset + c), (uint32_t)kglob,
                    (uint32_t)(t + i * TPC), z0, z1);
        if (!ok(2 * i)) z0 = 0.0;
        if (!ok(2 * i + 1)) z1 = 0.0;
      }
      p[2 * i] = z0; p[2 * i + 1] = z1;
    }
    double k0 = 0.0;
#pragma unroll
    for (int e = 0; e < E; ++e) {
      q[e] = qc[e];
      if constexpr (AUX) { if (has_mass) p[e] = __dmul_rn(sm[e], p[e]); }
      k0 = __dadd_rn(k0, __dmul_rn(p[e], dkdp(e)));
    }

    // ---- trajectory --------------------------------------------------------------
    gi = 0;
    if (fast_kind == TERM_NORMAL) run_schedule(A.S, eps, mom_normal, pos);
    else if (fast_kind == TERM_LAPLACE) run_schedule(A.S, eps, mom_laplace, pos);
    else run_schedule(A.S, eps, mom_generic, pos);

    // ---- energies and decision ----------------------------------------------------
    double k1 = 0.0, u1 = 0.0;
#pragma unroll
    for (int e = 0; e < E; ++e) {
      k1 = __dadd_rn(k1, __dmul_rn(p[e], dkdp(e)));
      const int kind = kind_of(e);
      if (kind) u1 = __dadd_rn(u1, term_misfit(kind, ta[e], tb[e], q[e]));
    }
    red.sum3(k0, k1, u1);
    const unsigned oob = T.n_checks ? violations() : 0u;
    double x1 = __dadd_rn(u1, T.const_sum);
    if (oob) x1 = __dadd_rn(x1, CUDART_INF);
    const double h0 = __dadd_rn(x, __dmul_rn(0.5, k0));
    const double h1 = __dadd_rn(x1, __dmul_rn(0.5, k1));
    const bool acc = metropolis_accept(h0, h1, u_acc);
    if (A.tune.enabled) eps0 = autotune_stepsize(A.tune, eps0, h0, h1, kglob);

    if (live) {
      if (A.out_q_prop) {
#pragma unroll
        for (int e = 0; e < E; ++e)
          if (ok(e)) {
            A.out_q_prop[kc * d + jj[e]] = q[e];
            A.out_p_prop[kc * d + jj[e]] = p[e];
          }
      }
      if (t == 0) {
        if (A.out_accept) A.out_accept[kc] = acc ? 1 : 0;
        if (A.out_h0) A.out_h0[kc] = h0;
        if (A.out_h1) A.out_h1[kc] = h1;
      }
    }
    if (acc) {
#pragma unroll
      for (int e = 0; e < E; ++e) qc[e] = q[e];
      x = x1;
      ++accepted;
    }
    const bool store_now = kglob == next_store;
    if (store_now) next_store += A.thinning;
    if (A.out_samples && live && store_now) {
      const size_t srow = (store_row * C + c) * (size_t)(d + 1);
#pragma unroll
      for (int e = 0; e < E; ++e)
        if (ok(e)) A.out_samples[srow + jj[e]] = qc[e];
      if (t == 0) A.out_samples[srow + d] = x;
    }
    if (store_now) ++store_row;
  }

  if (live) {
#pragma unroll
    for (int e = 0; e < E; ++e)
      if (ok(e)) A.q[row + jj[e]] = qc[e];
    if (t == 0) {
      A.x[c] = x;
      if (A.accepted_total) A.accepted_total[c] += accepted;
      if (A.tune.enabled && A.stepsize_chain) A.stepsize_chain[c] = eps0;
    }
  }
}

// Standalone batched misfit / gradient / reflection / mass-matrix kernels for the
// single-vector protocol entry points (one thread per coordinate, block per chain-slab).
template <int TPC>
__global__ void __launch_bounds__((TPC < 256 ? 256 : TPC))
prior_misfit_kernel(const DevTarget T, int chains, const double* __restrict__ q,
                    double* __restrict__ x, const double* __restrict__ lik_misfit) {
  constexpr int BLOCK = TPC < 256 ? 256 : TPC;
  constexpr int CPB = BLOCK / TPC;
  __shared__ double scratch[ChainReduce<TPC>::scratch_doubles(BLOCK)];
  const ChainReduce<TPC> red{scratch};
  const int t = threadIdx.x % TPC;
  int c = blockIdx.x * CPB + threadIdx.x / TPC;
  const bool live = c < chains;
  if (!live) c = chains - 1;
  double u = 0.0, z0 = 0.0, z1 = 0.0;
  unsigned oob = 0;
  for (int j = t; j < T.dims; j += TPC) {
    const double v = q[(size_t)c * T.dims + j];
    u = __dadd_rn(u, prior_misfit(T, j, v));
    oob |= bound_violations(T, j, v);
  }
  red.sum3(u, z0, z1);
  oob = red.any_bits(oob);
  if (live && t == 0) {
    double r = __dadd_rn(u, T.const_sum);
    if (lik_misfit) r = __dadd_rn(r, lik_misfit[c]);
    if (oob) r = __dadd_rn(r, CUDART_INF);
    x[c] = r;
  }
}

template <int TPC>
__global__ void __launch_bounds__((TPC < 256 ? 256 : TPC))
prior_gradient_kernel(const DevTarget T, int chains, const double* __restrict__ q,
                      double* __restrict__ g, int accumulate) {
  constexpr int BLOCK = TPC < 256 ? 256 : TPC;
  constexpr int CPB = BLOCK / TPC;
  __shared__ double scratch[ChainReduce<TPC>::scratch_doubles(BLOCK)];
  const ChainReduce<TPC> red{scratch};
  const int t = threadIdx.x % TPC;
  int c = blockIdx.x * CPB + threadIdx.x / TPC;
  const bool live = c < chains;
  if (!live) c = chains - 1;
  unsigned oob = 0;
  if (T.grad_check_mask) {
    for (int j = t; j < T.dims; j += TPC)
      oob |= bound_violations(T, j, q[(size_t)c * T.dims + j]);
    oob = red.any_bits(oob);
  }
  if (!live) return;
  for (int j = t; j < T.dims; j += TPC) {
    const size_t o = (size_t)c * T.dims + j;
    const double pg = prior_gradient(T, j, q[o], oob);
    g[o] = accumulate ? __dadd_rn(pg, g[o]) : pg;
  }
}

#ifdef HMCB_FUSED_AUX_KERNELS
__global__ void reflect_kernel(const DevTarget T, size_t total, double* __restrict__ q,
                               double* __restrict__ p) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int j = (int)(i % T.dims);
  double qq = q[i], pp = p[i];
  reflect(T, j, qq, pp);
  q[i] = qq; p[i] = pp;
}

// mode 0: p = sqrt(M) z ; mode 1: dK/dp
__global__ void mass_elementwise_kernel(const DevTarget T, size_t total, int mode,
                                        const double* __restrict__ in, double* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int j = (int)(i % T.dims);
  const double v = in[i];
  if (mode == 0) out[i] = T.sqrtm ? __dmul_rn(__ldg(T.sqrtm + j), v) : v;
  else out[i] = kinetic_gradient(T, j, v);
}
#endif  // HMCB_FUSED_AUX_KERNELS

template <int TPC>
__global__ void __launch_bounds__((TPC < 256 ? 256 : TPC))
kinetic_energy_kernel(const DevTarget T, int chains, const double* __restrict__ p,
                      double* __restrict__ k) {
  constexpr int BLOCK = TPC < 256 ? 256 : TPC;
  constexpr int CPB = BLOCK / TPC;
  __shared__ double scratch[ChainReduce<TPC>::scratch_doubles(BLOCK)];
  const ChainReduce<TPC> red{scratch};
  const int t = threadIdx.x % TPC;
  int c = blockIdx.x * CPB + threadIdx.x / TPC;
  const bool live = c < chains;
  if (!live) c = chains - 1;
  double s = 0.0, z0 = 0.0, z1 = 0.0;
  for (int j = t; j < T.dims; j += TPC)
    s = __dadd_rn(s, kinetic_term(T, j, p[(size_t)c * T.dims + j]));
  red.sum3(s, z0, z1);
  if (live && t == 0) k[c] = __dmul_rn(0.5, s);
}

}  // namespace hmcb

// srcloc.cuh -- SourceLocation3D travel-time misfit / gradient and the fused HMC kernel
// built on it (compute-bound elementwise-reduction; SourceLocation.py:482-540, 697-713).
//
// Thread mapping: a chain is owned by TPC threads; event e = t / LPE, the LPE lanes of an
// event split the stations (s = sub, sub+LPE, ...).  All LPE lanes of an event hold the
// event's 4 parameters (x, y, z, T) and integrate them redundantly (identical bits), so
// the only communication per gradient is an LPE-lane butterfly.  Station geometry and
// picks are staged in shared memory once per block.
#pragma once
#include "common.cuh"
#include "fused.cuh"

namespace hmcb {

// Chains that fit inside a warp are packed into blocks of 64 threads: with equal work per
// block and an fp64-bound SM, small blocks let the hardware balance the last wave (8192 chains
// of config 5 in 256-thread blocks ran in the time of 9472).
constexpr int SRCLOC_SMALL_BLOCK = 64;

struct SrcLocDev {
  int events, stations, infer_velocity;
  int np;  // parameters per event: 4 = (x, y, z, T) SourceLocation3D, 3 = (x, z, T) SourceLocation2D
  double velocity;
  const double* rx; const double* ry; const double* rz;  // [S]
  const double* tobs; const double* std;                 // [E x S]
};

struct SrcLocShared {
  const double2* xy;    // [S] station (x, y)
  const double* z;      // [S] station z
  const double2* pick;  // [E x S] (t_obs, 1/sigma^2); a missing pick is stored as (0, 0)
};

// Shared-memory doubles: 3 per station + 2 per event-station pair.
__host__ __device__ inline size_t srcloc_smem_doubles(int events, int stations) {
  return 3 * (size_t)stations + 2 * (size_t)events * stations + 2;
}

__device__ __forceinline__ SrcLocShared srcloc_stage(const SrcLocDev& L, double* smem) {
  const int S = L.stations, ES = L.events * L.stations;
  double2* pick = reinterpret_cast<double2*>(smem);          // 16-byte aligned
  double2* xy = pick + ES;
  double* z = reinterpret_cast<double*>(xy + S);
  for (int i = threadIdx.x; i < S; i += blockDim.x) {
    xy[i] = make_double2(L.rx[i], L.ry[i]);
    z[i] = L.rz[i];
  }
  for (int i = threadIdx.x; i < ES; i += blockDim.x) {
    const double sd = L.std[i], t = L.tobs[i];
    const double iv = __ddiv_rn(1.0, __dmul_rn(sd, sd));
    // nansum (SourceLocation.py:487-490, 519-523) drops a pair whose pick or uncertainty is
    // NaN from every sum: a zero weight does the same without a per-term test
    const bool missing = (t != t) || (iv != iv);
    pick[i] = missing ? make_double2(0.0, 0.0) : make_double2(t, iv);
  }
  __syncthreads();
  return SrcLocShared{xy, z, pick};
}

__device__ __forceinline__ double nan_to_zero(double v) { return (v != v) ? 0.0 : v; }

// 1/sqrt(d2) for d2 >= 0 without special-case paths; 0 for d2 == 0, so that a source on top of
// a station contributes no direction term (the reference drops its 0/0 through nansum).
__device__ __forceinline__ double rsqrt_or_zero(double d2) {
  double y = rsqrt_seed(d2);
  const double h = 0.5 * d2;
  y = fma(y, fma(-(h * y), y, 0.5), y);
  y = fma(y, fma(-(h * y), y, 0.5), y);
  return (d2 > 0.0) ? y : 0.0;
}

// The reference evaluates, per event-station pair, dist = sqrt(.), t = T + dist / v,
// w = (t - t_obs) / sigma^2 and the direction terms dx / (v * dist) (SourceLocation.py:495-524):
// six IEEE divisions and a square root.  Here one reciprocal square root per pair and
// precomputed reciprocals (1/v, 1/sigma^2) replace them and the sums are accumulated with FMAs;
// terms agree with the reference to a few ulp (the sums over stations already differ from
// numpy's pairwise order at that level; parity tolerance is 1e-10).
template <int LPE>
__device__ __forceinline__ void srcloc_gradient_partial(const SrcLocShared& M, int S, int e, int sub,
                                                        double x, double y, double z, double T,
                                                        double inv_v, bool want_gv, bool nansum,
                                                        double& gx, double& gy, double& gz, double& gT,
                                                        double& gv) {
  gx = gy = gz = gT = gv = 0.0;
  const double neg_inv_vv = -__dmul_rn(inv_v, inv_v);
  const double2* pick = M.pick + e * S;
#pragma unroll 3
  for (int s = sub; s < S; s += LPE) {
    const double2 r = M.xy[s];
    const double2 pk = pick[s];
    const double dx = __dsub_rn(x, r.x), dy = __dsub_rn(y, r.y), dz = __dsub_rn(z, M.z[s]);
    const double d2 = fma(dx, dx, fma(dy, dy, __dmul_rn(dz, dz)));
    const double rinv = rsqrt_or_zero(d2);
    const double dist = __dmul_rn(d2, rinv);
    const double tcalc = fma(dist, inv_v, T);
    const double w = __dmul_rn(__dsub_rn(tcalc, pk.x), pk.y);
    const double u = __dmul_rn(w, __dmul_rn(inv_v, rinv));  // w / (v * dist)
    gx = fma(u, dx, gx);
    gy = fma(u, dy, gy);
    gz = fma(u, dz, gz);
    gT = __dadd_rn(gT, w);
    if (want_gv) gv = fma(w, __dmul_rn(dist, neg_inv_vv), gv);
  }
  // A trajectory that has already diverged (non-finite coordinates) makes every term NaN;
  // nansum returns 0 for such a sum, and the bounds term then decides (+inf).
  // (the 2-D reference sums with a plain sum, SourceLocation.py:133-136: NaN propagates there)
  if (nansum) {
    gx = nan_to_zero(gx); gy = nan_to_zero(gy); gz = nan_to_zero(gz); gT = nan_to_zero(gT);
    gv = nan_to_zero(gv);
  }
}

// Partial sum of squared standardised residuals of event e (SourceLocation.py:482-493).
template <int LPE>
__device__ __forceinline__ double srcloc_misfit_partial(const SrcLocShared& M, int S, int e, int sub,
                                                        double x, double y, double z, double T,
                                                        double inv_v) {
  double acc = 0.0;
  const double2* pick = M.pick + e * S;
#pragma unroll 3
  for (int s = sub; s < S; s += LPE) {
    const double2 r = M.xy[s];
    const double2 pk = pick[s];
    const double dx = __dsub_rn(x, r.x), dy = __dsub_rn(y, r.y), dz = __dsub_rn(z, M.z[s]);
    const double d2 = fma(dx, dx, fma(dy, dy, __dmul_rn(dz, dz)));
    const double dist = __dmul_rn(d2, rsqrt_or_zero(d2));
    const double res = __dsub_rn(pk.x, fma(dist, inv_v, T));
    acc = fma(__dmul_rn(res, res), pk.y, acc);
  }
  return nan_to_zero(acc);
}

template <int LPE>
__device__ __forceinline__ double group_sum(double v) {
#pragma unroll
  for (int off = LPE / 2; off > 0; off >>= 1) v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, off));
  return v;
}

// Per-thread view of one chain: 4 event parameters (+ the shared velocity).
template <int TPC, int LPE, int NP = 4>
struct SrcLocLane {
  int e, sub, c;
  bool has_event, lead, vlead, live;

  __device__ __forceinline__ void init(int chains, int E) {
    constexpr int BLOCK = TPC <= 32 ? SRCLOC_SMALL_BLOCK : TPC;
    constexpr int CPB = BLOCK / TPC;
    const int t = threadIdx.x % TPC;
    c = blockIdx.x * CPB + threadIdx.x / TPC;
    live = c < chains;
    if (!live) c = chains - 1;
    e = t / LPE; sub = t % LPE;
    has_event = e < E;
    lead = has_event && sub == 0;  // the lane that accounts for / writes the event's coordinates
    vlead = t == 0;                // the lane that accounts for / writes the velocity
    if (!has_event) e = E - 1;
  }
  // external coordinate of internal slot i (x, y, z, T); the 2-D variant has no y (-1)
  // slot i carries a parameter (compile-time once the loops over i are unrolled)
  __device__ __forceinline__ static constexpr bool active(int i) { return NP == 4 || i != 1; }
  __device__ __forceinline__ int coord(int i) const {
    return NP == 4 ? 4 * e + i : (i == 1 ? -1 : 3 * e + (i == 0 ? 0 : i - 1));
  }
};

// Gradient of the full target at the lane's coordinates: prior terms + travel-time term.
// g[0..3] for (x,y,z,T) of the lane's event, gvel for the velocity coordinate.
template <int TPC, int LPE, int NP>
__device__ __forceinline__ void srcloc_total_gradient(const DevTarget& T, const SrcLocDev& L,
                                                      const SrcLocShared& M,
                                                      const SrcLocLane<TPC, LPE, NP>& ln,
                                                      const ChainReduce<TPC>& red, const double* q,
                                                      double qv, unsigned oob, double* g, double& gvel) {
  const double inv_v = __ddiv_rn(1.0, L.infer_velocity ? qv : L.velocity);
  double gx, gy, gz, gT, gv;
  srcloc_gradient_partial<LPE>(M, L.stations, ln.e, ln.sub, q[0], q[1], q[2], q[3], inv_v,
                               L.infer_velocity != 0, NP == 4, gx, gy, gz, gT, gv);
  gx = group_sum<LPE>(gx); gy = group_sum<LPE>(gy); gz = group_sum<LPE>(gz); gT = group_sum<LPE>(gT);
  const double lik[4] = {gx, gy, gz, gT};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int j = ln.coord(i);
    g[i] = ln.active(i) ? __dadd_rn(prior_gradient(T, j, q[i], oob), lik[i]) : 0.0;
  }
  gvel = 0.0;
  if (L.infer_velocity) {
    double a = ln.has_event ? gv : 0.0, b = 0.0, c2 = 0.0;
    red.sum3(a, b, c2);
    gvel = __dadd_rn(prior_gradient(T, NP * L.events, qv, oob), a);
  }
}

template <int TPC, int LPE, int NP>
__device__ __forceinline__ unsigned srcloc_violations(const DevTarget& T, const SrcLocDev& L,
                                                      const SrcLocLane<TPC, LPE, NP>& ln,
                                                      const ChainReduce<TPC>& red, const double* q,
                                                      double qv) {
  unsigned m = 0;
  if (T.n_checks) {
    if (ln.lead) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int j = ln.coord(i);
        if (ln.active(i)) m |= bound_violations(T, j, q[i]);
      }
    }
    if (L.infer_velocity && ln.vlead) m |= bound_violations(T, NP * L.events, qv);
    m = red.any_bits(m);
  }
  return m;
}

template <int TPC, int LPE, int NP>
__global__ void __launch_bounds__((TPC <= 32 ? SRCLOC_SMALL_BLOCK : TPC), (TPC <= 32 ? 8 : (TPC <= 256 ? 2 : 1)))
hmc_fused_srcloc_kernel(const FusedArgs A, const SrcLocDev L) {
  extern __shared__ __align__(16) double dyn_smem[];
  __shared__ double scratch[ChainReduce<TPC>::scratch_doubles(TPC <= 32 ? SRCLOC_SMALL_BLOCK : TPC)];
  const ChainReduce<TPC> red{scratch};
  const SrcLocShared M = srcloc_stage(L, dyn_smem);
  const DevTarget& T = A.T;
  const int d = T.dims, E = L.events;
  SrcLocLane<TPC, LPE, NP> ln;
  ln.init(A.chains, E);
  const int c = ln.c;
  const size_t row = (size_t)c * d, C = (size_t)A.chains;
  int cj[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) cj[i] = ln.coord(i);
  constexpr int kNP = NP;
  const int jv = kNP * E;
  const bool inferv = L.infer_velocity != 0;
  const bool grad_checks = T.grad_check_mask != 0u;

  double qc[4], q[4], p[4], qcv = 0.0, qv = 0.0, pv = 0.0;
#pragma unroll
  for (int i = 0; i < 4; ++i) qc[i] = ln.active(i) ? A.q[row + cj[i]] : 0.0;
  if (inferv) qcv = A.q[row + jv];
  double x = A.x[c];
  int accepted = 0;

  UniformPairCache<TPC> ucache;
  ucache.us = ucache.ua = 0.0;
  double eps0 = A.stepsize_chain ? A.stepsize_chain[c] : A.stepsize;  // this chain's step size
  // stored proposals: global index k with k % thinning == 0 (no division inside the loop)
  long long next_store = (A.proposal_offset + A.thinning - 1) / A.thinning * A.thinning;
  size_t store_row = 0;
  for (int kb = 0; kb < A.proposals; ++kb) {
    const long long kglob = A.proposal_offset + kb;
    const size_t kc = (size_t)kb * C + c;
    const uint32_t cg = (uint32_t)(A.chain_offset + c), kg = (uint32_t)kglob;
    double u_step, u_acc;
    if (A.u_step_in && A.u_accept_in) {
      u_step = A.u_step_in[kc]; u_acc = A.u_accept_in[kc];
    } else {
      constexpr int W = UniformPairCache<TPC>::W;
      if ((kb % W) == 0) ucache.fill(A.seed, cg, kglob);
      ucache.get(kb % W, u_step, u_acc);
      if (A.u_step_in) u_step = A.u_step_in[kc];
      if (A.u_accept_in) u_acc = A.u_accept_in[kc];
    }
    const double eps = A.randomize ? __dmul_rn(u_step, eps0) : eps0;
    if (A.out_stepsize && ln.live && ln.vlead) A.out_stepsize[kc] = eps0;

    if (A.z_in) {
      const double* zr = A.z_in + kc * d;
#pragma unroll
      for (int i = 0; i < 4; ++i) p[i] = ln.active(i) ? zr[cj[i]] : 0.0;
      if (inferv) pv = zr[jv];
    } else {
      // normals are keyed by coordinate pair (coordinates 2m, 2m+1 come from pair m)
      if constexpr (NP == 4) {
        normal_pair(A.seed, cg, kg, (uint32_t)(2 * ln.e), p[0], p[1]);
        normal_pair(A.seed, cg, kg, (uint32_t)(2 * ln.e + 1), p[2], p[3]);
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          p[i] = 0.0;
          if (ln.active(i)) {
            double z0, z1;
            normal_pair(A.seed, cg, kg, (uint32_t)(cj[i] >> 1), z0, z1);
            p[i] = (cj[i] & 1) ? z1 : z0;
          }
        }
      }
      if (inferv) {
        double z0, z1;
        normal_pair(A.seed, cg, kg, (uint32_t)(jv >> 1), z0, z1);
        pv = (jv & 1) ? z1 : z0;
      }
    }
    double k0 = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      q[i] = qc[i];
      if (!ln.active(i)) continue;
      if (T.sqrtm) p[i] = __dmul_rn(__ldg(T.sqrtm + cj[i]), p[i]);
      if (ln.lead) k0 = __dadd_rn(k0, kinetic_term(T, cj[i], p[i]));
    }
    if (inferv) {
      qv = qcv;
      if (T.sqrtm) pv = __dmul_rn(__ldg(T.sqrtm + jv), pv);
      if (ln.vlead) k0 = __dadd_rn(k0, kinetic_term(T, jv, pv));
    }

    int gi = 0;
    auto mom = [&](double cb) {
      unsigned oob = 0;
      if (grad_checks) oob = srcloc_violations<TPC, LPE, NP>(T, L, ln, red, q, qv);
      double g[4], gvel;
      srcloc_total_gradient<TPC, LPE, NP>(T, L, M, ln, red, q, qv, oob, g, gvel);
      if (A.trace_q && ln.live) {
        const size_t o = (((size_t)kb * A.S.grads_per_proposal + gi) * C + c) * d;
        if (ln.lead) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (ln.active(i)) { A.trace_q[o + cj[i]] = q[i]; A.trace_g[o + cj[i]] = g[i]; }
        }
        if (inferv && ln.vlead) { A.trace_q[o + jv] = qv; A.trace_g[o + jv] = gvel; }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (ln.active(i)) momentum_update(cb, g[i], p[i]);
      if (inferv) momentum_update(cb, gvel, pv);
      ++gi;
    };
    auto pos = [&](double ca) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (ln.active(i)) position_update(T, cj[i], ca, q[i], p[i]);
      if (inferv) position_update(T, jv, ca, qv, pv);
    };
    run_schedule(A.S, eps, mom, pos);

    // energies: kinetic, prior misfit, travel-time misfit
    double k1 = 0.0, u1 = 0.0;
    if (ln.lead) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (!ln.active(i)) continue;
        k1 = __dadd_rn(k1, kinetic_term(T, cj[i], p[i]));
        u1 = __dadd_rn(u1, prior_misfit(T, cj[i], q[i]));
      }
    }
    if (inferv && ln.vlead) {
      k1 = __dadd_rn(k1, kinetic_term(T, jv, pv));
      u1 = __dadd_rn(u1, prior_misfit(T, jv, qv));
    }
    const double inv_vel = __ddiv_rn(1.0, inferv ? qv : L.velocity);
    double lik = srcloc_misfit_partial<LPE>(M, L.stations, ln.e, ln.sub, q[0], q[1], q[2], q[3], inv_vel);
    if (!ln.has_event) lik = 0.0;
    red.sum3(k0, k1, u1);
    double z0 = 0.0, z1 = 0.0;
    red.sum3(lik, z0, z1);
    const unsigned oob = srcloc_violations<TPC, LPE, NP>(T, L, ln, red, q, qv);
    // BayesRule order: prior misfit first, then the likelihood, then the bounds
    double x1 = __dadd_rn(__dadd_rn(u1, T.const_sum), __dmul_rn(0.5, lik));
    if (oob) x1 = __dadd_rn(x1, CUDART_INF);
    const double h0 = __dadd_rn(x, __dmul_rn(0.5, k0));
    const double h1 = __dadd_rn(x1, __dmul_rn(0.5, k1));
    const bool acc = metropolis_accept(h0, h1, u_acc);
    if (A.tune.enabled) eps0 = autotune_stepsize(A.tune, eps0, h0, h1, kglob);

    if (ln.live) {
      if (A.out_q_prop) {
        if (ln.lead) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (ln.active(i)) { A.out_q_prop[kc * d + cj[i]] = q[i]; A.out_p_prop[kc * d + cj[i]] = p[i]; }
        }
        if (inferv && ln.vlead) { A.out_q_prop[kc * d + jv] = qv; A.out_p_prop[kc * d + jv] = pv; }
      }
      if (ln.vlead) {
        if (A.out_accept) A.out_accept[kc] = acc ? 1 : 0;
        if (A.out_h0) A.out_h0[kc] = h0;
        if (A.out_h1) A.out_h1[kc] = h1;
      }
    }
    if (acc) {
#pragma unroll
      for (int i = 0; i < 4; ++i) qc[i] = q[i];
      qcv = qv; x = x1; ++accepted;
    }
    const bool store_now = kglob == next_store;
    if (store_now) next_store += A.thinning;
    if (A.out_samples && ln.live && store_now) {
      const size_t srow = (store_row * C + c) * (size_t)(d + 1);
      if (ln.lead) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (ln.active(i)) A.out_samples[srow + cj[i]] = qc[i];
      }
      if (ln.vlead) {
        if (inferv) A.out_samples[srow + jv] = qcv;
        A.out_samples[srow + d] = x;
      }
    }
    if (store_now) ++store_row;
  }

  if (ln.live) {
    if (ln.lead) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (ln.active(i)) A.q[row + cj[i]] = qc[i];
    }
    if (ln.vlead) {
      if (inferv) A.q[row + jv] = qcv;
      A.x[c] = x;
      if (A.accepted_total) A.accepted_total[c] += accepted;
      if (A.tune.enabled && A.stepsize_chain) A.stepsize_chain[c] = eps0;
    }
  }
}

// mode 0: x[c] = misfit ; mode 1: g[c,:] = gradient
template <int TPC, int LPE, int NP>
__global__ void __launch_bounds__((TPC <= 32 ? SRCLOC_SMALL_BLOCK : TPC))
srcloc_eval_kernel(const DevTarget T, const SrcLocDev L, int chains, int mode,
                   const double* __restrict__ qin, double* __restrict__ out) {
  extern __shared__ __align__(16) double dyn_smem[];
  __shared__ double scratch[ChainReduce<TPC>::scratch_doubles(TPC <= 32 ? SRCLOC_SMALL_BLOCK : TPC)];
  const ChainReduce<TPC> red{scratch};
  const SrcLocShared M = srcloc_stage(L, dyn_smem);
  const int d = T.dims, E = L.events;
  SrcLocLane<TPC, LPE, NP> ln;
  ln.init(chains, E);
  const size_t row = (size_t)ln.c * d;
  int cj[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) cj[i] = ln.coord(i);
  constexpr int kNP = NP;
  const int jv = kNP * E;
  double q[4], qv = 0.0;
#pragma unroll
  for (int i = 0; i < 4; ++i) q[i] = ln.active(i) ? qin[row + cj[i]] : 0.0;
  if (L.infer_velocity) qv = qin[row + jv];
  const unsigned oob = srcloc_violations<TPC, LPE, NP>(T, L, ln, red, q, qv);
  if (mode == 0) {
    double u1 = 0.0, z0 = 0.0, z1 = 0.0;
    if (ln.lead)
      for (int i = 0; i < 4; ++i)
        if (ln.active(i)) u1 = __dadd_rn(u1, prior_misfit(T, cj[i], q[i]));
    if (L.infer_velocity && ln.vlead) u1 = __dadd_rn(u1, prior_misfit(T, jv, qv));
    const double inv_vel = __ddiv_rn(1.0, L.infer_velocity ? qv : L.velocity);
    double lik = srcloc_misfit_partial<LPE>(M, L.stations, ln.e, ln.sub, q[0], q[1], q[2], q[3], inv_vel);
    if (!ln.has_event) lik = 0.0;
    red.sum3(u1, lik, z0);
    (void)z1;
    double x1 = __dadd_rn(__dadd_rn(u1, T.const_sum), __dmul_rn(0.5, lik));
    if (oob) x1 = __dadd_rn(x1, CUDART_INF);
    if (ln.live && ln.vlead) out[ln.c] = x1;
  } else {
    double g[4], gvel;
    srcloc_total_gradient<TPC, LPE, NP>(T, L, M, ln, red, q, qv, oob, g, gvel);
    if (ln.live) {
      if (ln.lead)
        for (int i = 0; i < 4; ++i)
          if (ln.active(i)) out[row + cj[i]] = g[i];
      if (L.infer_velocity && ln.vlead) out[row + jv] = gvel;
    }
  }
}

}  // namespace hmcb

// common.cuh -- device-side building blocks shared by every HMC kernel.
//
//  * DevTarget: flattened target distribution (per-coordinate tables of elementwise
//    priors, bound checks, reflection bounds, mass matrix) passed by value to kernels.
//  * Elementwise arithmetic follows the reference's operation order and is written with
//    __dmul_rn/__dadd_rn/__dsub_rn so that nvcc never contracts it into FMAs: on separable
//    targets a trajectory is bit-identical to numpy's.
//  * Philox4x32-10 counter RNG keyed by (seed; chain, proposal, pair, stream) so that the
//    draws do not depend on how chains are distributed over GPUs or thread blocks.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

// bound checks are tracked per chain as bits of one 32-bit word
#define HMCB_MAX_CHECKS 32

namespace hmcb {

// Flattened target distribution as per-coordinate tables (built once by hmcb_finalize):
//   * prior terms: slot t of coordinate j holds the t-th elementwise prior covering j
//     (kind 0 = none, 1 = Normal: a = mean, b = inverse variance; 2 = Laplace: a = mean,
//     b = inverse dispersion), in the order the reference sums them;
//   * bound checks: check k has lb/ub rows with -inf/+inf where it does not apply, and
//     c_cover[j] has bit k set when coordinate j lies inside check k's range (that is
//     where a fired check adds +inf to the gradient);
//   * reflection bounds and the diagonal mass matrix.
struct DevTarget {
  int dims;
  int n_terms;   // max number of prior terms covering one coordinate
  int n_checks;
  int uniform_kind;  // TERM_* when every coordinate has exactly one prior term of this kind, else 0
  unsigned grad_check_mask;  // bit k set: check k adds +inf to the gradient on its range
  double const_sum;          // sum of the priors' normalisation constants
  const unsigned char* t_kind;  // [n_terms x dims]
  const double* t_a;            // [n_terms x dims]
  const double* t_b;            // [n_terms x dims]
  const double* c_lb;           // [n_checks x dims]
  const double* c_ub;           // [n_checks x dims]
  const unsigned* c_cover;      // [dims]
  const double* refl_lb;  // [dims] or null
  const double* refl_ub;  // [dims] or null
  const double* invm;     // [dims] 1/diagonal, null = unit mass
  const double* sqrtm;    // [dims] sqrt(diagonal), null = unit mass
};

enum { TERM_NONE = 0, TERM_NORMAL = 1, TERM_LAPLACE = 2 };

// ---------------------------------------------------------------- elementwise pieces ---

__device__ __forceinline__ double sign_np(double v) {
  // numpy.sign: -1, 0, +1, nan
  return (v > 0.0) ? 1.0 : ((v < 0.0) ? -1.0 : ((v == 0.0) ? 0.0 : v));
}

// d(chi)/dq of one prior term (base.py:564-570, 703-710)
__device__ __forceinline__ double term_gradient(int kind, double a, double b, double q) {
  return (kind == TERM_NORMAL) ? __dmul_rn(-b, __dsub_rn(a, q))           // -icov * (mu - q)
                               : __dmul_rn(sign_np(__dsub_rn(q, a)), b);  // sign(q - mu) * idisp
}
// misfit contribution of one prior term, constants excluded (base.py:539-550, 689-700):
// Normal 0.5*r*(icov*r), Laplace |q-mu|*idisp.
__device__ __forceinline__ double term_misfit(int kind, double a, double b, double q) {
  if (kind == TERM_NORMAL) {
    const double d = __dsub_rn(a, q);
    return __dmul_rn(0.5, __dmul_rn(d, __dmul_rn(b, d)));
  }
  return __dmul_rn(fabs(__dsub_rn(q, a)), b);
}

// Sum over prior terms of d(chi)/dq_j at coordinate j, plus +inf when a gradient-visible
// bound check covering j has fired somewhere in this chain (oob_mask).
__device__ __forceinline__ double prior_gradient(const DevTarget& T, int j, double q,
                                                 unsigned oob_mask) {
  double g = 0.0;
  for (int t = 0; t < T.n_terms; ++t) {
    const size_t o = (size_t)t * T.dims + j;
    const int kind = __ldg(T.t_kind + o);
    if (kind) g = __dadd_rn(g, term_gradient(kind, __ldg(T.t_a + o), __ldg(T.t_b + o), q));
  }
  oob_mask &= T.grad_check_mask;
  if (oob_mask && (oob_mask & __ldg(T.c_cover + j))) g = __dadd_rn(g, CUDART_INF);
  return g;
}

// Contribution of coordinate j to the prior misfit, constants excluded.
__device__ __forceinline__ double prior_misfit(const DevTarget& T, int j, double q) {
  double s = 0.0;
  for (int t = 0; t < T.n_terms; ++t) {
    const size_t o = (size_t)t * T.dims + j;
    const int kind = __ldg(T.t_kind + o);
    if (kind) s = __dadd_rn(s, term_misfit(kind, __ldg(T.t_a + o), __ldg(T.t_b + o), q));
  }
  return s;
}

// Bit k set iff coordinate j violates bound check k (misfit_bounds, base.py:361-374).
// NaN compares false on both sides, like numpy.
__device__ __forceinline__ unsigned bound_violations(const DevTarget& T, int j, double q) {
  unsigned m = 0;
  for (int k = 0; k < T.n_checks; ++k) {
    const size_t o = (size_t)k * T.dims + j;
    if (q < __ldg(T.c_lb + o) || q > __ldg(T.c_ub + o)) m |= (1u << k);
  }
  return m;
}

// One-shot mirror reflection (base.py:258-270): the upper test sees the corrected q.
__device__ __forceinline__ void reflect_on(double lb, double ub, double& q, double& p) {
  if (q < lb) { q = __dadd_rn(q, __dmul_rn(2.0, __dsub_rn(lb, q))); p = -p; }
  if (q > ub) { q = __dadd_rn(q, __dmul_rn(2.0, __dsub_rn(ub, q))); p = -p; }
}
__device__ __forceinline__ void reflect(const DevTarget& T, int j, double& q, double& p) {
  if (T.refl_lb) {
    const double lb = __ldg(T.refl_lb + j);
    if (q < lb) { q = __dadd_rn(q, __dmul_rn(2.0, __dsub_rn(lb, q))); p = -p; }
  }
  if (T.refl_ub) {
    const double ub = __ldg(T.refl_ub + j);
    if (q > ub) { q = __dadd_rn(q, __dmul_rn(2.0, __dsub_rn(ub, q))); p = -p; }
  }
}

// dK/dp_j (MassMatrices.py:116-133, 201-218)
__device__ __forceinline__ double kinetic_gradient(const DevTarget& T, int j, double p) {
  return T.invm ? __dmul_rn(__ldg(T.invm + j), p) : p;
}
// p_j * dK/dp_j ; K = 0.5 * sum (MassMatrices.py:100-114, 185-199)
__device__ __forceinline__ double kinetic_term(const DevTarget& T, int j, double p) {
  return __dmul_rn(p, kinetic_gradient(T, j, p));
}
// q += coeff * dK/dp ; reflect
__device__ __forceinline__ void position_update(const DevTarget& T, int j, double coeff,
                                                double& q, double& p) {
  q = __dadd_rn(q, __dmul_rn(coeff, kinetic_gradient(T, j, p)));
  reflect(T, j, q, p);
}
// p -= coeff * g
__device__ __forceinline__ void momentum_update(double coeff, double g, double& p) {
  p = __dsub_rn(p, __dmul_rn(coeff, g));
}

// Metropolis test (Samplers.py:1481-1486): exp(H0 - H1) > u, NaN compares false.
__device__ __forceinline__ bool metropolis_accept(double h0, double h1, double u) {
  return exp(__dsub_rn(h0, h1)) > u;
}

// Step-size autotuning (HMC.autotune, Samplers.py:1494-1522) after proposal k (global index).
struct AutotuneArgs {
  int enabled;
  double target, learning_rate;
};
__device__ __forceinline__ double autotune_stepsize(const AutotuneArgs& at, double stepsize, double h0,
                                                    double h1, long long k) {
  double rate = exp(__dsub_rn(h0, h1));
  if (rate != rate) rate = 0.0;
  const double weight = pow((double)(k + 1), -at.learning_rate);
  stepsize = __dsub_rn(stepsize, __dmul_rn(weight, __dsub_rn(at.target, fmin(rate, 1.0))));
  if (stepsize <= 0.0) stepsize = fmax(stepsize, 1e-18);
  return stepsize;
}

// -------------------------------------------------------------------------- Philox ---

struct Philox4 { uint32_t x, y, z, w; };

__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2,
                                                 uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return Philox4{c0, c1, c2, c3};
}

__device__ __forceinline__ double u53(uint32_t hi, uint32_t lo) {
  // top 53 bits of the 64-bit word (hi:lo) -> [0,1)
  const uint64_t w = ((uint64_t)hi << 32) | lo;
  return (double)(w >> 11) * 1.1102230246251565e-16;  // 2^-53
}

enum : uint32_t { STREAM_NORMAL = 0u, STREAM_UNIFORM = 1u };

// ---- Box-Muller transcendentals ---------------------------------------------------------
// The momentum draw is about a third of the fused kernels' instructions, so the three
// transcendentals are specialised to the ranges Box-Muller needs (no special-case paths)
// with their coefficients in constant memory, where an FMA reads them as an operand instead
// of materialising 64-bit immediates.  Accuracy: ln <= 1 ulp, sin/cos <= 2.3e-16 absolute,
// sqrt <= 1 ulp (checked against numpy in tests/test_gpu_rng.py via a CPU replica).

// fdlibm e_log.c minimax coefficients for log((1+s)/(1-s)) - 2s on |s| < 0.1716
static __constant__ double kLg[7] = {6.666666666666735130e-01, 3.999999999940941908e-01,
                                     2.857142874366239149e-01, 2.222219843214978396e-01,
                                     1.818357216161805012e-01, 1.531383769920937332e-01,
                                     1.479819860511658591e-01};
// Taylor coefficients of sin(pi f) / f and cos(pi f) in f^2, |f| <= 1/4
static __constant__ double kSinPi[9] = {3.141592653589793,      -5.16771278004997,      2.5501640398773455,
                                        -0.5992645293207921,    0.08214588661112823,    -0.0073704309457143504,
                                        0.00046630280576761255, -2.1915353447830217e-05, 7.952054001475513e-07};
static __constant__ double kCosPi[10] = {1.0,
                                         -4.934802200544679,
                                         4.0587121264167685,
                                         -1.3352627688545895,
                                         0.2353306303588932,
                                         -0.02580689139001406,
                                         0.0019295743094039231,
                                         -0.0001046381049248457,
                                         4.303069587032947e-06,
                                         -1.3878952462213771e-07};

__device__ __forceinline__ double rcp_seed(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  return r;
}
__device__ __forceinline__ double rsqrt_seed(double x) {
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  return r;
}

// ln(x) for a normal x in (0, 1] (x >= 2^-53 here)
__device__ __forceinline__ double log_unit_interval(double x) {
  int hi = __double2hiint(x);
  const int lo = __double2loint(x);
  int e = (hi >> 20) - 1023;
  hi = (hi & 0x000fffff) | 0x3ff00000;                 // mantissa m in [1, 2)
  if (hi >= 0x3ff6a09f) { hi -= 0x00100000; ++e; }     // m >= ~sqrt(2): use m / 2
  const double m = __hiloint2double(hi, lo);
  const double f = m - 1.0;
  const double den = 2.0 + f;
  double r = rcp_seed(den);
  r = fma(r, fma(-den, r, 1.0), r);
  r = fma(r, fma(-den, r, 1.0), r);
  double sq = f * r;
  sq = fma(fma(-sq, den, f), r, sq);                   // s = f / (2 + f)
  const double z = sq * sq;
  double R = kLg[6];
  R = fma(R, z, kLg[5]); R = fma(R, z, kLg[4]); R = fma(R, z, kLg[3]);
  R = fma(R, z, kLg[2]); R = fma(R, z, kLg[1]); R = fma(R, z, kLg[0]);
  R *= z;
  const double hfsq = 0.5 * f * f;
  const double k = (double)e;
  const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10;
  return k * ln2_hi - ((hfsq - fma(sq, hfsq + R, k * ln2_lo)) - f);
}

// sqrt(t) for t >= 0 (finite)
__device__ __forceinline__ double sqrt_nonneg(double t) {
  double y = rsqrt_seed(t);
  const double h = 0.5 * t;
  y = fma(y, fma(-(h * y), y, 0.5), y);
  y = fma(y, fma(-(h * y), y, 0.5), y);
  double r = t * y;
  r = fma(fma(-r, r, t), 0.5 * y, r);
  return (t > 0.0) ? r : 0.0;
}

// sin(2 pi u), cos(2 pi u) for u in [0, 1)
__device__ __forceinline__ void sincos_2pi(double u, double& sn, double& cs) {
  const double x = 2.0 * u;                            // angle = pi * x, x in [0, 2)
  const int q = __double2int_rn(2.0 * x);              // nearest multiple of 1/2
  const double f = fma((double)q, -0.5, x);            // |f| <= 1/4, exact
  const double f2 = f * f;
  double sp = kSinPi[8];
#pragma unroll
  for (int k = 7; k >= 0; --k) sp = fma(sp, f2, kSinPi[k]);
  sp *= f;
  double cp = kCosPi[9];
#pragma unroll
  for (int k = 8; k >= 0; --k) cp = fma(cp, f2, kCosPi[k]);
  const double a = (q & 1) ? cp : sp;                  // quadrant rotation
  const double b = (q & 1) ? sp : cp;
  sn = (q & 2) ? -a : a;
  cs = ((q + 1) & 2) ? -b : b;
}

// Standard normals for coordinates (2*pair, 2*pair+1) of `chain` at `proposal`.
__device__ __forceinline__ void normal_pair(uint64_t seed, uint32_t chain, uint32_t proposal,
                                            uint32_t pair, double& z0, double& z1) {
  const Philox4 r = philox4x32_10(chain, proposal, pair, STREAM_NORMAL, (uint32_t)seed,
                                  (uint32_t)(seed >> 32));
  const double u1 = 1.0 - u53(r.x, r.y);  // (0,1]
  const double u2 = u53(r.z, r.w);        // [0,1)
  const double rad = sqrt_nonneg(-2.0 * log_unit_interval(u1));
  double s, c;
  sincos_2pi(u2, s, c);
  z0 = rad * c;
  z1 = rad * s;
}

// (step-size factor in [0.5,1.5), acceptance uniform in [0,1)) of `chain` at `proposal`.
__device__ __forceinline__ void uniform_pair(uint64_t seed, uint32_t chain, uint32_t proposal,
                                             double& u_step, double& u_acc) {
  const Philox4 r = philox4x32_10(chain, proposal, 0u, STREAM_UNIFORM, (uint32_t)seed,
                                  (uint32_t)(seed >> 32));
  u_step = 0.5 + u53(r.x, r.y);
  u_acc = u53(r.z, r.w);
}

// The (step-size factor, acceptance uniform) pair is one Philox call per chain and
// proposal.  Instead of every thread of the chain repeating it, lane l of a group of
// W = min(TPC, 32) lanes evaluates the pair of proposal base + l, and each proposal then
// fetches its pair with two shuffles: one Philox call per W proposals per warp.
template <int TPC>
struct UniformPairCache {
  static constexpr int W = TPC < 32 ? TPC : 32;
  double us, ua;
  __device__ __forceinline__ void fill(uint64_t seed, uint32_t chain, long long first_proposal) {
    uniform_pair(seed, chain, (uint32_t)(first_proposal + (threadIdx.x % W)), us, ua);
  }
  __device__ __forceinline__ void get(int slot, double& u_step, double& u_acc) const {
    const int src = ((threadIdx.x & 31) & ~(W - 1)) + slot;
    u_step = __shfl_sync(0xffffffffu, us, src);
    u_acc = __shfl_sync(0xffffffffu, ua, src);
  }
};

// ---------------------------------------------------------------- chain reductions ---
// A chain is owned by TPC consecutive threads (TPC a power of two).  TPC <= 32: the group
// lives inside one warp and reduces with xor-shuffles (every lane ends with the same
// bits because each butterfly level adds the same two numbers on both partners).
// TPC > 32: the chain spans TPC/32 whole warps of the block (several chains may share a
// block); warps reduce, then every thread sums its chain's per-warp partials in warp
// order through a per-chain shared-memory segment.  Summation order is fixed =>
// deterministic.

template <int TPC>
struct ChainReduce {
  static constexpr int kWarps = (TPC + 31) / 32;
  // shared-memory doubles a block of BLOCK threads needs (TPC > 32 only)
  __host__ __device__ static constexpr int scratch_doubles(int block) { return TPC > 32 ? 3 * kWarps * (block / TPC) + 1 : 1; }
  double* scratch;

  __device__ __forceinline__ double* mine() const { return scratch + (threadIdx.x / TPC) * (3 * kWarps); }

  __device__ __forceinline__ void sum3(double& a, double& b, double& c) const {
    constexpr int W = TPC < 32 ? TPC : 32;
#pragma unroll
    for (int off = W / 2; off > 0; off >>= 1) {
      a = __dadd_rn(a, __shfl_xor_sync(0xffffffffu, a, off));
      b = __dadd_rn(b, __shfl_xor_sync(0xffffffffu, b, off));
      c = __dadd_rn(c, __shfl_xor_sync(0xffffffffu, c, off));
    }
    if constexpr (TPC > 32) {
      double* s = mine();
      const int warp = (threadIdx.x % TPC) >> 5;  // warp index inside the chain
      if ((threadIdx.x & 31) == 0) {
        s[warp] = a; s[kWarps + warp] = b; s[2 * kWarps + warp] = c;
      }
      __syncthreads();
      a = 0.0; b = 0.0; c = 0.0;
#pragma unroll 4
      for (int w = 0; w < kWarps; ++w) {
        a = __dadd_rn(a, s[w]);
        b = __dadd_rn(b, s[kWarps + w]);
        c = __dadd_rn(c, s[2 * kWarps + w]);
      }
      __syncthreads();
    }
  }

  __device__ __forceinline__ unsigned any_bits(unsigned m) const {
    constexpr int W = TPC < 32 ? TPC : 32;
#pragma unroll
    for (int off = W / 2; off > 0; off >>= 1) m |= __shfl_xor_sync(0xffffffffu, m, off);
    if constexpr (TPC > 32) {
      unsigned* s = reinterpret_cast<unsigned*>(mine());
      const int warp = (threadIdx.x % TPC) >> 5;
      if ((threadIdx.x & 31) == 0) s[warp] = m;
      __syncthreads();
      m = 0;
      for (int w = 0; w < kWarps; ++w) m |= s[w];
      __syncthreads();
    }
    return m;
  }
};

// ------------------------------------------------------------ integrator schedule ---
// A trajectory is: [lone position update]* then pairs (momentum update, position update);
// every momentum sub-step of lf/3s/4s is followed by a position sub-step
// (Samplers.py:1524-1584, 1586-1661, 1663-1726).  Coefficients are multiples of the
// chain's step size eps.

struct StageOp {
  double b;   // momentum coefficient multiplier (used iff has_b)
  double a;   // position coefficient multiplier
  int has_b;  // 1: gradient + momentum update + position update; 0: position update only
  int pad;
};

struct Schedule {
  int n_pre, n_body, reps, n_post;
  int grads_per_proposal;
  int kind;  // 0 lf, 1 3s, 2 4s (HMCB_INTEGRATOR_*)
  StageOp pre[2];
  StageOp body[6];
  StageOp post[2];
};

// Runs one trajectory with the integrator's stage sequence laid out at compile time, so
// the per-stage work is straight-line code: `pos(ca)` is a position sub-step
// q += ca * dK/dp (+ reflection), `mom(cb)` a momentum sub-step p -= cb * grad(q).
// Coefficients are the host's multipliers times the chain's step size, formed exactly as
// the reference forms them (Samplers.py:1539, 1562-1569, 1588-1603, 1666-1679).
template <class Mom, class Pos>
__device__ __forceinline__ void run_schedule(const Schedule& S, double eps, Mom&& mom, Pos&& pos) {
  if (S.kind == 0) {
    const double half = __dmul_rn(S.pre[0].a, eps);
    const double cb = __dmul_rn(S.body[0].b, eps), ca = __dmul_rn(S.body[0].a, eps);
    pos(half);
    for (int r = 0; r < S.reps; ++r) { mom(cb); pos(ca); }
    mom(__dmul_rn(S.post[0].b, eps));
    pos(__dmul_rn(S.post[0].a, eps));
  } else if (S.kind == 1) {
    const double a1 = __dmul_rn(S.body[0].a, eps), b1 = __dmul_rn(S.body[1].b, eps);
    const double a2 = __dmul_rn(S.body[1].a, eps), b2 = __dmul_rn(S.body[2].b, eps);
    for (int r = 0; r < S.reps; ++r) {
      pos(a1); mom(b1); pos(a2); mom(b2); pos(a2); mom(b1); pos(a1);
    }
  } else {
    const double a1 = __dmul_rn(S.body[0].a, eps), b1 = __dmul_rn(S.body[1].b, eps);
    const double a2 = __dmul_rn(S.body[1].a, eps), b2 = __dmul_rn(S.body[2].b, eps);
    const double a3 = __dmul_rn(S.body[2].a, eps);
    for (int r = 0; r < S.reps; ++r) {
      pos(a1); mom(b1); pos(a2); mom(b2); pos(a3); mom(b2); pos(a2); mom(b1); pos(a1);
    }
  }
}

}  // namespace hmcb

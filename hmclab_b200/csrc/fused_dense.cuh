// fused_dense.cuh -- whole-proposal fused HMC kernel for the dense LinearMatrix likelihood in
// its premultiplied form (GtG q - Gtd0, LinearMatrix.py:164-206) when the model is small
// enough for GtG to live in shared memory (dims <= 128): the shape of BASELINE.json's config 1
// and of most dense problems the reference is used for.
//
// One block = 32 chains for the whole block of proposals.  GtG (zero padded to 128 x 128)
// stays resident in shared memory; the chains' positions are the B operand of a
// 128 x 32 x 128 fp64 tensor-core product (DMMA m8n8k4) per gradient evaluation, and the
// chain state (q, q_current, p) lives in registers *in the accumulator fragment layout*, so
// the momentum / position update, the prior gradient, the bounds reflection and the energy
// partial sums are applied to the accumulators in place: HBM sees q once in, once out and the
// stored sample rows, like the priors-only kernel.
//
// Replaces, for these targets, the staged pipeline's ~25 launches per proposal.
#pragma once
#include "common.cuh"
#include "fused.cuh"
#include "gemm.cuh"

namespace hmcb {

constexpr int FD_M = 128;             // padded dims
constexpr int FD_BN = 32;             // chains per block
constexpr int FD_THREADS = 256;       // 8 warps, warp w owns rows [16 w, 16 w + 16)
constexpr int FD_LDA = FD_M + 4;      // 132 doubles: conflict-free A fragments
constexpr int FD_LDB = FD_BN + 4;     // 36 doubles: conflict-free B fragments
constexpr int FD_RED = 5;             // k0, k1, prior misfit, likelihood misfit, bound flags
// (+ 256 B static: the gradient's bound flags)
constexpr size_t FD_SMEM_BYTES =
    sizeof(double) * ((size_t)FD_M * FD_LDA + (size_t)FD_M * FD_LDB + (size_t)FD_RED * 8 * FD_BN + 4 * FD_BN) +
    sizeof(int) * FD_BN;

struct FusedDenseArgs {
  FusedArgs F;
  const double* GtG;   // [128 x 128] zero padded, row-major
  const double* Gtd0;  // [dims]
  double dtd;
};

__global__ void __launch_bounds__(FD_THREADS, 1)
hmc_fused_dense_kernel(const FusedDenseArgs D) {
  extern __shared__ __align__(16) double fd_smem[];
  double* As = fd_smem;                          // GtG
  double* Bs = As + FD_M * FD_LDA;               // positions of the block's chains, [dims x chains]
  double* red = Bs + FD_M * FD_LDB;              // [FD_RED][8 warps][32 chains]
  double* eps_s = red + FD_RED * 8 * FD_BN;      // per-chain step size of this proposal
  double* uacc_s = eps_s + FD_BN;
  double* x_s = uacc_s + FD_BN;                  // current misfit
  double* x1_s = x_s + FD_BN;                    // proposed misfit (scratch)
  int* acc_s = reinterpret_cast<int*>(x1_s + FD_BN);
  __shared__ unsigned gflags[2][FD_BN];          // per-chain bound violations seen by the gradient

  const FusedArgs& A = D.F;
  const DevTarget& T = A.T;
  const int d = T.dims;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int c0 = blockIdx.x * FD_BN;             // first chain of the block
  const size_t C = (size_t)A.chains;

  // ---- operands -------------------------------------------------------------------------
  for (int idx = tid; idx < FD_M * FD_M; idx += FD_THREADS)
    As[(idx >> 7) * FD_LDA + (idx & 127)] = D.GtG[idx];

  // fragment coordinates: rows r[i] = 16 w + lane/4 + 8 i ; columns cc[j][h] = 8 j + 2 (lane%4) + h
  int r[2];
  r[0] = 16 * warp + (lane >> 2);
  r[1] = r[0] + 8;
  const int cb = 2 * (lane & 3);
  auto col = [&](int j, int h) { return 8 * j + cb + h; };
  auto chain_ok = [&](int j, int h) { return c0 + col(j, h) < A.chains; };

  // per-row constants (one prior term per coordinate at most; rows >= dims are padding)
  int kind[2];
  double ta[2], tb[2], gtd[2], im[2], sm[2], rlb[2], rub[2];
  bool row_ok[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    row_ok[i] = r[i] < d;
    const int j = row_ok[i] ? r[i] : 0;
    kind[i] = (row_ok[i] && T.n_terms) ? (int)T.t_kind[j] : TERM_NONE;
    ta[i] = (row_ok[i] && T.n_terms) ? T.t_a[j] : 0.0;
    tb[i] = (row_ok[i] && T.n_terms) ? T.t_b[j] : 0.0;
    gtd[i] = row_ok[i] ? D.Gtd0[j] : 0.0;
    im[i] = (row_ok[i] && T.invm) ? T.invm[j] : 1.0;
    sm[i] = (row_ok[i] && T.sqrtm) ? T.sqrtm[j] : 1.0;
    rlb[i] = (row_ok[i] && T.refl_lb) ? T.refl_lb[j] : -CUDART_INF;
    rub[i] = (row_ok[i] && T.refl_ub) ? T.refl_ub[j] : CUDART_INF;
  }
  const bool has_mass = T.invm != nullptr;
  const bool has_refl = T.refl_lb != nullptr || T.refl_ub != nullptr;
  const bool grad_checks = T.grad_check_mask != 0u;
  unsigned cover[2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
    cover[i] = (grad_checks && row_ok[i]) ? ((unsigned)T.c_cover[r[i]] & T.grad_check_mask) : 0u;
  if (tid < 2 * FD_BN) gflags[tid / FD_BN][tid % FD_BN] = 0u;
  int gbuf = 0;

  double qc[2][4][2], q[2][4][2], p[2][4][2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int h = 0; h < 2; ++h)
        qc[i][j][h] = (row_ok[i] && chain_ok(j, h)) ? A.q[(size_t)(c0 + col(j, h)) * d + r[i]] : 0.0;

  // per-chain scalars are owned by threads 0..31 (thread t <-> chain c0 + t)
  const bool owner = tid < FD_BN;
  const bool owner_live = owner && (c0 + tid < A.chains);
  double eps0 = 0.0;
  int accepted = 0;
  if (owner) {
    x_s[tid] = owner_live ? A.x[c0 + tid] : 0.0;
    eps0 = (owner_live && A.stepsize_chain) ? A.stepsize_chain[c0 + tid] : A.stepsize;
  }

  auto store_positions = [&](const double (&v)[2][4][2]) {
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        *reinterpret_cast<double2*>(Bs + r[i] * FD_LDB + col(j, 0)) = make_double2(v[i][j][0], v[i][j][1]);
  };
  // Y = GtG * Q for the block's chains; acc[i][j][h] = Y[r[i]][col(j, h)]
  auto gemm = [&](double (&acc)[2][4][2]) {
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    const double* a0 = As + r[0] * FD_LDA + (lane & 3);
    const double* b0 = Bs + (lane & 3) * FD_LDB + (lane >> 2);
#pragma unroll 8
    for (int kk = 0; kk < FD_M; kk += 4) {
      const double a_0 = a0[kk], a_1 = a0[8 * FD_LDA + kk];
      double b[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = b0[kk * FD_LDB + 8 * j];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        dmma_8x8x4(acc[0][j][0], acc[0][j][1], a_0, b[j]);
        dmma_8x8x4(acc[1][j][0], acc[1][j][1], a_1, b[j]);
      }
    }
  };
  auto dkdp = [&](int i, double pv) { return has_mass ? __dmul_rn(im[i], pv) : pv; };

  // sums over the rows of this thread, per column -> sums over all rows via shuffles and the
  // per-warp partial table `red`; the owners add the 8 warp partials in warp order
  auto reduce_columns = [&](int slot, const double (&v)[4][2]) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        double s = v[j][h];
        s = __dadd_rn(s, __shfl_xor_sync(0xffffffffu, s, 4));
        s = __dadd_rn(s, __shfl_xor_sync(0xffffffffu, s, 8));
        s = __dadd_rn(s, __shfl_xor_sync(0xffffffffu, s, 16));
        if ((lane >> 2) == 0) red[(slot * 8 + warp) * FD_BN + col(j, h)] = s;
      }
  };
  auto column_total = [&](int slot) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) s = __dadd_rn(s, red[(slot * 8 + w) * FD_BN + tid]);
    return s;
  };

  long long next_store = (A.proposal_offset + A.thinning - 1) / A.thinning * A.thinning;
  size_t store_row = 0;
  __syncthreads();

  for (int kb = 0; kb < A.proposals; ++kb) {
    const long long kglob = A.proposal_offset + kb;

    // ---- per-chain draws (owners) -----------------------------------------------------
    if (owner) {
      const size_t kc = (size_t)kb * C + (c0 + tid);
      double u_step = 1.0, u_acc = 0.0;
      if (owner_live && A.u_step_in) u_step = A.u_step_in[kc];
      if (owner_live && A.u_accept_in) u_acc = A.u_accept_in[kc];
      if (!(A.u_step_in && A.u_accept_in)) {
        double us, ua;
        uniform_pair(A.seed, (uint32_t)(A.chain_offset + c0 + tid), (uint32_t)kglob, us, ua);
        if (!A.u_step_in) u_step = us;
        if (!A.u_accept_in) u_acc = ua;
      }
      eps_s[tid] = A.randomize ? __dmul_rn(u_step, eps0) : eps0;
      uacc_s[tid] = u_acc;
      if (A.out_stepsize && owner_live) A.out_stepsize[kc] = eps0;
    }

    // ---- momentum draw ----------------------------------------------------------------
    double k0[4][2];
#pragma unroll
    for (int j = 0; j < 4; ++j) k0[j][0] = k0[j][1] = 0.0;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        double z[2];
        if (A.z_in) {
#pragma unroll
          for (int h = 0; h < 2; ++h)
            z[h] = (row_ok[i] && chain_ok(j, h))
                       ? A.z_in[((size_t)kb * C + (c0 + col(j, h))) * d + r[i]] : 0.0;
        } else {
          // rows 2t and 2t+1 sit in lanes l and l^4: the even-row lane draws the pair of chain
          // col(j,0), the odd-row lane the pair of chain col(j,1), and they swap one normal
          const bool odd = (r[i] & 1) != 0;
          double z0, z1;
          normal_pair(A.seed, (uint32_t)(A.chain_offset + c0 + col(j, odd ? 1 : 0)), (uint32_t)kglob,
                      (uint32_t)(r[i] >> 1), z0, z1);
          const double give = odd ? z0 : z1;
          const double got = __shfl_xor_sync(0xffffffffu, give, 4);
          z[0] = odd ? got : z0;   // coordinate r[i] of chain col(j,0)
          z[1] = odd ? z1 : got;   // coordinate r[i] of chain col(j,1)
          if (!row_ok[i]) z[0] = z[1] = 0.0;
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          double pv = has_mass ? __dmul_rn(sm[i], z[h]) : z[h];
          if (!row_ok[i]) pv = 0.0;
          p[i][j][h] = pv;
          q[i][j][h] = qc[i][j][h];
          k0[j][h] = __dadd_rn(k0[j][h], __dmul_rn(pv, dkdp(i, pv)));
        }
      }
    __syncthreads();  // eps_s / uacc_s visible

    double eps[4][2];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int h = 0; h < 2; ++h) eps[j][h] = eps_s[col(j, h)];

    // ---- trajectory -------------------------------------------------------------------
    int gi = 0;
    auto pos = [&](double a_mult) {
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            q[i][j][h] = __dadd_rn(q[i][j][h], __dmul_rn(__dmul_rn(a_mult, eps[j][h]), dkdp(i, p[i][j][h])));
            if (has_refl) reflect_on(rlb[i], rub[i], q[i][j][h], p[i][j][h]);
          }
    };
    auto mom = [&](double b_mult) {
      __syncthreads();           // previous product has finished reading Bs
      store_positions(q);
      if (grad_checks) {
        // misfit_bounds of priors / containers adds +inf to the gradient of every coordinate of
        // a violated check's range (base.py:361-374, 564-570): per-chain flags through shared memory
        if (tid < FD_BN) gflags[gbuf ^ 1][tid] = 0u;   // the buffer of the next evaluation
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            unsigned m = 0;
#pragma unroll
            for (int i = 0; i < 2; ++i)
              if (row_ok[i]) m |= bound_violations(T, r[i], q[i][j][h]);
            if (m) atomicOr(&gflags[gbuf][col(j, h)], m);
          }
      }
      __syncthreads();
      double y[2][4][2];
      gemm(y);
      unsigned oob[4][2];
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int h = 0; h < 2; ++h) oob[j][h] = grad_checks ? gflags[gbuf][col(j, h)] : 0u;
      gbuf ^= 1;
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const double lik = __dsub_rn(y[i][j][h], gtd[i]);
            double pg = kind[i] ? __dadd_rn(0.0, term_gradient(kind[i], ta[i], tb[i], q[i][j][h])) : 0.0;
            if (oob[j][h] & cover[i]) pg = __dadd_rn(pg, CUDART_INF);
            const double g = __dadd_rn(pg, lik);
            if (A.trace_q && row_ok[i] && chain_ok(j, h)) {
              const size_t o = (((size_t)kb * A.S.grads_per_proposal + gi) * C + (c0 + col(j, h))) * d + r[i];
              A.trace_q[o] = q[i][j][h];
              A.trace_g[o] = g;
            }
            if (row_ok[i]) momentum_update(__dmul_rn(b_mult, eps[j][h]), g, p[i][j][h]);
          }
      ++gi;
    };
    // run_schedule multiplies its coefficients by a common eps; here eps differs per column,
    // so it is given 1.0 and the lambdas apply each chain's own step size (same products:
    // (multiplier * 1.0) * eps_c == multiplier * eps_c exactly)
    run_schedule(A.S, 1.0, mom, pos);

    // ---- energies ---------------------------------------------------------------------
    __syncthreads();
    store_positions(q);
    __syncthreads();
    double y[2][4][2];
    gemm(y);
    double k1[4][2], u1[4][2], lk[4][2], fl[4][2];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        k1[j][h] = u1[j][h] = lk[j][h] = fl[j][h] = 0.0;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          if (!row_ok[i]) continue;
          const double qv = q[i][j][h], pv = p[i][j][h];
          k1[j][h] = __dadd_rn(k1[j][h], __dmul_rn(pv, dkdp(i, pv)));
          if (kind[i]) u1[j][h] = __dadd_rn(u1[j][h], term_misfit(kind[i], ta[i], tb[i], qv));
          // LinearMatrix.py:185-191: m^T (GtG m - 2 Gtd0)
          lk[j][h] = __dadd_rn(lk[j][h], __dmul_rn(qv, __dsub_rn(y[i][j][h], __dmul_rn(2.0, gtd[i]))));
          if (T.n_checks && bound_violations(T, r[i], qv)) fl[j][h] = 1.0;
        }
      }
    reduce_columns(0, k0);
    reduce_columns(1, k1);
    reduce_columns(2, u1);
    reduce_columns(3, lk);
    if (T.n_checks) reduce_columns(4, fl);
    __syncthreads();
    if (owner) {
      const double K0 = column_total(0), K1 = column_total(1), U1 = column_total(2), L1 = column_total(3);
      const bool oob = T.n_checks ? column_total(4) > 0.0 : false;
      const double lik = __dmul_rn(0.5, __dadd_rn(L1, D.dtd));
      double x1 = __dadd_rn(__dadd_rn(U1, T.const_sum), lik);
      if (oob) x1 = __dadd_rn(x1, CUDART_INF);
      const double x0 = x_s[tid];
      const double h0 = __dadd_rn(x0, __dmul_rn(0.5, K0));
      const double h1 = __dadd_rn(x1, __dmul_rn(0.5, K1));
      const bool acc = metropolis_accept(h0, h1, uacc_s[tid]);
      if (A.tune.enabled) eps0 = autotune_stepsize(A.tune, eps0, h0, h1, kglob);
      acc_s[tid] = acc ? 1 : 0;
      if (acc) { x_s[tid] = x1; ++accepted; }
      if (owner_live) {
        const size_t kc = (size_t)kb * C + (c0 + tid);
        if (A.out_accept) A.out_accept[kc] = acc ? 1 : 0;
        if (A.out_h0) A.out_h0[kc] = h0;
        if (A.out_h1) A.out_h1[kc] = h1;
      }
    }
    __syncthreads();

    // ---- state update and outputs -----------------------------------------------------
    const bool store_now = kglob == next_store;
    if (store_now) next_store += A.thinning;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int cl = col(j, h);
          if (A.out_q_prop && row_ok[i] && chain_ok(j, h)) {
            const size_t o = ((size_t)kb * C + (c0 + cl)) * d + r[i];
            A.out_q_prop[o] = q[i][j][h];
            A.out_p_prop[o] = p[i][j][h];
          }
          if (acc_s[cl]) qc[i][j][h] = q[i][j][h];
          if (A.out_samples && store_now && row_ok[i] && chain_ok(j, h))
            A.out_samples[(store_row * C + (c0 + cl)) * (size_t)(d + 1) + r[i]] = qc[i][j][h];
        }
    if (A.out_samples && store_now && owner_live)
      A.out_samples[(store_row * C + (c0 + tid)) * (size_t)(d + 1) + d] = x_s[tid];
    if (store_now) ++store_row;
    __syncthreads();  // acc_s / red are rewritten by the next proposal
  }

#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int h = 0; h < 2; ++h)
        if (row_ok[i] && chain_ok(j, h)) A.q[(size_t)(c0 + col(j, h)) * d + r[i]] = qc[i][j][h];
  if (owner_live) {
    A.x[c0 + tid] = x_s[tid];
    if (A.accepted_total) A.accepted_total[c0 + tid] += accepted;
    if (A.tune.enabled && A.stepsize_chain) A.stepsize_chain[c0 + tid] = eps0;
  }
}

}  // namespace hmcb

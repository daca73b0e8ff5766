// hmcb.cu -- the C ABI of include/hmcb.h: engine object, target lowering to device
// constants, path selection and the host-side orchestration of a block of proposals.
//
// No arithmetic of the hot path happens on the host: this file validates, uploads model
// constants once, and enqueues kernels on the caller's stream.
#include "../../include/hmcb.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "launch.cuh"
#include "rwmh.cuh"

using namespace hmcb;

namespace {

thread_local std::string g_error;

int fail(const std::string& msg) {
  g_error = msg;
  return -1;
}

#define HMCB_CUDA(expr)                                                                  \
  do {                                                                                   \
    cudaError_t err__ = (expr);                                                          \
    if (err__ != cudaSuccess)                                                            \
      return fail(std::string(#expr) + " failed: " + cudaGetErrorString(err__));         \
  } while (0)

#define HMCB_CHECK(cond, msg) \
  do {                        \
    if (!(cond)) return fail(msg); \
  } while (0)

enum LikKind { LK_NONE = 0, LK_DENSE_PREMULT, LK_DENSE_DIRECT, LK_CSR_DIRECT, LK_CSR_PREMULT, LK_SRCLOC };

int round_up(int64_t v, int m) { return (int)((v + m - 1) / m * m); }

struct HostPrior {
  int kind;
  int64_t offset, len;
  std::vector<double> a, b;
  double constant;
};
struct HostCheck {
  int64_t offset, len;
  bool has_lb, has_ub;
  std::vector<double> lb, ub;
  int in_gradient;
};
struct HostCsr {
  int64_t rows = 0, cols = 0, nnz = 0;
  std::vector<int32_t> indptr, indices;
  std::vector<double> data;
};

}  // namespace

struct hmcb_engine {
  int device = 0;
  int64_t C = 0, d = 0;
  int integrator = HMCB_INTEGRATOR_LF;
  int steps = 10;
  bool mass_diag = false;
  std::vector<double> h_diag, h_invdiag;
  bool mass_full = false;              // MassMatrices.Full: p = L z, dK/dp = M^-1 p as GEMMs over the batch
  std::vector<double> h_L, h_Minv;     // [d x d] row-major: lower Cholesky factor, inverse
  double *dL = nullptr, *dMinv = nullptr, *v_w = nullptr;
  std::vector<HostPrior> priors;
  std::vector<HostCheck> checks;
  bool has_rlb = false, has_rub = false;
  std::vector<double> rlb, rub;

  int lik = LK_NONE;
  int64_t N = 0;                       // data dimension (direct forms)
  std::vector<double> h_A, h_At;       // dense: GtG or G ; Gt
  bool has_At = false;
  std::vector<double> h_Amis, h_vecmis;  // dense data covariance, direct form: U G and U d of the misfit pass
  double *dAmis = nullptr, *dvecmis = nullptr;
  // Ozaki-sliced tcgen05 path of the dense direct products (ozaki.cuh)
  bool oz = false;
  int oz_saG = 0, oz_saGt = 0;         // int8 digits of G / G^T (what their rows need, at most the orders kept)
  int oz_sb = OZ_SLICES_B, oz_orders = OZ_NUM_ORDERS;   // digits of the chain batch, orders kept
  int64_t oz_rows = 0;                                  // padded rows of the forward operator (npad, or dpad when premultiplied)
  // modular variant for G^T r (long contraction): residue planes of G^T, of the chain batch, of the product
  bool oz_crt = false;
  signed char *oz_crtA = nullptr, *oz_crtB = nullptr, *oz_crtC = nullptr;
  int* oz_crt_ea = nullptr;
  CUtensorMap oz_mapCrtA, oz_mapCrtB, oz_mapCrtBh;
  signed char *oz_AG = nullptr, *oz_AGt = nullptr, *oz_B = nullptr;   // int8 slices of G, G^T, the chain batch
  int *oz_eaG = nullptr, *oz_eaGt = nullptr, *oz_C = nullptr;         // row exponents, int32 order planes
  unsigned long long *oz_maxQ = nullptr, *oz_maxR = nullptr;          // per-chain max |.| (bit patterns)
  CUtensorMap oz_mapAG, oz_mapAGt, oz_mapBq, oz_mapBr, oz_mapBqh, oz_mapBrh, oz_mapCq, oz_mapCr;
  std::vector<double> h_vec, h_var, h_sigma;  // Gtd0 or d ; var ; sigma
  double dtd = 0.0;
  HostCsr csr, csr_t;
  // srcloc
  int64_t events = 0, stations = 0;
  int srcloc_np = 4;  // 4: SourceLocation3D, 3: SourceLocation2D
  int infer_velocity = 0;
  double velocity = 0.0;
  std::vector<double> h_rx, h_ry, h_rz, h_tobs, h_std;

  bool finalized = false;
  bool exact = false;      // hmcb_set_exact_arithmetic
  int path = -1;
  bool fused_dense = false;  // run_block uses the whole-proposal small-dense kernel
  int64_t launches = 0;

  // device state ---------------------------------------------------------------------
  std::vector<void*> allocs;
  DevTarget T{};
  Schedule S{};
  std::vector<StageOp> ops;  // flattened trajectory
  SrcLocDev L{};
  // staged path
  int ld = 0, dpad = 0, npad = 0, jtiles = 0, ltiles = 0;
  double *dA_rowmajor = nullptr;  // [128 x 128] GtG for the fused small-dense kernel
  double *dA = nullptr, *dAt = nullptr, *dvec = nullptr, *dvar = nullptr, *dsigma = nullptr;
  CsrDev csr_dev{}, csr_t_dev{};
  StripDev strip_dev{}, strip_t_dev{};  // shared-memory staged SpMM tables (spmm_strip.cuh)
  bool use_strips = false;
  CUtensorMap tmap_q[2], tmap_R;        // B operands of the strip SpMM: the two position planes and R
  double *q_cur = nullptr, *q_w[2] = {nullptr, nullptr}, *p_w = nullptr, *R = nullptr;
  double *eps = nullptr, *uacc = nullptr, *k0part = nullptr, *k1part = nullptr, *upart = nullptr,
         *lpart = nullptr;
  unsigned* flags[3] = {nullptr, nullptr, nullptr};
  unsigned char* accbuf = nullptr;
  double *rw_qp = nullptr, *rw_x1 = nullptr;  // hmcb_run_block_rwmh scratch
  // hmcb_sample_host: streams, events and device buffers are created once and reused
  cudaStream_t s_compute = nullptr, s_copy = nullptr;
  cudaEvent_t ev_produced[2] = {nullptr, nullptr}, ev_drained[2] = {nullptr, nullptr};
  double *sh_q = nullptr, *sh_x = nullptr, *sh_buf[2] = {nullptr, nullptr};
  int32_t* sh_acc = nullptr;
  size_t sh_buf_doubles = 0;
  // hmcb_kernel_timing_*: CUDA event pairs recorded around the launches of the dominant kernels
  bool ktiming = false;
  std::vector<cudaEvent_t> kev[2];   // class 0: gradient-pass / whole-block kernels, 1: misfit pass
  size_t kev_used[2] = {0, 0};
};

namespace {

template <class Tv>
int dev_alloc(hmcb_engine* e, size_t count, Tv** out, bool zero = true) {
  void* p = nullptr;
  const size_t bytes = std::max<size_t>(count, 1) * sizeof(Tv);
  HMCB_CUDA(cudaMalloc(&p, bytes));
  e->allocs.push_back(p);
  if (zero) HMCB_CUDA(cudaMemset(p, 0, bytes));
  *out = static_cast<Tv*>(p);
  return 0;
}

template <class Tv>
int dev_upload(hmcb_engine* e, const std::vector<Tv>& v, const Tv** out) {
  Tv* p = nullptr;
  if (dev_alloc(e, v.size(), &p, false)) return -1;
  if (!v.empty()) HMCB_CUDA(cudaMemcpy(p, v.data(), v.size() * sizeof(Tv), cudaMemcpyHostToDevice));
  *out = p;
  return 0;
}

// rows x cols host matrix -> zero padded rows_pad x cols_pad device matrix
int dev_upload_padded(hmcb_engine* e, const double* src, int64_t rows, int64_t cols, int64_t rows_pad,
                      int64_t cols_pad, double** out) {
  double* p = nullptr;
  if (dev_alloc(e, (size_t)rows_pad * cols_pad, &p, true)) return -1;
  HMCB_CUDA(cudaMemcpy2D(p, (size_t)cols_pad * sizeof(double), src, (size_t)cols * sizeof(double),
                         (size_t)cols * sizeof(double), (size_t)rows, cudaMemcpyHostToDevice));
  *out = p;
  return 0;
}

// rows x cols host matrix -> tile-major device copy for the DMMA GEMM's bulk-copy staging:
// [rows_pad/128][cols_pad/16] tiles of GEMM_BM x GEMM_LDA_S doubles (zero padded)
int dev_upload_tiled(hmcb_engine* e, const double* src, int64_t rows, int64_t cols, int64_t rows_pad,
                     int64_t cols_pad, double** out) {
  const int64_t mt = rows_pad / GEMM_BM, kt = cols_pad / GEMM_BK;
  std::vector<double> host((size_t)mt * kt * GEMM_A_STAGE, 0.0);
  for (int64_t i = 0; i < rows; ++i) {
    const int64_t tm = i / GEMM_BM, r = i % GEMM_BM;
    for (int64_t j = 0; j < cols; ++j) {
      const int64_t tk = j / GEMM_BK, c = j % GEMM_BK;
      host[(size_t)((tm * kt + tk) * GEMM_A_STAGE + r * GEMM_LDA_S + c)] = src[(size_t)i * cols + j];
    }
  }
  double* p = nullptr;
  if (dev_alloc(e, host.size(), &p, false)) return -1;
  HMCB_CUDA(cudaMemcpy(p, host.data(), host.size() * sizeof(double), cudaMemcpyHostToDevice));
  *out = p;
  return 0;
}

// Event pair around one launch of a dominant kernel (only while hmcb_kernel_timing_begin is active):
// the events go on the launching stream, so they time the kernel and nothing else.
struct KernelTimer {
  hmcb_engine* e; cudaStream_t s; int cls; bool on = false;
  static constexpr size_t kMaxPairs = 16384;
  KernelTimer(hmcb_engine* e_, cudaStream_t s_, int cls_) : e(e_), s(s_), cls(cls_) {
    if (!e->ktiming || e->kev_used[cls] + 2 > 2 * kMaxPairs) return;
    if (mark()) on = true;
  }
  ~KernelTimer() { if (on) mark(); }
  bool mark() {
    std::vector<cudaEvent_t>& v = e->kev[cls];
    size_t& used = e->kev_used[cls];
    if (used == v.size()) {
      cudaEvent_t ev;
      if (cudaEventCreate(&ev) != cudaSuccess) return false;
      v.push_back(ev);
    }
    cudaEventRecord(v[used++], s);
    return true;
  }
};

void free_device(hmcb_engine* e) {
  for (void* p : e->allocs) cudaFree(p);
  e->allocs.clear();
  e->finalized = false;
  // every pointer into the freed pool is reset; hmcb_finalize / first use allocates again
  e->fused_dense = false;
  e->rw_qp = e->rw_x1 = nullptr;
  e->dA = e->dAt = e->dA_rowmajor = e->dvec = e->dvar = e->dsigma = nullptr;
  e->dL = e->dMinv = e->v_w = nullptr;
  e->dAmis = e->dvecmis = nullptr;
  e->oz = false;
  e->oz_AG = e->oz_AGt = e->oz_B = nullptr; e->oz_eaG = e->oz_eaGt = e->oz_C = nullptr;
  e->oz_maxQ = e->oz_maxR = nullptr;
  e->q_cur = e->q_w[0] = e->q_w[1] = e->p_w = e->R = nullptr;
  e->eps = e->uacc = e->k0part = e->k1part = e->upart = e->lpart = nullptr;
  e->flags[0] = e->flags[1] = e->flags[2] = nullptr;
  e->accbuf = nullptr;
  e->N = 0;
}

int check_csr(const HostCsr& m, const char* what) {
  HMCB_CHECK((int64_t)m.indptr.size() == m.rows + 1, std::string(what) + ": indptr size");
  HMCB_CHECK(m.indptr[0] == 0 && m.indptr[m.rows] == m.nnz, std::string(what) + ": indptr range");
  for (int64_t i = 0; i < m.rows; ++i)
    HMCB_CHECK(m.indptr[i] <= m.indptr[i + 1], std::string(what) + ": indptr not monotone");
  for (int64_t k = 0; k < m.nnz; ++k)
    HMCB_CHECK(m.indices[k] >= 0 && m.indices[k] < m.cols, std::string(what) + ": column index out of range");
  return 0;
}

void copy_vec(std::vector<double>& dst, const double* src, int64_t n) { dst.assign(src, src + n); }

// Samplers.py:1524-1584 (lf), :1663-1726 (3s), :1586-1661 (4s); multipliers of eps.
void build_schedule(hmcb_engine* e) {
  Schedule& S = e->S;
  std::memset(&S, 0, sizeof(S));
  const int Lsteps = e->steps;
  S.kind = e->integrator;
  auto lone = [](double a) { return StageOp{0.0, a, 0, 0}; };
  auto pair = [](double b, double a) { return StageOp{b, a, 1, 0}; };
  if (e->integrator == HMCB_INTEGRATOR_LF) {
    S.n_pre = 1; S.pre[0] = lone(0.5);
    S.n_body = 1; S.body[0] = pair(1.0, 1.0); S.reps = Lsteps - 1;
    S.n_post = 1; S.post[0] = pair(1.0, 0.5);
    S.grads_per_proposal = Lsteps;
  } else if (e->integrator == HMCB_INTEGRATOR_3S) {
    const double a1 = 0.11888010966548, a2 = 1.0 / 2.0 - a1;
    const double b1 = 0.29619504261126, b2 = 1.0 - 2.0 * b1;
    S.n_body = 4; S.reps = Lsteps;
    S.body[0] = lone(a1); S.body[1] = pair(b1, a2); S.body[2] = pair(b2, a2); S.body[3] = pair(b1, a1);
    S.grads_per_proposal = 3 * Lsteps;
  } else {
    const double a1 = 0.071353913450279725904, a2 = 0.268548791161230105820;
    const double a3 = 1.0 - 2.0 * a1 - 2.0 * a2;
    const double b1 = 0.1916678, b2 = 1.0 / 2.0 - b1;
    S.n_body = 5; S.reps = Lsteps;
    S.body[0] = lone(a1); S.body[1] = pair(b1, a2); S.body[2] = pair(b2, a3);
    S.body[3] = pair(b2, a2); S.body[4] = pair(b1, a1);
    S.grads_per_proposal = 4 * Lsteps;
  }
  e->ops.clear();
  for (int s = 0; s < S.n_pre; ++s) e->ops.push_back(S.pre[s]);
  for (int r = 0; r < S.reps; ++r)
    for (int s = 0; s < S.n_body; ++s) e->ops.push_back(S.body[s]);
  for (int s = 0; s < S.n_post; ++s) e->ops.push_back(S.post[s]);
}

int upload_csr(hmcb_engine* e, const HostCsr& m, CsrDev* out) {
  const int32_t *ip = nullptr, *ix = nullptr;
  const double* dv = nullptr;
  if (dev_upload(e, m.indptr, &ip) || dev_upload(e, m.indices, &ix) || dev_upload(e, m.data, &dv)) return -1;
  out->indptr = ip; out->indices = ix; out->data = dv;
  out->rows = (int)m.rows;
  // Row chunks (one per block, 4 warps interleaving its rows) are sized so that the chain
  // slabs of the gathered operand in flight at any time fit in L2 together: a slab is
  // cols x 256 B, about 148 SMs x 16 blocks are resident, slabs in flight = resident / chunks.
  const double slab_bytes = (double)m.cols * SPMM_SLAB * sizeof(double);
  const double min_chunks = std::max(1.0, 2368.0 * slab_bytes / (48.0 * 1024 * 1024));
  int64_t rpc = (int64_t)((double)m.rows / min_chunks);
  rpc = std::max<int64_t>(SPMM_WARPS, std::min<int64_t>(64, rpc / SPMM_WARPS * SPMM_WARPS));
  out->rows_per_chunk = (int)rpc;
  out->chunks = (int)((m.rows + rpc - 1) / rpc);
  return 0;
}

int env_int(const char* name, int fallback) {
  const char* v = std::getenv(name);
  return (v && *v) ? std::atoi(v) : fallback;
}

// Host form of the tables of the shared-memory staged SpMM (spmm_strip.cuh, spmm_types.cuh).
struct StripHost {
  std::vector<unsigned char> ent;      // packed (chunk, strip) groups, 16-byte aligned each
  std::vector<SpmmStrip> strips;
  std::vector<int32_t> strip_ptr;      // [chunks + 1]
  int chunks = 0, cstride = 0, kb_box = 0, compact = 0;
};

// Per row chunk of RB rows the columns are dealt round-robin into T strips (column c -> strip
// c % T, local row c / T); per (chunk, strip) group: a header with the first slot of every warp,
// then per warp the nonzeros of its RW rows in row order and a sentinel.  Rows are sorted by
// column and duplicate entries summed first (scipy allows both).
int build_strip_tables(const HostCsr& m, const SpmmShape& shape, int kb, int emax, bool allow_compact,
                       StripHost* host) {
  const int S = 32 * shape.cpl, RB = shape.warps * shape.rw;
  const int64_t rows = m.rows, cols = m.cols;
  std::vector<int32_t> ip(rows + 1, 0), ix;
  std::vector<double> dv;
  ix.reserve(m.nnz); dv.reserve(m.nnz);
  {
    std::vector<std::pair<int32_t, double>> row;
    for (int64_t i = 0; i < rows; ++i) {
      row.clear();
      for (int32_t k = m.indptr[i]; k < m.indptr[i + 1]; ++k) row.emplace_back(m.indices[k], m.data[k]);
      std::stable_sort(row.begin(), row.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
      for (size_t k = 0; k < row.size(); ++k) {
        if (k && row[k].first == row[k - 1].first) dv.back() += row[k].second;
        else { ix.push_back(row[k].first); dv.push_back(row[k].second); }
      }
      ip[i + 1] = (int32_t)ix.size();
    }
  }
  const int chunks = (int)((rows + RB - 1) / RB);
  const int W = shape.warps, RW = shape.rw;
  // compact 8-byte nonzeros when every value survives a round trip through fp32 (the reference
  // rounds G to numpy.single by default) and HMCB_SPMM_COMPACT does not forbid it
  bool compact = allow_compact && (int64_t)kb * S * 8 < (1 << 24) && RW < 255;
  for (size_t k = 0; compact && k < dv.size(); ++k) compact = (double)(float)dv[k] == dv[k];
  const int esz = compact ? 8 : 16;                       // bytes per slot
  const int hdr = ((W * 4 + 15) / 16) * 16 / esz;         // header slots: first slot of every warp
  const int cap = emax * 16 / esz;                        // slots a stage can hold
  const int budget = cap - hdr - W - 1;   // left for nonzeros: header, one sentinel per warp, padding
  std::vector<SpmmStrip>& strips = host->strips;
  std::vector<int32_t>& strip_ptr = host->strip_ptr;
  std::vector<unsigned char>& ent = host->ent;
  strips.clear(); ent.clear(); strip_ptr.assign(chunks + 1, 0);
  ent.reserve((ix.size() + ix.size() / 4) * esz);
  // strips per chunk: column c belongs to strip c % T (local index c / T); T grows until the
  // fullest (chunk, strip) group fits in the slot budget
  int64_t T = std::max<int64_t>(1, (cols + kb - 1) / kb);
  std::vector<int32_t> cnt;
  for (;;) {
    cnt.assign((size_t)chunks * T, 0);
    int32_t worst = 0;
    for (int64_t i = 0; i < rows; ++i)
      for (int32_t k = ip[i]; k < ip[i + 1]; ++k) worst = std::max(worst, ++cnt[(size_t)(i / RB) * T + ix[k] % T]);
    if (worst <= budget) break;
    if (T >= cols) return fail("internal: SpMM strip does not fit");
    T = std::min<int64_t>(cols, T + T / 4 + 1);
  }
  std::vector<std::vector<SpmmEntry>> bucket(T);   // per strip, per warp streams built in one pass
  std::vector<int32_t> wstart((size_t)T * W);
  for (int b = 0; b < chunks; ++b) {
    const int64_t r0 = (int64_t)b * RB, r1 = std::min<int64_t>(rows, r0 + RB);
    for (auto& v : bucket) v.clear();
    for (int w = 0; w < W; ++w) {
      for (int64_t t = 0; t < T; ++t) wstart[t * W + w] = (int32_t)bucket[t].size();
      for (int rr = 0; rr < RW; ++rr) {
        const int64_t i = r0 + w * RW + rr;
        if (i >= r1) break;
        for (int32_t k = ip[i]; k < ip[i + 1]; ++k)
          bucket[ix[k] % T].push_back(SpmmEntry{dv[k], (int)((ix[k] / T) * S * 8), rr});
      }
      for (int64_t t = 0; t < T; ++t) bucket[t].push_back(SpmmEntry{0.0, 0, RW});
    }
    for (int64_t t = 0; t < T; ++t) {
      if ((int)bucket[t].size() == W) continue;   // only sentinels: nothing to do in this strip
      const size_t base = ent.size();
      const size_t bytes = (((size_t)hdr + bucket[t].size()) * esz + 15) / 16 * 16;
      if (bytes > (size_t)emax * 16) return fail("internal: strip nonzero count mismatch");
      ent.resize(base + bytes, 0);
      int32_t* first = reinterpret_cast<int32_t*>(&ent[base]);
      for (int w = 0; w < W; ++w) first[w] = hdr + wstart[t * W + w];
      unsigned char* dst = &ent[base + (size_t)hdr * esz];
      if (compact) {
        SpmmEntry32* o = reinterpret_cast<SpmmEntry32*>(dst);
        for (size_t k = 0; k < bucket[t].size(); ++k)
          o[k] = SpmmEntry32{(float)bucket[t][k].val, (unsigned)bucket[t][k].off | ((unsigned)bucket[t][k].row << 24)};
      } else {
        std::memcpy(dst, bucket[t].data(), bucket[t].size() * sizeof(SpmmEntry));
      }
      strips.push_back(SpmmStrip{(int)t, (int)((cols - t + T - 1) / T), (int)(base / 16), (int)(bytes / 16)});
    }
    strip_ptr[b + 1] = (int32_t)strips.size();
  }
  if ((cols + T - 1) / T > kb) return fail("internal: SpMM strip wider than its buffer");
  HMCB_CHECK(ent.size() / 16 < (size_t)1 << 31, "CSR matrix too large for the strip tables");
  host->chunks = chunks; host->cstride = (int)T; host->kb_box = (int)((cols + T - 1) / T);
  host->compact = compact ? 1 : 0;
  return 0;
}

// ---- row-blocked form ------------------------------------------------------------------------
// Rows sorted by column with duplicates summed (scipy allows unsorted rows and split entries).
void canonical_csr(const HostCsr& m, std::vector<int32_t>* ip, std::vector<int32_t>* ix, std::vector<double>* dv) {
  ip->assign(m.rows + 1, 0); ix->clear(); dv->clear();
  ix->reserve(m.nnz); dv->reserve(m.nnz);
  std::vector<std::pair<int32_t, double>> row;
  for (int64_t i = 0; i < m.rows; ++i) {
    row.clear();
    for (int32_t k = m.indptr[i]; k < m.indptr[i + 1]; ++k) row.emplace_back(m.indices[k], m.data[k]);
    std::stable_sort(row.begin(), row.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
    for (size_t k = 0; k < row.size(); ++k) {
      if (k && row[k].first == row[k - 1].first) dv->back() += row[k].second;
      else { ix->push_back(row[k].first); dv->push_back(row[k].second); }
    }
    (*ip)[i + 1] = (int32_t)ix->size();
  }
}

// Groups of R rows with similar column sets, so that one gathered B row feeds several rows of a
// group.  Greedy: the seed is the first unassigned row, its R-1 partners are the unassigned rows
// that share the most columns with it (counted through the transpose).  order: [groups x R]
// original row indices, -1 = padding.  Deterministic.
void cluster_rows(int64_t rows, int64_t cols, const std::vector<int32_t>& ip, const std::vector<int32_t>& ix,
                  int R, std::vector<int32_t>* order) {
  std::vector<int32_t> tptr(cols + 1, 0), tidx(ix.size());
  for (int32_t j : ix) ++tptr[j + 1];
  for (int64_t j = 0; j < cols; ++j) tptr[j + 1] += tptr[j];
  {
    std::vector<int32_t> fill(tptr.begin(), tptr.end() - 1);
    for (int64_t i = 0; i < rows; ++i)
      for (int32_t k = ip[i]; k < ip[i + 1]; ++k) tidx[fill[ix[k]]++] = (int32_t)i;
  }
  std::vector<char> assigned(rows, 0);
  std::vector<int32_t> cnt(rows, 0), touched, empty_rows;
  order->clear();
  order->reserve(rows + R);
  for (int64_t seed = 0; seed < rows; ++seed) {
    if (assigned[seed]) continue;
    assigned[seed] = 1;
    if (ip[seed] == ip[seed + 1]) { empty_rows.push_back((int32_t)seed); continue; }
    touched.clear();
    for (int32_t k = ip[seed]; k < ip[seed + 1]; ++k) {
      const int32_t j = ix[k];
      for (int32_t t = tptr[j]; t < tptr[j + 1]; ++t) {
        const int32_t i = tidx[t];
        if (!assigned[i] && cnt[i]++ == 0) touched.push_back(i);
      }
    }
    const size_t want = std::min<size_t>(R - 1, touched.size());
    std::partial_sort(touched.begin(), touched.begin() + want, touched.end(), [&](int32_t a, int32_t b) {
      return cnt[a] != cnt[b] ? cnt[a] > cnt[b] : a < b;
    });
    order->push_back((int32_t)seed);
    for (size_t k = 0; k < want; ++k) { assigned[touched[k]] = 1; order->push_back(touched[k]); }
    for (size_t k = want; k + 1 < (size_t)R; ++k) order->push_back(-1);
    for (int32_t i : touched) cnt[i] = 0;
  }
  for (size_t k = 0; k < empty_rows.size(); ++k) order->push_back(empty_rows[k]);
  while (order->size() % R) order->push_back(-1);
}

struct BlockHost {
  std::vector<unsigned char> ent;
  std::vector<SpmmStrip> strips;
  std::vector<int32_t> strip_ptr, perm;
  int chunks = 0, cstride = 0, kb_box = 0;
  int emax = 0;            // 16-byte units a stage reserves for the k-tiles
  int box_rows = 0, row_boxes = 0, compact = 0;
  int64_t blocks = 0, nnz = 0, ktiles = 0;
};

struct BlockShape { int warps, gw, nb; };

// Staging geometry of a strip of `kb` B rows: row_boxes TMA boxes of box_rows rows (a box holds at
// most 256 rows, and a multiple of 8 so that every box starts on a 1024-byte swizzle atom).
void block_box_geometry(int64_t kb, int* box_rows, int* row_boxes) {
  const int64_t nbr = (kb + 255) / 256;
  *row_boxes = (int)nbr;
  *box_rows = (int)(((kb + nbr - 1) / nbr + 7) / 8 * 8);
}

// Row-blocked tensor-core tables (spmm_types.cuh).  warps x gw groups of SPMM_BLOCK_R rows per
// chunk; strips are dealt round-robin over the columns as in build_strip_tables.  A stage of `cap16`
// 16-byte units holds the B rows of a strip and the k-tiles of one (chunk, strip) group: the number
// of strips T is the smallest for which the widest strip and the fullest group fit together.
int build_block_tables(const HostCsr& m, const BlockShape& shape, int cap16, bool allow_compact, BlockHost* host) {
  constexpr int R = SPMM_BLOCK_R;
  const int W = shape.warps, GW = shape.gw, NB = shape.nb, GPC = W * GW;
  const int64_t cols = m.cols;
  std::vector<int32_t> ip, ix;
  std::vector<double> dv;
  canonical_csr(m, &ip, &ix, &dv);
  std::vector<int32_t>& order = host->perm;
  cluster_rows(m.rows, cols, ip, ix, R, &order);
  const int64_t groups = (int64_t)order.size() / R;
  const int chunks = (int)((groups + GPC - 1) / GPC);
  {
    // Balance the warps of a block: a stage of the ring is released when its slowest warp is done, so
    // the warps of a chunk should carry the same work.  Groups are sorted by their number of nonzeros
    // (heavy chunks first: they also run first), and inside a chunk dealt to the warps in serpentine
    // order, so every warp gets one group of each weight class.
    std::vector<int64_t> weight(groups, 0);
    for (int64_t g = 0; g < groups; ++g)
      for (int r = 0; r < R; ++r) {
        const int32_t i = order[g * R + r];
        if (i >= 0) weight[g] += ip[i + 1] - ip[i];
      }
    std::vector<int32_t> rank(groups);
    for (int64_t g = 0; g < groups; ++g) rank[g] = (int32_t)g;
    std::stable_sort(rank.begin(), rank.end(), [&](int32_t a, int32_t b) { return weight[a] > weight[b]; });
    std::vector<int32_t> dealt((size_t)chunks * GPC * R, -1);
    for (int64_t i = 0; i < groups; ++i) {
      const int64_t b = i / GPC, k = i % GPC, g = k / W, w = (g & 1) ? W - 1 - k % W : k % W;
      std::copy(order.begin() + (size_t)rank[i] * R, order.begin() + (size_t)(rank[i] + 1) * R,
                dealt.begin() + (size_t)((b * W + w) * GW + g) * R);
    }
    order.swap(dealt);
  }
  bool compact = allow_compact;
  for (size_t k = 0; compact && k < dv.size(); ++k) compact = (double)(float)dv[k] == dv[k];
  host->compact = compact ? 1 : 0;
  // column blocks of every group slot (chunk, warp, group): distinct columns in ascending order, R values each
  const int64_t slots = (int64_t)chunks * GPC;
  std::vector<int64_t> gptr(slots + 1, 0);
  std::vector<int32_t> gcol;
  std::vector<double> gval;
  gcol.reserve(ix.size() / 2); gval.reserve(ix.size() / 2 * R);
  {
    int32_t cur[R];
    for (int64_t g = 0; g < slots; ++g) {
      for (int r = 0; r < R; ++r) { const int32_t i = order[g * R + r]; cur[r] = i < 0 ? -1 : ip[i]; }
      for (;;) {
        int32_t next = INT32_MAX;
        for (int r = 0; r < R; ++r) {
          const int32_t i = order[g * R + r];
          if (i >= 0 && cur[r] < ip[i + 1]) next = std::min(next, ix[cur[r]]);
        }
        if (next == INT32_MAX) break;
        gcol.push_back(next);
        for (int r = 0; r < R; ++r) {
          const int32_t i = order[g * R + r];
          if (i >= 0 && cur[r] < ip[i + 1] && ix[cur[r]] == next) gval.push_back(dv[cur[r]++]);
          else gval.push_back(0.0);
        }
      }
      gptr[g + 1] = (int64_t)gcol.size();
    }
  }
  host->blocks = (int64_t)gcol.size();
  host->nnz = (int64_t)ix.size();
  const int hdr16 = ((GPC + 1) * 4 + 15) / 16;
  // a k-tile record: per column {B row offset / 16 | row mask << 13 | values before it << 21, values
  // of the k-tile}, then the nonzero values column by column -- fp32 when the matrix allows, else fp64
  const int rec_hdr = 32, vsz = compact ? 4 : 8;
  std::vector<int32_t> ennz(gcol.size(), 0);          // nonzeros of every column block
  for (size_t k = 0; k < gcol.size(); ++k)
    for (int r = 0; r < R; ++r) ennz[k] += gval[k * R + r] != 0.0;
  auto size16 = [&](int64_t tiles, int64_t values) {
    return (int64_t)hdr16 + (tiles * (rec_hdr + (compact ? 4 : 0)) + values * vsz + 15) / 16;   // records are 8-byte aligned
  };
  // strips per chunk: T grows until the widest strip and the fullest (chunk, strip) group fit into a stage
  int64_t T = std::max<int64_t>(1, (cols * NB * 8 + cap16 - 1) / cap16);
  std::vector<int32_t> per_strip, kt, nz;
  int box_rows = 0, row_boxes = 0;
  for (;;) {
    int64_t worst = 0;
    per_strip.assign((size_t)T, 0);
    kt.assign((size_t)T, 0);
    nz.assign((size_t)T, 0);
    for (int b = 0; b < chunks; ++b) {
      std::fill(kt.begin(), kt.end(), 0);
      std::fill(nz.begin(), nz.end(), 0);
      for (int64_t g = (int64_t)b * GPC; g < (int64_t)(b + 1) * GPC; ++g) {
        for (int64_t k = gptr[g]; k < gptr[g + 1]; ++k) { ++per_strip[gcol[k] % T]; nz[gcol[k] % T] += ennz[k]; }
        for (int64_t k = gptr[g]; k < gptr[g + 1]; ++k) {
          int32_t& n = per_strip[gcol[k] % T];
          if (n) { kt[gcol[k] % T] += (n + 3) / 4; n = 0; }
        }
      }
      for (int64_t t = 0; t < T; ++t) worst = std::max<int64_t>(worst, size16(kt[t], nz[t]));
    }
    block_box_geometry((cols + T - 1) / T, &box_rows, &row_boxes);
    const int64_t box16 = (int64_t)box_rows * row_boxes * NB * 8;
    if ((int64_t)box_rows * row_boxes * 128 * NB >= (1 << 17)) return fail("internal: SpMM block stage too large for 13-bit offsets");
    if (std::getenv("HMCB_DEBUG_TABLES"))
      fprintf(stderr, "block tables: T=%lld box16=%lld worst size16=%lld cap=%d\n", (long long)T, (long long)box16,
              (long long)worst, cap16);
    if (box16 + worst <= cap16 && worst * 4 < (1 << 22)) { host->emax = (int)(cap16 - box16); break; }
    if (T >= cols) return fail("internal: SpMM block strip does not fit");
    T = std::min<int64_t>(cols, T + std::max<int64_t>(1, T / 16));
  }
  const int emax = host->emax;
  const uint32_t region = (uint32_t)NB * box_rows * 128u;      // bytes of one row box (NB chain boxes of 128-byte rows)
  std::vector<unsigned char>& ent = host->ent;
  std::vector<SpmmStrip>& strips = host->strips;
  std::vector<int32_t>& strip_ptr = host->strip_ptr;
  ent.clear(); strips.clear(); strip_ptr.assign(chunks + 1, 0);
  // per strip of the current chunk: the record stream, per (warp, group) its first word and k-tile count
  std::vector<std::vector<unsigned char>> rec(T);
  std::vector<uint32_t> first((size_t)T * GPC), count((size_t)T * GPC);
  std::vector<std::vector<int64_t>> lists(T);     // entries of the current group per strip
  std::vector<int64_t> q[4], tile;
  int64_t ktiles = 0;
  for (int b = 0; b < chunks; ++b) {
    for (int64_t t = 0; t < T; ++t) rec[t].clear();
    for (int s = 0; s < GPC; ++s) {
      const int64_t g = (int64_t)b * GPC + s;
      for (int64_t t = 0; t < T; ++t) { first[t * GPC + s] = (uint32_t)(rec[t].size() / 4); count[t * GPC + s] = 0; }
      for (int64_t k = gptr[g]; k < gptr[g + 1]; ++k) lists[gcol[k] % T].push_back(k);
      for (int64_t k0 = gptr[g]; k0 < gptr[g + 1]; ++k0) {
        const int64_t t = gcol[k0] % T;
        std::vector<int64_t>& L = lists[t];
        if (L.empty()) continue;      // this strip of the group has been emitted already
        // k-tiles of 4 columns in column order
        auto emit = [&]() {
          std::vector<unsigned char>& out = rec[t];
          uint32_t words[8] = {0, 0, 0, 0, 0, 0, 0, 0};
          uint32_t before = 0;
          for (int c = 0; c < (int)tile.size(); ++c) {
            const int64_t r = gcol[tile[c]] / T, rl = r % box_rows;
            const uint32_t off = (uint32_t)(r / box_rows) * region + (uint32_t)rl * 128u + (uint32_t)(rl & 7) * 16u;
            uint32_t colmask = 0;
            for (int rr = 0; rr < R; ++rr) colmask |= (gval[tile[c] * R + rr] != 0.0 ? 1u : 0u) << rr;
            words[2 * c] = (off >> 4) | (colmask << 13) | (before << 21);
            before += (uint32_t)__builtin_popcount(colmask);
          }
          for (int c = (int)tile.size(); c < 4; ++c) words[2 * c] = before << 21;
          for (int c = 0; c < 4; ++c) words[2 * c + 1] = before;
          const unsigned char* wb = reinterpret_cast<const unsigned char*>(words);
          out.insert(out.end(), wb, wb + rec_hdr);
          for (int c = 0; c < (int)tile.size(); ++c)
            for (int rr = 0; rr < R; ++rr) {
              const double v = gval[tile[c] * R + rr];
              if (v == 0.0) continue;
              if (compact) {
                const float f = (float)v;
                const unsigned char* fb = reinterpret_cast<const unsigned char*>(&f);
                out.insert(out.end(), fb, fb + 4);
              } else {
                const unsigned char* vb = reinterpret_cast<const unsigned char*>(&v);
                out.insert(out.end(), vb, vb + 8);
              }
            }
          while (out.size() % 8) out.push_back(0);     // the next record starts 8-byte aligned
          ++count[t * GPC + s];
          tile.clear();
        };
        // The lanes of a quarter warp read the same 32-byte column of the 4 gathered rows; with the
        // 128-byte swizzle row r lands in bank group (r >> 1) & 3 of that quarter: a k-tile whose 4 rows
        // fall into 4 different classes loads its B fragment without a bank conflict.  One column of
        // each class while all four last, then the fullest classes first.
        for (auto& v : q) v.clear();
        for (int64_t k : L) q[((gcol[k] / T) % box_rows >> 1) & 3].push_back(k);
        size_t head[4] = {0, 0, 0, 0};
        tile.clear();
        for (;;) {
          int order4[4] = {0, 1, 2, 3};
          std::sort(order4, order4 + 4, [&](int a, int b) {
            const size_t ra = q[a].size() - head[a], rb = q[b].size() - head[b];
            return ra != rb ? ra > rb : a < b;
          });
          if (q[order4[0]].size() == head[order4[0]]) break;      // nothing left
          for (int c = 0; c < 4 && tile.size() < 4; ++c)
            if (head[order4[c]] < q[order4[c]].size()) tile.push_back(q[order4[c]][head[order4[c]]++]);
          // fewer than 4 classes left: go round the classes again (2-way conflicts before 3-way ones)
          for (bool took = true; took && tile.size() < 4;) {
            took = false;
            for (int c = 0; c < 4 && tile.size() < 4; ++c)
              if (head[order4[c]] < q[order4[c]].size()) { tile.push_back(q[order4[c]][head[order4[c]]++]); took = true; }
          }
          emit();
        }
        L.clear();
      }
    }
    for (int64_t t = 0; t < T; ++t) {
      if (rec[t].empty()) continue;
      int64_t n = 0;
      for (int s = 0; s < GPC; ++s) n += count[t * GPC + s];
      const size_t base = ent.size(), bytes = ((size_t)hdr16 * 16 + rec[t].size() + 15) / 16 * 16;
      if (bytes > (size_t)emax * 16) return fail("internal: block strip size mismatch");
      ent.resize(base + bytes, 0);
      uint32_t* hdr = reinterpret_cast<uint32_t*>(&ent[base]);
      for (int s = 0; s < GPC; ++s) {
        if (count[t * GPC + s] >= 1024u || first[t * GPC + s] >= (1u << 22)) return fail("internal: block strip header overflow");
        hdr[s] = (first[t * GPC + s] << 10) | count[t * GPC + s];
      }
      hdr[GPC] = (uint32_t)n;
      std::memcpy(&ent[base + (size_t)hdr16 * 16], rec[t].data(), rec[t].size());
      strips.push_back(SpmmStrip{(int)t, (int)((cols - t + T - 1) / T), (int)(base / 16), (int)(bytes / 16)});
      ktiles += n;
    }
    strip_ptr[b + 1] = (int32_t)strips.size();
  }
  HMCB_CHECK(ent.size() / 16 < (size_t)1 << 31, "CSR matrix too large for the block tables");
  host->chunks = chunks; host->cstride = (int)T; host->kb_box = (int)((cols + T - 1) / T);
  host->box_rows = box_rows; host->row_boxes = row_boxes; host->ktiles = ktiles;
  return 0;
}

int upload_csr_strips(hmcb_engine* e, const HostCsr& m, const SpmmShape& shape, int kb, int emax,
                      int stages, StripDev* out) {
  StripHost host;
  if (build_strip_tables(m, shape, kb, emax, env_int("HMCB_SPMM_COMPACT", 1) != 0, &host)) return -1;
  const int S = 32 * shape.cpl;
  const int64_t rows = m.rows;
  const int chunks = host.chunks;
  const int64_t T = host.cstride;
  const unsigned char* d_ent = nullptr; const SpmmStrip* d_strips = nullptr;
  const int32_t* d_ptr = nullptr;
  if (dev_upload(e, host.ent, &d_ent) || dev_upload(e, host.strips, &d_strips) ||
      dev_upload(e, host.strip_ptr, &d_ptr)) return -1;
  out->ent = d_ent; out->compact = host.compact; out->strips = d_strips; out->strip_ptr = d_ptr;
  out->rows = (int)rows; out->chunks = chunks; out->cstride = (int)T;
  out->warps = shape.warps; out->rw = shape.rw; out->cpl = shape.cpl;
  out->kb = kb; out->emax = emax; out->stages = stages;
  out->b_bytes = kb * S * 8;
  out->stage_bytes = (out->b_bytes + emax * 16 + 127) / 128 * 128;
  out->kb_box = host.kb_box;
  HMCB_CUDA(spmm_strip_init(*out));
  return 0;
}

int upload_csr_blocks(hmcb_engine* e, const BlockHost& host, int rows, const BlockShape& shape, int stages,
                      int cap16, StripDev* out) {
  const unsigned char* d_ent = nullptr; const SpmmStrip* d_strips = nullptr;
  const int32_t *d_ptr = nullptr, *d_perm = nullptr;
  if (dev_upload(e, host.ent, &d_ent) || dev_upload(e, host.strips, &d_strips) ||
      dev_upload(e, host.strip_ptr, &d_ptr) || dev_upload(e, host.perm, &d_perm)) return -1;
  out->ent = d_ent; out->compact = host.compact; out->strips = d_strips; out->strip_ptr = d_ptr;
  out->rows = rows; out->chunks = host.chunks; out->cstride = host.cstride;
  out->warps = shape.warps; out->rw = SPMM_BLOCK_R * shape.gw; out->cpl = 0;
  out->kb = host.kb_box; out->emax = host.emax; out->stages = stages;
  out->b_bytes = host.box_rows * host.row_boxes * shape.nb * 128;
  out->stage_bytes = cap16 * 16;
  out->kb_box = host.kb_box;
  out->blocked = 1; out->perm = d_perm;
  out->gw = shape.gw; out->nb = shape.nb; out->box_rows = host.box_rows; out->row_boxes = host.row_boxes;
  HMCB_CUDA(spmm_strip_init(*out));
  return 0;
}

// Chooses the SpMM path for the CSR likelihoods: the shared-memory staged kernel in the mapping
// and strip limits that measured best on the tomography workload (profiles/spmm_lab_r01.json).
// profiles/tools/spmm_lab.py overrides them through the environment: HMCB_SPMM_SHAPE = 0..n-1 picks
// a thread mapping of launch_spmm.cu (-1: the L2-gather kernel it replaced, kept as the lab's
// baseline), HMCB_SPMM_KB / _EMAX / _STAGES the strip limits.
int upload_csr_both(hmcb_engine* e, const HostCsr& m, CsrDev* gather, StripDev* strip) {
  const SpmmShape* shapes = nullptr;
  const int n_shapes = spmm_strip_shapes(&shapes);
  const int which = env_int("HMCB_SPMM_SHAPE", 0);
  e->use_strips = which >= 0;
  if (!e->use_strips) return upload_csr(e, m, gather);
  HMCB_CHECK(which < n_shapes, "HMCB_SPMM_SHAPE out of range");
  // Row-blocked form first (HMCB_SPMM_BLOCKED: -1 = when it pays, 0 = never, 1 = always): rows are
  // regrouped into groups of 8 with similar column sets; it pays when a gathered B row then feeds
  // enough rows of its group, i.e. when the 9 shared-memory wavefronts of a column block buy more
  // than the 5 per nonzero of the plain strip kernel.
  const int want_blocked = env_int("HMCB_SPMM_BLOCKED", -1);
  if (want_blocked != 0) {
    const BlockShape bs{env_int("HMCB_SPMM_BLOCK_WARPS", 31), env_int("HMCB_SPMM_BLOCK_GW", 4),
                        env_int("HMCB_SPMM_BLOCK_NB", 1)};
    HMCB_CHECK(spmm_block_shape_supported(bs.warps, bs.gw, bs.nb), "HMCB_SPMM_BLOCK_*: not a built mapping");
    const int stages = env_int("HMCB_SPMM_STAGES", 2);
    HMCB_CHECK(stages >= 2 && stages <= SPMM_MAX_STAGES, "HMCB_SPMM_STAGES out of range");
    // a stage = the B rows of a strip + the k-tiles of one (chunk, strip) group; stages start on
    // 1024-byte swizzle atoms
    const int cap16 = ((225 * 1024) / stages) / 1024 * 64;
    BlockHost host;
    if (build_block_tables(m, bs, cap16, env_int("HMCB_SPMM_COMPACT", 1) != 0, &host)) return -1;
    // it pays when a gathered B row feeds enough rows of its group: a k-tile (4 columns x 8 rows on the
    // tensor pipe) has to carry more nonzeros than the plain kernel turns over in the same time
    if (want_blocked > 0 || (host.ktiles > 0 && (double)host.nnz >= 6.0 * (double)host.ktiles))
      return upload_csr_blocks(e, host, (int)m.rows, bs, stages, cap16, strip);
  }
  const SpmmShape sh = shapes[which];
  const int S = 32 * sh.cpl, RB = sh.warps * sh.rw;
  // two stages of (96 KB of B rows + 14 KB of nonzero slots): the widest strips that fit measured
  // best once the B rows arrive by tensor-map TMA (profiles/spmm_lab_r01.json)
  const int kb = env_int("HMCB_SPMM_KB", 98304 / (S * 8));
  // a single column can hold RB nonzeros of the chunk: the slot limit must leave room for them
  const int emax = std::max(RB + sh.warps + (sh.warps + 3) / 4 + 1, env_int("HMCB_SPMM_EMAX", 896));
  const int stage_bytes = kb * S * 8 + emax * 16 + 128;
  // 227 KB of dynamic shared memory per block, minus the static barriers
  int stages = env_int("HMCB_SPMM_STAGES", std::min(3, (226 * 1024) / stage_bytes));
  HMCB_CHECK(kb >= 1 && stages >= 2 && stages <= SPMM_MAX_STAGES && stages * stage_bytes <= 226 * 1024,
             "SpMM strip configuration does not fit in shared memory");
  return upload_csr_strips(e, m, sh, kb, emax, stages, strip);
}

// Digits of a model matrix for the Ozaki-sliced products: the fewest digit planes whose row-wise
// representation error stays below 2^-45 of the row's absolute sum (at most `cap`), uploaded with the
// row exponents.
int oz_upload_matrix(hmcb_engine* e, const double* A, int64_t rows, int64_t cols, int64_t rows_pad, int64_t cols_pad,
                     int cap, int* S_out, signed char** slices_out, int** ea_out) {
  std::vector<int> ea((size_t)rows_pad);
  int S = std::min(5, cap);
  while (S < cap && oz_slice_rows_host(A, rows, cols, rows_pad, cols_pad, S, nullptr, ea.data()) > 0x1p-45) ++S;
  std::vector<signed char> sl((size_t)S * rows_pad * cols_pad);
  oz_slice_rows_host(A, rows, cols, rows_pad, cols_pad, S, sl.data(), ea.data());
  const signed char* p = nullptr; const int* q = nullptr;
  if (dev_upload(e, sl, &p) || dev_upload(e, ea, &q)) return -1;
  *S_out = S; *slices_out = const_cast<signed char*>(p); *ea_out = const_cast<int*>(q);
  return 0;
}

// Sets up the Ozaki-sliced tcgen05 path for the dense products (direct: G q and G^T r; premultiplied: GtG q)
// when it is valid and pays: int32 accumulation must not overflow (K * pairs * 128^2 < 2^31) and the
// products must be large.
int oz_setup(hmcb_engine* e) {
  const int want = env_int("HMCB_OZAKI", -1);   // -1: when it pays, 0: never, 1: whenever valid
  if (want == 0) return 0;
  const bool premult = e->lik == LK_DENSE_PREMULT;
  e->oz_orders = std::min(std::max(env_int("HMCB_OZAKI_ORDERS", OZ_NUM_ORDERS), 4), OZ_SLICES_MAX);
  e->oz_sb = e->oz_orders;
  e->oz_rows = premult ? e->dpad : e->npad;     // rows of the forward operator (padded)
  const int64_t Kmax = std::max<int64_t>(e->dpad, premult ? 0 : e->npad);
  if (Kmax * e->oz_orders * 128 * 128 >= (1ll << 31)) return 0;
  if (want < 0 && ((int64_t)e->dpad * e->oz_rows < (1ll << 20) || e->C < 1024)) return 0;
  if (premult && e->dpad == 128) return 0;      // the whole-proposal kernel with GtG in shared memory owns that size
  for (double v : e->h_A) if (!std::isfinite(v)) return 0;
  for (double v : e->h_At) if (!std::isfinite(v)) return 0;
  if (want < 0) {
    // The digits of a chain are cut 48 bits below the CHAIN's largest |value|.  With a fully populated
    // operator every output mixes all coordinates and that is far below its own rounding; an operator with
    // structural zeros (block structure) can have rows that only see coordinates many orders of magnitude
    // below the chain's largest one, which the native fp64 path keeps and this one would lose: such
    // operators stay on DMMA unless the path is forced.
    size_t zeros = 0;
    for (double v : e->h_A) zeros += v == 0.0;
    if (zeros * 100 > e->h_A.size()) return 0;
  }
  const int d = (int)e->d;
  if (oz_upload_matrix(e, e->h_A.data(), premult ? d : e->N, d, e->oz_rows, e->dpad, e->oz_orders, &e->oz_saG, &e->oz_AG,
                       &e->oz_eaG))
    return -1;
  if (!premult && oz_upload_matrix(e, e->h_At.data(), d, e->N, e->dpad, e->npad, e->oz_orders, &e->oz_saGt, &e->oz_AGt,
                                   &e->oz_eaGt))
    return -1;
  if (dev_alloc(e, (size_t)e->oz_sb * e->ld * Kmax, &e->oz_B) ||
      dev_alloc(e, (size_t)e->oz_orders * e->oz_rows * e->ld, &e->oz_C) ||
      dev_alloc(e, (size_t)e->ld, &e->oz_maxQ) || dev_alloc(e, (size_t)e->ld, &e->oz_maxR)) return -1;
  HMCB_CUDA(ozaki_init());
  HMCB_CUDA(ozaki_slice_map(e->oz_AG, e->dpad, e->oz_rows, e->oz_saG, 128, &e->oz_mapAG));
  HMCB_CUDA(ozaki_slice_map(e->oz_B, e->dpad, e->ld, e->oz_sb, 256, &e->oz_mapBq));
  HMCB_CUDA(ozaki_slice_map(e->oz_B, e->dpad, e->ld, e->oz_sb, 128, &e->oz_mapBqh));   // half tiles: CTA pairs
  HMCB_CUDA(ozaki_plane_map(e->oz_C, e->ld, e->oz_rows, e->oz_orders, (long long)e->oz_rows * e->ld, &e->oz_mapCq));
  // G^T r as 13 modular products + Chinese-remainder reconstruction instead of 21 digit products: opt-in
  // (HMCB_OZAKI_CRT=1).  Measured at config 3: 2.14 ms against 2.39 ms for the products, but no two modular
  // products share an operand tile (2 tile loads per product instead of 1.1: the L2 -> SM feed binds, tensor
  // pipe 48 % active) and reducing the chain batch modulo 13 moduli costs 0.73 ms against 0.21 ms of digit
  // slicing -- 3.1 ms against 2.8 ms for the whole G^T r side.
  e->oz_crt = !premult && env_int("HMCB_OZAKI_CRT", 0) > 0 && e->npad * 128ll * 128ll < (1ll << 31) && e->npad < 40000;
  if (e->oz_crt) {
    std::vector<signed char> res((size_t)OZ_NUM_MODULI * e->dpad * e->npad);
    std::vector<int> ea((size_t)e->dpad);
    oz_residue_rows_host(e->h_At.data(), d, e->N, e->dpad, e->npad, res.data(), ea.data());
    const signed char* pa = nullptr; const int* pe = nullptr;
    if (dev_upload(e, res, &pa) || dev_upload(e, ea, &pe)) return -1;
    e->oz_crtA = const_cast<signed char*>(pa); e->oz_crt_ea = const_cast<int*>(pe);
    if (dev_alloc(e, (size_t)OZ_NUM_MODULI * e->ld * e->npad, &e->oz_crtB) ||
        dev_alloc(e, (size_t)OZ_NUM_MODULI * e->dpad * e->ld, &e->oz_crtC)) return -1;
    HMCB_CUDA(ozaki_slice_map(e->oz_crtA, e->npad, e->dpad, OZ_NUM_MODULI, 128, &e->oz_mapCrtA));
    HMCB_CUDA(ozaki_slice_map(e->oz_crtB, e->npad, e->ld, OZ_NUM_MODULI, 256, &e->oz_mapCrtB));
    HMCB_CUDA(ozaki_slice_map(e->oz_crtB, e->npad, e->ld, OZ_NUM_MODULI, 128, &e->oz_mapCrtBh));
  }
  if (!premult) {
    HMCB_CUDA(ozaki_slice_map(e->oz_AGt, e->npad, e->dpad, e->oz_saGt, 128, &e->oz_mapAGt));
    HMCB_CUDA(ozaki_slice_map(e->oz_B, e->npad, e->ld, e->oz_sb, 256, &e->oz_mapBr));
    HMCB_CUDA(ozaki_slice_map(e->oz_B, e->npad, e->ld, e->oz_sb, 128, &e->oz_mapBrh));
    HMCB_CUDA(ozaki_plane_map(e->oz_C, e->ld, e->dpad, e->oz_orders, (long long)e->dpad * e->ld, &e->oz_mapCr));
  }
  e->oz = true;
  return 0;
}

// Y = G q (premultiplied: GtG q) as int8 slice products on the tcgen05 tensor cores (ozaki.cuh): per-chain scale, digits of the
// chain batch, the exact slice products in int32 order planes -> oz_C (recombined by the caller's epilogue)
int oz_forward_product(hmcb_engine* e, const double* q_in, cudaStream_t s) {
  HMCB_CUDA(launch_oz_colmax(q_in, e->dpad, e->ld, e->oz_maxQ, s));
  HMCB_CUDA(launch_oz_slice_chains(q_in, e->dpad, e->ld, e->oz_sb, e->oz_maxQ, e->oz_B, s));
  HMCB_CUDA(launch_i8_gemm_orders(e->oz_mapAG, e->oz_mapBq, e->oz_mapBqh, e->oz_mapCq, e->oz_rows, e->ld, e->dpad, e->oz_saG,
                                  e->oz_sb, e->oz_orders, e->ld, s));
  e->launches += 3;
  return 0;
}

StagedCommon staged_common(const hmcb_engine* e, const hmcb_block* b) {
  StagedCommon S{};
  S.T = e->T; S.C = (int)e->C; S.ld = e->ld; S.jtiles = e->jtiles;
  if (b) {
    S.chain_offset = b->chain_offset; S.seed = b->seed; S.stepsize = b->stepsize;
    S.randomize = b->randomize_stepsize;
    S.stepsize_chain = b->stepsize_chain;
  }
  return S;
}

inline int staged_lik_mode(const hmcb_engine* e) {
  switch (e->lik) {
    case LK_DENSE_PREMULT: case LK_CSR_PREMULT: return LIK_PREMULT;
    case LK_DENSE_DIRECT: case LK_CSR_DIRECT: return LIK_DIRECT;
    default: return LIK_NONE;
  }
}

// tensor map of a position plane handed to the strip SpMM (only the two ping-pong planes are)
inline const CUtensorMap& q_map(const hmcb_engine* e, const double* q) {
  return e->tmap_q[q == e->q_w[1] ? 1 : 0];
}

// total gradient at q_in fused with the update described by `epi` (q_in -> epi.q_out)
int staged_gradient_pass(hmcb_engine* e, const double* q_in, UpdateEpi epi, cudaStream_t s) {
  epi.q_in = q_in;
  KernelTimer timer(e, s, 0);
  switch (e->lik) {
    case LK_NONE: {
      HMCB_CUDA(launch_st_update(staged_common(e, nullptr), epi, s));
      e->launches += 1;
      break;
    }
    case LK_DENSE_PREMULT: {
      epi.sub = e->dvec;
      if (e->oz) {   // GtG q as int8 slice products on tcgen05, the update fused into the recombination
        if (oz_forward_product(e, q_in, s)) return -1;
        HMCB_CUDA(launch_oz_combine_update(e->oz_C, (long long)e->oz_rows * e->ld, e->dpad, e->ld, e->oz_orders, e->oz_eaG,
                                           e->oz_maxQ, epi, s));
        e->launches += 1;
        break;
      }
      HMCB_CUDA(launch_gemm_update(e->dA, e->dpad, e->dpad, q_in, e->ld, e->dpad, epi, s));
      e->launches += 1;
      break;
    }
    case LK_DENSE_DIRECT: {
      ResidualEpi r{(int)e->N, (int)e->C, e->ld, e->dvec, e->dvar, e->R};
      epi.sub = nullptr;
      if (e->oz) {
        // both products as int8 slice products on the tcgen05 tensor cores (ozaki.cuh), the fp64
        // recombination fused with the residual / update epilogue
        const long long plane_q = (long long)e->npad * e->ld, plane_r = (long long)e->dpad * e->ld;
        if (oz_forward_product(e, q_in, s)) return -1;
        HMCB_CUDA(launch_oz_combine_residual(e->oz_C, plane_q, e->npad, e->ld, e->oz_orders, e->oz_eaG, e->oz_maxQ, r,
                                             e->oz_maxR, s));
        if (e->oz_crt) {   // residues of R, 13 modular products, Chinese-remainder reconstruction + update
          HMCB_CUDA(launch_oz_residue_chains(e->R, (int)e->npad, e->ld, e->oz_maxR, e->oz_crtB, s));
          HMCB_CUDA(launch_i8_gemm_moduli(e->oz_mapCrtA, e->oz_mapCrtB, e->oz_mapCrtBh, e->dpad, e->ld, e->npad, e->ld,
                                          e->oz_crtC, s));
          HMCB_CUDA(launch_oz_crt_update(e->oz_crtC, (long long)e->dpad * e->ld, (int)e->dpad, e->ld, e->oz_crt_ea,
                                         e->oz_maxR, epi, s));
          e->launches += 4;
          break;
        }
        HMCB_CUDA(launch_oz_slice_chains(e->R, e->npad, e->ld, e->oz_sb, e->oz_maxR, e->oz_B, s));
        HMCB_CUDA(launch_i8_gemm_orders(e->oz_mapAGt, e->oz_mapBr, e->oz_mapBrh, e->oz_mapCr, e->dpad, e->ld, e->npad, e->oz_saGt,
                                        e->oz_sb, e->oz_orders, e->ld, s));
        HMCB_CUDA(launch_oz_combine_update(e->oz_C, plane_r, e->dpad, e->ld, e->oz_orders, e->oz_eaGt, e->oz_maxR, epi,
                                           s));
        e->launches += 4;
        break;
      }
      HMCB_CUDA(launch_gemm_residual(e->dA, e->dpad, e->npad, q_in, e->ld, e->dpad, r, s));
      HMCB_CUDA(launch_gemm_update(e->dAt, e->npad, e->dpad, e->R, e->ld, e->npad, epi, s));
      e->launches += 2;
      break;
    }
    case LK_CSR_DIRECT: {
      ResidualEpi r{(int)e->N, (int)e->C, e->ld, e->dvec, e->dvar, e->R};
      epi.sub = nullptr;
      if (e->use_strips) {
        HMCB_CUDA(launch_spmm_strip_residual(e->strip_dev, q_map(e, q_in), q_in, e->ld, r, s));
        HMCB_CUDA(launch_spmm_strip_update(e->strip_t_dev, e->tmap_R, e->R, e->ld, epi, s));
      } else {
        HMCB_CUDA(launch_spmm_residual(e->csr_dev, q_in, e->ld, r, s));
        HMCB_CUDA(launch_spmm_update(e->csr_t_dev, e->R, e->ld, epi, s));
      }
      e->launches += 2;
      break;
    }
    case LK_CSR_PREMULT: {
      epi.sub = e->dvec;
      if (e->use_strips) HMCB_CUDA(launch_spmm_strip_update(e->strip_dev, q_map(e, q_in), q_in, e->ld, epi, s));
      else HMCB_CUDA(launch_spmm_update(e->csr_dev, q_in, e->ld, epi, s));
      e->launches += 1;
      break;
    }
    default: return fail("internal: bad likelihood kind on the staged path");
  }
  return 0;
}

// per-chain partial sums of the likelihood misfit at q -> lpart
int staged_misfit_pass(hmcb_engine* e, const double* q, cudaStream_t s) {
  MisfitEpi m{};
  m.mode = staged_lik_mode(e); m.C = (int)e->C; m.ld = e->ld; m.q = q; m.part = e->lpart;
  KernelTimer timer(e, s, 1);
  switch (e->lik) {
    case LK_NONE: return 0;
    case LK_DENSE_PREMULT:
      m.rows = (int)e->d; m.vec = e->dvec;
      if (e->oz) {
        if (oz_forward_product(e, q, s)) return -1;
        HMCB_CUDA(launch_oz_combine_misfit(e->oz_C, (long long)e->oz_rows * e->ld, e->dpad, e->ld, e->oz_orders, e->oz_eaG,
                                           e->oz_maxQ, m, s));
        break;
      }
      HMCB_CUDA(launch_gemm_misfit(e->dA, e->dpad, e->dpad, q, e->ld, e->dpad, m, s));
      break;
    case LK_DENSE_DIRECT:
      m.rows = (int)e->N; m.vec = e->dAmis ? e->dvecmis : e->dvec; m.sigma = e->dsigma;
      if (e->oz && !e->dAmis) {   // G q on tcgen05, the misfit sums in the recombination (one partial per 128 rows)
        if (oz_forward_product(e, q, s)) return -1;
        HMCB_CUDA(launch_oz_combine_misfit(e->oz_C, (long long)e->npad * e->ld, e->npad, e->ld, e->oz_orders, e->oz_eaG,
                                           e->oz_maxQ, m, s));
        break;
      }
      HMCB_CUDA(launch_gemm_misfit(e->dAmis ? e->dAmis : e->dA, e->dpad, e->npad, q, e->ld, e->dpad, m, s));
      break;
    case LK_CSR_DIRECT:
      m.rows = (int)e->N; m.vec = e->dvec; m.sigma = e->dsigma;
      if (e->use_strips) HMCB_CUDA(launch_spmm_strip_misfit(e->strip_dev, q_map(e, q), q, e->ld, m, s));
      else HMCB_CUDA(launch_spmm_misfit(e->csr_dev, q, e->ld, m, s));
      break;
    case LK_CSR_PREMULT:
      m.rows = (int)e->d; m.vec = e->dvec;
      if (e->use_strips) HMCB_CUDA(launch_spmm_strip_misfit(e->strip_dev, q_map(e, q), q, e->ld, m, s));
      else HMCB_CUDA(launch_spmm_misfit(e->csr_dev, q, e->ld, m, s));
      break;
    default: return fail("internal: bad likelihood kind on the staged path");
  }
  e->launches += 1;
  return 0;
}

DecideArgs decide_args(const hmcb_engine* e) {
  DecideArgs D{};
  D.C = (int)e->C; D.ld = e->ld; D.jtiles = e->jtiles; D.ltiles = e->ltiles;
  D.lik_mode = staged_lik_mode(e);
  D.dtd = e->dtd; D.const_sum = e->T.const_sum;
  D.k0part = e->k0part; D.k1part = e->k1part; D.upart = e->upart; D.lpart = e->lpart;
  D.uacc = e->uacc; D.acc = e->accbuf;
  return D;
}

// dK/dp = M^-1 p for the whole batch (Full mass matrix): v_w = Minv . p_w
int full_mass_velocity(hmcb_engine* e, cudaStream_t s) {
  StoreEpi st{(int)e->d, (int)e->C, e->ld, e->v_w};
  HMCB_CUDA(launch_gemm_store(e->dMinv, e->dpad, e->dpad, e->p_w, e->ld, e->dpad, st, s));
  e->launches += 1;
  return 0;
}

// One trajectory with a Full mass matrix (MassMatrices.py:241-327): p = L z and dK/dp = M^-1 p are DMMA
// GEMMs over the chain batch; the momentum update stays fused into the likelihood GEMM / SpMM epilogue,
// the position update (which needs M^-1 p of the updated momentum over all coordinates) follows as a
// GEMM + an elementwise kernel.  Works in place on q_w[0]; q_w[1] is the scratch plane of the draws.
int full_mass_trajectory(hmcb_engine* e, const hmcb_block* b, const StagedCommon& SC, int64_t kglob, size_t kc,
                         bool grad_checks, int* cur_out, int* fcur_out, cudaStream_t s) {
  const int C = (int)e->C, d = (int)e->d, ld = e->ld;
  const size_t flag_bytes = (size_t)ld * sizeof(unsigned);
  const int G = e->S.grads_per_proposal;
  const bool reflects = e->T.refl_lb != nullptr || e->T.refl_ub != nullptr;
  int fcur = *fcur_out;
  double* q = e->q_w[0];
  StagedCommon SD = SC;
  SD.draw_only = 1;
  HMCB_CUDA(launch_st_begin(SD, kglob, 0.0, e->q_cur, q, e->q_w[1], b->z_in ? b->z_in + kc * d : nullptr,
                            b->u_step_in ? b->u_step_in + kc : nullptr,
                            b->u_accept_in ? b->u_accept_in + kc : nullptr, e->eps, e->uacc, e->k0part, nullptr,
                            b->out_stepsize ? b->out_stepsize + kc : nullptr, s));
  StoreEpi to_p{d, C, ld, e->p_w};
  HMCB_CUDA(launch_gemm_store(e->dL, e->dpad, e->dpad, e->q_w[1], ld, e->dpad, to_p, s));   // p = L z
  e->launches += 2;
  if (full_mass_velocity(e, s)) return -1;
  StagedCommon SV = SC;
  SV.v = e->v_w;
  bool v_valid = true;
  // first (lone) position update + the kinetic energy of the drawn momentum
  HMCB_CUDA(launch_st_kpos(SV, e->ops[0].a, e->q_cur, q, e->p_w, e->eps, e->k0part,
                           grad_checks ? e->flags[fcur] : nullptr, s));
  e->launches += 1;
  if (reflects) v_valid = false;
  int gi = 0;
  for (size_t o = 1; o < e->ops.size(); ++o) {
    const StageOp& op = e->ops[o];
    if (op.has_b) {
      UpdateEpi epi{};
      epi.T = e->T; epi.C = C; epi.ld = ld;
      epi.q_out = q; epi.p = e->p_w; epi.eps = e->eps;
      epi.b_mult = op.b; epi.a_mult = 0.0; epi.momentum_only = 1;
      if (grad_checks) epi.flags_in = e->flags[fcur];
      if (b->trace_q) {
        const size_t off = (((size_t)(kglob - b->proposal_offset) * G + gi) * C) * d;
        epi.trace_q = b->trace_q + off;
        epi.trace_g = b->trace_g + off;
      }
      if (staged_gradient_pass(e, q, epi, s)) return -1;
      ++gi;
      v_valid = false;
    }
    if (!v_valid && full_mass_velocity(e, s)) return -1;
    v_valid = true;
    if (grad_checks) {
      fcur ^= 1;
      HMCB_CUDA(cudaMemsetAsync(e->flags[fcur], 0, flag_bytes, s));
    }
    HMCB_CUDA(launch_st_kpos(SV, op.a, q, q, e->p_w, e->eps, nullptr, grad_checks ? e->flags[fcur] : nullptr, s));
    e->launches += 1;
    if (reflects) v_valid = false;
  }
  if (!v_valid && full_mass_velocity(e, s)) return -1;   // the energies need M^-1 p of the final momentum
  *cur_out = 0;
  *fcur_out = fcur;
  return 0;
}

int staged_run_block(hmcb_engine* e, const hmcb_block* b, cudaStream_t s) {
  const int C = (int)e->C, d = (int)e->d, ld = e->ld;
  const bool grad_checks = e->T.grad_check_mask != 0u;
  const bool any_checks = e->T.n_checks > 0;
  const size_t flag_bytes = (size_t)ld * sizeof(unsigned);
  const int G = e->S.grads_per_proposal;
  StagedCommon SC = staged_common(e, b);

  // chain-major API tensor -> transposed working layout
  HMCB_CUDA(launch_st_transpose(b->q, C, d, d, e->q_cur, ld, s));
  e->launches += 1;

  const int64_t first_row = (b->proposal_offset + b->thinning - 1) / b->thinning;
  for (int64_t kb = 0; kb < b->proposals; ++kb) {
    const int64_t kglob = b->proposal_offset + kb;
    const size_t kc = (size_t)kb * C;
    int cur = 0, fcur = 0;
    if (grad_checks) HMCB_CUDA(cudaMemsetAsync(e->flags[fcur], 0, flag_bytes, s));
    if (e->mass_full) {
      if (full_mass_trajectory(e, b, SC, kglob, kc, grad_checks, &cur, &fcur, s)) return -1;
    } else {
    HMCB_CUDA(launch_st_begin(SC, kglob, e->ops[0].a, e->q_cur, e->q_w[cur], e->p_w,
                              b->z_in ? b->z_in + kc * d : nullptr,
                              b->u_step_in ? b->u_step_in + kc : nullptr,
                              b->u_accept_in ? b->u_accept_in + kc : nullptr, e->eps, e->uacc,
                              e->k0part, grad_checks ? e->flags[fcur] : nullptr,
                              b->out_stepsize ? b->out_stepsize + kc : nullptr, s));
    e->launches += 1;
    int gi = 0;
    for (size_t o = 1; o < e->ops.size(); ++o) {
      const StageOp& op = e->ops[o];
      if (op.has_b) {
        UpdateEpi epi{};
        epi.T = e->T; epi.C = C; epi.ld = ld;
        epi.q_out = e->q_w[cur ^ 1]; epi.p = e->p_w; epi.eps = e->eps;
        epi.b_mult = op.b; epi.a_mult = op.a;
        if (grad_checks) {
          HMCB_CUDA(cudaMemsetAsync(e->flags[fcur ^ 1], 0, flag_bytes, s));
          epi.flags_in = e->flags[fcur];
          epi.flags_out = e->flags[fcur ^ 1];
        }
        if (b->trace_q) {
          const size_t off = (((size_t)kb * G + gi) * C) * d;
          epi.trace_q = b->trace_q + off;
          epi.trace_g = b->trace_g + off;
        }
        if (staged_gradient_pass(e, e->q_w[cur], epi, s)) return -1;
        cur ^= 1; fcur ^= 1; ++gi;
      } else {
        if (grad_checks) HMCB_CUDA(cudaMemsetAsync(e->flags[fcur], 0, flag_bytes, s));
        HMCB_CUDA(launch_st_position(SC, op.a, e->q_w[cur], e->p_w, e->eps,
                                     grad_checks ? e->flags[fcur] : nullptr, s));
        e->launches += 1;
      }
    }
    }   // !mass_full
    // energies, decision, state update
    if (any_checks) HMCB_CUDA(cudaMemsetAsync(e->flags[2], 0, flag_bytes, s));
    StagedCommon SE = SC;
    if (e->mass_full) SE.v = e->v_w;      // K = 0.5 p . M^-1 p (full_mass_trajectory left v current)
    HMCB_CUDA(launch_st_energy(SE, e->q_w[cur], e->p_w, e->k1part, e->upart,
                               any_checks ? e->flags[2] : nullptr, s));
    e->launches += 1;
    if (staged_misfit_pass(e, e->q_w[cur], s)) return -1;
    DecideArgs D = decide_args(e);
    D.flags = any_checks ? e->flags[2] : nullptr;
    D.x = b->x;
    D.out_accept = b->out_accept ? b->out_accept + kc : nullptr;
    D.out_h0 = b->out_h0 ? b->out_h0 + kc : nullptr;
    D.out_h1 = b->out_h1 ? b->out_h1 + kc : nullptr;
    D.accepted_total = b->accepted_total;
    D.stepsize_chain = b->stepsize_chain;
    D.tune = AutotuneArgs{b->autotune ? 1 : 0, b->target_acceptance_rate, b->learning_rate};
    D.kglob = kglob;
    double* sample_rows = nullptr;
    if (b->out_samples && (kglob % b->thinning) == 0) {
      sample_rows = b->out_samples + (size_t)(kglob / b->thinning - first_row) * C * (size_t)(d + 1);
      D.sample_misfit = sample_rows + d;
      D.sample_stride = d + 1;
    }
    HMCB_CUDA(launch_st_decide(D, s));
    HMCB_CUDA(launch_st_commit(C, d, ld, e->accbuf, e->q_w[cur], e->p_w, e->q_cur, sample_rows,
                               b->out_q_prop ? b->out_q_prop + kc * d : nullptr,
                               b->out_p_prop ? b->out_p_prop + kc * d : nullptr, s));
    e->launches += 2;
  }
  HMCB_CUDA(launch_st_transpose(e->q_cur, d, C, ld, b->q, d, s));
  e->launches += 1;
  return 0;
}

FusedArgs fused_args(const hmcb_engine* e, const hmcb_block* b) {
  FusedArgs A{};
  A.T = e->T; A.S = e->S; A.chains = (int)e->C; A.proposals = (int)b->proposals;
  A.thinning = b->thinning; A.proposal_offset = b->proposal_offset; A.chain_offset = b->chain_offset;
  A.seed = b->seed; A.stepsize = b->stepsize; A.randomize = b->randomize_stepsize;
  A.q = b->q; A.x = b->x; A.z_in = b->z_in; A.u_step_in = b->u_step_in; A.u_accept_in = b->u_accept_in;
  A.out_samples = b->out_samples; A.out_accept = b->out_accept; A.out_h0 = b->out_h0; A.out_h1 = b->out_h1;
  A.accepted_total = b->accepted_total; A.out_q_prop = b->out_q_prop; A.out_p_prop = b->out_p_prop;
  A.trace_q = b->trace_q; A.trace_g = b->trace_g;
  A.stepsize_chain = b->stepsize_chain; A.out_stepsize = b->out_stepsize;
  A.tune = AutotuneArgs{b->autotune ? 1 : 0, b->target_acceptance_rate, b->learning_rate};
  A.exact = e->exact ? 1 : 0;
  return A;
}

}  // namespace

// ========================================================================== C ABI =====

extern "C" {

int hmcb_abi_version(void) { return HMCB_ABI_VERSION; }

// Host-side self check of the strip tables: builds them exactly as hmcb_finalize does and walks
// them the way csr_spmm_strip_kernel does (strip by strip, warp stream by warp stream, up to the
// sentinel), checking the format invariants on the way.  No GPU involved.
int hmcb_debug_spmm_tables(int64_t rows, int64_t cols, int64_t nnz, const int32_t* indptr,
                           const int32_t* indices, const double* data, int warps, int rw, int cpl, int kb,
                           int emax, int allow_compact, int64_t chains, const double* B, double* Y,
                           int64_t* info) {
  HMCB_CHECK(indptr && B && Y && info && (nnz == 0 || (indices && data)), "hmcb_debug_spmm_tables: NULL argument");
  HMCB_CHECK(rows > 0 && cols > 0 && nnz >= 0 && chains > 0 && warps > 0 && rw > 0 && (cpl == 1 || cpl == 2) &&
                 kb > 0 && emax > 0, "hmcb_debug_spmm_tables: bad sizes");
  HostCsr m;
  m.rows = rows; m.cols = cols; m.nnz = nnz;
  m.indptr.assign(indptr, indptr + rows + 1);
  m.indices.assign(indices, indices + nnz);
  m.data.assign(data, data + nnz);
  if (check_csr(m, "matrix")) return -1;
  const SpmmShape shape{warps, rw, cpl};
  const int S = 32 * cpl, RB = warps * rw;
  emax = std::max(emax, RB + warps + (warps + 3) / 4 + 1);   // the floor hmcb_finalize applies
  StripHost host;
  if (build_strip_tables(m, shape, kb, emax, allow_compact != 0, &host)) return -1;
  const int esz = host.compact ? 8 : 16;
  const int64_t T = host.cstride;
  std::fill(Y, Y + rows * chains, 0.0);
  HMCB_CHECK((int64_t)host.strip_ptr.size() == host.chunks + 1 && host.kb_box <= kb, "strip tables: bad sizes");
  for (int b = 0; b < host.chunks; ++b) {
    for (int32_t s = host.strip_ptr[b]; s < host.strip_ptr[b + 1]; ++s) {
      const SpmmStrip& st = host.strips[s];
      HMCB_CHECK(st.ent_cnt > 0 && st.ent_cnt <= emax && st.col0 >= 0 && st.col0 < T && st.ncols <= host.kb_box &&
                     (size_t)(st.ent_off + st.ent_cnt) * 16 <= host.ent.size(), "strip tables: bad strip descriptor");
      const unsigned char* group = host.ent.data() + (size_t)st.ent_off * 16;
      const int32_t* first = reinterpret_cast<const int32_t*>(group);
      const int slots = st.ent_cnt * 16 / esz;
      for (int w = 0; w < warps; ++w) {
        int slot = first[w], last_row = 0;
        for (;;) {
          HMCB_CHECK(slot >= 0 && slot < slots, "strip tables: stream runs out of its group");
          double val; int64_t off; int row;
          if (host.compact) {
            const SpmmEntry32* en = reinterpret_cast<const SpmmEntry32*>(group) + slot;
            val = (double)en->val; off = en->meta & 0xFFFFFFu; row = (int)(en->meta >> 24);
          } else {
            const SpmmEntry* en = reinterpret_cast<const SpmmEntry*>(group) + slot;
            val = en->val; off = en->off; row = en->row;
          }
          ++slot;
          if (row == rw) break;   // sentinel
          HMCB_CHECK(row >= last_row && row < rw && off % (S * 8) == 0 && off / (S * 8) < st.ncols,
                     "strip tables: bad nonzero");
          last_row = row;
          const int64_t i = (int64_t)b * RB + (int64_t)w * rw + row, j = st.col0 + off / (S * 8) * T;
          HMCB_CHECK(i < rows && j < cols, "strip tables: nonzero outside the matrix");
          for (int64_t c = 0; c < chains; ++c) Y[i * chains + c] += val * B[j * chains + c];
        }
      }
    }
  }
  info[0] = T; info[1] = (int64_t)host.strips.size(); info[2] = host.compact; info[3] = (int64_t)host.ent.size();
  return 0;
}
// Same self check for the row-blocked tensor-core tables: clusters the rows, builds the tables as
// hmcb_finalize does and walks them the way csr_spmm_block_kernel does (strip by strip, (warp, group)
// by (warp, group), k-tile by k-tile, undoing the swizzled offsets).  No GPU involved.
int hmcb_debug_spmm_block_tables(int64_t rows, int64_t cols, int64_t nnz, const int32_t* indptr,
                                 const int32_t* indices, const double* data, int warps, int groups_per_warp,
                                 int chain_boxes, int cap16, int allow_compact, int64_t chains, const double* B,
                                 double* Y, int64_t* info) {
  HMCB_CHECK(indptr && B && Y && info && (nnz == 0 || (indices && data)), "hmcb_debug_spmm_block_tables: NULL argument");
  HMCB_CHECK(rows > 0 && cols > 0 && nnz >= 0 && chains > 0 && warps > 0 && groups_per_warp > 0 && chain_boxes > 0 &&
                 cap16 > 0, "hmcb_debug_spmm_block_tables: bad sizes");
  constexpr int R = SPMM_BLOCK_R;
  HostCsr m;
  m.rows = rows; m.cols = cols; m.nnz = nnz;
  m.indptr.assign(indptr, indptr + rows + 1);
  m.indices.assign(indices, indices + nnz);
  m.data.assign(data, data + nnz);
  if (check_csr(m, "matrix")) return -1;
  const BlockShape shape{warps, groups_per_warp, chain_boxes};
  const int GPC = warps * groups_per_warp;
  BlockHost host;
  if (build_block_tables(m, shape, cap16, allow_compact != 0, &host)) return -1;
  const int64_t T = host.cstride;
  const int hdr16 = ((GPC + 1) * 4 + 15) / 16;
  const int rec_hdr = 32, vsz = host.compact ? 4 : 8;
  const int64_t region = (int64_t)chain_boxes * host.box_rows * 128;
  HMCB_CHECK(host.box_rows % 8 == 0 && host.box_rows <= 256 && host.box_rows * host.row_boxes >= host.kb_box &&
                 (int64_t)host.box_rows * host.row_boxes * chain_boxes * 8 + host.emax <= cap16,
             "block tables: stage overflows");
  std::fill(Y, Y + rows * chains, 0.0);
  HMCB_CHECK((int64_t)host.strip_ptr.size() == host.chunks + 1 &&
                 (int64_t)host.perm.size() == (int64_t)host.chunks * GPC * R, "block tables: bad sizes");
  {   // the permutation holds every row exactly once
    std::vector<char> seen(rows, 0);
    for (int32_t i : host.perm) {
      if (i < 0) continue;
      HMCB_CHECK(i < rows && !seen[i], "block tables: bad permutation");
      seen[i] = 1;
    }
    for (int64_t i = 0; i < rows; ++i) HMCB_CHECK(seen[i], "block tables: a row is missing from the permutation");
  }
  int64_t ktiles = 0, full_tiles = 0;
  for (int b = 0; b < host.chunks; ++b) {
    for (int32_t s = host.strip_ptr[b]; s < host.strip_ptr[b + 1]; ++s) {
      const SpmmStrip& st = host.strips[s];
      HMCB_CHECK(st.ent_cnt > 0 && st.ent_cnt <= host.emax && st.col0 >= 0 && st.col0 < T && st.ncols <= host.kb_box &&
                     (size_t)(st.ent_off + st.ent_cnt) * 16 <= host.ent.size(), "block tables: bad strip descriptor");
      const unsigned char* group = host.ent.data() + (size_t)st.ent_off * 16;
      const uint32_t* hdr = reinterpret_cast<const uint32_t*>(group);
      const int64_t total = hdr[GPC];
      const unsigned char* stream = group + (size_t)hdr16 * 16;
      const int64_t stream_bytes = (int64_t)(st.ent_cnt - hdr16) * 16;
      int64_t sum = 0, pos = 0;
      for (int wg = 0; wg < GPC; ++wg) {
        const int64_t first = hdr[wg] >> 10, n = hdr[wg] & 1023u;
        HMCB_CHECK(first * 4 == pos, "block tables: (warp, group) streams are not consecutive");
        sum += n;
        for (int64_t k = 0; k < n; ++k) {
          HMCB_CHECK(pos + rec_hdr <= stream_bytes, "block tables: record runs out of its group");
          const uint32_t* w = reinterpret_cast<const uint32_t*>(stream + pos);
          const uint32_t values = w[1];
          const unsigned char* vals = stream + pos + rec_hdr;
          HMCB_CHECK(pos + rec_hdr + (int64_t)values * vsz <= stream_bytes, "block tables: values run out of their group");
          uint32_t before = 0, classes = 0;
          int used = 0;
          for (int c = 0; c < 4; ++c) {
            const uint32_t word = w[2 * c];
            const int64_t o = (int64_t)(word & 0x1FFFu) << 4;
            const uint32_t colmask = (word >> 13) & 0xFFu;
            HMCB_CHECK(w[2 * c + 1] == values && (word >> 21) == before, "block tables: bad record header");
            if (!colmask) continue;
            ++used;
            const int64_t rb = o / region, rem = o % region, rl = rem / 128;
            HMCB_CHECK(rb < host.row_boxes && rl < host.box_rows && rem % 128 == (rl & 7) * 16,
                       "block tables: bad swizzled B row offset");
            classes |= 1u << ((rl >> 1) & 3);
            const int64_t j = st.col0 + (rb * host.box_rows + rl) * T;
            HMCB_CHECK(j < cols, "block tables: column outside the matrix");
            for (int r = 0; r < R; ++r) {
              if (!(colmask >> r & 1u)) continue;
              const double v = host.compact ? (double)reinterpret_cast<const float*>(vals)[before]
                                            : reinterpret_cast<const double*>(vals)[before];
              ++before;
              const int32_t i = host.perm[((size_t)b * GPC + wg) * R + r];
              HMCB_CHECK(v != 0.0 && i >= 0, "block tables: value on a padding row");
              for (int64_t ch = 0; ch < chains; ++ch) Y[i * chains + ch] = std::fma(v, B[j * chains + ch], Y[i * chains + ch]);
            }
          }
          HMCB_CHECK(before == values && used > 0, "block tables: value count mismatch");
          if (used == 4 && classes == 15u) ++full_tiles;    // 4 columns in 4 bank classes: conflict free
          pos += rec_hdr + ((int64_t)values * vsz + 7) / 8 * 8;
        }
      }
      HMCB_CHECK(sum == total && total > 0, "block tables: k-tile count mismatch");
      ktiles += total;
    }
  }
  const int64_t balanced = full_tiles;
  HMCB_CHECK(ktiles == host.ktiles, "block tables: total k-tile count mismatch");
  info[0] = T; info[1] = (int64_t)host.strips.size(); info[2] = host.blocks; info[3] = (int64_t)host.ent.size();
  info[4] = host.nnz; info[5] = host.ktiles; info[6] = balanced; info[7] = host.compact;
  return 0;
}
const char* hmcb_last_error(void) { return g_error.c_str(); }

int hmcb_create(int device, int64_t chains, int64_t dims, hmcb_engine** out) {
  HMCB_CHECK(out != nullptr, "hmcb_create: out is NULL");
  *out = nullptr;
  HMCB_CHECK(chains > 0 && chains < (1ll << 30), "hmcb_create: chains must be in [1, 2^30)");
  HMCB_CHECK(dims > 0 && dims < (1ll << 24), "hmcb_create: dims must be in [1, 2^24)");
  int count = 0;
  HMCB_CUDA(cudaGetDeviceCount(&count));
  HMCB_CHECK(device >= 0 && device < count, "hmcb_create: no such CUDA device");
  cudaDeviceProp prop;
  HMCB_CUDA(cudaGetDeviceProperties(&prop, device));
  HMCB_CHECK(prop.major == 10,
             std::string("hmcb_create: kernels are built for sm_100a (B200) only; device is ") + prop.name);
  HMCB_CUDA(cudaSetDevice(device));
  hmcb_engine* e = new hmcb_engine();
  e->device = device; e->C = chains; e->d = dims;
  e->exact = env_int("HMCB_EXACT", 0) != 0;
  *out = e;
  return 0;
}

int hmcb_destroy(hmcb_engine* e) {
  if (!e) return 0;
  cudaSetDevice(e->device);
  free_device(e);
  if (e->s_compute) cudaStreamDestroy(e->s_compute);
  if (e->s_copy) cudaStreamDestroy(e->s_copy);
  cudaFree(e->sh_q); cudaFree(e->sh_x); cudaFree(e->sh_acc); cudaFree(e->sh_buf[0]); cudaFree(e->sh_buf[1]);
  for (int i = 0; i < 2; ++i) {
    if (e->ev_produced[i]) cudaEventDestroy(e->ev_produced[i]);
    if (e->ev_drained[i]) cudaEventDestroy(e->ev_drained[i]);
    for (cudaEvent_t ev : e->kev[i]) cudaEventDestroy(ev);
  }
  delete e;
  return 0;
}

int hmcb_set_integrator(hmcb_engine* e, int integrator, int amount_of_steps) {
  HMCB_CHECK(e, "engine is NULL");
  HMCB_CHECK(integrator >= HMCB_INTEGRATOR_LF && integrator <= HMCB_INTEGRATOR_4S,
             "Unknown integrator used. Choices are: lf, 3s, 4s");
  HMCB_CHECK(amount_of_steps > 0, "amount_of_steps must be a positive integer");
  e->integrator = integrator; e->steps = amount_of_steps;
  build_schedule(e);  // cheap; lets the integrator change after finalize
  return 0;
}

int hmcb_set_exact_arithmetic(hmcb_engine* e, int on) {
  HMCB_CHECK(e, "engine is NULL");
  e->exact = on != 0;
  return 0;
}

int hmcb_set_mass_unit(hmcb_engine* e) {
  HMCB_CHECK(e, "engine is NULL");
  HMCB_CHECK(!e->finalized, "mass matrix must be set before hmcb_finalize");
  e->mass_diag = false; e->mass_full = false;
  return 0;
}

int hmcb_set_mass_diagonal(hmcb_engine* e, const double* diagonal, const double* inverse_diagonal) {
  HMCB_CHECK(e && diagonal && inverse_diagonal, "hmcb_set_mass_diagonal: NULL argument");
  HMCB_CHECK(!e->finalized, "mass matrix must be set before hmcb_finalize");
  for (int64_t j = 0; j < e->d; ++j)
    HMCB_CHECK(diagonal[j] > 0.0, "hmcb_set_mass_diagonal: diagonal entries must be positive");
  copy_vec(e->h_diag, diagonal, e->d);
  copy_vec(e->h_invdiag, inverse_diagonal, e->d);
  e->mass_diag = true; e->mass_full = false;
  return 0;
}

int hmcb_set_mass_full(hmcb_engine* e, const double* cholesky_lower, const double* inverse) {
  HMCB_CHECK(e && cholesky_lower && inverse, "hmcb_set_mass_full: NULL argument");
  HMCB_CHECK(!e->finalized, "mass matrix must be set before hmcb_finalize");
  for (int64_t j = 0; j < e->d; ++j)
    HMCB_CHECK(cholesky_lower[(size_t)j * e->d + j] > 0.0, "hmcb_set_mass_full: the Cholesky factor needs a positive diagonal");
  copy_vec(e->h_L, cholesky_lower, e->d * e->d);
  copy_vec(e->h_Minv, inverse, e->d * e->d);
  for (int64_t i = 0; i < e->d; ++i)          // the factor is lower triangular: make that exact
    for (int64_t j = i + 1; j < e->d; ++j) e->h_L[(size_t)i * e->d + j] = 0.0;
  e->mass_full = true; e->mass_diag = false;
  return 0;
}

int hmcb_clear_target(hmcb_engine* e) {
  HMCB_CHECK(e, "engine is NULL");
  cudaSetDevice(e->device);
  free_device(e);
  e->priors.clear(); e->checks.clear();
  e->has_rlb = e->has_rub = false;
  e->lik = LK_NONE;
  e->h_A.clear(); e->h_At.clear(); e->h_vec.clear(); e->h_var.clear(); e->h_sigma.clear();
  e->h_Amis.clear(); e->h_vecmis.clear();
  e->csr = HostCsr(); e->csr_t = HostCsr();
  return 0;
}

int hmcb_add_prior(hmcb_engine* e, int kind, int64_t offset, int64_t len, const double* a,
                   const double* b, double constant) {
  HMCB_CHECK(e && a && b, "hmcb_add_prior: NULL argument");
  HMCB_CHECK(!e->finalized, "target must be described before hmcb_finalize");
  HMCB_CHECK(kind == HMCB_PRIOR_NORMAL || kind == HMCB_PRIOR_LAPLACE, "hmcb_add_prior: unknown kind");
  HMCB_CHECK(offset >= 0 && len > 0 && offset + len <= e->d, "hmcb_add_prior: range outside [0, dims)");
  HostPrior p;
  p.kind = kind; p.offset = offset; p.len = len; p.constant = constant;
  copy_vec(p.a, a, len); copy_vec(p.b, b, len);
  e->priors.push_back(std::move(p));
  return 0;
}

int hmcb_add_bound_check(hmcb_engine* e, int64_t offset, int64_t len, const double* lb,
                         const double* ub, int in_gradient) {
  HMCB_CHECK(e, "engine is NULL");
  HMCB_CHECK(!e->finalized, "target must be described before hmcb_finalize");
  HMCB_CHECK(offset >= 0 && len > 0 && offset + len <= e->d, "hmcb_add_bound_check: range outside [0, dims)");
  if (!lb && !ub) return 0;
  HMCB_CHECK((int)e->checks.size() < HMCB_MAX_CHECKS, "hmcb_add_bound_check: too many bound checks (max 32 distinct bounded objects per posterior)");
  HostCheck c;
  c.offset = offset; c.len = len; c.has_lb = lb != nullptr; c.has_ub = ub != nullptr;
  c.in_gradient = in_gradient ? 1 : 0;
  if (lb) copy_vec(c.lb, lb, len);
  if (ub) copy_vec(c.ub, ub, len);
  e->checks.push_back(std::move(c));
  return 0;
}

int hmcb_set_reflection(hmcb_engine* e, const double* lb, const double* ub) {
  HMCB_CHECK(e, "engine is NULL");
  HMCB_CHECK(!e->finalized, "target must be described before hmcb_finalize");
  e->has_rlb = lb != nullptr; e->has_rub = ub != nullptr;
  if (lb) copy_vec(e->rlb, lb, e->d);
  if (ub) copy_vec(e->rub, ub, e->d);
  return 0;
}

int hmcb_set_likelihood_dense_premult(hmcb_engine* e, const double* GtG, const double* Gtd0, double dtd) {
  HMCB_CHECK(e && GtG && Gtd0, "hmcb_set_likelihood_dense_premult: NULL argument");
  HMCB_CHECK(!e->finalized && e->lik == LK_NONE, "one likelihood per engine, set before hmcb_finalize");
  copy_vec(e->h_A, GtG, e->d * e->d);
  copy_vec(e->h_vec, Gtd0, e->d);
  e->dtd = dtd;
  e->lik = LK_DENSE_PREMULT;
  return 0;
}

int hmcb_set_likelihood_dense_direct(hmcb_engine* e, int64_t N, const double* G, const double* Gt,
                                     const double* d, const double* var, const double* sigma) {
  HMCB_CHECK(e && G && d && var && sigma, "hmcb_set_likelihood_dense_direct: NULL argument");
  HMCB_CHECK(!e->finalized && e->lik == LK_NONE, "one likelihood per engine, set before hmcb_finalize");
  HMCB_CHECK(N > 0 && N < (1ll << 24), "hmcb_set_likelihood_dense_direct: bad N");
  e->N = N;
  copy_vec(e->h_A, G, N * e->d);
  e->h_At.resize((size_t)N * e->d);
  if (Gt) {
    std::memcpy(e->h_At.data(), Gt, sizeof(double) * (size_t)N * e->d);
  } else {
    for (int64_t i = 0; i < N; ++i)
      for (int64_t j = 0; j < e->d; ++j) e->h_At[(size_t)j * N + i] = G[(size_t)i * e->d + j];
  }
  copy_vec(e->h_vec, d, N); copy_vec(e->h_var, var, N); copy_vec(e->h_sigma, sigma, N);
  e->lik = LK_DENSE_DIRECT;
  return 0;
}

int hmcb_set_likelihood_dense_direct_cov(hmcb_engine* e, int64_t N, const double* G, const double* GtCinv,
                                         const double* d, const double* UG, const double* Ud) {
  HMCB_CHECK(e && G && GtCinv && d && UG && Ud, "hmcb_set_likelihood_dense_direct_cov: NULL argument");
  const std::vector<double> ones((size_t)std::max<int64_t>(N, 1), 1.0);
  if (hmcb_set_likelihood_dense_direct(e, N, G, GtCinv, d, ones.data(), ones.data())) return -1;
  copy_vec(e->h_Amis, UG, N * e->d);
  copy_vec(e->h_vecmis, Ud, N);
  return 0;
}

static int fill_csr(HostCsr& m, int64_t rows, int64_t cols, int64_t nnz, const int32_t* indptr,
                    const int32_t* indices, const double* data, const char* what) {
  m.rows = rows; m.cols = cols; m.nnz = nnz;
  m.indptr.assign(indptr, indptr + rows + 1);
  m.indices.assign(indices, indices + nnz);
  m.data.assign(data, data + nnz);
  return check_csr(m, what);
}

int hmcb_set_likelihood_csr_direct(hmcb_engine* e, int64_t N, int64_t nnz, const int32_t* indptr,
                                   const int32_t* indices, const double* data, const int32_t* t_indptr,
                                   const int32_t* t_indices, const double* t_data, const double* d,
                                   const double* var, const double* sigma) {
  HMCB_CHECK(e && indptr && t_indptr && d && var && sigma, "hmcb_set_likelihood_csr_direct: NULL argument");
  HMCB_CHECK(nnz == 0 || (indices && data && t_indices && t_data), "hmcb_set_likelihood_csr_direct: NULL argument");
  HMCB_CHECK(!e->finalized && e->lik == LK_NONE, "one likelihood per engine, set before hmcb_finalize");
  HMCB_CHECK(N > 0 && N < (1ll << 30) && nnz >= 0 && nnz < (1ll << 31), "hmcb_set_likelihood_csr_direct: bad sizes");
  if (fill_csr(e->csr, N, e->d, nnz, indptr, indices, data, "G")) return -1;
  if (fill_csr(e->csr_t, e->d, N, nnz, t_indptr, t_indices, t_data, "G^T")) return -1;
  e->N = N;
  copy_vec(e->h_vec, d, N); copy_vec(e->h_var, var, N); copy_vec(e->h_sigma, sigma, N);
  e->lik = LK_CSR_DIRECT;
  return 0;
}

int hmcb_set_likelihood_csr_premult(hmcb_engine* e, int64_t nnz, const int32_t* indptr,
                                    const int32_t* indices, const double* data, const double* Gtd0,
                                    double dtd) {
  HMCB_CHECK(e && indptr && Gtd0, "hmcb_set_likelihood_csr_premult: NULL argument");
  HMCB_CHECK(nnz == 0 || (indices && data), "hmcb_set_likelihood_csr_premult: NULL argument");
  HMCB_CHECK(!e->finalized && e->lik == LK_NONE, "one likelihood per engine, set before hmcb_finalize");
  HMCB_CHECK(nnz >= 0 && nnz < (1ll << 31), "hmcb_set_likelihood_csr_premult: bad nnz");
  if (fill_csr(e->csr, e->d, e->d, nnz, indptr, indices, data, "GtG")) return -1;
  copy_vec(e->h_vec, Gtd0, e->d);
  e->dtd = dtd;
  e->lik = LK_CSR_PREMULT;
  return 0;
}

int hmcb_set_likelihood_srcloc3d(hmcb_engine* e, int64_t events, int64_t stations, const double* rx,
                                 const double* ry, const double* rz, const double* tobs,
                                 const double* std, int infer_velocity, double velocity) {
  HMCB_CHECK(e && rx && ry && rz && tobs && std, "hmcb_set_likelihood_srcloc3d: NULL argument");
  HMCB_CHECK(!e->finalized && e->lik == LK_NONE, "one likelihood per engine, set before hmcb_finalize");
  HMCB_CHECK(events > 0 && stations > 0, "hmcb_set_likelihood_srcloc3d: bad sizes");
  HMCB_CHECK(e->d == 4 * events + (infer_velocity ? 1 : 0),
             "hmcb_set_likelihood_srcloc3d: dims must be 4*events (+1 when the velocity is inferred)");
  HMCB_CHECK(srcloc_supported((int)events, (int)stations),
             "hmcb_set_likelihood_srcloc3d: events x stations exceeds the kernel's shared-memory staging");
  e->events = events; e->stations = stations; e->infer_velocity = infer_velocity ? 1 : 0;
  e->velocity = velocity;
  e->srcloc_np = 4;
  copy_vec(e->h_rx, rx, stations); copy_vec(e->h_ry, ry, stations); copy_vec(e->h_rz, rz, stations);
  copy_vec(e->h_tobs, tobs, events * stations); copy_vec(e->h_std, std, events * stations);
  e->lik = LK_SRCLOC;
  return 0;
}

int hmcb_set_likelihood_srcloc2d(hmcb_engine* e, int64_t events, int64_t stations, const double* rx,
                                 const double* rz, const double* tobs, const double* std,
                                 int infer_velocity, double velocity) {
  HMCB_CHECK(e && rx && rz && tobs && std, "hmcb_set_likelihood_srcloc2d: NULL argument");
  HMCB_CHECK(!e->finalized && e->lik == LK_NONE, "one likelihood per engine, set before hmcb_finalize");
  HMCB_CHECK(events > 0 && stations > 0, "hmcb_set_likelihood_srcloc2d: bad sizes");
  HMCB_CHECK(e->d == 3 * events + (infer_velocity ? 1 : 0),
             "hmcb_set_likelihood_srcloc2d: dims must be 3*events (+1 when the velocity is inferred)");
  HMCB_CHECK(srcloc_supported((int)events, (int)stations),
             "hmcb_set_likelihood_srcloc2d: events x stations exceeds the kernel's shared-memory staging");
  e->events = events; e->stations = stations; e->infer_velocity = infer_velocity ? 1 : 0;
  e->velocity = velocity;
  e->srcloc_np = 3;
  copy_vec(e->h_rx, rx, stations); copy_vec(e->h_rz, rz, stations);
  e->h_ry.assign((size_t)stations, 0.0);   // the 2-D problem is the 3-D one in the plane y = 0
  copy_vec(e->h_tobs, tobs, events * stations); copy_vec(e->h_std, std, events * stations);
  e->lik = LK_SRCLOC;
  return 0;
}

int hmcb_finalize(hmcb_engine* e) {
  HMCB_CHECK(e, "engine is NULL");
  HMCB_CHECK(!e->finalized, "hmcb_finalize called twice (use hmcb_clear_target to start over)");
  HMCB_CUDA(cudaSetDevice(e->device));
  const int d = (int)e->d;
  build_schedule(e);

  DevTarget& T = e->T;
  std::memset(&T, 0, sizeof(T));
  T.dims = d;
  T.n_checks = (int)e->checks.size();
  {
    // per-coordinate prior-term table: slot t of coordinate j = t-th term covering j
    std::vector<int> used((size_t)d, 0);
    int n_terms = 0;
    for (const HostPrior& P : e->priors)
      for (int64_t r = 0; r < P.len; ++r) n_terms = std::max(n_terms, ++used[(size_t)(P.offset + r)]);
    T.n_terms = n_terms;
    std::vector<unsigned char> kind((size_t)std::max(n_terms, 1) * d, (unsigned char)TERM_NONE);
    std::vector<double> ta(kind.size(), 0.0), tb(kind.size(), 0.0);
    std::fill(used.begin(), used.end(), 0);
    double const_sum = 0.0;
    for (const HostPrior& P : e->priors) {
      for (int64_t r = 0; r < P.len; ++r) {
        const size_t j = (size_t)(P.offset + r);
        const size_t o = (size_t)used[j]++ * d + j;
        kind[o] = (unsigned char)(P.kind == HMCB_PRIOR_NORMAL ? TERM_NORMAL : TERM_LAPLACE);
        ta[o] = P.a[(size_t)r];
        tb[o] = P.b[(size_t)r];
      }
      const_sum += P.constant;
    }
    T.const_sum = const_sum;
    if (n_terms == 1) {
      int uk = kind[0];
      for (int j = 0; j < d; ++j)
        if (kind[(size_t)j] != uk) uk = TERM_NONE;
      T.uniform_kind = uk;
    }
    if (dev_upload(e, kind, &T.t_kind) || dev_upload(e, ta, &T.t_a) || dev_upload(e, tb, &T.t_b)) return -1;
    // bound checks: -inf / +inf where a check does not apply
    const double inf = std::numeric_limits<double>::infinity();
    const size_t nc = (size_t)std::max(T.n_checks, 1);
    std::vector<double> clb(nc * d, -inf), cub(nc * d, inf);
    std::vector<unsigned> cover((size_t)d, 0u);
    for (int k = 0; k < T.n_checks; ++k) {
      const HostCheck& Ck = e->checks[k];
      for (int64_t r = 0; r < Ck.len; ++r) {
        const size_t j = (size_t)(Ck.offset + r);
        if (Ck.has_lb) clb[(size_t)k * d + j] = Ck.lb[(size_t)r];
        if (Ck.has_ub) cub[(size_t)k * d + j] = Ck.ub[(size_t)r];
        cover[j] |= (1u << k);
      }
      if (Ck.in_gradient) T.grad_check_mask |= (1u << k);
    }
    if (dev_upload(e, clb, &T.c_lb) || dev_upload(e, cub, &T.c_ub) || dev_upload(e, cover, &T.c_cover)) return -1;
  }
  if (e->has_rlb && dev_upload(e, e->rlb, &T.refl_lb)) return -1;
  if (e->has_rub && dev_upload(e, e->rub, &T.refl_ub)) return -1;
  if (e->mass_diag) {
    std::vector<double> sq(e->h_diag.size());
    for (size_t j = 0; j < sq.size(); ++j) sq[j] = std::sqrt(e->h_diag[j]);  // MassMatrices.py:226
    if (dev_upload(e, e->h_invdiag, &T.invm) || dev_upload(e, sq, &T.sqrtm)) return -1;
  }

  // ---- path ------------------------------------------------------------------------
  if (e->lik == LK_SRCLOC) {
    HMCB_CHECK(!e->mass_full, "a Full mass matrix is not supported together with a SourceLocation likelihood");
    e->path = HMCB_PATH_FUSED_SRCLOC;
    SrcLocDev& L = e->L;
    L.events = (int)e->events; L.stations = (int)e->stations; L.infer_velocity = e->infer_velocity;
    L.np = e->srcloc_np;
    L.velocity = e->velocity;
    if (dev_upload(e, e->h_rx, &L.rx) || dev_upload(e, e->h_ry, &L.ry) || dev_upload(e, e->h_rz, &L.rz) ||
        dev_upload(e, e->h_tobs, &L.tobs) || dev_upload(e, e->h_std, &L.std))
      return -1;
  } else if (e->lik == LK_NONE && T.n_terms <= 1 && fused_priors_supported(d) && !e->mass_full &&
             !std::getenv("HMCB_FORCE_STAGED")) {
    e->path = HMCB_PATH_FUSED_PRIORS;
  } else {
    e->path = HMCB_PATH_STAGED;
    HMCB_CUDA(staged_init());
    const int C = (int)e->C;
    e->ld = round_up(C, 128);
    e->dpad = round_up(d, 128);
    e->npad = e->N ? round_up(e->N, 128) : 0;
    e->jtiles = (d + ST_DT - 1) / ST_DT;
    const size_t plane = (size_t)e->dpad * e->ld;
    if (dev_alloc(e, plane, &e->q_cur) || dev_alloc(e, plane, &e->p_w)) return -1;
    if (dev_alloc(e, (size_t)e->ld, &e->eps) || dev_alloc(e, (size_t)e->ld, &e->uacc) ||
        dev_alloc(e, (size_t)e->ld, &e->accbuf))
      return -1;
    for (int f = 0; f < 3; ++f)
      if (dev_alloc(e, (size_t)e->ld, &e->flags[f])) return -1;
    const size_t parts = (size_t)e->jtiles * e->ld;
    if (dev_alloc(e, parts, &e->k0part) || dev_alloc(e, parts, &e->k1part) || dev_alloc(e, parts, &e->upart))
      return -1;
    const double* tmp = nullptr;
    switch (e->lik) {
      case LK_NONE: e->ltiles = 0; break;
      case LK_DENSE_PREMULT:
        if (dev_upload_tiled(e, e->h_A.data(), d, d, e->dpad, e->dpad, &e->dA)) return -1;
        if (e->dpad == 128 && dev_upload_padded(e, e->h_A.data(), d, d, 128, 128, &e->dA_rowmajor)) return -1;
        if (dev_upload(e, e->h_vec, &tmp)) return -1;
        e->dvec = const_cast<double*>(tmp);
        if (oz_setup(e)) return -1;
        e->ltiles = e->dpad / GEMM_BM;
        break;
      case LK_DENSE_DIRECT:
        if (dev_upload_tiled(e, e->h_A.data(), e->N, d, e->npad, e->dpad, &e->dA)) return -1;
        if (dev_upload_tiled(e, e->h_At.data(), d, e->N, e->dpad, e->npad, &e->dAt)) return -1;
        if (oz_setup(e)) return -1;
        if (!e->h_Amis.empty()) {   // dense data covariance: the misfit pass applies U G, U d
          const double* um = nullptr;
          if (dev_upload_tiled(e, e->h_Amis.data(), e->N, d, e->npad, e->dpad, &e->dAmis) ||
              dev_upload(e, e->h_vecmis, &um)) return -1;
          e->dvecmis = const_cast<double*>(um);
        }
        e->ltiles = e->npad / GEMM_BM;
        break;
      case LK_CSR_DIRECT:
        if (upload_csr_both(e, e->csr, &e->csr_dev, &e->strip_dev) ||
            upload_csr_both(e, e->csr_t, &e->csr_t_dev, &e->strip_t_dev)) return -1;
        e->ltiles = e->use_strips ? e->strip_dev.chunks * e->strip_dev.warps : e->csr_dev.chunks * SPMM_WARPS;
        break;
      case LK_CSR_PREMULT:
        if (upload_csr_both(e, e->csr, &e->csr_dev, &e->strip_dev)) return -1;
        if (dev_upload(e, e->h_vec, &tmp)) return -1;
        e->dvec = const_cast<double*>(tmp);
        e->ltiles = e->use_strips ? e->strip_dev.chunks * e->strip_dev.warps : e->csr_dev.chunks * SPMM_WARPS;
        break;
      default: return fail("internal: bad likelihood kind");
    }
    if (e->lik == LK_DENSE_DIRECT || e->lik == LK_CSR_DIRECT) {
      const double *dv = nullptr, *vv = nullptr, *sv = nullptr;
      if (dev_upload(e, e->h_vec, &dv) || dev_upload(e, e->h_var, &vv) || dev_upload(e, e->h_sigma, &sv)) return -1;
      e->dvec = const_cast<double*>(dv); e->dvar = const_cast<double*>(vv); e->dsigma = const_cast<double*>(sv);
      // (the strip SpMM reads R through a tensor map that needs one strip stride of slack rows)
      const size_t r_rows = (size_t)e->npad + (e->use_strips && e->lik == LK_CSR_DIRECT ? e->strip_t_dev.cstride : 0);
      if (dev_alloc(e, r_rows * e->ld, &e->R)) return -1;
      if (e->use_strips && e->lik == LK_CSR_DIRECT)
        HMCB_CUDA(spmm_strip_tensor_map(e->strip_t_dev, e->R, e->ld, (long long)r_rows, &e->tmap_R));
    }
    {
      const bool strips = e->use_strips && (e->lik == LK_CSR_DIRECT || e->lik == LK_CSR_PREMULT);
      const size_t q_rows = (size_t)e->dpad + (strips ? e->strip_dev.cstride : 0);
      for (int k = 0; k < 2; ++k) {
        if (dev_alloc(e, q_rows * e->ld, &e->q_w[k])) return -1;
        if (strips) HMCB_CUDA(spmm_strip_tensor_map(e->strip_dev, e->q_w[k], e->ld, (long long)q_rows, &e->tmap_q[k]));
      }
    }
    if (e->ltiles && dev_alloc(e, (size_t)e->ltiles * e->ld, &e->lpart)) return -1;
    // small premultiplied dense models: the whole block of proposals runs in one kernel with
    // GtG resident in shared memory (the staged workspaces still serve hmcb_misfit/gradient)
    e->fused_dense = e->lik == LK_DENSE_PREMULT && e->dpad == 128 && T.n_terms <= 1 && !e->mass_full &&
                     !std::getenv("HMCB_FORCE_STAGED");
    if (e->mass_full) {
      if (dev_upload_tiled(e, e->h_L.data(), d, d, e->dpad, e->dpad, &e->dL) ||
          dev_upload_tiled(e, e->h_Minv.data(), d, d, e->dpad, e->dpad, &e->dMinv) ||
          dev_alloc(e, plane, &e->v_w)) return -1;
    }
    // the host copies of the big operands are no longer needed
    std::vector<double>().swap(e->h_A);
    std::vector<double>().swap(e->h_At);
    std::vector<double>().swap(e->h_Amis);
  }
  HMCB_CUDA(cudaDeviceSynchronize());
  e->finalized = true;
  return 0;
}

int hmcb_path(const hmcb_engine* e) {
  if (!e) return -1;
  return (e->path == HMCB_PATH_STAGED && e->fused_dense) ? HMCB_PATH_FUSED_DENSE : e->path;
}
int hmcb_dense_products_on_tcgen05(const hmcb_engine* e) {
  if (!e || !e->oz) return 0;
  // slice pairs of the products of one gradient evaluation (two in the direct form, one premultiplied): digits s of the matrix, t of the batch, s + t < orders
  int pairs = 0;
  for (int sa : {e->oz_saG, (e->lik == LK_DENSE_PREMULT || e->oz_crt) ? 0 : e->oz_saGt})
    for (int a = 0; a < sa; ++a)
      for (int b = 0; b < e->oz_sb; ++b) pairs += (a + b < e->oz_orders) ? 1 : 0;
  return pairs + (e->oz_crt ? OZ_NUM_MODULI : 0);   // G^T r as modular products
}
int64_t hmcb_grads_per_proposal(const hmcb_engine* e) { return e ? e->S.grads_per_proposal : -1; }
int64_t hmcb_launch_count(const hmcb_engine* e) { return e ? e->launches : -1; }

int hmcb_kernel_timing_begin(hmcb_engine* e) {
  HMCB_CHECK(e, "engine is NULL");
  e->ktiming = true;
  e->kev_used[0] = e->kev_used[1] = 0;
  return 0;
}

int hmcb_kernel_timing_end(hmcb_engine* e, double* total_ms, int64_t* passes) {
  HMCB_CHECK(e && total_ms && passes, "hmcb_kernel_timing_end: NULL argument");
  HMCB_CUDA(cudaSetDevice(e->device));
  e->ktiming = false;
  for (int c = 0; c < 2; ++c) {
    double sum = 0.0;
    const size_t pairs = e->kev_used[c] / 2;
    for (size_t k = 0; k < pairs; ++k) {
      float ms = 0.f;
      HMCB_CUDA(cudaEventSynchronize(e->kev[c][2 * k + 1]));
      HMCB_CUDA(cudaEventElapsedTime(&ms, e->kev[c][2 * k], e->kev[c][2 * k + 1]));
      sum += ms;
    }
    total_ms[c] = sum;
    passes[c] = (int64_t)pairs;
    e->kev_used[c] = 0;
  }
  return 0;
}

#define HMCB_READY(e)                                                      \
  HMCB_CHECK((e) != nullptr, "engine is NULL");                            \
  HMCB_CHECK((e)->finalized, "engine is not finalized (call hmcb_finalize)"); \
  HMCB_CUDA(cudaSetDevice((e)->device))

int hmcb_misfit(hmcb_engine* e, const double* q, double* x, void* stream) {
  HMCB_READY(e);
  HMCB_CHECK(q && x, "hmcb_misfit: NULL argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int C = (int)e->C, d = (int)e->d;
  if (e->path == HMCB_PATH_FUSED_PRIORS) {
    HMCB_CUDA(launch_prior_misfit(e->T, C, q, x, nullptr, s));
    e->launches += 1;
  } else if (e->path == HMCB_PATH_FUSED_SRCLOC) {
    HMCB_CUDA(launch_srcloc_eval(e->T, e->L, C, 0, q, x, s));
    e->launches += 1;
  } else {
    const bool any_checks = e->T.n_checks > 0;
    HMCB_CUDA(launch_st_transpose(q, C, d, d, e->q_w[0], e->ld, s));
    if (any_checks) HMCB_CUDA(cudaMemsetAsync(e->flags[2], 0, (size_t)e->ld * sizeof(unsigned), s));
    HMCB_CUDA(launch_st_energy(staged_common(e, nullptr), e->q_w[0], nullptr, nullptr, e->upart,
                               any_checks ? e->flags[2] : nullptr, s));
    e->launches += 2;
    if (staged_misfit_pass(e, e->q_w[0], s)) return -1;
    DecideArgs D = decide_args(e);
    D.flags = any_checks ? e->flags[2] : nullptr;
    D.x = x; D.misfit_only = 1;
    HMCB_CUDA(launch_st_decide(D, s));
    e->launches += 1;
  }
  return 0;
}

int hmcb_gradient(hmcb_engine* e, const double* q, double* g, void* stream) {
  HMCB_READY(e);
  HMCB_CHECK(q && g, "hmcb_gradient: NULL argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int C = (int)e->C, d = (int)e->d;
  if (e->path == HMCB_PATH_FUSED_PRIORS) {
    HMCB_CUDA(launch_prior_gradient(e->T, C, q, g, 0, s));
    e->launches += 1;
  } else if (e->path == HMCB_PATH_FUSED_SRCLOC) {
    HMCB_CUDA(launch_srcloc_eval(e->T, e->L, C, 1, q, g, s));
    e->launches += 1;
  } else {
    const bool grad_checks = e->T.grad_check_mask != 0u;
    HMCB_CUDA(launch_st_transpose(q, C, d, d, e->q_w[0], e->ld, s));
    e->launches += 1;
    if (grad_checks) {
      HMCB_CUDA(cudaMemsetAsync(e->flags[2], 0, (size_t)e->ld * sizeof(unsigned), s));
      HMCB_CUDA(launch_st_energy(staged_common(e, nullptr), e->q_w[0], nullptr, nullptr, e->upart,
                                 e->flags[2], s));
      e->launches += 1;
    }
    UpdateEpi epi{};
    epi.T = e->T; epi.C = C; epi.ld = e->ld; epi.p = e->p_w; epi.q_out = e->q_w[1];
    epi.flags_in = grad_checks ? e->flags[2] : nullptr;
    epi.grad_only = 1;
    if (staged_gradient_pass(e, e->q_w[0], epi, s)) return -1;
    HMCB_CUDA(launch_st_transpose(e->p_w, d, C, e->ld, g, d, s));
    e->launches += 1;
  }
  return 0;
}

int hmcb_reflect(hmcb_engine* e, double* q, double* p, void* stream) {
  HMCB_READY(e);
  HMCB_CHECK(q && p, "hmcb_reflect: NULL argument");
  HMCB_CUDA(launch_reflect(e->T, (int)e->C, q, p, static_cast<cudaStream_t>(stream)));
  e->launches += 1;
  return 0;
}

// Full mass matrix: out = A . in for chain-major [C x d] batches through the transposed planes
static int full_mass_apply(hmcb_engine* e, const double* A_tiled, const double* in, double* out, cudaStream_t s) {
  const int C = (int)e->C, d = (int)e->d;
  HMCB_CUDA(launch_st_transpose(in, C, d, d, e->q_w[1], e->ld, s));
  StoreEpi st{d, C, e->ld, e->v_w};
  HMCB_CUDA(launch_gemm_store(A_tiled, e->dpad, e->dpad, e->q_w[1], e->ld, e->dpad, st, s));
  if (out) HMCB_CUDA(launch_st_transpose(e->v_w, d, C, e->ld, out, d, s));
  e->launches += out ? 3 : 2;
  return 0;
}

int hmcb_scale_momentum(hmcb_engine* e, const double* z, double* p, void* stream) {
  HMCB_READY(e);
  HMCB_CHECK(z && p, "hmcb_scale_momentum: NULL argument");
  if (e->mass_full) return full_mass_apply(e, e->dL, z, p, static_cast<cudaStream_t>(stream));
  HMCB_CUDA(launch_mass_elementwise(e->T, (int)e->C, 0, z, p, static_cast<cudaStream_t>(stream)));
  e->launches += 1;
  return 0;
}

int hmcb_kinetic_energy(hmcb_engine* e, const double* p, double* k, void* stream) {
  HMCB_READY(e);
  HMCB_CHECK(p && k, "hmcb_kinetic_energy: NULL argument");
  if (e->mass_full) {   // K = 0.5 p . M^-1 p: v_w = M^-1 p, p is still in q_w[1]
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (full_mass_apply(e, e->dMinv, p, nullptr, s)) return -1;
    StagedCommon SV = staged_common(e, nullptr);
    SV.v = e->v_w;
    HMCB_CUDA(launch_st_energy(SV, e->q_w[1], e->q_w[1], e->k1part, e->upart, nullptr, s));
    HMCB_CUDA(launch_st_colsum(e->k1part, e->jtiles, e->ld, (int)e->C, 0.5, k, s));
    e->launches += 2;
    return 0;
  }
  HMCB_CUDA(launch_kinetic_energy(e->T, (int)e->C, p, k, static_cast<cudaStream_t>(stream)));
  e->launches += 1;
  return 0;
}

int hmcb_kinetic_gradient(hmcb_engine* e, const double* p, double* dk, void* stream) {
  HMCB_READY(e);
  HMCB_CHECK(p && dk, "hmcb_kinetic_gradient: NULL argument");
  if (e->mass_full) return full_mass_apply(e, e->dMinv, p, dk, static_cast<cudaStream_t>(stream));
  HMCB_CUDA(launch_mass_elementwise(e->T, (int)e->C, 1, p, dk, static_cast<cudaStream_t>(stream)));
  e->launches += 1;
  return 0;
}

int hmcb_run_block(hmcb_engine* e, const hmcb_block* b, void* stream) {
  HMCB_READY(e);
  HMCB_CHECK(b, "hmcb_run_block: block is NULL");
  HMCB_CHECK(b->proposals > 0 && b->proposals < (1ll << 31), "hmcb_run_block: proposals must be positive");
  HMCB_CHECK(b->thinning > 0, "hmcb_run_block: thinning must be positive");
  HMCB_CHECK(b->proposal_offset >= 0 && b->chain_offset >= 0, "hmcb_run_block: negative offset");
  HMCB_CHECK(b->stepsize > 0.0 || b->stepsize_chain, "hmcb_run_block: stepsize must be positive");
  HMCB_CHECK(!b->autotune || b->stepsize_chain, "hmcb_run_block: autotune needs stepsize_chain");
  HMCB_CHECK(!b->autotune || (b->learning_rate > 0.5 && b->learning_rate <= 1.0),
             "The learning rate should be larger than 0.5 and smaller than or equal to 1.0");
  HMCB_CHECK(b->q && b->x, "hmcb_run_block: q and x are required");
  HMCB_CHECK((b->trace_q == nullptr) == (b->trace_g == nullptr), "hmcb_run_block: trace_q and trace_g go together");
  HMCB_CHECK((b->out_q_prop == nullptr) == (b->out_p_prop == nullptr),
             "hmcb_run_block: out_q_prop and out_p_prop go together");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (e->path == HMCB_PATH_FUSED_PRIORS) {
    KernelTimer timer(e, s, 0);
    HMCB_CUDA(launch_fused_priors(fused_args(e, b), s));
    e->launches += 1;
    return 0;
  }
  if (e->path == HMCB_PATH_FUSED_SRCLOC) {
    KernelTimer timer(e, s, 0);
    HMCB_CUDA(launch_fused_srcloc(fused_args(e, b), e->L, s));
    e->launches += 1;
    return 0;
  }
  if (e->fused_dense) {
    KernelTimer timer(e, s, 0);
    HMCB_CUDA(launch_fused_dense(fused_args(e, b), e->dA_rowmajor, e->dvec, e->dtd, s));
    e->launches += 1;
    return 0;
  }
  return staged_run_block(e, b, s);
}

int hmcb_run_block_rwmh(hmcb_engine* e, const hmcb_block* b, const double* step_vector, void* stream) {
  HMCB_READY(e);
  HMCB_CHECK(b, "hmcb_run_block_rwmh: block is NULL");
  HMCB_CHECK(b->proposals > 0 && b->proposals < (1ll << 31), "hmcb_run_block_rwmh: proposals must be positive");
  HMCB_CHECK(b->thinning > 0, "hmcb_run_block_rwmh: thinning must be positive");
  HMCB_CHECK(b->proposal_offset >= 0 && b->chain_offset >= 0, "hmcb_run_block_rwmh: negative offset");
  HMCB_CHECK(b->stepsize > 0.0 || b->stepsize_chain, "hmcb_run_block_rwmh: stepsize must be positive");
  HMCB_CHECK(b->q && b->x, "hmcb_run_block_rwmh: q and x are required");
  HMCB_CHECK(!b->autotune || b->stepsize_chain, "hmcb_run_block_rwmh: autotune needs stepsize_chain");
  HMCB_CHECK(!b->autotune || (b->learning_rate > 0.5 && b->learning_rate <= 1.0),
             "The learning rate should be larger than 0.5 and smaller than or equal to 1.0");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int C = (int)e->C, d = (int)e->d;
  if (!e->rw_qp) {
    if (dev_alloc(e, (size_t)C * d, &e->rw_qp) || dev_alloc(e, (size_t)C, &e->rw_x1)) return -1;
  }
  const int64_t first_row = (b->proposal_offset + b->thinning - 1) / b->thinning;
  for (int64_t kb = 0; kb < b->proposals; ++kb) {
    const int64_t kglob = b->proposal_offset + kb;
    const size_t kc = (size_t)kb * C;
    HMCB_CUDA(launch_rwmh_propose(C, d, b->q, e->rw_qp, step_vector, b->stepsize_chain, b->stepsize,
                                  b->z_in ? b->z_in + kc * d : nullptr, b->seed, b->chain_offset, kglob, s));
    e->launches += 1;
    if (hmcb_misfit(e, e->rw_qp, e->rw_x1, stream)) return -1;
    if (b->out_q_prop)
      HMCB_CUDA(cudaMemcpyAsync(b->out_q_prop + kc * d, e->rw_qp, sizeof(double) * (size_t)C * d,
                                cudaMemcpyDeviceToDevice, s));
    RwmhDecide D{};
    D.chains = C; D.dims = d; D.q = b->q; D.qp = e->rw_qp; D.x = b->x; D.x1 = e->rw_x1;
    D.u_in = b->u_accept_in ? b->u_accept_in + kc : nullptr;
    D.seed = b->seed; D.chain_offset = b->chain_offset; D.kglob = kglob;
    if (b->out_samples && (kglob % b->thinning) == 0)
      D.sample_rows = b->out_samples + (size_t)(kglob / b->thinning - first_row) * C * (size_t)(d + 1);
    D.out_accept = b->out_accept ? b->out_accept + kc : nullptr;
    D.out_h0 = b->out_h0 ? b->out_h0 + kc : nullptr;
    D.out_h1 = b->out_h1 ? b->out_h1 + kc : nullptr;
    D.accepted_total = b->accepted_total;
    D.stepsize_chain = b->stepsize_chain;
    D.out_stepsize = b->out_stepsize ? b->out_stepsize + kc : nullptr;
    D.tune = AutotuneArgs{b->autotune ? 1 : 0, b->target_acceptance_rate, b->learning_rate};
    HMCB_CUDA(launch_rwmh_decide(D, s));
    e->launches += 1;
  }
  return 0;
}

int hmcb_sample_host(hmcb_engine* e, const double* q0_host, int64_t proposals, int64_t thinning,
                     int64_t block_proposals, double stepsize, int randomize_stepsize, uint64_t seed,
                     int64_t chain_offset, double* samples_host, int32_t* accept_host,
                     double* final_q_host, double* final_x_host) {
  HMCB_READY(e);
  HMCB_CHECK(q0_host, "hmcb_sample_host: q0_host is NULL");
  HMCB_CHECK(proposals > 0 && thinning > 0 && proposals % thinning == 0,
             "hmcb_sample_host: proposals must be a positive multiple of thinning");
  HMCB_CHECK(stepsize > 0.0, "hmcb_sample_host: stepsize must be positive");
  if (block_proposals <= 0) block_proposals = thinning;
  block_proposals = (block_proposals + thinning - 1) / thinning * thinning;  // whole stored rows per block
  const size_t C = (size_t)e->C, d = (size_t)e->d;
  const size_t row_doubles = C * (d + 1);
  const size_t rows_per_block = (size_t)(block_proposals / thinning);
  if (!e->s_compute) HMCB_CUDA(cudaStreamCreateWithFlags(&e->s_compute, cudaStreamNonBlocking));
  if (!e->s_copy) HMCB_CUDA(cudaStreamCreateWithFlags(&e->s_copy, cudaStreamNonBlocking));

  int rc = 0;
  auto cleanup = [&]() {
    cudaStreamSynchronize(e->s_compute);
    cudaStreamSynchronize(e->s_copy);
  };
#define HMCB_CUDA_C(expr)                                                             \
  do {                                                                                \
    cudaError_t err__ = (expr);                                                       \
    if (err__ != cudaSuccess) {                                                       \
      fail(std::string(#expr) + " failed: " + cudaGetErrorString(err__));             \
      cleanup();                                                                      \
      return -1;                                                                      \
    }                                                                                 \
  } while (0)
  if (!e->sh_q) {
    HMCB_CUDA_C(cudaMalloc(&e->sh_q, C * d * sizeof(double)));
    HMCB_CUDA_C(cudaMalloc(&e->sh_x, C * sizeof(double)));
    HMCB_CUDA_C(cudaMalloc(&e->sh_acc, C * sizeof(int32_t)));
    for (int i = 0; i < 2; ++i) {
      HMCB_CUDA_C(cudaEventCreateWithFlags(&e->ev_produced[i], cudaEventDisableTiming));
      HMCB_CUDA_C(cudaEventCreateWithFlags(&e->ev_drained[i], cudaEventDisableTiming));
    }
  }
  if (samples_host && e->sh_buf_doubles < rows_per_block * row_doubles) {
    for (int i = 0; i < 2; ++i) {
      if (e->sh_buf[i]) cudaFree(e->sh_buf[i]);
      e->sh_buf[i] = nullptr;
    }
    e->sh_buf_doubles = 0;
    HMCB_CUDA_C(cudaMalloc(&e->sh_buf[0], rows_per_block * row_doubles * sizeof(double)));
    HMCB_CUDA_C(cudaMalloc(&e->sh_buf[1], rows_per_block * row_doubles * sizeof(double)));
    e->sh_buf_doubles = rows_per_block * row_doubles;
  }
  double *dq = e->sh_q, *dx = e->sh_x, **dbuf = e->sh_buf;
  int32_t* dacc = e->sh_acc;
  cudaEvent_t *produced = e->ev_produced, *drained = e->ev_drained;
  HMCB_CUDA_C(cudaMemcpyAsync(dq, q0_host, C * d * sizeof(double), cudaMemcpyHostToDevice, e->s_compute));
  HMCB_CUDA_C(cudaMemsetAsync(dacc, 0, C * sizeof(int32_t), e->s_compute));
  rc = hmcb_misfit(e, dq, dx, e->s_compute);
  if (rc) { cleanup(); return rc; }

  int64_t done = 0, nblock = 0;
  size_t rows_done = 0;
  while (done < proposals) {
    const int64_t B = std::min<int64_t>(block_proposals, proposals - done);
    const int slot = (int)(nblock & 1);
    hmcb_block blk;
    std::memset(&blk, 0, sizeof(blk));
    blk.proposals = B; blk.thinning = thinning; blk.proposal_offset = done; blk.chain_offset = chain_offset;
    blk.seed = seed; blk.stepsize = stepsize; blk.randomize_stepsize = randomize_stepsize;
    blk.q = dq; blk.x = dx; blk.accepted_total = dacc;
    blk.out_samples = samples_host ? dbuf[slot] : nullptr;
    if (samples_host && nblock >= 2) HMCB_CUDA_C(cudaStreamWaitEvent(e->s_compute, drained[slot], 0));
    rc = hmcb_run_block(e, &blk, e->s_compute);
    if (rc) { cleanup(); return rc; }
    if (samples_host) {
      const size_t rows = (size_t)(B / thinning);
      HMCB_CUDA_C(cudaEventRecord(produced[slot], e->s_compute));
      HMCB_CUDA_C(cudaStreamWaitEvent(e->s_copy, produced[slot], 0));
      HMCB_CUDA_C(cudaMemcpyAsync(samples_host + rows_done * row_doubles, dbuf[slot],
                                  rows * row_doubles * sizeof(double), cudaMemcpyDeviceToHost, e->s_copy));
      HMCB_CUDA_C(cudaEventRecord(drained[slot], e->s_copy));
      rows_done += rows;
    }
    done += B;
    ++nblock;
  }
  if (accept_host)
    HMCB_CUDA_C(cudaMemcpyAsync(accept_host, dacc, C * sizeof(int32_t), cudaMemcpyDeviceToHost, e->s_compute));
  if (final_q_host)
    HMCB_CUDA_C(cudaMemcpyAsync(final_q_host, dq, C * d * sizeof(double), cudaMemcpyDeviceToHost, e->s_compute));
  if (final_x_host)
    HMCB_CUDA_C(cudaMemcpyAsync(final_x_host, dx, C * sizeof(double), cudaMemcpyDeviceToHost, e->s_compute));
  HMCB_CUDA_C(cudaStreamSynchronize(e->s_compute));
  HMCB_CUDA_C(cudaStreamSynchronize(e->s_copy));
  cleanup();
#undef HMCB_CUDA_C
  return 0;
}

}  // extern "C"

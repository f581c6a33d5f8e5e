// Fused dual-ascent evaluation for block-structured matching LPs on sm_100a.
//
// One launch of matching_pass_kernel replaces the ~15 eager ATen passes of the reference's
// MatchingSolverDualObjectiveFunction.calculate (reference src/dualip/objectives/matching.py:116-188):
//   v = a*(-lambda[row]/gamma) + (-c/gamma)        matching.py:136-142, utils/sparse_utils.py:54-85,26-51
//   x = Proj_column(v)                              sparse_utils.py:133-220 -> projections/{box,cone,simplex}.py
//   grad[row] += a*x ; cx += c*x ; xx += x*x        matching.py:153-160, sparse_utils.py:223-243
//   grad -= b ; dual_obj = cx + gamma/2*xx + lambda.grad ; slacks      matching.py:25-34,164-178
//
// Layout ("pass table"): the nonzeros of consecutive short columns (1..32 entries) are packed greedily
// into passes of <= 32 nonzeros that begin and end on column boundaries.  A warp handles one pass with
// one lane per nonzero: loads of a, c, row are fully coalesced, every per-column reduction is a
// segmented warp scan, and nothing per nonzero ever touches shared memory except the lambda gather
// and the gradient scatter.  ccol is never read by the hot kernel: a pass entry is
// {first nnz, end-of-column bit mask}.  Columns with more than 32 entries go to a warp-per-column kernel.
#include <cub/device/device_scan.cuh>
#include <algorithm>
#include <new>
#include <type_traits>
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace dualip {

static thread_local std::string g_last_error;
void set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
}

constexpr int kPassWidth = 32;        // nonzeros per pass = lanes per warp
constexpr int kChunkCols = 2048;      // columns walked by one builder thread; passes never straddle chunks
constexpr int kPrefetch = 4;          // passes in flight per warp
constexpr int kMaxClasses = 255;
constexpr size_t kSmemBudget = 227 * 1024;

struct LongCol {
  int64_t start;
  int32_t len;
  int32_t cls;
};

struct PassEntryU {  // uniform projection map: 8 bytes / pass
  uint32_t start;
  uint32_t endmask;
};
struct PassEntryC {  // mixed projection map: 16 bytes / pass
  uint32_t start;
  uint32_t endmask;
  uint32_t colbase;  // ordinal (among non-empty short columns) of the pass's first column
  uint32_t pad;
};

}  // namespace dualip

using namespace dualip;

struct dualip_plan {
  int device = 0;
  int64_t n_cols = 0, nnz = 0;
  int32_t m = 0;
  const float* a = nullptr;  // borrowed
  const float* c = nullptr;  // borrowed
  void* row = nullptr;       // owned, narrowed
  int row_bits = 32;
  bool uniform = true;
  void* entries = nullptr;   // PassEntryU[] or PassEntryC[], padded with empty passes
  int64_t n_passes = 0;      // real passes
  int64_t n_runs = 0;
  int run_len = 32;
  uint8_t* cls_ne = nullptr;  // class id per non-empty short column (mixed maps only)
  LongCol* longcols = nullptr;
  int64_t n_long = 0;
  dualip_proj_class classes_host[kMaxClasses];
  dualip_proj_class* classes_dev = nullptr;
  int n_classes = 0;
  float* acc = nullptr;        // m floats, zero between calls
  double* acc_scal = nullptr;  // [c.x, ||x||^2], zero between calls
  unsigned int* counter = nullptr;
  float* lambda_stage = nullptr;  // m floats, for *_calc_host
  float* grad_stage = nullptr;
  dualip_scalars* scal_stage = nullptr;
  int n_sms = 0, n_ctas = 0, threads = 512, smode = 0;
  size_t smem_bytes = 0;
  size_t owned_bytes = 0;
  int flush_bulk = 1;
};

namespace dualip {

// ------------------------------------------------------------------------------------------
// Plan construction (setup time)
// ------------------------------------------------------------------------------------------
template <typename IdxT, typename OutT>
__global__ void narrow_rows_kernel(const IdxT* __restrict__ in, OutT* __restrict__ out, int64_t n, int32_t m,
                                   unsigned int* bad) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    IdxT v = in[i];
    if (v < 0 || v >= (IdxT)m) atomicOr(bad, 1u);
    out[i] = (OutT)v;
  }
}

// Walks one chunk of columns and either counts or emits passes / long columns / class ids.
template <typename IdxT, bool EMIT, bool UNIFORM>
__global__ void build_passes_kernel(const IdxT* __restrict__ ccol, const uint8_t* __restrict__ col_class, int64_t n_cols,
                                    int64_t n_chunks, unsigned long long* __restrict__ counts /* 3*n_chunks */,
                                    const unsigned long long* __restrict__ offsets /* 3*n_chunks */, void* entries,
                                    uint8_t* cls_ne, LongCol* longcols, unsigned int* bad) {
  const int64_t chunk = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (chunk >= n_chunks) return;
  const int64_t j0 = chunk * kChunkCols;
  const int64_t j1 = min(j0 + (int64_t)kChunkCols, n_cols);
  unsigned long long n_pass = 0, n_ne = 0, n_long = 0;
  unsigned long long o_pass = 0, o_ne = 0, o_long = 0;
  if (EMIT) {
    o_pass = offsets[3 * chunk + 0];
    o_ne = offsets[3 * chunk + 1];
    o_long = offsets[3 * chunk + 2];
  }
  int cur = 0;  // nonzeros in the open pass
  uint32_t endmask = 0;
  int64_t pass_start = 0;
  unsigned long long pass_colbase = 0;
  IdxT lo = ccol[j0];
  auto close_pass = [&]() {
    if (cur == 0) return;
    if (EMIT) {
      if (UNIFORM) {
        PassEntryU e{(uint32_t)pass_start, endmask};
        reinterpret_cast<PassEntryU*>(entries)[o_pass + n_pass] = e;
      } else {
        PassEntryC e{(uint32_t)pass_start, endmask, (uint32_t)pass_colbase, 0u};
        reinterpret_cast<PassEntryC*>(entries)[o_pass + n_pass] = e;
      }
    }
    ++n_pass;
    cur = 0;
    endmask = 0;
  };
  for (int64_t j = j0; j < j1; ++j) {
    const IdxT hi = ccol[j + 1];
    const int64_t d = (int64_t)hi - (int64_t)lo;
    if (d < 0) atomicOr(bad, 2u);
    if (d > kPassWidth) {
      close_pass();
      if (EMIT) {
        LongCol lc;
        lc.start = (int64_t)lo;
        lc.len = (int32_t)d;
        lc.cls = col_class ? (int32_t)col_class[j] : 0;
        longcols[o_long + n_long] = lc;
      }
      if (d > 0x7fffffffLL) atomicOr(bad, 4u);
      ++n_long;
    } else if (d > 0) {
      if (cur + (int)d > kPassWidth) close_pass();
      if (cur == 0) {
        pass_start = (int64_t)lo;
        pass_colbase = o_ne + n_ne;
      }
      cur += (int)d;
      endmask |= 1u << (cur - 1);
      if (EMIT && !UNIFORM) cls_ne[o_ne + n_ne] = col_class[j];
      ++n_ne;
    }
    lo = hi;
  }
  close_pass();
  if (!EMIT) {
    counts[3 * chunk + 0] = n_pass;
    counts[3 * chunk + 1] = n_ne;
    counts[3 * chunk + 2] = n_long;
  }
}

// ------------------------------------------------------------------------------------------
// Hot kernel
// ------------------------------------------------------------------------------------------
struct KArgs {
  const float* a;
  const float* c;
  const void* row;
  const void* entries;
  const uint8_t* cls_ne;
  const dualip_proj_class* classes;
  int n_classes;
  const float* lambda;
  const float* b;            // may be null
  float* acc;                // m floats (global accumulator across CTAs)
  double* acc_scal;          // 2 doubles
  unsigned int* counter;
  float* grad_out;           // calc mode
  dualip_scalars* scalars_out;
  float* partial_out;        // partial mode: m+2 floats
  float* x_out;              // may be null
  uint8_t* diag;             // may be null
  int64_t n_runs;
  int run_len;
  int m;
  double gamma;
  float s;                   // fl32(-1/gamma)   (matching.py:136: `-1.0 / self.gamma * dual_val`)
  int flush_bulk;
  int do_epilogue;           // 1: calc (grad/scalars), 0: partial (packed sums)
};

struct Slot {
  float a, c;
  uint32_t r;
  uint32_t cls;
  uint32_t start, endmask;
};

template <bool ROW16, bool UNIFORM>
__device__ __forceinline__ void load_slot(Slot& s, const KArgs& k, uint32_t start, uint32_t endmask, uint32_t colbase,
                                          int lane) {
  s.start = start;
  s.endmask = endmask;
  const int cnt = 32 - __clz(endmask);
  s.a = 0.f;
  s.c = 0.f;
  s.r = 0u;
  s.cls = 0u;
  if (lane < cnt) {
    const size_t idx = (size_t)start + lane;
    s.a = __ldg(k.a + idx);
    s.c = __ldg(k.c + idx);
    if (ROW16)
      s.r = __ldg(reinterpret_cast<const unsigned short*>(k.row) + idx);
    else
      s.r = __ldg(reinterpret_cast<const uint32_t*>(k.row) + idx);
    if (!UNIFORM) {
      const uint32_t headbits = (endmask << 1) | 1u;
      const uint32_t le = 0xffffffffu >> (31 - lane);
      s.cls = __ldg(k.cls_ne + (size_t)colbase + (__popc(headbits & le) - 1));
    }
  }
}

// Everything per pass.  SMODE 0: scaled lambda and grad accumulator in smem; 1: grad accumulator in smem,
// lambda gathered from global (L2); 2: neither (global atomics; very large m).
template <bool UNIFORM, int SMODE>
__device__ __forceinline__ void process_pass(const Slot& sl, const KArgs& k, const float* s_lam, float* s_grad,
                                             const dualip_proj_class* s_cls, int lane, double& cx, double& xx) {
  const unsigned FULL = 0xffffffffu;
  const uint32_t endmask = sl.endmask;
  const int cnt = 32 - __clz(endmask);
  if (cnt == 0) return;  // warp-uniform (padding pass)
  const bool valid = lane < cnt;
  const uint32_t headbits = (endmask << 1) | 1u;
  const uint32_t le = 0xffffffffu >> (31 - lane);
  int seg_start = 31 - __clz(headbits & le);
  int seg_end = __ffs(endmask & (0xffffffffu << lane)) - 1;
  if (!valid) {
    seg_start = lane;
    seg_end = lane;
  }
  const float av = sl.a, cv = sl.c;
  const uint32_t rv = sl.r;
  float lam_s;
  if (SMODE == 0)
    lam_s = s_lam[rv];
  else
    lam_s = __fmul_rn(k.s, __ldg(k.lambda + rv));
  // v = fl(fl(a * fl(s*lambda_r)) + fl(s*c)): same operation order as the reference, no FMA contraction.
  const float v = __fadd_rn(__fmul_rn(av, lam_s), __fmul_rn(k.s, cv));

  const dualip_proj_class pc = s_cls[UNIFORM ? 0u : sl.cls];
  const bool is_sx = valid && (pc.kind != DUALIP_PROJ_CLAMP);
  float x = fminf(fmaxf(v, pc.lo), pc.hi);  // box.py:16, cone.py:22-28 (lo/hi = -+inf when open)
  int branch = 0, rho = 0;

  if (__any_sync(FULL, is_sx)) {
    // ---- batched Duchi with pre-clamp, one column per lane segment (simplex.py:143-236) ----
    const float u = is_sx ? fmaxf(v, 0.f) : 0.f;               // simplex.py:148
    const float un = is_sx ? __fdiv_rn(u, pc.z) : 0.f;         // simplex.py:172 (top-2 test on u/z)
    float S = u, m1 = un, m2 = 0.f;                            // m2 starts at the zero padding value
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const float tS = __shfl_up_sync(FULL, S, d);
      const float t1 = __shfl_up_sync(FULL, m1, d);
      const float t2 = __shfl_up_sync(FULL, m2, d);
      if (lane - d >= seg_start) {
        S = __fadd_rn(tS, S);
        const float mn = fminf(m1, t1);
        m1 = fmaxf(m1, t1);
        m2 = fmaxf(fmaxf(m2, t2), mn);
      }
    }
    S = __shfl_sync(FULL, S, seg_end);
    m1 = __shfl_sync(FULL, m1, seg_end);
    m2 = __shfl_sync(FULL, m2, seg_end);
    const int deg = seg_end - seg_start + 1;
    const uint32_t segmask = (0xffffffffu >> (31 - seg_end)) & (0xffffffffu << seg_start);
    const bool feasible = (pc.kind == DUALIP_PROJ_SIMPLEX) && (S <= pc.z_thr);                 // simplex.py:153-155
    const bool padded = (deg > 1) || !(pc.flags & DUALIP_PROJ_FLAG_D1_UNPADDED);                 // simplex.py:166
    const bool shortcut = !feasible && padded && (__fsub_rn(m1, m2) > 1.0f);                     // simplex.py:178
    const uint32_t eqb = __ballot_sync(FULL, is_sx && (un == m1)) & segmask;
    const int amax = __ffs(eqb) - 1;
    const bool need = is_sx && !feasible && !shortcut;
    float theta = 0.f;
    if (__any_sync(FULL, need)) {
      // Rank every entry inside its column by all-pairs comparison (ties broken by position), and take
      // the prefix sum in sorted order in fp64 like torch's CPU cumsum does (acc_type<float> = double).
      const int maxlen = __reduce_max_sync(FULL, need ? deg : 0);
      int rank = 0;
      double cs = 0.0;
      for (int t = 0; t < maxlen; ++t) {
        const int p = seg_start + t;
        const float up = __shfl_sync(FULL, u, p & 31);
        if (p <= seg_end && (up > u || (up == u && p <= lane))) {
          ++rank;
          cs += (double)up;
        }
      }
      const float css = (float)cs;
      // cond_i = u_(i) - (css_i - z)/i > 0 ; rho = max i with cond   (simplex.py:221-225)
      const bool cond = need && (__fsub_rn(u, __fdiv_rn(__fsub_rn(css, pc.z), (float)rank)) > 0.f);
      int rr = cond ? rank : 0;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(FULL, rr, d);
        if (lane - d >= seg_start) rr = max(rr, t);
      }
      rho = max(__shfl_sync(FULL, rr, seg_end), 1);
      const uint32_t rb = __ballot_sync(FULL, need && rank == rho) & segmask;
      const int src = __ffs(rb) - 1;
      const float cssr = __shfl_sync(FULL, css, src < 0 ? lane : src);
      theta = __fdiv_rn(__fsub_rn(cssr, pc.z), (float)rho);                                      // simplex.py:228-230
    }
    if (is_sx) {
      if (feasible) {
        x = u;
        branch = 0;
      } else if (shortcut) {
        x = (lane == amax) ? pc.z : 0.f;                                                         // simplex.py:185-190
        branch = 1;
        rho = 1;
      } else {
        x = fmaxf(__fsub_rn(u, theta), 0.f);                                                     // simplex.py:233
        branch = 2;
      }
    }
  }

  if (valid) {
    const float g = __fmul_rn(av, x);  // matching.py:153 (elementwise mul), then row sums
    if (g != 0.f) {
      if (SMODE <= 1)
        atomicAdd(&s_grad[rv], g);
      else
        atomicAdd(&k.acc[rv], g);
    }
    const double xd = (double)x;
    cx = fma((double)cv, xd, cx);
    xx = fma(xd, xd, xx);
    const size_t idx = (size_t)sl.start + lane;
    if (k.x_out) k.x_out[idx] = x;
    if (k.diag && is_sx && lane == seg_start) k.diag[idx] = (uint8_t)(branch | (min(rho, 63) << 2));
  }
}

// m-length tail, executed by one whole CTA.  `sum` = sum_j a_rj x_rj (m floats), cxv = c.x, xxv = ||x||^2.
// Reference: calc_grad (matching.py:25-34) and matching.py:164-178 / :280-299.
__device__ void cta_epilogue(const float* sum, double cxv, double xxv, const float* lambda, const float* b, int m,
                             double gamma, float* grad_out, dualip_scalars* out, double* dscratch, float* fscratch) {
  double lg = 0.0, sp = 0.0, g2 = 0.0;
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < m; i += blockDim.x) {
    const float raw = __ldcg(sum + i);
    const float g = b ? __fsub_rn(raw, b[i]) : raw;
    grad_out[i] = g;
    lg = fma((double)lambda[i], (double)g, lg);
    sp += (double)fmaxf(g, 0.f);
    g2 = fma((double)g, (double)g, g2);
    mx = fmaxf(mx, g);
  }
  lg = block_sum(lg, dscratch);
  sp = block_sum(sp, dscratch);
  g2 = block_sum(g2, dscratch);
  mx = block_max(mx, fscratch);
  if (threadIdx.x == 0) {
    const double reg = 0.5 * gamma * xxv;
    dualip_scalars r;
    r.primal_objective = cxv;
    r.reg_penalty = reg;
    r.dual_val_times_grad = lg;
    r.dual_objective = cxv + reg + lg;
    r.max_pos_slack = (double)fmaxf(mx, 0.f);
    r.sum_pos_slack = sp;
    r.x_sq_norm = xxv;
    r.grad_sq_norm = g2;
    *out = r;
  }
}

template <bool ROW16, bool UNIFORM, int SMODE, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) matching_pass_kernel(const KArgs k) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // carve: [mbarrier 16B][classes][scratch 32 doubles][s_lam m_pad floats][s_grad m floats]
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  dualip_proj_class* s_cls = reinterpret_cast<dualip_proj_class*>(smem_raw + 16);
  const int n_cls_bytes = ((k.n_classes * (int)sizeof(dualip_proj_class)) + 15) & ~15;
  double* dscratch = reinterpret_cast<double*>(smem_raw + 16 + n_cls_bytes);
  float* fscratch = reinterpret_cast<float*>(dscratch + 32);
  float* s_lam = fscratch + 32;
  const int m = k.m;
  const int m_pad = (m + 3) & ~3;
  float* s_grad = s_lam + (SMODE == 0 ? m_pad : 0);
  __shared__ unsigned int s_ticket;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = THREADS / 32;

  // ---- stage lambda into shared memory with a bulk async copy (TMA engine), then scale by -1/gamma ----
  bool bulk_ok = false;
  if (SMODE == 0) {
    const uint32_t bulk_bytes = (uint32_t)(m & ~3) * 4u;
    bulk_ok = ((reinterpret_cast<uintptr_t>(k.lambda) & 15u) == 0) && bulk_bytes >= 16;
    if (tid == 0) {
      mbar_init(bar, 1);
      fence_mbar_init();
    }
    __syncthreads();
    if (bulk_ok && tid == 0) {
      mbar_expect_tx(bar, bulk_bytes);
      bulk_g2s(s_lam, k.lambda, bulk_bytes, bar);
    }
  }
  for (int i = tid; i < k.n_classes * (int)(sizeof(dualip_proj_class) / 4); i += THREADS)
    reinterpret_cast<uint32_t*>(s_cls)[i] = reinterpret_cast<const uint32_t*>(k.classes)[i];
  if (SMODE <= 1)
    for (int i = tid; i < m; i += THREADS) s_grad[i] = 0.f;
  if (SMODE == 0) {
    if (bulk_ok) {
      mbar_wait(bar, 0);
      for (int i = (m & ~3) + tid; i < m; i += THREADS) s_lam[i] = k.lambda[i];
      __syncthreads();
      for (int i = tid; i < m; i += THREADS) s_lam[i] = __fmul_rn(k.s, s_lam[i]);
    } else {
      for (int i = tid; i < m; i += THREADS) s_lam[i] = __fmul_rn(k.s, k.lambda[i]);
    }
  }
  __syncthreads();

  // ---- stream this CTA's share of the pass table ----
  double cx = 0.0, xx = 0.0;
  {
    const int64_t run_begin = (k.n_runs * (int64_t)blockIdx.x) / gridDim.x;
    const int64_t run_end = (k.n_runs * (int64_t)(blockIdx.x + 1)) / gridDim.x;
    const int RL = k.run_len;
    using Entry = typename std::conditional<UNIFORM, PassEntryU, PassEntryC>::type;
    const Entry* entries = reinterpret_cast<const Entry*>(k.entries);
    auto load_entry = [&](int64_t run) -> Entry {
      Entry e;
      memset(&e, 0, sizeof(e));
      if (run < run_end && lane < RL) {
        if (UNIFORM) {
          const uint2 t = __ldg(reinterpret_cast<const uint2*>(entries) + run * RL + lane);
          memcpy(&e, &t, sizeof(t));
        } else {
          const uint4 t = __ldg(reinterpret_cast<const uint4*>(entries) + run * RL + lane);
          memcpy(&e, &t, sizeof(t));
        }
      }
      return e;
    };
    int64_t run = run_begin + warp;
    if (run < run_end) {
      Entry ent_cur = load_entry(run);
      Entry ent_next = load_entry(run + NW);
      Slot slot[kPrefetch];
      auto fetch = [&](Slot& s, const Entry& e, int idx) {
        const uint32_t st = __shfl_sync(0xffffffffu, e.start, idx);
        const uint32_t em = __shfl_sync(0xffffffffu, e.endmask, idx);
        uint32_t cb = 0;
        if constexpr (!UNIFORM) cb = __shfl_sync(0xffffffffu, e.colbase, idx);
        load_slot<ROW16, UNIFORM>(s, k, st, em, cb, lane);
      };
#pragma unroll
      for (int u = 0; u < kPrefetch; ++u) fetch(slot[u], ent_cur, u);
      while (true) {
        for (int i = 0; i < RL; i += kPrefetch) {
          const bool from_next = (i + kPrefetch >= RL);
#pragma unroll
          for (int u = 0; u < kPrefetch; ++u) {
            const Slot cur = slot[u];
            const int nidx = from_next ? (i + kPrefetch - RL + u) : (i + kPrefetch + u);
            if (from_next)
              fetch(slot[u], ent_next, nidx);
            else
              fetch(slot[u], ent_cur, nidx);
            process_pass<UNIFORM, SMODE>(cur, k, s_lam, s_grad, s_cls, lane, cx, xx);
          }
        }
        run += NW;
        if (run >= run_end) break;
        ent_cur = ent_next;
        ent_next = load_entry(run + NW);
      }
    }
  }

  // ---- flush per-CTA partial sums ----
  cx = block_sum(cx, dscratch);
  xx = block_sum(xx, dscratch);
  if (tid == 0) {
    if (cx != 0.0) atomicAdd(&k.acc_scal[0], cx);
    if (xx != 0.0) atomicAdd(&k.acc_scal[1], xx);
  }
  __syncthreads();
  if (SMODE <= 1) {
    const uint32_t bulk_bytes = (uint32_t)(m & ~3) * 4u;
    if (k.flush_bulk && bulk_bytes >= 16 && ((reinterpret_cast<uintptr_t>(k.acc) & 15u) == 0)) {
      // one TMA bulk reduction: acc[0..m) += s_grad[0..m) performed at L2 (SASS UBLKRED)
      fence_proxy_async_smem();
      __syncthreads();
      if (tid == 0) {
        bulk_reduce_add_f32_s2g(k.acc, s_grad, bulk_bytes);
        bulk_commit();
        bulk_wait_all();
      }
      for (int i = (m & ~3) + tid; i < m; i += THREADS) {
        const float g = s_grad[i];
        if (g != 0.f) atomicAdd(&k.acc[i], g);
      }
    } else {
      // staggered start so that CTAs do not walk the same addresses in lock step
      const int off = (int)(((int64_t)blockIdx.x * m) / gridDim.x);
      for (int i = tid; i < m; i += THREADS) {
        int j = i + off;
        if (j >= m) j -= m;
        const float g = s_grad[j];
        if (g != 0.f) atomicAdd(&k.acc[j], g);
      }
    }
  }
  // ---- last CTA to finish runs the m-length tail and leaves the accumulators zeroed ----
  __threadfence();
  __syncthreads();
  if (tid == 0) s_ticket = atomicAdd(k.counter, 1u);
  __syncthreads();
  if (s_ticket != gridDim.x - 1) return;
  __threadfence();
  const double cxv = __ldcg(&k.acc_scal[0]);
  const double xxv = __ldcg(&k.acc_scal[1]);
  if (k.do_epilogue) {
    cta_epilogue(k.acc, cxv, xxv, k.lambda, k.b, m, k.gamma, k.grad_out, k.scalars_out, dscratch, fscratch);
  } else {
    for (int i = tid; i < m; i += THREADS) k.partial_out[i] = __ldcg(k.acc + i);
    if (tid == 0) {
      k.partial_out[m] = (float)cxv;
      k.partial_out[m + 1] = (float)xxv;
    }
  }
  __syncthreads();
  for (int i = tid; i < m; i += THREADS) k.acc[i] = 0.f;
  if (tid == 0) {
    k.acc_scal[0] = 0.0;
    k.acc_scal[1] = 0.0;
    *k.counter = 0u;
  }
}

// ------------------------------------------------------------------------------------------
// Long columns (> 32 entries): one warp per column, multi-sweep, straight from global memory.
// The threshold is found with Michelot's fixed point (same support set and theta formula as Duchi's
// sorted scan in exact arithmetic); sums over the support are fp64 like the reference's CPU cumsum.
// ------------------------------------------------------------------------------------------
template <bool ROW16>
__global__ void __launch_bounds__(256) matching_long_kernel(const KArgs k, const LongCol* __restrict__ cols, int64_t n_long) {
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  double cx = 0.0, xx = 0.0;
  for (int64_t ci = warp_global; ci < n_long; ci += n_warps) {
    const LongCol lc = cols[ci];
    const dualip_proj_class pc = k.classes[lc.cls];
    const float* a = k.a + lc.start;
    const float* c = k.c + lc.start;
    auto row_at = [&](int e) -> uint32_t {
      if (ROW16) return __ldg(reinterpret_cast<const unsigned short*>(k.row) + lc.start + e);
      return __ldg(reinterpret_cast<const uint32_t*>(k.row) + lc.start + e);
    };
    auto v_at = [&](int e, float& av, float& cv, uint32_t& rv) -> float {
      av = __ldg(a + e);
      cv = __ldg(c + e);
      rv = row_at(e);
      const float lam_s = __fmul_rn(k.s, __ldg(k.lambda + rv));
      return __fadd_rn(__fmul_rn(av, lam_s), __fmul_rn(k.s, cv));
    };
    const bool is_sx = pc.kind != DUALIP_PROJ_CLAMP;
    float theta = 0.f;
    int branch = 0, rho = 0, amax = -1;
    if (is_sx) {
      // sweep 1: sum, top-2 of u/z, position of the maximum
      double S = 0.0;
      float m1 = 0.f, m2 = 0.f;
      int am = 0x7fffffff;
      for (int e = lane; e < lc.len; e += 32) {
        float av, cv;
        uint32_t rv;
        const float u = fmaxf(v_at(e, av, cv, rv), 0.f);
        const float un = __fdiv_rn(u, pc.z);
        S += (double)u;
        if (un > m1) {
          m2 = m1;
          m1 = un;
          am = e;
        } else if (un > m2) {
          m2 = un;
        }
      }
      S = warp_sum(S);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float o1 = __shfl_xor_sync(FULL, m1, o), o2 = __shfl_xor_sync(FULL, m2, o);
        const int oa = __shfl_xor_sync(FULL, am, o);
        const float mn = fminf(m1, o1);
        if (o1 > m1 || (o1 == m1 && oa < am)) am = oa;
        m1 = fmaxf(m1, o1);
        m2 = fmaxf(fmaxf(m2, o2), mn);
      }
      const bool feasible = (pc.kind == DUALIP_PROJ_SIMPLEX) && ((float)S <= pc.z_thr);
      const bool shortcut = !feasible && (__fsub_rn(m1, m2) > 1.0f);
      if (feasible) {
        branch = 0;
      } else if (shortcut) {
        branch = 1;
        rho = 1;
        amax = am;
      } else {
        branch = 2;
        // Michelot: t <- (sum_{u>t} u - z)/|{u>t}| until the support stops shrinking.
        double t = (S - (double)pc.z) / (double)lc.len;
        int cnt_prev = lc.len;
        double Ssup = S;
        for (int it = 0; it < 64; ++it) {
          double s2 = 0.0;
          int n2 = 0;
          for (int e = lane; e < lc.len; e += 32) {
            float av, cv;
            uint32_t rv;
            const float u = fmaxf(v_at(e, av, cv, rv), 0.f);
            if ((double)u > t) {
              s2 += (double)u;
              ++n2;
            }
          }
          s2 = warp_sum(s2);
          n2 = __reduce_add_sync(FULL, n2);
          if (n2 == 0) break;
          Ssup = s2;
          const bool done = (n2 == cnt_prev);
          cnt_prev = n2;
          t = (s2 - (double)pc.z) / (double)n2;
          if (done) break;
        }
        rho = cnt_prev;
        theta = __fdiv_rn(__fsub_rn((float)Ssup, pc.z), (float)rho);
      }
    }
    // final sweep: x, gradient scatter, scalars
    for (int e = lane; e < lc.len; e += 32) {
      float av, cv;
      uint32_t rv;
      const float v = v_at(e, av, cv, rv);
      float x;
      if (!is_sx) {
        x = fminf(fmaxf(v, pc.lo), pc.hi);
      } else {
        const float u = fmaxf(v, 0.f);
        x = branch == 0 ? u : (branch == 1 ? (e == amax ? pc.z : 0.f) : fmaxf(__fsub_rn(u, theta), 0.f));
      }
      const float g = __fmul_rn(av, x);
      if (g != 0.f) atomicAdd(&k.acc[rv], g);
      const double xd = (double)x;
      cx = fma((double)cv, xd, cx);
      xx = fma(xd, xd, xx);
      if (k.x_out) k.x_out[lc.start + e] = x;
      if (k.diag && is_sx && e == 0) k.diag[lc.start] = (uint8_t)(branch | (min(rho, 63) << 2));
    }
  }
  cx = warp_sum(cx);
  xx = warp_sum(xx);
  if (lane == 0) {
    if (cx != 0.0) atomicAdd(&k.acc_scal[0], cx);
    if (xx != 0.0) atomicAdd(&k.acc_scal[1], xx);
  }
}

__global__ void __launch_bounds__(1024) epilogue_kernel(const float* sum, int m, const float* lambda, const float* b,
                                                        double gamma, float* grad_out, dualip_scalars* out) {
  __shared__ double dscratch[32];
  __shared__ float fscratch[32];
  cta_epilogue(sum, (double)sum[m], (double)sum[m + 1], lambda, b, m, gamma, grad_out, out, dscratch, fscratch);
}

// ------------------------------------------------------------------------------------------
// Launch plumbing
// ------------------------------------------------------------------------------------------
typedef void (*PassKernel)(const KArgs);

template <int THREADS, int MINB>
static PassKernel pick_kernel(bool row16, bool uniform, int smode) {
#define DUALIP_PICK(R, U, S) \
  if (row16 == R && uniform == U && smode == S) return matching_pass_kernel<R, U, S, THREADS, MINB>;
  DUALIP_PICK(true, true, 0)
  DUALIP_PICK(true, true, 1)
  DUALIP_PICK(true, true, 2)
  DUALIP_PICK(true, false, 0)
  DUALIP_PICK(true, false, 1)
  DUALIP_PICK(true, false, 2)
  DUALIP_PICK(false, true, 0)
  DUALIP_PICK(false, true, 1)
  DUALIP_PICK(false, true, 2)
  DUALIP_PICK(false, false, 0)
  DUALIP_PICK(false, false, 1)
  DUALIP_PICK(false, false, 2)
#undef DUALIP_PICK
  return nullptr;
}

static PassKernel plan_kernel(const dualip_plan* p) {
  const bool row16 = p->row_bits == 16;
  if (p->threads == 512) return pick_kernel<512, 2>(row16, p->uniform, p->smode);
  return pick_kernel<1024, 1>(row16, p->uniform, p->smode);
}

static size_t smem_fixed_bytes(int n_classes) {
  return 16 + (((size_t)n_classes * sizeof(dualip_proj_class) + 15) & ~(size_t)15) + 32 * sizeof(double) + 32 * sizeof(float);
}

static int launch_eval(dualip_plan* p, const float* lambda, const float* b, double gamma, float* grad_out,
                       dualip_scalars* scalars_out, float* partial_out, float* x_out, uint8_t* diag, int do_epilogue,
                       cudaStream_t stream) {
  if (!(gamma > 0.0) && !(gamma < 0.0)) {
    set_error("gamma must be non-zero");
    return DUALIP_EINVAL;
  }
  KArgs k;
  k.a = p->a;
  k.c = p->c;
  k.row = p->row;
  k.entries = p->entries;
  k.cls_ne = p->cls_ne;
  k.classes = p->classes_dev;
  k.n_classes = p->n_classes;
  k.lambda = lambda;
  k.b = b;
  k.acc = p->acc;
  k.acc_scal = p->acc_scal;
  k.counter = p->counter;
  k.grad_out = grad_out;
  k.scalars_out = scalars_out;
  k.partial_out = partial_out;
  k.x_out = x_out;
  k.diag = diag;
  k.n_runs = p->n_runs;
  k.run_len = p->run_len;
  k.m = p->m;
  k.gamma = gamma;
  k.s = (float)(-1.0 / gamma);
  k.flush_bulk = p->flush_bulk;
  k.do_epilogue = do_epilogue;
  if (p->n_long > 0) {
    const int64_t warps = p->n_long;
    int blocks = (int)std::min<int64_t>((warps + 7) / 8, (int64_t)p->n_sms * 8);
    if (p->row_bits == 16)
      matching_long_kernel<true><<<blocks, 256, 0, stream>>>(k, p->longcols, p->n_long);
    else
      matching_long_kernel<false><<<blocks, 256, 0, stream>>>(k, p->longcols, p->n_long);
  }
  PassKernel kern = plan_kernel(p);
  kern<<<p->n_ctas, p->threads, p->smem_bytes, stream>>>(k);
  DUALIP_CUDA_TRY(cudaGetLastError());
  return DUALIP_OK;
}

template <typename IdxT>
static int build_passes(dualip_plan* p, const dualip_csc_desc* d, cudaStream_t stream) {
  const int64_t n_chunks = (p->n_cols + kChunkCols - 1) / kChunkCols;
  unsigned long long *counts = nullptr, *offsets = nullptr;
  unsigned int* bad = nullptr;
  DUALIP_CUDA_TRY(cudaMalloc(&counts, sizeof(unsigned long long) * 3 * (n_chunks + 1)));
  DUALIP_CUDA_TRY(cudaMalloc(&offsets, sizeof(unsigned long long) * 3 * (n_chunks + 1)));
  DUALIP_CUDA_TRY(cudaMalloc(&bad, sizeof(unsigned int)));
  DUALIP_CUDA_TRY(cudaMemsetAsync(bad, 0, sizeof(unsigned int), stream));
  DUALIP_CUDA_TRY(cudaMemsetAsync(counts, 0, sizeof(unsigned long long) * 3 * (n_chunks + 1), stream));
  const IdxT* ccol = reinterpret_cast<const IdxT*>(d->ccol_dev);
  const int tb = 128;
  const int nb = (int)((n_chunks + tb - 1) / tb);
  if (n_chunks > 0) {
    if (p->uniform)
      build_passes_kernel<IdxT, false, true><<<nb, tb, 0, stream>>>(ccol, d->col_class_dev, p->n_cols, n_chunks, counts,
                                                                    nullptr, nullptr, nullptr, nullptr, bad);
    else
      build_passes_kernel<IdxT, false, false><<<nb, tb, 0, stream>>>(ccol, d->col_class_dev, p->n_cols, n_chunks, counts,
                                                                     nullptr, nullptr, nullptr, nullptr, bad);
  }
  // exclusive scan over the interleaved triples: scan each of the three streams with stride-3 via a
  // single scan of 3*(n_chunks+1) values is wrong, so scan three strided views with a custom iterator.
  // Simpler: copy to host when small, else three cub scans over de-interleaved temporaries.
  const int64_t N = n_chunks + 1;
  unsigned long long *tmp_in = nullptr, *tmp_out = nullptr;
  DUALIP_CUDA_TRY(cudaMalloc(&tmp_in, sizeof(unsigned long long) * N));
  DUALIP_CUDA_TRY(cudaMalloc(&tmp_out, sizeof(unsigned long long) * N));
  void* cub_tmp = nullptr;
  size_t cub_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, tmp_in, tmp_out, (int)N, stream);
  DUALIP_CUDA_TRY(cudaMalloc(&cub_tmp, cub_bytes ? cub_bytes : 16));
  unsigned long long totals[3] = {0, 0, 0};
  for (int f = 0; f < 3; ++f) {
    DUALIP_CUDA_TRY(cudaMemcpy2DAsync(tmp_in, sizeof(unsigned long long), counts + f, 3 * sizeof(unsigned long long),
                                      sizeof(unsigned long long), N, cudaMemcpyDeviceToDevice, stream));
    cub::DeviceScan::ExclusiveSum(cub_tmp, cub_bytes, tmp_in, tmp_out, (int)N, stream);
    DUALIP_CUDA_TRY(cudaMemcpy2DAsync(offsets + f, 3 * sizeof(unsigned long long), tmp_out, sizeof(unsigned long long),
                                      sizeof(unsigned long long), N, cudaMemcpyDeviceToDevice, stream));
    DUALIP_CUDA_TRY(cudaMemcpyAsync(&totals[f], tmp_out + n_chunks, sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
  }
  DUALIP_CUDA_TRY(cudaStreamSynchronize(stream));
  p->n_passes = (int64_t)totals[0];
  const int64_t n_ne = (int64_t)totals[1];
  p->n_long = (int64_t)totals[2];

  // run length: shorter runs for small problems so that every warp gets work
  int rl = 32;
  const int64_t warps_total = (int64_t)p->n_ctas * (p->threads / 32);
  while (rl > kPrefetch && p->n_passes / rl < warps_total * 4) rl >>= 1;
  p->run_len = rl;
  p->n_runs = (p->n_passes + rl - 1) / rl;
  const int64_t padded = (p->n_runs + 2) * rl + 32;
  const size_t esz = p->uniform ? sizeof(PassEntryU) : sizeof(PassEntryC);
  DUALIP_CUDA_TRY(cudaMalloc(&p->entries, esz * padded));
  DUALIP_CUDA_TRY(cudaMemsetAsync(p->entries, 0, esz * padded, stream));
  p->owned_bytes += esz * padded;
  if (!p->uniform) {
    DUALIP_CUDA_TRY(cudaMalloc(&p->cls_ne, (size_t)n_ne + 16));
    p->owned_bytes += (size_t)n_ne + 16;
  }
  if (p->n_long > 0) {
    DUALIP_CUDA_TRY(cudaMalloc(&p->longcols, sizeof(LongCol) * p->n_long));
    p->owned_bytes += sizeof(LongCol) * p->n_long;
  }
  if (n_chunks > 0) {
    if (p->uniform)
      build_passes_kernel<IdxT, true, true><<<nb, tb, 0, stream>>>(ccol, d->col_class_dev, p->n_cols, n_chunks, nullptr,
                                                                   offsets, p->entries, p->cls_ne, p->longcols, bad);
    else
      build_passes_kernel<IdxT, true, false><<<nb, tb, 0, stream>>>(ccol, d->col_class_dev, p->n_cols, n_chunks, nullptr,
                                                                    offsets, p->entries, p->cls_ne, p->longcols, bad);
  }
  unsigned int bad_h = 0;
  DUALIP_CUDA_TRY(cudaMemcpyAsync(&bad_h, bad, sizeof(bad_h), cudaMemcpyDeviceToHost, stream));
  DUALIP_CUDA_TRY(cudaStreamSynchronize(stream));
  cudaFree(counts);
  cudaFree(offsets);
  cudaFree(tmp_in);
  cudaFree(tmp_out);
  cudaFree(cub_tmp);
  cudaFree(bad);
  if (bad_h) {
    set_error("ccol_indices must be non-decreasing (flags=%u)", bad_h);
    return DUALIP_EINVAL;
  }
  return DUALIP_OK;
}

}  // namespace dualip

// ------------------------------------------------------------------------------------------
// C-ABI
// ------------------------------------------------------------------------------------------
extern "C" {

int dualip_abi_version(void) { return DUALIP_B200_ABI_VERSION; }
const char* dualip_last_error(void) { return g_last_error.c_str(); }

void dualip_plan_destroy(dualip_plan* p) {
  if (!p) return;
  DeviceGuard g(p->device);
  cudaFree(p->row);
  cudaFree(p->entries);
  cudaFree(p->cls_ne);
  cudaFree(p->longcols);
  cudaFree(p->classes_dev);
  cudaFree(p->acc);
  cudaFree(p->acc_scal);
  cudaFree(p->counter);
  cudaFree(p->lambda_stage);
  cudaFree(p->grad_stage);
  cudaFree(p->scal_stage);
  delete p;
}

int dualip_plan_create(dualip_plan** out, const dualip_csc_desc* d) {
  if (!out || !d) {
    set_error("null argument");
    return DUALIP_EINVAL;
  }
  *out = nullptr;
  if (d->n_cols < 0 || d->nnz < 0 || d->n_rows <= 0) {
    set_error("bad shape: n_cols=%lld nnz=%lld n_rows=%d", (long long)d->n_cols, (long long)d->nnz, d->n_rows);
    return DUALIP_EINVAL;
  }
  if (d->index_bits != 32 && d->index_bits != 64) {
    set_error("index_bits must be 32 or 64");
    return DUALIP_EINVAL;
  }
  if (d->n_classes < 1 || d->n_classes > kMaxClasses || !d->classes) {
    set_error("n_classes must be in 1..%d", kMaxClasses);
    return DUALIP_EINVAL;
  }
  if (d->nnz >= (1LL << 32)) {
    set_error("nnz >= 2^32 per shard is not supported by this build; shard the columns");
    return DUALIP_ERANGE;
  }
  if (d->nnz > 0 && (!d->ccol_dev || !d->row_dev || !d->a_dev || !d->c_dev)) {
    set_error("null CSC array");
    return DUALIP_EINVAL;
  }
  for (int i = 0; i < d->n_classes; ++i) {
    const dualip_proj_class& pc = d->classes[i];
    if (pc.kind < DUALIP_PROJ_CLAMP || pc.kind > DUALIP_PROJ_SIMPLEX_EQ) {
      set_error("class %d: unknown projection kind %d", i, pc.kind);
      return DUALIP_EINVAL;
    }
    if (pc.kind != DUALIP_PROJ_CLAMP && !(pc.z > 0.f)) {
      set_error("class %d: simplex radius z must be positive", i);  // simplex.py:145
      return DUALIP_EINVAL;
    }
  }
  DeviceGuard g(d->device);
  if (!g.ok) {
    set_error("cannot select CUDA device %d", d->device);
    return DUALIP_ECUDA;
  }
  dualip_plan* p = new (std::nothrow) dualip_plan();
  if (!p) return DUALIP_ENOMEM;
  p->device = d->device;
  p->n_cols = d->n_cols;
  p->nnz = d->nnz;
  p->m = d->n_rows;
  p->a = d->a_dev;
  p->c = d->c_dev;
  p->uniform = (d->col_class_dev == nullptr);
  p->n_classes = d->n_classes;
  memcpy(p->classes_host, d->classes, sizeof(dualip_proj_class) * d->n_classes);
  const char* fb = getenv("DUALIP_FLUSH");
  p->flush_bulk = (fb && strcmp(fb, "atomic") == 0) ? 0 : 1;

  auto fail = [&](int rc) {
    dualip_plan_destroy(p);
    return rc;
  };
  cudaStream_t stream = 0;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, d->device) != cudaSuccess) {
    set_error("cudaGetDeviceProperties failed");
    return fail(DUALIP_ECUDA);
  }
  p->n_sms = prop.multiProcessorCount;

  // shared-memory mode and CTA shape
  const size_t fixed = smem_fixed_bytes(p->n_classes);
  const size_t m_pad = ((size_t)p->m + 3) & ~(size_t)3;
  const size_t need0 = fixed + 4 * m_pad + 4 * (size_t)p->m;
  const size_t need1 = fixed + 4 * (size_t)p->m;
  const size_t smem_max = std::min<size_t>(kSmemBudget, prop.sharedMemPerBlockOptin);
  if (need0 <= smem_max) {
    p->smode = 0;
    p->smem_bytes = need0;
  } else if (need1 <= smem_max) {
    p->smode = 1;
    p->smem_bytes = need1;
  } else {
    p->smode = 2;
    p->smem_bytes = fixed;
  }
  // two 512-thread CTAs per SM when both fit (each SM has 228 KB; 1 KB per CTA is reserved)
  if (2 * (p->smem_bytes + 1024) <= prop.sharedMemPerMultiprocessor) {
    p->threads = 512;
    p->n_ctas = 2 * p->n_sms;
  } else {
    p->threads = 1024;
    p->n_ctas = p->n_sms;
  }
  const char* env_ctas = getenv("DUALIP_CTAS");
  if (env_ctas && atoi(env_ctas) > 0) p->n_ctas = atoi(env_ctas);

#define DUALIP_TRY_FAIL(expr)                                                                    \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess) {                                                                     \
      set_error("%s failed: %s", #expr, cudaGetErrorString(_e));                                 \
      return fail(_e == cudaErrorMemoryAllocation ? DUALIP_ENOMEM : DUALIP_ECUDA);               \
    }                                                                                            \
  } while (0)

  // narrowed row indices (owned)
  p->row_bits = (p->m <= 65536) ? 16 : 32;
  {
    const size_t rb = (size_t)(p->row_bits / 8) * (size_t)std::max<int64_t>(p->nnz, 1) + 64;
    DUALIP_TRY_FAIL(cudaMalloc(&p->row, rb));
    p->owned_bytes += rb;
    unsigned int* bad = nullptr;
    DUALIP_TRY_FAIL(cudaMalloc(&bad, sizeof(unsigned int)));
    DUALIP_TRY_FAIL(cudaMemsetAsync(bad, 0, sizeof(unsigned int), stream));
    if (p->nnz > 0) {
      const int tb = 256;
      const int nb = (int)std::min<int64_t>((p->nnz + tb - 1) / tb, (int64_t)p->n_sms * 16);
      if (d->index_bits == 64) {
        if (p->row_bits == 16)
          narrow_rows_kernel<long long, unsigned short><<<nb, tb, 0, stream>>>((const long long*)d->row_dev, (unsigned short*)p->row, p->nnz, p->m, bad);
        else
          narrow_rows_kernel<long long, uint32_t><<<nb, tb, 0, stream>>>((const long long*)d->row_dev, (uint32_t*)p->row, p->nnz, p->m, bad);
      } else {
        if (p->row_bits == 16)
          narrow_rows_kernel<int, unsigned short><<<nb, tb, 0, stream>>>((const int*)d->row_dev, (unsigned short*)p->row, p->nnz, p->m, bad);
        else
          narrow_rows_kernel<int, uint32_t><<<nb, tb, 0, stream>>>((const int*)d->row_dev, (uint32_t*)p->row, p->nnz, p->m, bad);
      }
    }
    unsigned int bad_h = 0;
    DUALIP_TRY_FAIL(cudaMemcpyAsync(&bad_h, bad, sizeof(bad_h), cudaMemcpyDeviceToHost, stream));
    DUALIP_TRY_FAIL(cudaStreamSynchronize(stream));
    cudaFree(bad);
    if (bad_h) {
      set_error("row index out of range [0,%d)", p->m);
      return fail(DUALIP_EINVAL);
    }
  }
  // pass table
  {
    int rc = (d->index_bits == 64) ? build_passes<long long>(p, d, stream) : build_passes<int>(p, d, stream);
    if (rc != DUALIP_OK) return fail(rc);
  }
  // small state
  DUALIP_TRY_FAIL(cudaMalloc(&p->classes_dev, sizeof(dualip_proj_class) * p->n_classes));
  DUALIP_TRY_FAIL(cudaMemcpy(p->classes_dev, p->classes_host, sizeof(dualip_proj_class) * p->n_classes, cudaMemcpyHostToDevice));
  DUALIP_TRY_FAIL(cudaMalloc(&p->acc, sizeof(float) * (m_pad + 4)));
  DUALIP_TRY_FAIL(cudaMemset(p->acc, 0, sizeof(float) * (m_pad + 4)));
  DUALIP_TRY_FAIL(cudaMalloc(&p->acc_scal, sizeof(double) * 2));
  DUALIP_TRY_FAIL(cudaMemset(p->acc_scal, 0, sizeof(double) * 2));
  DUALIP_TRY_FAIL(cudaMalloc(&p->counter, sizeof(unsigned int)));
  DUALIP_TRY_FAIL(cudaMemset(p->counter, 0, sizeof(unsigned int)));
  DUALIP_TRY_FAIL(cudaMalloc(&p->lambda_stage, sizeof(float) * (m_pad + 4)));
  DUALIP_TRY_FAIL(cudaMalloc(&p->grad_stage, sizeof(float) * (m_pad + 4)));
  DUALIP_TRY_FAIL(cudaMalloc(&p->scal_stage, sizeof(dualip_scalars)));
  p->owned_bytes += sizeof(float) * 3 * (m_pad + 4) + 64;

  PassKernel kern = plan_kernel(p);
  if (!kern) {
    set_error("no kernel variant");
    return fail(DUALIP_EINVAL);
  }
  DUALIP_TRY_FAIL(cudaFuncSetAttribute((const void*)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem_bytes));
  DUALIP_TRY_FAIL(cudaDeviceSynchronize());
#undef DUALIP_TRY_FAIL
  *out = p;
  return DUALIP_OK;
}

int dualip_plan_info(const dualip_plan* p, int64_t* out, int cap) {
  if (!p || !out) {
    set_error("null argument");
    return DUALIP_EINVAL;
  }
  const int64_t v[10] = {p->n_passes, p->n_long, p->n_ctas, p->threads, (int64_t)p->smem_bytes, p->row_bits,
                         p->smode,    p->run_len, (p->n_long > 0) ? 2 : 1, (int64_t)p->owned_bytes};
  for (int i = 0; i < cap && i < 10; ++i) out[i] = v[i];
  return DUALIP_OK;
}

int dualip_matching_calc(dualip_plan* p, const float* lambda_dev, const float* b_dev, double gamma, float* grad_out_dev,
                         dualip_scalars* scalars_out_dev, float* x_out_dev, uint8_t* diag_out_dev, uint32_t flags,
                         void* stream) {
  (void)flags;
  if (!p || !lambda_dev || !grad_out_dev || !scalars_out_dev) {
    set_error("null argument");
    return DUALIP_EINVAL;
  }
  DeviceGuard g(p->device);
  return launch_eval(p, lambda_dev, b_dev, gamma, grad_out_dev, scalars_out_dev, nullptr, x_out_dev, diag_out_dev, 1,
                     (cudaStream_t)stream);
}

int dualip_matching_partial(dualip_plan* p, const float* lambda_dev, double gamma, float* partial_out_dev,
                            float* x_out_dev, uint8_t* diag_out_dev, uint32_t flags, void* stream) {
  (void)flags;
  if (!p || !lambda_dev || !partial_out_dev) {
    set_error("null argument");
    return DUALIP_EINVAL;
  }
  DeviceGuard g(p->device);
  return launch_eval(p, lambda_dev, nullptr, gamma, nullptr, nullptr, partial_out_dev, x_out_dev, diag_out_dev, 0,
                     (cudaStream_t)stream);
}

int dualip_matching_epilogue(const float* partial_sum_dev, int32_t m, const float* lambda_dev, const float* b_dev,
                             double gamma, float* grad_out_dev, dualip_scalars* scalars_out_dev, void* stream) {
  if (!partial_sum_dev || !lambda_dev || !grad_out_dev || !scalars_out_dev || m <= 0) {
    set_error("null argument");
    return DUALIP_EINVAL;
  }
  epilogue_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(partial_sum_dev, m, lambda_dev, b_dev, gamma, grad_out_dev,
                                                        scalars_out_dev);
  DUALIP_CUDA_TRY(cudaGetLastError());
  return DUALIP_OK;
}

int dualip_matching_calc_host(dualip_plan* p, const float* lambda_host, const float* b_dev, double gamma,
                              float* grad_out_host, dualip_scalars* scalars_out_host, void* stream) {
  if (!p || !lambda_host || !grad_out_host || !scalars_out_host) {
    set_error("null argument");
    return DUALIP_EINVAL;
  }
  DeviceGuard g(p->device);
  cudaStream_t st = (cudaStream_t)stream;
  DUALIP_CUDA_TRY(cudaMemcpyAsync(p->lambda_stage, lambda_host, sizeof(float) * p->m, cudaMemcpyHostToDevice, st));
  int rc = launch_eval(p, p->lambda_stage, b_dev, gamma, p->grad_stage, p->scal_stage, nullptr, nullptr, nullptr, 1, st);
  if (rc != DUALIP_OK) return rc;
  DUALIP_CUDA_TRY(cudaMemcpyAsync(grad_out_host, p->grad_stage, sizeof(float) * p->m, cudaMemcpyDeviceToHost, st));
  DUALIP_CUDA_TRY(cudaMemcpyAsync(scalars_out_host, p->scal_stage, sizeof(dualip_scalars), cudaMemcpyDeviceToHost, st));
  DUALIP_CUDA_TRY(cudaStreamSynchronize(st));
  return DUALIP_OK;
}

}  // extern "C"

// Fused dual-ascent evaluation for block-structured matching LPs on sm_100a.
//
// One launch of matching_slab_kernel replaces the ~15 eager ATen passes of the reference's
// MatchingSolverDualObjectiveFunction.calculate (reference src/dualip/objectives/matching.py:116-188):
//   v = a*(-lambda[row]/gamma) + (-c/gamma)        matching.py:136-142, utils/sparse_utils.py:54-85,26-51
//   x = Proj_column(v)                              sparse_utils.py:133-220 -> projections/{box,cone,simplex}.py
//   grad[row] += a*x ; cx += c*x ; xx += x*x        matching.py:153-160, sparse_utils.py:223-243
//   grad -= b ; dual_obj = cx + gamma/2*xx + lambda.grad ; slacks      matching.py:25-34,164-178
//
// HBM layout ("slabs", a sliced-ELL format sorted by projection class and column length): at plan time the
// columns are stably sorted by (class, nnz) and cut into slabs of 32 columns of EQUAL length d.  A slab stores
// its 32*d values in CHUNKS of four entries per column: entry k of column `lane` sits at
//     (k & ~3)*32 + lane*4 + (k & 3)                      for k < (d & ~3)        (one 16-byte vector per lane and chunk)
// and the d & 3 trailing entries form a 2-entry chunk (lane*2 + j) and/or a 1-entry chunk (lane); see slab_elem().
// The three arrays of a slab are stored back to back -- [a: 32 d floats][c: 32 d floats][row: 32 d ids], 320 d bytes with
// uint16 ids -- so that ONE bulk copy (TMA engine) moves a whole slab into shared memory; slabs follow each other in
// (class, length) order, and a small group table {first slab, first row, d, class} replaces per-slab headers in the hot loop.
// A warp owns a slab and a lane owns a column: one LDG.128 per lane fetches four entries of a (or c), one LDG.64
// four uint16 row ids, every request is fully coalesced (512 contiguous bytes per warp instruction), all per-column
// reductions are register arithmetic without shuffles, and lanes never diverge on column length or projection
// type.  Columns of up to kRegDeg entries are processed entirely in registers by code specialised on d (one load
// phase, no second pass); longer ones stream through the generic path.  The only shared-memory traffic per nonzero is
// the lambda gather and the gradient scatter.  Columns longer than kMaxThreadDeg go to a warp-per-column kernel.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_run_length_encode.cuh>
#include <algorithm>
#include <new>
#include <type_traits>
#include <vector>
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "agd_step.cuh"

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

constexpr int kSlabW = 32;            // columns per slab = lanes per warp
constexpr int kMaxThreadDeg = 1024;   // longest column handled one-lane-per-column
constexpr int kStashDeg = 16;          // simplex columns up to this length keep u in the warp's shared-memory stash
// Sort key of a column: length-major, key = length << 8 | class.  Groups of equal length and different class are
// neighbours in the slab order, so a contiguous range of slabs holds the same mix of classes as the whole problem.
constexpr int kClsBits = 8;
__host__ __device__ constexpr uint32_t make_key(uint32_t cls, uint32_t d) { return (d << kClsBits) | cls; }
__host__ __device__ constexpr int key_d(uint32_t key) { return (int)(key >> kClsBits); }
__host__ __device__ constexpr int key_cls(uint32_t key) { return (int)(key & ((1u << kClsBits) - 1)); }
constexpr uint32_t kKeyEmpty = make_key(255u, 2047u);
constexpr uint32_t kKeyLong = make_key(255u, 2046u);
constexpr int kKeyBits = 19;
constexpr int kMaxClasses = 255;
constexpr size_t kSmemBudget = 227 * 1024;
constexpr int kStageSlots = 2;        // slots of a warp's staging region = its mbarriers
constexpr int kThreads = 512;
constexpr int kBarBytes = (kThreads / 32) * kStageSlots * 8;  // slot mbarriers of all warps         // one CTA of 16 warps per SM: up to 128 registers per thread for the register path

struct SlabHdr {      // 8 bytes per slab (per 32*d nonzeros)
  uint32_t off32;     // first row of the slab in units of 32 elements
  uint16_t d;         // entries per column
  uint8_t cls;        // projection class
  uint8_t ncols;      // active lanes (32 except in the last slab of a (class, d) group)
};

// Position of entry k of column `lane` inside a slab of column length d, in elements from the slab's start.
__host__ __device__ __forceinline__ uint32_t slab_elem(int k, int d, int lane) {
  const int full = d & ~3;
  if (k < full) return (uint32_t)((k & ~3) * 32 + lane * 4 + (k & 3));
  const int two = d & 2, kr = k - full;
  if (kr < two) return (uint32_t)(full * 32 + lane * 2 + kr);
  return (uint32_t)(full * 32 + two * 32 + lane);
}

// One (class, column length) group: n_slabs consecutive slabs of 32 columns of d entries each (the last one may hold
// fewer columns).  Slab j of the group starts at row32 = off32_begin + j * d, i.e. at byte (off32_begin + j*d) * row_bytes.
struct SlabGroup {       // 16 bytes
  uint32_t slab_begin;   // index of the group's first slab
  uint32_t off32_begin;  // its first row of 32 elements
  uint16_t d;            // entries per column
  uint8_t cls;           // projection class
  uint8_t last_ncols;    // columns in the group's last slab (1..32)
  uint32_t n_slabs;
};
// The part of a group that falls into one CTA's slab range; built in shared memory at kernel start.
struct Seg {             // 16 bytes
  int a, b;              // slabs [a, b)
  uint32_t off32_a;      // row of 32 elements where slab a starts
  uint16_t d;
  uint8_t cls;
  uint8_t last_ncols;    // columns in slab b-1 (32 unless it is the group's last slab)
};
constexpr int kMaxSeg = 128;  // segments per batch (a CTA's range rarely touches more than a handful of groups)

// bytes of one row of 32 entries: a + c + row id
__host__ __device__ constexpr int row32_bytes(int row_bits) { return 32 * (8 + row_bits / 8); }

struct LongCol {
  int64_t src_start;  // position of the column's first entry in the caller's CSC value order
  int64_t off;        // offset into the plan's compact long-column arrays
  int32_t len;
  int32_t cls;
};

}  // namespace dualip

using namespace dualip;

struct dualip_plan {
  int device = 0;
  int64_t n_cols = 0, nnz = 0;
  int32_t m = 0;
  int row_bits = 32;
  // slabs (owned)
  unsigned char* data = nullptr;  // slabs back to back, row32_bytes(row_bits) bytes per row of 32 entries
  SlabGroup* groups = nullptr;    // n_groups entries
  int n_groups = 0;
  std::vector<SlabGroup> groups_host;
  int2* cta_range = nullptr;      // n_ctas + 1 entries {first slab, first group} of every CTA's contiguous slab range
  std::vector<int2> ranges_host;
  unsigned int* cta_ns = nullptr; // per CTA: duration of its main loop in the last launch (ns), input of dualip_plan_rebalance
  int rebalances = 0;
  SlabHdr* hdr = nullptr;         // per-slab headers (plan-time kernels, tests); not read by the hot kernel
  int64_t* orig_start = nullptr;  // per slab lane: first nnz position of the column in the caller's order, or -1
  int64_t n_slabs = 0;
  int64_t rows32 = 0;             // total slab rows (32 elements each)
  int64_t n_short = 0;            // columns stored in slabs
  // long columns (owned, compact copies)
  LongCol* longcols = nullptr;    // sorted by length: [0, n_mid) mid columns (<= kMaxThreadDeg entries, warp per column inside the
                                  // slab kernel), [n_mid, n_long) long columns (matching_long_kernel)
  int64_t n_long = 0;
  int64_t n_mid = 0;
  int64_t n_ctalong = 0;          // of the long columns, those that fit the CTA-per-column kernel's stash: [n_mid, n_mid + n_ctalong)
  bool mid_separate = false;      // many mid columns: they take the team-per-column kernel (32 threads per column, 40 warps per SM)
                                  // in front of the slab kernel instead of the slab kernel's own warp-per-column path
  int* mid_range = nullptr;       // n_ctas + 1: every CTA's contiguous share of the mid columns (equal cost)
  int64_t long_total = 0;         // entries of all long columns
  float* long_a = nullptr;
  float* long_c = nullptr;
  uint32_t* long_row = nullptr;
  dualip_proj_class classes_host[kMaxClasses];
  dualip_proj_class* classes_dev = nullptr;
  int n_classes = 0;
  int* pad_dev = nullptr;      // n_classes x DUALIP_PAD_BUCKETS padded block lengths (simplex_eq), or null
  float* acc = nullptr;        // m floats, zero between calls
  int* acc_lo = nullptr;       // fixed-point mode: m ints each, zero between calls
  int* acc_hi = nullptr;
  bool class_used[256] = {};   // classes that own at least one non-empty column
  int fixed_point = 0;         // 1: 32-bit fixed-point gradient accumulation (native shared-memory integer adds)
  int fx_bits = 0;             // F
  float fx_scale = 1.f;        // 2^F
  double fx_inv = 1.0;         // 2^-F
  double fx_bound = 0.0;       // largest possible |row sum| inside one CTA
  double fx_relerr = 0.0;      // worst-row rounding error estimate relative to the row's largest possible sum
  bool fx_bounded = false;     // every class bounds x and the kernel variant supports fixed point (only the resolution can fail)
  std::vector<float> row_total_host;  // per row: sum |a| * xmax over all its entries (the row's largest possible sum)
  float* row_unscale = nullptr;  // m floats 2^-k_r, or null: rows are stored scaled by 2^k_r (power-of-two row equilibration)
  double* acc_scal = nullptr;  // [c.x, ||x||^2], zero between calls
  unsigned int* counter = nullptr;
  unsigned int* grid_bar = nullptr;  // grid barrier of the all-CTA tail: {arrive counter, generation}
  double* tail_part = nullptr;       // n_ctas x kTailPart doubles
  int* grid_status = nullptr;
  int* grid_status_host = nullptr;      // the same word in mapped host memory (written by the kernel on a time-out only)
  int* grid_status_host_dev = nullptr;  // its device-side address
  int grid_tail = 1;                 // all-CTA tail for this plan's own launches (size rule, or forced by DUALIP_GRID_TAIL=0|1)
  int grid_tail_auto = 1;            // 1: not forced; sharded launches may still turn it on from the world size
  int last_grid_tail = 0;            // what the most recent launch of the slab kernel used (introspection)
  float* lambda_stage = nullptr;  // m floats, for *_calc_host
  float* grad_stage = nullptr;
  dualip_scalars* scal_stage = nullptr;
  int n_sms = 0, n_ctas = 0, threads = 512, smode = 0;
  size_t smem_bytes = 0;
  size_t owned_bytes = 0;
  int flush_bulk = 1;
  int stage_region = 0;        // staging bytes per warp
  unsigned long long* timeline = nullptr;  // debug only
  int stage = 0;               // longest column whose slab is TMA-staged (0: staging off / does not fit)
};

namespace dualip {

// ------------------------------------------------------------------------------------------
// Plan construction (setup time)
// ------------------------------------------------------------------------------------------
// Histogram of the column lengths 0..kMaxThreadDeg (longer ones in the last bin): the plan chooses from it where the slab layout ends.
template <typename IdxT>
__global__ void length_hist_kernel(const IdxT* __restrict__ ccol, int64_t n_cols, unsigned int* __restrict__ hist) {
  __shared__ unsigned int s_h[kMaxThreadDeg + 2];
  for (int i = threadIdx.x; i < kMaxThreadDeg + 2; i += blockDim.x) s_h[i] = 0u;
  __syncthreads();
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; j < n_cols; j += (int64_t)gridDim.x * blockDim.x) {
    const int64_t d = (int64_t)ccol[j + 1] - (int64_t)ccol[j];
    atomicAdd(&s_h[d < 0 ? 0 : (d > kMaxThreadDeg ? kMaxThreadDeg + 1 : (int)d)], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kMaxThreadDeg + 2; i += blockDim.x)
    if (s_h[i]) atomicAdd(&hist[i], s_h[i]);
}

// slab_max_deg: columns longer than this leave the slab layout for the compact column-contiguous arrays (warp per column):
// kRegDeg for plans with the register path, kMaxThreadDeg otherwise.  cls_slab_only[c] != 0 keeps class c in slabs up to
// kMaxThreadDeg whatever the plan (bisection classes: only the slab kernel's generic path implements them).
template <typename IdxT>
__global__ void column_keys_kernel(const IdxT* __restrict__ ccol, const uint8_t* __restrict__ col_class, int64_t n_cols,
                                   int n_classes, int slab_max_deg, const uint8_t* __restrict__ cls_slab_only,
                                   uint32_t* __restrict__ keys, uint32_t* __restrict__ vals, unsigned int* bad) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; j < n_cols; j += stride) {
    const int64_t d = (int64_t)ccol[j + 1] - (int64_t)ccol[j];
    const uint32_t cls = col_class ? col_class[j] : 0u;
    if (d < 0) atomicOr(bad, 2u);
    if (d > 0x7fffffffLL) atomicOr(bad, 4u);
    if ((int)cls >= n_classes) atomicOr(bad, 8u);
    uint32_t key;
    if (d <= 0)
      key = kKeyEmpty;
    else if (d > kMaxThreadDeg || (d > slab_max_deg && !((int)cls < n_classes && cls_slab_only[cls])))
      key = kKeyLong;
    else
      key = make_key(cls, (uint32_t)d);
    keys[j] = key;
    vals[j] = (uint32_t)j;
  }
}

// One thread per short column (in sorted order): copies the column into its slab lane.
template <typename IdxT, typename RowT>
__global__ void fill_slabs_kernel(const IdxT* __restrict__ ccol, const IdxT* __restrict__ row, const float* __restrict__ a,
                                  const float* __restrict__ c, const uint32_t* __restrict__ perm, int64_t n_short,
                                  const int64_t* __restrict__ g_start, const int64_t* __restrict__ g_slab_base,
                                  const int64_t* __restrict__ g_off32, const uint32_t* __restrict__ g_key, int n_groups,
                                  unsigned char* __restrict__ data,
                                  SlabHdr* __restrict__ hdr, int64_t* __restrict__ orig_start, int32_t m,
                                  unsigned int* bad) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_short) return;
  int lo = 0, hi = n_groups - 1;  // largest g with g_start[g] <= i
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (g_start[mid] <= i)
      lo = mid;
    else
      hi = mid - 1;
  }
  const int g = lo;
  const uint32_t key = g_key[g];
  const int d = key_d(key);
  const int64_t rel = i - g_start[g];
  const int64_t slab = g_slab_base[g] + rel / kSlabW;
  const int lane = (int)(rel % kSlabW);
  const int64_t off32 = g_off32[g] + (rel / kSlabW) * d;
  const uint32_t col = perm[i];
  const int64_t e0 = (int64_t)ccol[col];
  unsigned char* slab_bytes = data + (size_t)off32 * (size_t)row32_bytes(8 * (int)sizeof(RowT));
  float* a_t = reinterpret_cast<float*>(slab_bytes);
  float* c_t = a_t + (size_t)d * kSlabW;
  RowT* row_t = reinterpret_cast<RowT*>(c_t + (size_t)d * kSlabW);
  for (int k = 0; k < d; ++k) {
    const uint32_t dst = slab_elem(k, d, lane);
    const IdxT r = row[e0 + k];
    if (r < 0 || r >= (IdxT)m) atomicOr(bad, 1u);
    a_t[dst] = a[e0 + k];
    c_t[dst] = c[e0 + k];
    row_t[dst] = (RowT)r;
  }
  orig_start[slab * kSlabW + lane] = e0;
  if (lane == 0) {
    const int64_t g_count = (g + 1 < n_groups ? g_start[g + 1] : n_short) - g_start[g];
    const int64_t left = g_count - (rel / kSlabW) * kSlabW;
    SlabHdr h;
    h.off32 = (uint32_t)off32;
    h.d = (uint16_t)d;
    h.cls = (uint8_t)key_cls(key);
    h.ncols = (uint8_t)(left < kSlabW ? left : kSlabW);
    hdr[slab] = h;
  }
}

template <typename IdxT>
__global__ void long_meta_kernel(const IdxT* __restrict__ ccol, const uint8_t* __restrict__ col_class,
                                 const uint32_t* __restrict__ perm_long, int64_t n_long, LongCol* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_long) return;
  const uint32_t col = perm_long[i];
  LongCol lc;
  lc.src_start = (int64_t)ccol[col];
  lc.off = 0;
  lc.len = (int32_t)((int64_t)ccol[col + 1] - (int64_t)ccol[col]);
  lc.cls = col_class ? (int32_t)col_class[col] : 0;
  out[i] = lc;
}

template <typename IdxT>
__global__ void long_copy_kernel(const LongCol* __restrict__ cols, int64_t n_long, const IdxT* __restrict__ row,
                                 const float* __restrict__ a, const float* __restrict__ c, float* __restrict__ la,
                                 float* __restrict__ lcv, uint32_t* __restrict__ lrow, int32_t m, unsigned int* bad) {
  const int lane = threadIdx.x & 31;
  int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (; w < n_long; w += nw) {
    const LongCol lc = cols[w];
    for (int e = lane; e < lc.len; e += 32) {
      const IdxT r = row[lc.src_start + e];
      if (r < 0 || r >= (IdxT)m) atomicOr(bad, 1u);
      la[lc.off + e] = a[lc.src_start + e];
      lcv[lc.off + e] = c[lc.src_start + e];
      lrow[lc.off + e] = (uint32_t)r;
    }
  }
}

// ------------------------------------------------------------------------------------------
// Hot kernel
// ------------------------------------------------------------------------------------------
struct KArgs {
  const unsigned char* data;   // slabs back to back
  const SlabGroup* groups;
  int n_groups;
  const int2* cta_range;       // gridDim.x + 1 entries {first slab, first group}
  unsigned int* cta_ns;        // per CTA: main-loop duration of this launch, ns
  const int64_t* orig_start;
  int64_t n_slabs;
  const dualip_proj_class* classes;
  int n_classes;
  const int* pad;            // n_classes x DUALIP_PAD_BUCKETS padded lengths of the reference's blocks, or null
  const float* lambda;
  const float* b;            // may be null
  float* acc;                // m floats (global accumulator across CTAs; fp32 mode)
  int* acc_lo;               // fixed-point mode: low 16 bits / high part of every CTA's 32-bit row sums, m ints each
  int* acc_hi;
  float fx_scale;            // 2^F
  double fx_inv;             // 2^-F
  const float* row_unscale;  // m floats 2^-k_r (rows are stored as a * 2^k_r), or null
  double* acc_scal;          // 2 doubles
  unsigned int* counter;
  float* grad_out;           // calc mode
  dualip_scalars* scalars_out;
  float* partial_out;        // partial mode: m+2 floats
  float* x_out;              // may be null
  uint8_t* diag;             // may be null
  int m;
  double gamma;
  float s;                   // fl32(-1/gamma)   (matching.py:136: `-1.0 / self.gamma * dual_val`)
  int flush_bulk;
  int do_epilogue;           // 1: calc (grad/scalars), 0: partial (packed sums)
  int stage;                 // slabs of up to this many entries per column are TMA-staged (per-warp buffer + mbarrier); 0: off
  unsigned long long* timeline;  // debug (DUALIP_TIMELINE=1): per CTA 5 x {clock64, globaltimer}, or null
  int stage_region;          // bytes of staging buffer per warp (multiple of 512), cut into 4 / 2 / 1 slots
  const float* long_a;
  const float* long_c;
  const uint32_t* long_row;
  const LongCol* mid_cols;   // columns of kRegDeg+1 .. kMaxThreadDeg entries, warp per column (mid_col.cuh), or null
  const int* mid_range;      // gridDim.x + 1 entries
  // fused tail: the CTA that finishes last also takes the accelerated step on the optimizer state (one launch per
  // iteration).  0: none; 1: after the objective's tail (single device); 2: sharded -- this rank's packed sums go into
  // its exchange slot, the peers' sums are fetched over NVLink, then the tail and the step (dualip_agd_step_peer's work).
  int fuse;
  AgdStepArgs agd;
  PeerArgs peer;
  // scheduled launch (fuse != 0 only): gamma, momentum, decay flag, log slot and exchange step number come from a
  // device-resident schedule at index *agd.pushes, so the kernel arguments are the same for every iteration and a sequence
  // of launches can be captured once in a CUDA graph and replayed (dualip_ascent_graph_*).  sched.gamma == null: off.
  SchedArgs sched;
  // grid-wide tail (grid_tail.cuh): all CTAs share the m-length tail, the exchange and the step.  0: the last CTA does it.
  int grid_tail;
  unsigned int* grid_bar;   // {arrive counter, generation}
  double* tail_part;        // gridDim.x x kTailPart doubles
  int* grid_status;         // set to 2 if a grid barrier timed out
  int* grid_status_host;    // the same, in mapped host memory
};

template <bool ROW16>
__device__ __forceinline__ uint32_t ld_row(const void* row_t, size_t idx) {
  if (ROW16) return __ldg(reinterpret_cast<const unsigned short*>(row_t) + idx);
  return __ldg(reinterpret_cast<const uint32_t*>(row_t) + idx);
}

// SMODE 0: scaled lambda and the gradient accumulator live in shared memory; 1: accumulator only (lambda is
// gathered from global/L2); 2: neither (global atomics; only for very large m).
template <int SMODE>
__device__ __forceinline__ float lam_scaled(const KArgs& k, float s, const float* s_lam, uint32_t r) {
  if (SMODE == 0) return s_lam[r];
  return __fmul_rn(s, __ldg(k.lambda + r));
}
template <int SMODE, int ACC>
__device__ __forceinline__ void grad_add(const KArgs& k, float* s_grad, uint32_t r, float g) {
  if (ACC == 1) {
    const int gi = __float2int_rn(g * k.fx_scale);
    if (gi != 0) atomicAdd(reinterpret_cast<int*>(s_grad) + r, gi);
  } else if (SMODE <= 1) {
    atomicAdd(&s_grad[r], g);
  } else {
    atomicAdd(&k.acc[r], g);
  }
}

// v = fl(fl(a * fl(s*lambda_r)) + fl(s*c)): the reference's operation order in fp32, no FMA contraction.
__device__ __forceinline__ float make_v(float a, float lam_s, float s, float c) {
  return __fadd_rn(__fmul_rn(a, lam_s), __fmul_rn(s, c));
}

// Largest double below x, for x >= 0 (so that {u > t} becomes {u >= x}).
__device__ __forceinline__ double just_below(double x) {
  return x > 0.0 ? __longlong_as_double(__double_as_longlong(x) - 1LL) : -4.9406564584124654e-324;
}

// Padded length L of the reference's [L x K] block for a column of class `cls` with d entries (sparse_utils.py:197,207):
// only simplex_eq depends on it.  Table index = ceil(log2(d)); 0 or absent = no padding.
__device__ __forceinline__ int pad_len_of(const KArgs& k, int cls, int d) {
  if (k.pad == nullptr) return d;
  const int bkt = (d <= 1) ? 0 : 32 - __clz(d - 1);
  return max(__ldg(k.pad + cls * DUALIP_PAD_BUCKETS + bkt), d);
}

}  // namespace dualip
#include "slab_fast.cuh"
#include "mid_col.cuh"
#include "long_col.cuh"
#include "grid_tail.cuh"
namespace dualip {

// Generic-path loads of N consecutive entries k0 .. k0+N-1 of this lane's column (N = 8/4: whole 4-entry chunks,
// N = 2 / 1: the tail chunks; k0 is where the plan's layout puts that chunk, see slab_elem()).
template <int N, typename RowT>
__device__ __forceinline__ void load_batch(const float* __restrict__ pa0, const float* __restrict__ pc0,
                                           const RowT* __restrict__ pr0, int k0, int lane, float (&a4)[N], float (&c4)[N],
                                           uint32_t (&r4)[N]) {
  if (N >= 4) {
#pragma unroll
    for (int j = 0; j < N / 4; ++j) {
      const size_t o = (size_t)(k0 + 4 * j) * kSlabW + (size_t)lane * 4;
      const float4 va = __ldg(reinterpret_cast<const float4*>(pa0 + o));
      const float4 vc = __ldg(reinterpret_cast<const float4*>(pc0 + o));
      a4[4 * j] = va.x, a4[4 * j + 1] = va.y, a4[4 * j + 2] = va.z, a4[4 * j + 3] = va.w;
      c4[4 * j] = vc.x, c4[4 * j + 1] = vc.y, c4[4 * j + 2] = vc.z, c4[4 * j + 3] = vc.w;
      if (sizeof(RowT) == 2) {
        const uint2 vr = __ldg(reinterpret_cast<const uint2*>(pr0 + o));
        r4[4 * j] = vr.x & 0xffffu, r4[4 * j + 1] = vr.x >> 16, r4[4 * j + 2] = vr.y & 0xffffu, r4[4 * j + 3] = vr.y >> 16;
      } else {
        const uint4 vr = __ldg(reinterpret_cast<const uint4*>(pr0 + o));
        r4[4 * j] = vr.x, r4[4 * j + 1] = vr.y, r4[4 * j + 2] = vr.z, r4[4 * j + 3] = vr.w;
      }
    }
  } else if (N == 2) {
    const size_t o = (size_t)k0 * kSlabW + (size_t)lane * 2;
    const float2 va = __ldg(reinterpret_cast<const float2*>(pa0 + o));
    const float2 vc = __ldg(reinterpret_cast<const float2*>(pc0 + o));
    a4[0] = va.x, a4[1] = va.y, c4[0] = vc.x, c4[1] = vc.y;
    if (sizeof(RowT) == 2) {
      const uint32_t vr = __ldg(reinterpret_cast<const uint32_t*>(pr0 + o));
      r4[0] = vr & 0xffffu, r4[1] = vr >> 16;
    } else {
      const uint2 vr = __ldg(reinterpret_cast<const uint2*>(pr0 + o));
      r4[0] = vr.x, r4[1] = vr.y;
    }
  } else {
    const size_t o = (size_t)k0 * kSlabW + (size_t)lane;
    a4[0] = __ldg(pa0 + o);
    c4[0] = __ldg(pc0 + o);
    r4[0] = (uint32_t)__ldg(pr0 + o);
  }
}

// m-length tail, executed by one whole CTA.  sum_load(i) = sum_j a_ij x_ij, cxv = c.x, xxv = ||x||^2.
// Reference: calc_grad (matching.py:25-34) and matching.py:164-178 / :280-299.
// sum_load(i) returns the grid-wide sum of row i, sum_clear(i) zeroes its accumulator slot for the next launch.  The loop
// is unrolled by four with all loads of a round issued before the first store or use: this runs on ONE CTA at the very
// end of the launch, so its latency is exposed.
template <typename SumFn, typename ClearFn>
__device__ __forceinline__ void cta_epilogue(SumFn sum_load, ClearFn sum_clear, double cxv, double xxv, const float* lambda, const float* b, int m,
                                             double gamma, float* grad_out, dualip_scalars* out, double* dscratch,
                                             float* fscratch, TailScratch& T) {
  double lg = 0.0, sp = 0.0, g2 = 0.0;
  float mx = -INFINITY;
  const int nt = blockDim.x;
  for (int base = threadIdx.x; base < m; base += 4 * nt) {
    float raw[4], bb[4], ll[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = base + u * nt;
      raw[u] = 0.f, bb[u] = 0.f, ll[u] = 0.f;
      if (i < m) {
        raw[u] = sum_load(i);
        bb[u] = b ? __ldg(b + i) : 0.f;
        ll[u] = lambda[i];
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = base + u * nt;
      if (i < m) {
        sum_clear(i);
        const float g = b ? __fsub_rn(raw[u], bb[u]) : raw[u];
        grad_out[i] = g;
        lg = fma((double)ll[u], (double)g, lg);
        sp += (double)fmaxf(g, 0.f);
        g2 = fma((double)g, (double)g, g2);
        mx = fmaxf(mx, g);
      }
    }
  }
  // one combined block reduction (dscratch: 32 doubles, reused three times; fscratch: 32 floats)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  lg = warp_sum(lg);
  sp = warp_sum(sp);
  g2 = warp_sum(g2);
  mx = warp_max(mx);
  double (&s_red)[5][32] = T.red;
  __syncthreads();
  if (lane == 0) {
    s_red[0][warp] = lg;
    s_red[1][warp] = sp;
    s_red[2][warp] = g2;
    fscratch[warp] = mx;
  }
  __syncthreads();
  if (warp == 0) {
    lg = warp_sum(lane < nw ? s_red[0][lane] : 0.0);
    sp = warp_sum(lane < nw ? s_red[1][lane] : 0.0);
    g2 = warp_sum(lane < nw ? s_red[2][lane] : 0.0);
    mx = warp_max(lane < nw ? fscratch[lane] : -INFINITY);
  }
  (void)dscratch;
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

// OUT: the launch writes the primal (save_primal) and / or the per-column diagnostics; a separate instantiation keeps
// those tests out of the hot loop of ordinary iterations.
template <bool ROW16, int SMODE, int ACC, int THREADS, int MINB, bool OUT>
__global__ void __launch_bounds__(THREADS, MINB) matching_slab_kernel(const KArgs k) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // carve: [mbarrier 16B][per-warp slot mbarriers 512B][classes][scratch 32 doubles + 32 floats][s_lam m_pad floats]
  //        [s_grad m_pad floats][per-warp stash (generic path) or per-warp staging buffers (register path)]
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* wbar = reinterpret_cast<uint64_t*>(smem_raw + 16);
  dualip_proj_class* s_cls = reinterpret_cast<dualip_proj_class*>(smem_raw + 16 + kBarBytes);
  const int n_cls_bytes = ((k.n_classes * (int)sizeof(dualip_proj_class)) + 15) & ~15;
  Seg* s_seg = reinterpret_cast<Seg*>(smem_raw + 16 + kBarBytes + n_cls_bytes);
  double* dscratch = reinterpret_cast<double*>(smem_raw + 16 + kBarBytes + n_cls_bytes + kMaxSeg * (int)sizeof(Seg));
  float* fscratch = reinterpret_cast<float*>(dscratch + 32);
  float* s_lam = fscratch + 32;
  const int m = k.m;
  const int m_pad = (m + 3) & ~3;
  float* s_grad = s_lam + (SMODE == 0 ? m_pad : 0);
  float* s_stash = s_grad + (SMODE <= 1 ? m_pad : 0);  // kStashDeg x 32 floats per warp (generic path)
  unsigned char* s_stage = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(s_stash) + 127) & ~(uintptr_t)127);  // kStageBytes per warp (register path)
  __shared__ unsigned int s_ticket;
  __shared__ int s_nseg;
  __shared__ TailScratch s_tail;  // the one static scratch of the tail code (cta_epilogue, agd_step_body, grid_tail)

  const unsigned FULL = 0xffffffffu;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = THREADS / 32;
  auto stamp = [&](int slot) {  // debug timeline
    if (k.timeline != nullptr && tid == 0) {
      unsigned long long g;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
      k.timeline[((size_t)blockIdx.x * 5 + slot) * 2] = (unsigned long long)clock64();
      k.timeline[((size_t)blockIdx.x * 5 + slot) * 2 + 1] = g;
    }
  };
  stamp(0);
  // per-iteration scalars: kernel arguments, or (scheduled launch) the schedule's entry for the iteration the optimizer
  // state is at.  The counter is advanced by the last CTA of a launch, after every CTA has passed this point.
  float s_run = k.s;
  if (k.sched.gamma != nullptr) {
    const long long it = __ldcg(k.agd.pushes);
    const double g = __ldg(k.sched.gamma + (it < (long long)k.sched.n ? it : (long long)k.sched.n - 1));
    s_run = (float)(-1.0 / g);  // matching.py:136 in Python double, rounded once when it meets the fp32 tensor
  }
  constexpr bool FAST = ROW16 && (SMODE == 0);  // register path (slab_fast.cuh)
  const uint32_t s_grad_u32 = pin_u32(smem_u32(s_grad));

  // ---- stage lambda into shared memory with a bulk async copy (TMA engine), then scale by -1/gamma ----
  bool bulk_ok = false;
  if (SMODE == 0) {
    const uint32_t bulk_bytes = (uint32_t)(m & ~3) * 4u;
    bulk_ok = ((reinterpret_cast<uintptr_t>(k.lambda) & 15u) == 0) && bulk_bytes >= 16;
    if (tid == 0) {
      mbar_init(bar, 1);
      for (int w = 0; w < (THREADS / 32) * kStageSlots; ++w) mbar_init(wbar + w, 1);
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
      for (int i = tid; i < m; i += THREADS) s_lam[i] = __fmul_rn(s_run, s_lam[i]);
    } else {
      for (int i = tid; i < m; i += THREADS) s_lam[i] = __fmul_rn(s_run, k.lambda[i]);
    }
    // rows stored as a * 2^k_r: fl(a*2^k * fl(s*lambda)*2^-k) == fl(a * fl(s*lambda)), powers of two commute with rounding
    if (k.row_unscale != nullptr)
      for (int i = tid; i < m; i += THREADS) s_lam[i] = __fmul_rn(s_lam[i], __ldg(k.row_unscale + i));
  }
  __syncthreads();

  stamp(1);
  unsigned long long t_loop0 = 0;
  if (tid == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_loop0));
  // ---- stream slabs.  Every CTA owns a contiguous range of the (length, class)-ordered slab sequence, cut at plan time so
  //      that all ranges cost about the same: an SM then runs the code of one or two column lengths only (instruction
  //      cache), and since groups of equal length and different class are neighbours, every range holds the problem's mix
  //      of classes.  The range is split into segments (one per group it touches); warp w takes slabs a+w, a+w+16, ... of a
  //      segment.  Inside a block of segments of equal length the warps start at different segments (rotation by warp
  //      index): at any time some warps of the SM run compute-heavy simplex slabs and others bandwidth-heavy clamp slabs. ----
  double cx = 0.0, xx = 0.0;
  const float s = s_run;
  using RowT = typename std::conditional<ROW16, unsigned short, uint32_t>::type;
  constexpr uint32_t RB = (uint32_t)row32_bytes(ROW16 ? 16 : 32);  // bytes per row of 32 entries
  const int2 range0 = k.cta_range[blockIdx.x];
  const int rb = range0.x, re = k.cta_range[blockIdx.x + 1].x;
  int g_first = range0.y;
  // TMA staging (register path): every warp owns a region of shared memory and two mbarriers.  Slabs of up to half the
  // region are staged two deep (slot = t mod 2, t = the warp's running slab count): while the warp works on slab t the
  // engine copies slabs t+1 and t+2; longer slabs take the whole region, one at a time.  ONE bulk copy per slab.
  const bool use_stage = FAST && (k.stage != 0);
  const uint32_t region = (uint32_t)k.stage_region;  // bytes per warp, multiple of 512
  unsigned char* my_stage = s_stage + (size_t)warp * region;
  uint64_t* my_bars = wbar + warp * kStageSlots;
  uint32_t phases = 0;  // phase bit per slot barrier
  int t = 0;            // running count of this warp's slabs
  for (bool more_batches = rb < re; more_batches;) {
    // ---- this CTA's segments (one batch, unless its range touches more than kMaxSeg groups) ----
    __syncthreads();
    if (warp == 0) {
      int count = 0;
      for (int round = 0; round < kMaxSeg / 32; ++round) {
        const int gi = g_first + count + lane;
        bool valid = gi < k.n_groups;
        SlabGroup G = {};
        if (valid) G = k.groups[gi];
        valid = valid && (int)G.slab_begin < re;
        if (valid) {
          const int g_end = (int)(G.slab_begin + G.n_slabs);
          Seg sg;
          sg.a = max(rb, (int)G.slab_begin);
          sg.b = min(re, g_end);
          sg.off32_a = G.off32_begin + (uint32_t)(sg.a - (int)G.slab_begin) * G.d;
          sg.d = G.d;
          sg.cls = G.cls;
          sg.last_ncols = (sg.b == g_end) ? G.last_ncols : (uint8_t)32;
          s_seg[count + lane] = sg;
        }
        const int nv = __popc(__ballot_sync(FULL, valid));
        count += nv;
        if (nv < 32) break;
      }
      if (lane == 0) s_nseg = count;
    }
    __syncthreads();
    const int n_seg = s_nseg;
    g_first += n_seg;
    more_batches = (n_seg == kMaxSeg);
    // ---- segment walk: blocks of segments of equal length, entered at a warp-dependent position ----
    struct Walk {
      int bs, bn, i, rot;  // current block [bs, bs+bn), position inside it, rotation
    };
    auto block_len = [&](int bs) {
      int n = 1;
      while (bs + n < n_seg && s_seg[bs + n].d == s_seg[bs].d) ++n;
      return n;
    };
    auto walk_next = [&](Walk& w) -> int {  // index of the next segment in which this warp has slabs, or -1
      for (;;) {
        ++w.i;
        if (w.i >= w.bn) {
          w.bs += w.bn;
          if (w.bs >= n_seg) return -1;
          w.bn = block_len(w.bs);
          w.rot = warp % w.bn;
          w.i = 0;
        }
        int idx = w.i + w.rot;
        if (idx >= w.bn) idx -= w.bn;
        idx += w.bs;
        if (s_seg[idx].a + warp < s_seg[idx].b) return idx;
      }
    };
    // producer cursor: the next slab to hand to the engine (it runs ahead of the consumer, across segments)
    Walk pw{0, 0, 0, 0};
    int p_sl = 0, p_t = t, p_end = 0, p_big = -1;
    uint32_t p_bytes = 0, p_off = 0, p_stride = 0;  // offsets in rows of 32 entries (RB bytes each)
    bool p_done = false;
    auto p_next_seg = [&]() {
      const int idx = walk_next(pw);
      if (idx < 0) {
        p_done = true;
        return;
      }
      const Seg S = s_seg[idx];
      p_sl = S.a + warp;
      p_end = S.b;
      // generic path (long columns, bisection classes): never staged
      p_bytes = (S.d > kRegDeg || s_cls[S.cls].kind >= DUALIP_PROJ_SIMPLEX_BISECT) ? 0xffffffffu : (uint32_t)S.d * RB;
      p_off = S.off32_a + (uint32_t)warp * S.d;
      p_stride = (uint32_t)NW * S.d;
    };
    // t_free: slabs < t_free (running count) have left their slots
    auto top_up = [&](int t_free) {
      while (!p_done) {
        if (p_sl >= p_end) {
          p_next_seg();
          continue;
        }
        if (!use_stage || p_bytes > region) {  // not staged: the cursor just moves on (at most one slab ahead)
          if (p_t > t_free) break;
        } else if (p_bytes * 2u <= region) {   // two deep
          if (p_big >= t_free || p_t > t_free + 1) break;
          if (lane == 0) stage_issue(my_stage + (p_t & 1) * (region >> 1), my_bars + (p_t & 1), k.data + (size_t)p_off * RB, p_bytes);
        } else {                               // whole region: only into an empty pipeline
          if (p_t != t_free) break;
          if (lane == 0) stage_issue(my_stage, my_bars, k.data + (size_t)p_off * RB, p_bytes);
          p_big = p_t;
        }
        ++p_t;
        p_sl += NW;
        p_off += p_stride;
      }
    };
    Walk cw{0, 0, 0, 0};
    for (int seg_idx = walk_next(cw); seg_idx >= 0; seg_idx = walk_next(cw)) {
      const Seg S = s_seg[seg_idx];
      const int d = (int)S.d, cls = (int)S.cls;
      const int g_end = S.b;
      int sl = S.a + warp;
      uint32_t boff = S.off32_a + (uint32_t)warp * S.d;  // in rows of 32 entries
      const uint32_t stride = (uint32_t)NW * S.d;
      const dualip_proj_class pc = s_cls[cls];
      if (FAST && d <= kRegDeg && pc.kind <= DUALIP_PROJ_SIMPLEX_EQ) {
        // ---- register path: the whole column lives in registers, code specialised on d ----
        const unsigned char* s_lam_b = reinterpret_cast<const unsigned char*>(s_lam);
        StageCtx st{use_stage, region, my_stage, my_bars};
        auto ensure_issued = [&](int tt) {
          if (p_t <= tt) top_up(tt);
        };
        auto after_load = [&](int tt) {  // slab tt is in registers: its slot is free
          __syncwarp();
          top_up(tt + 1);
        };
        switch (d) {
#define DUALIP_FAST_CASE(DD)                                                                                              \
  case DD:                                                                                                                \
    for (; sl < g_end; sl += NW, boff += stride, ++t) {                                                                   \
      const bool active = lane < ((sl == g_end - 1) ? (int)S.last_ncols : 32);                                            \
      fast_slab<DD, SMODE, ACC, OUT>(k, pc, cls, k.data + (size_t)boff * RB, lane, active, s_lam_b, s_grad_u32, s, sl, t, \
                                     cx, xx, st, phases, ensure_issued, after_load);                                      \
    }                                                                                                                     \
    break;
          DUALIP_FAST_CASE(1)
          DUALIP_FAST_CASE(2)
          DUALIP_FAST_CASE(3)
          DUALIP_FAST_CASE(4)
          DUALIP_FAST_CASE(5)
          DUALIP_FAST_CASE(6)
          DUALIP_FAST_CASE(7)
          DUALIP_FAST_CASE(8)
          DUALIP_FAST_CASE(9)
          DUALIP_FAST_CASE(10)
          DUALIP_FAST_CASE(11)
          DUALIP_FAST_CASE(12)
          DUALIP_FAST_CASE(13)
          DUALIP_FAST_CASE(14)
          DUALIP_FAST_CASE(15)
          DUALIP_FAST_CASE(16)
          DUALIP_FAST_CASE(17)
          DUALIP_FAST_CASE(18)
          DUALIP_FAST_CASE(19)
          DUALIP_FAST_CASE(20)
#undef DUALIP_FAST_CASE
          default:
            break;
        }
        continue;
      }
      // ---- generic path ----
      for (; sl < g_end; sl += NW, boff += stride, ++t) {
    const bool active = lane < ((sl == g_end - 1) ? (int)S.last_ncols : 32);
    const float* __restrict__ pa = reinterpret_cast<const float*>(k.data + (size_t)boff * RB);
    const float* __restrict__ pcv = pa + (size_t)d * kSlabW;
    const RowT* __restrict__ pr = reinterpret_cast<const RowT*>(pcv + (size_t)d * kSlabW);
    auto issue_next = [&]() {
      __syncwarp();
      top_up(t + 1);
    };
    issue_next();  // the generic path does not use the staging buffer
    auto ld1 = [&](int kq, float& av, float& cv, uint32_t& rv) {
      const uint32_t o = slab_elem(kq, d, lane);
      av = __ldg(pa + o);
      cv = __ldg(pcv + o);
      rv = (uint32_t)__ldg(pr + o);
    };
    float cxs = 0.f, xxs = 0.f;
    // per-lane outcome, also what the (rare) output pass needs: x_k = branch 0: u_k, 1: z*[k == i1], 2: max(u_k - theta, 0)
    int branch = -1, rho = 0, i1 = 0;
    float theta = 0.f;

    if (pc.kind == DUALIP_PROJ_CLAMP) {
      // ---- box / cone / identity: one streaming pass (box.py:16, cone.py:21-28) ----
      const float lo = active ? pc.lo : 0.f, hi = active ? pc.hi : 0.f;  // padding lanes produce x = 0
      auto body = [&](float a, float c, uint32_t r) {
        const float v = make_v(a, lam_scaled<SMODE>(k, s, s_lam, r), s, c);
        const float x = fminf(fmaxf(v, lo), hi);
        const float g = __fmul_rn(a, x);
        if (g != 0.f) grad_add<SMODE, ACC>(k, s_grad, r, g);
        cxs = fmaf(c, x, cxs);
        xxs = fmaf(x, x, xxs);
      };
      // batches of 8, then 4/2/1 for the tail: every batch issues all of its loads before the first use
      auto batch = [&](auto n_tag, int k0) {
        constexpr int N = decltype(n_tag)::value;
        float a4[N], c4[N];
        uint32_t r4[N];
        load_batch<N, RowT>(pa, pcv, pr, k0, lane, a4, c4, r4);
#pragma unroll
        for (int q = 0; q < N; ++q) body(a4[q], c4[q], r4[q]);
      };
      int kk = 0;
      for (; kk + 8 <= d; kk += 8) {
        batch(std::integral_constant<int, 8>{}, kk);
        if ((kk & 63) == 56) {  // keep the fp32 partials short
          cx += (double)cxs;
          xx += (double)xxs;
          cxs = 0.f;
          xxs = 0.f;
        }
      }
      if (d & 4) {
        batch(std::integral_constant<int, 4>{}, kk);
        kk += 4;
      }
      if (d & 2) {
        batch(std::integral_constant<int, 2>{}, kk);
        kk += 2;
      }
      if (d & 1) batch(std::integral_constant<int, 1>{}, kk);
    } else if (pc.kind >= DUALIP_PROJ_SIMPLEX_BISECT) {
      // ---- method="bisection_search" (simplex.py:6-123) on the zero-padded column of the reference's block: no pre-clamp;
      //      every sum runs down the block's rows in fp32 (torch's outer-dimension reduction), the padding after the entries ----
      const float z = pc.z;
      const int n0 = pad_len_of(k, cls, d) - d;  // zero rows below this column in its bucket's block
      auto vq = [&](int kq) -> float {
        float a, c;
        uint32_t r;
        ld1(kq, a, c, r);
        return make_v(a, lam_scaled<SMODE>(k, s, s_lam, r), s, c);
      };
      float ssum = 0.f, vmin = INFINITY, t0 = -INFINITY, t1 = -INFINITY;  // column sum, smallest entry, two largest of x/z
      int am = -1;
      for (int kq = 0; kq < d; ++kq) {
        const float v = vq(kq);
        ssum = __fadd_rn(ssum, v);
        vmin = fminf(vmin, v);
        const float xn = __fdiv_rn(v, z);
        if (xn > t0) {
          t1 = t0, t0 = xn, am = kq;
        } else if (xn > t1) {
          t1 = xn;
        }
      }
      if (n0 > 0) {  // the padding's zeros take part in the top-2 (a stable sort puts an entry before an equal padding zero)
        if (0.f > t0) {
          t1 = (n0 > 1) ? 0.f : t0, t0 = 0.f, am = -1;
        } else if (0.f > t1) {
          t1 = 0.f;
        }
      }
      const bool feas = (pc.kind == DUALIP_PROJ_SIMPLEX_BISECT) && (ssum <= pc.z_thr) && (vmin >= -1e-6f);  // :40 (padding: 0 >= -tol)
      const bool shortc = !feas && (d + n0 > 1) && (__fsub_rn(t0, t1) > 1.0f);                                                 // :52-67
      float nu = 0.f;
      if (feas) {
        branch = 4;  // x = v, unclamped (:41)
      } else if (shortc) {
        branch = 1, rho = 1, i1 = am;  // one-hot at the maximum (:69-75); am < 0: it fell on a padding row
      } else {
        branch = 3;
        theta = t0;  // max(x/z): the shift (:86-89)
        float lo = -1.f, hi = 0.f, prev = 0.f;
        bool act = true;
        for (int it = 0; it < 50 && act; ++it) {  // :95-118; every column of a block halves the same interval
          const float mid = __fmul_rn(__fadd_rn(lo, hi), 0.5f);
          if (it > 0 && fabsf(__fsub_rn(mid, prev)) < 1e-6f) break;
          float sm = 0.f;
          for (int kq = 0; kq < d; ++kq) sm = __fadd_rn(sm, fmaxf(__fsub_rn(__fsub_rn(vq(kq), t0), mid), 0.f));
          const float tp = fmaxf(__fsub_rn(__fsub_rn(0.f, t0), mid), 0.f);
          for (int q0 = 0; q0 < n0; ++q0) sm = __fadd_rn(sm, tp);
          const bool high = sm > 1.0f;
          lo = high ? mid : lo;
          hi = high ? hi : mid;
          act = !(__fsub_rn(hi, lo) < 1e-6f);
          prev = mid;
        }
        nu = __fmul_rn(__fadd_rn(lo, hi), 0.5f);
      }
      if (active) {
        for (int kq = 0; kq < d; ++kq) {
          float a, c;
          uint32_t r;
          ld1(kq, a, c, r);
          const float v = make_v(a, lam_scaled<SMODE>(k, s, s_lam, r), s, c);
          const float x = branch == 4 ? v : (branch == 1 ? (kq == i1 ? z : 0.f) : __fmul_rn(fmaxf(__fsub_rn(__fsub_rn(v, theta), nu), 0.f), z));
          const float g = __fmul_rn(a, x);
          if (g != 0.f) grad_add<SMODE, ACC>(k, s_grad, r, g);
          cxs = fmaf(c, x, cxs);
          xxs = fmaf(x, x, xxs);
          if (OUT && k.x_out) k.x_out[k.orig_start[(int64_t)sl * kSlabW + lane] + kq] = x;
        }
      }
      branch = -2;  // the output pass below is done
    } else {
      // ---- simplex (simplex.py:143-236 per column at its true length) ----
      // Pass 1 streams the column once: column sum, the two largest entries with their positions, the third largest
      // value, and (for d <= kStashDeg) a copy of u in this warp's shared-memory stash.  That decides feasible / top-2
      // shortcut and resolves the sorted scan in closed form when the support has at most two entries.  Lanes with a
      // larger support run a Michelot threshold search on the stash; lanes with more than two non-zeros take pass 2.
      const float z = pc.z;
      // The stash holds u of the whole slab (d x 32 floats).  Kernels with the register path reach this code only for
      // columns longer than it handles; those sit at the END of a CTA's range (the slab order is length-major), so no staged
      // copy is in flight any more and the warp's staging region serves as the stash: the threshold search and pass 2 then
      // read shared memory instead of re-streaming the column from L2.
      const int stash_cap = FAST ? (use_stage ? (int)(region / (kSlabW * sizeof(float))) : 0) : kStashDeg;
      const bool stash = d <= stash_cap;
      float* __restrict__ su = (FAST ? reinterpret_cast<float*>(my_stage) : s_stash + (size_t)warp * (kStashDeg * kSlabW)) + lane;
      float S = 0.f, m1 = -1.f, m2 = -1.f, m3 = -1.f;
      int i2 = 0;
      const bool is_eq = pc.kind == DUALIP_PROJ_SIMPLEX_EQ;
      double Sd = 0.0;  // simplex_eq: the column sum in double (css_d of the reference's scan)
      auto track = [&](float a, float c, uint32_t r, int kq) {
        const float u = fmaxf(make_v(a, lam_scaled<SMODE>(k, s, s_lam, r), s, c), 0.f);
        if (stash) su[kq * kSlabW] = u;
        if (is_eq) Sd += (double)u;
        S = __fadd_rn(S, u);  // column sum in entry order
        const bool g1 = u > m1, g2 = u > m2;
        m3 = fmaxf(m3, fminf(m2, u));
        m2 = g1 ? m1 : (g2 ? u : m2);
        i2 = g1 ? i1 : (g2 ? kq : i2);
        m1 = g1 ? u : m1;
        i1 = g1 ? kq : i1;
      };
      auto batch = [&](auto n_tag, int k0) {
        constexpr int N = decltype(n_tag)::value;
        float a4[N], c4[N];
        uint32_t r4[N];
        load_batch<N, RowT>(pa, pcv, pr, k0, lane, a4, c4, r4);
#pragma unroll
        for (int q = 0; q < N; ++q) track(a4[q], c4[q], r4[q], k0 + q);
      };
      int kk = 0;
      for (; kk + 8 <= d; kk += 8) batch(std::integral_constant<int, 8>{}, kk);
      if (d & 4) {
        batch(std::integral_constant<int, 4>{}, kk);
        kk += 4;
      }
      if (d & 2) {
        batch(std::integral_constant<int, 2>{}, kk);
        kk += 2;
      }
      if (d & 1) batch(std::integral_constant<int, 1>{}, kk);

      const bool feasible = (pc.kind == DUALIP_PROJ_SIMPLEX) && (S <= pc.z_thr);                    // simplex.py:153-155
      const bool padded = (d > 1) || !(pc.flags & DUALIP_PROJ_FLAG_D1_UNPADDED);                    // simplex.py:166
      const float m2p = fmaxf(m2, 0.f), m3p = fmaxf(m3, 0.f);  // the reference's zero padding takes part in its top-2
      const float un1 = (z == 1.0f) ? m1 : __fdiv_rn(m1, z), un2 = (z == 1.0f) ? m2p : __fdiv_rn(m2p, z);
      const bool shortcut = !feasible && padded && (__fsub_rn(un1, un2) > 1.0f);                    // simplex.py:172-178
      float t3lo = -1.f;            // lower bound on theta from the top-3 scan (threshold search start)
      float x1 = 0.f, x2 = 0.f;     // results for the two largest entries when they are the only non-zeros
      bool need_theta = false;      // support of three or more: threshold search
      bool need_p2 = false;         // more than two non-zeros: second streaming pass
      // simplex_eq with a clamped sum below z: in the reference's scan over the zero-padded column every padded position
      // satisfies cond_i (0 - (css_d - z)/i > 0), so rho = L (the bucket's padded length) and theta = (css_d - z)/L < 0:
      // every entry grows by -theta (simplex.py:160-161,207-233).  The shortcut needs u_(1) > z: excluded.
      const float t_below = is_eq ? __fsub_rn((float)Sd, z) : 0.f;
      if (is_eq && t_below < 0.f) {
        branch = 2;
        rho = pad_len_of(k, cls, d);
        theta = __fdiv_rn(t_below, (float)rho);
        need_p2 = true;
      } else if (feasible) {
        branch = 0;
        if (m3p > 0.f) {
          need_p2 = true;
        } else {
          x1 = m1;
          x2 = m2p;
        }
      } else if (shortcut) {
        branch = 1;
        rho = 1;
        x1 = z;                                                                                     // simplex.py:185-190
      } else {
        branch = 2;
        // sorted scan restricted to the three largest: css_i = fl32(prefix sum in fp64), cond_i = u_(i) - (css_i - z)/i > 0
        const float css2 = (float)((double)m1 + (double)m2p);
        const float t2 = __fmul_rn(__fsub_rn(css2, z), 0.5f);  // division by 2 is exact
        const bool cond2 = (d >= 2) && (__fsub_rn(m2, t2) > 0.f);
        bool cond3 = false;
        if (d >= 3 && m3p > __fsub_rn(__fsub_rn(m1, z), 1e-3f * fabsf(m1))) {  // cond_3 needs m3 > m1 - z (exactly)
          const float css3 = (float)((double)m1 + (double)m2p + (double)m3p);
          const float t3 = __fdiv_rn(__fsub_rn(css3, z), 3.0f);
          cond3 = __fsub_rn(m3p, t3) > 0.f;
          t3lo = __fsub_rd(t3, 4e-7f * fabsf(t3));
        }
        if (cond3) {
          need_theta = true;
          need_p2 = true;
        } else {
          rho = cond2 ? 2 : 1;
          theta = cond2 ? t2 : __fsub_rn(m1, z);                                                    // simplex.py:228-230
          if (m3p > theta) {
            need_p2 = true;  // rounding leaves a third entry above the threshold: take the general pass
          } else {
            x1 = fmaxf(__fsub_rn(m1, theta), 0.f);                                                  // simplex.py:233
            x2 = (d >= 2) ? fmaxf(__fsub_rn(m2, theta), 0.f) : 0.f;
          }
        }
      }
      need_theta = need_theta && active;
      need_p2 = need_p2 && active;
      if (!active) {
        x1 = 0.f;
        x2 = 0.f;
      }
      // non-zeros of the closed-form lanes: their entries are re-read from lines this warp has just streamed
      auto emit = [&](int kq, float x) {
        if (x != 0.f) {
          float a, c;
          uint32_t r;
          ld1(kq, a, c, r);
          const float g = __fmul_rn(a, x);
          if (g != 0.f) grad_add<SMODE, ACC>(k, s_grad, r, g);
          cxs = fmaf(c, x, cxs);
          xxs = fmaf(x, x, xxs);
        }
      };
      if (__any_sync(FULL, x1 != 0.f)) emit(i1, x1);
      if (__any_sync(FULL, x2 != 0.f)) emit(i2, x2);

      if (__any_sync(FULL, need_theta)) {
        // ---- Michelot fixed point, then alignment with the reference's fp32 conditions (simplex.py:207-231) ----
        // The sequence t_0 = (sum - z)/d <= t_1 <= ... increases to theta*, and theta* >= max - z, so the search
        // starts from the larger of the two.  (double)u > t  <=>  u > round_down_to_float(t).
        auto search = [&](auto stash_tag) {
          constexpr bool ST = decltype(stash_tag)::value;
          auto uq = [&](int kq) -> float {
            if (ST) return su[kq * kSlabW];
            float a, c;
            uint32_t r;
            ld1(kq, a, c, r);
            return fmaxf(make_v(a, lam_scaled<SMODE>(k, s, s_lam, r), s, c), 0.f);
          };
          // Newton steps from below on f(t) = sum max(u - t, 0) - z (Michelot): t <- (sum_{u>t} u - z)/#{u>t}.  Lower
          // bounds to start from: (S - z)/d, max - z, and the top-3 scan value t3 (the scan thresholds increase while
          // their conditions hold).  fp32 sums are scaled down by `guard` so that no step overshoots theta*.
          const float guard = 1.0f - 2.4e-7f * (float)d;
          float tf = fmaxf(fmaxf(__fdiv_rd(__fsub_rd(__fmul_rd(S, guard), z), (float)d), __fsub_rd(m1, z)), t3lo);
          tf = (tf > 0.f) ? __uint_as_float(__float_as_uint(tf) - 1u) : -1.f;  // strictly below the bound
          int cnt = 0;
          for (int it = 0; it < 64; ++it) {
            cnt = 0;
            float fsum = 0.f, umin = INFINITY;
#pragma unroll 4
            for (int kq = 0; kq < d; ++kq) {
              const float u = uq(kq);
              const bool in = u > tf;
              cnt += in ? 1 : 0;
              fsum += in ? u : 0.f;
              umin = in ? fminf(umin, u) : umin;
            }
            // next step; converged when it removes nothing, i.e. the smallest support value stays above it
            const float tn = __fdiv_rd(__fsub_rd(__fmul_rd(fsum, guard), z), (float)max(cnt, 1));
            const bool done = !need_theta || cnt == 0 || !(tn > tf) || umin > tn;
            if (!done) tf = tn;
            if (__all_sync(FULL, done)) break;
          }
          // exact sums over the support and the two boundary values, then the reference's own conditions
          float th = 0.f;
          for (int fix = 0; fix < 64; ++fix) {
            double ssum = 0.0;
            float umin = INFINITY, uout = -INFINITY;
            cnt = 0;
#pragma unroll 2
            for (int kq = 0; kq < d; ++kq) {
              const float u = uq(kq);
              const bool in = u > tf;
              cnt += in ? 1 : 0;
              ssum += in ? (double)u : 0.0;
              umin = in ? fminf(umin, u) : umin;
              uout = in ? uout : fmaxf(uout, u);
            }
            th = __fdiv_rn(__fsub_rn((float)ssum, z), (float)max(cnt, 1));
            bool changed = false;
            if (need_theta && cnt > 1 && !(__fsub_rn(umin, th) > 0.f)) {
              // cond_rho fails in the fp32 formula: every support value <= th goes (a Michelot step with exact sums; th >= umin
              // here, and the largest value stays in: th < max unless rounding at huge magnitudes, hence the clamp).  Dropping
              // only the smallest value per round needs one round per value within the guard band below theta, which long
              // columns with closely spaced values (ratings data) exceed.
              tf = fmaxf(umin, fminf(th, __uint_as_float(__float_as_uint(m1) - 1u)));
              changed = true;
            } else if (need_theta && uout > -INFINITY) {
              const float t1 = __fdiv_rn(__fsub_rn((float)(ssum + (double)uout), z), (float)(cnt + 1));
              if (__fsub_rn(uout, t1) > 0.f) {  // cond_{rho+1} holds: the support grows
                tf = (uout > 0.f) ? __uint_as_float(__float_as_uint(uout) - 1u) : -1.f;
                changed = true;
              }
            }
            if (!__any_sync(FULL, changed)) break;
          }
          if (need_theta) {
            theta = th;
            rho = max(cnt, 1);
          }
        };
        if (stash)
          search(std::true_type{});
        else
          search(std::false_type{});
      }
      if (__any_sync(FULL, need_p2)) {
        // ---- pass 2: x_k = u_k (feasible) or max(u_k - theta, 0); re-reads lines this warp has just streamed ----
        auto pass2 = [&](auto stash_tag) {
          constexpr bool ST = decltype(stash_tag)::value;
          auto scatter = [&](float a, float c, uint32_t r, int kq) {
            float u;
            if (ST)
              u = su[kq * kSlabW];
            else
              u = fmaxf(make_v(a, lam_scaled<SMODE>(k, s, s_lam, r), s, c), 0.f);
            const float x = (branch == 0) ? u : fmaxf(__fsub_rn(u, theta), 0.f);
            if (need_p2 && x != 0.f) {
              const float g = __fmul_rn(a, x);
              if (g != 0.f) grad_add<SMODE, ACC>(k, s_grad, r, g);
              cxs = fmaf(c, x, cxs);
              xxs = fmaf(x, x, xxs);
            }
          };
          auto batch2 = [&](auto n_tag, int k0) {
            constexpr int N = decltype(n_tag)::value;
            float a4[N], c4[N];
            uint32_t r4[N];
            load_batch<N, RowT>(pa, pcv, pr, k0, lane, a4, c4, r4);
#pragma unroll
            for (int q = 0; q < N; ++q) scatter(a4[q], c4[q], r4[q], k0 + q);
          };
          int k2 = 0;
          for (; k2 + 8 <= d; k2 += 8) {
            batch2(std::integral_constant<int, 8>{}, k2);
            if ((k2 & 63) == 56) {
              cx += (double)cxs;
              xx += (double)xxs;
              cxs = 0.f;
              xxs = 0.f;
            }
          }
          if (d & 4) {
            batch2(std::integral_constant<int, 4>{}, k2);
            k2 += 4;
          }
          if (d & 2) {
            batch2(std::integral_constant<int, 2>{}, k2);
            k2 += 2;
          }
          if (d & 1) batch2(std::integral_constant<int, 1>{}, k2);
        };
        if (stash)
          pass2(std::true_type{});
        else
          pass2(std::false_type{});
      }
    }
    cx += (double)cxs;
    xx += (double)xxs;

    if (OUT && active && branch != -2) {
      // ---- primal / diagnostics output (save_primal on the last iteration, tests): plain re-stream ----
      const int64_t os = k.orig_start[(int64_t)sl * kSlabW + lane];
      if (k.x_out) {
        for (int kq = 0; kq < d; ++kq) {
          float a, c;
          uint32_t r;
          ld1(kq, a, c, r);
          const float v = make_v(a, lam_scaled<SMODE>(k, s, s_lam, r), s, c);
          float x;
          if (branch < 0)
            x = fminf(fmaxf(v, pc.lo), pc.hi);
          else if (branch == 0)
            x = fmaxf(v, 0.f);
          else if (branch == 1)
            x = (kq == i1) ? pc.z : 0.f;
          else
            x = fmaxf(__fsub_rn(fmaxf(v, 0.f), theta), 0.f);
          k.x_out[os + kq] = x;
        }
      }
      if (k.diag && branch >= 0) k.diag[os] = (uint8_t)(branch | (min(rho, 63) << 2));
    }
      }  // slabs of the segment (generic path)
    }    // segments
  }      // batches

  // ---- this CTA's share of the mid columns, a warp per column (plans with the register path only) ----
  if (FAST && k.mid_cols != nullptr)
    mid_columns_of_cta<ACC, OUT, NW>(k, k.mid_cols, k.mid_range[blockIdx.x], k.mid_range[blockIdx.x + 1], warp, lane, s_cls, s_lam,
                                     s_grad, s, cx, xx);

  // ---- flush per-CTA partial sums ----
  __syncthreads();
  if (tid == 0 && k.cta_ns != nullptr) {  // how long this CTA's range took: feedback for dualip_plan_rebalance
    unsigned long long t_loop1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_loop1));
    k.cta_ns[blockIdx.x] = (unsigned int)min(t_loop1 - t_loop0, 0xffffffffull);
  }
  stamp(2);
  cx = block_sum(cx, dscratch);
  xx = block_sum(xx, dscratch);
  if (tid == 0) {
    // one slot per CTA, added up in CTA order by the tail: the objective's scalars are reproducible bit for bit, like the
    // fixed-point gradient (a floating-point atomic per CTA would add them in arrival order)
    k.tail_part[(size_t)blockIdx.x * kTailPart + 6] = cx;
    k.tail_part[(size_t)blockIdx.x * kTailPart + 7] = xx;
  }
  __syncthreads();
  if (ACC == 1) {
    // fixed point: every CTA sum s (|s| < 2^31) is split as s = hi*65536 + lo with 0 <= lo < 65536, so that the
    // grid-wide totals of both parts fit 32 bits; lo stays in place, hi reuses the (now dead) lambda array; two TMA
    // bulk reductions (SASS UBLKRED, element-wise s32 add performed at L2) flush them.
    int* s_lo = reinterpret_cast<int*>(s_grad);
    int* s_hi = reinterpret_cast<int*>(s_lam);
    for (int i = tid; i < m; i += THREADS) {
      const int v = s_lo[i];
      s_lo[i] = v & 0xffff;
      s_hi[i] = v >> 16;
    }
    const uint32_t bulk_bytes = (uint32_t)(m & ~3) * 4u;
    const bool bulk = k.flush_bulk && bulk_bytes >= 16 && ((reinterpret_cast<uintptr_t>(k.acc_lo) & 15u) == 0) &&
                      ((reinterpret_cast<uintptr_t>(k.acc_hi) & 15u) == 0);
    fence_proxy_async_smem();
    __syncthreads();
    if (bulk) {
      if (tid == 0) {
        bulk_reduce_add_s32_s2g(k.acc_lo, s_lo, bulk_bytes);
        bulk_reduce_add_s32_s2g(k.acc_hi, s_hi, bulk_bytes);
        bulk_commit();
        bulk_wait_all();
      }
    }
    for (int i = (bulk ? (m & ~3) : 0) + tid; i < m; i += THREADS) {
      if (s_lo[i] != 0) atomicAdd(&k.acc_lo[i], s_lo[i]);
      if (s_hi[i] != 0) atomicAdd(&k.acc_hi[i], s_hi[i]);
    }
  } else if (SMODE <= 1) {
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
  stamp(3);
  __threadfence();
  auto sum_load = [&](int i) -> float {
    const double un = k.row_unscale ? (double)__ldg(k.row_unscale + i) : 1.0;  // exact: a power of two
    if (ACC == 1) {
      const long long v = (long long)__ldcg(k.acc_hi + i) * 65536LL + (long long)__ldcg(k.acc_lo + i);
      return (float)((double)v * k.fx_inv * un);
    }
    return (float)((double)__ldcg(k.acc + i) * un);
  };
  auto sum_clear = [&](int i) {  // leaves the accumulators zeroed for the next launch
    if (ACC == 1) {
      k.acc_lo[i] = 0;
      k.acc_hi[i] = 0;
    } else {
      k.acc[i] = 0.f;
    }
  };
  const bool scheduled = k.sched.gamma != nullptr;
  // (read again rather than kept in registers across the main loop; the counter moves only at the very end of the tail, behind
  // barriers that every thread reaches after this point)
  long long sched_it = 0;
  double gamma_run = k.gamma;
  if (scheduled) {
    sched_it = __ldcg(k.agd.pushes);
    gamma_run = __ldg(k.sched.gamma + (sched_it < (long long)k.sched.n ? sched_it : (long long)k.sched.n - 1));
  }
  if (k.grid_tail) {
    // ---- all CTAs share the tail, the exchange and the step (grid_tail.cuh) ----
    const StepDyn dyn = scheduled ? step_dyn_sched(k.agd, k.sched, sched_it) : step_dyn_of(k.agd);
    const unsigned long long seq_g = scheduled ? (unsigned long long)(k.sched.seq_base + sched_it + 1) : k.peer.seq;
    if (k.fuse == 2)
      grid_tail<true, true>(k, sum_load, sum_clear, dyn, gamma_run, seq_g, s_tail);
    else if (k.fuse == 1)
      grid_tail<false, true>(k, sum_load, sum_clear, dyn, gamma_run, seq_g, s_tail);
    else if (k.fuse == 3)
      grid_tail<true, false>(k, sum_load, sum_clear, dyn, gamma_run, seq_g, s_tail);
    else
      grid_tail<false, false>(k, sum_load, sum_clear, dyn, gamma_run, seq_g, s_tail);
    stamp(4);
    return;
  }
  // ---- last CTA to finish runs the m-length tail and leaves the accumulators zeroed ----
  __syncthreads();
  if (tid == 0) s_ticket = atomicAdd(k.counter, 1u);
  __syncthreads();
  if (s_ticket != gridDim.x - 1) return;
  __threadfence();
  double cxv, xxv;
  load_scalar_sums(k, s_tail, cxv, xxv);
  if (k.do_epilogue) {
    cta_epilogue(sum_load, sum_clear, cxv, xxv, k.lambda, k.b, m, gamma_run, k.grad_out, k.scalars_out, dscratch, fscratch, s_tail);
    if (k.fuse == 1) {
      __syncthreads();  // grad_out / scalars_out written above are read by other threads of this CTA
      agd_step_body<false>(k.agd, scheduled ? step_dyn_sched(k.agd, k.sched, sched_it) : step_dyn_of(k.agd), s_tail);
    }
  } else {
    // sharded + scheduled: the exchange step number follows the device-side iteration count, and so does the slot
    const unsigned long long seq = scheduled ? (unsigned long long)(k.sched.seq_base + sched_it + 1) : k.peer.seq;
    float* partial_out = k.partial_out;
    if (k.fuse == 2 && scheduled)  // (fuse == 3 is never scheduled)
      partial_out = reinterpret_cast<float*>(k.peer.win[k.peer.rank] + kPeerFlagBytes + (size_t)(seq & 1ull) * k.peer.slot_bytes);
    const bool push = (k.fuse == 2 || k.fuse == 3) && k.peer.push != 0;
    float* pslot[DUALIP_PEER_MAX_WORLD];  // push: this rank's slot in every rank's window
#pragma unroll
    for (int r = 0; r < DUALIP_PEER_MAX_WORLD; ++r) pslot[r] = (push && r < k.peer.world) ? peer_push_slot(k.peer, r, k.peer.rank, seq) : nullptr;
    for (int base = tid; base < m; base += 4 * THREADS) {
      float raw[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) raw[u] = (base + u * THREADS < m) ? sum_load(base + u * THREADS) : 0.f;
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (base + u * THREADS < m) {
          sum_clear(base + u * THREADS);
          if (push) {
#pragma unroll
            for (int r = 0; r < DUALIP_PEER_MAX_WORLD; ++r)
              if (r < k.peer.world) pslot[r][base + u * THREADS] = raw[u];  // posted stores over NVLink (own window: local)
          } else {
            partial_out[base + u * THREADS] = raw[u];
          }
        }
    }
    if (tid == 0) {
      if (push) {
#pragma unroll
        for (int r = 0; r < DUALIP_PEER_MAX_WORLD; ++r)
          if (r < k.peer.world) pslot[r][m] = (float)cxv, pslot[r][m + 1] = (float)xxv;
      } else {
        partial_out[m] = (float)cxv;
        partial_out[m + 1] = (float)xxv;
      }
    }
    if (k.fuse == 2 || k.fuse == 3) {
      if (push)
        peer_push_exchange_cta(k.peer, m + 2, seq);
      else
        peer_exchange_cta(k.peer, m + 2, seq);
      if (k.fuse == 2) {
        agd_step_body<true>(k.agd, scheduled ? step_dyn_sched(k.agd, k.sched, sched_it) : step_dyn_of(k.agd), s_tail);
      } else {
        // sharded evaluation for a caller that keeps the iterate itself (host-buffer path): the m-length tail on the summed
        // vector, no optimizer step
        const float* sum = k.peer.sum;
        cta_epilogue([&](int i) { return sum[i]; }, [](int) {}, (double)sum[m], (double)sum[m + 1], k.lambda, k.b, m, gamma_run,
                     k.grad_out, k.scalars_out, dscratch, fscratch, s_tail);
      }
    }
  }
  if (tid == 0) {
    k.acc_scal[0] = 0.0;
    k.acc_scal[1] = 0.0;
    *k.counter = 0u;
  }
  __syncthreads();
  stamp(4);
}

// ------------------------------------------------------------------------------------------
// Very long columns (> kMaxThreadDeg entries): one warp per column, multi-sweep over the plan's compact copy.
// Threshold by Michelot's fixed point (same support set and theta formula as the sorted scan in exact arithmetic);
// sums over the support are fp64 like the reference's CPU cumsum.
// ------------------------------------------------------------------------------------------
template <int ACC>
__global__ void __launch_bounds__(256) matching_long_kernel(const KArgs k, const LongCol* __restrict__ cols, int64_t n_long) {
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  double cx = 0.0, xx = 0.0;
  float s_run = k.s;
  if (k.sched.gamma != nullptr) {  // scheduled launch: see matching_slab_kernel
    const long long it = __ldcg(k.agd.pushes);
    s_run = (float)(-1.0 / __ldg(k.sched.gamma + (it < (long long)k.sched.n ? it : (long long)k.sched.n - 1)));
  }
  for (int64_t ci = warp_global; ci < n_long; ci += n_warps) {
    const LongCol lc = cols[ci];
    const dualip_proj_class pc = k.classes[lc.cls];
    const float* a = k.long_a + lc.off;
    const float* c = k.long_c + lc.off;
    const uint32_t* row = k.long_row + lc.off;
    auto v_at = [&](int e, float& av, float& cv, uint32_t& rv) -> float {
      av = __ldg(a + e);
      cv = __ldg(c + e);
      rv = __ldg(row + e);
      float ls = __fmul_rn(s_run, __ldg(k.lambda + rv));
      if (k.row_unscale != nullptr) ls = __fmul_rn(ls, __ldg(k.row_unscale + rv));  // long_a holds a * 2^k_r
      return make_v(av, ls, s_run, cv);
    };
    const bool is_sx = pc.kind != DUALIP_PROJ_CLAMP;
    float theta = 0.f;
    int branch = 0, rho = 0, amax = -1;
    if (is_sx) {
      double S = 0.0;
      float m1 = -1.f, m2 = 0.f;
      int am = 0x7fffffff;
      for (int e = lane; e < lc.len; e += 32) {
        float av, cv;
        uint32_t rv;
        const float u = fmaxf(v_at(e, av, cv, rv), 0.f);
        const float un = __fdiv_rn(u, pc.z);
        S += (double)u;
        if (un > m1) am = e;
        m2 = fmaxf(m2, fminf(m1, un));
        m1 = fmaxf(m1, un);
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
      const float t_below = __fsub_rn((float)S, pc.z);
      if (pc.kind == DUALIP_PROJ_SIMPLEX_EQ && t_below < 0.f) {  // see the generic path of matching_slab_kernel
        branch = 2;
        rho = pad_len_of(k, lc.cls, lc.len);
        theta = __fdiv_rn(t_below, (float)rho);
      } else if (feasible) {
        branch = 0;
      } else if (shortcut) {
        branch = 1;
        rho = 1;
        amax = am;
      } else {
        branch = 2;
        double t = -INFINITY, Ssup = S;
        int cnt_prev = -1;
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
          if (done) break;
          t = (s2 - (double)pc.z) / (double)n2;
        }
        rho = max(cnt_prev, 1);
        theta = __fdiv_rn(__fsub_rn((float)Ssup, pc.z), (float)rho);
      }
    }
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
      if (ACC == 1) {
        const long long gi = __double2ll_rn((double)g * (double)k.fx_scale);
        if (gi != 0) {
          atomicAdd(&k.acc_lo[rv], (int)(gi & 0xffff));
          atomicAdd(&k.acc_hi[rv], (int)(gi >> 16));
        }
      } else if (g != 0.f) {
        atomicAdd(&k.acc[rv], g);
      }
      const double xd = (double)x;
      cx = fma((double)cv, xd, cx);
      xx = fma(xd, xd, xx);
      if (k.x_out) k.x_out[lc.src_start + e] = x;
      if (k.diag && is_sx && e == 0) k.diag[lc.src_start] = (uint8_t)(branch | (min(rho, 63) << 2));
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
  __shared__ TailScratch s_tail;
  const double cxv = (double)sum[m], xxv = (double)sum[m + 1];
  cta_epilogue([&](int i) { return __ldcg(sum + i); }, [](int) {}, cxv, xxv, lambda, b, m, gamma, grad_out, out, dscratch,
               fscratch, s_tail);
}

// ------------------------------------------------------------------------------------------
// Plan time: bound on every CTA's row sums, for the fixed-point accumulator.  Same slab -> CTA assignment as the hot
// kernel (CTA b owns the slabs of its range).  |a * x| <= |a| * xmax(class).
// ------------------------------------------------------------------------------------------
template <typename RowT>
__global__ void cta_row_bound_kernel(const unsigned char* __restrict__ data,
                                     const SlabHdr* __restrict__ hdr, const int2* __restrict__ cta_range, int64_t n_slabs,
                                     const float* __restrict__ cls_xmax,
                                     int m, float* __restrict__ table, unsigned int* __restrict__ row_cnt,
                                     const LongCol* __restrict__ mid_cols, const int* __restrict__ mid_range,
                                     const float* __restrict__ long_a, const uint32_t* __restrict__ long_row) {
  extern __shared__ __align__(16) unsigned char bound_smem[];
  float* s_bound = reinterpret_cast<float*>(bound_smem);              // m floats: this CTA's row bounds
  unsigned int* s_cnt = reinterpret_cast<unsigned int*>(s_bound + m);  // m counters
  for (int i = threadIdx.x; i < m; i += blockDim.x) {
    s_bound[i] = 0.f;
    s_cnt[i] = 0u;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, NW = blockDim.x >> 5;
  (void)n_slabs;
  for (int sl = cta_range[blockIdx.x].x + warp; sl < cta_range[blockIdx.x + 1].x; sl += NW) {
    const SlabHdr h = hdr[sl];
    if (lane >= (int)h.ncols) continue;
    const float xmax = cls_xmax[h.cls];
    const float* a_t = reinterpret_cast<const float*>(data + (size_t)h.off32 * (size_t)row32_bytes(8 * (int)sizeof(RowT)));
    const RowT* row_t = reinterpret_cast<const RowT*>(a_t + 2 * (size_t)h.d * kSlabW);
    for (int k = 0; k < (int)h.d; ++k) {
      const uint32_t idx = slab_elem(k, (int)h.d, lane);
      const uint32_t r = (uint32_t)row_t[idx];
      atomicAdd(&s_bound[r], fabsf(a_t[idx]) * xmax);
      atomicAdd(&s_cnt[r], 1u);
    }
  }
  if (mid_cols != nullptr) {  // the CTA's mid columns add to the same shared-memory accumulator
    for (int ci = mid_range[blockIdx.x] + warp; ci < mid_range[blockIdx.x + 1]; ci += NW) {
      const LongCol lc = mid_cols[ci];
      const float xmax = cls_xmax[lc.cls];
      for (int e = lane; e < lc.len; e += 32) {
        const uint32_t r = long_row[lc.off + e];
        atomicAdd(&s_bound[r], fabsf(long_a[lc.off + e]) * xmax);
        atomicAdd(&s_cnt[r], 1u);
      }
    }
  }
  __syncthreads();
  float* my = table + (size_t)blockIdx.x * m;
  for (int i = threadIdx.x; i < m; i += blockDim.x) {
    my[i] = s_bound[i];
    if (s_cnt[i]) atomicAdd(&row_cnt[i], s_cnt[i]);
  }
}

__global__ void long_row_bound_kernel(const LongCol* __restrict__ cols, int64_t n_long, const float* __restrict__ la,
                                      const uint32_t* __restrict__ lrow, const float* __restrict__ cls_xmax,
                                      float* __restrict__ long_bound, unsigned int* __restrict__ row_cnt) {
  const int lane = threadIdx.x & 31;
  int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (; w < n_long; w += nw) {
    const LongCol lc = cols[w];
    const float xmax = cls_xmax[lc.cls];
    for (int e = lane; e < lc.len; e += 32) {
      atomicAdd(&long_bound[lrow[lc.off + e]], fabsf(la[lc.off + e]) * xmax);
      atomicAdd(&row_cnt[lrow[lc.off + e]], 1u);
    }
  }
}

// per row: largest CTA bound and the sum over CTAs
__global__ void bound_reduce_kernel(const float* __restrict__ table, int n_ctas, int m, float* __restrict__ row_max,
                                    float* __restrict__ row_sum) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= m) return;
  float mx = 0.f, sm = 0.f;
  for (int c = 0; c < n_ctas; ++c) {
    const float v = table[(size_t)c * m + r];
    mx = fmaxf(mx, v);
    sm += v;
  }
  row_max[r] = mx;
  row_sum[r] = sm;
}

// ------------------------------------------------------------------------------------------
// Launch plumbing
// ------------------------------------------------------------------------------------------
typedef void (*SlabKernel)(const KArgs);

// Variants: the register path and fixed-point accumulation exist for uint16 rows with lambda and the accumulator in
// shared memory (m up to ~24k); larger m streams through the generic path with fp32 accumulation.
template <bool OUT>
static SlabKernel plan_kernel_out(const dualip_plan* p) {
  const bool row16 = p->row_bits == 16;
  if (row16 && p->smode == 0)
    return p->fixed_point ? matching_slab_kernel<true, 0, 1, kThreads, 1, OUT> : matching_slab_kernel<true, 0, 0, kThreads, 1, OUT>;
  if (row16 && p->smode == 1) return matching_slab_kernel<true, 1, 0, kThreads, 1, OUT>;
  if (row16 && p->smode == 2) return matching_slab_kernel<true, 2, 0, kThreads, 1, OUT>;
  if (!row16 && p->smode == 1) return matching_slab_kernel<false, 1, 0, kThreads, 1, OUT>;
  if (!row16 && p->smode == 2) return matching_slab_kernel<false, 2, 0, kThreads, 1, OUT>;
  return nullptr;
}
static SlabKernel plan_kernel(const dualip_plan* p, bool out) { return out ? plan_kernel_out<true>(p) : plan_kernel_out<false>(p); }

static size_t smem_fixed_bytes(int n_classes) {
  return 16 + kBarBytes + (((size_t)n_classes * sizeof(dualip_proj_class) + 15) & ~(size_t)15) + kMaxSeg * sizeof(Seg) +
         32 * sizeof(double) + 32 * sizeof(float);
}

struct FuseSpec {
  int mode = 0;
  AgdStepArgs agd = {};
  PeerArgs peer = {};
  SchedArgs sched = {};
};

static int launch_eval(dualip_plan* p, const float* lambda, const float* b, double gamma, float* grad_out,
                       dualip_scalars* scalars_out, float* partial_out, float* x_out, uint8_t* diag, int do_epilogue,
                       cudaStream_t stream, const FuseSpec* fuse = nullptr) {
  if (!(gamma > 0.0) && !(gamma < 0.0)) {
    set_error("gamma must be non-zero");
    return DUALIP_EINVAL;
  }
  KArgs k;
  k.data = p->data;
  k.groups = p->groups;
  k.n_groups = p->n_groups;
  k.cta_range = p->cta_range;
  k.cta_ns = p->cta_ns;
  k.orig_start = p->orig_start;
  k.n_slabs = p->n_slabs;
  k.classes = p->classes_dev;
  k.n_classes = p->n_classes;
  k.pad = p->pad_dev;
  k.lambda = lambda;
  k.b = b;
  k.acc = p->acc;
  k.acc_lo = p->acc_lo;
  k.acc_hi = p->acc_hi;
  k.fx_scale = p->fx_scale;
  k.fx_inv = p->fx_inv;
  k.row_unscale = p->row_unscale;
  k.acc_scal = p->acc_scal;
  k.counter = p->counter;
  k.grad_out = grad_out;
  k.scalars_out = scalars_out;
  k.partial_out = partial_out;
  k.x_out = x_out;
  k.diag = diag;
  k.m = p->m;
  k.gamma = gamma;
  k.s = (float)(-1.0 / gamma);
  k.flush_bulk = p->flush_bulk;
  k.do_epilogue = do_epilogue;
  k.stage_region = p->stage_region;
  k.timeline = p->timeline;
  k.stage = p->stage;
  k.long_a = p->long_a;
  k.long_c = p->long_c;
  k.long_row = p->long_row;
  k.mid_cols = (p->n_mid > 0 && !p->mid_separate) ? p->longcols : nullptr;
  k.mid_range = p->mid_range;
  k.fuse = fuse ? fuse->mode : 0;
  // the all-CTA tail needs every CTA of the grid resident at once (one CTA per SM) and, sharded, the push exchange
  {
    const int mode = fuse ? fuse->mode : 0;
    const bool ok_mode = (mode == 0 && do_epilogue) || mode == 1 || ((mode == 2 || mode == 3) && fuse->peer.push != 0);
    // sharded, the single CTA's tail also grows with the world size (it stores its sums into W windows and adds W slots): at
    // 8 ranks and m = 10 000 sharing it is worth 6.8 % of the step (2935 -> 3135 it/s, A/B/A/B on one box), at 2 ranks nothing
    const bool wide_world = (mode == 2 || mode == 3) && fuse->peer.world >= 4 && p->m >= 8192;
    const bool want = p->grid_tail || (p->grid_tail_auto && wide_world);
    k.grid_tail = (want && p->n_ctas <= p->n_sms && x_out == nullptr && diag == nullptr && ok_mode) ? 1 : 0;
    p->last_grid_tail = k.grid_tail;
  }
  k.grid_bar = p->grid_bar;
  k.tail_part = p->tail_part;
  k.grid_status = p->grid_status;
  k.grid_status_host = p->grid_status_host_dev;
  if (fuse) {
    k.agd = fuse->agd;
    k.peer = fuse->peer;
    k.sched = fuse->sched;
  } else {
    memset(&k.agd, 0, sizeof(k.agd));
    memset(&k.peer, 0, sizeof(k.peer));
    memset(&k.sched, 0, sizeof(k.sched));
  }
  if (p->n_mid > 0 && p->mid_separate) {  // many mid columns: 32 threads per column, u in a 4 KB shared-memory stash per team
    const int team_bytes = (kLongThreads / 32) * kTeamStash * (int)sizeof(float);
    const int blocks = (int)std::min<int64_t>((p->n_mid + 7) / 8, (int64_t)p->n_sms * 5);
    if (p->fixed_point)
      matching_long_cta_kernel<1, 32><<<blocks, kLongThreads, team_bytes, stream>>>(k, p->longcols, (int)p->n_mid);
    else
      matching_long_cta_kernel<0, 32><<<blocks, kLongThreads, team_bytes, stream>>>(k, p->longcols, (int)p->n_mid);
  }
  if (p->n_ctalong > 0) {  // 1025 .. kLongStash entries: a CTA per column, u in shared memory
    const int stash_bytes = kLongStash * (int)sizeof(float);  // the attribute was raised at plan creation (per device)
    const int blocks = (int)std::min<int64_t>(p->n_ctalong, (int64_t)p->n_sms * 4);
    if (p->fixed_point)
      matching_long_cta_kernel<1, kLongThreads><<<blocks, kLongThreads, stash_bytes, stream>>>(k, p->longcols + p->n_mid, (int)p->n_ctalong);
    else
      matching_long_cta_kernel<0, kLongThreads><<<blocks, kLongThreads, stash_bytes, stream>>>(k, p->longcols + p->n_mid, (int)p->n_ctalong);
  }
  if (p->n_long > p->n_mid + p->n_ctalong) {  // longer still: warp per column, re-streamed
    const int64_t nl = p->n_long - p->n_mid - p->n_ctalong;
    const int blocks = (int)std::min<int64_t>((nl + 7) / 8, (int64_t)p->n_sms * 8);
    if (p->fixed_point)
      matching_long_kernel<1><<<blocks, 256, 0, stream>>>(k, p->longcols + p->n_mid + p->n_ctalong, nl);
    else
      matching_long_kernel<0><<<blocks, 256, 0, stream>>>(k, p->longcols + p->n_mid + p->n_ctalong, nl);
  }
  SlabKernel kern = plan_kernel(p, x_out != nullptr || diag != nullptr);
  kern<<<p->n_ctas, p->threads, p->smem_bytes, stream>>>(k);
  DUALIP_CUDA_TRY(cudaGetLastError());
  return DUALIP_OK;
}

struct Group {
  uint32_t key;
  int64_t start, count, slab_base, off32;
};

template <typename IdxT>
static int build_slabs(dualip_plan* p, const dualip_csc_desc* d, cudaStream_t stream) {
  const int64_t n = p->n_cols;
  const IdxT* ccol = reinterpret_cast<const IdxT*>(d->ccol_dev);
  const IdxT* row = reinterpret_cast<const IdxT*>(d->row_dev);
  unsigned int* bad = nullptr;
  uint32_t *keys = nullptr, *vals = nullptr, *keys_s = nullptr, *perm = nullptr, *uniq = nullptr, *counts = nullptr;
  int* n_runs_dev = nullptr;
  void* tmp = nullptr;
  int64_t *g_start_d = nullptr, *g_slab_d = nullptr, *g_off_d = nullptr;
  uint32_t* g_key_d = nullptr;
  uint8_t* slab_only_d = nullptr;
  int rc = DUALIP_OK;
  auto cleanup = [&]() {
    cudaFree(slab_only_d);
    cudaFree(bad);
    cudaFree(keys);
    cudaFree(vals);
    cudaFree(keys_s);
    cudaFree(perm);
    cudaFree(uniq);
    cudaFree(counts);
    cudaFree(n_runs_dev);
    cudaFree(tmp);
    cudaFree(g_start_d);
    cudaFree(g_slab_d);
    cudaFree(g_off_d);
    cudaFree(g_key_d);
  };
#define BS_TRY(expr)                                                            \
  do {                                                                          \
    cudaError_t _e = (expr);                                                    \
    if (_e != cudaSuccess) {                                                    \
      set_error("%s failed: %s", #expr, cudaGetErrorString(_e));                \
      cleanup();                                                                \
      return _e == cudaErrorMemoryAllocation ? DUALIP_ENOMEM : DUALIP_ECUDA;    \
    }                                                                           \
  } while (0)
  const int64_t n_alloc = std::max<int64_t>(n, 1);
  const int64_t max_runs = std::min<int64_t>(n_alloc, (int64_t)256 * 2048) + 2;
  BS_TRY(cudaMalloc(&bad, sizeof(unsigned int)));
  BS_TRY(cudaMemsetAsync(bad, 0, sizeof(unsigned int), stream));
  BS_TRY(cudaMalloc(&keys, sizeof(uint32_t) * n_alloc));
  BS_TRY(cudaMalloc(&vals, sizeof(uint32_t) * n_alloc));
  BS_TRY(cudaMalloc(&keys_s, sizeof(uint32_t) * n_alloc));
  BS_TRY(cudaMalloc(&perm, sizeof(uint32_t) * n_alloc));
  BS_TRY(cudaMalloc(&uniq, sizeof(uint32_t) * max_runs));
  BS_TRY(cudaMalloc(&counts, sizeof(uint32_t) * max_runs));
  BS_TRY(cudaMalloc(&n_runs_dev, sizeof(int)));
  BS_TRY(cudaMemsetAsync(n_runs_dev, 0, sizeof(int), stream));
  std::vector<Group> groups;
  int64_t n_short = 0, n_long = 0, long_start = 0;
  if (n > 0) {
    const int tb = 256;
    const int nb = (int)std::min<int64_t>((n + tb - 1) / tb, (int64_t)p->n_sms * 32);
    {
      uint8_t slab_only[kMaxClasses + 1] = {};
      for (int i = 0; i < p->n_classes; ++i) slab_only[i] = p->classes_host[i].kind >= DUALIP_PROJ_SIMPLEX_BISECT ? 1 : 0;
      BS_TRY(cudaMalloc(&slab_only_d, kMaxClasses + 1));
      BS_TRY(cudaMemcpyAsync(slab_only_d, slab_only, kMaxClasses + 1, cudaMemcpyHostToDevice, stream));
      BS_TRY(cudaStreamSynchronize(stream));
    }
    // Where the slab layout ends.  Plans with the register path hand longer columns to the warp-per-column path -- except
    // that MANY columns just above kRegDeg are cheaper lane-per-column (32 columns advance together in a warp, and up to
    // stage_region / 128 entries per lane fit the warp's shared-memory stash), while a FEW of them would leave each warp a
    // serial chain of slow slabs.  The length histogram decides.  DUALIP_MID=0: slabs up to 1024 entries (generic path);
    // DUALIP_MID=<d>: slabs up to d entries.
    const bool fast_plan = p->row_bits == 16 && p->smode == 0;
    const char* env_mid = getenv("DUALIP_MID");
    int slab_max_deg = kMaxThreadDeg;
    if (fast_plan) {
      slab_max_deg = kRegDeg;
      unsigned int* hist_d = nullptr;
      std::vector<unsigned int> hist(kMaxThreadDeg + 2, 0u);
      BS_TRY(cudaMalloc(&hist_d, sizeof(unsigned int) * hist.size()));
      cudaError_t he = cudaMemsetAsync(hist_d, 0, sizeof(unsigned int) * hist.size(), stream);
      length_hist_kernel<IdxT><<<nb, tb, 0, stream>>>(ccol, n, hist_d);
      if (he == cudaSuccess) he = cudaMemcpyAsync(hist.data(), hist_d, sizeof(unsigned int) * hist.size(), cudaMemcpyDeviceToHost, stream);
      if (he == cudaSuccess) he = cudaStreamSynchronize(stream);
      cudaFree(hist_d);
      BS_TRY(he);
      const int stash_cap = p->stage ? std::min(p->stage_region / (kSlabW * (int)sizeof(float)), 64) : 0;
      int64_t short_cols = 0;
      for (int dd = kRegDeg + 1; dd <= stash_cap; ++dd) short_cols += hist[dd];
      if (short_cols > (int64_t)8 * p->n_sms * (p->threads / 32)) slab_max_deg = stash_cap;
      if (env_mid) slab_max_deg = atoi(env_mid) <= 0 ? kMaxThreadDeg : std::max(kRegDeg, std::min(atoi(env_mid), kMaxThreadDeg));
    }
    column_keys_kernel<IdxT><<<nb, tb, 0, stream>>>(ccol, d->col_class_dev, n, p->n_classes, slab_max_deg, slab_only_d, keys, vals, bad);
    size_t b1 = 0, b2 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, b1, keys, keys_s, vals, perm, (int)n, 0, kKeyBits, stream);
    cub::DeviceRunLengthEncode::Encode(nullptr, b2, keys_s, uniq, counts, n_runs_dev, (int)n, stream);
    BS_TRY(cudaMalloc(&tmp, std::max(b1, b2) + 16));
    size_t tb1 = std::max(b1, b2) + 16;
    cub::DeviceRadixSort::SortPairs(tmp, tb1, keys, keys_s, vals, perm, (int)n, 0, kKeyBits, stream);
    tb1 = std::max(b1, b2) + 16;
    cub::DeviceRunLengthEncode::Encode(tmp, tb1, keys_s, uniq, counts, n_runs_dev, (int)n, stream);
    int n_runs = 0;
    unsigned int bad_h = 0;
    BS_TRY(cudaMemcpyAsync(&n_runs, n_runs_dev, sizeof(int), cudaMemcpyDeviceToHost, stream));
    BS_TRY(cudaMemcpyAsync(&bad_h, bad, sizeof(bad_h), cudaMemcpyDeviceToHost, stream));
    BS_TRY(cudaStreamSynchronize(stream));
    if (bad_h) {
      set_error(bad_h & 8u ? "col_class id out of range" : "ccol_indices must be non-decreasing (flags=%u)", bad_h);
      cleanup();
      return DUALIP_EINVAL;
    }
    std::vector<uint32_t> uk(n_runs), uc(n_runs);
    if (n_runs > 0) {
      BS_TRY(cudaMemcpy(uk.data(), uniq, sizeof(uint32_t) * n_runs, cudaMemcpyDeviceToHost));
      BS_TRY(cudaMemcpy(uc.data(), counts, sizeof(uint32_t) * n_runs, cudaMemcpyDeviceToHost));
    }
    int64_t pos = 0, slab = 0, off32 = 0;
    for (int r = 0; r < n_runs; ++r) {
      if (uk[r] == kKeyLong) {
        long_start = pos;
        n_long = uc[r];
      } else if (uk[r] != kKeyEmpty) {
        Group g;
        g.key = uk[r];
        g.start = pos;
        g.count = uc[r];
        g.slab_base = slab;
        g.off32 = off32;
        const int64_t ns = (g.count + kSlabW - 1) / kSlabW;
        slab += ns;
        off32 += ns * (int64_t)key_d(g.key);
        groups.push_back(g);
        p->class_used[key_cls(g.key)] = true;
        n_short = pos + g.count;
      }
      pos += uc[r];
    }
    p->n_slabs = slab;
    p->rows32 = off32;
  }
  p->n_short = n_short;
  p->n_long = n_long;
  if (p->rows32 >= (1LL << 32)) {
    set_error("shard too large for 32-bit slab offsets; shard the columns");
    cleanup();
    return DUALIP_ERANGE;
  }
  // slab storage
  const size_t data_bytes = (size_t)std::max<int64_t>(p->rows32, 1) * (size_t)row32_bytes(p->row_bits);
  BS_TRY(cudaMalloc(&p->data, data_bytes + 512));  // + slack: vector loads of a partially filled last slab stay inside
  BS_TRY(cudaMalloc(&p->hdr, sizeof(SlabHdr) * std::max<int64_t>(p->n_slabs, 1)));
  BS_TRY(cudaMalloc(&p->orig_start, sizeof(int64_t) * std::max<int64_t>(p->n_slabs, 1) * kSlabW));
  p->owned_bytes += data_bytes + (sizeof(SlabHdr) + 8 * kSlabW) * (size_t)std::max<int64_t>(p->n_slabs, 1);
  BS_TRY(cudaMemsetAsync(p->data, 0, data_bytes + 512, stream));
  BS_TRY(cudaMemsetAsync(p->orig_start, 0xff, sizeof(int64_t) * std::max<int64_t>(p->n_slabs, 1) * kSlabW, stream));
  if (n_short > 0) {
    const int G = (int)groups.size();
    std::vector<int64_t> gs(G), gb(G), go(G);
    std::vector<uint32_t> gk(G);
    for (int i = 0; i < G; ++i) {
      gs[i] = groups[i].start;
      gb[i] = groups[i].slab_base;
      go[i] = groups[i].off32;
      gk[i] = groups[i].key;
    }
    BS_TRY(cudaMalloc(&g_start_d, sizeof(int64_t) * G));
    BS_TRY(cudaMalloc(&g_slab_d, sizeof(int64_t) * G));
    BS_TRY(cudaMalloc(&g_off_d, sizeof(int64_t) * G));
    BS_TRY(cudaMalloc(&g_key_d, sizeof(uint32_t) * G));
    BS_TRY(cudaMemcpyAsync(g_start_d, gs.data(), sizeof(int64_t) * G, cudaMemcpyHostToDevice, stream));
    BS_TRY(cudaMemcpyAsync(g_slab_d, gb.data(), sizeof(int64_t) * G, cudaMemcpyHostToDevice, stream));
    BS_TRY(cudaMemcpyAsync(g_off_d, go.data(), sizeof(int64_t) * G, cudaMemcpyHostToDevice, stream));
    BS_TRY(cudaMemcpyAsync(g_key_d, gk.data(), sizeof(uint32_t) * G, cudaMemcpyHostToDevice, stream));
    const int tb = 128;
    const unsigned nb = (unsigned)((n_short + tb - 1) / tb);
    if (p->row_bits == 16)
      fill_slabs_kernel<IdxT, unsigned short><<<nb, tb, 0, stream>>>(ccol, row, d->a_dev, d->c_dev, perm, n_short, g_start_d,
                                                                      g_slab_d, g_off_d, g_key_d, G, p->data, p->hdr,
                                                                      p->orig_start, p->m, bad);
    else
      fill_slabs_kernel<IdxT, uint32_t><<<nb, tb, 0, stream>>>(ccol, row, d->a_dev, d->c_dev, perm, n_short, g_start_d, g_slab_d,
                                                                g_off_d, g_key_d, G, p->data, p->hdr, p->orig_start, p->m,
                                                                bad);
    // group table of the hot kernel
    std::vector<SlabGroup> gt(G);
    for (int i = 0; i < G; ++i) {
      const int64_t ns = (groups[i].count + kSlabW - 1) / kSlabW;
      gt[i].slab_begin = (uint32_t)groups[i].slab_base;
      gt[i].off32_begin = (uint32_t)groups[i].off32;
      gt[i].d = (uint16_t)key_d(groups[i].key);
      gt[i].cls = (uint8_t)key_cls(groups[i].key);
      gt[i].last_ncols = (uint8_t)(groups[i].count - (ns - 1) * kSlabW);
      gt[i].n_slabs = (uint32_t)ns;
    }
    BS_TRY(cudaMalloc(&p->groups, sizeof(SlabGroup) * G));
    BS_TRY(cudaMemcpyAsync(p->groups, gt.data(), sizeof(SlabGroup) * G, cudaMemcpyHostToDevice, stream));
    p->n_groups = G;
    p->groups_host = gt;
    BS_TRY(cudaStreamSynchronize(stream));  // host vectors go out of scope
  }
  if (n_long > 0) {
    BS_TRY(cudaMalloc(&p->longcols, sizeof(LongCol) * n_long));
    const int tb = 128;
    long_meta_kernel<IdxT><<<(unsigned)((n_long + tb - 1) / tb), tb, 0, stream>>>(ccol, d->col_class_dev, perm + long_start,
                                                                                  n_long, p->longcols);
    std::vector<LongCol> lc(n_long);
    BS_TRY(cudaMemcpyAsync(lc.data(), p->longcols, sizeof(LongCol) * n_long, cudaMemcpyDeviceToHost, stream));
    BS_TRY(cudaStreamSynchronize(stream));
    // by length: the mid columns (warp per column inside the slab kernel) come first, and a warp's consecutive columns take
    // the same specialisation
    std::stable_sort(lc.begin(), lc.end(), [](const LongCol& x, const LongCol& y) { return x.len < y.len; });
    int64_t tot = 0;
    p->n_mid = 0;
    const char* env_lc = getenv("DUALIP_LONG_CTA");  // DUALIP_LONG_CTA=0: all long columns take the warp-per-column kernel
    const bool use_ctalong = !(env_lc && atoi(env_lc) == 0);
    p->n_ctalong = 0;
    for (auto& c : lc) {
      if (c.len <= kMaxThreadDeg) ++p->n_mid;
      else if (use_ctalong && c.len <= kLongStash) ++p->n_ctalong;
      c.off = tot;
      tot += c.len;
      p->class_used[c.cls & 0xff] = true;
      if (p->classes_host[c.cls & 0xff].kind >= DUALIP_PROJ_SIMPLEX_BISECT) {
        set_error("bisection-search classes are limited to columns of at most %d entries (a column has %d)", kMaxThreadDeg, c.len);
        cleanup();
        return DUALIP_ERANGE;
      }
    }
    {
      // few mid columns (the tail of a short-column problem): inside the slab kernel, sharing lambda and the accumulator in
      // shared memory; many (ratings-like data): their own launch at 2.5x the occupancy.  DUALIP_MID_KERNEL=0|1 forces it.
      const char* env_mk = getenv("DUALIP_MID_KERNEL");
      p->mid_separate = env_mk ? atoi(env_mk) != 0 : p->n_mid > (int64_t)4 * p->n_sms * (p->threads / 32);
    }
    BS_TRY(cudaMemcpyAsync(p->longcols, lc.data(), sizeof(LongCol) * n_long, cudaMemcpyHostToDevice, stream));
    p->long_total = tot;
    BS_TRY(cudaMalloc(&p->long_a, sizeof(float) * tot));
    BS_TRY(cudaMalloc(&p->long_c, sizeof(float) * tot));
    BS_TRY(cudaMalloc(&p->long_row, sizeof(uint32_t) * tot));
    p->owned_bytes += 12 * (size_t)tot + sizeof(LongCol) * n_long;
    const int blocks = (int)std::min<int64_t>((n_long + 3) / 4, (int64_t)p->n_sms * 16);
    long_copy_kernel<IdxT><<<blocks, 128, 0, stream>>>(p->longcols, n_long, row, d->a_dev, d->c_dev, p->long_a, p->long_c,
                                                       p->long_row, p->m, bad);
    BS_TRY(cudaStreamSynchronize(stream));
  }
  unsigned int bad_h = 0;
  BS_TRY(cudaMemcpyAsync(&bad_h, bad, sizeof(bad_h), cudaMemcpyDeviceToHost, stream));
  BS_TRY(cudaStreamSynchronize(stream));
  cleanup();
#undef BS_TRY
  if (bad_h) {
    set_error("row index out of range [0,%d)", p->m);
    return DUALIP_EINVAL;
  }
  return rc;
}

// Cost of a slab by projection kind and column length d, fitted on B200 to per-CTA main-loop times of the C3 workload at a
// late iterate (100M- and 12.5M-entity shards agree within 10 %; residual of the fit 2 % mean, 6 % max: profiles/r2), in
// units of 7.5 ns per CTA:
//   simplex  4.0 + d (d <= 14, staged two deep)   1.27 d (15..16)   1.61 d (17..20, a and c are read twice)
//   clamp    2.8 + 0.56 d (d <= 14)               0.76 d (15..20)
//   generic path (longer columns, or a plan without the register path): simplex 6.7 d (iterative threshold search), clamp 2.7 d.
// The table only seeds the partition: dualip_plan_rebalance corrects it with measured per-CTA times.
static double slab_cost(const dualip_plan* p, const SlabGroup& g) {
  const bool fast = p->row_bits == 16 && p->smode == 0;
  const bool simplex = p->classes_host[g.cls].kind != DUALIP_PROJ_CLAMP;
  const double d = (double)g.d;
  if (p->classes_host[g.cls].kind >= DUALIP_PROJ_SIMPLEX_BISECT) return 30.0 * d;  // ~20 passes over the column (generic path)
  if (!fast || g.d > kRegDeg) return simplex ? 6.7 * d : 2.7 * d;
  if (simplex) return g.d <= 14 ? 4.0 + d : (g.d <= 16 ? 1.27 * d : 1.61 * d);
  return g.d <= 14 ? 2.8 + 0.56 * d : 0.76 * d;
}

// A run of consecutive slabs of one group with one cost per slab.
struct CostPiece {
  int group;
  int64_t slab_begin, n_slabs;
  double cost;
};

// Cuts the slab sequence (given as cost pieces in slab order) into n_ctas contiguous ranges of about equal cost.
static std::vector<int2> cut_ranges(const dualip_plan* p, const std::vector<CostPiece>& pieces) {
  double total = 0.0;
  for (const CostPiece& c : pieces) total += c.cost * (double)c.n_slabs;
  std::vector<int2> r((size_t)p->n_ctas + 1);
  const int P = (int)pieces.size(), G = (int)p->groups_host.size();
  int pi = 0;
  int64_t used = 0;
  double acc = 0.0;
  for (int c = 0; c < p->n_ctas; ++c) {
    while (pi < P && used >= pieces[pi].n_slabs) {
      ++pi;
      used = 0;
    }
    r[c].x = pi < P ? (int)(pieces[pi].slab_begin + used) : (int)p->n_slabs;
    r[c].y = pi < P ? pieces[pi].group : G;
    const double target = total * (double)(c + 1) / (double)p->n_ctas;
    while (pi < P && acc < target) {
      const CostPiece& pc = pieces[pi];
      const int64_t left = pc.n_slabs - used;
      int64_t take = pc.cost > 0.0 ? (int64_t)ceil((target - acc) / pc.cost) : left;
      take = std::max<int64_t>(0, std::min(take, left));
      used += take;
      acc += pc.cost * (double)take;
      if (used >= pc.n_slabs) {
        ++pi;
        used = 0;
      } else {
        break;
      }
    }
  }
  r[p->n_ctas].x = (int)p->n_slabs;
  r[p->n_ctas].y = G;
  return r;
}

static int upload_ranges(dualip_plan* p, const std::vector<int2>& r, cudaStream_t stream) {
  if (!p->cta_range && cudaMalloc(&p->cta_range, sizeof(int2) * r.size()) != cudaSuccess) {
    set_error("allocating the CTA range table failed");
    return DUALIP_ECUDA;
  }
  p->ranges_host = r;  // the copy below reads this vector: it lives as long as the plan, and a later upload synchronises first
  if (cudaMemcpyAsync(p->cta_range, p->ranges_host.data(), sizeof(int2) * r.size(), cudaMemcpyHostToDevice, stream) != cudaSuccess ||
      cudaStreamSynchronize(stream) != cudaSuccess) {
    set_error("uploading the CTA range table failed");
    return DUALIP_ECUDA;
  }
  return DUALIP_OK;
}

static int build_cta_ranges(dualip_plan* p) {
  std::vector<CostPiece> pieces;
  for (int g = 0; g < (int)p->groups_host.size(); ++g) {
    const SlabGroup& sg = p->groups_host[g];
    pieces.push_back({g, (int64_t)sg.slab_begin, (int64_t)sg.n_slabs, slab_cost(p, sg)});
  }
  if (!p->cta_ns) {
    if (cudaMalloc(&p->cta_ns, sizeof(unsigned int) * (size_t)std::max(p->n_ctas, 1)) != cudaSuccess) {
      set_error("allocating the per-CTA timers failed");
      return DUALIP_ECUDA;
    }
    cudaMemset(p->cta_ns, 0, sizeof(unsigned int) * (size_t)std::max(p->n_ctas, 1));
  }
  return upload_ranges(p, cut_ranges(p, pieces), 0);
}

// Every CTA's contiguous share of the (length-sorted) mid columns, cut to equal cost: a column costs its length plus a fixed
// part (two warp-wide passes over its registers per search round, the reductions, the column header).
static int build_mid_ranges(dualip_plan* p) {
  if (p->n_mid <= 0 || p->mid_separate) return DUALIP_OK;
  std::vector<LongCol> lc((size_t)p->n_mid);
  if (cudaMemcpy(lc.data(), p->longcols, sizeof(LongCol) * lc.size(), cudaMemcpyDeviceToHost) != cudaSuccess) {
    set_error("reading the mid-column table failed");
    return DUALIP_ECUDA;
  }
  auto cost = [](const LongCol& c) { return (double)c.len + 48.0; };
  double total = 0.0;
  for (const LongCol& c : lc) total += cost(c);
  std::vector<int> r((size_t)p->n_ctas + 1, 0);
  double acc = 0.0;
  size_t i = 0;
  for (int c = 0; c < p->n_ctas; ++c) {
    r[c] = (int)i;
    const double target = total * (double)(c + 1) / (double)p->n_ctas;
    while (i < lc.size() && acc + 0.5 * cost(lc[i]) < target) acc += cost(lc[i++]);
  }
  r[p->n_ctas] = (int)lc.size();
  if (cudaMalloc(&p->mid_range, sizeof(int) * r.size()) != cudaSuccess ||
      cudaMemcpy(p->mid_range, r.data(), sizeof(int) * r.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
    set_error("uploading the mid-column ranges failed");
    return DUALIP_ECUDA;
  }
  return DUALIP_OK;
}

// Chooses the accumulation mode of a plan.  Fixed point needs (a) a finite bound on x for every class, (b) the register /
// shared-memory configuration it is built for (uint16 rows, lambda + accumulator in shared memory), and (c) enough
// resolution: with quantum q = 2^-F the rounding error of a row sum of N terms is ~ q*sqrt(N/12); it must stay below a
// few fp32 ulps of the row's largest possible sum, otherwise (heavy-tailed row bounds) the plan keeps fp32 atomics.
// keep_bits < 0: plan construction.  keep_bits >= 0: re-check after the CTA ranges moved; F stays at keep_bits if it still fits.
static int choose_accumulator(dualip_plan* p, cudaStream_t stream, int keep_bits = -1) {
  p->fixed_point = 0;
  p->fx_bounded = false;
  const char* env = getenv("DUALIP_ACCUM");
  if (env && strcmp(env, "f32") == 0) return DUALIP_OK;
  if (p->row_bits != 16 || p->smode != 0) return DUALIP_OK;
  std::vector<float> xmax(p->n_classes);
  for (int i = 0; i < p->n_classes; ++i) {
    const dualip_proj_class& pc = p->classes_host[i];
    xmax[i] = (pc.kind == DUALIP_PROJ_CLAMP) ? fmaxf(fabsf(pc.lo), fabsf(pc.hi)) : pc.z;
    if (pc.kind >= DUALIP_PROJ_SIMPLEX_BISECT) xmax[i] = INFINITY;  // (v - max(v/z) + 1) * z: no bound for z != 1
    if (!p->class_used[i]) xmax[i] = 0.f;
    if (!(xmax[i] < INFINITY)) return DUALIP_OK;  // open cone / identity: no bound
  }
  const int m = p->m;
  float *xmax_d = nullptr, *table = nullptr, *row_max = nullptr, *row_sum = nullptr, *long_bound = nullptr;
  unsigned int* row_cnt = nullptr;
  auto cleanup = [&]() {
    cudaFree(xmax_d);
    cudaFree(table);
    cudaFree(row_max);
    cudaFree(row_sum);
    cudaFree(long_bound);
    cudaFree(row_cnt);
  };
#define CA_TRY(expr)                                               \
  do {                                                             \
    cudaError_t _e = (expr);                                       \
    if (_e != cudaSuccess) {                                       \
      cleanup();                                                   \
      if (_e == cudaErrorMemoryAllocation) {                       \
        cudaGetLastError();                                        \
        return DUALIP_OK; /* no room for the table: fp32 mode */   \
      }                                                            \
      set_error("%s failed: %s", #expr, cudaGetErrorString(_e));   \
      return DUALIP_ECUDA;                                         \
    }                                                              \
  } while (0)
  const size_t tab = (size_t)p->n_ctas * m;
  CA_TRY(cudaMalloc(&xmax_d, sizeof(float) * p->n_classes));
  CA_TRY(cudaMalloc(&table, sizeof(float) * tab));
  CA_TRY(cudaMalloc(&row_max, sizeof(float) * m));
  CA_TRY(cudaMalloc(&row_sum, sizeof(float) * m));
  CA_TRY(cudaMalloc(&long_bound, sizeof(float) * m));
  CA_TRY(cudaMalloc(&row_cnt, sizeof(unsigned int) * m));
  CA_TRY(cudaMemcpyAsync(xmax_d, xmax.data(), sizeof(float) * p->n_classes, cudaMemcpyHostToDevice, stream));
  CA_TRY(cudaMemsetAsync(table, 0, sizeof(float) * tab, stream));
  CA_TRY(cudaMemsetAsync(long_bound, 0, sizeof(float) * m, stream));
  CA_TRY(cudaMemsetAsync(row_cnt, 0, sizeof(unsigned int) * m, stream));
  if (p->n_slabs > 0 || (p->n_mid > 0 && !p->mid_separate)) {
    const size_t bsm = 8 * (size_t)m;  // fits: mode 0 already keeps 8*m bytes of lambda + accumulator in shared memory
    CA_TRY(cudaFuncSetAttribute((const void*)cta_row_bound_kernel<unsigned short>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bsm));
    cta_row_bound_kernel<unsigned short><<<p->n_ctas, p->threads, bsm, stream>>>(
        p->data, p->hdr, p->cta_range, p->n_slabs, xmax_d, m, table, row_cnt,
        (p->n_mid > 0 && !p->mid_separate) ? p->longcols : nullptr, p->mid_range, p->long_a, p->long_row);
  }
  {
    // columns whose sums go straight to the global accumulators: the long ones, and the mid ones when they have their own launch
    const int64_t first = p->mid_separate ? 0 : p->n_mid, nl = p->n_long - first;
    if (nl > 0) {
      const int blocks = (int)std::min<int64_t>((nl + 7) / 8, (int64_t)p->n_sms * 8);
      long_row_bound_kernel<<<blocks, 256, 0, stream>>>(p->longcols + first, nl, p->long_a, p->long_row, xmax_d, long_bound, row_cnt);
    }
  }
  bound_reduce_kernel<<<(m + 255) / 256, 256, 0, stream>>>(table, p->n_ctas, m, row_max, row_sum);
  std::vector<float> h_max(m), h_sum(m), h_long(m);
  std::vector<unsigned int> h_cnt(m);
  CA_TRY(cudaMemcpyAsync(h_max.data(), row_max, sizeof(float) * m, cudaMemcpyDeviceToHost, stream));
  CA_TRY(cudaMemcpyAsync(h_sum.data(), row_sum, sizeof(float) * m, cudaMemcpyDeviceToHost, stream));
  CA_TRY(cudaMemcpyAsync(h_long.data(), long_bound, sizeof(float) * m, cudaMemcpyDeviceToHost, stream));
  CA_TRY(cudaMemcpyAsync(h_cnt.data(), row_cnt, sizeof(unsigned int) * m, cudaMemcpyDeviceToHost, stream));
  CA_TRY(cudaStreamSynchronize(stream));
  CA_TRY(cudaGetLastError());
  cleanup();
#undef CA_TRY
  double bmax = 0.0, total_max = 0.0;
  p->row_total_host.assign((size_t)m, 0.f);
  for (int r = 0; r < m; ++r) {
    bmax = std::max(bmax, (double)h_max[r]);
    total_max = std::max(total_max, (double)h_sum[r] + (double)h_long[r]);
    p->row_total_host[r] = h_sum[r] + h_long[r];
  }
  if (!(bmax < INFINITY) || !(total_max < INFINITY)) return DUALIP_OK;
  p->fx_bounded = true;
  // the float table itself carries rounding error: 1.001 covers it.  B * 2^F <= 2^30 leaves 2^30 of headroom for the
  // +-1/2 per term of the integer rounding (up to 2^31 terms).
  // One more bit is left free at plan time so that dualip_plan_rebalance can move the CTA ranges without changing F (the
  // sums stay bit-identical across rebalances); a rebalance lowers F only if a range's bound grew past that.
  int F = 20;
  if (bmax > 0.0) F = (int)floor(log2(1073741824.0 / (bmax * 1.001))) - (keep_bits < 0 ? 1 : 0);
  if (keep_bits >= 0 && F > keep_bits) F = keep_bits;
  F = std::max(-64, std::min(64, F));
  const double q = ldexp(1.0, -F);
  if (total_max * 1.001 * ldexp(1.0, F) >= 7.0e13) return DUALIP_OK;  // grid-wide high parts must fit 32 bits (2^46 = 7.04e13)
  double relerr = 0.0;
  for (int r = 0; r < m; ++r) {
    const double A = (double)h_sum[r] + (double)h_long[r];
    if (h_cnt[r] > 0 && A > 0.0) relerr = std::max(relerr, q * sqrt((double)h_cnt[r] / 12.0) / A);
  }
  p->fx_bits = F;
  p->fx_scale = (float)ldexp(1.0, F);
  p->fx_inv = q;
  p->fx_bound = bmax;
  p->fx_relerr = relerr;
  const bool forced = env && strcmp(env, "fixed") == 0;
  if (relerr <= 2.4e-7 || forced) p->fixed_point = 1;
  return DUALIP_OK;
}

// ------------------------------------------------------------------------------------------
// Power-of-two row equilibration (plan time).  The fixed-point accumulator has ONE quantum 2^-F for all rows, sized by the
// largest row sum a CTA can see; rows whose entries are orders of magnitude smaller (no Jacobi scaling, heavy-tailed row
// scales) would lose relative precision and the plan used to fall back to fp32 compare-and-swap atomics.  Instead the plan
// stores row r as a * 2^k_r, k_r = -floor(log2(sum_r |a| xmax)), and hands the kernel lambda' = fl(s*lambda_r) * 2^-k_r:
// fl(a*2^k * lambda') == fl(a * fl(s*lambda_r)) bit for bit (a power of two commutes with rounding; |k_r| <= 40 keeps every
// factor normal for data within 2^+-40 of 1), so v, the projection and x are unchanged, while the scatter adds
// fl(a*x) * 2^(k_r + F): every row gets its own quantum.  The m-length tail multiplies the row sums by 2^-k_r (exact).
// ------------------------------------------------------------------------------------------
template <typename RowT>
__global__ void scale_slab_rows_kernel(unsigned char* __restrict__ data, const SlabHdr* __restrict__ hdr, int64_t n_slabs,
                                       const float* __restrict__ row_scale) {
  const int lane = threadIdx.x & 31;
  int64_t sl = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (; sl < n_slabs; sl += nw) {
    const SlabHdr h = hdr[sl];
    float* a_t = reinterpret_cast<float*>(data + (size_t)h.off32 * (size_t)row32_bytes(8 * (int)sizeof(RowT)));
    const RowT* row_t = reinterpret_cast<const RowT*>(a_t + 2 * (size_t)h.d * kSlabW);
    for (int i = lane; i < (int)h.d * kSlabW; i += 32) a_t[i] = __fmul_rn(a_t[i], row_scale[(uint32_t)row_t[i]]);  // padding: a = 0, row 0
  }
}
__global__ void scale_long_rows_kernel(float* __restrict__ la, const uint32_t* __restrict__ lrow, int64_t total,
                                       const float* __restrict__ row_scale) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i < total; i += (int64_t)gridDim.x * blockDim.x) la[i] = __fmul_rn(la[i], row_scale[lrow[i]]);
}

// Called when the single-quantum accumulator lacks resolution: scales the stored rows and re-runs the choice.
static int equilibrate_rows(dualip_plan* p, cudaStream_t stream) {
  const int m = p->m;
  std::vector<float> scale((size_t)m, 1.f), unscale((size_t)m, 1.f);
  bool any = false;
  for (int r = 0; r < m; ++r) {
    const float A = p->row_total_host[r];
    if (!(A > 0.f) || !(A < INFINITY)) continue;
    int e = 0;
    frexpf(A, &e);  // A = f * 2^e, f in [0.5, 1)
    const int kr = std::max(-40, std::min(40, 1 - e));  // A * 2^kr in [1, 2)
    if (kr != 0) any = true;
    scale[r] = ldexpf(1.f, kr);
    unscale[r] = ldexpf(1.f, -kr);
  }
  if (!any) return DUALIP_OK;
  float* scale_d = nullptr;
  if (cudaMalloc(&scale_d, sizeof(float) * m) != cudaSuccess || cudaMalloc(&p->row_unscale, sizeof(float) * m) != cudaSuccess) {
    cudaFree(scale_d);
    cudaFree(p->row_unscale);
    p->row_unscale = nullptr;
    cudaGetLastError();
    return DUALIP_OK;  // no room: the plan keeps its fp32 accumulator
  }
  cudaMemcpyAsync(scale_d, scale.data(), sizeof(float) * m, cudaMemcpyHostToDevice, stream);
  cudaMemcpyAsync(p->row_unscale, unscale.data(), sizeof(float) * m, cudaMemcpyHostToDevice, stream);
  if (p->n_slabs > 0) {
    const int blocks = (int)std::min<int64_t>((p->n_slabs + 7) / 8, (int64_t)p->n_sms * 16);
    scale_slab_rows_kernel<unsigned short><<<blocks, 256, 0, stream>>>(p->data, p->hdr, p->n_slabs, scale_d);
  }
  if (p->n_long > 0 && p->long_total > 0)
    scale_long_rows_kernel<<<p->n_sms * 8, 256, 0, stream>>>(p->long_a, p->long_row, p->long_total, scale_d);
  cudaError_t e = cudaStreamSynchronize(stream);  // the host vectors are read by the copies above
  cudaFree(scale_d);
  if (e != cudaSuccess || (e = cudaGetLastError()) != cudaSuccess) {
    set_error("row equilibration failed: %s", cudaGetErrorString(e));
    return DUALIP_ECUDA;
  }
  return choose_accumulator(p, stream);
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
  cudaFree(p->data);
  cudaFree(p->groups);
  cudaFree(p->cta_range);
  cudaFree(p->hdr);
  cudaFree(p->orig_start);
  cudaFree(p->longcols);
  cudaFree(p->mid_range);
  cudaFree(p->long_a);
  cudaFree(p->long_c);
  cudaFree(p->long_row);
  cudaFree(p->classes_dev);
  cudaFree(p->pad_dev);
  cudaFree(p->row_unscale);
  cudaFree(p->acc);
  cudaFree(p->acc_lo);
  cudaFree(p->timeline);
  cudaFree(p->acc_hi);
  cudaFree(p->acc_scal);
  cudaFree(p->counter);
  cudaFree(p->grid_bar);
  cudaFree(p->tail_part);
  cudaFree(p->grid_status);
  if (p->grid_status_host) cudaFreeHost(p->grid_status_host);
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
  if (d->n_cols >= (1LL << 31)) {
    set_error("more than 2^31-1 columns per shard is not supported by this build; shard the columns");
    return DUALIP_ERANGE;
  }
  if (!d->ccol_dev || (d->nnz > 0 && (!d->row_dev || !d->a_dev || !d->c_dev))) {
    set_error("null CSC array");
    return DUALIP_EINVAL;
  }
  for (int i = 0; i < d->n_classes; ++i) {
    const dualip_proj_class& pc = d->classes[i];
    if (pc.kind < DUALIP_PROJ_CLAMP || pc.kind > DUALIP_PROJ_SIMPLEX_EQ_BISECT) {
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
  p->n_classes = d->n_classes;
  memcpy(p->classes_host, d->classes, sizeof(dualip_proj_class) * d->n_classes);
  const char* fb = getenv("DUALIP_FLUSH");
  p->flush_bulk = (fb && strcmp(fb, "atomic") == 0) ? 0 : 1;
  if (getenv("DUALIP_TIMELINE")) {
    if (cudaMalloc(&p->timeline, sizeof(unsigned long long) * 12 * 4096) != cudaSuccess) p->timeline = nullptr;
    if (p->timeline) cudaMemset(p->timeline, 0, sizeof(unsigned long long) * 12 * 4096);
  }

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

  // shared-memory mode and CTA shape: ONE CTA per SM.  lambda and the gradient accumulator are per CTA, so a single
  // CTA halves their footprint (and the flush traffic) against two smaller CTAs; 16 warps with all loads of a slab in
  // flight at once (plus the L2 prefetch of the next slab) cover the HBM latency.
  p->threads = kThreads;
  p->n_ctas = p->n_sms;
  const size_t fixed = smem_fixed_bytes(p->n_classes);
  const size_t m_pad = ((size_t)p->m + 3) & ~(size_t)3;
  const size_t stash_bytes = (size_t)(p->threads / 32) * kStashDeg * kSlabW * sizeof(float);
  // TMA staging of the register-path slabs: every warp gets an equal share of the shared memory that lambda and the
  // accumulator leave free, cut into 4 / 2 / 1 slots by column length (see the kernel).  DUALIP_STAGE=0 disables it.
  const char* env_stage = getenv("DUALIP_STAGE");
  int stage_deg = env_stage ? atoi(env_stage) : kRegDeg;
  stage_deg = std::max(0, std::min(stage_deg, kRegDeg));
  p->row_bits = (p->m <= 65536) ? 16 : 32;
  const size_t smem_max = std::min<size_t>(kSmemBudget, prop.sharedMemPerBlockOptin) - 4096;  // static shared memory (block reductions) + slack
  const size_t need1 = fixed + 4 * m_pad + stash_bytes;
  const size_t n_warps = (size_t)(p->threads / 32);
  size_t region = 0;
  if (stage_deg > 0 && smem_max > fixed + 8 * m_pad + 128)
    region = std::min<size_t>(((smem_max - fixed - 8 * m_pad - 128) / n_warps) & ~(size_t)511, ((size_t)2 * 320 * kRegDeg + 511) & ~(size_t)511);
  if (const char* er = getenv("DUALIP_STAGE_REGION")) region = std::min(region, (size_t)atoi(er) & ~(size_t)511);
  // mode 0 (register path): lambda + accumulator, plus the per-warp TMA staging regions when at least one slab fits
  if (p->row_bits == 16 && region >= 320 && fixed + 8 * m_pad + 128 + n_warps * region <= smem_max) {
    p->smode = 0;
    p->stage = stage_deg;
    p->stage_region = (int)region;
    p->smem_bytes = fixed + 8 * m_pad + 128 + n_warps * region;
  } else if (p->row_bits == 16 && fixed + 8 * m_pad + 128 <= smem_max) {
    p->smode = 0;
    p->stage = 0;
    p->smem_bytes = fixed + 8 * m_pad + 128;
  } else if (need1 <= smem_max) {
    p->smode = 1;
    p->smem_bytes = need1;
  } else {
    p->smode = 2;
    p->smem_bytes = fixed + stash_bytes;
  }
  if (const char* pad = getenv("DUALIP_SMEM_PAD")) {  // experiments: shrink L1 by requesting unused shared memory
    p->smem_bytes = std::min(smem_max, p->smem_bytes + (size_t)atoi(pad));
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

  {
    int rc = (d->index_bits == 64) ? build_slabs<long long>(p, d, stream) : build_slabs<int>(p, d, stream);
    if (rc != DUALIP_OK) return fail(rc);
  }
  // do not launch more warps than there are slabs (tiny problems)
  {
    const int64_t warps_per_cta = p->threads / 32;
    const int64_t want = std::max<int64_t>(1, std::max((p->n_slabs + warps_per_cta - 1) / warps_per_cta,
                                                       ((p->mid_separate ? 0 : p->n_mid) + warps_per_cta - 1) / warps_per_cta));
    if (!(env_ctas && atoi(env_ctas) > 0) && want < p->n_ctas) p->n_ctas = (int)want;
  }
  {
    int rc = build_cta_ranges(p);
    if (rc != DUALIP_OK) return fail(rc);
    rc = build_mid_ranges(p);
    if (rc != DUALIP_OK) return fail(rc);
  }
  {
    int rc = choose_accumulator(p, stream);
    if (rc != DUALIP_OK) return fail(rc);
    const char* ers = getenv("DUALIP_ROW_SCALE");
    if (!p->fixed_point && p->fx_bounded && !(ers && strcmp(ers, "0") == 0)) {
      rc = equilibrate_rows(p, stream);  // one quantum does not resolve every row: per-row powers of two
      if (rc != DUALIP_OK) return fail(rc);
    }
  }
  // small state
  DUALIP_TRY_FAIL(cudaMalloc(&p->acc_lo, sizeof(int) * (m_pad + 4)));
  DUALIP_TRY_FAIL(cudaMemset(p->acc_lo, 0, sizeof(int) * (m_pad + 4)));
  DUALIP_TRY_FAIL(cudaMalloc(&p->acc_hi, sizeof(int) * (m_pad + 4)));
  DUALIP_TRY_FAIL(cudaMemset(p->acc_hi, 0, sizeof(int) * (m_pad + 4)));
  DUALIP_TRY_FAIL(cudaMalloc(&p->classes_dev, sizeof(dualip_proj_class) * p->n_classes));
  DUALIP_TRY_FAIL(cudaMemcpy(p->classes_dev, p->classes_host, sizeof(dualip_proj_class) * p->n_classes, cudaMemcpyHostToDevice));
  if (d->pad_len != nullptr) {
    bool any_eq = false;
    for (int i = 0; i < p->n_classes; ++i) any_eq = any_eq || p->classes_host[i].kind >= DUALIP_PROJ_SIMPLEX_EQ;
    if (any_eq) {
      const size_t bytes = sizeof(int) * (size_t)p->n_classes * DUALIP_PAD_BUCKETS;
      DUALIP_TRY_FAIL(cudaMalloc(&p->pad_dev, bytes));
      DUALIP_TRY_FAIL(cudaMemcpy(p->pad_dev, d->pad_len, bytes, cudaMemcpyHostToDevice));
    }
  }
  DUALIP_TRY_FAIL(cudaMalloc(&p->acc, sizeof(float) * (m_pad + 4)));
  DUALIP_TRY_FAIL(cudaMemset(p->acc, 0, sizeof(float) * (m_pad + 4)));
  DUALIP_TRY_FAIL(cudaMalloc(&p->acc_scal, sizeof(double) * 2));
  DUALIP_TRY_FAIL(cudaMemset(p->acc_scal, 0, sizeof(double) * 2));
  DUALIP_TRY_FAIL(cudaMalloc(&p->counter, sizeof(unsigned int)));
  DUALIP_TRY_FAIL(cudaMemset(p->counter, 0, sizeof(unsigned int)));
  DUALIP_TRY_FAIL(cudaMalloc(&p->grid_bar, 2 * sizeof(unsigned int)));
  DUALIP_TRY_FAIL(cudaMemset(p->grid_bar, 0, 2 * sizeof(unsigned int)));
  DUALIP_TRY_FAIL(cudaMalloc(&p->tail_part, sizeof(double) * kTailPart * (size_t)std::max(p->n_ctas, 1)));
  DUALIP_TRY_FAIL(cudaMalloc(&p->grid_status, sizeof(int)));
  DUALIP_TRY_FAIL(cudaMemset(p->grid_status, 0, sizeof(int)));
  DUALIP_TRY_FAIL(cudaHostAlloc(reinterpret_cast<void**>(&p->grid_status_host), sizeof(int), cudaHostAllocMapped));
  *p->grid_status_host = 0;
  DUALIP_TRY_FAIL(cudaHostGetDevicePointer(reinterpret_cast<void**>(&p->grid_status_host_dev), p->grid_status_host, 0));
  // the all-CTA tail pays two to four grid barriers (~3 us each): a gain once the m-length passes of a single CTA cost more than
  // that (measured: +17 % at m = 26 744, neutral at m = 10 000 on one and two GPUs, -9 % at m = 1 000); DUALIP_GRID_TAIL=0|1
  // forces it
  p->grid_tail = p->m >= 16384 ? 1 : 0;
  p->grid_tail_auto = 1;
  if (const char* gt = getenv("DUALIP_GRID_TAIL")) {
    p->grid_tail = atoi(gt) != 0 ? 1 : 0;
    p->grid_tail_auto = 0;
  }
  DUALIP_TRY_FAIL(cudaMalloc(&p->lambda_stage, sizeof(float) * (m_pad + 4)));
  DUALIP_TRY_FAIL(cudaMalloc(&p->grad_stage, sizeof(float) * (m_pad + 4)));
  DUALIP_TRY_FAIL(cudaMalloc(&p->scal_stage, sizeof(dualip_scalars)));
  p->owned_bytes += sizeof(float) * 3 * (m_pad + 4) + 64;

  // the attribute is per function, not per plan: always allow the device maximum so that plans of different m coexist
  for (int out = 0; out < 2; ++out) {
    SlabKernel kern = plan_kernel(p, out != 0);
    if (!kern) {
      set_error("no kernel variant");
      return fail(DUALIP_EINVAL);
    }
    cudaFuncAttributes fa;
    DUALIP_TRY_FAIL(cudaFuncGetAttributes(&fa, (const void*)kern));
    const int max_dyn = (int)prop.sharedMemPerBlockOptin - (int)fa.sharedSizeBytes;
    DUALIP_TRY_FAIL(cudaFuncSetAttribute((const void*)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dyn));
  }
  if (p->n_ctalong > 0) {
    DUALIP_TRY_FAIL(cudaFuncSetAttribute((const void*)matching_long_cta_kernel<0, kLongThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         kLongStash * (int)sizeof(float)));
    DUALIP_TRY_FAIL(cudaFuncSetAttribute((const void*)matching_long_cta_kernel<1, kLongThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         kLongStash * (int)sizeof(float)));
  }
  DUALIP_TRY_FAIL(cudaDeviceSynchronize());
#undef DUALIP_TRY_FAIL
  *out = p;
  return DUALIP_OK;
}

/* Debug only (not declared in the public header): copies the per-CTA timeline of the last launch, 10 values per CTA. */
int dualip_debug_timeline(dualip_plan* p, unsigned long long* out_host, int n_ctas) {
  if (!p || !p->timeline || !out_host) return DUALIP_EINVAL;
  DeviceGuard g(p->device);
  DUALIP_CUDA_TRY(cudaDeviceSynchronize());
  DUALIP_CUDA_TRY(cudaMemcpy(out_host, p->timeline, sizeof(unsigned long long) * (n_ctas < 0 ? 12 * 4096 : 10 * std::min(n_ctas, 4096)), cudaMemcpyDeviceToHost));
  return DUALIP_OK;
}

/* Debug only: group table (6 values per group: slab_begin, n_slabs, d, cls, off32_begin, last_ncols) and the CTA ranges
 * (first slab of every CTA, n_ctas + 1 values).  Returns the number of groups. */
int dualip_debug_layout(dualip_plan* p, int64_t* groups_out, int cap_groups, int64_t* ranges_out, int cap_ranges) {
  if (!p) return DUALIP_EINVAL;
  DeviceGuard g(p->device);
  const int G = (int)p->groups_host.size();
  for (int i = 0; i < G && i < cap_groups; ++i) {
    const SlabGroup& sg = p->groups_host[i];
    int64_t* o = groups_out + 6 * i;
    o[0] = sg.slab_begin, o[1] = sg.n_slabs, o[2] = sg.d, o[3] = sg.cls, o[4] = sg.off32_begin, o[5] = sg.last_ncols;
  }
  std::vector<int2> r((size_t)p->n_ctas + 1);
  if (p->cta_range) cudaMemcpy(r.data(), p->cta_range, sizeof(int2) * r.size(), cudaMemcpyDeviceToHost);
  for (int i = 0; i <= p->n_ctas && i < cap_ranges; ++i) ranges_out[i] = r[i].x;
  return G;
}

int dualip_plan_rebalance(dualip_plan* p, void* stream_v) {
  if (!p) {
    set_error("null argument");
    return DUALIP_EINVAL;
  }
  if (p->n_ctas < 2 || p->n_slabs == 0 || !p->cta_ns || p->ranges_host.empty()) return DUALIP_OK;
  DeviceGuard g(p->device);
  cudaStream_t stream = (cudaStream_t)stream_v;
  std::vector<unsigned int> ns((size_t)p->n_ctas);
  DUALIP_CUDA_TRY(cudaStreamSynchronize(stream));
  DUALIP_CUDA_TRY(cudaMemcpy(ns.data(), p->cta_ns, sizeof(unsigned int) * ns.size(), cudaMemcpyDeviceToHost));
  // measured time per unit of table cost in every current range; ranges the table got right have the same ratio
  const std::vector<int2> old = p->ranges_host;
  const int G = (int)p->groups_host.size();
  std::vector<CostPiece> pieces;
  std::vector<double> model((size_t)p->n_ctas, 0.0);
  std::vector<int> owner;
  for (int c = 0; c < p->n_ctas; ++c) {
    const int64_t a = old[c].x, b = old[c + 1].x;
    for (int gi = old[c].y; gi < G && (int64_t)p->groups_host[gi].slab_begin < b; ++gi) {
      const SlabGroup& sg = p->groups_host[gi];
      const int64_t lo = std::max<int64_t>(a, sg.slab_begin), hi = std::min<int64_t>(b, (int64_t)sg.slab_begin + sg.n_slabs);
      if (hi <= lo) continue;
      const double cs = slab_cost(p, sg);
      pieces.push_back({gi, lo, hi - lo, cs});
      owner.push_back(c);
      model[c] += cs * (double)(hi - lo);
    }
  }
  double sum_t = 0.0, sum_m = 0.0;
  for (int c = 0; c < p->n_ctas; ++c) {
    if (ns[c] == 0 && model[c] > 0.0) return DUALIP_OK;  // no launch since the last upload: nothing measured
    sum_t += (double)ns[c];
    sum_m += model[c];
  }
  if (!(sum_t > 0.0) || !(sum_m > 0.0)) return DUALIP_OK;
  const double mean_scale = sum_t / sum_m;
  for (size_t i = 0; i < pieces.size(); ++i) {
    const int c = owner[i];
    double rel = model[c] > 0.0 ? ((double)ns[c] / model[c]) / mean_scale : 1.0;
    rel = std::min(4.0, std::max(0.25, rel));
    pieces[i].cost *= pow(rel, 0.85);  // damped: moving a boundary changes what the neighbours contend for
  }
  const std::vector<int2> fresh = cut_ranges(p, pieces);
  const int was_fixed = p->fixed_point, old_bits = p->fx_bits;
  const float old_scale = p->fx_scale;
  const double old_inv = p->fx_inv, old_bound = p->fx_bound, old_relerr = p->fx_relerr;
  int rc = upload_ranges(p, fresh, stream);
  if (rc != DUALIP_OK) return rc;
  if (was_fixed) {
    rc = choose_accumulator(p, stream, old_bits);
    if (rc != DUALIP_OK || !p->fixed_point) {  // the new ranges do not admit the fixed-point sums: keep the old partition
      p->fixed_point = was_fixed, p->fx_bits = old_bits, p->fx_scale = old_scale, p->fx_inv = old_inv;
      p->fx_bound = old_bound, p->fx_relerr = old_relerr;
      return upload_ranges(p, old, stream);
    }
  }
  cudaMemsetAsync(p->cta_ns, 0, sizeof(unsigned int) * ns.size(), stream);
  ++p->rebalances;
  return DUALIP_OK;
}

int dualip_plan_info(const dualip_plan* p, int64_t* out, int cap) {
  if (!p || !out) {
    set_error("null argument");
    return DUALIP_EINVAL;
  }
  const int64_t v[21] = {p->n_slabs, p->n_long - p->n_mid, p->n_ctas, p->threads, (int64_t)p->smem_bytes, p->row_bits,
                         p->smode,   p->rows32 * kSlabW, 1 + (p->n_ctalong > 0 ? 1 : 0) + (p->n_long > p->n_mid + p->n_ctalong ? 1 : 0) + ((p->mid_separate && p->n_mid > 0) ? 1 : 0), (int64_t)p->owned_bytes, p->n_short, p->nnz,
                         p->fixed_point, p->fx_bits, (int64_t)(p->fx_relerr * 1e12), p->stage, p->row_unscale ? 1 : 0, p->n_mid,
                         (p->grid_tail && p->n_ctas <= p->n_sms) ? 1 : 0,
                         p->grid_status_host ? (int64_t)*reinterpret_cast<volatile int*>(p->grid_status_host) : 0,
                         p->last_grid_tail};
  for (int i = 0; i < cap && i < 21; ++i) out[i] = v[i];
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

int dualip_matching_ascent_step(dualip_plan* p, dualip_agd* a, const float* b_dev, double gamma, float* grad_out_dev,
                                dualip_scalars* scalars_out_dev, float* x_out_dev, float beta, int32_t decay_now,
                                double decay_factor, int32_t iter_index, void* stream) {
  if (!p || !a || !grad_out_dev || !scalars_out_dev) {
    set_error("null argument");
    return DUALIP_EINVAL;
  }
  if (a->m != p->m || a->device != p->device) {
    set_error("optimizer state does not match the plan (m or device)");
    return DUALIP_EINVAL;
  }
  DeviceGuard g(p->device);
  FuseSpec f;
  f.mode = 1;
  f.agd = step_args(a, grad_out_dev, scalars_out_dev, beta, decay_now, decay_factor, iter_index, nullptr, 0.0, nullptr, nullptr);
  return launch_eval(p, a->x, b_dev, gamma, grad_out_dev, scalars_out_dev, nullptr, x_out_dev, nullptr, 1, (cudaStream_t)stream, &f);
}

int dualip_matching_ascent_step_peer(dualip_plan* p, dualip_agd* a, dualip_peer* peer, const float* b_dev, double gamma,
                                     float* grad_out_dev, dualip_scalars* scalars_out_dev, float beta, int32_t decay_now,
                                     double decay_factor, int32_t iter_index, void* stream) {
  if (!p || !a || !peer || !grad_out_dev || !scalars_out_dev) {
    set_error("null argument");
    return DUALIP_EINVAL;
  }
  if (a->m != p->m || a->device != p->device || !peer->connected || peer->m != a->m || peer->device != a->device) {
    set_error("optimizer state / exchange window do not match the plan, or the window is not connected");
    return DUALIP_EINVAL;
  }
  DeviceGuard g(p->device);
  FuseSpec f;
  f.mode = 2;
  f.peer = peer_args(peer, true);
  float* slot = reinterpret_cast<float*>(peer->window + kPeerFlagBytes + (size_t)(f.peer.seq & 1ull) * peer->slot_bytes);
  f.agd = step_args(a, peer->sum, nullptr, beta, decay_now, decay_factor, iter_index, b_dev, gamma, grad_out_dev, scalars_out_dev);
  return launch_eval(p, a->x, nullptr, gamma, nullptr, nullptr, slot, nullptr, nullptr, 0, (cudaStream_t)stream, &f);
}

// Scheduled launch: dualip_matching_ascent_step[_peer] with gamma / beta / decay / log slot / exchange step number taken
// from the schedule installed by dualip_agd_set_schedule at the device-side iteration count.  The kernel arguments do not
// depend on the iteration, so the launch can be captured once and replayed.
static int launch_scheduled(dualip_plan* p, dualip_agd* a, dualip_peer* peer, const float* b_dev, float* grad_out_dev,
                            dualip_scalars* scalars_out_dev, cudaStream_t stream) {
  FuseSpec f;
  f.sched.gamma = a->sched_gamma;
  f.sched.beta = a->sched_beta;
  f.sched.decay = a->sched_decay;
  f.sched.n = a->sched_n;
  f.sched.seq_base = 0;
  if (peer) {
    f.mode = 2;
    f.sched.seq_base = (long long)peer->seq - a->launched;  // invariant: both advance by one per step
    f.peer = peer_args(peer, true);
    f.agd = step_args(a, peer->sum, nullptr, 0.f, 0, a->sched_factor, 0, b_dev, 1.0, grad_out_dev, scalars_out_dev);
  } else {
    f.mode = 1;
    f.agd = step_args(a, grad_out_dev, scalars_out_dev, 0.f, 0, a->sched_factor, 0, nullptr, 0.0, nullptr, nullptr);
  }
  f.agd.log_obj = a->log_obj;  // the kernel checks the iteration index against log_cap
  f.agd.log_step = a->log_step;
  // gamma = 1.0 below is a placeholder that passes launch_eval's argument check; scheduled kernels never read it
  return launch_eval(p, a->x, peer ? nullptr : b_dev, 1.0, peer ? nullptr : grad_out_dev, peer ? nullptr : scalars_out_dev,
                     nullptr, nullptr, nullptr, peer ? 0 : 1, stream, &f);
}

static int check_scheduled(dualip_plan* p, dualip_agd* a, dualip_peer* peer, float* grad_out_dev, dualip_scalars* scalars_out_dev) {
  if (!p || !a || !grad_out_dev || !scalars_out_dev) {
    set_error("null argument");
    return DUALIP_EINVAL;
  }
  if (a->m != p->m || a->device != p->device) {
    set_error("optimizer state does not match the plan (m or device)");
    return DUALIP_EINVAL;
  }
  if (a->sched_n <= 0) {
    set_error("no schedule installed (dualip_agd_set_schedule)");
    return DUALIP_EINVAL;
  }
  if (peer && (!peer->connected || peer->m != a->m || peer->device != a->device)) {
    set_error("exchange window does not match the plan, or is not connected");
    return DUALIP_EINVAL;
  }
  return DUALIP_OK;
}

int dualip_matching_ascent_step_scheduled(dualip_plan* p, dualip_agd* a, dualip_peer* peer, const float* b_dev,
                                          float* grad_out_dev, dualip_scalars* scalars_out_dev, void* stream) {
  int rc = check_scheduled(p, a, peer, grad_out_dev, scalars_out_dev);
  if (rc != DUALIP_OK) return rc;
  if (a->launched >= (long long)a->sched_n) {
    set_error("the schedule holds %d iterations and all of them have been launched", a->sched_n);
    return DUALIP_ERANGE;
  }
  DeviceGuard g(p->device);
  return launch_scheduled(p, a, peer, b_dev, grad_out_dev, scalars_out_dev, (cudaStream_t)stream);
}

}  // extern "C"

struct dualip_ascent_graph {
  int device = 0;
  int chunk = 0;
  dualip_agd* agd = nullptr;
  dualip_peer* peer = nullptr;
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
};

extern "C" {

void dualip_ascent_graph_destroy(dualip_ascent_graph* g) {
  if (!g) return;
  DeviceGuard dg(g->device);
  if (g->exec) cudaGraphExecDestroy(g->exec);
  if (g->graph) cudaGraphDestroy(g->graph);
  delete g;
}

int dualip_ascent_graph_create(dualip_ascent_graph** out, dualip_plan* p, dualip_agd* a, dualip_peer* peer, const float* b_dev,
                               float* grad_out_dev, dualip_scalars* scalars_out_dev, int32_t chunk) {
  if (!out || chunk <= 0) {
    set_error("bad argument");
    return DUALIP_EINVAL;
  }
  *out = nullptr;
  int rc = check_scheduled(p, a, peer, grad_out_dev, scalars_out_dev);
  if (rc != DUALIP_OK) return rc;
  DeviceGuard dg(p->device);
  dualip_ascent_graph* g = new (std::nothrow) dualip_ascent_graph();
  if (!g) return DUALIP_ENOMEM;
  g->device = p->device;
  g->chunk = chunk;
  g->agd = a;
  g->peer = peer;
  // capture on a private stream (the caller's may be the legacy default stream, which cannot be captured); the captured
  // launches are not executed, so the host-side step counters are restored afterwards
  cudaStream_t cs = nullptr;
  const long long launched0 = a->launched;
  const unsigned long long seq0 = peer ? peer->seq : 0ull;
  cudaError_t e = cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal);
  if (e != cudaSuccess) {
    set_error("starting the stream capture failed: %s", cudaGetErrorString(e));
    if (cs) cudaStreamDestroy(cs);
    delete g;
    return DUALIP_ECUDA;
  }
  for (int i = 0; i < chunk && rc == DUALIP_OK; ++i) {
    rc = launch_scheduled(p, a, peer, b_dev, grad_out_dev, scalars_out_dev, cs);
    // every captured launch must carry the SAME arguments: undo the counters so that seq_base / slot stay what they are
    a->launched = launched0;
    if (peer) peer->seq = seq0;
  }
  e = cudaStreamEndCapture(cs, &g->graph);
  if (rc == DUALIP_OK && e != cudaSuccess) {
    set_error("ending the stream capture failed: %s", cudaGetErrorString(e));
    rc = DUALIP_ECUDA;
  }
  if (rc == DUALIP_OK) {
    e = cudaGraphInstantiate(&g->exec, g->graph, 0);
    if (e != cudaSuccess) {
      set_error("cudaGraphInstantiate failed: %s", cudaGetErrorString(e));
      rc = DUALIP_ECUDA;
    }
  }
  cudaStreamDestroy(cs);
  if (rc != DUALIP_OK) {
    cudaGetLastError();
    dualip_ascent_graph_destroy(g);
    return rc;
  }
  *out = g;
  return DUALIP_OK;
}

int dualip_ascent_graph_launch(dualip_ascent_graph* g, void* stream) {
  if (!g || !g->exec) {
    set_error("null argument");
    return DUALIP_EINVAL;
  }
  if (g->agd->launched + g->chunk > (long long)g->agd->sched_n) {
    set_error("the schedule holds %d iterations: %lld launched, the graph takes %d more", g->agd->sched_n, g->agd->launched,
              g->chunk);
    return DUALIP_ERANGE;
  }
  DeviceGuard dg(g->device);
  DUALIP_CUDA_TRY(cudaGraphLaunch(g->exec, (cudaStream_t)stream));
  g->agd->launched += g->chunk;
  if (g->peer) g->peer->seq += (unsigned long long)g->chunk;
  return DUALIP_OK;
}

int dualip_matching_calc_peer(dualip_plan* p, dualip_peer* peer, const float* lambda_dev, const float* b_dev, double gamma,
                              float* grad_out_dev, dualip_scalars* scalars_out_dev, void* stream) {
  if (!p || !peer || !lambda_dev || !grad_out_dev || !scalars_out_dev) {
    set_error("null argument");
    return DUALIP_EINVAL;
  }
  if (!peer->connected || peer->m != p->m || peer->device != p->device) {
    set_error("exchange window does not match the plan, or is not connected");
    return DUALIP_EINVAL;
  }
  DeviceGuard g(p->device);
  FuseSpec f;
  f.mode = 3;
  f.peer = peer_args(peer, true);
  float* slot = reinterpret_cast<float*>(peer->window + kPeerFlagBytes + (size_t)(f.peer.seq & 1ull) * peer->slot_bytes);
  return launch_eval(p, lambda_dev, b_dev, gamma, grad_out_dev, scalars_out_dev, slot, nullptr, nullptr, 0, (cudaStream_t)stream, &f);
}

int dualip_matching_calc_peer_host(dualip_plan* p, dualip_peer* peer, const float* lambda_host, const float* b_dev, double gamma,
                                   float* grad_out_host, dualip_scalars* scalars_out_host, void* stream) {
  if (!p || !peer || !lambda_host || !grad_out_host || !scalars_out_host) {
    set_error("null argument");
    return DUALIP_EINVAL;
  }
  DeviceGuard g(p->device);
  cudaStream_t st = (cudaStream_t)stream;
  DUALIP_CUDA_TRY(cudaMemcpyAsync(p->lambda_stage, lambda_host, sizeof(float) * p->m, cudaMemcpyHostToDevice, st));
  int rc = dualip_matching_calc_peer(p, peer, p->lambda_stage, b_dev, gamma, p->grad_stage, p->scal_stage, stream);
  if (rc != DUALIP_OK) return rc;
  DUALIP_CUDA_TRY(cudaMemcpyAsync(grad_out_host, p->grad_stage, sizeof(float) * p->m, cudaMemcpyDeviceToHost, st));
  DUALIP_CUDA_TRY(cudaMemcpyAsync(scalars_out_host, p->scal_stage, sizeof(dualip_scalars), cudaMemcpyDeviceToHost, st));
  DUALIP_CUDA_TRY(cudaStreamSynchronize(st));
  if (*peer->status_host != 0) {
    set_error("peer exchange: a rank did not arrive within the time-out");
    return DUALIP_ECUDA;
  }
  return DUALIP_OK;
}

int dualip_matching_step_host(dualip_plan* p, dualip_peer* peer, dualip_agd_host* agd, const float* b_dev, double gamma, float beta,
                              int32_t decay_now, double decay_factor, float* grad_out_host, dualip_scalars* scalars_out_host,
                              double* step_out, void* stream) {
  if (!p || !agd || !grad_out_host || !scalars_out_host) {
    set_error("null argument");
    return DUALIP_EINVAL;
  }
  const float* x_host = dualip_agd_host_x(agd);  // the evaluation point (pinned when a CUDA device is present)
  int rc = peer ? dualip_matching_calc_peer_host(p, peer, x_host, b_dev, gamma, grad_out_host, scalars_out_host, stream)
                : dualip_matching_calc_host(p, x_host, b_dev, gamma, grad_out_host, scalars_out_host, stream);
  if (rc != DUALIP_OK) return rc;
  return dualip_agd_host_step(agd, grad_out_host, beta, decay_now, decay_factor, step_out);
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
  if (p->grid_status_host && *reinterpret_cast<volatile int*>(p->grid_status_host) != 0) {
    set_error("all-CTA tail: a grid-wide barrier timed out (the CTAs of the launch were not co-resident)");
    return DUALIP_ECUDA;
  }
  return DUALIP_OK;
}

}  // extern "C"

// Device-resident Maximizer state and the fused per-iteration update (one CTA, no host sync).
//
// Restates AcceleratedGradientDescent.maximize's rank-0 update (reference src/dualip/optimizers/agd.py:163-187)
// and calculate_step_size (optimizers/agd_utils.py:4-89) with the history ring and the Lipschitz ratios kept on
// the device.  The reference recomputes all <=14 ratios every iteration (agd_utils.py:86-88); they only depend on
// stored history entries, so caching them and computing the newest pair gives the same numbers.
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <new>

#include "common.cuh"

using namespace dualip;

struct dualip_agd {
  int device = 0;
  int m = 0;
  int H = 15;
  float* x = nullptr;
  float* y = nullptr;
  float* gh = nullptr;      // H x m ring of gradients
  float* yh = nullptr;      // H x m ring of y iterates (the reference stores y, not x: agd.py:170-172)
  float* ratios = nullptr;  // H-1 ring: ratio of pair (j, j+1) at slot j % (H-1)
  long long* pushes = nullptr;  // number of history pushes so far
  double* dstate = nullptr;     // [0] max_step_size (mutable: gamma decay), [1] initial_step_size
  uint8_t* eqmask = nullptr;
  double* log_obj = nullptr;
  double* log_step = nullptr;
  int log_cap = 0;
};

namespace dualip {

// One CTA.  FROM_PARTIAL: `grad` points at the all-reduced packed sums [sum_j a_rj x_rj (m) | c.x | ||x||^2] of the sharded
// path; the kernel then also does the m-length tail of the objective (grad = sum - b, lambda.grad, slacks, dual objective:
// matching.py:280-299), writes grad_out / scal_out, and saves the separate epilogue launch.  Both loops are unrolled by
// four with every load of a round issued before the first use: a single CTA has no other warps to hide L2 latency.
struct AgdStepArgs {
  float* x;
  float* y;
  float* gh;
  float* yh;
  float* ratios;
  long long* pushes;
  double* dstate;
  const uint8_t* eqmask;
  const float* grad;  // FROM_PARTIAL: packed sums, m+2 floats
  const dualip_scalars* scal;
  int m, H;
  float beta;
  int decay_now;
  double decay_factor;
  double* log_obj;
  double* log_step;
  int iter_index;
  const float* b;
  double gamma;
  float* grad_out;
  dualip_scalars* scal_out;
};

template <bool FROM_PARTIAL>
__device__ __forceinline__ void agd_step_body(const AgdStepArgs& A) {
  float* __restrict__ x = A.x;
  float* __restrict__ y = A.y;
  float* __restrict__ gh = A.gh;
  float* __restrict__ yh = A.yh;
  float* __restrict__ ratios = A.ratios;
  long long* __restrict__ pushes = A.pushes;
  double* __restrict__ dstate = A.dstate;
  const uint8_t* __restrict__ eqmask = A.eqmask;
  const float* grad = A.grad;
  const dualip_scalars* __restrict__ scal = A.scal;
  const int m = A.m, H = A.H;
  const float beta = A.beta;
  const int decay_now = A.decay_now, iter_index = A.iter_index;
  const double decay_factor = A.decay_factor, gamma = A.gamma;
  double* log_obj = A.log_obj;
  double* log_step = A.log_step;
  const float* __restrict__ b = A.b;
  float* grad_out = A.grad_out;
  dualip_scalars* __restrict__ scal_out = A.scal_out;
  __shared__ double s_red[5][32];
  __shared__ float s_mx[32];
  __shared__ double s_step;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int lane = tid & 31, warp = tid >> 5, nw = (nt + 31) >> 5;
  const long long t = *pushes;  // index of the entry pushed now
  const int slot = (int)(t % H);
  const int prev = (int)((t + H - 1) % H);
  const bool have_prev = t > 0;
  // 1) gradient (sharded path: the objective's tail), push (grad, y), measure the newest pair
  //    agd_utils.py:11-27, :30-41 ; matching.py:280-299
  double dg2 = 0.0, dy2 = 0.0, lg = 0.0, sp = 0.0, g2 = 0.0;
  float mx = -INFINITY;
  for (int base = tid; base < m; base += 4 * nt) {
    float g4[4], y4[4], gp4[4], yp4[4], b4[4], x4[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = base + u * nt;
      g4[u] = y4[u] = gp4[u] = yp4[u] = b4[u] = x4[u] = 0.f;
      if (i < m) {
        g4[u] = grad[i];
        y4[u] = y[i];
        if (have_prev) {
          gp4[u] = gh[(size_t)prev * m + i];
          yp4[u] = yh[(size_t)prev * m + i];
        }
        if (FROM_PARTIAL) {
          b4[u] = b ? b[i] : 0.f;
          x4[u] = x[i];
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = base + u * nt;
      if (i < m) {
        float g = g4[u];
        if (FROM_PARTIAL) {
          g = b ? __fsub_rn(g, b4[u]) : g;
          grad_out[i] = g;
          lg = fma((double)x4[u], (double)g, lg);
          sp += (double)fmaxf(g, 0.f);
          g2 = fma((double)g, (double)g, g2);
          mx = fmaxf(mx, g);
        }
        gh[(size_t)slot * m + i] = g;
        yh[(size_t)slot * m + i] = y4[u];
        if (have_prev) {
          const float dg = __fsub_rn(gp4[u], g);
          const float dy = __fsub_rn(yp4[u], y4[u]);
          dg2 = fma((double)dg, (double)dg, dg2);
          dy2 = fma((double)dy, (double)dy, dy2);
        }
      }
    }
  }
  dg2 = warp_sum(dg2);
  dy2 = warp_sum(dy2);
  if (FROM_PARTIAL) {
    lg = warp_sum(lg);
    sp = warp_sum(sp);
    g2 = warp_sum(g2);
    mx = warp_max(mx);
  }
  if (lane == 0) {
    s_red[0][warp] = dg2;
    s_red[1][warp] = dy2;
    if (FROM_PARTIAL) {
      s_red[2][warp] = lg;
      s_red[3][warp] = sp;
      s_red[4][warp] = g2;
      s_mx[warp] = mx;
    }
  }
  __syncthreads();
  if (warp == 0) {
    dg2 = warp_sum(lane < nw ? s_red[0][lane] : 0.0);
    dy2 = warp_sum(lane < nw ? s_red[1][lane] : 0.0);
    if (FROM_PARTIAL) {
      lg = warp_sum(lane < nw ? s_red[2][lane] : 0.0);
      sp = warp_sum(lane < nw ? s_red[3][lane] : 0.0);
      g2 = warp_sum(lane < nw ? s_red[4][lane] : 0.0);
      mx = warp_max(lane < nw ? s_mx[lane] : -INFINITY);
    }
  }
  if (tid == 0) {
    double dual_obj = scal ? scal->dual_objective : 0.0;
    if (FROM_PARTIAL) {
      const double cxv = (double)grad[m], xxv = (double)grad[m + 1];
      dualip_scalars r;
      r.primal_objective = cxv;
      r.reg_penalty = 0.5 * gamma * xxv;
      r.dual_val_times_grad = lg;
      r.dual_objective = cxv + r.reg_penalty + lg;
      r.max_pos_slack = (double)fmaxf(mx, 0.f);
      r.sum_pos_slack = sp;
      r.x_sq_norm = xxv;
      r.grad_sq_norm = g2;
      *scal_out = r;
      dual_obj = r.dual_objective;
    }
    if (have_prev) ratios[(t - 1) % (H - 1)] = __fdiv_rn((float)sqrt(dg2), (float)sqrt(dy2));
    // 2) step size                                                      agd_utils.py:44-62
    const long long n_pairs = t < (long long)(H - 1) ? t : (long long)(H - 1);
    const double max_step = dstate[0], init_step = dstate[1];
    double step = init_step;
    if (n_pairs >= H - 1) {
      // Python max() over the list in chronological order: the first element wins unless a later one is greater
      const long long j0 = t - (H - 1);
      float lmax = ratios[j0 % (H - 1)];
      for (long long j = j0 + 1; j < t; ++j) {
        const float v = ratios[j % (H - 1)];
        if (v > lmax) lmax = v;
      }
      if (!(isnan(lmax) || isinf(lmax))) {
        const double cand = (lmax != 0.f) ? 1.0 / (double)lmax : max_step;
        step = cand < max_step ? cand : max_step;
      }
    }
    s_step = step;
    if (log_obj) log_obj[iter_index] = dual_obj;
    if (log_step) log_step[iter_index] = step;
    if (decay_now) dstate[0] = step * decay_factor;  // agd.py:107
    *pushes = t + 1;
  }
  __syncthreads();
  // 3) ascent step, projection on the dual cone, momentum              agd.py:181-185, :13-21
  const float step32 = (float)s_step;
  const float omb = __fsub_rn(1.0f, beta);
  const float* gsrc = FROM_PARTIAL ? grad_out : grad;
  for (int base = tid; base < m; base += 4 * nt) {
    float g4[4], y4[4], x4[4];
    uint8_t e4[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = base + u * nt;
      g4[u] = y4[u] = x4[u] = 0.f;
      e4[u] = 0;
      if (i < m) {
        g4[u] = gsrc[i];
        y4[u] = y[i];
        x4[u] = x[i];
        e4[u] = eqmask ? eqmask[i] : (uint8_t)0;
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = base + u * nt;
      if (i < m) {
        float yn = __fadd_rn(x4[u], __fmul_rn(g4[u], step32));
        if (!e4[u]) yn = fmaxf(yn, 0.f);
        x[i] = __fadd_rn(__fmul_rn(yn, omb), __fmul_rn(y4[u], beta));
        y[i] = yn;
      }
    }
  }
}

template <bool FROM_PARTIAL>
__global__ void __launch_bounds__(1024) agd_step_kernel(const AgdStepArgs A) {
  agd_step_body<FROM_PARTIAL>(A);
}

// ---- peer-memory exchange: arrival flags and slots in every rank's window (include/dualip_b200.h) ----
constexpr int kPeerFlagBytes = 256;  // DUALIP_PEER_MAX_WORLD x 8-byte arrival flags, padded

struct PeerArgs {
  unsigned char* win[DUALIP_PEER_MAX_WORLD];  // window base of every rank, as mapped in this process
  int rank, world;
  unsigned long long seq;  // number of this step (1, 2, ...): the flag value, and seq & 1 the slot
  size_t slot_bytes;
  unsigned long long timeout_ns;
  unsigned int* ticket;  // local word, 0 between steps: which CTA of the step kernel finishes last
  float* sum;   // local m+2 floats (padded to a multiple of 4): the reduced packed sums
  int* status;  // local word: set to 1 when a wait timed out
};

__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ld_relaxed_sys_f32(const float* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ unsigned long long ld_relaxed_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_relaxed_sys_f4(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}

// ceil((m+2) / 4096) CTAs of 1024 threads.  CTA 0 tells every peer that this rank's partial sums (written by the preceding
// kernel on the same stream) are in their slot; every CTA waits until all peers have said the same, then each thread
// fetches ONE float4 from every rank's slot -- all W loads in flight together, a single NVLink round trip -- and adds
// them in rank order, so that every rank computes bit-identical sums.  The CTA that finishes last (atomic ticket) runs the
// objective's tail and the accelerated update on the reduced vector.
__global__ void __launch_bounds__(1024) agd_step_peer_kernel(const AgdStepArgs A, const PeerArgs P) {
  __shared__ int s_last;
  const int tid = threadIdx.x;
  const int m2 = A.m + 2;
  if (tid < P.world) {
    // flags[r] of rank t's window is written by rank r only; the release makes the slot visible before the flag
    if (blockIdx.x == 0) st_release_sys_u64(reinterpret_cast<unsigned long long*>(P.win[tid]) + P.rank, P.seq);
    const unsigned long long* mine = reinterpret_cast<const unsigned long long*>(P.win[P.rank]) + tid;
    const unsigned long long t0 = global_timer_ns();
    // once a wait has timed out the run is invalid (the host raises): later steps must not wait the full time-out again
    while (*reinterpret_cast<volatile int*>(P.status) == 0 && ld_relaxed_sys_u64(mine) < P.seq) {
      if (global_timer_ns() - t0 > P.timeout_ns) {
        *P.status = 1;
        break;
      }
    }
    asm volatile("fence.acq_rel.sys;" ::: "memory");  // the peers' slots are read after their flags
  }
  __syncthreads();
  const size_t slot_off = (size_t)kPeerFlagBytes + (size_t)(P.seq & 1ull) * P.slot_bytes;
  const int i4 = (blockIdx.x * 1024 + tid) * 4;  // slots are padded to 128 bytes: a float4 never leaves the slot
  if (i4 < m2) {
    float4 v[DUALIP_PEER_MAX_WORLD];
#pragma unroll
    for (int r = 0; r < DUALIP_PEER_MAX_WORLD; ++r)
      if (r < P.world) v[r] = ld_relaxed_sys_f4(reinterpret_cast<const float*>(P.win[r] + slot_off) + i4);
    float4 acc = v[0];
#pragma unroll
    for (int r = 1; r < DUALIP_PEER_MAX_WORLD; ++r) {
      if (r < P.world) {
        acc.x = __fadd_rn(acc.x, v[r].x);
        acc.y = __fadd_rn(acc.y, v[r].y);
        acc.z = __fadd_rn(acc.z, v[r].z);
        acc.w = __fadd_rn(acc.w, v[r].w);
      }
    }
    if (i4 + 3 < m2) {
      *reinterpret_cast<float4*>(P.sum + i4) = acc;
    } else {
      P.sum[i4] = acc.x;
      if (i4 + 1 < m2) P.sum[i4 + 1] = acc.y;
      if (i4 + 2 < m2) P.sum[i4 + 2] = acc.z;
    }
  }
  if (gridDim.x > 1) {
    __threadfence();
    __syncthreads();
    if (tid == 0) {
      const unsigned int ticket = atomicAdd(P.ticket, 1u);
      s_last = (ticket == gridDim.x - 1) ? 1 : 0;
      if (s_last) *P.ticket = 0u;  // ready for the next step
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
  } else {
    __syncthreads();  // P.sum is read below by other threads of this CTA
  }
  agd_step_body<true>(A);
}

}  // namespace dualip

extern "C" {

void dualip_agd_destroy(dualip_agd* a) {
  if (!a) return;
  DeviceGuard g(a->device);
  cudaFree(a->x);
  cudaFree(a->y);
  cudaFree(a->gh);
  cudaFree(a->yh);
  cudaFree(a->ratios);
  cudaFree(a->pushes);
  cudaFree(a->dstate);
  cudaFree(a->eqmask);
  cudaFree(a->log_obj);
  cudaFree(a->log_step);
  delete a;
}

int dualip_agd_reserve_log(dualip_agd* a, int32_t capacity) {
  if (!a || capacity < 0) {
    set_error("bad argument");
    return DUALIP_EINVAL;
  }
  if (capacity <= a->log_cap) return DUALIP_OK;
  DeviceGuard g(a->device);
  double *lo = nullptr, *ls = nullptr;
  DUALIP_CUDA_TRY(cudaMalloc(&lo, sizeof(double) * capacity));
  DUALIP_CUDA_TRY(cudaMalloc(&ls, sizeof(double) * capacity));
  DUALIP_CUDA_TRY(cudaMemset(lo, 0, sizeof(double) * capacity));
  DUALIP_CUDA_TRY(cudaMemset(ls, 0, sizeof(double) * capacity));
  if (a->log_cap > 0) {
    DUALIP_CUDA_TRY(cudaMemcpy(lo, a->log_obj, sizeof(double) * a->log_cap, cudaMemcpyDeviceToDevice));
    DUALIP_CUDA_TRY(cudaMemcpy(ls, a->log_step, sizeof(double) * a->log_cap, cudaMemcpyDeviceToDevice));
  }
  cudaFree(a->log_obj);
  cudaFree(a->log_step);
  a->log_obj = lo;
  a->log_step = ls;
  a->log_cap = capacity;
  return DUALIP_OK;
}

int dualip_agd_create(dualip_agd** out, int32_t m, int32_t device, const float* initial_dev,
                      const uint8_t* equality_mask_dev, double initial_step_size, double max_step_size,
                      int32_t history_len) {
  if (!out || m <= 0 || history_len < 2) {
    set_error("bad argument");
    return DUALIP_EINVAL;
  }
  *out = nullptr;
  DeviceGuard g(device);
  if (!g.ok) {
    set_error("cannot select CUDA device %d", device);
    return DUALIP_ECUDA;
  }
  dualip_agd* a = new (std::nothrow) dualip_agd();
  if (!a) return DUALIP_ENOMEM;
  a->device = device;
  a->m = m;
  a->H = history_len;
#define AGD_TRY(expr)                                                 \
  do {                                                                \
    cudaError_t _e = (expr);                                          \
    if (_e != cudaSuccess) {                                          \
      set_error("%s failed: %s", #expr, cudaGetErrorString(_e));      \
      dualip_agd_destroy(a);                                          \
      return DUALIP_ECUDA;                                            \
    }                                                                 \
  } while (0)
  AGD_TRY(cudaMalloc(&a->x, sizeof(float) * (m + 4)));
  AGD_TRY(cudaMalloc(&a->y, sizeof(float) * (m + 4)));
  AGD_TRY(cudaMalloc(&a->gh, sizeof(float) * (size_t)m * a->H));
  AGD_TRY(cudaMalloc(&a->yh, sizeof(float) * (size_t)m * a->H));
  AGD_TRY(cudaMalloc(&a->ratios, sizeof(float) * a->H));
  AGD_TRY(cudaMalloc(&a->pushes, sizeof(long long)));
  AGD_TRY(cudaMalloc(&a->dstate, sizeof(double) * 2));
  AGD_TRY(cudaMemset(a->pushes, 0, sizeof(long long)));
  AGD_TRY(cudaMemset(a->ratios, 0, sizeof(float) * a->H));
  if (initial_dev) {
    AGD_TRY(cudaMemcpy(a->x, initial_dev, sizeof(float) * m, cudaMemcpyDeviceToDevice));
    AGD_TRY(cudaMemcpy(a->y, initial_dev, sizeof(float) * m, cudaMemcpyDeviceToDevice));
  } else {
    AGD_TRY(cudaMemset(a->x, 0, sizeof(float) * m));
    AGD_TRY(cudaMemset(a->y, 0, sizeof(float) * m));
  }
  const double ds[2] = {max_step_size, initial_step_size};
  AGD_TRY(cudaMemcpy(a->dstate, ds, sizeof(ds), cudaMemcpyHostToDevice));
  if (equality_mask_dev) {
    AGD_TRY(cudaMalloc(&a->eqmask, m));
    AGD_TRY(cudaMemcpy(a->eqmask, equality_mask_dev, m, cudaMemcpyDeviceToDevice));
  }
#undef AGD_TRY
  *out = a;
  return DUALIP_OK;
}

const float* dualip_agd_x(const dualip_agd* a) { return a ? a->x : nullptr; }
const float* dualip_agd_y(const dualip_agd* a) { return a ? a->y : nullptr; }

int dualip_agd_get(dualip_agd* a, float* x_out_dev, float* y_out_dev, void* stream) {
  if (!a) {
    set_error("null argument");
    return DUALIP_EINVAL;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (x_out_dev) DUALIP_CUDA_TRY(cudaMemcpyAsync(x_out_dev, a->x, sizeof(float) * a->m, cudaMemcpyDeviceToDevice, st));
  if (y_out_dev) DUALIP_CUDA_TRY(cudaMemcpyAsync(y_out_dev, a->y, sizeof(float) * a->m, cudaMemcpyDeviceToDevice, st));
  return DUALIP_OK;
}

static AgdStepArgs step_args(dualip_agd* a, const float* grad, const dualip_scalars* scal, float beta, int decay_now,
                             double decay_factor, int iter_index, const float* b, double gamma, float* grad_out,
                             dualip_scalars* scal_out) {
  const bool log = iter_index >= 0 && iter_index < a->log_cap;
  AgdStepArgs A;
  A.x = a->x;
  A.y = a->y;
  A.gh = a->gh;
  A.yh = a->yh;
  A.ratios = a->ratios;
  A.pushes = a->pushes;
  A.dstate = a->dstate;
  A.eqmask = a->eqmask;
  A.grad = grad;
  A.scal = scal;
  A.m = a->m;
  A.H = a->H;
  A.beta = beta;
  A.decay_now = decay_now;
  A.decay_factor = decay_factor;
  A.log_obj = log ? a->log_obj : nullptr;
  A.log_step = log ? a->log_step : nullptr;
  A.iter_index = log ? iter_index : 0;
  A.b = b;
  A.gamma = gamma;
  A.grad_out = grad_out;
  A.scal_out = scal_out;
  return A;
}

int dualip_agd_step(dualip_agd* a, const float* grad_dev, const dualip_scalars* scalars_dev, float beta,
                    int32_t decay_now, double decay_factor, int32_t iter_index, void* stream) {
  if (!a || !grad_dev) {
    set_error("null argument");
    return DUALIP_EINVAL;
  }
  agd_step_kernel<false><<<1, 1024, 0, (cudaStream_t)stream>>>(
      step_args(a, grad_dev, scalars_dev, beta, decay_now, decay_factor, iter_index, nullptr, 0.0, nullptr, nullptr));
  DUALIP_CUDA_TRY(cudaGetLastError());
  return DUALIP_OK;
}

int dualip_agd_step_sharded(dualip_agd* a, const float* partial_sum_dev, const float* b_dev, double gamma, float* grad_out_dev,
                            dualip_scalars* scalars_out_dev, float beta, int32_t decay_now, double decay_factor,
                            int32_t iter_index, void* stream) {
  if (!a || !partial_sum_dev || !grad_out_dev || !scalars_out_dev) {
    set_error("null argument");
    return DUALIP_EINVAL;
  }
  agd_step_kernel<true><<<1, 1024, 0, (cudaStream_t)stream>>>(step_args(
      a, partial_sum_dev, nullptr, beta, decay_now, decay_factor, iter_index, b_dev, gamma, grad_out_dev, scalars_out_dev));
  DUALIP_CUDA_TRY(cudaGetLastError());
  return DUALIP_OK;
}

// ---- exchange windows ----
struct dualip_peer {
  int device = 0, m = 0, rank = 0, world = 1;
  size_t slot_bytes = 0, window_bytes = 0;
  unsigned char* window = nullptr;                    // own window (cudaMalloc: exportable through CUDA IPC)
  unsigned char* win[DUALIP_PEER_MAX_WORLD] = {};     // all windows as mapped here
  bool opened[DUALIP_PEER_MAX_WORLD] = {};            // mapped with cudaIpcOpenMemHandle
  bool connected = false;
  float* sum = nullptr;
  int* status = nullptr;  // [0] status, [1] ticket of the step kernel
  unsigned long long seq = 0;  // steps taken
  unsigned long long timeout_ns = 20ull * 1000ull * 1000ull * 1000ull;  // DUALIP_PEER_TIMEOUT_MS overrides
};

void dualip_peer_destroy(dualip_peer* p) {
  if (!p) return;
  DeviceGuard g(p->device);
  for (int r = 0; r < DUALIP_PEER_MAX_WORLD; ++r)
    if (p->opened[r]) cudaIpcCloseMemHandle(p->win[r]);
  cudaFree(p->window);
  cudaFree(p->sum);
  cudaFree(p->status);
  delete p;
}

int dualip_peer_create(dualip_peer** out, int32_t m, int32_t rank, int32_t world, int32_t device) {
  if (!out || m <= 0 || world < 1 || world > DUALIP_PEER_MAX_WORLD || rank < 0 || rank >= world) {
    set_error("bad argument (world must be 1..%d)", DUALIP_PEER_MAX_WORLD);
    return DUALIP_EINVAL;
  }
  *out = nullptr;
  DeviceGuard g(device);
  if (!g.ok) {
    set_error("cannot select CUDA device %d", device);
    return DUALIP_ECUDA;
  }
  dualip_peer* p = new (std::nothrow) dualip_peer();
  if (!p) return DUALIP_ENOMEM;
  p->device = device;
  p->m = m;
  p->rank = rank;
  p->world = world;
  if (const char* env = getenv("DUALIP_PEER_TIMEOUT_MS")) {
    const long long ms = atoll(env);
    if (ms > 0) p->timeout_ns = (unsigned long long)ms * 1000000ull;
  }
  p->slot_bytes = (sizeof(float) * (size_t)(m + 2) + 127) & ~(size_t)127;
  p->window_bytes = kPeerFlagBytes + 2 * p->slot_bytes;
  cudaError_t e = cudaMalloc(&p->window, p->window_bytes);
  if (e == cudaSuccess) e = cudaMemset(p->window, 0, p->window_bytes);
  if (e == cudaSuccess) e = cudaMalloc(&p->sum, sizeof(float) * (m + 8));
  if (e == cudaSuccess) e = cudaMalloc(&p->status, 2 * sizeof(int));
  if (e == cudaSuccess) e = cudaMemset(p->status, 0, 2 * sizeof(int));
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    set_error("allocating the exchange window failed: %s", cudaGetErrorString(e));
    dualip_peer_destroy(p);
    return DUALIP_ECUDA;
  }
  p->win[rank] = p->window;
  p->connected = world == 1;
  *out = p;
  return DUALIP_OK;
}

int dualip_peer_export(dualip_peer* p, uint8_t* handle_out) {
  static_assert(sizeof(cudaIpcMemHandle_t) == DUALIP_PEER_HANDLE_BYTES, "IPC handle size");
  if (!p || !handle_out) {
    set_error("null argument");
    return DUALIP_EINVAL;
  }
  DeviceGuard g(p->device);
  cudaIpcMemHandle_t h;
  DUALIP_CUDA_TRY(cudaIpcGetMemHandle(&h, p->window));
  memcpy(handle_out, &h, sizeof(h));
  return DUALIP_OK;
}

int dualip_peer_connect_ipc(dualip_peer* p, const uint8_t* handles) {
  if (!p || !handles) {
    set_error("null argument");
    return DUALIP_EINVAL;
  }
  DeviceGuard g(p->device);
  for (int r = 0; r < p->world; ++r) {
    if (r == p->rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + (size_t)r * DUALIP_PEER_HANDLE_BYTES, sizeof(h));
    void* ptr = nullptr;
    DUALIP_CUDA_TRY(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    p->win[r] = static_cast<unsigned char*>(ptr);
    p->opened[r] = true;
  }
  p->connected = true;
  return DUALIP_OK;
}

int dualip_peer_connect_ptrs(dualip_peer* p, void* const* windows) {
  if (!p || !windows) {
    set_error("null argument");
    return DUALIP_EINVAL;
  }
  for (int r = 0; r < p->world; ++r) {
    if (r == p->rank) continue;
    if (!windows[r]) {
      set_error("window %d is null", r);
      return DUALIP_EINVAL;
    }
    p->win[r] = static_cast<unsigned char*>(windows[r]);
  }
  p->connected = true;
  return DUALIP_OK;
}

void* dualip_peer_window(dualip_peer* p) { return p ? p->window : nullptr; }

float* dualip_peer_next_slot(dualip_peer* p) {
  if (!p) return nullptr;
  return reinterpret_cast<float*>(p->window + kPeerFlagBytes + (size_t)((p->seq + 1) & 1ull) * p->slot_bytes);
}

int dualip_peer_status(dualip_peer* p, int32_t* status_out, void* stream) {
  if (!p || !status_out) {
    set_error("null argument");
    return DUALIP_EINVAL;
  }
  DeviceGuard g(p->device);
  int v = 0;
  DUALIP_CUDA_TRY(cudaMemcpyAsync(&v, p->status, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  DUALIP_CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
  *status_out = v;
  return DUALIP_OK;
}

int dualip_agd_step_peer(dualip_agd* a, dualip_peer* p, const float* b_dev, double gamma, float* grad_out_dev,
                         dualip_scalars* scalars_out_dev, float beta, int32_t decay_now, double decay_factor,
                         int32_t iter_index, void* stream) {
  if (!a || !p || !grad_out_dev || !scalars_out_dev) {
    set_error("null argument");
    return DUALIP_EINVAL;
  }
  if (!p->connected || p->m != a->m || p->device != a->device) {
    set_error("exchange window is not connected or does not match the optimizer state");
    return DUALIP_EINVAL;
  }
  PeerArgs P;
  for (int r = 0; r < DUALIP_PEER_MAX_WORLD; ++r) P.win[r] = r < p->world ? p->win[r] : nullptr;
  P.rank = p->rank;
  P.world = p->world;
  P.seq = ++p->seq;
  P.slot_bytes = p->slot_bytes;
  P.timeout_ns = p->timeout_ns;
  P.sum = p->sum;
  P.status = p->status;
  P.ticket = reinterpret_cast<unsigned int*>(p->status + 1);
  const int n_ctas = (a->m + 2 + 4095) / 4096;
  agd_step_peer_kernel<<<n_ctas, 1024, 0, (cudaStream_t)stream>>>(
      step_args(a, p->sum, nullptr, beta, decay_now, decay_factor, iter_index, b_dev, gamma, grad_out_dev, scalars_out_dev), P);
  DUALIP_CUDA_TRY(cudaGetLastError());
  return DUALIP_OK;
}

// ---- host-resident twin of the optimizer state: the same update for callers that keep the iterate in host memory ----
struct dualip_agd_host {
  int m = 0, H = 15;
  float* x = nullptr;   // pinned when a CUDA device is present (the caller copies it host->device every iteration)
  float* y = nullptr;
  bool pinned = false;
  float* gh = nullptr;
  float* yh = nullptr;
  float ratios[64] = {};
  long long pushes = 0;
  double max_step = 0.1, init_step = 1e-5;
  uint8_t* eqmask = nullptr;
};

void dualip_agd_host_destroy(dualip_agd_host* h) {
  if (!h) return;
  if (h->pinned) {
    cudaFreeHost(h->x);
  } else {
    free(h->x);
  }
  free(h->y);
  free(h->gh);
  free(h->yh);
  free(h->eqmask);
  delete h;
}

int dualip_agd_host_create(dualip_agd_host** out, int32_t m, const float* initial_host, const uint8_t* equality_mask_host,
                           double initial_step_size, double max_step_size, int32_t history_len) {
  if (!out || m <= 0 || history_len < 2 || history_len > 64) {
    set_error("bad argument");
    return DUALIP_EINVAL;
  }
  *out = nullptr;
  dualip_agd_host* h = new (std::nothrow) dualip_agd_host();
  if (!h) return DUALIP_ENOMEM;
  h->m = m;
  h->H = history_len;
  h->max_step = max_step_size;
  h->init_step = initial_step_size;
  void* px = nullptr;
  if (cudaHostAlloc(&px, sizeof(float) * m, cudaHostAllocDefault) == cudaSuccess) {
    h->x = static_cast<float*>(px);
    h->pinned = true;
  } else {
    cudaGetLastError();  // no device / no driver: plain memory (the update itself needs no GPU)
    h->x = static_cast<float*>(malloc(sizeof(float) * m));
  }
  h->y = static_cast<float*>(malloc(sizeof(float) * m));
  h->gh = static_cast<float*>(malloc(sizeof(float) * (size_t)m * h->H));
  h->yh = static_cast<float*>(malloc(sizeof(float) * (size_t)m * h->H));
  if (equality_mask_host) h->eqmask = static_cast<uint8_t*>(malloc(m));
  if (!h->x || !h->y || !h->gh || !h->yh || (equality_mask_host && !h->eqmask)) {
    dualip_agd_host_destroy(h);
    set_error("out of host memory");
    return DUALIP_ENOMEM;
  }
  for (int i = 0; i < m; ++i) h->x[i] = h->y[i] = initial_host ? initial_host[i] : 0.f;
  if (equality_mask_host) memcpy(h->eqmask, equality_mask_host, m);
  *out = h;
  return DUALIP_OK;
}

float* dualip_agd_host_x(dualip_agd_host* h) { return h ? h->x : nullptr; }
float* dualip_agd_host_y(dualip_agd_host* h) { return h ? h->y : nullptr; }

// Same arithmetic as agd_step_kernel<false>: float32 differences and iterates, norms accumulated in double.
int dualip_agd_host_step(dualip_agd_host* h, const float* grad_host, float beta, int32_t decay_now, double decay_factor,
                         double* step_out) {
  if (!h || !grad_host) {
    set_error("null argument");
    return DUALIP_EINVAL;
  }
  const int m = h->m, H = h->H;
  const long long t = h->pushes;
  const int slot = (int)(t % H), prev = (int)((t + H - 1) % H);
  float* gs = h->gh + (size_t)slot * m;
  float* ys = h->yh + (size_t)slot * m;
  const float* gp = h->gh + (size_t)prev * m;
  const float* yp = h->yh + (size_t)prev * m;
  double dg2 = 0.0, dy2 = 0.0;
  if (t > 0) {
    for (int i = 0; i < m; ++i) {
      const float g = grad_host[i], yv = h->y[i];
      const float dg = gp[i] - g, dy = yp[i] - yv;
      dg2 += (double)dg * (double)dg;
      dy2 += (double)dy * (double)dy;
      gs[i] = g;
      ys[i] = yv;
    }
    h->ratios[(t - 1) % (H - 1)] = (float)sqrt(dg2) / (float)sqrt(dy2);
  } else {
    memcpy(gs, grad_host, sizeof(float) * m);
    memcpy(ys, h->y, sizeof(float) * m);
  }
  const long long n_pairs = t < (long long)(H - 1) ? t : (long long)(H - 1);
  double step = h->init_step;
  if (n_pairs >= H - 1) {
    const long long j0 = t - (H - 1);
    float lmax = h->ratios[j0 % (H - 1)];
    for (long long j = j0 + 1; j < t; ++j) {
      const float v = h->ratios[j % (H - 1)];
      if (v > lmax) lmax = v;
    }
    if (!(isnan(lmax) || isinf(lmax))) {
      const double cand = (lmax != 0.f) ? 1.0 / (double)lmax : h->max_step;
      step = cand < h->max_step ? cand : h->max_step;
    }
  }
  if (decay_now) h->max_step = step * decay_factor;
  h->pushes = t + 1;
  const float step32 = (float)step, omb = 1.0f - beta;
  for (int i = 0; i < m; ++i) {
    float yn = h->x[i] + grad_host[i] * step32;
    if (!(h->eqmask && h->eqmask[i])) yn = yn > 0.f ? yn : 0.f;
    const float a = yn * omb, b = h->y[i] * beta;
    h->x[i] = a + b;
    h->y[i] = yn;
  }
  if (step_out) *step_out = step;
  return DUALIP_OK;
}

int dualip_agd_read_log(dualip_agd* a, int32_t count, double* dual_obj_host, double* step_host, void* stream) {
  if (!a || count < 0 || count > a->log_cap) {
    set_error("bad log range");
    return DUALIP_EINVAL;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (count > 0) {
    if (dual_obj_host) DUALIP_CUDA_TRY(cudaMemcpyAsync(dual_obj_host, a->log_obj, sizeof(double) * count, cudaMemcpyDeviceToHost, st));
    if (step_host) DUALIP_CUDA_TRY(cudaMemcpyAsync(step_host, a->log_step, sizeof(double) * count, cudaMemcpyDeviceToHost, st));
  }
  DUALIP_CUDA_TRY(cudaStreamSynchronize(st));
  return DUALIP_OK;
}

}  // extern "C"

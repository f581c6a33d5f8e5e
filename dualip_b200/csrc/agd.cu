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
#include "agd_step.cuh"

using namespace dualip;

namespace dualip {

template <bool FROM_PARTIAL>
__global__ void __launch_bounds__(1024) agd_step_kernel(const AgdStepArgs A) {
  __shared__ TailScratch s_tail;
  agd_step_body<FROM_PARTIAL>(A, step_dyn_of(A), s_tail);
}

// ceil((m+2) / 4096) CTAs of 1024 threads.  CTA 0 tells every peer that this rank's partial sums (written by the preceding
// kernel on the same stream) are in their slot; every CTA waits until all peers have said the same, then each thread
// fetches ONE float4 from every rank's slot -- all W loads in flight together, a single NVLink round trip -- and adds
// them in rank order, so that every rank computes bit-identical sums.  The CTA that finishes last (atomic ticket) runs the
// objective's tail and the accelerated update on the reduced vector.
__global__ void __launch_bounds__(1024) agd_step_peer_kernel(const AgdStepArgs A, const PeerArgs P) {
  __shared__ int s_last;
  __shared__ TailScratch s_tail;
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
        *reinterpret_cast<volatile int*>(P.status_host) = 1;
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
  agd_step_body<true>(A, step_dyn_of(A), s_tail);
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
  cudaFree(a->sched_gamma);
  cudaFree(a->sched_beta);
  cudaFree(a->sched_decay);
  delete a;
}

int dualip_agd_set_schedule(dualip_agd* a, int32_t n_iters, const double* gamma_host, const float* beta_host,
                            const uint8_t* decay_now_host, double decay_factor) {
  if (!a || n_iters <= 0 || !gamma_host || !beta_host) {
    set_error("bad argument");
    return DUALIP_EINVAL;
  }
  for (int i = 0; i < n_iters; ++i)
    if (!(gamma_host[i] > 0.0) && !(gamma_host[i] < 0.0)) {
      set_error("gamma must be non-zero (iteration %d)", i);
      return DUALIP_EINVAL;
    }
  DeviceGuard g(a->device);
  DUALIP_CUDA_TRY(cudaDeviceSynchronize());  // launches that read an earlier schedule have finished
  cudaFree(a->sched_gamma);
  cudaFree(a->sched_beta);
  cudaFree(a->sched_decay);
  a->sched_gamma = nullptr, a->sched_beta = nullptr, a->sched_decay = nullptr, a->sched_n = 0;
  DUALIP_CUDA_TRY(cudaMalloc(&a->sched_gamma, sizeof(double) * n_iters));
  DUALIP_CUDA_TRY(cudaMalloc(&a->sched_beta, sizeof(float) * n_iters));
  DUALIP_CUDA_TRY(cudaMalloc(&a->sched_decay, n_iters));
  DUALIP_CUDA_TRY(cudaMemcpy(a->sched_gamma, gamma_host, sizeof(double) * n_iters, cudaMemcpyHostToDevice));
  DUALIP_CUDA_TRY(cudaMemcpy(a->sched_beta, beta_host, sizeof(float) * n_iters, cudaMemcpyHostToDevice));
  if (decay_now_host)
    DUALIP_CUDA_TRY(cudaMemcpy(a->sched_decay, decay_now_host, n_iters, cudaMemcpyHostToDevice));
  else
    DUALIP_CUDA_TRY(cudaMemset(a->sched_decay, 0, n_iters));
  a->sched_n = n_iters;
  a->sched_factor = decay_factor;
  return DUALIP_OK;
}

long long dualip_agd_steps_launched(const dualip_agd* a) { return a ? a->launched : -1; }

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
void dualip_peer_destroy(dualip_peer* p) {
  if (!p) return;
  DeviceGuard g(p->device);
  for (int r = 0; r < DUALIP_PEER_MAX_WORLD; ++r)
    if (p->opened[r]) cudaIpcCloseMemHandle(p->win[r]);
  cudaFree(p->window);
  cudaFree(p->sum);
  cudaFree(p->ticket);
  cudaFree(p->status);
  if (p->status_host) cudaFreeHost(p->status_host);
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
  if (const char* env = getenv("DUALIP_PEER_PUSH")) p->push = atoi(env) != 0 ? 1 : 0;
  if (const char* env = getenv("DUALIP_PEER_TIMEOUT_MS")) {
    const long long ms = atoll(env);
    if (ms > 0) p->timeout_ns = (unsigned long long)ms * 1000000ull;
  }
  p->slot_bytes = (sizeof(float) * (size_t)(m + 2) + 127) & ~(size_t)127;
  // [arrival flags | two pull slots (this rank's sums, read by the peers) | 2 x world push slots (every rank's sums, written
  //  by that rank: the one-launch path, see peer_push_exchange_cta)]
  p->window_bytes = kPeerFlagBytes + 2 * p->slot_bytes + 2 * (size_t)world * p->slot_bytes;
  cudaError_t e = cudaMalloc(&p->window, p->window_bytes);
  if (e == cudaSuccess) e = cudaMemset(p->window, 0, p->window_bytes);
  if (e == cudaSuccess) e = cudaMalloc(&p->sum, sizeof(float) * (m + 8));
  // the status word lives in mapped host memory: the kernel sets it over PCIe only when a wait times out, and the host can
  // look at it at any time without synchronising the stream
  if (e == cudaSuccess) e = cudaHostAlloc(reinterpret_cast<void**>(&p->status_host), sizeof(int), cudaHostAllocMapped);
  if (e == cudaSuccess) {
    *p->status_host = 0;
    e = cudaHostGetDevicePointer(reinterpret_cast<void**>(&p->status_host_dev), p->status_host, 0);
  }
  if (e == cudaSuccess) e = cudaMalloc(&p->status, sizeof(int));
  if (e == cudaSuccess) e = cudaMemset(p->status, 0, sizeof(int));
  if (e == cudaSuccess) e = cudaMalloc(&p->ticket, sizeof(unsigned int));
  if (e == cudaSuccess) e = cudaMemset(p->ticket, 0, sizeof(unsigned int));
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
  DUALIP_CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
  *status_out = *reinterpret_cast<volatile int*>(p->status_host);
  return DUALIP_OK;
}

int dualip_peer_status_nowait(dualip_peer* p) { return p ? *reinterpret_cast<volatile int*>(p->status_host) : 0; }

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
  const PeerArgs P = peer_args(p, true);
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

// Device code shared by the update kernels (agd.cu) and the slab kernel's fused tail (calc.cu): the accelerated step on the
// device-resident optimizer state, and the peer-memory exchange of the sharded path.
#pragma once
#include "common.cuh"

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
  int log_cap;  // entries of log_obj / log_step (scheduled launches check the device-side iteration index against it)
};

// What changes from one iteration to the next.  Launches issued one by one pass these as kernel arguments (AgdStepArgs);
// scheduled launches -- identical kernel arguments every iteration, hence capturable in a CUDA graph that is replayed --
// read them from a device-resident schedule at index `*pushes` (the number of steps the state has taken).
struct StepDyn {
  float beta;
  int decay_now;
  int iter_index;
  double gamma;
  bool log;
};
struct SchedArgs {
  const double* gamma;         // n entries: gamma of iteration i (null: not scheduled)
  const float* beta;           // n entries: momentum of iteration i (agd.py:93-100)
  const unsigned char* decay;  // n entries: 1 where the step cap is lowered after the iteration (agd.py:102-109)
  int n;
  long long seq_base;          // sharded: exchange step number of iteration i is seq_base + i + 1
};
__device__ __forceinline__ StepDyn step_dyn_of(const AgdStepArgs& A) {
  return StepDyn{A.beta, A.decay_now, A.iter_index, A.gamma, A.log_obj != nullptr};
}
__device__ __forceinline__ StepDyn step_dyn_sched(const AgdStepArgs& A, const SchedArgs& S, long long it) {
  const int i = (int)(it < (long long)S.n ? it : (long long)S.n - 1);
  return StepDyn{S.beta[i], (int)S.decay[i], (int)it, S.gamma[i], A.log_obj != nullptr && it < (long long)A.log_cap};
}

// Shared-memory scratch of the m-length tail / step code.  ONE object per kernel, handed to every helper: static __shared__
// arrays inside templated device functions would be replicated per instantiation, and the slab kernel's dynamic shared
// memory is sized up to the device limit minus a fixed allowance for its static part.
struct TailScratch {
  double red[5][32];
  double tot[8];
  double step;
  float mx[32];
};

template <bool FROM_PARTIAL>
__device__ __forceinline__ void agd_step_body(const AgdStepArgs& A, const StepDyn& D, TailScratch& T) {
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
  const float beta = D.beta;
  const int decay_now = D.decay_now, iter_index = D.iter_index;
  const double decay_factor = A.decay_factor, gamma = D.gamma;
  double* log_obj = D.log ? A.log_obj : nullptr;
  double* log_step = D.log ? A.log_step : nullptr;
  const float* __restrict__ b = A.b;
  float* grad_out = A.grad_out;
  dualip_scalars* __restrict__ scal_out = A.scal_out;
  double (&s_red)[5][32] = T.red;
  float (&s_mx)[32] = T.mx;
  double& s_step = T.step;
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
  int* status;       // local device word: set to 1 when a wait timed out (checked before every wait)
  int* status_host;  // the same flag in mapped host memory, written on a time-out only: the host polls it without a sync
  int push;          // one-launch path: 1 = sums are pushed into every peer's window (stores), 0 = peers pull them (loads)
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


// The exchange by ONE CTA (the slab kernel's last CTA, after it has written this rank's packed sums into its slot): tell
// every peer that the slot is ready, wait until all peers have said the same, fetch all slots -- W float4 loads in flight
// per thread -- and add them in rank order, so that every rank computes bit-identical sums.  Result in P.sum.
__device__ __forceinline__ void peer_exchange_cta(const PeerArgs& P, int m2, unsigned long long seq) {
  const int tid = threadIdx.x, nt = blockDim.x;
  __threadfence_system();  // the slot written by this CTA's threads is visible system-wide before the flag
  __syncthreads();
  if (tid < P.world) {
    st_release_sys_u64(reinterpret_cast<unsigned long long*>(P.win[tid]) + P.rank, seq);
    const unsigned long long* mine = reinterpret_cast<const unsigned long long*>(P.win[P.rank]) + tid;
    const unsigned long long t0 = global_timer_ns();
    // once a wait has timed out the run is invalid (the host raises): later steps must not wait the full time-out again
    while (*reinterpret_cast<volatile int*>(P.status) == 0 && ld_relaxed_sys_u64(mine) < seq) {
      if (global_timer_ns() - t0 > P.timeout_ns) {
        *reinterpret_cast<volatile int*>(P.status) = 1;
        *reinterpret_cast<volatile int*>(P.status_host) = 1;
        break;
      }
    }
    asm volatile("fence.acq_rel.sys;" ::: "memory");  // the peers' slots are read after their flags
  }
  __syncthreads();
  const size_t slot_off = (size_t)kPeerFlagBytes + (size_t)(seq & 1ull) * P.slot_bytes;
  for (int i4 = tid * 4; i4 < m2; i4 += nt * 4) {  // slots are padded to 128 bytes: a float4 never leaves the slot
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
  __syncthreads();  // P.sum is read by other threads of this CTA
}

// Push variant used by the slab kernel's fused tail.  The CTA has ALREADY written this rank's packed sums into slot
// [parity][rank] of EVERY rank's window (posted stores over NVLink: no read round trip, peer_push_slot() gives the address);
// here it tells the peers so, waits for theirs, and adds the W slots of its OWN window -- local memory -- in rank order.
// One NVLink one-way latency per step instead of ceil(m / (4 threads)) dependent remote-read rounds.
__device__ __forceinline__ float* peer_push_slot(const PeerArgs& P, int target, int writer, unsigned long long seq) {
  return reinterpret_cast<float*>(P.win[target] + kPeerFlagBytes + 2 * P.slot_bytes +
                                  ((size_t)(seq & 1ull) * (size_t)P.world + (size_t)writer) * P.slot_bytes);
}
__device__ __forceinline__ void peer_push_exchange_cta(const PeerArgs& P, int m2, unsigned long long seq) {
  const int tid = threadIdx.x, nt = blockDim.x;
  __threadfence_system();  // this CTA's stores into the peers' windows are performed before the flags
  __syncthreads();
  if (tid < P.world) {
    st_release_sys_u64(reinterpret_cast<unsigned long long*>(P.win[tid]) + P.rank, seq);
    const unsigned long long* mine = reinterpret_cast<const unsigned long long*>(P.win[P.rank]) + tid;
    const unsigned long long t0 = global_timer_ns();
    while (*reinterpret_cast<volatile int*>(P.status) == 0 && ld_relaxed_sys_u64(mine) < seq) {
      if (global_timer_ns() - t0 > P.timeout_ns) {
        *reinterpret_cast<volatile int*>(P.status) = 1;
        *reinterpret_cast<volatile int*>(P.status_host) = 1;
        break;
      }
    }
    asm volatile("fence.acq_rel.sys;" ::: "memory");  // the slots the peers wrote are read after their flags
  }
  __syncthreads();
  for (int i4 = tid * 4; i4 < m2; i4 += nt * 4) {  // slots are padded to 128 bytes: a float4 never leaves the slot
    float4 v[DUALIP_PEER_MAX_WORLD];
#pragma unroll
    for (int r = 0; r < DUALIP_PEER_MAX_WORLD; ++r)
      if (r < P.world) v[r] = ld_relaxed_sys_f4(peer_push_slot(P, P.rank, r, seq) + i4);
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
  __syncthreads();  // P.sum is read by other threads of this CTA
}

}  // namespace dualip

// ---- host-side state behind the opaque C handles (shared by agd.cu and calc.cu) ----
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
  long long launched = 0;  // steps enqueued so far (host count; the device's `pushes` reaches it when the stream drains)
  // device-resident schedule (dualip_agd_set_schedule)
  double* sched_gamma = nullptr;
  float* sched_beta = nullptr;
  unsigned char* sched_decay = nullptr;
  int sched_n = 0;
  double sched_factor = 1.0;
};

struct dualip_peer {
  int device = 0, m = 0, rank = 0, world = 1;
  size_t slot_bytes = 0, window_bytes = 0;
  unsigned char* window = nullptr;                    // own window (cudaMalloc: exportable through CUDA IPC)
  unsigned char* win[DUALIP_PEER_MAX_WORLD] = {};     // all windows as mapped here
  bool opened[DUALIP_PEER_MAX_WORLD] = {};            // mapped with cudaIpcOpenMemHandle
  bool connected = false;
  float* sum = nullptr;
  int* status = nullptr;       // device word: a wait timed out
  int* status_host = nullptr;  // the same flag in mapped host memory (host address)
  int* status_host_dev = nullptr;  // its device address
  unsigned int* ticket = nullptr;  // ticket of the multi-CTA step kernel
  unsigned long long seq = 0;  // steps taken
  unsigned long long timeout_ns = 20ull * 1000ull * 1000ull * 1000ull;  // DUALIP_PEER_TIMEOUT_MS overrides
  int push = 1;  // DUALIP_PEER_PUSH=0: the one-launch path pulls like dualip_agd_step_peer
};

static inline dualip::AgdStepArgs step_args(dualip_agd* a, const float* grad, const dualip_scalars* scal, float beta, int decay_now,
                             double decay_factor, int iter_index, const float* b, double gamma, float* grad_out,
                             dualip_scalars* scal_out) {
  const bool log = iter_index >= 0 && iter_index < a->log_cap;
  dualip::AgdStepArgs A;
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
  A.log_cap = a->log_cap;
  ++a->launched;
  return A;
}


// Kernel arguments of one exchange step; `advance` takes the next step number (the flag value; its parity selects the slot).
static inline dualip::PeerArgs peer_args(dualip_peer* p, bool advance) {
  dualip::PeerArgs P;
  for (int r = 0; r < DUALIP_PEER_MAX_WORLD; ++r) P.win[r] = r < p->world ? p->win[r] : nullptr;
  P.rank = p->rank;
  P.world = p->world;
  P.seq = advance ? ++p->seq : p->seq;
  P.slot_bytes = p->slot_bytes;
  P.timeout_ns = p->timeout_ns;
  P.sum = p->sum;
  P.status = p->status;
  P.status_host = p->status_host_dev;
  P.ticket = p->ticket;
  P.push = p->push;
  return P;
}

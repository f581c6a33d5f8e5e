// Grid-wide tail of the fused iteration.  Included by calc.cu after agd_step.cuh (needs KArgs).
//
// The one-launch iteration used to end on ONE CTA: the CTA that finished last ran the m-length tail of the objective
// (grad = sums - b, lambda.grad, slacks), the exchange with the peers, and the accelerated step -- three to five dependent
// passes over m-vectors by 512 threads while 147 SMs idle: ~16 us at m = 10 000, ~60 us at m = 26 744, a fifth of a small
// iteration.  All CTAs of the slab kernel are co-resident (one per SM, never more CTAs than SMs), so the tail can be done by
// ALL of them: every CTA owns a slice of ceil(m / CTAs) rows, and the few points where the slices depend on each other are
// grid-wide barriers (an arrive counter and a generation word in global memory, bounded spins):
//
//   B1  every CTA's flush into the global accumulators is complete
//       [sharded: each CTA stores its slice of the shard's sums into every peer's window; B2; CTA 0 raises the peers' arrival
//        flags and waits for theirs; B3; each CTA adds its slice of the W slots of its own window in rank order]
//       slice: grad = sums - b, the objective's partial sums, history push, the newest Lipschitz pair's partial sums
//   B_last  per-CTA partials are added in a FIXED order by every CTA (deterministic: replicas of a sharded run stay
//       bit-identical), every CTA derives the same step size, CTA 0 writes scalars / ratio ring / logs / step cap
//       slice: y_new = proj(x + step * grad), x = y_new (1 - beta) + y beta
//
// Arithmetic per element is that of cta_epilogue + agd_step_body (reference optimizers/agd.py:163-187, agd_utils.py:4-89,
// objectives/matching.py:25-34,164-178); only the order in which the double-precision partial sums are added differs.
#pragma once

namespace dualip {

constexpr int kTailPart = 8;  // doubles per CTA in the partials table: lg, sp, g2, dg2, dy2, mx

// status / status_host: the same failure word in device memory (read here, so that after one time-out the later barriers of
// the run do not wait again) and in mapped host memory (written on a time-out only; the host reads it without a sync).
// timeout_ns: 4 s for a grid on its own; a sharded grid also waits (in B3) for CTA 0's wait on the peers, which is bounded by
// the exchange's own time-out, so its barriers allow that much more.
__device__ __forceinline__ void grid_barrier(unsigned int* bar, unsigned int n, int* status, int* status_host,
                                             unsigned long long timeout_ns) {
  __syncthreads();
  if (threadIdx.x == 0 && n > 1) {
    volatile unsigned int* gen = bar + 1;
    const unsigned int g = *gen;
    const int failed = status ? *reinterpret_cast<volatile int*>(status) : 0;
    __threadfence();
    if (atomicAdd(bar, 1u) == n - 1) {
      bar[0] = 0u;
      __threadfence();
      atomicAdd(bar + 1, 1u);
    } else if (!failed) {
      const unsigned long long t0 = global_timer_ns();
      while (*gen == g) {
        if (global_timer_ns() - t0 > timeout_ns) {  // a CTA of this grid never arrived (not co-resident?)
          if (status) *reinterpret_cast<volatile int*>(status) = 2;
          if (status_host) *reinterpret_cast<volatile int*>(status_host) = 2;
          break;
        }
      }
    }
    __threadfence();
  }
  __syncthreads();
}

// STEP false: evaluation only (dualip_matching_calc / dualip_matching_calc_peer: the host-buffer path) -- the tail without
// the optimizer's part; the evaluation point is k.lambda and the outputs are k.grad_out / k.scalars_out.
// c.x and ||x||^2 of this launch: the slab kernel's per-CTA slots added in CTA order (warp 0 / warp 1: lanes stride over the
// CTAs, then a fixed shuffle tree) on top of what the column kernels launched before it left in acc_scal.  Every thread of the
// calling CTA gets the result; all CTAs that call it compute the same bits.
__device__ __forceinline__ void load_scalar_sums(const KArgs& k, TailScratch& T, double& cxv, double& xxv) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp < 2) {
    double acc = 0.0;
    for (int c = lane; c < (int)gridDim.x; c += 32) acc += __ldcg(k.tail_part + (size_t)c * kTailPart + 6 + warp);
    acc = warp_sum(acc);
    if (lane == 0) T.tot[6 + warp] = acc + __ldcg(&k.acc_scal[warp]);
  }
  __syncthreads();
  cxv = T.tot[6];
  xxv = T.tot[7];
  __syncthreads();
}

template <bool SHARDED, bool STEP, typename SumFn, typename ClearFn>
__device__ __forceinline__ void grid_tail(const KArgs& k, SumFn sum_load, ClearFn sum_clear, const StepDyn D, double gamma,
                                          unsigned long long seq, TailScratch& T) {
  double (&s_red)[5][32] = T.red;
  float (&s_mx)[32] = T.mx;
  double (&s_tot)[8] = T.tot;
  double& s_step = T.step;
  const AgdStepArgs& A = k.agd;
  const unsigned FULL = 0xffffffffu;
  const int tid = threadIdx.x, nt = blockDim.x, nb = gridDim.x, bid = blockIdx.x;
  const int lane = tid & 31, warp = tid >> 5, nw = (nt + 31) >> 5;
  const int m = k.m, H = STEP ? A.H : 2;  // (evaluation only: the optimizer arguments are not set)
  const int S = (m + nb - 1) / nb;
  const int r0 = min(m, bid * S), r1 = min(m, r0 + S);
  int* status = SHARDED ? k.peer.status : k.grid_status;
  int* status_host = SHARDED ? k.peer.status_host : k.grid_status_host;
  const unsigned long long bar_timeout = 4000000000ull + (SHARDED ? k.peer.timeout_ns : 0ull);
  // optimizer state that CTA 0 changes at the very end: read by everyone before the first barrier
  const long long t = STEP ? __ldcg(A.pushes) : 0;
  const int slot = (int)(t % H), prev = (int)((t + H - 1) % H);
  const bool have_prev = STEP && t > 0;
  const double max_step = STEP ? __ldcg(&A.dstate[0]) : 0.0, init_step = STEP ? __ldcg(&A.dstate[1]) : 0.0;
  float* __restrict__ x = A.x;
  float* __restrict__ y = A.y;
  float* __restrict__ gh = A.gh;
  float* __restrict__ yh = A.yh;
  const float* __restrict__ lam_p = STEP ? A.x : k.lambda;  // the evaluation point
  float* grad_out = (SHARDED && STEP) ? A.grad_out : k.grad_out;
  dualip_scalars* scal_out = (SHARDED && STEP) ? A.scal_out : k.scalars_out;
  const float* __restrict__ b = (SHARDED && STEP) ? A.b : k.b;

  grid_barrier(k.grid_bar, nb, status, status_host, bar_timeout);  // B1: the accumulators hold this rank's complete sums
  double cx_local, xx_local;
  load_scalar_sums(k, T, cx_local, xx_local);
  double cxv = cx_local, xxv = xx_local;
  if (SHARDED) {
    const PeerArgs& P = k.peer;
    for (int i = r0 + tid; i < r1; i += nt) {
      const float raw = sum_load(i);
      sum_clear(i);
#pragma unroll
      for (int r = 0; r < DUALIP_PEER_MAX_WORLD; ++r)
        if (r < P.world) peer_push_slot(P, r, P.rank, seq)[i] = raw;  // posted stores over NVLink, from every SM
    }
    if (bid == 0 && tid == 0) {
#pragma unroll
      for (int r = 0; r < DUALIP_PEER_MAX_WORLD; ++r)
        if (r < P.world) {
          float* ps = peer_push_slot(P, r, P.rank, seq);
          ps[m] = (float)cx_local;
          ps[m + 1] = (float)xx_local;
        }
    }
    __threadfence_system();
    grid_barrier(k.grid_bar, nb, status, status_host, bar_timeout);  // B2: every CTA's stores into the peers' windows are performed
    if (bid == 0) {
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
        asm volatile("fence.acq_rel.sys;" ::: "memory");
      }
    }
    grid_barrier(k.grid_bar, nb, status, status_host, bar_timeout);  // B3: every peer's sums are in this rank's window
    asm volatile("fence.acq_rel.sys;" ::: "memory");
    if (tid == 0) {  // the two scalars, added in rank order like the rows
      float c0 = 0.f, c1 = 0.f;
      for (int r = 0; r < P.world; ++r) {
        const float* ps = peer_push_slot(P, P.rank, r, seq);
        const float v0 = ld_relaxed_sys_f32(ps + m), v1 = ld_relaxed_sys_f32(ps + m + 1);
        c0 = r == 0 ? v0 : __fadd_rn(c0, v0);
        c1 = r == 0 ? v1 : __fadd_rn(c1, v1);
      }
      s_tot[0] = (double)c0;
      s_tot[1] = (double)c1;
    }
    __syncthreads();
    cxv = s_tot[0];
    xxv = s_tot[1];
    __syncthreads();
  }

  // ---- this CTA's rows: gradient, objective partials, history push, newest Lipschitz pair ----
  double lg = 0.0, sp = 0.0, g2 = 0.0, dg2 = 0.0, dy2 = 0.0;
  float mx = -INFINITY;
  for (int i = r0 + tid; i < r1; i += nt) {
    float tot;
    if (SHARDED) {
      const PeerArgs& P = k.peer;
      tot = ld_relaxed_sys_f32(peer_push_slot(P, P.rank, 0, seq) + i);
      for (int r = 1; r < P.world; ++r) tot = __fadd_rn(tot, ld_relaxed_sys_f32(peer_push_slot(P, P.rank, r, seq) + i));
    } else {
      tot = sum_load(i);
      sum_clear(i);
    }
    const float g = b ? __fsub_rn(tot, __ldg(b + i)) : tot;
    grad_out[i] = g;
    const float lam = lam_p[i];
    lg = fma((double)lam, (double)g, lg);
    sp += (double)fmaxf(g, 0.f);
    g2 = fma((double)g, (double)g, g2);
    mx = fmaxf(mx, g);
    if (STEP) {
      const float yv = y[i];
      if (have_prev) {
        const float dg = __fsub_rn(gh[(size_t)prev * m + i], g);
        const float dy = __fsub_rn(yh[(size_t)prev * m + i], yv);
        dg2 = fma((double)dg, (double)dg, dg2);
        dy2 = fma((double)dy, (double)dy, dy2);
      }
      gh[(size_t)slot * m + i] = g;
      yh[(size_t)slot * m + i] = yv;
    }
  }
  lg = warp_sum(lg), sp = warp_sum(sp), g2 = warp_sum(g2), dg2 = warp_sum(dg2), dy2 = warp_sum(dy2);
  mx = warp_max(mx);
  if (lane == 0) {
    s_red[0][warp] = lg, s_red[1][warp] = sp, s_red[2][warp] = g2, s_red[3][warp] = dg2, s_red[4][warp] = dy2;
    s_mx[warp] = mx;
  }
  __syncthreads();
  if (warp == 0) {
    lg = warp_sum(lane < nw ? s_red[0][lane] : 0.0);
    sp = warp_sum(lane < nw ? s_red[1][lane] : 0.0);
    g2 = warp_sum(lane < nw ? s_red[2][lane] : 0.0);
    dg2 = warp_sum(lane < nw ? s_red[3][lane] : 0.0);
    dy2 = warp_sum(lane < nw ? s_red[4][lane] : 0.0);
    mx = warp_max(lane < nw ? s_mx[lane] : -INFINITY);
    if (lane == 0) {
      double* mine = k.tail_part + (size_t)bid * kTailPart;
      mine[0] = lg, mine[1] = sp, mine[2] = g2, mine[3] = dg2, mine[4] = dy2, mine[5] = (double)mx;
    }
  }
  __threadfence();
  grid_barrier(k.grid_bar, nb, status, status_host, bar_timeout);  // B_last: every CTA's partials are in the table

  // ---- totals, in the same fixed order on every CTA (and on every rank) ----
  if (warp < 6) {
    double acc = (warp == 5) ? -INFINITY : 0.0;
    for (int c = lane; c < nb; c += 32) {
      const double v = __ldcg(k.tail_part + (size_t)c * kTailPart + warp);
      acc = (warp == 5) ? fmax(acc, v) : acc + v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double v = __shfl_xor_sync(FULL, acc, o);
      acc = (warp == 5) ? fmax(acc, v) : acc + v;
    }
    if (lane == 0) s_tot[warp] = acc;
  }
  __syncthreads();
  if (tid == 0) {
    lg = s_tot[0], sp = s_tot[1], g2 = s_tot[2], dg2 = s_tot[3], dy2 = s_tot[4];
    mx = (float)s_tot[5];
    dualip_scalars r;
    r.primal_objective = cxv;
    r.reg_penalty = 0.5 * gamma * xxv;
    r.dual_val_times_grad = lg;
    r.dual_objective = cxv + r.reg_penalty + lg;
    r.max_pos_slack = (double)fmaxf(mx, 0.f);
    r.sum_pos_slack = sp;
    r.x_sq_norm = xxv;
    r.grad_sq_norm = g2;
    if (!STEP) {
      if (bid == 0) {
        *scal_out = r;
        k.acc_scal[0] = 0.0;
        k.acc_scal[1] = 0.0;
      }
    } else {
      const float rnew = have_prev ? __fdiv_rn((float)sqrt(dg2), (float)sqrt(dy2)) : 0.f;
      // step size (agd_utils.py:44-62): Python max() over the ratios in chronological order; the newest one is rnew
      const long long n_pairs = t < (long long)(H - 1) ? t : (long long)(H - 1);
      double step = init_step;
      if (n_pairs >= H - 1) {
        const long long j0 = t - (H - 1);
        float lmax = (j0 == t - 1) ? rnew : A.ratios[j0 % (H - 1)];
        for (long long j = j0 + 1; j < t; ++j) {
          const float v = (j == t - 1) ? rnew : A.ratios[j % (H - 1)];
          if (v > lmax) lmax = v;
        }
        if (!(isnan(lmax) || isinf(lmax))) {
          const double cand = (lmax != 0.f) ? 1.0 / (double)lmax : max_step;
          step = cand < max_step ? cand : max_step;
        }
      }
      s_step = step;
      if (bid == 0) {  // one writer for everything that is not sliced
        *scal_out = r;
        if (have_prev) A.ratios[(t - 1) % (H - 1)] = rnew;
        if (D.log) {
          A.log_obj[D.iter_index] = r.dual_objective;
          A.log_step[D.iter_index] = step;
        }
        if (D.decay_now) A.dstate[0] = step * A.decay_factor;  // agd.py:107
        *A.pushes = t + 1;
        k.acc_scal[0] = 0.0;
        k.acc_scal[1] = 0.0;
      }
    }
  }
  if (!STEP) return;
  __syncthreads();
  // ---- ascent step, projection on the dual cone, momentum on this CTA's rows (agd.py:181-185, :13-21) ----
  const float step32 = (float)s_step;
  const float beta = D.beta, omb = __fsub_rn(1.0f, beta);
  const uint8_t* __restrict__ eqmask = A.eqmask;
  for (int i = r0 + tid; i < r1; i += nt) {
    const float g = grad_out[i], yv = y[i], xv = x[i];
    float yn = __fadd_rn(xv, __fmul_rn(g, step32));
    if (!(eqmask && eqmask[i])) yn = fmaxf(yn, 0.f);
    x[i] = __fadd_rn(__fmul_rn(yn, omb), __fmul_rn(yv, beta));
    y[i] = yn;
  }
}

}  // namespace dualip

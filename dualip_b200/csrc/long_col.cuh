// CTA-per-column processing of long columns (more than kMaxThreadDeg entries, up to kLongStash).  Included by calc.cu after
// mid_col.cuh.  Same reference semantics as every other path (matching.py:116-188, simplex.py:143-236, box.py:16,
// cone.py:21-28).
//
// The first long-column kernel gave a WARP a column and re-streamed it from L2 once per round of the threshold search: a
// 4000-entry column is 140 dependent load rounds per sweep, and the kernel took longer than everything else of a
// MovieLens-shaped iteration together.  Here 256 threads own a column: every thread holds len/256 entries, u = max(v, 0) is
// written once to a shared-memory stash, all rounds of the search read the stash, and a/c/row are read a second time only for
// the entries with x != 0.  The scatter goes to the plan's global accumulators (the slab kernel, launched afterwards on the
// same stream, folds them into its flush and runs the m-length tail), as before.
#pragma once

namespace dualip {

constexpr int kLongThreads = 256;
constexpr int kLongStash = 12288;  // floats of u per CTA (48 KB: four CTAs per SM); longer columns keep the warp-per-column kernel
// The same kernel with a TEAM of 32 threads per column (eight columns per CTA, 4 KB of stash each): mid columns (up to
// kMaxThreadDeg entries) of plans that hold MANY of them.  Inside the slab kernel such columns are a chain of dependent
// reductions at 16 warps per SM (128 registers per thread); here nothing but a few scalars lives in registers, so 40 warps per
// SM keep two and a half times as many columns in flight.
constexpr int kTeamStash = 1024;   // floats of u per 32-thread team

struct LongRed {  // one block-wide reduction round
  double s;
  int c;
  float mn, mx;
};

// All threads of a team obtain the team-wide {sum s, sum c, min mn, max mx}.  TEAM == 32: warp shuffles only.  TEAM ==
// kLongThreads: `scratch` holds 2 x 8 entries; consecutive calls alternate between the halves, so one barrier per call
// suffices (a thread cannot be two calls ahead of another).
template <int TEAM>
__device__ __forceinline__ LongRed long_team_reduce(LongRed v, LongRed* scratch, int& phase) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v.s = warp_sum(v.s);
  v.c = __reduce_add_sync(0xffffffffu, v.c);
  v.mn = warp_min_f(v.mn);
  v.mx = warp_max(v.mx);
  if (TEAM == 32) return v;
  LongRed* buf = scratch + phase * (kLongThreads / 32);
  if (lane == 0) buf[warp] = v;
  __syncthreads();
  LongRed r = buf[0];
#pragma unroll
  for (int w = 1; w < kLongThreads / 32; ++w) {
    r.s += buf[w].s;
    r.c += buf[w].c;
    r.mn = fminf(r.mn, buf[w].mn);
    r.mx = fmaxf(r.mx, buf[w].mx);
  }
  phase ^= 1;
  return r;
}
template <int TEAM>
__device__ __forceinline__ void team_sync() {
  if (TEAM == 32)
    __syncwarp();
  else
    __syncthreads();
}

template <int ACC, int TEAM>
__global__ void __launch_bounds__(kLongThreads, TEAM == 32 ? 5 : 1) matching_long_cta_kernel(const KArgs k, const LongCol* __restrict__ cols,
                                                                                          int n_cols) {
  extern __shared__ __align__(16) unsigned char long_smem[];
  constexpr int kTeams = kLongThreads / TEAM;  // columns in flight per CTA
  // the team's stash: kLongStash floats for a whole-CTA team, kTeamStash per 32-thread team
  float* s_u = reinterpret_cast<float*>(long_smem) + (TEAM == 32 ? (threadIdx.x / TEAM) * kTeamStash : 0);
  __shared__ LongRed s_red[2 * (kLongThreads / 32)];
  __shared__ float s_m1[kLongThreads / 32], s_m2[kLongThreads / 32];
  __shared__ int s_am[kLongThreads / 32];
  __shared__ float s_seq;
  __shared__ double s_scal[32];
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int tid = threadIdx.x % TEAM;  // position inside the team
  int phase = 0;
  float s_run = k.s;
  if (k.sched.gamma != nullptr) {  // scheduled launch: see matching_slab_kernel
    const long long it = __ldcg(k.agd.pushes);
    s_run = (float)(-1.0 / __ldg(k.sched.gamma + (it < (long long)k.sched.n ? it : (long long)k.sched.n - 1)));
  }
  const float s = s_run;
  double cx = 0.0, xx = 0.0;
  auto lam_of = [&](uint32_t r) -> float {
    float ls = __fmul_rn(s, __ldg(k.lambda + r));
    if (k.row_unscale != nullptr) ls = __fmul_rn(ls, __ldg(k.row_unscale + r));  // long_a holds a * 2^k_r
    return ls;
  };
  auto scatter = [&](float a, float c, uint32_t r, float x) {
    const float g = __fmul_rn(a, x);
    if (ACC == 1) {
      const long long gi = __double2ll_rn((double)g * (double)k.fx_scale);
      if (gi != 0) {
        atomicAdd(&k.acc_lo[r], (int)(gi & 0xffff));
        atomicAdd(&k.acc_hi[r], (int)(gi >> 16));
      }
    } else if (g != 0.f) {
      atomicAdd(&k.acc[r], g);
    }
    cx = fma((double)c, (double)x, cx);
    xx = fma((double)x, (double)x, xx);
  };

  for (int ci = blockIdx.x * kTeams + threadIdx.x / TEAM; ci < n_cols; ci += gridDim.x * kTeams) {
    const LongCol lc = cols[ci];
    const dualip_proj_class pc = k.classes[lc.cls];
    const int len = lc.len;
    const float* __restrict__ pa = k.long_a + lc.off;
    const float* __restrict__ pcv = k.long_c + lc.off;
    const uint32_t* __restrict__ pr = k.long_row + lc.off;
    if (pc.kind == DUALIP_PROJ_CLAMP) {
#pragma unroll 4
      for (int e = tid; e < len; e += TEAM) {
        const float a = __ldg(pa + e), c = __ldg(pcv + e);
        const uint32_t r = __ldg(pr + e);
        const float x = fminf(fmaxf(make_v(a, lam_of(r), s, c), pc.lo), pc.hi);
        scatter(a, c, r, x);
        if (k.x_out) k.x_out[lc.src_start + e] = x;
      }
      continue;
    }
    // ---- simplex / simplex_eq: u into the stash; column sum, the two largest values and the position of the largest ----
    team_sync<TEAM>();  // the previous column's stash is no longer read
    const float z = pc.z;
    double Sd = 0.0;
    float m1 = -1.f, m2 = -1.f;
    int am = 0x7fffffff;
#pragma unroll 4
    for (int e = tid; e < len; e += TEAM) {
      const float a = __ldg(pa + e), c = __ldg(pcv + e);
      const uint32_t r = __ldg(pr + e);
      const float u = fmaxf(make_v(a, lam_of(r), s, c), 0.f);  // simplex.py:148
      s_u[e] = u;
      Sd += (double)u;
      if (u > m1) am = e;
      m2 = fmaxf(m2, fminf(m1, u));
      m1 = fmaxf(m1, u);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float o1 = __shfl_xor_sync(FULL, m1, o), o2 = __shfl_xor_sync(FULL, m2, o);
      const int oa = __shfl_xor_sync(FULL, am, o);
      const float mn = fminf(m1, o1);
      if (o1 > m1 || (o1 == m1 && oa < am)) am = oa;
      m1 = fmaxf(m1, o1);
      m2 = fmaxf(fmaxf(m2, o2), mn);
    }
    if (TEAM != 32 && lane == 0) s_m1[warp] = m1, s_m2[warp] = m2, s_am[warp] = am;
    LongRed r0 = long_team_reduce<TEAM>(LongRed{Sd, 0, 0.f, 0.f}, s_red, phase);  // its barrier also publishes s_m1 / s_m2 / s_am and s_u
    Sd = r0.s;
    if (TEAM == 32) {
      __syncwarp();  // the team's stash is complete
    } else {
      m1 = s_m1[0], m2 = s_m2[0], am = s_am[0];
#pragma unroll
      for (int w = 1; w < kLongThreads / 32; ++w) {
        const float o1 = s_m1[w], o2 = s_m2[w];
        const int oa = s_am[w];
        const float mn = fminf(m1, o1);
        if (o1 > m1 || (o1 == m1 && oa < am)) am = oa;
        m1 = fmaxf(m1, o1);
        m2 = fmaxf(fmaxf(m2, o2), mn);
      }
    }
    // feasibility (simplex.py:153-155): the reference compares its fp32 entry-order sum; away from the threshold (further than
    // the worst rounding error of such a sum) the fp64 sum decides, inside the band one thread forms the sequential sum
    bool feasible = false;
    if (pc.kind == DUALIP_PROJ_SIMPLEX) {
      const double thr = (double)pc.z_thr;
      if (fabs(Sd - thr) > 2.2 * (double)len * 5.97e-8 * fmax(Sd, thr)) {
        feasible = Sd <= thr;
      } else {
        if (TEAM == 32) {  // every lane forms the same sequential sum from the stash (broadcast reads)
          float S = 0.f;
          for (int e = 0; e < len; ++e) S = __fadd_rn(S, s_u[e]);
          feasible = S <= pc.z_thr;
        } else {
          if (tid == 0) {
            float S = 0.f;
            for (int e = 0; e < len; ++e) S = __fadd_rn(S, s_u[e]);
            s_seq = S;
          }
          __syncthreads();
          feasible = s_seq <= pc.z_thr;
        }
      }
    }
    const float m2p = fmaxf(m2, 0.f);
    const float un1 = (z == 1.0f) ? m1 : __fdiv_rn(m1, z), un2 = (z == 1.0f) ? m2p : __fdiv_rn(m2p, z);
    const bool shortcut = !feasible && (__fsub_rn(un1, un2) > 1.0f);  // simplex.py:166-178
    const float t_below = (pc.kind == DUALIP_PROJ_SIMPLEX_EQ) ? __fsub_rn((float)Sd, z) : 0.f;
    int branch, rho = 0;
    float theta = 0.f;
    if (pc.kind == DUALIP_PROJ_SIMPLEX_EQ && t_below < 0.f) {
      branch = 2;
      rho = pad_len_of(k, lc.cls, len);
      theta = __fdiv_rn(t_below, (float)rho);
    } else if (feasible) {
      branch = 0;
    } else if (shortcut) {
      branch = 1;
      rho = 1;
    } else {
      branch = 2;
      // Michelot's fixed point from below (see mid_col.cuh), sums in fp64, every step rounded down
      float tf = fmaxf(__double2float_rd((Sd - (double)z) / (double)len), __fsub_rd(m1, z));
      tf = (tf > 0.f) ? prev_float(tf) : -1.f;
      int cnt = 0;
      for (int it = 0; it < 64; ++it) {
        LongRed v{0.0, 0, INFINITY, -INFINITY};
        for (int e = tid; e < len; e += TEAM) {
          const float u = s_u[e];
          const bool in = u > tf;
          v.c += in ? 1 : 0;
          v.s += in ? (double)u : 0.0;
          v.mn = in ? fminf(v.mn, u) : v.mn;
        }
        v = long_team_reduce<TEAM>(v, s_red, phase);
        cnt = v.c;
        if (v.c == 0) break;
        const float tn = __double2float_rd((v.s - (double)z) / (double)v.c);
        if (!(tn > tf) || v.mn > tn) break;
        tf = tn;
      }
      // the reference's own fp32 conditions on the support and its two boundary values (simplex.py:207-231)
      float th = 0.f;
      for (int fix = 0; fix < 64; ++fix) {
        LongRed v{0.0, 0, INFINITY, -INFINITY};
        for (int e = tid; e < len; e += TEAM) {
          const float u = s_u[e];
          const bool in = u > tf;
          v.c += in ? 1 : 0;
          v.s += in ? (double)u : 0.0;
          v.mn = in ? fminf(v.mn, u) : v.mn;
          v.mx = in ? v.mx : fmaxf(v.mx, u);
        }
        v = long_team_reduce<TEAM>(v, s_red, phase);
        cnt = v.c;
        th = __fdiv_rn(__fsub_rn((float)v.s, z), (float)max(v.c, 1));
        bool changed = false;
        if (v.c > 1 && !(__fsub_rn(v.mn, th) > 0.f)) {
          tf = fmaxf(v.mn, fminf(th, prev_float(m1)));
          changed = true;
        } else if (v.mx > -INFINITY) {
          const float t1 = __fdiv_rn(__fsub_rn((float)(v.s + (double)v.mx), z), (float)(v.c + 1));
          if (__fsub_rn(v.mx, t1) > 0.f) {
            tf = (v.mx > 0.f) ? prev_float(v.mx) : -1.f;
            changed = true;
          }
        }
        if (!changed) break;
      }
      theta = th;
      rho = max(cnt, 1);
    }
    // ---- x and the scatter: a, c and the row id are read again only where x != 0 ----
    for (int e = tid; e < len; e += TEAM) {
      const float u = s_u[e];
      const float x = branch == 0 ? u : (branch == 1 ? (e == am ? z : 0.f) : fmaxf(__fsub_rn(u, theta), 0.f));
      if (x != 0.f) scatter(__ldg(pa + e), __ldg(pcv + e), __ldg(pr + e), x);
      if (k.x_out) k.x_out[lc.src_start + e] = x;
    }
    if (k.diag && tid == 0) k.diag[lc.src_start] = (uint8_t)(branch | (min(rho, 63) << 2));
  }
  cx = block_sum(cx, s_scal);
  xx = block_sum(xx, s_scal);
  if (tid == 0) {
    if (cx != 0.0) atomicAdd(&k.acc_scal[0], cx);
    if (xx != 0.0) atomicAdd(&k.acc_scal[1], xx);
  }
}

}  // namespace dualip

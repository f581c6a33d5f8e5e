// Warp-per-column processing of "mid" columns: longer than the register path handles (kRegDeg) and up to kMaxThreadDeg
// entries.  Included by calc.cu after slab_fast.cuh.  Reference semantics as everywhere: matching.py:116-188 per column with
// projections/simplex.py:143-236, box.py:16, cone.py:21-28.
//
// Such columns used to share the lane-per-column slab layout: a lane then walks its column serially and the threshold search
// re-streams it from L2 once per round -- a single slab of 1000-entry columns kept a warp busy for milliseconds (MovieLens-
// shaped data, mean degree 120-140, ran at 1 % of the HBM roofline).  Here the plan keeps these columns contiguous (the
// compact arrays of the long-column kernel) and a WARP owns a column: lane l holds entries l, l+32, ... in registers (at most 32
// per lane), every load is a coalesced 128-byte request, the column is read from memory twice (projection input, then a, c
// and the row ids again for the scatter: an L1/L2 hit) and all rounds of the threshold search run on registers with warp
// reductions.  The code runs inside matching_slab_kernel after a CTA's slab range, so lambda, the fixed-point accumulator, the
// flush, the m-length tail and the fused optimizer step are shared: still one launch per iteration.
#pragma once

namespace dualip {

__device__ __forceinline__ float warp_min_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float prev_float(float x) {  // largest float below x, for x > 0
  return __uint_as_float(__float_as_uint(x) - 1u);
}

// One column of `len` entries (len <= 32 * NPL).  cx / xx: this lane's running c.x and ||x||^2 partials.
template <int NPL, int ACC, bool OUT>
__device__ __forceinline__ void mid_column(const KArgs& k, const LongCol& lc, const dualip_proj_class& pc, int lane,
                                           const float* __restrict__ s_lam, float* __restrict__ s_grad, float s, double& cx,
                                           double& xx) {
  const unsigned FULL = 0xffffffffu;
  const int len = lc.len;
  const float* __restrict__ pa = k.long_a + lc.off;
  const float* __restrict__ pcv = k.long_c + lc.off;
  const uint32_t* __restrict__ pr = k.long_row + lc.off;
  float cxs = 0.f, xxs = 0.f;
  auto emit = [&](float a, float c, uint32_t r, float x, int e) {
    const float g = __fmul_rn(a, x);  // matching.py:153
    if (ACC == 1) {
      const int gi = __float2int_rn(g * k.fx_scale);
      if (gi != 0) atomicAdd(reinterpret_cast<int*>(s_grad) + r, gi);
    } else if (g != 0.f) {
      atomicAdd(&s_grad[r], g);
    }
    cxs = fmaf(c, x, cxs);
    xxs = fmaf(x, x, xxs);
    if (OUT && k.x_out) k.x_out[lc.src_start + e] = x;
  };

  if (pc.kind == DUALIP_PROJ_CLAMP) {
    // ---- box / cone / identity: one pass ----
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
      const int e = lane + 32 * i;
      if (e < len) {
        const float a = __ldg(pa + e), c = __ldg(pcv + e);
        const uint32_t r = __ldg(pr + e);
        emit(a, c, r, fminf(fmaxf(make_v(a, s_lam[r], s, c), pc.lo), pc.hi), e);
      }
    }
    cx += (double)cxs;
    xx += (double)xxs;
    return;
  }

  // ---- simplex / simplex_eq ----
  const float z = pc.z;
  float u[NPL];
  double Sd = 0.0;
  float m1 = -1.f, m2 = -1.f;  // two largest values of this lane (duplicates of the maximum count)
  int am = 0x7fffffff;
#pragma unroll
  for (int i = 0; i < NPL; ++i) {
    const int e = lane + 32 * i;
    u[i] = 0.f;  // positions past the column's end behave like the reference's zero padding
    if (e < len) {
      const float a = __ldg(pa + e), c = __ldg(pcv + e);
      const uint32_t r = __ldg(pr + e);
      u[i] = fmaxf(make_v(a, s_lam[r], s, c), 0.f);  // simplex.py:148
      if (u[i] > m1) am = e;
    }
    Sd += (double)u[i];
    m2 = fmaxf(m2, fminf(m1, u[i]));
    m1 = fmaxf(m1, u[i]);
  }
  Sd = warp_sum(Sd);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float o1 = __shfl_xor_sync(FULL, m1, o), o2 = __shfl_xor_sync(FULL, m2, o);
    const int oa = __shfl_xor_sync(FULL, am, o);
    const float mn = fminf(m1, o1);
    if (o1 > m1 || (o1 == m1 && oa < am)) am = oa;
    m1 = fmaxf(m1, o1);
    m2 = fmaxf(fmaxf(m2, o2), mn);
  }
  // Feasibility (simplex.py:153-155) compares the fp32 column sum accumulated in entry order with fl32(z + 1e-6).  A sum of
  // up to 1024 non-negative floats differs from the exact one by at most 1023 * 2^-24 relative: away from the threshold the
  // fp64 tree sum decides; within that band the warp forms the reference's sequential sum itself.
  bool feasible = false;
  if (pc.kind == DUALIP_PROJ_SIMPLEX) {
    const double thr = (double)pc.z_thr;
    if (fabs(Sd - thr) > 1.3e-4 * fmax(Sd, thr)) {
      feasible = Sd <= thr;
    } else {
      float S = 0.f;
#pragma unroll
      for (int i = 0; i < NPL; ++i) {
#pragma unroll 1
        for (int l = 0; l < 32; ++l) S = __fadd_rn(S, __shfl_sync(FULL, u[i], l));  // entry l + 32 i, in order (rare: not unrolled)
      }
      feasible = S <= pc.z_thr;
    }
  }
  const float m2p = fmaxf(m2, 0.f);
  const float un1 = (z == 1.0f) ? m1 : __fdiv_rn(m1, z), un2 = (z == 1.0f) ? m2p : __fdiv_rn(m2p, z);
  const bool shortcut = !feasible && (__fsub_rn(un1, un2) > 1.0f);  // simplex.py:166-178 (columns here have L > 1)
  const float t_below = (pc.kind == DUALIP_PROJ_SIMPLEX_EQ) ? __fsub_rn((float)Sd, z) : 0.f;
  int branch, rho = 0;
  float theta = 0.f;
  if (pc.kind == DUALIP_PROJ_SIMPLEX_EQ && t_below < 0.f) {
    // every zero-padded position of the reference's block satisfies cond_i: rho = L, theta = (css_d - z)/L < 0 (App. A #4)
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
    // Michelot's fixed point from below on registers: t <- (sum_{u > t} u - z) / #{u > t}; the sequence increases to theta*,
    // and theta* >= max - z and >= (S - z)/d.  Sums in fp64; every step is rounded DOWN so that no step overshoots.
    float tf = fmaxf(__double2float_rd((Sd - (double)z) / (double)len), __fsub_rd(m1, z));
    tf = (tf > 0.f) ? prev_float(tf) : -1.f;
    int cnt = 0;
    for (int it = 0; it < 64; ++it) {
      double ssum = 0.0;
      int c2 = 0;
      float umin = INFINITY;
#pragma unroll
      for (int i = 0; i < NPL; ++i) {
        const bool in = u[i] > tf;
        c2 += in ? 1 : 0;
        ssum += in ? (double)u[i] : 0.0;
        umin = in ? fminf(umin, u[i]) : umin;
      }
      ssum = warp_sum(ssum);
      c2 = __reduce_add_sync(FULL, c2);
      umin = warp_min_f(umin);
      cnt = c2;
      if (c2 == 0) break;
      const float tn = __double2float_rd((ssum - (double)z) / (double)c2);
      if (!(tn > tf) || umin > tn) break;  // the step removes nothing: converged
      tf = tn;
    }
    // exact sums over the support and its two boundary values, then the reference's own fp32 conditions (simplex.py:207-231):
    // css_rho = fl32(fp64 prefix sum), cond_i = u_(i) - fl32((css_i - z)/i) > 0, rho = max{i: cond_i}
    float th = 0.f;
    for (int fix = 0; fix < 64; ++fix) {
      double ssum = 0.0;
      int c2 = 0;
      float umin = INFINITY, uout = -INFINITY;
#pragma unroll
      for (int i = 0; i < NPL; ++i) {
        const bool in = u[i] > tf;
        c2 += in ? 1 : 0;
        ssum += in ? (double)u[i] : 0.0;
        umin = in ? fminf(umin, u[i]) : umin;
        uout = in ? uout : fmaxf(uout, u[i]);
      }
      ssum = warp_sum(ssum);
      c2 = __reduce_add_sync(FULL, c2);
      umin = warp_min_f(umin);
      uout = warp_max(uout);
      cnt = c2;
      th = __fdiv_rn(__fsub_rn((float)ssum, z), (float)max(c2, 1));
      bool changed = false;
      if (c2 > 1 && !(__fsub_rn(umin, th) > 0.f)) {
        // cond_rho fails in the fp32 formula: every support value <= th leaves (the largest stays: th < max unless rounding
        // at huge magnitudes, hence the clamp)
        tf = fmaxf(umin, fminf(th, prev_float(m1)));
        changed = true;
      } else if (uout > -INFINITY) {
        const float t1 = __fdiv_rn(__fsub_rn((float)(ssum + (double)uout), z), (float)(c2 + 1));
        if (__fsub_rn(uout, t1) > 0.f) {  // cond_{rho+1} holds: the support grows
          tf = (uout > 0.f) ? prev_float(uout) : -1.f;
          changed = true;
        }
      }
      if (!changed) break;
    }
    theta = th;
    rho = max(cnt, 1);
  }

  // ---- second read of a, c and the row ids (just streamed: L1 / L2), scatter ----
#pragma unroll
  for (int i = 0; i < NPL; ++i) {
    const int e = lane + 32 * i;
    if (e < len) {
      const float x = branch == 0 ? u[i] : (branch == 1 ? (e == am ? z : 0.f) : fmaxf(__fsub_rn(u[i], theta), 0.f));
      if (x != 0.f || (OUT && k.x_out)) {
        const float a = __ldg(pa + e), c = __ldg(pcv + e);
        const uint32_t r = __ldg(pr + e);
        emit(a, c, r, x, e);
      }
    }
  }
  if (OUT && k.diag && lane == 0) k.diag[lc.src_start] = (uint8_t)(branch | (min(rho, 63) << 2));
  cx += (double)cxs;
  xx += (double)xxs;
}

// All mid columns of this CTA: warp w takes columns begin + w, begin + w + NW, ...; the instantiation follows the column's
// length (the list is sorted by length, so consecutive columns of a warp take the same one).
template <int ACC, bool OUT, int NW>
__device__ __forceinline__ void mid_columns_of_cta(const KArgs& k, const LongCol* __restrict__ cols, int begin, int end, int warp,
                                                   int lane, const dualip_proj_class* s_cls, const float* s_lam, float* s_grad,
                                                   float s, double& cx, double& xx) {
  // The work on a column is a chain of dependent steps (header -> a/c/row -> lambda gather -> reductions), and a warp has
  // only 15 others to hide behind: the header of the column after next is loaded, and the data of the next column is pulled
  // into L2 (one prefetch per lane covers up to 4 KB of each array), while the current column is processed.
  const LongCol none = {0, 0, 0, 0};
  LongCol nxt = (begin + warp < end) ? cols[begin + warp] : none;
  LongCol nxt2 = (begin + warp + NW < end) ? cols[begin + warp + NW] : none;
  for (int ci = begin + warp; ci < end; ci += NW) {
    const LongCol lc = nxt;
    nxt = nxt2;
    if (ci + NW < end) {
      const size_t span = (size_t)nxt.len * 4 + 127;
      if ((size_t)lane * 128 < span) {
        asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(k.long_a + nxt.off) + lane * 128));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(k.long_c + nxt.off) + lane * 128));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(k.long_row + nxt.off) + lane * 128));
      }
    }
    nxt2 = (ci + 2 * NW < end) ? cols[ci + 2 * NW] : none;
    const dualip_proj_class pc = s_cls[lc.cls];
    const int npl = (lc.len + 31) >> 5;
    if (npl <= 1)
      mid_column<1, ACC, OUT>(k, lc, pc, lane, s_lam, s_grad, s, cx, xx);
    else if (npl <= 2)
      mid_column<2, ACC, OUT>(k, lc, pc, lane, s_lam, s_grad, s, cx, xx);
    else if (npl <= 4)
      mid_column<4, ACC, OUT>(k, lc, pc, lane, s_lam, s_grad, s, cx, xx);
    else if (npl <= 8)
      mid_column<8, ACC, OUT>(k, lc, pc, lane, s_lam, s_grad, s, cx, xx);
    else if (npl <= 16)
      mid_column<16, ACC, OUT>(k, lc, pc, lane, s_lam, s_grad, s, cx, xx);
    else
      mid_column<32, ACC, OUT>(k, lc, pc, lane, s_lam, s_grad, s, cx, xx);
  }
}

}  // namespace dualip

// Register-resident processing of one slab (32 columns of equal length D <= kRegDeg), specialised on D.
// Included by calc.cu after KArgs / make_v.  Reference semantics: matching.py:116-188 per column, with
// projections/simplex.py:143-236 (batched Duchi with pre-clamp), box.py:16, cone.py:21-28.
//
// One load phase: D/4 LDG.128 for a, D/4 LDG.128 for c, D/4 LDG.64 for the uint16 row ids (plus the short tail
// chunks), everything issued before the first use.  The column then lives in registers: u = max(v, 0), the
// feasibility / top-2 tests, the threshold search and the scatter all run on fully unrolled register arrays, so
// there is no second pass over memory, no stash traffic and no per-entry index arithmetic.
#pragma once

namespace dualip {

constexpr int kRegDeg = 20;   // longest column handled by the register path
constexpr int kKeepDeg = 16;  // up to here a and c stay in registers across the simplex scan; longer columns drop them
                              // after u is formed and read them again (L1/L2 hit) for the scatter: the sorted copy,
                              // u, the row offsets, a and c (5 D registers) would not fit 128 registers

// ---- streaming loads: the slab arrays are read exactly once per launch, so they bypass L1 allocation ----
__device__ __forceinline__ float4 ldg_stream_f4(const float* p) {
  float4 v;
  asm("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float2 ldg_stream_f2(const float* p) {
  float2 v;
  asm("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ float ldg_stream_f1(const float* p) {
  float v;
  asm("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ uint2 ldg_stream_u2(const void* p) {
  uint2 v;
  asm("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ uint32_t ldg_stream_u1(const void* p) {
  uint32_t v;
  asm("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ uint32_t ldg_stream_h1(const void* p) {
  unsigned short v;
  asm("ld.global.nc.L1::no_allocate.u16 %0, [%1];" : "=h"(v) : "l"(p));
  return (uint32_t)v;
}
// One thread asks the TMA engine to pull a contiguous range into L2 (no register or shared-memory destination).
__device__ __forceinline__ void prefetch_l2_bulk(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

__device__ __forceinline__ float lds_off(const unsigned char* base, uint32_t byte_off) {
  return *reinterpret_cast<const float*>(base + byte_off);
}
// Pins a value in a register (the compiler would otherwise rematerialise shared-window addresses at every use).
__device__ __forceinline__ uint32_t pin_u32(uint32_t v) {
  uint32_t r;
  asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(v));
  return r;
}

// Sorting networks on register arrays (descending), generated and zero-one verified by tools/gen_sort_networks.py.
template <int D>
struct SortNet {
  static_assert(D == 1, "sort_networks.inc must cover 2..kRegDeg (tools/gen_sort_networks.py)");
  static __device__ __forceinline__ void run(float (&)[D]) {}  // D = 1
};
#define DUALIP_CS(i, j)                     \
  {                                         \
    const float hi_ = fmaxf(w[i], w[j]);    \
    w[j] = fminf(w[i], w[j]);               \
    w[i] = hi_;                             \
  }
#include "sort_networks.inc"
#undef DUALIP_CS

// fl32(t / n) for a small positive integer n, correctly rounded without the division routine: q0 = t * fl(1/n),
// exact residual r = t - q0*n by FMA, q = q0 + r * fl(1/n) (Markstein's correction step).
template <int N>
__device__ __forceinline__ float div_by_int(float t) {
  constexpr float rn = 1.0f / (float)N;
  const float q0 = __fmul_rn(t, rn);
  const float r = __fmaf_rn(-q0, (float)N, t);
  return __fmaf_rn(r, rn, q0);
}

// Compile-time loop with the index available as a template constant.
template <int I, int N, typename F>
__device__ __forceinline__ void static_for(F&& f) {
  if constexpr (I < N) {
    f(std::integral_constant<int, I>{});
    static_for<I + 1, N>(f);
  }
}

// A lane's column.  Row ids are kept as byte offsets (row * 4) into an m-float shared-memory array; columns longer than
// kKeepDeg keep them packed instead, two uint16 per register as loaded, and unpack at each use (two instructions).
template <int D>
struct ColRegs {
  static constexpr bool PACKED = D > kKeepDeg;
  float a[D], c[D];
  uint32_t rr[PACKED ? (D + 1) / 2 : D];
  // W-th 32-bit word of the column's row ids (rows 2W and 2W+1)
  template <int W>
  __device__ __forceinline__ void set_word(uint32_t word) {
    if constexpr (PACKED) {
      rr[W] = word;
    } else {
      rr[2 * W] = (word << 2) & 0x3fffcu;
      if constexpr (2 * W + 1 < D) rr[2 * W + 1] = (word >> 14) & 0x3fffcu;
    }
  }
  template <int Q>
  __device__ __forceinline__ uint32_t off() const {
    if constexpr (PACKED)
      return (Q & 1) ? ((rr[Q / 2] >> 14) & 0x3fffcu) : ((rr[Q / 2] << 2) & 0x3fffcu);
    else
      return rr[Q];
  }
};

template <int D>
__device__ __forceinline__ void load_cols(ColRegs<D>& R, const float* __restrict__ a_s, const float* __restrict__ c_s,
                                          const unsigned short* __restrict__ r_s, int lane) {
  // a_s / c_s / r_s point at the slab's first element
  constexpr int NF = D / 4;
  float4 va[NF > 0 ? NF : 1], vc[NF > 0 ? NF : 1];
  uint2 vr[NF > 0 ? NF : 1];
#pragma unroll
  for (int q = 0; q < NF; ++q) {
    va[q] = ldg_stream_f4(a_s + q * 128 + lane * 4);
    vc[q] = ldg_stream_f4(c_s + q * 128 + lane * 4);
    vr[q] = ldg_stream_u2(r_s + q * 128 + lane * 4);
  }
  float2 ta = make_float2(0.f, 0.f), tc = make_float2(0.f, 0.f);
  uint32_t tr = 0;
  if (D & 2) {
    ta = ldg_stream_f2(a_s + NF * 128 + lane * 2);
    tc = ldg_stream_f2(c_s + NF * 128 + lane * 2);
    tr = ldg_stream_u1(r_s + NF * 128 + lane * 2);
  }
  float sa = 0.f, sc = 0.f;
  uint32_t sr = 0;
  if (D & 1) {
    sa = ldg_stream_f1(a_s + NF * 128 + (D & 2) * 32 + lane);
    sc = ldg_stream_f1(c_s + NF * 128 + (D & 2) * 32 + lane);
    sr = ldg_stream_h1(r_s + NF * 128 + (D & 2) * 32 + lane);
  }
#pragma unroll
  for (int q = 0; q < NF; ++q) {
    R.a[4 * q + 0] = va[q].x;
    R.a[4 * q + 1] = va[q].y;
    R.a[4 * q + 2] = va[q].z;
    R.a[4 * q + 3] = va[q].w;
    R.c[4 * q + 0] = vc[q].x;
    R.c[4 * q + 1] = vc[q].y;
    R.c[4 * q + 2] = vc[q].z;
    R.c[4 * q + 3] = vc[q].w;
  }
  static_for<0, NF>([&](auto qc) {
    constexpr int q = decltype(qc)::value;
    R.template set_word<2 * q>(vr[q].x);
    R.template set_word<2 * q + 1>(vr[q].y);
  });
  if constexpr ((D & 2) != 0) {
    R.a[4 * NF + 0] = ta.x;
    R.a[4 * NF + 1] = ta.y;
    R.c[4 * NF + 0] = tc.x;
    R.c[4 * NF + 1] = tc.y;
    R.template set_word<2 * NF>(tr);
  }
  if constexpr ((D & 1) != 0) {
    R.a[D - 1] = sa;
    R.c[D - 1] = sc;
    R.template set_word<(D - 1) / 2>(sr);
  }
}

// Second read of a and c (columns longer than kKeepDeg).  volatile: must not be merged with the first read, which
// would keep the registers alive in between.
template <int D>
__device__ __forceinline__ void reload_ac(ColRegs<D>& R, const float* __restrict__ a_s, const float* __restrict__ c_s, int lane) {
  constexpr int NF = D / 4;
#pragma unroll
  for (int q = 0; q < NF; ++q) {
    float4 va, vc;
    asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(va.x), "=f"(va.y), "=f"(va.z), "=f"(va.w) : "l"(a_s + q * 128 + lane * 4));
    asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(vc.x), "=f"(vc.y), "=f"(vc.z), "=f"(vc.w) : "l"(c_s + q * 128 + lane * 4));
    R.a[4 * q + 0] = va.x, R.a[4 * q + 1] = va.y, R.a[4 * q + 2] = va.z, R.a[4 * q + 3] = va.w;
    R.c[4 * q + 0] = vc.x, R.c[4 * q + 1] = vc.y, R.c[4 * q + 2] = vc.z, R.c[4 * q + 3] = vc.w;
  }
  if (D & 2) {
    asm volatile("ld.global.nc.v2.f32 {%0,%1}, [%2];" : "=f"(R.a[4 * NF]), "=f"(R.a[4 * NF + 1]) : "l"(a_s + NF * 128 + lane * 2));
    asm volatile("ld.global.nc.v2.f32 {%0,%1}, [%2];" : "=f"(R.c[4 * NF]), "=f"(R.c[4 * NF + 1]) : "l"(c_s + NF * 128 + lane * 2));
  }
  if (D & 1) {
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(R.a[D - 1]) : "l"(a_s + NF * 128 + (D & 2) * 32 + lane));
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(R.c[D - 1]) : "l"(c_s + NF * 128 + (D & 2) * 32 + lane));
  }
}

// Same layout, read from this warp's shared-memory staging buffer (filled by the TMA engine, see stage_issue()):
// [a: 32*D floats][c: 32*D floats][row: 32*D uint16].  LDS.128 / LDS.64 with lane-contiguous addresses: conflict-free.
__device__ __forceinline__ float4 lds_f4(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr) : "memory");
  return v;
}
__device__ __forceinline__ float2 lds_f2(uint32_t saddr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(saddr) : "memory");
  return v;
}
__device__ __forceinline__ float lds_f1(uint32_t saddr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr) : "memory");
  return v;
}
__device__ __forceinline__ uint2 lds_u2(uint32_t saddr) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(saddr) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t lds_u1(uint32_t saddr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t lds_h1(uint32_t saddr) {
  unsigned short v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(saddr) : "memory");
  return (uint32_t)v;
}
template <int D>
__device__ __forceinline__ void load_cols_staged(ColRegs<D>& R, uint32_t buf, int lane) {
  constexpr int NF = D / 4;
  const uint32_t a_s = buf + lane * 16, c_s = buf + 128 * D + lane * 16, r_s = buf + 256 * D + lane * 8;
  static_for<0, NF>([&](auto qc) {
    constexpr int q = decltype(qc)::value;
    const float4 va = lds_f4(a_s + q * 512);
    const float4 vc = lds_f4(c_s + q * 512);
    const uint2 vr = lds_u2(r_s + q * 256);
    R.a[4 * q + 0] = va.x, R.a[4 * q + 1] = va.y, R.a[4 * q + 2] = va.z, R.a[4 * q + 3] = va.w;
    R.c[4 * q + 0] = vc.x, R.c[4 * q + 1] = vc.y, R.c[4 * q + 2] = vc.z, R.c[4 * q + 3] = vc.w;
    R.template set_word<2 * q>(vr.x);
    R.template set_word<2 * q + 1>(vr.y);
  });
  if constexpr ((D & 2) != 0) {
    const float2 ta = lds_f2(buf + NF * 512 + lane * 8);
    const float2 tc = lds_f2(buf + 128 * D + NF * 512 + lane * 8);
    R.a[4 * NF + 0] = ta.x, R.a[4 * NF + 1] = ta.y;
    R.c[4 * NF + 0] = tc.x, R.c[4 * NF + 1] = tc.y;
    R.template set_word<2 * NF>(lds_u1(buf + 256 * D + NF * 256 + lane * 4));
  }
  if constexpr ((D & 1) != 0) {
    R.a[D - 1] = lds_f1(buf + (NF * 128 + (D & 2) * 32 + lane) * 4);
    R.c[D - 1] = lds_f1(buf + 128 * D + (NF * 128 + (D & 2) * 32 + lane) * 4);
    R.template set_word<(D - 1) / 2>(lds_h1(buf + 256 * D + (NF * 128 + (D & 2) * 32 + lane) * 2));
  }
}

// Staging of one of a warp's upcoming slabs: one lane arms the slot's mbarrier with the slab's byte count and issues ONE
// bulk async copy (TMA engine, SASS UBLKCP) global -> shared; it lands while the warp works on earlier slabs.
struct StageCtx {
  bool use_stage;
  uint32_t region;        // bytes of this warp's staging region
  unsigned char* base;    // its first byte
  uint64_t* bars;         // its two slot barriers
};
__device__ __forceinline__ void stage_issue(unsigned char* buf, uint64_t* bar, const unsigned char* slab, uint32_t bytes) {
  fence_proxy_async_smem();  // the warp's reads of the slot (generic proxy) precede the engine's writes
  mbar_expect_tx(bar, bytes);
  bulk_g2s(buf, slab, bytes, bar);
}

// Shared tail of both projections: scatter a*x into the CTA's gradient accumulator and form the c.x and ||x||^2
// partials.  ACC 1: the accumulator is 32-bit fixed point (value * 2^F, F chosen at plan time from a bound on every
// CTA's row sums): one native fire-and-forget ATOMS.ADD per non-zero, order-independent and therefore bitwise
// reproducible.  ACC 0: fp32 atomicAdd, which sm_100a implements as a load / add / compare-and-swap loop
// (ATOMS.CAST.SPIN); kept for unbounded projection classes (open cones), where no overflow bound exists.
template <int D, int SMODE, int ACC>
__device__ __forceinline__ void emit_cols(const KArgs& k, const ColRegs<D>& R, const float (&x)[D], uint32_t s_grad_u32,
                                          float& cxs, float& xxs) {
  static_for<0, D>([&](auto qc) {
    constexpr int q = decltype(qc)::value;
    const float g = __fmul_rn(R.a[q], x[q]);  // matching.py:153 (A.values * x.values, then row sums)
    if (ACC == 1) {
      const int gi = __float2int_rn(g * k.fx_scale);
#ifdef DUALIP_COND_ADD
      if (gi != 0)
#endif
      // unconditional: adding 0 is harmless, and ptxas would wrap a conditional ATOMS in a branch (4 instructions
      // instead of 1); the extra shared-memory wavefronts fit the MIO budget (profiles/r1_ubench_smem2.txt)
      asm volatile("red.shared.add.s32 [%0], %1;" ::"r"(s_grad_u32 + R.template off<q>()), "r"(gi) : "memory");
    } else if (g != 0.f) {
      if (SMODE <= 1)
        asm volatile("red.shared.add.f32 [%0], %1;" ::"r"(s_grad_u32 + R.template off<q>()), "f"(g) : "memory");
      else
        atomicAdd(k.acc + (R.template off<q>() >> 2), g);
    }
    cxs = fmaf(R.c[q], x[q], cxs);
    xxs = fmaf(x[q], x[q], xxs);
  });
}

template <int D, int SMODE>
__device__ __forceinline__ void make_v_cols(const KArgs& k, const ColRegs<D>& R, const unsigned char* s_lam_b, float s,
                                            float (&v)[D]) {
  static_for<0, D>([&](auto qc) {
    constexpr int q = decltype(qc)::value;
    const uint32_t ro = R.template off<q>();
    const float lam_s = (SMODE == 0) ? lds_off(s_lam_b, ro) : __fmul_rn(s, __ldg(k.lambda + (ro >> 2)));
    v[q] = make_v(R.a[q], lam_s, s, R.c[q]);
  });
}

// box / cone / identity (box.py:16, cone.py:21-28): x = min(max(v, lo), hi).
template <int D, int SMODE>
__device__ __forceinline__ void fast_clamp(const KArgs& k, const dualip_proj_class& pc, const ColRegs<D>& R, bool active,
                                           const unsigned char* s_lam_b, float s, float (&x)[D]) {
  const float lo = active ? pc.lo : 0.f, hi = active ? pc.hi : 0.f;  // padding lanes produce x = 0
  make_v_cols<D, SMODE>(k, R, s_lam_b, s, x);
#pragma unroll
  for (int q = 0; q < D; ++q) x[q] = fminf(fmaxf(x[q], lo), hi);
}

// The reference's scan over the sorted column.  Zero entries cannot satisfy cond_i once the prefix sum has reached z
// (q_i = (css_i - z)/i >= 0 = u_(i)), and they sit at the end of the sorted order, so the scan may stop at a position where
// no lane of the warp that needs theta has a positive entry left.  `more` carries that per lane; lanes whose column sum is
// within rounding distance of z (css_i >= z not certain: `near`) keep the warp scanning to the end, which reproduces the
// reference there too.  The vote is taken at two positions (about 0.6 D and 0.8 D: at late iterates the longest support
// in a warp is about D/2), not at every step, because it costs three instructions.
template <int D, int I>
__device__ __forceinline__ void scan_sorted(const float (&w)[D], float z, double& acc, float& t_sel, int& rho_sel, bool need,
                                            bool near) {
  if constexpr (I < D) {
    constexpr int kStop1 = (3 * D + 4) / 5, kStop2 = (4 * D + 4) / 5;
    if constexpr (D >= 5 && (I == kStop1 || (I == kStop2 && kStop2 != kStop1))) {
      if (!__any_sync(0xffffffffu, need && (near || w[I] > 0.f))) return;
    }
    acc += (double)w[I];
    const float t = __fsub_rn((float)acc, z);
    const bool cond = w[I] > div_by_int<I + 1>(t);
    t_sel = cond ? t : t_sel;
    rho_sel = cond ? (I + 1) : rho_sel;
    scan_sorted<D, I + 1>(w, z, acc, t_sel, rho_sel, need, near);
  }
}

// simplex / simplex_eq (simplex.py:143-236 per column at its true length).  On return x holds the projection;
// branch: 0 feasible, 1 top-2 shortcut, 2 sorted scan ("Duchi"); rho: support size for branches 1 and 2.
template <int D, int SMODE>
__device__ __forceinline__ void fast_simplex(const KArgs& k, const dualip_proj_class& pc, int cls, ColRegs<D>& R, bool active,
                                             const unsigned char* s_lam_b, float s, float (&u)[D], int& branch, int& rho) {
  const unsigned FULL = 0xffffffffu;
  make_v_cols<D, SMODE>(k, R, s_lam_b, s, u);
  const float z = pc.z;
#ifdef DUALIP_SORT_FIRST
  // Late iterates: nearly every warp holds a column that needs the sorted scan, so the sort is not gated on the top-2
  // test; the two largest values are then read off the sorted copy instead of being tracked entry by entry.
  float S = 0.f;
#pragma unroll
  for (int q = 0; q < D; ++q) {
    u[q] = fmaxf(u[q], 0.f);             // simplex.py:148
    S = __fadd_rn(S, u[q]);              // column sum in entry order
  }
  const bool feasible = (pc.kind == DUALIP_PROJ_SIMPLEX) && (S <= pc.z_thr);                          // simplex.py:153-155
  bool shortcut = false;
  float m1 = 0.f;
  float theta = 0.f;
  bool below = false;
  float t_below = 0.f;
  branch = 0;
  rho = 0;
  if (__any_sync(FULL, active && !feasible)) {
    float w[D];
#pragma unroll
    for (int q = 0; q < D; ++q) w[q] = u[q];
    SortNet<D>::run(w);
    m1 = w[0];
    const float m2p = (D > 1) ? w[1] : 0.f;  // the reference's zero padding takes part in its top-2
    const bool padded = (D > 1) || !(pc.flags & DUALIP_PROJ_FLAG_D1_UNPADDED);                        // simplex.py:166
    const float un1 = (z == 1.0f) ? m1 : __fdiv_rn(m1, z), un2 = (z == 1.0f) ? m2p : __fdiv_rn(m2p, z);
    shortcut = !feasible && padded && (__fsub_rn(un1, un2) > 1.0f);                                   // simplex.py:172-178
    branch = feasible ? 0 : (shortcut ? 1 : 2);
    rho = shortcut ? 1 : 0;
    if (pc.kind == DUALIP_PROJ_SIMPLEX_EQ) {  // see below
      double sd = 0.0;
#pragma unroll
      for (int q = 0; q < D; ++q) sd += (double)u[q];
      t_below = __fsub_rn((float)sd, z);
      below = t_below < 0.f;
    }
    const bool need_theta = active && branch == 2 && !below;
    if (__any_sync(FULL, need_theta)) {
      double acc = 0.0;
      float t_sel = __fsub_rn(w[0], z);
      int rho_sel = 1;
      const bool near = !(S > __fmul_rn(z, 1.0001f));
      scan_sorted<D, 0>(w, z, acc, t_sel, rho_sel, need_theta, near);
      if (need_theta) {
        theta = __fdiv_rn(t_sel, (float)rho_sel);                                                     // simplex.py:228-230
        rho = rho_sel;
      }
    }
  }
#else
  float S = 0.f, m1 = -1.f, m2 = -1.f;
#pragma unroll
  for (int q = 0; q < D; ++q) {
    u[q] = fmaxf(u[q], 0.f);             // simplex.py:148
    S = __fadd_rn(S, u[q]);              // column sum in entry order
    m2 = fmaxf(m2, fminf(m1, u[q]));     // second largest (duplicates of the maximum count)
    m1 = fmaxf(m1, u[q]);
  }
  const bool feasible = (pc.kind == DUALIP_PROJ_SIMPLEX) && (S <= pc.z_thr);                          // simplex.py:153-155
  const bool padded = (D > 1) || !(pc.flags & DUALIP_PROJ_FLAG_D1_UNPADDED);                          // simplex.py:166
  const float m2p = fmaxf(m2, 0.f);  // the reference's zero padding takes part in its top-2
  const float un1 = (z == 1.0f) ? m1 : __fdiv_rn(m1, z), un2 = (z == 1.0f) ? m2p : __fdiv_rn(m2p, z);
  const bool shortcut = !feasible && padded && (__fsub_rn(un1, un2) > 1.0f);                          // simplex.py:172-178
  branch = feasible ? 0 : (shortcut ? 1 : 2);
  rho = shortcut ? 1 : 0;
  float theta = 0.f;
  // simplex_eq with a clamped sum below z: every zero-padded position of the reference's block satisfies cond_i, so
  // rho = L (the bucket's padded length) and theta = (css_d - z)/L < 0 (simplex.py:160-161,207-233).
  // The shortcut needs u_(1) > z, which excludes it.  Uniform per slab: only simplex_eq slabs pay for the double sum.
  bool below = false;
  float t_below = 0.f;
  if (pc.kind == DUALIP_PROJ_SIMPLEX_EQ) {
    double sd = 0.0;
#pragma unroll
    for (int q = 0; q < D; ++q) sd += (double)u[q];
    t_below = __fsub_rn((float)sd, z);
    below = t_below < 0.f;
  }
  const bool need_theta = active && branch == 2 && !below;

  if (__any_sync(FULL, need_theta)) {
    // ---- the reference's sorted scan itself (simplex.py:207-231), on a sorted copy of the column in registers ----
    //   css_i = fl32(prefix sum accumulated in fp64)            (torch's CPU cumsum of float32 accumulates in double)
    //   cond_i = u_(i) - fl((css_i - z)/i) > 0  <=>  u_(i) > fl((css_i - z)/i)
    //   rho = max{i : cond_i},  theta = (css_rho - z)/rho
    // A sorting network (no branches, no data-dependent trip counts) replaces torch.sort; every lane of the warp runs
    // it, lanes that do not need theta ignore the result.
    float w[D];
#pragma unroll
    for (int q = 0; q < D; ++q) w[q] = u[q];
    SortNet<D>::run(w);
    double acc = 0.0;
    float t_sel = __fsub_rn(w[0], z);
    int rho_sel = 1;  // no cond true: torch's max over an all-zero mask gives index 0
    // lanes that do not need theta never hold the warp in the scan; a lane whose sum is not clearly above z (within 1e-4 z,
    // or below it: simplex_eq) scans everything
    const bool near = !(S > __fmul_rn(z, 1.0001f));
    scan_sorted<D, 0>(w, z, acc, t_sel, rho_sel, need_theta, near);
    if (need_theta) {
      theta = __fdiv_rn(t_sel, (float)rho_sel);                                                     // simplex.py:228-230
      rho = rho_sel;
    }
  }
#endif
  if (below) {
    rho = pad_len_of(k, cls, D);
    theta = __fdiv_rn(t_below, (float)rho);
  }
  if (__any_sync(FULL, shortcut)) {
    // x = z at the (unique) maximum, 0 elsewhere (simplex.py:185-190); theta stays 0 so the subtraction below is exact
#pragma unroll
    for (int q = 0; q < D; ++q) u[q] = shortcut ? ((u[q] == m1) ? z : 0.f) : u[q];
  }
  if (!active) theta = INFINITY;  // padding lanes (and nothing else) produce x = 0
#pragma unroll
  for (int q = 0; q < D; ++q) u[q] = fmaxf(__fsub_rn(u[q], theta), 0.f);                             // simplex.py:233
}

// One slab of column length D, whole life cycle.  `slab`: the slab in global memory ([a][c][row ids] back to back); when the
// slab is staged the same bytes are waiting in (or on their way to) one of the warp's slots.  ensure_issued(t) makes sure
// the copy of slab t has been requested before the warp waits for it; after_load(t) is called as soon as the slot has been
// read out, to request further slabs.
template <int D, int SMODE, int ACC, bool OUT, typename Ensure, typename AfterLoad>
__device__ __forceinline__ void fast_slab(const KArgs& k, const dualip_proj_class& pc, int cls, const unsigned char* slab,
                                          int lane, bool active, const unsigned char* s_lam_b, uint32_t s_grad_u32, float s,
                                          int slab_index, int t, double& cx, double& xx, const StageCtx& st, uint32_t& phases,
                                          Ensure ensure_issued, AfterLoad after_load) {
  constexpr uint32_t kBytes = 320u * D;  // uint16 row ids (the register path exists for them only)
  const float* __restrict__ a_s = reinterpret_cast<const float*>(slab);
  const float* __restrict__ c_s = a_s + 32 * D;
  const unsigned short* __restrict__ r_s = reinterpret_cast<const unsigned short*>(c_s + 32 * D);
  ColRegs<D> R;
  if (st.use_stage && kBytes <= st.region) {
    const bool two_deep = kBytes * 2u <= st.region;
    const int slot = two_deep ? (t & 1) : 0;
    ensure_issued(t);
    mbar_wait(st.bars + slot, (phases >> slot) & 1u);
    phases ^= 1u << slot;
    load_cols_staged<D>(R, smem_u32(st.base) + (uint32_t)slot * (st.region >> 1), lane);
  } else {
    load_cols<D>(R, a_s, c_s, r_s, lane);
  }
  after_load(t);
  float x[D];
  int branch = -1, rho = 0;
  if (pc.kind == DUALIP_PROJ_CLAMP)
    fast_clamp<D, SMODE>(k, pc, R, active, s_lam_b, s, x);
  else
    fast_simplex<D, SMODE>(k, pc, cls, R, active, s_lam_b, s, x, branch, rho);
  if (OUT) {  // save_primal / diagnostics: straight from registers
    if (active) {
      const int64_t os = k.orig_start[(int64_t)slab_index * 32 + lane];
      if (k.x_out) {
#pragma unroll
        for (int q = 0; q < D; ++q) k.x_out[os + q] = x[q];
      }
      if (k.diag && branch >= 0) k.diag[os] = (uint8_t)(branch | (min(rho, 63) << 2));
    }
  }
  if constexpr (D > kKeepDeg) {
    if (pc.kind != DUALIP_PROJ_CLAMP) reload_ac<D>(R, a_s, c_s, lane);
  }
  float cxs = 0.f, xxs = 0.f;
  emit_cols<D, SMODE, ACC>(k, R, x, s_grad_u32, cxs, xxs);
  cx += (double)cxs;
  xx += (double)xxs;
}

}  // namespace dualip

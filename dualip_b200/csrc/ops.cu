// Stand-alone CSC operators of the reference's public extension recipe, and the fairness-row objective built with them.
//
// The reference documents how users extend the matching objective (docs/demo/matching_complex.rst:82-168): a subclass
// overrides calculate() and composes left_multiply_sparse / elementwise_csc / apply_F_to_columns / row_sums_csc
// (src/dualip/utils/sparse_utils.py:26-51,54-85,133-220,223-243) with calc_grad (objectives/matching.py:25-34).  The fused
// kernel of calc.cu replaces that chain for the stock objective; the kernels below are the device implementations of the
// individual operators, so that such a recipe runs on CUDA tensors through this library, and dualip_fair_calc is the
// recipe of the demo itself (two dense fairness rows on top of the matching rows) as ONE kernel over the caller's CSC
// arrays -- one thread per column, no plan.
#include <math.h>
#include <algorithm>

#include "common.cuh"

using namespace dualip;

namespace dualip {

static int grid_for(int64_t n, int tb, int device) {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  return (int)std::max<int64_t>(1, std::min<int64_t>((n + tb - 1) / tb, (int64_t)sms * 16));
}

// out[e] = vals[e] * v[row[e]]                                   (sparse_utils.py:79: vals * v[row_idx])
template <typename IdxT>
__global__ void left_multiply_kernel(const float* __restrict__ vals, const IdxT* __restrict__ row, int64_t nnz,
                                     const float* __restrict__ v, float* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < nnz; i += stride) out[i] = __fmul_rn(vals[i], __ldg(v + row[i]));
}

// out[r] += sum of vals at row r (sparse_utils.py:240-242).  Every CTA sums its part in shared memory first (when the m
// floats fit), then adds its non-zero rows to the global vector: m-way contention stays on chip.
template <typename IdxT, bool SMEM>
__global__ void row_sums_kernel(const float* __restrict__ vals, const IdxT* __restrict__ row, int64_t nnz, int m,
                                float* __restrict__ out) {
  extern __shared__ float s_sum[];
  if (SMEM) {
    for (int i = threadIdx.x; i < m; i += blockDim.x) s_sum[i] = 0.f;
    __syncthreads();
  }
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < nnz; i += stride) {
    const float v = vals[i];
    if (v != 0.f) atomicAdd(SMEM ? &s_sum[row[i]] : &out[row[i]], v);
  }
  if (SMEM) {
    __syncthreads();
    for (int r = threadIdx.x; r < m; r += blockDim.x)
      if (s_sum[r] != 0.f) atomicAdd(&out[r], s_sum[r]);
  }
}

// Zero-padded [L x K] block of the columns cols[0..K) (sparse_utils.py:185-201), and its inverse (:207-210).
template <typename IdxT>
__global__ void gather_block_kernel(const IdxT* __restrict__ ccol, const float* __restrict__ vals,
                                    const long long* __restrict__ cols, long long K, long long L, float* __restrict__ block) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = K * L, stride = (long long)gridDim.x * blockDim.x;
  for (; t < total; t += stride) {
    const long long i = t / K, k = t - i * K;  // block[i][k]: consecutive threads walk along a block row
    const long long j = cols ? cols[k] : k;
    const long long e0 = (long long)ccol[j], len = (long long)ccol[j + 1] - e0;
    block[t] = i < len ? vals[e0 + i] : 0.f;
  }
}
template <typename IdxT>
__global__ void scatter_block_kernel(const IdxT* __restrict__ ccol, const float* __restrict__ block,
                                     const long long* __restrict__ cols, long long K, long long L, float* __restrict__ vals_out) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = K * L, stride = (long long)gridDim.x * blockDim.x;
  for (; t < total; t += stride) {
    const long long i = t / K, k = t - i * K;
    const long long j = cols ? cols[k] : k;
    const long long e0 = (long long)ccol[j], len = (long long)ccol[j + 1] - e0;
    if (i < len) vals_out[e0 + i] = block[t];
  }
}

// ------------------------------------------------------------------------------------------
// Fairness-row objective (docs/demo/matching_complex.rst:82-168), one thread per column.
//   scaled = fl(s * lambda), s = fl32(-1/gamma)                                         (rst:104)
//   v = fl(fl(fl(fl(a*scaled_r) + fl(scaled_m * f)) + fl(-scaled_{m+1} * f)) + fl(s*c))  (rst:107-116, in this order)
//   x = Proj_column(v)  -- box / cone / simplex / simplex_eq on the zero-padded block of the column's length bucket
//   sums[r] += fl(a*x) (r < m);  sums[m] = sum fl(f*x);  sums[m+1] = -sums[m];  c.x;  ||x||^2      (rst:126-136)
// The m+2-length tail (grad = sums - b, dual objective, slacks: calc_grad, matching.py:25-34) is dualip_matching_epilogue.
// ------------------------------------------------------------------------------------------
struct FairArgs {
  const void* ccol;
  const void* row;
  const float* a;
  const float* c;
  const float* f;
  const uint8_t* col_class;
  const dualip_proj_class* classes;
  const int* pad;  // n_classes x DUALIP_PAD_BUCKETS or null
  const float* lambda;  // m + 2
  long long n_cols;
  int m;
  float s;
  float* xbuf;      // nnz floats: v, then x
  float* sums;      // m + 4 floats, zeroed: [row sums (m) | f.x | -f.x | c.x | ||x||^2]
  double* dacc;     // 3 doubles, zeroed: f.x, c.x, ||x||^2
  unsigned int* ticket;
};

template <typename IdxT, bool SMEM>
__global__ void __launch_bounds__(256) fair_calc_kernel(const FairArgs k) {
  extern __shared__ float s_sum[];  // m floats when SMEM
  __shared__ double s_red[32];
  __shared__ unsigned int s_last;
  const IdxT* __restrict__ ccol = reinterpret_cast<const IdxT*>(k.ccol);
  const IdxT* __restrict__ row = reinterpret_cast<const IdxT*>(k.row);
  const int m = k.m;
  if (SMEM) {
    for (int i = threadIdx.x; i < m; i += blockDim.x) s_sum[i] = 0.f;
    __syncthreads();
  }
  const float s = k.s;
  const float sl_m = __fmul_rn(s, __ldg(k.lambda + m));           // scaled[-2]
  const float nsl_m1 = -__fmul_rn(s, __ldg(k.lambda + m + 1));    // -1 * scaled[-1]
  double fx = 0.0, cx = 0.0, xx = 0.0;
  long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; j < k.n_cols; j += stride) {
    const long long e0 = (long long)ccol[j];
    const int d = (int)((long long)ccol[j + 1] - e0);
    if (d <= 0) continue;
    const int cls = k.col_class ? (int)k.col_class[j] : 0;
    const dualip_proj_class pc = k.classes[cls];
    const float* __restrict__ a = k.a + e0;
    const float* __restrict__ c = k.c + e0;
    const float* __restrict__ f = k.f + e0;
    const IdxT* __restrict__ r = row + e0;
    float* __restrict__ xb = k.xbuf + e0;
    // pass 1: v (clamp classes: x at once); simplex: column sum in entry order and the two largest normalised values
    float S = 0.f, m1 = -1.f, m2 = -1.f;
    double Sd = 0.0;
    int am = 0;
    const bool clampk = pc.kind == DUALIP_PROJ_CLAMP;
    for (int q = 0; q < d; ++q) {
      const float sl = __fmul_rn(s, __ldg(k.lambda + r[q]));
      float v = __fmul_rn(a[q], sl);
      v = __fadd_rn(v, __fmul_rn(sl_m, f[q]));
      v = __fadd_rn(v, __fmul_rn(nsl_m1, f[q]));
      v = __fadd_rn(v, __fmul_rn(s, c[q]));
      if (clampk) {
        xb[q] = fminf(fmaxf(v, pc.lo), pc.hi);
      } else {
        const float u = fmaxf(v, 0.f);  // simplex.py:148
        xb[q] = u;
        S = __fadd_rn(S, u);
        Sd += (double)u;
        const float un = __fdiv_rn(u, pc.z);
        if (un > m1) {
          m2 = m1, m1 = un, am = q;
        } else if (un > m2) {
          m2 = un;
        }
      }
    }
    if (!clampk) {
      const int bkt = (d <= 1) ? 0 : 32 - __clz(d - 1);
      const int L = k.pad ? max(__ldg(k.pad + cls * DUALIP_PAD_BUCKETS + bkt), d) : d;  // padded length of the bucket
      const bool padded = (d > 1) || !(pc.flags & DUALIP_PROJ_FLAG_D1_UNPADDED);
      const float m2p = fmaxf(m2, 0.f);  // the zero padding takes part in the reference's top-2
      const float t_below = __fsub_rn((float)Sd, pc.z);
      if (pc.kind == DUALIP_PROJ_SIMPLEX && S <= pc.z_thr) {
        // feasible: x = u                                                                    simplex.py:153-155
      } else if (pc.kind == DUALIP_PROJ_SIMPLEX_EQ && t_below < 0.f) {
        const float theta = __fdiv_rn(t_below, (float)L);  // rho = L: every padded position satisfies cond_i (App. A #4)
        for (int q = 0; q < d; ++q) xb[q] = fmaxf(__fsub_rn(xb[q], theta), 0.f);
      } else if (padded && __fsub_rn(m1, m2p) > 1.0f) {
        for (int q = 0; q < d; ++q) xb[q] = (q == am) ? pc.z : 0.f;                                // simplex.py:166-190
      } else {
        // the sorted scan by ranks (ties in entry order, like a stable sort): css_i in fp64, cond_i, rho = max{i: cond_i}
        int rho = 1;
        float css_rho = 0.f;
        bool have = false;
        for (int q = 0; q < d; ++q) {
          const float u = xb[q];
          int rank = 0;
          double cs = 0.0;
          for (int t = 0; t < d; ++t) {
            const float ut = xb[t];
            if (ut > u || (ut == u && t <= q)) {
              ++rank;
              cs += (double)ut;
            }
          }
          const float css = (float)cs;
          const bool cond = __fsub_rn(u, __fdiv_rn(__fsub_rn(css, pc.z), (float)rank)) > 0.f;
          if (cond && (!have || rank > rho)) rho = rank, css_rho = css, have = true;
          if (!have && rank == 1) css_rho = css;  // no cond true: index 0 (simplex.py:225)
        }
        const float theta = __fdiv_rn(__fsub_rn(css_rho, pc.z), (float)rho);
        for (int q = 0; q < d; ++q) xb[q] = fmaxf(__fsub_rn(xb[q], theta), 0.f);
      }
    }
    // pass 2: sums
    for (int q = 0; q < d; ++q) {
      const float x = xb[q];
      const float g = __fmul_rn(a[q], x);
      if (g != 0.f) atomicAdd(SMEM ? &s_sum[r[q]] : &k.sums[r[q]], g);
      fx += (double)__fmul_rn(f[q], x);
      cx = fma((double)c[q], (double)x, cx);
      xx = fma((double)x, (double)x, xx);
    }
  }
  if (SMEM) {
    __syncthreads();
    for (int i = threadIdx.x; i < m; i += blockDim.x)
      if (s_sum[i] != 0.f) atomicAdd(&k.sums[i], s_sum[i]);
  }
  fx = block_sum(fx, s_red);
  cx = block_sum(cx, s_red);
  xx = block_sum(xx, s_red);
  if (threadIdx.x == 0) {
    if (fx != 0.0) atomicAdd(&k.dacc[0], fx);
    if (cx != 0.0) atomicAdd(&k.dacc[1], cx);
    if (xx != 0.0) atomicAdd(&k.dacc[2], xx);
    __threadfence();
    s_last = atomicAdd(k.ticket, 1u);
  }
  __syncthreads();
  if (s_last == gridDim.x - 1 && threadIdx.x == 0) {
    __threadfence();
    const float fsum = (float)__ldcg(&k.dacc[0]);
    k.sums[m] = fsum;        // grad[-2] = sum(A_fairness * x)      (rst:127)
    k.sums[m + 1] = -fsum;   // grad[-1] = sum(-A_fairness * x)     (rst:128)
    k.sums[m + 2] = (float)__ldcg(&k.dacc[1]);
    k.sums[m + 3] = (float)__ldcg(&k.dacc[2]);
    k.dacc[0] = 0.0, k.dacc[1] = 0.0, k.dacc[2] = 0.0;
    *k.ticket = 0u;
  }
}

}  // namespace dualip

extern "C" {

int dualip_csc_left_multiply(const float* vals_dev, const void* row_dev, int32_t index_bits, int64_t nnz, const float* v_dev,
                             float* out_dev, int32_t device, void* stream) {
  if ((nnz > 0 && (!vals_dev || !row_dev || !v_dev || !out_dev)) || nnz < 0 || (index_bits != 32 && index_bits != 64)) {
    set_error("bad argument");
    return DUALIP_EINVAL;
  }
  if (nnz == 0) return DUALIP_OK;
  DeviceGuard g(device);
  const int nb = grid_for(nnz, 256, device);
  if (index_bits == 64)
    left_multiply_kernel<long long><<<nb, 256, 0, (cudaStream_t)stream>>>(vals_dev, (const long long*)row_dev, nnz, v_dev, out_dev);
  else
    left_multiply_kernel<int><<<nb, 256, 0, (cudaStream_t)stream>>>(vals_dev, (const int*)row_dev, nnz, v_dev, out_dev);
  DUALIP_CUDA_TRY(cudaGetLastError());
  return DUALIP_OK;
}

int dualip_csc_row_sums(const float* vals_dev, const void* row_dev, int32_t index_bits, int64_t nnz, int32_t m,
                        float* out_dev, int32_t device, void* stream) {
  if (!out_dev || m <= 0 || nnz < 0 || (nnz > 0 && (!vals_dev || !row_dev)) || (index_bits != 32 && index_bits != 64)) {
    set_error("bad argument");
    return DUALIP_EINVAL;
  }
  DeviceGuard g(device);
  cudaStream_t st = (cudaStream_t)stream;
  DUALIP_CUDA_TRY(cudaMemsetAsync(out_dev, 0, sizeof(float) * m, st));
  if (nnz == 0) return DUALIP_OK;
  const int nb = std::min(grid_for(nnz, 256, device), 4 * 148);
  const size_t sm = sizeof(float) * (size_t)m;
  const bool smem = sm <= 48 * 1024;
  if (index_bits == 64) {
    if (smem)
      row_sums_kernel<long long, true><<<nb, 256, sm, st>>>(vals_dev, (const long long*)row_dev, nnz, m, out_dev);
    else
      row_sums_kernel<long long, false><<<nb, 256, 0, st>>>(vals_dev, (const long long*)row_dev, nnz, m, out_dev);
  } else {
    if (smem)
      row_sums_kernel<int, true><<<nb, 256, sm, st>>>(vals_dev, (const int*)row_dev, nnz, m, out_dev);
    else
      row_sums_kernel<int, false><<<nb, 256, 0, st>>>(vals_dev, (const int*)row_dev, nnz, m, out_dev);
  }
  DUALIP_CUDA_TRY(cudaGetLastError());
  return DUALIP_OK;
}

int dualip_csc_gather_block(const void* ccol_dev, int32_t index_bits, const float* vals_dev, const int64_t* cols_dev, int64_t K,
                            int64_t L, float* block_dev, int32_t device, void* stream) {
  if (!ccol_dev || !block_dev || K < 0 || L < 0 || (index_bits != 32 && index_bits != 64)) {
    set_error("bad argument");
    return DUALIP_EINVAL;
  }
  if (K == 0 || L == 0) return DUALIP_OK;
  DeviceGuard g(device);
  const int nb = grid_for(K * L, 256, device);
  if (index_bits == 64)
    gather_block_kernel<long long><<<nb, 256, 0, (cudaStream_t)stream>>>((const long long*)ccol_dev, vals_dev,
                                                                           (const long long*)cols_dev, K, L, block_dev);
  else
    gather_block_kernel<int><<<nb, 256, 0, (cudaStream_t)stream>>>((const int*)ccol_dev, vals_dev, (const long long*)cols_dev, K, L,
                                                                     block_dev);
  DUALIP_CUDA_TRY(cudaGetLastError());
  return DUALIP_OK;
}

int dualip_csc_scatter_block(const void* ccol_dev, int32_t index_bits, const float* block_dev, const int64_t* cols_dev, int64_t K,
                             int64_t L, float* vals_out_dev, int32_t device, void* stream) {
  if (!ccol_dev || !block_dev || !vals_out_dev || K < 0 || L < 0 || (index_bits != 32 && index_bits != 64)) {
    set_error("bad argument");
    return DUALIP_EINVAL;
  }
  if (K == 0 || L == 0) return DUALIP_OK;
  DeviceGuard g(device);
  const int nb = grid_for(K * L, 256, device);
  if (index_bits == 64)
    scatter_block_kernel<long long><<<nb, 256, 0, (cudaStream_t)stream>>>((const long long*)ccol_dev, block_dev,
                                                                            (const long long*)cols_dev, K, L, vals_out_dev);
  else
    scatter_block_kernel<int><<<nb, 256, 0, (cudaStream_t)stream>>>((const int*)ccol_dev, block_dev, (const long long*)cols_dev, K, L,
                                                                      vals_out_dev);
  DUALIP_CUDA_TRY(cudaGetLastError());
  return DUALIP_OK;
}

int dualip_fair_calc(const dualip_csc_desc* d, const float* f_dev, const float* lambda_dev, const float* b_dev, double gamma,
                     float* grad_out_dev, dualip_scalars* scalars_out_dev, float* x_out_dev, float* work_dev, void* stream) {
  if (!d || !lambda_dev || !grad_out_dev || !scalars_out_dev || !x_out_dev || !work_dev || !d->classes || !d->ccol_dev ||
      (d->nnz > 0 && (!d->row_dev || !d->a_dev || !d->c_dev || !f_dev))) {
    set_error("null argument");
    return DUALIP_EINVAL;
  }
  if (d->n_rows <= 0 || d->n_cols < 0 || (d->index_bits != 32 && d->index_bits != 64) || d->n_classes < 1) {
    set_error("bad shape");
    return DUALIP_EINVAL;
  }
  if (!(gamma > 0.0) && !(gamma < 0.0)) {
    set_error("gamma must be non-zero");
    return DUALIP_EINVAL;
  }
  for (int i = 0; i < d->n_classes; ++i)
    if (d->classes[i].kind < DUALIP_PROJ_CLAMP || d->classes[i].kind > DUALIP_PROJ_SIMPLEX_EQ) {
      set_error("class %d: the fairness objective projects with box / cone / simplex / simplex_eq (Duchi) only", i);
      return DUALIP_EINVAL;
    }
  DeviceGuard g(d->device);
  cudaStream_t st = (cudaStream_t)stream;
  const int m = d->n_rows;
  // work_dev: [sums: m + 4 floats | pad to 8 bytes | 3 doubles | ticket | classes | pad table]; zeroed by the caller ONCE
  // (the kernel leaves the doubles and the ticket zeroed; the sums are cleared here)
  float* sums = work_dev;
  const size_t off_d = (((size_t)(m + 4) * sizeof(float)) + 15) & ~(size_t)15;
  double* dacc = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(work_dev) + off_d);
  unsigned int* ticket = reinterpret_cast<unsigned int*>(dacc + 3);
  dualip_proj_class* cls_dev = reinterpret_cast<dualip_proj_class*>(reinterpret_cast<unsigned char*>(work_dev) + off_d + 32);
  int* pad_dev = reinterpret_cast<int*>(cls_dev + d->n_classes);
  DUALIP_CUDA_TRY(cudaMemsetAsync(sums, 0, sizeof(float) * (m + 4), st));
  DUALIP_CUDA_TRY(cudaMemcpyAsync(cls_dev, d->classes, sizeof(dualip_proj_class) * d->n_classes, cudaMemcpyHostToDevice, st));
  if (d->pad_len)
    DUALIP_CUDA_TRY(cudaMemcpyAsync(pad_dev, d->pad_len, sizeof(int) * (size_t)d->n_classes * DUALIP_PAD_BUCKETS,
                                    cudaMemcpyHostToDevice, st));
  FairArgs k;
  k.ccol = d->ccol_dev, k.row = d->row_dev, k.a = d->a_dev, k.c = d->c_dev, k.f = f_dev;
  k.col_class = d->col_class_dev, k.classes = cls_dev, k.pad = d->pad_len ? pad_dev : nullptr;
  k.lambda = lambda_dev, k.n_cols = d->n_cols, k.m = m, k.s = (float)(-1.0 / gamma);
  k.xbuf = x_out_dev, k.sums = sums, k.dacc = dacc, k.ticket = ticket;
  const size_t sm = sizeof(float) * (size_t)m;
  const bool smem = sm <= 40 * 1024;
  const int nb = std::max(1, std::min(grid_for(d->n_cols, 256, d->device), 148 * 4));
  if (d->index_bits == 64) {
    if (smem)
      fair_calc_kernel<long long, true><<<nb, 256, sm, st>>>(k);
    else
      fair_calc_kernel<long long, false><<<nb, 256, 0, st>>>(k);
  } else {
    if (smem)
      fair_calc_kernel<int, true><<<nb, 256, sm, st>>>(k);
    else
      fair_calc_kernel<int, false><<<nb, 256, 0, st>>>(k);
  }
  DUALIP_CUDA_TRY(cudaGetLastError());
  // m+2-length tail: grad = sums - b, lambda.grad, slacks, dual objective (calc_grad; rst:139-156)
  return dualip_matching_epilogue(sums, m + 2, lambda_dev, b_dev, gamma, grad_out_dev, scalars_out_dev, stream);
}

int64_t dualip_fair_work_bytes(int32_t n_rows, int32_t n_classes) {
  const size_t off_d = (((size_t)(n_rows + 4) * sizeof(float)) + 15) & ~(size_t)15;
  return (int64_t)(off_d + 32 + sizeof(dualip_proj_class) * (size_t)n_classes + sizeof(int) * (size_t)n_classes * DUALIP_PAD_BUCKETS);
}

}  // extern "C"

// Setup-time kernels: Jacobi row preconditioning (reference src/dualip/preprocessing/precondition.py:8-29,
// utils/sparse_utils.py:429-450).  One pass to accumulate row norms, one pass to scale.
#include <math.h>
#include <algorithm>

#include "common.cuh"

using namespace dualip;

namespace dualip {

template <typename IdxT>
__global__ void row_sq_norm_kernel(const float* __restrict__ a, const IdxT* __restrict__ row, int64_t nnz,
                                   double* __restrict__ sq) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < nnz; i += stride) {
    const float v = a[i];
    atomicAdd(&sq[row[i]], (double)__fmul_rn(v, v));  // vals.pow(2) in fp32, then summed (sparse_utils.py:443-449)
  }
}

__global__ void finish_norms_kernel(const double* __restrict__ sq, float* __restrict__ norms, float* __restrict__ rec,
                                    float* __restrict__ b, int m) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const float nr = sqrtf((float)sq[i]);   // row_sums.pow(1/2) in fp32
  norms[i] = nr;
  const float r = __fdiv_rn(1.0f, nr);    // reciprocal = 1 / row_norms        (precondition.py:25)
  rec[i] = r;
  if (b) b[i] = __fmul_rn(b[i], r);       // b.mul_(reciprocal)                (precondition.py:29)
}

template <typename IdxT>
__global__ void scale_rows_kernel(float* __restrict__ a, const IdxT* __restrict__ row, int64_t nnz,
                                  const float* __restrict__ rec) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < nnz; i += stride) a[i] = __fmul_rn(a[i], __ldg(rec + row[i]));  // vals * v[row_idx] (sparse_utils.py:79)
}

}  // namespace dualip

extern "C" int dualip_jacobi_precondition(float* a_dev, const void* row_dev, int32_t index_bits, int64_t nnz,
                                          float* b_dev, int32_t m, float* norms_out_dev, int32_t device, void* stream) {
  if (!a_dev || !row_dev || !norms_out_dev || m <= 0 || nnz < 0 || (index_bits != 32 && index_bits != 64)) {
    set_error("bad argument");
    return DUALIP_EINVAL;
  }
  DeviceGuard g(device);
  cudaStream_t st = (cudaStream_t)stream;
  double* sq = nullptr;
  float* rec = nullptr;
  DUALIP_CUDA_TRY(cudaMalloc(&sq, sizeof(double) * m));
  DUALIP_CUDA_TRY(cudaMalloc(&rec, sizeof(float) * m));
  DUALIP_CUDA_TRY(cudaMemsetAsync(sq, 0, sizeof(double) * m, st));
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  const int tb = 256;
  const int nb = (int)std::max<int64_t>(1, std::min<int64_t>((nnz + tb - 1) / tb, (int64_t)sms * 16));
  if (nnz > 0) {
    if (index_bits == 64)
      row_sq_norm_kernel<long long><<<nb, tb, 0, st>>>(a_dev, (const long long*)row_dev, nnz, sq);
    else
      row_sq_norm_kernel<int><<<nb, tb, 0, st>>>(a_dev, (const int*)row_dev, nnz, sq);
  }
  finish_norms_kernel<<<(m + 255) / 256, 256, 0, st>>>(sq, norms_out_dev, rec, b_dev, m);
  if (nnz > 0) {
    if (index_bits == 64)
      scale_rows_kernel<long long><<<nb, tb, 0, st>>>(a_dev, (const long long*)row_dev, nnz, rec);
    else
      scale_rows_kernel<int><<<nb, tb, 0, st>>>(a_dev, (const int*)row_dev, nnz, rec);
  }
  DUALIP_CUDA_TRY(cudaGetLastError());
  DUALIP_CUDA_TRY(cudaStreamSynchronize(st));
  cudaFree(sq);
  cudaFree(rec);
  return DUALIP_OK;
}

// Sharded variant of the above: (1) local squared row norms, (2) caller all-reduces them, (3) scale with 1/norm.
extern "C" int dualip_row_sq_norms(const float* a_dev, const void* row_dev, int32_t index_bits, int64_t nnz, int32_t m,
                                   double* sq_out_dev, int32_t device, void* stream) {
  if (!a_dev || !row_dev || !sq_out_dev || m <= 0 || nnz < 0 || (index_bits != 32 && index_bits != 64)) {
    set_error("bad argument");
    return DUALIP_EINVAL;
  }
  DeviceGuard g(device);
  cudaStream_t st = (cudaStream_t)stream;
  DUALIP_CUDA_TRY(cudaMemsetAsync(sq_out_dev, 0, sizeof(double) * m, st));
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  const int tb = 256;
  const int nb = (int)std::max<int64_t>(1, std::min<int64_t>((nnz + tb - 1) / tb, (int64_t)sms * 16));
  if (nnz > 0) {
    if (index_bits == 64)
      row_sq_norm_kernel<long long><<<nb, tb, 0, st>>>(a_dev, (const long long*)row_dev, nnz, sq_out_dev);
    else
      row_sq_norm_kernel<int><<<nb, tb, 0, st>>>(a_dev, (const int*)row_dev, nnz, sq_out_dev);
  }
  DUALIP_CUDA_TRY(cudaGetLastError());
  return DUALIP_OK;
}

extern "C" int dualip_scale_rows(float* a_dev, const void* row_dev, int32_t index_bits, int64_t nnz,
                                 const float* scale_dev, int32_t device, void* stream) {
  if (!a_dev || !row_dev || !scale_dev || nnz < 0 || (index_bits != 32 && index_bits != 64)) {
    set_error("bad argument");
    return DUALIP_EINVAL;
  }
  DeviceGuard g(device);
  cudaStream_t st = (cudaStream_t)stream;
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  const int tb = 256;
  const int nb = (int)std::max<int64_t>(1, std::min<int64_t>((nnz + tb - 1) / tb, (int64_t)sms * 16));
  if (nnz > 0) {
    if (index_bits == 64)
      scale_rows_kernel<long long><<<nb, tb, 0, st>>>(a_dev, (const long long*)row_dev, nnz, scale_dev);
    else
      scale_rows_kernel<int><<<nb, tb, 0, st>>>(a_dev, (const int*)row_dev, nnz, scale_dev);
  }
  DUALIP_CUDA_TRY(cudaGetLastError());
  return DUALIP_OK;
}

// ------------------------------------------------------------------------------------------
// ProjectionOperator.__call__ on a zero-padded dense [L x K] block (reference projections/base.py:30-36 contract,
// box.py:15-16, cone.py:21-28, simplex.py:126-236).  One thread per column; API surface, not the hot path
// (the matching objective projects inside matching_pass_kernel).
// ------------------------------------------------------------------------------------------
namespace dualip {

__global__ void project_block_kernel(const float* __restrict__ x, float* __restrict__ out, long long L, long long K,
                                     dualip_proj_class pc) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= K) return;
  if (pc.kind == DUALIP_PROJ_CLAMP) {
    for (long long i = 0; i < L; ++i) out[i * K + j] = fminf(fmaxf(x[i * K + j], pc.lo), pc.hi);
    return;
  }
  if (pc.kind >= DUALIP_PROJ_SIMPLEX_BISECT) {
    // method="bisection_search" (simplex.py:6-123): no pre-clamp; sums run down the rows in fp32 like torch's outer-dimension
    // reduction; every column of a block halves the same interval [-1, 0] (19 halvings until 2^-20 < tol)
    float ssum = 0.f, vmin = INFINITY, t0 = -INFINITY, t1 = -INFINITY;
    long long am = 0;
    for (long long i = 0; i < L; ++i) {
      const float v = x[i * K + j];
      ssum = __fadd_rn(ssum, v);
      vmin = fminf(vmin, v);
      const float xn = __fdiv_rn(v, pc.z);
      if (xn > t0) {
        t1 = t0, t0 = xn, am = i;
      } else if (xn > t1) {
        t1 = xn;
      }
    }
    if (pc.kind == DUALIP_PROJ_SIMPLEX_BISECT && ssum <= pc.z_thr && vmin >= -1e-6f) {  // :40-41
      for (long long i = 0; i < L; ++i) out[i * K + j] = x[i * K + j];
      return;
    }
    if (L > 1 && __fsub_rn(t0, t1) > 1.0f) {  // :52-75
      for (long long i = 0; i < L; ++i) out[i * K + j] = (i == am) ? pc.z : 0.f;
      return;
    }
    float lo = -1.f, hi = 0.f, prev = 0.f;
    bool act = true;
    for (int it = 0; it < 50 && act; ++it) {  // :95-118
      const float mid = __fmul_rn(__fadd_rn(lo, hi), 0.5f);
      if (it > 0 && fabsf(__fsub_rn(mid, prev)) < 1e-6f) break;
      float sm = 0.f;
      for (long long i = 0; i < L; ++i) sm = __fadd_rn(sm, fmaxf(__fsub_rn(__fsub_rn(x[i * K + j], t0), mid), 0.f));
      const bool high = sm > 1.0f;
      lo = high ? mid : lo;
      hi = high ? hi : mid;
      act = !(__fsub_rn(hi, lo) < 1e-6f);
      prev = mid;
    }
    const float nu = __fmul_rn(__fadd_rn(lo, hi), 0.5f);
    for (long long i = 0; i < L; ++i)
      out[i * K + j] = __fmul_rn(fmaxf(__fsub_rn(__fsub_rn(x[i * K + j], t0), nu), 0.f), pc.z);  // :120-122
    return;
  }
  float S = 0.f, m1 = -1.f, m2 = -1.f;
  long long am = 0;
  for (long long i = 0; i < L; ++i) {
    const float u = fmaxf(x[i * K + j], 0.f);
    S = __fadd_rn(S, u);
    const float un = __fdiv_rn(u, pc.z);
    if (un > m1) {
      m2 = m1;
      m1 = un;
      am = i;
    } else if (un > m2) {
      m2 = un;
    }
  }
  if (pc.kind == DUALIP_PROJ_SIMPLEX && S <= pc.z_thr) {
    for (long long i = 0; i < L; ++i) out[i * K + j] = fmaxf(x[i * K + j], 0.f);
    return;
  }
  if (L > 1 && __fsub_rn(m1, m2) > 1.0f) {
    for (long long i = 0; i < L; ++i) out[i * K + j] = (i == am) ? pc.z : 0.f;
    return;
  }
  long long rho = 1;
  float css_rho = 0.f;
  if (L <= 2048) {
    // exact restatement of the sorted scan: rank by all-pairs comparison, prefix sums in fp64
    bool have = false;
    for (long long i = 0; i < L; ++i) {
      const float u = fmaxf(x[i * K + j], 0.f);
      long long rank = 0;
      double cs = 0.0;
      for (long long t = 0; t < L; ++t) {
        const float ut = fmaxf(x[t * K + j], 0.f);
        if (ut > u || (ut == u && t <= i)) {
          ++rank;
          cs += (double)ut;
        }
      }
      const float css = (float)cs;
      const bool cond = __fsub_rn(u, __fdiv_rn(__fsub_rn(css, pc.z), (float)rank)) > 0.f;
      if (cond && (!have || rank > rho)) {
        rho = rank;
        css_rho = css;
        have = true;
      }
      if (!have && rank == 1) css_rho = css;  // rho0 = 0 fallback of the reference (simplex.py:225)
    }
  } else {
    // Michelot fixed point for very long columns
    double t = ((double)S - (double)pc.z) / (double)L, ssum = S;
    long long cnt_prev = L;
    for (int it = 0; it < 128; ++it) {
      double s2 = 0.0;
      long long n2 = 0;
      for (long long i = 0; i < L; ++i) {
        const float u = fmaxf(x[i * K + j], 0.f);
        if ((double)u > t) {
          s2 += (double)u;
          ++n2;
        }
      }
      if (n2 == 0) break;
      ssum = s2;
      const bool done = (n2 == cnt_prev);
      cnt_prev = n2;
      t = (s2 - (double)pc.z) / (double)n2;
      if (done) break;
    }
    rho = cnt_prev;
    css_rho = (float)ssum;
  }
  const float theta = __fdiv_rn(__fsub_rn(css_rho, pc.z), (float)rho);
  for (long long i = 0; i < L; ++i) out[i * K + j] = fmaxf(__fsub_rn(fmaxf(x[i * K + j], 0.f), theta), 0.f);
}

}  // namespace dualip

extern "C" int dualip_project_block(const float* x_dev, float* out_dev, int64_t L, int64_t K,
                                    const dualip_proj_class* cls, void* stream) {
  if (!x_dev || !out_dev || !cls || L < 0 || K < 0) {
    set_error("bad argument");
    return DUALIP_EINVAL;
  }
  if (cls->kind != DUALIP_PROJ_CLAMP && !(cls->z > 0.f)) {
    set_error("Simplex radius z must be positive.");
    return DUALIP_EINVAL;
  }
  if (L == 0 || K == 0) return DUALIP_OK;
  const int tb = 128;
  project_block_kernel<<<(unsigned)((K + tb - 1) / tb), tb, 0, (cudaStream_t)stream>>>(x_dev, out_dev, L, K, *cls);
  DUALIP_CUDA_TRY(cudaGetLastError());
  return DUALIP_OK;
}

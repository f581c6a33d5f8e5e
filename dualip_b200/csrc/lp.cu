// Generic-LP dual evaluation (no block structure): reference src/dualip/objectives/miplib.py:60-109
//   z = -1/gamma * (A^T lambda + c) ; x = clamp(z, lower, upper) per variable (box / cone entries of the projection map)
//   grad = A x - b ; reg = gamma/2 ||x||^2 ; dual_obj = c.x + reg + lambda.(A x - b)
// Three launches on the caller's stream, no host synchronisation: a warp per variable over the CSC copy of A, a warp
// per constraint row over the CSR copy, and a one-thread finalisation that also re-arms the scratch accumulators.
// The shipped MIPLIB instance is tiny (150 variables x 7822 rows), so this path is latency-bound; it exists so that the
// second ObjectiveFunction of the reference runs on the device behind the same Maximizer.
#include <math.h>
#include <algorithm>

#include "common.cuh"

using namespace dualip;

namespace dualip {

// scratch (doubles): [0] c.x  [1] ||x||^2  [2] lambda'.(Ax-b)  [3] sum relu(grad)  [4] ||grad||^2  [5] max relu(grad) as
// float bits in the low word (non-negative floats order like unsigned integers)
__global__ void __launch_bounds__(256) lp_primal_kernel(const dualip_lp_desc d, const float* __restrict__ lambda, float s,
                                                        float* __restrict__ x_out, double* __restrict__ scratch) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  double cx = 0.0, xx = 0.0;
  for (int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; j < d.n; j += warps) {
    float dot = 0.f;
    for (int e = d.csc_colptr_dev[j] + lane; e < d.csc_colptr_dev[j + 1]; e += 32) {
      const int r = d.csc_row_dev[e];
      const float lam = d.row_scale_dev ? __fmul_rn(d.row_scale_dev[r], lambda[r]) : lambda[r];  // miplib.py:73-74
      dot = fmaf(d.csc_val_dev[e], lam, dot);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    const float z = __fmul_rn(s, __fadd_rn(dot, d.c_dev[j]));                                    // miplib.py:76
    const float x = fminf(fmaxf(z, d.lo_dev[j]), d.hi_dev[j]);                                   // miplib.py:80-90
    if (lane == 0) {
      x_out[j] = x;
      cx = fma((double)d.c_dev[j], (double)x, cx);
      xx = fma((double)x, (double)x, xx);
    }
  }
  if (lane == 0) {
    if (cx != 0.0) atomicAdd(&scratch[0], cx);
    if (xx != 0.0) atomicAdd(&scratch[1], xx);
  }
}

__global__ void __launch_bounds__(256) lp_dual_kernel(const dualip_lp_desc d, const float* __restrict__ lambda,
                                                      const float* __restrict__ x, float* __restrict__ grad_out,
                                                      double* __restrict__ scratch) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  double lg = 0.0, sp = 0.0, g2 = 0.0;
  float mx = 0.f;
  for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < d.m; r += warps) {
    float ax = 0.f;
    for (int e = d.csr_rowptr_dev[r] + lane; e < d.csr_rowptr_dev[r + 1]; e += 32)
      ax = fmaf(d.csr_val_dev[e], x[d.csr_col_dev[e]], ax);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ax += __shfl_xor_sync(0xffffffffu, ax, o);
    if (lane == 0) {
      const float resid = __fsub_rn(ax, d.b_dev[r]);                                             // A x - b
      const float sc = d.row_scale_dev ? d.row_scale_dev[r] : 1.0f;
      const float g = d.row_scale_dev ? __fmul_rn(sc, resid) : resid;                            // miplib.py:92-95
      grad_out[r] = g;
      lg = fma((double)(d.row_scale_dev ? __fmul_rn(sc, lambda[r]) : lambda[r]), (double)resid, lg);  // miplib.py:99
      sp += (double)fmaxf(g, 0.f);
      g2 = fma((double)g, (double)g, g2);
      mx = fmaxf(mx, g);
    }
  }
  if (lane == 0) {
    if (lg != 0.0) atomicAdd(&scratch[2], lg);
    if (sp != 0.0) atomicAdd(&scratch[3], sp);
    if (g2 != 0.0) atomicAdd(&scratch[4], g2);
    if (mx > 0.f) atomicMax(reinterpret_cast<unsigned int*>(&scratch[5]), __float_as_uint(mx));
  }
}

__global__ void lp_finalize_kernel(double* __restrict__ scratch, double gamma, dualip_scalars* __restrict__ out) {
  const double cx = scratch[0], xx = scratch[1], lg = scratch[2];
  dualip_scalars r;
  r.primal_objective = cx;
  r.reg_penalty = 0.5 * gamma * xx;
  r.dual_val_times_grad = lg;
  r.dual_objective = cx + r.reg_penalty + lg;
  r.sum_pos_slack = scratch[3];
  r.grad_sq_norm = scratch[4];
  r.max_pos_slack = (double)__uint_as_float(*reinterpret_cast<unsigned int*>(&scratch[5]));
  r.x_sq_norm = xx;
  *out = r;
  for (int i = 0; i < 6; ++i) scratch[i] = 0.0;
}

}  // namespace dualip

extern "C" int dualip_lp_calc(const dualip_lp_desc* d, const float* lambda_dev, double gamma, float* x_out_dev,
                              float* grad_out_dev, dualip_scalars* scalars_out_dev, double* scratch_dev, void* stream) {
  if (!d || !lambda_dev || !x_out_dev || !grad_out_dev || !scalars_out_dev || !scratch_dev) {
    set_error("null argument");
    return DUALIP_EINVAL;
  }
  if (d->m <= 0 || d->n <= 0 || !d->csr_rowptr_dev || !d->csc_colptr_dev || !d->c_dev || !d->b_dev || !d->lo_dev || !d->hi_dev ||
      (d->nnz > 0 && (!d->csr_col_dev || !d->csr_val_dev || !d->csc_row_dev || !d->csc_val_dev))) {
    set_error("bad generic-LP description");
    return DUALIP_EINVAL;
  }
  if (!(gamma > 0.0) && !(gamma < 0.0)) {
    set_error("gamma must be non-zero");
    return DUALIP_EINVAL;
  }
  DeviceGuard g(d->device);
  if (!g.ok) {
    set_error("cannot select CUDA device %d", d->device);
    return DUALIP_ECUDA;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const float s = (float)(-1.0 / gamma);
  const int bn = (int)std::min<int64_t>(((int64_t)d->n * 32 + 255) / 256, 148 * 8);
  const int bm = (int)std::min<int64_t>(((int64_t)d->m * 32 + 255) / 256, 148 * 8);
  lp_primal_kernel<<<bn, 256, 0, st>>>(*d, lambda_dev, s, x_out_dev, scratch_dev);
  lp_dual_kernel<<<bm, 256, 0, st>>>(*d, lambda_dev, x_out_dev, grad_out_dev, scratch_dev);
  lp_finalize_kernel<<<1, 1, 0, st>>>(scratch_dev, gamma, scalars_out_dev);
  DUALIP_CUDA_TRY(cudaGetLastError());
  return DUALIP_OK;
}

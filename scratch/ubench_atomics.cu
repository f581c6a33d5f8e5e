// Microbenchmark: throughput of scatter-add flavours into an m-bin vector (m = 10k),
// to pick the dual-gradient accumulation strategy. Not product code.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ unsigned hash32(unsigned x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}

template <int MODE>
__global__ void __launch_bounds__(512) bench(float* gout, double* gout64, int m, int iters, int active_lanes, long long* cyc) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* sf = reinterpret_cast<float*>(smem_raw);
  unsigned long long* s64 = reinterpret_cast<unsigned long long*>(smem_raw);
  unsigned* s32 = reinterpret_cast<unsigned*>(smem_raw);
  int words = (MODE == 2) ? 2 * m : m;
  for (int i = threadIdx.x; i < words; i += blockDim.x) s32[i] = 0;
  __syncthreads();
  long long t0 = clock64();
  unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
  int lane = threadIdx.x & 31;
  float acc = 0.f;
  if (lane < active_lanes) {
    #pragma unroll 4
    for (int it = 0; it < iters; ++it) {
      unsigned h = hash32(tid * 7919u + it * 104729u);
      int idx = h % (unsigned)m;
      float v = __uint_as_float(0x3f800000u | (h >> 9)) - 1.0f;  // [0,1)
      if (MODE == 0) atomicAdd(&sf[idx], v);                                  // smem f32 (CAS loop)
      else if (MODE == 1) atomicAdd(&s32[idx], (unsigned)(h & 0xff));            // smem u32 native
      else if (MODE == 2) atomicAdd(&s64[idx], (unsigned long long)(__float2ll_rn(v * 1099511627776.0f))); // smem u64 fixed point
      else if (MODE == 3) atomicAdd(&gout[idx], v);                             // global f32 RED
      else if (MODE == 4) atomicAdd(&gout64[idx], (double)v);                   // global f64 RED
      else if (MODE == 5) acc += sf[idx] * v;                                   // smem random gather only
      else if (MODE == 6) { if (it & 1) atomicAdd(&sf[idx], v); else atomicAdd(&gout[idx], v); } // mix
      else if (MODE == 7) { if ((h >> 3) % 10 == 0) atomicAdd(&sf[idx], v); }   // 10% density smem f32
    }
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  if (MODE == 5 && acc == 12345.f) gout[0] = acc;
  if (MODE == 0 || MODE == 6 || MODE == 7) for (int i = threadIdx.x; i < m; i += blockDim.x) atomicAdd(&gout[i], sf[i]);
}

template <int MODE>
void run(const char* name, int m, int iters, int active, int ctas_per_sm) {
  int dev = 0, sms = 0; CK(cudaGetDevice(&dev)); CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  int grid = sms * ctas_per_sm;
  float* g; double* g64; long long* cyc;
  CK(cudaMalloc(&g, m * 4)); CK(cudaMalloc(&g64, m * 8)); CK(cudaMalloc(&cyc, grid * 8));
  CK(cudaMemset(g, 0, m * 4)); CK(cudaMemset(g64, 0, m * 8));
  size_t smem = (MODE == 2 ? 8 : 4) * (size_t)m;
  CK(cudaFuncSetAttribute(bench<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int w = 0; w < 2; ++w) bench<MODE><<<grid, 512, smem>>>(g, g64, m, iters, active, cyc);
  CK(cudaEventRecord(e0));
  bench<MODE><<<grid, 512, smem>>>(g, g64, m, iters, active, cyc);
  CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  long long* h = (long long*)malloc(grid * 8); CK(cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost));
  long long mx = 0; double avg = 0; for (int i = 0; i < grid; ++i) { if (h[i] > mx) mx = h[i]; avg += h[i]; } avg /= grid;
  double ops = (double)grid * 16 * active * iters;  // 16 warps per CTA
  printf("%-28s m=%d active=%2d cta/sm=%d : %.3f ms  %.1f Gops/s  ops/cycle/SM(avg cyc)=%.3f  (max cyc %lld avg %.0f)\n",
         name, m, active, ctas_per_sm, ms, ops / ms * 1e-6, (double)ctas_per_sm * 16 * active * iters / avg, mx, avg);
  cudaFree(g); cudaFree(g64); cudaFree(cyc); free(h);
}

int main() {
  int iters = 4096;
  for (int m : {10000, 1000}) {
    for (int cps : {1, 2}) {
      run<0>("smem f32 atomicAdd (CAS)", m, iters, 32, cps);
      run<0>("smem f32 atomicAdd (CAS)", m, iters, 20, cps);
      run<1>("smem u32 atomicAdd", m, iters, 32, cps);
      run<2>("smem u64 fixed-point add", m, iters, 32, cps);
      run<3>("global f32 RED", m, iters, 32, cps);
      run<4>("global f64 RED", m, iters, 32, cps);
      run<5>("smem random gather", m, iters, 32, cps);
      run<6>("mix smem/global f32", m, iters, 32, cps);
      run<7>("smem f32 10% density", m, iters, 32, cps);
    }
  }
  return 0;
}

// Microbenchmark v2: MIO cost of the two shared-memory operations of the hot kernel with random row addresses:
// 32-bit gather (LDS), native integer add (ATOMS.ADD), fp32 add (ATOMS.CAST.SPIN loop).  Addresses are kept in
// registers and advanced with one add+and per use, so the loop is MIO-bound, not ALU-bound.  Not product code.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
__device__ __forceinline__ unsigned hash32(unsigned x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

template <int MODE>
__global__ void __launch_bounds__(512) bench(float* gout, int iters, int active_pct, long long* cyc) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned* s32 = reinterpret_cast<unsigned*>(smem_raw);
  float* sf = reinterpret_cast<float*>(smem_raw);
  const int M = 8192;
  for (int i = threadIdx.x; i < M; i += blockDim.x) s32[i] = 0;
  __syncthreads();
  unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned idx[8];
  for (int j = 0; j < 8; ++j) idx[j] = (hash32(tid * 8u + j) & (M - 1)) * 4u;
  const bool act = (hash32(tid ^ 0x9e3779b9u) % 100u) < (unsigned)active_pct;
  const unsigned step = (hash32(tid + 77u) | 1u) * 4u;
  float acc = 0.f;
  const unsigned base = (unsigned)__cvta_generic_to_shared(smem_raw);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const unsigned a = base + idx[j];
      if (MODE == 0) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); acc += v; }
      else if (MODE == 1) { asm volatile("red.shared.add.s32 [%0], %1;" ::"r"(a), "r"(it + 1) : "memory"); }
      else if (MODE == 2) { if (act) asm volatile("red.shared.add.s32 [%0], %1;" ::"r"(a), "r"(it + 1) : "memory"); }
      else if (MODE == 3) { if (act) atomicAdd(reinterpret_cast<float*>(smem_raw + idx[j]), 1.0f); }
      else if (MODE == 4) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); acc += v;
                            if (act) asm volatile("red.shared.add.s32 [%0], %1;" ::"r"(a ^ 64u), "r"(it + 1) : "memory"); }
      idx[j] = (idx[j] + step) & (M * 4 - 4);
    }
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  if (acc == 12345.f) gout[0] = acc;
  if (threadIdx.x == 1) gout[1 + (blockIdx.x & 3)] = sf[threadIdx.x];
}

template <int MODE>
void run(const char* name, int iters, int active_pct) {
  int grid = 148; float* g; long long* cyc;
  CK(cudaMalloc(&g, 64)); CK(cudaMalloc(&cyc, grid * 8));
  size_t smem = 8192 * 4;
  bench<MODE><<<grid, 512, smem>>>(g, iters, active_pct, cyc);
  bench<MODE><<<grid, 512, smem>>>(g, iters, active_pct, cyc);
  CK(cudaDeviceSynchronize());
  long long h[148]; CK(cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost));
  double avg = 0; for (int i = 0; i < grid; ++i) avg += h[i]; avg /= grid;
  double winst = 16.0 * iters * 8;  // warp-level operations per SM
  printf("%-40s active %3d%% : %.2f cycles per warp-op per SM\n", name, active_pct, avg / winst);
  cudaFree(g); cudaFree(cyc);
}

int main() {
  int iters = 2048;
  run<0>("LDS random gather", iters, 100);
  run<1>("ATOMS.ADD.S32 random (unconditional)", iters, 100);
  for (int p : {100, 60, 30, 10}) run<2>("ATOMS.ADD.S32 random (predicated lanes)", iters, p);
  for (int p : {100, 60, 30, 10}) run<3>("f32 atomicAdd (CAS loop) random", iters, p);
  for (int p : {100, 30}) run<4>("LDS gather + ATOMS.ADD", iters, p);
  return 0;
}

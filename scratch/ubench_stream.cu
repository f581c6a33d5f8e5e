// Read-bandwidth ceilings on B200 for the slab access pattern (not product code).
//   mode 0: grid-stride LDG.128 over one 10 GB array (read-only streaming peak)
//   mode 1: the slab kernel's pattern: 148 x 16 warps, warp w reads slabs w, w+W, ...: three chunks of 128d, 128d, 64d bytes
//           from three arrays, LDG.128/LDG.64 with L1::no_allocate, all loads of a slab issued before the first use
//   mode 2: as 1 but the slab staged by cp.async.bulk into a per-warp buffer (one slab in flight), then LDS.128
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_stream ubench_stream.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ float4 ldg_stream_f4(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ uint2 ldg_stream_u2(const void* p) {
  uint2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
  return v;
}
__global__ void k_linear(const float4* __restrict__ a, size_t n4, float* out) {
  float acc = 0.f;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n4; i += 4 * stride) {
    float4 v0 = ldg_stream_f4((const float*)(a + i)), v1 = ldg_stream_f4((const float*)(a + i + stride));
    float4 v2 = ldg_stream_f4((const float*)(a + i + 2 * stride)), v3 = ldg_stream_f4((const float*)(a + i + 3 * stride));
    acc += v0.x + v1.y + v2.z + v3.w;
  }
  if (acc == 123.456f) out[0] = acc;
}
template <int D>
__global__ void __launch_bounds__(512, 1) k_slab(const float* __restrict__ a, const float* __restrict__ c, const unsigned short* __restrict__ r,
                                                 long long n_slabs, float* out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long W = (long long)gridDim.x * 16;
  float acc = 0.f;
  for (long long sl = (long long)blockIdx.x * 16 + warp; sl < n_slabs; sl += W) {
    const size_t base = (size_t)sl * 32 * D;
    float4 va[D / 4 + 1], vc[D / 4 + 1];
    uint2 vr[D / 4 + 1];
#pragma unroll
    for (int q = 0; q < (D + 3) / 4; ++q) {
      va[q] = ldg_stream_f4(a + base + q * 128 + lane * 4);
      vc[q] = ldg_stream_f4(c + base + q * 128 + lane * 4);
      vr[q] = ldg_stream_u2(r + base + q * 128 + lane * 4);
    }
#pragma unroll
    for (int q = 0; q < (D + 3) / 4; ++q) acc += va[q].x + vc[q].w + (float)vr[q].y;
  }
  if (acc == 123.456f) out[0] = acc;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int D>
__global__ void __launch_bounds__(512, 1) k_slab_tma(const float* __restrict__ a, const float* __restrict__ c, const unsigned short* __restrict__ r,
                                                     long long n_slabs, float* out, int depth) {
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr int SB = 320 * D;  // bytes per slab
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);                  // 16 warps x 4 barriers
  unsigned char* buf = smem + 1024 + (size_t)warp * depth * SB;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 64; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bars + i)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const long long W = (long long)gridDim.x * 16;
  auto issue = [&](long long sl, int slot) {
    if (lane == 0) {
      const size_t base = (size_t)sl * 32 * D;
      uint32_t bar = smem_u32(bars + warp * 4 + slot), dst = smem_u32(buf + slot * SB);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(SB) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(a + base), "r"(128 * D), "r"(bar) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst + 128 * D), "l"(c + base), "r"(128 * D), "r"(bar) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst + 256 * D), "l"(r + base), "r"(64 * D), "r"(bar) : "memory");
    }
  };
  long long sl0 = (long long)blockIdx.x * 16 + warp;
  for (int s = 0; s < depth; ++s)
    if (sl0 + s * W < n_slabs) issue(sl0 + s * W, s);
  float acc = 0.f;
  uint32_t phase_bits = 0;
  int it = 0;
  for (long long sl = sl0; sl < n_slabs; sl += W, ++it) {
    const int slot = it % depth;
    const uint32_t bar = smem_u32(bars + warp * 4 + slot);
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"((phase_bits >> slot) & 1u) : "memory");
    }
    phase_bits ^= 1u << slot;
    const unsigned char* b = buf + slot * SB;
#pragma unroll
    for (int q = 0; q < (D + 3) / 4; ++q) {
      float4 va = *reinterpret_cast<const float4*>(b + q * 512 + lane * 16);
      float4 vc = *reinterpret_cast<const float4*>(b + 128 * D + q * 512 + lane * 16);
      uint2 vr = *reinterpret_cast<const uint2*>(b + 256 * D + q * 256 + lane * 8);
      acc += va.x + vc.w + (float)vr.y;
    }
    __syncwarp();
    if (sl + (long long)depth * W < n_slabs) issue(sl + (long long)depth * W, slot);
  }
  if (acc == 123.456f) out[0] = acc;
}

int main() {
  constexpr int D = 12;  // 3.84 KB per slab (the benchmark's mean column length is 10)
  const long long n_slabs = 2600000;  // ~10 GB
  const size_t n = (size_t)n_slabs * 32 * D;
  float *a, *c, *out;
  unsigned short* r;
  CK(cudaMalloc(&a, n * 4));
  CK(cudaMalloc(&c, n * 4));
  CK(cudaMalloc(&r, n * 2));
  CK(cudaMalloc(&out, 16));
  CK(cudaMemset(a, 0, n * 4));
  CK(cudaMemset(c, 0, n * 4));
  CK(cudaMemset(r, 0, n * 2));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  auto time = [&](const char* name, double bytes, auto launch) {
    for (int i = 0; i < 2; ++i) launch();
    cudaEventRecord(e0);
    for (int i = 0; i < 5; ++i) launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("%-44s %8.3f ms  %7.1f GB/s  (%s)\n", name, ms / 5, bytes / (ms / 5 * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
  };
  const double total = (double)n * 10.0;
  time("linear LDG.128, one array (4 B/elem)", (double)n * 4.0, [&] { k_linear<<<148 * 8, 256>>>((const float4*)a, n / 4, out); });
  time("linear LDG.128, 148x2 CTAs x 1024", (double)n * 4.0, [&] { k_linear<<<148 * 2, 1024>>>((const float4*)a, n / 4, out); });
  time("slab pattern, LDG, 148 x 512", total, [&] { k_slab<D><<<148, 512>>>(a, c, r, n_slabs, out); });
  for (int depth = 1; depth <= 3; ++depth) {
    const size_t smem = 1024 + (size_t)16 * depth * 320 * D;
    cudaFuncSetAttribute(k_slab_tma<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    char nm[64];
    snprintf(nm, sizeof nm, "slab pattern, TMA depth %d, 148 x 512", depth);
    time(nm, total, [&] { k_slab_tma<D><<<148, 512, smem>>>(a, c, r, n_slabs, out, depth); });
  }
  // same with 80 KB of extra shared memory requested (as the product kernel: lambda + accumulator)
  for (int depth = 1; depth <= 2; ++depth) {
    const size_t smem = 1024 + (size_t)16 * depth * 320 * D + 81920;
    cudaFuncSetAttribute(k_slab_tma<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    char nm[64];
    snprintf(nm, sizeof nm, "slab TMA depth %d + 80 KB smem", depth);
    time(nm, total, [&] { k_slab_tma<D><<<148, 512, smem>>>(a, c, r, n_slabs, out, depth); });
  }
  return 0;
}

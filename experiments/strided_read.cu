// Micro-benchmark: cost of reading 16 B out of every 256-B row (the key dims of a [n,64] fp32 state).
// Variants: 4 scalar __ldg, one float4 ld.global.nc, L2 prefetch-size hints, and the device-wide
// cudaLimitMaxL2FetchGranularity setting.   nvcc -arch=sm_100a -O3 strided_read.cu -o strided_read
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>

template <int MODE>
__global__ void __launch_bounds__(256) k_read(const float* __restrict__ y, uint64_t n, int ld, float* out) {
  float lo = 1e30f, hi = -1e30f;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const float* r = y + i * ld;
    float4 v;
    if (MODE == 0) { v.x = __ldg(r); v.y = __ldg(r + 1); v.z = __ldg(r + 2); v.w = __ldg(r + 3); }
    else if (MODE == 1) { v = __ldg(reinterpret_cast<const float4*>(r)); }
    else if (MODE == 2) { asm volatile("ld.global.nc.L1::no_allocate.L2::64B.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(r)); }
    else if (MODE == 3) { asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(r)); }
    else if (MODE == 4) { asm volatile("ld.global.cs.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(r)); }
    else { uint32_t a0,a1,a2,a3,a4,a5,a6,a7; asm volatile("ld.global.L2::evict_first.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3), "=r"(a4), "=r"(a5), "=r"(a6), "=r"(a7) : "l"(r)); v.x=__uint_as_float(a0); v.y=__uint_as_float(a1); v.z=__uint_as_float(a2); v.w=__uint_as_float(a3^a4^a5^a6^a7); }
    lo = fminf(lo, fminf(fminf(v.x, v.y), fminf(v.z, v.w)));
    hi = fmaxf(hi, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
  }
  if (lo > hi) out[0] = lo;  // never
  if (threadIdx.x == 0 && blockIdx.x == 0) out[1] = lo + hi;
}

template <int MODE>
float run(const float* y, uint64_t n, int ld, float* out, int grid) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 2; ++i) k_read<MODE><<<grid, 256>>>(y, n, ld, out);
  cudaEventRecord(a);
  for (int i = 0; i < 5; ++i) k_read<MODE><<<grid, 256>>>(y, n, ld, out);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  return ms / 5;
}

int main(int argc, char** argv) {
  const uint64_t n = 1ull << 24; const int ld = 64;
  float* y; float* out;
  cudaMalloc(&y, n * ld * 4); cudaMalloc(&out, 16);
  cudaMemset(y, 0, n * ld * 4);
  size_t g = 0; cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity);
  printf("default L2 fetch granularity = %zu\n", g);
  for (int pass = 0; pass < 3; ++pass) {
    if (pass == 1) { cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, 32); cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity); printf("set 32 -> %s, now %zu\n", cudaGetErrorString(e), g); }
    if (pass == 2) { cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, 128); cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity); printf("set 128 -> %s, now %zu\n", cudaGetErrorString(e), g); }
    for (int grid : {148 * 4, 148 * 8, 148 * 16}) {
      printf("grid %5d: ldg x4 %.3f | ldg.v4 %.3f | nc.L2::64B %.3f | nc.noalloc %.3f | cs %.3f | evict_first %.3f ms\n", grid,
             run<0>(y, n, ld, out, grid), run<1>(y, n, ld, out, grid), run<2>(y, n, ld, out, grid),
             run<3>(y, n, ld, out, grid), run<4>(y, n, ld, out, grid), run<5>(y, n, ld, out, grid));
    }
  }
  printf("ideal at 32 B/row: %.3f ms; 64 B/row: %.3f; 128 B/row: %.3f (6.5 TB/s)\n", n * 32 / 6.5e9, n * 64 / 6.5e9, n * 128 / 6.5e9);
  return 0;
}

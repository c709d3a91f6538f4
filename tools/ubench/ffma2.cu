// Micro-benchmark (GPU box): packed fp32 FMA (fma.rn.f32x2 -> FFMA2) against scalar FFMA, cycles per 32 results per warp,
// for 1 / 2 / 4 warps per scheduler.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2 ffma2.cu && ./ffma2
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ uint64_t pk(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
template <int MODE>
__global__ void k(float* out, long long* cyc, int iters, float c0, float c1) {
  float a[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) a[i] = threadIdx.x * 0.001f + i;
  uint64_t p[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) p[i] = pk(a[2 * i], a[2 * i + 1]);
  const uint64_t m = pk(c0, c0), b = pk(c1, c1);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 32; ++i) a[i] = fmaf(a[i], c0, c1);
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(m), "l"(b));
    }
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) s += a[i];
#pragma unroll
  for (int i = 0; i < 16; ++i) s += __uint_as_float(static_cast<uint32_t>(p[i])) + __uint_as_float(static_cast<uint32_t>(p[i] >> 32));
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMallocManaged(&cyc, 8);
  const int iters = 4000;
  for (int mode = 0; mode < 2; ++mode)
    for (int warps : {4, 8, 16}) {
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) k<0><<<1, warps * 32>>>(out, cyc, iters, 1.0001f, 0.5f); else k<1><<<1, warps * 32>>>(out, cyc, iters, 1.0001f, 0.5f);
        cudaDeviceSynchronize();
      }
      printf("%-22s %d warp(s)/scheduler: %6.2f cycles per 32 results per warp\n", mode ? "16 x FFMA2 (f32x2)" : "32 x FFMA", warps / 4, double(*cyc) / iters);
    }
  return 0;
}

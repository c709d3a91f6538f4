// Micro-benchmark (GPU box): the attention softmax's exponential phase in isolation -- 64 FFMA, then 64 ex2 with the
// 32 fp16 packs trailing by LAG pairs -- for 1 and 2 warps per scheduler.  Also a dependent ex2 chain (latency).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o exps exps.cu && ./exps
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t pack(float a, float b) { uint32_t r; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a)); return r; }
template <int LAG>
__global__ void k(uint32_t* out, long long* cyc, int iters, float m) {
  float v[64];
  uint32_t acc = 0;
#pragma unroll
  for (int c = 0; c < 64; ++c) v[c] = -(threadIdx.x * 0.01f + c * 0.05f);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t pk[32];
    float e[64];
#pragma unroll
    for (int c = 0; c < 64; ++c) e[c] = fmaf(v[c], 1.4426950408889634f, -m);
#pragma unroll
    for (int c2 = 0; c2 < 32 + LAG; ++c2) {
      if (c2 < 32) { e[2 * c2] = ex2(e[2 * c2]); e[2 * c2 + 1] = ex2(e[2 * c2 + 1]); }
      if (c2 >= LAG) pk[c2 - LAG] = pack(e[2 * (c2 - LAG)], e[2 * (c2 - LAG) + 1]);
    }
#pragma unroll
    for (int c = 0; c < 32; ++c) acc ^= pk[c];
    m += 0.001f;
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_chain(float* out, long long* cyc, int iters) {
  float a = -0.5f - threadIdx.x * 0.001f;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) { a = ex2(a); a = -a; }
  }
  long long t1 = clock64();
  out[threadIdx.x] = a;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  uint32_t* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMallocManaged(&cyc, 8);
  const int iters = 2000;
  for (int lag : {0, 1, 2, 4, 8, 32})
    for (int warps : {4, 8}) {
      for (int rep = 0; rep < 2; ++rep) {
        switch (lag) {
          case 0: k<0><<<1, warps * 32>>>(out, cyc, iters, 1.f); break; case 1: k<1><<<1, warps * 32>>>(out, cyc, iters, 1.f); break;
          case 2: k<2><<<1, warps * 32>>>(out, cyc, iters, 1.f); break; case 4: k<4><<<1, warps * 32>>>(out, cyc, iters, 1.f); break;
          case 8: k<8><<<1, warps * 32>>>(out, cyc, iters, 1.f); break; case 32: k<32><<<1, warps * 32>>>(out, cyc, iters, 1.f); break;
        }
        cudaDeviceSynchronize();
      }
      printf("64 FFMA + 64 ex2 + 32 packs, source lag %2d pairs, %d warp(s)/scheduler: %7.1f cycles per round per warp\n", lag, warps / 4, double(*cyc) / iters);
    }
  for (int rep = 0; rep < 2; ++rep) { k_chain<<<1, 32>>>(reinterpret_cast<float*>(out), cyc, iters); cudaDeviceSynchronize(); }
  printf("dependent ex2 + FADD(neg) chain: %.1f cycles per link\n", double(*cyc) / iters / 16);
  return 0;
}

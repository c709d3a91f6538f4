// Micro-benchmark (GPU box): issue cost in cycles per warp-instruction of the instructions the attention softmax is made
// of, for 1, 2 and 4 warps per scheduler.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#define REP 64
template <int OP>
__global__ void k(float* out, long long* cyc, int iters) {
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
  float m0 = 0.f, m1 = 1.f, m2 = 2.f, m3 = 3.f;
  uint32_t u0 = threadIdx.x, u1 = 7;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < REP / 16; ++r) {
      if (OP == 0) {  // MUFU.EX2, 16 independent
#pragma unroll
        for (int i = 0; i < 16; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      } else if (OP == 1) {  // FMNMX3 (4 chains)
#pragma unroll
        for (int i = 0; i < 16; i += 8) {
          m0 = fmaxf(fmaxf(m0, a[i]), a[i + 4]); m1 = fmaxf(fmaxf(m1, a[i + 1]), a[i + 5]);
          m2 = fmaxf(fmaxf(m2, a[i + 2]), a[i + 6]); m3 = fmaxf(fmaxf(m3, a[i + 3]), a[i + 7]);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] += 1.0f;   // keep values changing (16 FADD)
      } else if (OP == 2) {  // FFMA, 16 independent
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], 1.0001f, 0.5f);
      } else if (OP == 3) {  // F2FP pack, 8 per 16 values
#pragma unroll
        for (int i = 0; i < 16; i += 2) { uint32_t r; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a[i]), "f"(a[i + 1])); u0 ^= r; }
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] += 1.0f;
      } else if (OP == 4) {  // FADD only (baseline for 1 and 3)
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] += 1.0f;
      } else if (OP == 5) {  // FMNMX 2-input, 8 per round in 4 chains via asm (no fusing: integer op in between)
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          asm volatile("max.f32 %0, %0, %1;" : "+f"(m0) : "f"(a[i])); asm volatile("max.f32 %0, %0, %1;" : "+f"(m1) : "f"(a[i + 1]));
          asm volatile("max.f32 %0, %0, %1;" : "+f"(m2) : "f"(a[i + 2])); asm volatile("max.f32 %0, %0, %1;" : "+f"(m3) : "f"(a[i + 3]));
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] += 1.0f;
      }
    }
  }
  long long t1 = clock64();
  float s = m0 + m1 + m2 + m3 + __uint_as_float(u0 ^ u1);
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMallocManaged(&cyc, 8);
  const char* names[] = {"MUFU.EX2 x16", "FMNMX3 x8 + FADD x16", "FFMA x16", "F2FP x8 + FADD x16", "FADD x16", "FMNMX x16 + FADD x16"};
  const int iters = 2000;
  for (int op = 0; op < 6; ++op)
    for (int warps : {4, 8, 16}) {   // 1, 2, 4 warps per scheduler
      for (int rep = 0; rep < 2; ++rep) {
        switch (op) {
          case 0: k<0><<<1, warps * 32>>>(out, cyc, iters); break; case 1: k<1><<<1, warps * 32>>>(out, cyc, iters); break;
          case 2: k<2><<<1, warps * 32>>>(out, cyc, iters); break; case 3: k<3><<<1, warps * 32>>>(out, cyc, iters); break;
          case 4: k<4><<<1, warps * 32>>>(out, cyc, iters); break; case 5: k<5><<<1, warps * 32>>>(out, cyc, iters); break;
        }
        cudaDeviceSynchronize();
      }
      printf("%-24s %2d warps/SM: %7.2f cycles per round of 16 values per warp\n", names[op], warps, double(*cyc) / (iters * (REP / 16)));
    }
  return 0;
}

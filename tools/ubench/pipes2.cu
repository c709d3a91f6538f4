// Micro-benchmark (GPU box), second set: what the attention softmax warps compete for.
//   1. ex2.approx.f16x2 vs ex2.approx.ftz.f32: cycles per warp instruction (two results per lane vs one)
//   2. does a MUFU-bound warp block the issue port of its scheduler?  warps 0-3 run ex2 only, warps 4-7 (same four
//      schedulers) run FFMA only; each class is timed alone and together
//   3. tcgen05.ld 32x32b.x32 / .x64 throughput: 4 and 8 warps streaming 64 fp32 columns per round from TMEM
//   4. tcgen05.st 32x32b.x32 throughput
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes2 pipes2.cu && ./pipes2
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// MODE 0: ex2.f32 x16 per round; 1: ex2.f16x2 x16 per round (32 results); 2: cvt.f16x2 + ex2.f16x2 (8 + 8 per 16 values)
template <int MODE>
__global__ void k_ex2(float* out, long long* cyc, int iters) {
  float a[16];
  uint32_t h[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) { a[i] = -(threadIdx.x * 0.001f + i * 0.1f); h[i] = 0xB800B400u + i; }
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
    } else if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 16; ++i) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[i]));
    } else {
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        uint32_t r;
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a[i]), "f"(a[i + 1]));
        asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(r));
        h[i >> 1] ^= r;
      }
    }
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i] + __uint_as_float(h[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

// warps 0-3: ex2 (f32 when F16 == 0, f16x2 otherwise), warps 4-7: FFMA.  `who` bit 0 runs the ex2 warps, bit 1 the FFMA warps.
template <int F16>
__global__ void k_coissue(float* out, long long* cyc, int iters_mufu, int iters_fma, int who) {
  float a[16];
  uint32_t h[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) { a[i] = -(threadIdx.x * 0.001f + i * 0.1f); h[i] = 0xB800B400u + i; }
  const int warp = threadIdx.x >> 5;
  __syncthreads();
  long long t0 = clock64();
  if (warp < 4) {
    if (who & 1)
      for (int it = 0; it < iters_mufu; ++it) {
        if (F16) {
#pragma unroll
          for (int i = 0; i < 16; ++i) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[i]));
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
        }
      }
  } else if (who & 2) {
    for (int it = 0; it < iters_fma; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], 1.0001f, 0.5f);
    }
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i] + __uint_as_float(h[i]);
  out[threadIdx.x] = s;
  if ((threadIdx.x & 31) == 0) cyc[warp] = t1 - t0;
}

// TMEM streaming: each warp reads (LD) or writes (ST) 64 columns x its 32 lanes per round.
// SHAPE 0: two .x32 loads then one wait; 1: one .x64 load; 2: one .x32 store + wait::st; 3: four .x16 loads
template <int SHAPE>
__global__ void k_tmem(float* out, long long* cyc, int iters) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16) + (warp >> 2) * 128;
  uint32_t acc = 0;
  uint32_t r[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) r[i] = threadIdx.x + i;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const uint32_t ta = base + (it & 1) * 64;
    if (SHAPE == 0 || SHAPE == 3) {
      if (SHAPE == 0) {
#pragma unroll
        for (int g = 0; g < 2; ++g)
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
              "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
              : "=r"(r[g * 32 + 0]), "=r"(r[g * 32 + 1]), "=r"(r[g * 32 + 2]), "=r"(r[g * 32 + 3]), "=r"(r[g * 32 + 4]), "=r"(r[g * 32 + 5]),
                "=r"(r[g * 32 + 6]), "=r"(r[g * 32 + 7]), "=r"(r[g * 32 + 8]), "=r"(r[g * 32 + 9]), "=r"(r[g * 32 + 10]), "=r"(r[g * 32 + 11]),
                "=r"(r[g * 32 + 12]), "=r"(r[g * 32 + 13]), "=r"(r[g * 32 + 14]), "=r"(r[g * 32 + 15]), "=r"(r[g * 32 + 16]), "=r"(r[g * 32 + 17]),
                "=r"(r[g * 32 + 18]), "=r"(r[g * 32 + 19]), "=r"(r[g * 32 + 20]), "=r"(r[g * 32 + 21]), "=r"(r[g * 32 + 22]), "=r"(r[g * 32 + 23]),
                "=r"(r[g * 32 + 24]), "=r"(r[g * 32 + 25]), "=r"(r[g * 32 + 26]), "=r"(r[g * 32 + 27]), "=r"(r[g * 32 + 28]), "=r"(r[g * 32 + 29]),
                "=r"(r[g * 32 + 30]), "=r"(r[g * 32 + 31])
              : "r"(ta + g * 32) : "memory");
      } else {
#pragma unroll
        for (int g = 0; g < 4; ++g)
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
              : "=r"(r[g * 16 + 0]), "=r"(r[g * 16 + 1]), "=r"(r[g * 16 + 2]), "=r"(r[g * 16 + 3]), "=r"(r[g * 16 + 4]), "=r"(r[g * 16 + 5]),
                "=r"(r[g * 16 + 6]), "=r"(r[g * 16 + 7]), "=r"(r[g * 16 + 8]), "=r"(r[g * 16 + 9]), "=r"(r[g * 16 + 10]), "=r"(r[g * 16 + 11]),
                "=r"(r[g * 16 + 12]), "=r"(r[g * 16 + 13]), "=r"(r[g * 16 + 14]), "=r"(r[g * 16 + 15])
              : "r"(ta + g * 16) : "memory");
      }
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int i = 0; i < 64; i += 8) acc ^= r[i];
    } else if (SHAPE == 1) {
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
          "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,"
          "%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
            "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
            "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
            "=r"(r[30]), "=r"(r[31]), "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]),
            "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]), "=r"(r[49]),
            "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]),
            "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
          : "r"(ta) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int i = 0; i < 64; i += 8) acc ^= r[i];
    } else {
      asm volatile(
          "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
          "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
          ::"r"(ta), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
            "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
            "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
            "r"(r[30]), "r"(r[31]) : "memory");
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      r[it & 31] += 1;
    }
  }
  long long t1 = clock64();
  out[threadIdx.x] = __uint_as_float(acc ^ r[5]);
  if ((threadIdx.x & 31) == 0) cyc[warp] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512u) : "memory");
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 1 << 20); cudaMallocManaged(&cyc, 64 * 8);
  const int iters = 2000;
  const char* n1[] = {"ex2.f32 x16", "ex2.f16x2 x16 (32 results)", "cvt.f16x2 x8 + ex2.f16x2 x8 (16 results)"};
  for (int m = 0; m < 3; ++m)
    for (int warps : {4, 8, 16}) {
      for (int rep = 0; rep < 2; ++rep) {
        if (m == 0) k_ex2<0><<<1, warps * 32>>>(out, cyc, iters);
        if (m == 1) k_ex2<1><<<1, warps * 32>>>(out, cyc, iters);
        if (m == 2) k_ex2<2><<<1, warps * 32>>>(out, cyc, iters);
        cudaDeviceSynchronize();
      }
      printf("%-42s %2d warps/SM: %7.2f cycles per round per warp\n", n1[m], warps, double(cyc[0]) / iters);
    }
  for (int f16 = 0; f16 < 2; ++f16)
    for (int who = 1; who <= 3; ++who) {
      for (int rep = 0; rep < 2; ++rep) {
        if (f16) k_coissue<1><<<1, 256>>>(out, cyc, iters, iters * 7, who); else k_coissue<0><<<1, 256>>>(out, cyc, iters, iters * 7, who);
        cudaDeviceSynchronize();
      }
      printf("co-issue %s who=%d: ex2 warp 0: %9lld cycles (%.1f / round of 16)   FFMA warp 4: %9lld cycles (%.1f / round of 16)\n",
             f16 ? "ex2.f16x2" : "ex2.f32  ", who, cyc[0], double(cyc[0]) / iters, cyc[4], double(cyc[4]) / (iters * 7));
    }
  const char* n3[] = {"tcgen05.ld 2 x .x32 + wait", "tcgen05.ld .x64 + wait", "tcgen05.st .x32 + wait::st", "tcgen05.ld 4 x .x16 + wait"};
  for (int s = 0; s < 4; ++s)
    for (int warps : {4, 8, 16}) {
      for (int rep = 0; rep < 2; ++rep) {
        if (s == 0) k_tmem<0><<<1, warps * 32>>>(out, cyc, iters);
        if (s == 1) k_tmem<1><<<1, warps * 32>>>(out, cyc, iters);
        if (s == 2) k_tmem<2><<<1, warps * 32>>>(out, cyc, iters);
        if (s == 3) k_tmem<3><<<1, warps * 32>>>(out, cyc, iters);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: %s\n", n3[s], cudaGetErrorString(e)); return 1; }
      }
      const double c = double(cyc[0]) / iters;
      const double bytes = (s == 2 ? 32 : 64) * 32 * 4.0;
      printf("%-30s %2d warps/SM: %7.1f cycles per round per warp -> %6.1f B/clk per warp, %7.1f B/clk per SM\n", n3[s], warps, c,
             bytes / c, bytes * warps / c);
    }
  return 0;
}

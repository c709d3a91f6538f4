// Micro-benchmark (GPU box), third set: hand-off latencies of the attention pipeline, one CTA.
//   A. mbarrier arrive -> try_wait returns in another warp (ping-pong between warp 0 and warp W; W = 1: another
//      scheduler, W = 4: the same scheduler), with the partner either spinning on try_wait or on test_wait
//   B. n x tcgen05.mma (SS, M=128, N=64|128, K=16) + tcgen05.commit -> mbarrier seen by the issuing thread
//   C. the same with the A operand in TMEM (N = 80, the PV shape)
//   D. softmax hand-over round trip: warp 4 tcgen05.st x32 + wait::st + fence + arrive -> warp 1 waits, fences, issues
//      4 TS MMAs + commit -> warp 4 waits for the commit
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../protein_gibbs_sampler_b200/csrc -o latency latency.cu -lcuda
#include <cstdio>
#include <cstdint>
#include "ptx.cuh"
using namespace pg;

__device__ __forceinline__ void spin_wait(uint64_t* bar, uint32_t parity, int use_test) {
  if (use_test) { while (!mbar_test_wait(bar, parity)) { } }
  else { while (!mbar_try_wait(bar, parity)) { } }
}

__global__ void k_pingpong(long long* cyc, int iters, int partner, int use_test) {
  __shared__ uint64_t bars[2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); fence_barrier_init(); }
  __syncthreads();
  long long t0 = clock64();
  if (warp == 0) {
    for (int it = 0; it < iters; ++it) {
      if (lane == 0) mbar_arrive(&bars[0]);
      __syncwarp();
      spin_wait(&bars[1], it & 1, use_test);
    }
  } else if (warp == partner) {
    for (int it = 0; it < iters; ++it) {
      spin_wait(&bars[0], it & 1, use_test);
      if (lane == 0) mbar_arrive(&bars[1]);
      __syncwarp();
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) *cyc = t1 - t0;
}

// MODE 0: SS S-shape (N = nn); 1: TS PV-shape (N = 80, B MN-major)
__global__ void __launch_bounds__(256, 1) k_mma(long long* cyc, int iters, int n_mma, int nn, int mode) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = slot;
  if (warp == 1) {
    const bool issuer = elect_one();
    const uint64_t ad = make_smem_desc_sw128(smem_u32(smem), 1024);
    const uint64_t bd = make_smem_desc_sw128(smem_u32(smem + 16384), 1024);
    const uint64_t vd = make_smem_desc_sw128(smem_u32(smem + 16384), 1024, 16384);
    const uint32_t idS = make_idesc_f16(128, nn, false, false), idPV = make_idesc_f16(128, 80, false, true);
    long long tot = 0;
    for (int it = 0; it < iters; ++it) {
      __syncwarp();
      long long t0 = clock64();
      if (issuer) {
        for (int k = 0; k < n_mma; ++k) {
          if (mode == 0) umma_f16_ss(tb, ad + 2 * (k & 3), bd + 2 * (k & 3), idS, k ? 1u : 0u);
          else umma_f16_ts(tb + 256, tb + 8 * (k & 3), vd + (k & 3) * (2048 >> 4), idPV, k ? 1u : 0u);
        }
        umma_commit(&bar);
      }
      __syncwarp();
      while (!mbar_try_wait(&bar, it & 1)) { }
      tc_fence_after();
      long long t1 = clock64();
      tot += t1 - t0;
    }
    if (issuer) *cyc = tot;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

// E. TMEM hazards between consecutive MMAs: rounds of [4 TS PV reading A at columns X, 4 SS S writing columns X (hazard=1:
//    the attention kernel's in-place P over S) or other columns (hazard=0)], 8 rounds per commit.
__global__ void __launch_bounds__(256, 1) k_hazard(long long* cyc, int iters, int hazard, int which) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = slot;
  if (warp == 1) {
    const bool issuer = elect_one();
    const uint64_t ad = make_smem_desc_sw128(smem_u32(smem), 1024);
    const uint64_t bd = make_smem_desc_sw128(smem_u32(smem + 16384), 1024);
    const uint64_t vd = make_smem_desc_sw128(smem_u32(smem + 16384), 1024, 16384);
    const uint32_t idS = make_idesc_f16(128, 64, false, false), idPV = make_idesc_f16(128, 80, false, true);
    long long tot = 0;
    for (int it = 0; it < iters; ++it) {
      __syncwarp();
      long long t0 = clock64();
      if (issuer) {
#pragma unroll 1
        for (int r = 0; r < 8; ++r) {
          const uint32_t pcol = tb + (r & 1) * 64;                          // P(i) lives in S buffer i & 1
          const uint32_t scol = hazard ? pcol : tb + 128 + (r & 1) * 64;    // S(i+2) overwrites it, or goes elsewhere
          if (which & 1) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16_ts(tb + 256, pcol + 8 * k, vd + k * (2048 >> 4), idPV, (r | k) ? 1u : 0u);
          }
          if (which & 2) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16_ss(scol, ad + 2 * k, bd + 2 * k, idS, k ? 1u : 0u);
          }
        }
        umma_commit(&bar);
      }
      __syncwarp();
      while (!mbar_try_wait(&bar, it & 1)) { }
      tc_fence_after();
      long long t1 = clock64();
      tot += t1 - t0;
    }
    if (issuer) *cyc = tot;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

// F. The attention issuers' steady state without the softmax: warp 1 (and, with two = 1, warp 3 on its own TMEM columns)
//    issue rounds of [4 TS PV + commit, 4 SS S + commit] back to back; cycles per round per issuer.
__global__ void __launch_bounds__(256, 1) k_issuers(long long* cyc, int iters, int two, int commits) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bars[8];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u;
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bars[i], 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = slot;
  if (warp == 1 || (two && warp == 3)) {
    const int t = warp >> 1;
    const bool issuer = elect_one();
    const uint64_t ad = make_smem_desc_sw128(smem_u32(smem + t * 16384), 1024);
    const uint64_t bd = make_smem_desc_sw128(smem_u32(smem + 32768), 1024);
    const uint64_t vd = make_smem_desc_sw128(smem_u32(smem + 32768), 1024, 16384);
    const uint32_t idS = make_idesc_f16(128, 64, false, false), idPV = make_idesc_f16(128, 80, false, true);
    const uint32_t tS0 = tb + t * 128, tO = tb + 256 + t * 80;
    __syncwarp();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint32_t pcol = tS0 + (it & 1) * 64;
      if (issuer) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16_ts(tO, pcol + 8 * k, vd + k * (2048 >> 4), idPV, (it | k) ? 1u : 0u);
        if (commits) umma_commit(&bars[4 * t]);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16_ss(pcol, ad + 2 * k, bd + 2 * k, idS, k ? 1u : 0u);
        if (commits) umma_commit(&bars[4 * t + 1]);
      }
      __syncwarp();
    }
    if (issuer) umma_commit(&bars[4 * t + 2]);
    __syncwarp();
    while (!mbar_try_wait(&bars[4 * t + 2], 0)) { }
    long long t1 = clock64();
    if (issuer) cyc[t] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

__global__ void __launch_bounds__(256, 1) k_handover(long long* cyc, int iters) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bars[2];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u;
  if (threadIdx.x == 0) { mbar_init(&bars[0], 4); mbar_init(&bars[1], 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = slot;
  if (warp == 1) {
    const bool issuer = elect_one();
    const uint64_t vd = make_smem_desc_sw128(smem_u32(smem + 16384), 1024, 16384);
    const uint32_t idPV = make_idesc_f16(128, 80, false, true);
    for (int it = 0; it < iters; ++it) {
      while (!mbar_try_wait(&bars[0], it & 1)) { }
      tc_fence_after();
      if (issuer) {
        for (int k = 0; k < 4; ++k) umma_f16_ts(tb + 256, tb + 8 * k, vd + k * (2048 >> 4), idPV, k ? 1u : 0u);
        umma_commit(&bars[1]);
      }
      __syncwarp();
    }
  } else if (warp >= 4) {
    const uint32_t ta = tb + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    uint32_t r[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) r[i] = 0x3C003C00u;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      tmem_st32(ta, r);
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[0]);
      while (!mbar_try_wait(&bars[1], it & 1)) { }
      tc_fence_after();
    }
    long long t1 = clock64();
    if (threadIdx.x == 128) *cyc = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

int main() {
  long long* cyc; cudaMallocManaged(&cyc, 64);
  const int iters = 2000;
  for (int partner : {1, 4})
    for (int use_test = 0; use_test < 2; ++use_test) {
      for (int rep = 0; rep < 2; ++rep) { k_pingpong<<<1, 256>>>(cyc, iters, partner, use_test); cudaDeviceSynchronize(); }
      printf("mbarrier ping-pong warp 0 <-> warp %d (%s): %.1f cycles one way\n", partner, use_test ? "test_wait" : "try_wait ",
             double(*cyc) / iters / 2);
    }
  cudaFuncSetAttribute(k_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000);
  cudaFuncSetAttribute(k_handover, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000);
  for (int mode = 0; mode < 2; ++mode)
    for (int nn : {64, 128}) {
      if (mode == 1 && nn == 128) continue;
      for (int n : {1, 2, 4, 8, 16}) {
        for (int rep = 0; rep < 2; ++rep) {
          k_mma<<<1, 256, 70000>>>(cyc, iters, n, nn, mode);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("k_mma: %s\n", cudaGetErrorString(e)); return 1; }
        }
        printf("%s x %2d + commit -> seen: %.1f cycles\n", mode ? "TS MMA 128x80x16 (PV) " : (nn == 64 ? "SS MMA 128x64x16  (S) " : "SS MMA 128x128x16 (S) "),
               n, double(*cyc) / iters);
      }
    }
  cudaFuncSetAttribute(k_hazard, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000);
  for (int which = 1; which <= 3; ++which)
    for (int hazard = 0; hazard < 2; ++hazard) {
      for (int rep = 0; rep < 2; ++rep) {
        k_hazard<<<1, 256, 70000>>>(cyc, iters, hazard, which);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("k_hazard: %s\n", cudaGetErrorString(e)); return 1; }
      }
      printf("8 rounds of [%s%s] %s + commit -> seen: %.1f cycles (%.1f per round)\n", which & 1 ? "4 TS PV " : "", which & 2 ? "4 SS S" : "",
             hazard ? "S overwrites the P just read" : "S into other columns       ", double(*cyc) / iters, double(*cyc) / iters / 8);
    }
  cudaFuncSetAttribute(k_issuers, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000);
  for (int two = 0; two < 2; ++two)
    for (int commits = 0; commits < 2; ++commits) {
      for (int rep = 0; rep < 2; ++rep) {
        k_issuers<<<1, 256, 70000>>>(cyc, iters, two, commits);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("k_issuers: %s\n", cudaGetErrorString(e)); return 1; }
      }
      printf("%d issuer warp(s), rounds of [4 TS PV%s, 4 SS S%s]: %.1f cycles per round (issuer A)%s\n", two + 1, commits ? " + commit" : "",
             commits ? " + commit" : "", double(cyc[0]) / iters, two ? "" : "");
      if (two) printf("   issuer B: %.1f cycles per round\n", double(cyc[1]) / iters);
    }
  for (int rep = 0; rep < 2; ++rep) {
    k_handover<<<1, 256, 70000>>>(cyc, iters);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("k_handover: %s\n", cudaGetErrorString(e)); return 1; }
  }
  printf("P hand-over round trip (st x32, wait::st, arrive x4 -> issuer -> 4 TS MMAs + commit -> softmax warp): %.1f cycles\n", double(*cyc) / iters);
  return 0;
}

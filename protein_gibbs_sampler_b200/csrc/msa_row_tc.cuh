// MSA Transformer tied row attention on tcgen05 (head_dim 64, alignments of up to 256 columns).
//
//     s[b,h,i,j] = sum_r sum_d q[b,r,i,h,d] k[b,r,j,h,d]        (q pre-scaled by Dh^-1/2 / sqrt(R) in the QKV GEMM)
//     p = softmax_j(s) ;  ctx[b,r,i,h,:] = sum_j p[b,h,i,j] v[b,r,j,h,:]
//
// One CTA per (128-query tile, head, MSA).  Both halves are GEMMs whose reduction / batch dimension is the MSA depth R:
//   phase 1   S[128, NK] = sum_r Q_r[128, 64] K_r[NK, 64]^T     one accumulator in TMEM, R x 4 MMAs (K = 16 each),
//             Q_r / K_r tiles of alignment row r streamed by TMA through a shared-memory ring
//   softmax   one thread per query row, two passes over the NK score columns in TMEM (max, then 2^(s - max));
//             P overwrites S as packed fp16; the row sum is taken from the ROUNDED P (exact convex weights)
//   phase 2   O_r[128, 64] = P[128, NK] V_r[NK, 64] for every r    A = P from TMEM, B = V_r (MN-major) from the same
//             ring; O double-buffered in TMEM so the epilogue of row r (1/l, fp16, swizzled staging, TMA store)
//             overlaps the MMAs of row r + 1
// NK = C rounded up to 16 (keys >= C are zero-filled by the tensor map and masked in the softmax).
// Warps: 0 TMA producer, 1 MMA issuer, 2 TMEM allocator, 4-7 softmax + epilogue.
// Replaces fair-esm RowSelfAttention (esm/axial_attention.py; reference call sites
// /root/reference/src/pgen/esm_msa_sampler.py:136,236); the mma.sync kernels in msa_attention.cuh remain for other
// head sizes and wider alignments.
#pragma once
#include "ptx.cuh"
#include "attention_fa.cuh"  // fa_ex2, fa_bar_sync, tma_store_3d

namespace pg {

struct MsaRowParams {
  int R, C, H;   // alignment depth, columns (tokens per row), heads
  int NK;        // C rounded up to a multiple of 16 (<= 256)
  int stages;    // ring depth
  int reverse;   // walk the MSAs from the last one down (pgibbs_engine::zigzag)
};

constexpr int kMrThreads = 256;
constexpr int kMrQBytes = 128 * 64 * 2;  // 16 KB
__host__ __device__ constexpr int mr_stage_bytes(int NK) { return kMrQBytes + ((NK * 128 + 1023) & ~1023); }
__host__ __device__ constexpr int mr_smem_bytes(int NK, int stages) {
  return stages * mr_stage_bytes(NK) + 2 * kMrQBytes /*output staging*/ + 1024 /*align*/ + 256 /*barriers*/;
}

__global__ void __launch_bounds__(kMrThreads, 1)
msa_row_attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                            const __grid_constant__ CUtensorMap tmCtx, const MsaRowParams p) {
  extern __shared__ __align__(1024) uint8_t mr_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(mr_smem_raw) + 1023) & ~uintptr_t(1023));
  const int stage_bytes = mr_stage_bytes(p.NK);
  uint8_t* ring = smem;                                  // [stages][Q 16 KB | K or V: NK rows x 128 B]
  uint8_t* stage_out = smem + p.stages * stage_bytes;    // [2][16 KB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage_out + 2 * kMrQBytes);
  uint64_t* full = bars;               // [8]
  uint64_t* empty = full + 8;          // [8]
  uint64_t* s_full = empty + 8;        // scores complete
  uint64_t* p_ready = s_full + 1;      // P written (count 4 = warps)
  uint64_t* o_full = p_ready + 1;      // [2]
  uint64_t* o_free = o_full + 2;       // [2] (count 4)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_free + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i0 = blockIdx.x * 128, head = blockIdx.y, b = p.reverse ? gridDim.z - 1 - blockIdx.z : blockIdx.z;
  const int R = p.R, NK = p.NK, d = p.H * 64;
  const int kv_bytes = NK * 128;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmQ);
    prefetch_tmap(&tmKV);
    prefetch_tmap(&tmCtx);
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(s_full, 1);
    mbar_init(p_ready, 4);
    for (int s = 0; s < 2; ++s) { mbar_init(&o_full[s], 1); mbar_init(&o_free[s], 4); }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();               // set-up above overlaps the previous kernel's tail (programmatic dependent launch)
  pdl_launch_dependents();
  const uint32_t tS = tmem_base, tO = tmem_base + 256;  // S / P: columns [0, NK); O: 2 x 64 columns

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (whole warp, elected lane)
    const bool issuer = elect_one();
    for (int it = 0; it < 2 * R; ++it) {  // R score steps (Q_r, K_r), then R value steps (V_r)
      const int st = it % p.stages;
      mbar_wait(&empty[st], ((it / p.stages) & 1) ^ 1);
      uint8_t* slot = ring + st * stage_bytes;
      if (issuer) {
        if (it < R) {
          const int row = b * R + it;
          mbar_arrive_expect_tx(&full[st], kMrQBytes + kv_bytes);
          tma_load_3d(slot, &tmQ, &full[st], head * 64, i0, row);
          tma_load_3d(slot + kMrQBytes, &tmKV, &full[st], d + head * 64, 0, row);
        } else {
          mbar_arrive_expect_tx(&full[st], kv_bytes);
          tma_load_3d(slot + kMrQBytes, &tmKV, &full[st], 2 * d + head * 64, 0, b * R + (it - R));
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (whole warp, elected lane)
    const bool issuer = elect_one();
    const uint32_t idesc_s = make_idesc_f16(128, NK, false, false);
    constexpr uint32_t kIdescPV = make_idesc_f16(128, 64, false, true);  // B = V is MN-major
    const uint32_t desc_hi = static_cast<uint32_t>(make_smem_desc_sw128(0, 1024) >> 32);
    const uint32_t desc_hi_v = static_cast<uint32_t>(make_smem_desc_sw128(0, 1024, 1024) >> 32);
    auto desc = [](uint32_t hi, uint32_t lo) { return (static_cast<uint64_t>(hi) << 32) | lo; };
    for (int r = 0; r < R; ++r) {
      const int st = r % p.stages;
      mbar_wait(&full[st], (r / p.stages) & 1);
      tc_fence_after();
      const uint32_t slot = smem_u32(ring + st * stage_bytes);
      const uint32_t q_lo = static_cast<uint32_t>(make_smem_desc_sw128(slot, 1024));
      const uint32_t k_lo = static_cast<uint32_t>(make_smem_desc_sw128(slot + kMrQBytes, 1024));
      if (issuer) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16_ss(tS, desc(desc_hi, q_lo + 2 * k), desc(desc_hi, k_lo + 2 * k), idesc_s, (r | k) ? 1u : 0u);
        umma_commit(&empty[st]);
        if (r + 1 == R) umma_commit(s_full);
      }
      __syncwarp();
    }
    mbar_wait(p_ready, 0);
    tc_fence_after();
    for (int r = 0; r < R; ++r) {
      const int it = R + r, st = it % p.stages, ob = r & 1;
      mbar_wait(&full[st], (it / p.stages) & 1);
      mbar_wait(&o_free[ob], ((r >> 1) & 1) ^ 1);
      tc_fence_after();
      // V_r: rows = keys (the MMA's K), 64 contiguous head-dim values per row (the MMA's N): MN-major, 128B swizzle;
      // a K=16 step is two 8-row swizzle atoms = 2048 B.
      const uint32_t v_lo = static_cast<uint32_t>(make_smem_desc_sw128(smem_u32(ring + st * stage_bytes + kMrQBytes), 1024, 1024));
      if (issuer) {
        for (int k = 0; k < NK / 16; ++k)
          umma_f16_ts(tO + ob * 64, tS + 8 * k, desc(desc_hi_v, v_lo + k * (2048 >> 4)), kIdescPV, k ? 1u : 0u);
        umma_commit(&empty[st]);
        umma_commit(&o_full[ob]);
      }
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ softmax, then the epilogue of every row
    const int quad = warp & 3;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const bool warp_live = i0 + quad * 32 < p.C;  // warp-uniform: any valid query row in this warp
    const bool leader = (threadIdx.x & 127) == 0;
    constexpr float kLog2e = 1.4426950408889634f;
    float inv = 0.f;
    mbar_wait(s_full, 0);
    tc_fence_after();
    if (warp_live) {
      const int n16 = NK >> 4;  // 16-column groups
      float mx = -INFINITY;
      for (int c = 0; c < n16; ++c) {
        uint32_t r[16];
        tmem_ld16(tS + lane_off + c * 16, r);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (c * 16 + i < p.C) mx = fmaxf(mx, __uint_as_float(r[i]));
      }
      mx *= kLog2e;
      float l = 0.f;
      for (int c = 0; c < n16; c += 2) {  // 32 score columns -> 16 packed P columns
        uint32_t r0[16], r1[16], pk[16];
        tmem_ld16(tS + lane_off + c * 16, r0);
        if (c + 1 < n16) tmem_ld16(tS + lane_off + (c + 1) * 16, r1);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
          const float p0 = c * 16 + i < p.C ? fa_ex2(fmaf(__uint_as_float(r0[i]), kLog2e, -mx)) : 0.f;
          const float p1 = c * 16 + i + 1 < p.C ? fa_ex2(fmaf(__uint_as_float(r0[i + 1]), kLog2e, -mx)) : 0.f;
          __half2 h = __floats2half2_rn(p0, p1);
          const float2 f = __half22float2(h);
          l += f.x + f.y;
          pk[i >> 1] = *reinterpret_cast<uint32_t*>(&h);
        }
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
          const int col = (c + 1) * 16 + i;
          const bool have = c + 1 < n16;
          const float p0 = have && col < p.C ? fa_ex2(fmaf(__uint_as_float(r1[i]), kLog2e, -mx)) : 0.f;
          const float p1 = have && col + 1 < p.C ? fa_ex2(fmaf(__uint_as_float(r1[i + 1]), kLog2e, -mx)) : 0.f;
          __half2 h = __floats2half2_rn(p0, p1);
          const float2 f = __half22float2(h);
          l += f.x + f.y;
          pk[8 + (i >> 1)] = *reinterpret_cast<uint32_t*>(&h);
        }
        // P (fp16 x2 per column) over score columns that have already been consumed (16 c / 2 <= 16 c)
        tmem_st16(tS + lane_off + (c >> 1) * 16, pk);
      }
      tmem_wait_st();
      inv = 1.0f / l;
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(p_ready);

    for (int r = 0; r < R; ++r) {
      const int ob = r & 1;
      uint8_t* stage = stage_out + ob * kMrQBytes;
      if (leader) tma_store_wait_read<1>();  // the store issued from this staging buffer two rows ago has read it
      fa_bar_sync(1, 128);
      mbar_wait(&o_full[ob], (r >> 1) & 1);
      tc_fence_after();
      if (warp_live) {
        const int row = quad * 32 + lane;
        uint8_t* rowp = stage + row * 128;
#pragma unroll
        for (int hlf = 0; hlf < 2; ++hlf) {
          uint32_t o[32];
          tmem_ld32(tO + ob * 64 + lane_off + hlf * 32, o);
          tmem_wait_ld();
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 w;
            __half2 h0 = __floats2half2_rn(__uint_as_float(o[8 * q]) * inv, __uint_as_float(o[8 * q + 1]) * inv);
            __half2 h1 = __floats2half2_rn(__uint_as_float(o[8 * q + 2]) * inv, __uint_as_float(o[8 * q + 3]) * inv);
            __half2 h2 = __floats2half2_rn(__uint_as_float(o[8 * q + 4]) * inv, __uint_as_float(o[8 * q + 5]) * inv);
            __half2 h3 = __floats2half2_rn(__uint_as_float(o[8 * q + 6]) * inv, __uint_as_float(o[8 * q + 7]) * inv);
            w.x = *reinterpret_cast<uint32_t*>(&h0); w.y = *reinterpret_cast<uint32_t*>(&h1);
            w.z = *reinterpret_cast<uint32_t*>(&h2); w.w = *reinterpret_cast<uint32_t*>(&h3);
            *reinterpret_cast<uint4*>(rowp + (((hlf * 4 + q) ^ (row & 7)) << 4)) = w;  // 128B swizzle, matches tmCtx
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_free[ob]);
      fence_proxy_async();  // generic-proxy smem writes -> visible to the TMA (async proxy)
      fa_bar_sync(2, 128);
      if (leader) {
        tma_store_3d(&tmCtx, stage, head * 64, i0, b * R + r);  // query rows >= C are clipped by the tensor map
        tma_store_commit();
      }
    }
    if (leader) tma_store_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace pg

// Self-attention on the 5th-generation tensor cores (tcgen05) for head_dim 64 -- ESM-1b / ESM-2 650M.
//
//   ctx[s, i, h, :] = softmax_j( q[s,i,h,:] . k[s,j,h,:] ) v[s,j,h,:]      q pre-scaled by Dh^-1/2 (+RoPE) by the
//                                                                           QKV GEMM epilogue; no padding mask
// One CTA = one 128-query tile of one (sequence, head); two CTAs are resident per SM so that one CTA's softmax
// (MUFU-bound) overlaps the other's MMAs.  Keys/values stream through a 2-stage TMA ring in blocks of 128:
//   warp 0     TMA producer (3-D tensor map over the fused qkv activation: rows t >= T are zero-filled)
//   warp 1     MMA issuer: S = Q K_j^T (SS, 128 x nk x 64) into TMEM, then O += P_j V_j (A = P from TMEM, B = V_j
//              as an MN-major shared-memory operand); also owns the 256-column TMEM allocation
//   warps 2-5  one thread per query row: online softmax in fp32 straight out of TMEM (tcgen05.ld), P written back
//              over S as packed fp16 (tcgen05.st), O rescaled in TMEM only when the running maximum moves,
//              final O / l -> fp16 ctx
// TMEM columns: S/P [0,128), O [128,192).  tcgen05.mma executes in issue order, so S_j (which overwrites P_{j-1})
// cannot start before O += P_{j-1} V_{j-1} has consumed P_{j-1}; s_full(j) therefore also tells the softmax warps
// that O is quiescent for the rescale.
// Replaces fair-esm MultiheadAttention's bmm / softmax / bmm (call site /root/reference/src/pgen/esm_sampler.py:223).
#pragma once
#include "ptx.cuh"

namespace pg {

struct AttnTcParams {
  __half* ctx;  // [n_seq*T, ldc]
  int T, ldc, d;  // d = heads * 64 (column offset of k is d, of v is 2d inside the qkv row)
};

constexpr int kAtThreads = 192;
constexpr int kAtStages = 2;
constexpr int kAtTile = 128 * 64 * 2;  // 16 KB: 128 rows x 64 fp16
constexpr int kAtSmemBytes = kAtTile * (1 + 2 * kAtStages) + 1024 /*align*/ + 256 /*barriers*/;

__global__ void __launch_bounds__(kAtThreads, 2)
attention_tcgen05_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnTcParams p) {
  extern __shared__ __align__(1024) uint8_t at_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(at_smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sKV = smem + kAtTile;  // stage s: K at sKV + s*2*kAtTile, V right after it
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kAtTile * (1 + 2 * kAtStages));
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;
  uint64_t* kv_empty = kv_full + kAtStages;
  uint64_t* s_full = kv_empty + kAtStages;
  uint64_t* p_ready = s_full + 1;
  uint64_t* o_done = p_ready + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, head = blockIdx.y, seq = blockIdx.z;
  const int T = p.T;
  const int nkb = (T + 127) >> 7;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmQKV);
    mbar_init(q_full, 1);
    for (int s = 0; s < kAtStages; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
    mbar_init(s_full, 1);
    mbar_init(p_ready, 128);
    mbar_init(o_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base, tmem_O = tmem_base + 128;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, kAtTile);
      tma_load_3d(sQ, &tmQKV, q_full, head * 64, qt * 128, seq);
      for (int j = 0; j < nkb; ++j) {
        const int st = j % kAtStages;
        mbar_wait(&kv_empty[st], ((j / kAtStages) & 1) ^ 1);
        uint8_t* sk = sKV + st * 2 * kAtTile;
        mbar_arrive_expect_tx(&kv_full[st], 2 * kAtTile);
        tma_load_3d(sk, &tmQKV, &kv_full[st], p.d + head * 64, j * 128, seq);
        tma_load_3d(sk + kAtTile, &tmQKV, &kv_full[st], 2 * p.d + head * 64, j * 128, seq);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      mbar_wait(q_full, 0);
      const uint64_t qdesc = make_smem_desc_sw128(smem_u32(sQ), 1024);
      constexpr uint32_t kIdescPV = make_idesc_f16(128, 64, false, true);  // B = V is MN-major
      for (int j = 0; j < nkb; ++j) {
        const int st = j % kAtStages;
        const int rem = T - j * 128;
        const int nk = rem >= 128 ? 128 : ((rem + 15) & ~15);  // keys in this block, rounded to the UMMA N/K step
        mbar_wait(&kv_full[st], (j / kAtStages) & 1);
        tc_fence_after();
        const uint32_t sk = smem_u32(sKV + st * 2 * kAtTile);
        const uint64_t kdesc = make_smem_desc_sw128(sk, 1024);
        const uint32_t idesc_s = make_idesc_f16(128, nk, false, false);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16_ss(tmem_S, qdesc + 2 * k, kdesc + 2 * k, idesc_s, k ? 1u : 0u);
        umma_commit(s_full);
        mbar_wait(p_ready, j & 1);
        tc_fence_after();
        // V_j: rows = keys (the MMA's K), 64 contiguous head-dim values per row (the MMA's N): MN-major, 128B swizzle;
        // a K=16 step is two 8-row swizzle atoms = 2048 B.
        const uint64_t vdesc = make_smem_desc_sw128(sk + kAtTile, 1024, 1024);
        for (int k = 0; k < nk / 16; ++k)
          umma_f16_ts(tmem_O, tmem_S + 8 * k, vdesc + static_cast<uint64_t>(k) * (2048 >> 4), kIdescPV,
                      (j | k) ? 1u : 0u);
        umma_commit(&kv_empty[st]);
        umma_commit(o_done);
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax / correction / output
    const int quad = warp & 3;
    const int row = qt * 128 + quad * 32 + lane;          // query index inside the sequence
    const bool warp_live = qt * 128 + quad * 32 < T;      // warp-uniform: any valid query row in this warp
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    constexpr float kLog2e = 1.4426950408889634f;
    float m = -INFINITY, l = 0.f;                         // running max (log2 domain) and sum
    for (int j = 0; j < nkb; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      if (warp_live) {
        const int rem = T - j * 128;                      // valid keys in this block (>= 1)
        const int nch = rem >= 128 ? 4 : (rem + 31) >> 5; // 32-column chunks holding valid keys
        float mx = m;
        for (int c = 0; c < nch; ++c) {
          uint32_t r[32];
          tmem_ld32(tmem_S + lane_off + c * 32, r);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float s = (c * 32 + i < rem) ? __uint_as_float(r[i]) * kLog2e : -INFINITY;
            mx = fmaxf(mx, s);
          }
        }
        const float alpha = exp2f(m - mx);                // 0 on the first block (m = -inf)
        if (j > 0 && __any_sync(0xffffffffu, alpha != 1.0f)) {
#pragma unroll
          for (int hlf = 0; hlf < 2; ++hlf) {
            uint32_t o[32];
            tmem_ld32(tmem_O + lane_off + hlf * 32, o);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st32(tmem_O + lane_off + hlf * 32, o);
          }
        }
        l *= alpha;
        m = mx;
        const int nst = rem >= 128 ? 4 : (((rem + 15) & ~15) + 31) >> 5;  // chunks covering the MMA's K extent
        for (int c = 0; c < nst; ++c) {
          uint32_t r[32], pk[16];
          tmem_ld32(tmem_S + lane_off + c * 32, r);
          tmem_wait_ld();
          float sum = 0.f;
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float p0 = (c * 32 + i < rem) ? exp2f(fmaf(__uint_as_float(r[i]), kLog2e, -mx)) : 0.f;
            const float p1 = (c * 32 + i + 1 < rem) ? exp2f(fmaf(__uint_as_float(r[i + 1]), kLog2e, -mx)) : 0.f;
            sum += p0 + p1;
            __half2 h = __floats2half2_rn(p0, p1);
            pk[i >> 1] = *reinterpret_cast<uint32_t*>(&h);
          }
          l += sum;
          tmem_st16(tmem_S + lane_off + c * 16, pk);      // P (fp16 x2 per column) over the consumed part of S
        }
        tmem_wait_st();
      }
      tc_fence_before();
      mbar_arrive(p_ready);
    }
    mbar_wait(o_done, (nkb - 1) & 1);
    tc_fence_after();
    if (warp_live) {
      const float inv = 1.0f / l;
      __half* out = p.ctx + (static_cast<long long>(seq) * T + row) * p.ldc + head * 64;
#pragma unroll
      for (int hlf = 0; hlf < 2; ++hlf) {
        uint32_t o[32];
        tmem_ld32(tmem_O + lane_off + hlf * 32, o);
        tmem_wait_ld();
        if (row < T) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 w;
            __half2 h0 = __floats2half2_rn(__uint_as_float(o[8 * q]) * inv, __uint_as_float(o[8 * q + 1]) * inv);
            __half2 h1 = __floats2half2_rn(__uint_as_float(o[8 * q + 2]) * inv, __uint_as_float(o[8 * q + 3]) * inv);
            __half2 h2 = __floats2half2_rn(__uint_as_float(o[8 * q + 4]) * inv, __uint_as_float(o[8 * q + 5]) * inv);
            __half2 h3 = __floats2half2_rn(__uint_as_float(o[8 * q + 6]) * inv, __uint_as_float(o[8 * q + 7]) * inv);
            w.x = *reinterpret_cast<uint32_t*>(&h0); w.y = *reinterpret_cast<uint32_t*>(&h1);
            w.z = *reinterpret_cast<uint32_t*>(&h2); w.w = *reinterpret_cast<uint32_t*>(&h3);
            *reinterpret_cast<uint4*>(out + hlf * 32 + q * 8) = w;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

}  // namespace pg

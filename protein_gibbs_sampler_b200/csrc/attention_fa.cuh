// Self-attention on tcgen05 for head_dim 64 (ESM-1b / ESM-2 650M), second generation: a persistent,
// warp-specialised kernel that keeps TWO 128-query tiles of one (sequence, head) in flight so that the tensor
// core works on one tile while the other tile's softmax runs (the softmax is MUFU-bound: 16 ex2 / clk / SM).
//
//   ctx[s, i, h, :] = softmax_j( q[s,i,h,:] . k[s,j,h,:] ) v[s,j,h,:]     q pre-scaled by Dh^-1/2 (+RoPE) by the
//                                                                          QKV GEMM epilogue; no padding mask
// Work item = (sequence, head, pair of query tiles).  One CTA per SM walks items  blockIdx.x + i*gridDim.x.
// Roles (384 threads):
//   warp 0      TMA producer: Q tiles (double-buffered across items) and 128-key K/V blocks (4-stage ring),
//               3-D tensor map over the fused qkv activation -- rows t >= T are zero-filled
//   warp 1      MMA issuer:  S_t = Q_t K_j^T (SS) into TMEM, O_t += P_t V_j (A = P from TMEM, B = V MN-major)
//   warp 2      TMEM allocator (512 columns: S_A, S_B 128 each; O_A, O_B 80 each)
//   warps 4-7   softmax of tile A: one thread per query row, S read ONCE from TMEM (two 64-column halves),
//   warps 8-11  softmax of tile B  P = 2^(s - m) written over S as packed fp16; O is rescaled in TMEM only when
//               the reference maximum m grows by more than 2^8 (lazy rescale: P <= 256 in fp16);
//               final O / l -> fp16 -> swizzled staging (the tile's own Q buffer) -> one TMA store
// The row sum l is computed by the tensor core: V's MN-major operand is given a second 64-wide panel (LBO) that
// points at a constant all-ones tile, so PV runs with N = 80 and O[:, 64] = sum_j P_j of the ROUNDED fp16 P --
// the normalised weights are then an exact convex combination (a stale reference maximum would otherwise leave
// the rounding error of the dominant P in the output), and the softmax warps need no adds for the sum.
// Issue order per item: S_A0 S_B0 | PV_A0 S_A1 | PV_B0 S_B1 | PV_A1 S_A2 | ...  tcgen05.mma executes in issue
// order, so s_full(t, j+1) also tells tile t's softmax warps that PV(t, j) has retired (O is quiescent).
// Replaces fair-esm MultiheadAttention's bmm / softmax / bmm (call site /root/reference/src/pgen/esm_sampler.py:223).
#pragma once
#include <type_traits>

#include "ptx.cuh"

namespace pg {

struct AttnFaParams {
  int T;        // tokens per sequence (keys: always all T)
  int H;        // heads (head_dim 64)
  int n_seq;
  int n_tiles;  // 128-row query tiles handled here: ceil(T/128), or floor(T/128) when a tail kernel takes the rest
};

constexpr int kFaThreads = 384;
constexpr int kFaKvStages = 4;
constexpr int kFaTile = 128 * 64 * 2;  // 16 KB: 128 rows x 64 fp16
constexpr int kFaSmemBytes = kFaTile * (4 + 2 * kFaKvStages + 1) + 1024 /*align*/ + 256 /*barriers*/;
constexpr int kFaOCols = 80;           // 64 head dims + 16 copies of the row sum
constexpr float kFaRescaleThreshold = 8.0f;  // log2 units

__device__ __forceinline__ float fa_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void fa_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

// Multiply the packed-fp16 P columns [0, ncols) of this thread's TMEM lane by alpha (rare slow path).
__device__ __forceinline__ void fa_rescale_p(uint32_t taddr, float alpha) {
  uint32_t r[32];
  tmem_ld32(taddr, r);
  tmem_wait_ld();
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const float2 f = __half22float2(*reinterpret_cast<__half2*>(&r[i]));
    __half2 h = __floats2half2_rn(f.x * alpha, f.y * alpha);
    r[i] = *reinterpret_cast<uint32_t*>(&h);
  }
  tmem_st32(taddr, r);
}
__device__ __forceinline__ void fa_rescale_o(uint32_t taddr, float alpha) {
#pragma unroll
  for (int hlf = 0; hlf < 2; ++hlf) {
    uint32_t o[32];
    tmem_ld32(taddr + hlf * 32, o);
    tmem_wait_ld();
#pragma unroll
    for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
    tmem_st32(taddr + hlf * 32, o);
  }
  uint32_t o[16];
  tmem_ld16(taddr + 64, o);  // the row sums
  tmem_wait_ld();
#pragma unroll
  for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
  tmem_st16(taddr + 64, o);
}

__global__ void __launch_bounds__(kFaThreads, 1)
attention_fa_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmCtx,
                    const AttnFaParams p) {
  extern __shared__ __align__(1024) uint8_t fa_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(fa_smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                   // [stage 2][tile 2][16 KB]
  uint8_t* sKV = smem + 4 * kFaTile;    // [kFaKvStages][K 16 KB | V 16 KB]
  uint8_t* sOnes = smem + kFaTile * (4 + 2 * kFaKvStages);  // 128 x 128 B of fp16 1.0 (any swizzle reads ones)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kFaTile * (4 + 2 * kFaKvStages + 1));
  uint64_t* q_full = bars;                      // [2]  producer -> MMA (tx bytes)
  uint64_t* q_empty = q_full + 2;               // [2]  softmax groups -> producer (count 2)
  uint64_t* kv_full = q_empty + 2;              // [stages]
  uint64_t* kv_empty = kv_full + kFaKvStages;   // [stages] tcgen05.commit
  uint64_t* s_full = kv_empty + kFaKvStages;    // [2]  per tile: S ready
  uint64_t* p_ready = s_full + 2;               // [2]  per tile: P written (count 4 = warps)
  uint64_t* o_full = p_ready + 2;               // [2]  per tile: last PV of the item retired
  uint64_t* o_free = o_full + 2;                // [2]  per tile: O read out (count 4)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_free + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = p.T;
  const int nkb = (T + 127) >> 7;
  const int n_pairs = (p.n_tiles + 1) >> 1;
  const int n_sh = p.n_seq * p.H;
  const int n_items = n_sh * n_pairs;
  const int d = p.H * 64;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmQKV);
    prefetch_tmap(&tmCtx);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&q_full[s], 1);
      mbar_init(&q_empty[s], 2);
      mbar_init(&s_full[s], 1);
      mbar_init(&p_ready[s], 4);
      mbar_init(&o_full[s], 1);
      mbar_init(&o_free[s], 4);
    }
    for (int s = 0; s < kFaKvStages; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < kFaTile / 16; i += kFaThreads)
    reinterpret_cast<uint4*>(sOnes)[i] = make_uint4(0x3C003C00u, 0x3C003C00u, 0x3C003C00u, 0x3C003C00u);
  fence_proxy_async();  // the MMA reads shared memory through the async proxy
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // item -> (pair, seq, head): pair-major so that every CTA gets the same mix of full and partial pairs
  auto decode = [&](int item, int& pair, int& seq, int& head) {
    pair = item / n_sh;
    const int r = item - pair * n_sh;
    seq = r / p.H;
    head = r - seq * p.H;
  };

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t kv_cnt = 0, it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        int pair, seq, head;
        decode(item, pair, seq, head);
        const bool has_b = 2 * pair + 1 < p.n_tiles;
        const int qs = it & 1;
        mbar_wait(&q_empty[qs], ((it >> 1) & 1) ^ 1);
        uint8_t* q = sQ + qs * 2 * kFaTile;
        mbar_arrive_expect_tx(&q_full[qs], has_b ? 2 * kFaTile : kFaTile);
        tma_load_3d(q, &tmQKV, &q_full[qs], head * 64, 2 * pair * 128, seq);
        if (has_b) tma_load_3d(q + kFaTile, &tmQKV, &q_full[qs], head * 64, (2 * pair + 1) * 128, seq);
        for (int j = 0; j < nkb; ++j, ++kv_cnt) {
          const int st = kv_cnt % kFaKvStages;
          mbar_wait(&kv_empty[st], ((kv_cnt / kFaKvStages) & 1) ^ 1);
          uint8_t* sk = sKV + st * 2 * kFaTile;
          mbar_arrive_expect_tx(&kv_full[st], 2 * kFaTile);
          tma_load_3d(sk, &tmQKV, &kv_full[st], d + head * 64, j * 128, seq);
          tma_load_3d(sk + kFaTile, &tmQKV, &kv_full[st], 2 * d + head * 64, j * 128, seq);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t kIdescPV = make_idesc_f16(128, kFaOCols, false, true);  // B = [V | ones] is MN-major
      uint32_t kv_cnt = 0, it = 0;
      uint32_t p_cnt[2] = {0, 0}, o_cnt[2] = {0, 0};
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        int pair, seq, head;
        decode(item, pair, seq, head);
        const int nt = (2 * pair + 1 < p.n_tiles) ? 2 : 1;
        const int qs = it & 1;
        mbar_wait(&q_full[qs], (it >> 1) & 1);
        tc_fence_after();
        uint64_t qdesc[2];
        qdesc[0] = make_smem_desc_sw128(smem_u32(sQ + qs * 2 * kFaTile), 1024);
        qdesc[1] = make_smem_desc_sw128(smem_u32(sQ + qs * 2 * kFaTile + kFaTile), 1024);
        auto issue_s = [&](int t, int j) {
          const int st = (kv_cnt + j) % kFaKvStages;
          const int rem = T - j * 128;
          const int nk = rem >= 128 ? 128 : ((rem + 15) & ~15);
          const uint64_t kdesc = make_smem_desc_sw128(smem_u32(sKV + st * 2 * kFaTile), 1024);
          const uint32_t idesc = make_idesc_f16(128, nk, false, false);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16_ss(tmem_base + t * 128, qdesc[t] + 2 * k, kdesc + 2 * k, idesc, k ? 1u : 0u);
          umma_commit(&s_full[t]);
        };
        {
          const int st = kv_cnt % kFaKvStages;
          mbar_wait(&kv_full[st], (kv_cnt / kFaKvStages) & 1);
          tc_fence_after();
          for (int t = 0; t < nt; ++t) issue_s(t, 0);
        }
        for (int j = 0; j < nkb; ++j) {
          const int st = (kv_cnt + j) % kFaKvStages;
          const int rem = T - j * 128;
          const int nk = rem >= 128 ? 128 : ((rem + 15) & ~15);
          // V_j: rows = keys (the MMA's K), 64 contiguous head-dim values per row (the MMA's N): MN-major, 128B
          // swizzle; a K=16 step is two 8-row swizzle atoms = 2048 B.
          // LBO = distance to the next 64-wide N panel = the all-ones tile (columns 64..79 of the operand).
          const uint32_t sv = smem_u32(sKV + st * 2 * kFaTile + kFaTile);
          const uint64_t vdesc = make_smem_desc_sw128(sv, 1024, smem_u32(sOnes) - sv);
          for (int t = 0; t < nt; ++t) {
            mbar_wait(&p_ready[t], p_cnt[t] & 1);
            ++p_cnt[t];
            if (j == 0) mbar_wait(&o_free[t], (o_cnt[t] & 1) ^ 1);  // previous item's O has been read out
            tc_fence_after();
            const uint32_t tS = tmem_base + t * 128, tO = tmem_base + 256 + t * kFaOCols;
            for (int k = 0; k < nk / 16; ++k)
              umma_f16_ts(tO, tS + 8 * k, vdesc + static_cast<uint64_t>(k) * (2048 >> 4), kIdescPV, (j | k) ? 1u : 0u);
            if (j + 1 < nkb) {
              if (t == 0) {
                const uint32_t c = kv_cnt + j + 1;
                mbar_wait(&kv_full[c % kFaKvStages], (c / kFaKvStages) & 1);
                tc_fence_after();
              }
              issue_s(t, j + 1);
            } else {
              umma_commit(&o_full[t]);
              ++o_cnt[t];
            }
          }
          umma_commit(&kv_empty[st]);  // K_j / V_j fully consumed once everything issued so far retires
        }
        kv_cnt += nkb;
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ softmax / correction / output
    const int t = (warp - 4) >> 2;     // tile slot: 0 = A, 1 = B
    const int quad = warp & 3;         // TMEM lane quadrant
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t tS = tmem_base + t * 128 + lane_off, tO = tmem_base + 256 + t * kFaOCols + lane_off;
    const bool leader = (threadIdx.x & 127) == 0;
    constexpr float kLog2e = 1.4426950408889634f;
    uint32_t s_cnt = 0, o_cnt = 0, it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      int pair, seq, head;
      decode(item, pair, seq, head);
      const int qs = it & 1;
      const int tile = 2 * pair + t;
      if (tile >= p.n_tiles) {  // no second tile in this item: only keep the Q-stage handshake going
        if (leader) {
          // paced by the producer (q_full of THIS item), or two early arrivals could complete one phase
          mbar_wait(&q_full[qs], (it >> 1) & 1);
          mbar_arrive(&q_empty[qs]);
        }
        continue;
      }
      const bool warp_live = tile * 128 + quad * 32 < T;  // warp-uniform: any valid query row in this warp
      float m_used = -INFINITY;                           // reference maximum (log2 domain)
      for (int j = 0; j < nkb; ++j) {
        mbar_wait(&s_full[t], s_cnt & 1);
        ++s_cnt;
        tc_fence_after();
        if (warp_live) {
          const int rem = T - j * 128;                    // valid keys in this block (>= 1)
          const int nk = rem >= 128 ? 128 : ((rem + 15) & ~15);
          // `W` consecutive score columns starting at `c0` (a multiple of 64): reference update, P = 2^(s - m)
          auto chunk = [&](auto wtag, int c0) {
            constexpr int W = decltype(wtag)::value;
            float v[W];
#pragma unroll
            for (int g = 0; g < W / 32; ++g) {
              uint32_t r[32];
              tmem_ld32(tS + c0 + g * 32, r);
              tmem_wait_ld();
#pragma unroll
              for (int i = 0; i < 32; ++i) v[g * 32 + i] = __uint_as_float(r[i]);
            }
            if (rem < c0 + W) {                           // block edge: keys >= T do not exist
#pragma unroll
              for (int i = 0; i < W; ++i)
                if (c0 + i >= rem) v[i] = -INFINITY;
            }
            float mx = v[0];
#pragma unroll
            for (int i = 1; i < W; ++i) mx = fmaxf(mx, v[i]);
            mx *= kLog2e;
            if (__any_sync(0xffffffffu, mx > m_used + kFaRescaleThreshold)) {
              const float m_new = fmaxf(m_used, mx);
              const float alpha = fa_ex2(m_used - m_new);  // 0 on the very first chunk (m_used = -inf)
              if (j > 0) fa_rescale_o(tO, alpha);           // PV(t, j-1) has retired (see header)
              if (c0 > 0) fa_rescale_p(tS, alpha);          // first half of this block used the old reference
              if (j > 0 || c0 > 0) tmem_wait_st();
              m_used = m_new;
            }
            uint32_t pk[W / 2];
#pragma unroll
            for (int i = 0; i < W; i += 2) {
              __half2 h = __floats2half2_rn(fa_ex2(fmaf(v[i], kLog2e, -m_used)), fa_ex2(fmaf(v[i + 1], kLog2e, -m_used)));
              pk[i >> 1] = *reinterpret_cast<uint32_t*>(&h);
            }
            // P (fp16 x2 per column) over the consumed part of S
            if constexpr (W == 64) tmem_st32(tS + (c0 >> 1), pk); else tmem_st16(tS + (c0 >> 1), pk);
          };
          if (nk > 32) chunk(std::integral_constant<int, 64>{}, 0); else chunk(std::integral_constant<int, 32>{}, 0);
          if (nk > 96) chunk(std::integral_constant<int, 64>{}, 64);
          else if (nk > 64) chunk(std::integral_constant<int, 32>{}, 64);
          tmem_wait_st();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_ready[t]);
      }
      // ---- output: O / l -> fp16 -> staging (this tile's Q buffer: every S MMA of the item has retired) -> TMA store
      mbar_wait(&o_full[t], o_cnt & 1);
      ++o_cnt;
      tc_fence_after();
      uint8_t* stage = sQ + (qs * 2 + t) * kFaTile;
      if (warp_live) {
        float inv;
        {
          uint32_t ls[16];
          tmem_ld16(tO + 64, ls);
          tmem_wait_ld();
          inv = 1.0f / __uint_as_float(ls[0]);
        }
        const int r = quad * 32 + lane;
        uint8_t* rowp = stage + r * 128;
#pragma unroll
        for (int hlf = 0; hlf < 2; ++hlf) {
          uint32_t o[32];
          tmem_ld32(tO + hlf * 32, o);
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
            *reinterpret_cast<uint4*>(rowp + (((hlf * 4 + q) ^ (r & 7)) << 4)) = w;  // 128B swizzle, matches tmCtx
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_free[t]);
      fence_proxy_async();            // generic-proxy smem writes -> visible to the TMA (async proxy)
      fa_bar_sync(1 + t, 128);
      if (leader) {
        tma_store_3d(&tmCtx, stage, head * 64, tile * 128, seq);  // rows >= T are clipped by the tensor map
        tma_store_commit();
        tma_store_wait_read<0>();
        mbar_arrive(&q_empty[qs]);
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

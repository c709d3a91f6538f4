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
//   warps 1, 3  MMA issuers (tile A, tile B): S_t = Q_t K_j^T (SS) into TMEM, O_t += P_t V_j (A = P from TMEM,
//               B = V MN-major)
//   warp 2      TMEM allocator (512 columns: S_A[2], S_B[2] 64 each; O_A, O_B 80 each); also the TAIL warp: when
//               T = 128 k + r with 1 <= r <= 8 (T = L + 2 for the usual L) the r trailing query rows would cost a
//               whole 128-row tile (or a second kernel that re-reads every K/V row from HBM); instead this warp
//               computes them with mma.sync straight from the K/V blocks of the ring, transposed (keys are the MMA
//               M dimension, the <= 8 queries its N = 8), once per (sequence, head)
//   warps 4-7   softmax of tile A: one thread per query row, each S sub-block read ONCE from TMEM,
//   warps 8-11  softmax of tile B  P = 2^(s - m) written over S as packed fp16; O is rescaled in TMEM only when
//               the reference maximum m grows by more than 2^11 (lazy rescale: P <= 2048 in fp16);
//               final O / l -> fp16 -> swizzled staging (the tile's own Q buffer) -> one TMA store
// The row sum l is computed by the tensor core: V's MN-major operand is given a second 64-wide panel (LBO) that
// points at a constant all-ones tile, so PV runs with N = 80 and O[:, 64] = sum_j P_j of the ROUNDED fp16 P --
// the normalised weights are then an exact convex combination (a stale reference maximum would otherwise leave
// the rounding error of the dominant P in the output), and the softmax warps need no adds for the sum.
// Scores are produced in 64-key sub-blocks into a per-tile DOUBLE buffer (S_t[2] x 64 columns): the issuer
// queues S(t, i+2) right behind PV(t, i), two sub-blocks ahead of the softmax, so a softmax group never waits for
// the tensor core in steady state and the two groups together keep the MUFU pipe busy.  Issue order per tile:
//   S_0 S_1 | PV_0 S_2 | PV_1 S_3 | PV_2 S_4 | ...     (one thread's tcgen05.mma execute in issue order)
// s_full(t, i) therefore implies PV(t, i-2) has retired; the rare O rescale at sub-block i additionally waits
// for PV(t, i-1) on pv_done[t].
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
  // Optional timeline (debug): CTA 0 appends (clock64 << 8 | event code) words, kFaTraceCap per traced warp
  // (0: issuer A, 1: softmax warp 4, 2: issuer B, 3: softmax warp 8).  nullptr = off; only -DPGIBBS_FA_TRACE=1 builds
  // write it.
  unsigned long long* trace;
  int stagger_cycles;  // initial lag of tile B's softmax group behind tile A's (see the softmax role)
  // Trailing query rows [n_tiles*128, T) (at most 8) handled by warp 2 with mma.sync from the K/V blocks that
  // are in shared memory anyway (0 = none: the tiles cover every row).
  int tail_rows;
  const __half* qkv;  // [n_seq*T, 3*H*64]
  __half* ctx;        // [n_seq*T, H*64]
  int reverse;        // walk the (sequence, head) items from the last one down (see pgibbs_engine::zigzag)
  int ctx_ld;         // ctx row pitch in elements (H*64, or 2*H*64 when rows are [hi | lo])
  int ctx_lo_off;     // split-operand mode: also store the fp16 rounding residual of ctx at column + ctx_lo_off (0 = off)
};
constexpr int kFaTraceCap = 2048;

constexpr int kFaThreads = 384;
constexpr int kFaKvStages = 4;
constexpr int kFaTile = 128 * 64 * 2;  // 16 KB: 128 rows x 64 fp16
constexpr int kFaSmemBytes = kFaTile * (4 + 2 * kFaKvStages + 1) + 1024 /*align*/ + 256 /*barriers*/ + 1024 /*tail P*/;
constexpr int kFaOCols = 80;           // 64 head dims + 16 copies of the row sum
constexpr float kFaRescaleThreshold = 11.0f;  // log2 units: P <= 2^11 (fp16 overflows at 2^16)

__device__ __forceinline__ float fa_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
#ifndef PGIBBS_FA_TRACE
#define PGIBBS_FA_TRACE 0
#endif
// Rare path (the reference maximum of a row grew by more than the threshold): O *= alpha.  Out of line and 16 columns at
// a time, so that it does not raise the register pressure of the softmax loop it is called from.
__device__ __noinline__ void fa_rescale_o(uint32_t taddr, float alpha) {
#pragma unroll 1
  for (int c0 = 0; c0 < kFaOCols; c0 += 16) {
    uint32_t o[16];
    tmem_ld16(taddr + c0, o);
    tmem_wait_ld();
#pragma unroll
    for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
    tmem_st16(taddr + c0, o);
  }
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
  uint64_t* s_full = kv_empty + kFaKvStages;    // [tile 2][buf 2]  S sub-block ready
  uint64_t* p_ready = s_full + 4;               // [tile 2][buf 2]  P written (count 4 = warps)
  uint64_t* pv_done = p_ready + 4;              // [2]  per tile: one completion per PV sub-block
  uint64_t* o_full = pv_done + 2;               // [2]  per tile: last PV of the item retired
  uint64_t* o_free = o_full + 2;                // [2]  per tile: O read out (count 4)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_free + 2);
  __half* sTailP = reinterpret_cast<__half*>(reinterpret_cast<uint8_t*>(bars) + 256);  // [8 queries][64 keys]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = p.T;
  const int nkb = (T + 127) >> 7;   // 128-key K/V blocks (TMA / smem stage granularity)
  const int nsub = (T + 63) >> 6;   // 64-key sub-blocks (MMA / softmax granularity)
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
      mbar_init(&s_full[2 * s], 1);
      mbar_init(&s_full[2 * s + 1], 1);
      mbar_init(&p_ready[2 * s], 4);
      mbar_init(&p_ready[2 * s + 1], 4);
      mbar_init(&pv_done[s], 1);
      mbar_init(&o_full[s], 1);
      mbar_init(&o_free[s], 4);
    }
    // K/V blocks are released by both issuers and, when there are trailing rows, by the tail warp
    for (int s = 0; s < kFaKvStages; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], p.tail_rows > 0 ? 3 : 2); }
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
  pdl_wait();               // set-up above overlaps the previous kernel's tail (programmatic dependent launch)
  pdl_launch_dependents();

  // The timeline costs registers and ~6 instructions per point in the softmax loop even when it is switched off (at the
  // 168-register cap that pushed loop-carried state into local memory, re-loaded in the middle of every hand-over:
  // 12 % of the kernel), so it only exists in -DPGIBBS_FA_TRACE=1 builds (PGIBBS_NVCC_EXTRA; tools/attn_trace.py).
#if PGIBBS_FA_TRACE
  const int trace_slot = warp == 1 ? 0 : warp == 4 ? 1 : warp == 3 ? 2 : warp == 8 ? 3 : -1;
  const bool tracing = p.trace != nullptr && blockIdx.x == 0 && lane == 0 && trace_slot >= 0;
  int trace_n = 0;
  auto trace = [&](int code) {
    if (tracing && trace_n < kFaTraceCap) p.trace[trace_slot * kFaTraceCap + trace_n++] = (static_cast<unsigned long long>(clock64()) << 8) | code;
  };
#else
  auto trace = [](int) {};
#endif

  // item -> (pair, seq, head): pair-major so that every CTA gets the same mix of full and partial pairs
  auto decode = [&](int item, int& pair, int& seq, int& head) {
    pair = item / n_sh;
    int r = item - pair * n_sh;
    if (p.reverse) r = n_sh - 1 - r;
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
        mbar_wait_lean(&q_empty[qs], ((it >> 1) & 1) ^ 1);
        uint8_t* q = sQ + qs * 2 * kFaTile;
        mbar_arrive_expect_tx(&q_full[qs], has_b ? 2 * kFaTile : kFaTile);
        tma_load_3d(q, &tmQKV, &q_full[qs], head * 64, 2 * pair * 128, seq);
        if (has_b) tma_load_3d(q + kFaTile, &tmQKV, &q_full[qs], head * 64, (2 * pair + 1) * 128, seq);
        for (int j = 0; j < nkb; ++j, ++kv_cnt) {
          const int st = kv_cnt % kFaKvStages;
          mbar_wait_lean(&kv_empty[st], ((kv_cnt / kFaKvStages) & 1) ^ 1);
          uint8_t* sk = sKV + st * 2 * kFaTile;
          mbar_arrive_expect_tx(&kv_full[st], 2 * kFaTile);
          tma_load_3d(sk, &tmQKV, &kv_full[st], d + head * 64, j * 128, seq);
          tma_load_3d(sk + kFaTile, &tmQKV, &kv_full[st], 2 * d + head * 64, j * 128, seq);
        }
      }
    }
  } else if (warp == 1 || warp == 3) {
    // ------------------------------------------------------------------ MMA issuers: warp 1 -> tile A, warp 3 -> tile B
    // One issuer per tile: the 64-key MMAs are small (32-40 tensor cycles) and cost the issuing warp ~50-100 cycles
    // each (plus ~100 per commit), so a single issuer serving both tiles in turn was the kernel's critical path.
    // The whole warp walks the schedule (waits, address arithmetic) so that every operand is warp-uniform and
    // lives in uniform registers; one elected lane issues the tcgen05 instructions.  (With a `lane == 0` branch
    // around the loop the compiler wraps EVERY tcgen05.mma in an elect/R2UR waterfall loop, ~10 instructions each.)
    // Descriptors are built once; per MMA only the low word changes (start address, and for V the LBO that points
    // at the all-ones panel).
    const int t = warp >> 1;  // 0 or 1
    const bool issuer = elect_one();
    constexpr uint32_t kIdescPV = make_idesc_f16(128, kFaOCols, false, true);  // B = [V | ones] is MN-major
    constexpr uint32_t kIdescS = make_idesc_f16(128, 64, false, false);
    const int last_keys = T - (nsub - 1) * 64;                                  // 1..64 valid keys
    const int last_nk = last_keys >= 64 ? 64 : ((last_keys + 15) & ~15);       // rounded to the UMMA N / K step
    const uint32_t idesc_s_last = make_idesc_f16(128, last_nk, false, false);
    const uint64_t kd0 = make_smem_desc_sw128(smem_u32(sKV), 1024);
    const uint64_t vd0 = make_smem_desc_sw128(smem_u32(sKV + kFaTile), 1024, smem_u32(sOnes) - smem_u32(sKV + kFaTile));
    const uint32_t desc_hi = static_cast<uint32_t>(kd0 >> 32);                  // same for every descriptor here
    const uint32_t k_lo0 = static_cast<uint32_t>(kd0), v_lo0 = static_cast<uint32_t>(vd0);
    const uint32_t q_lo0 = static_cast<uint32_t>(make_smem_desc_sw128(smem_u32(sQ + t * kFaTile), 1024));
    auto desc = [&](uint32_t lo) { return (static_cast<uint64_t>(desc_hi) << 32) | lo; };
    const uint32_t tS0 = tmem_base + t * 128, tO = tmem_base + 256 + t * kFaOCols;
    uint32_t kv_cnt = 0, it = 0;
    uint32_t p_par = 0, o_par = 0;  // phase parities: bit b of p_ready[t][b], o_free[t] (waits done so far)
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it, kv_cnt += nkb) {
      const int pair = item / n_sh;
      if (2 * pair + t >= p.n_tiles) {
        // No tile for this issuer in this item: it still owes kv_empty its arrival for every K/V block (paced by
        // kv_full, so that the arrival lands in the right phase of the ring).
        for (int j = 0; j < nkb; ++j) {
          const uint32_t c = kv_cnt + j;
          mbar_wait_lean(&kv_full[c % kFaKvStages], (c / kFaKvStages) & 1);
          if (issuer) mbar_arrive(&kv_empty[c % kFaKvStages]);
          __syncwarp();
        }
        continue;
      }
      const int qs = it & 1;
      trace(0x01);  // item begins (issuer)
      mbar_wait_lean(&q_full[qs], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t q_lo = q_lo0 + qs * (2 * kFaTile >> 4);
      // S(i) = Q K_i^T into buffer i & 1.  K rows of sub-block i: stage of block i/2, +8 KB for the odd half.
      auto issue_s = [&](int i) {
        const uint32_t delta = ((kv_cnt + (i >> 1)) % kFaKvStages) * (2 * kFaTile >> 4) + (i & 1) * (kFaTile >> 5);
        const uint32_t kd = k_lo0 + delta;
        const uint32_t idesc = (i + 1 == nsub) ? idesc_s_last : kIdescS;
        const uint32_t tS = tS0 + (i & 1) * 64;
        if (issuer) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16_ss(tS, desc(q_lo + 2 * k), desc(kd + 2 * k), idesc, k ? 1u : 0u);
          umma_commit(&s_full[2 * t + (i & 1)]);
        }
        __syncwarp();
      };
      mbar_wait_lean(&kv_full[kv_cnt % kFaKvStages], (kv_cnt / kFaKvStages) & 1);
      tc_fence_after();
      issue_s(0);
      if (nsub > 1) issue_s(1);
      for (int i = 0; i < nsub; ++i) {
        const int b = i & 1;
        const int st = (kv_cnt + (i >> 1)) % kFaKvStages;
        const int ksteps = (i + 1 == nsub) ? last_nk >> 4 : 4;
        // V rows of sub-block i: keys are the MMA's K, 64 contiguous head-dim values per key row (the MMA's N):
        // MN-major, 128B swizzle; a K=16 step is two 8-row swizzle atoms = 2048 B.  The LBO field (bits 16..29)
        // is the distance to the next 64-wide N panel = the all-ones tile (columns 64..79 of the operand), so
        // moving the start address by `delta` moves the LBO by -delta.
        const uint32_t delta = st * (2 * kFaTile >> 4) + b * (kFaTile >> 5);
        const uint32_t vd = v_lo0 + delta - (delta << 16);
        mbar_wait_lean(&p_ready[2 * t + b], (p_par >> b) & 1);
        p_par ^= 1u << b;
        trace(0x10 + t);  // P(i) seen
        if (i == 0) mbar_wait_lean(&o_free[t], (o_par & 1) ^ 1);  // previous item's O has been read out
        tc_fence_after();
        trace(0x30);  // fences done
        if (issuer) {
          for (int k = 0; k < ksteps; ++k)
            umma_f16_ts(tO, tS0 + b * 64 + 8 * k, desc(vd + k * (2048 >> 4)), kIdescPV, (i | k) ? 1u : 0u);
        }
        __syncwarp();
        trace(0x31);  // PV MMAs issued
        if (issuer) umma_commit(&pv_done[t]);
        __syncwarp();
        trace(0x32);  // pv_done commit issued
        if (i + 2 < nsub) {
          if (b == 0) {  // sub-block i+2 opens the next 128-key block
            const uint32_t c = kv_cnt + (i >> 1) + 1;
            mbar_wait_lean(&kv_full[c % kFaKvStages], (c / kFaKvStages) & 1);
            tc_fence_after();
          }
          issue_s(i + 2);
        }
        if (i + 1 == nsub) {
          if (issuer) umma_commit(&o_full[t]);
          o_par ^= 1u;
        }
        // this tile is done with K/V block i/2 (kv_empty counts both issuers) once everything issued so far retires
        if ((b || i + 1 == nsub) && issuer) umma_commit(&kv_empty[st]);
        __syncwarp();
        trace(0x18 + t);  // PV(i) [+ S(i+2)] issued
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------------ trailing query rows (see header)
    //   S^T[key, q] = K[key, :] . Q[q, :]        A = K rows (ldmatrix from the 128B-swizzled tile), B = Q from global
    //   O^T[dim, q] = sum_key V[key, dim] P^T[key, q]   A = V^T (ldmatrix.trans), B = P^T via a 1 KB smem transpose
    // C-fragment: thread (g = lane/4, t4 = lane%4) holds rows (g, g+8) x queries (2 t4, 2 t4 + 1).
    if (p.tail_rows > 0) {
      const int g = lane >> 2, t4 = lane & 3;
      const int q_first = p.n_tiles * 128;
      constexpr float kLog2e = 1.4426950408889634f;
      uint32_t kv_cnt = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, kv_cnt += nkb) {
        int pair, seq, head;
        decode(item, pair, seq, head);
        if (pair != 0) {  // the tail of this (sequence, head) belongs to its first item; still release the blocks
          for (int j = 0; j < nkb; ++j) {
            const uint32_t c = kv_cnt + j;
            mbar_wait_lean(&kv_full[c % kFaKvStages], (c / kFaKvStages) & 1);
            if (lane == 0) mbar_arrive(&kv_empty[c % kFaKvStages]);
            __syncwarp();
          }
          continue;
        }
        // Q as the B operand: b0 = Q[q = g][16 ks + 2 t4, +1], b1 = Q[g][16 ks + 2 t4 + 8, +9]; rows >= tail are 0
        uint32_t qb[4][2];
        {
          const __half* qrow = p.qkv + (static_cast<long long>(seq) * T + q_first + g) * (3 * d) + head * 64;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            qb[ks][0] = g < p.tail_rows ? *reinterpret_cast<const uint32_t*>(qrow + ks * 16 + 2 * t4) : 0u;
            qb[ks][1] = g < p.tail_rows ? *reinterpret_cast<const uint32_t*>(qrow + ks * 16 + 2 * t4 + 8) : 0u;
          }
        }
        float o[4][4];  // O^T: 4 tiles of 16 dims; c0,c2 <-> query 2 t4, c1,c3 <-> query 2 t4 + 1
#pragma unroll
        for (int j = 0; j < 4; ++j) { o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f; }
        float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;  // queries 2 t4, 2 t4 + 1
        for (int j = 0; j < nkb; ++j) {
          const uint32_t c = kv_cnt + j;
          const int st = c % kFaKvStages;
          mbar_wait_lean(&kv_full[st], (c / kFaKvStages) & 1);
          const uint32_t sk = smem_u32(sKV + st * 2 * kFaTile), sv = sk + kFaTile;
          const int nsb = min(2, (T - j * 128 + 63) >> 6);
          for (int sb = 0; sb < nsb; ++sb) {
            const int key0 = sb * 64, rem = T - j * 128 - key0;  // valid keys in this 64-key sub-block (>= 1)
            float sc[4][4];  // S^T: 4 tiles of 16 keys
#pragma unroll
            for (int mt = 0; mt < 4; ++mt) {
              sc[mt][0] = sc[mt][1] = sc[mt][2] = sc[mt][3] = 0.f;
              const int key = key0 + mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                uint32_t a[4];
                const uint32_t addr = sk + key * 128 + (((ks * 2 + (lane >> 4)) ^ (key & 7)) << 4);
                asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                             : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]) : "r"(addr));
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(sc[mt][0]), "+f"(sc[mt][1]), "+f"(sc[mt][2]), "+f"(sc[mt][3])
                             : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(qb[ks][0]), "r"(qb[ks][1]));
              }
            }
            // keys >= T do not exist; running maximum per query over the keys (registers, then the 8 lane groups)
            float mx0 = m0, mx1 = m1;
#pragma unroll
            for (int mt = 0; mt < 4; ++mt) {
              if (mt * 16 + g >= rem) { sc[mt][0] = -INFINITY; sc[mt][1] = -INFINITY; }
              if (mt * 16 + g + 8 >= rem) { sc[mt][2] = -INFINITY; sc[mt][3] = -INFINITY; }
              mx0 = fmaxf(mx0, fmaxf(sc[mt][0], sc[mt][2]));
              mx1 = fmaxf(mx1, fmaxf(sc[mt][1], sc[mt][3]));
            }
#pragma unroll
            for (int off = 4; off < 32; off <<= 1) {
              mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, off));
              mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, off));
            }
            const float a0 = exp2f((m0 - mx0) * kLog2e), a1 = exp2f((m1 - mx1) * kLog2e);  // 0 on the first sub-block
            m0 = mx0; m1 = mx1;
            float r0 = 0.f, r1 = 0.f;
            __syncwarp();  // the previous sub-block's P^T has been consumed by every lane
#pragma unroll
            for (int mt = 0; mt < 4; ++mt) {
              const float p00 = exp2f((sc[mt][0] - mx0) * kLog2e), p01 = exp2f((sc[mt][1] - mx1) * kLog2e);
              const float p10 = exp2f((sc[mt][2] - mx0) * kLog2e), p11 = exp2f((sc[mt][3] - mx1) * kLog2e);
              r0 += p00 + p10; r1 += p01 + p11;
              sTailP[(2 * t4) * 64 + mt * 16 + g] = __float2half_rn(p00);
              sTailP[(2 * t4 + 1) * 64 + mt * 16 + g] = __float2half_rn(p01);
              sTailP[(2 * t4) * 64 + mt * 16 + g + 8] = __float2half_rn(p10);
              sTailP[(2 * t4 + 1) * 64 + mt * 16 + g + 8] = __float2half_rn(p11);
            }
#pragma unroll
            for (int off = 4; off < 32; off <<= 1) {
              r0 += __shfl_xor_sync(0xffffffffu, r0, off);
              r1 += __shfl_xor_sync(0xffffffffu, r1, off);
            }
            l0 = l0 * a0 + r0; l1 = l1 * a1 + r1;
#pragma unroll
            for (int dt = 0; dt < 4; ++dt) { o[dt][0] *= a0; o[dt][2] *= a0; o[dt][1] *= a1; o[dt][3] *= a1; }
            __syncwarp();
            // O^T += V^T P^T.  B operand from the transpose buffer: b0 = P[q = g][16 ks + 2 t4, +1], b1 = .. + 8
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint32_t b0 = *reinterpret_cast<const uint32_t*>(&sTailP[g * 64 + ks * 16 + 2 * t4]);
              const uint32_t b1 = *reinterpret_cast<const uint32_t*>(&sTailP[g * 64 + ks * 16 + 2 * t4 + 8]);
              const int key = key0 + ks * 16 + (lane & 7) + ((lane >> 4) & 1) * 8;
#pragma unroll
              for (int dt = 0; dt < 4; ++dt) {
                uint32_t a[4];
                const uint32_t addr = sv + key * 128 + (((dt * 2 + ((lane >> 3) & 1)) ^ (key & 7)) << 4);
                asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                             : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]) : "r"(addr));
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(o[dt][0]), "+f"(o[dt][1]), "+f"(o[dt][2]), "+f"(o[dt][3])
                             : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
              }
            }
          }
          __syncwarp();  // every lane has finished reading this K/V block
          if (lane == 0) mbar_arrive(&kv_empty[st]);
        }
        // ctx[q][dim] = O^T[dim][q] / l[q]
        const float i0 = 1.0f / l0, i1 = 1.0f / l1;
        const long long ldc = p.ctx_ld;
        __half* out = p.ctx + (static_cast<long long>(seq) * T + q_first) * ldc + head * 64;
        auto put = [&](int q, int col, float v) {
          const __half hi = __float2half_rn(v);
          out[q * ldc + col] = hi;
          if (p.ctx_lo_off) out[q * ldc + p.ctx_lo_off + col] = __float2half_rn(v - __half2float(hi));
        };
#pragma unroll
        for (int dt = 0; dt < 4; ++dt) {
          if (2 * t4 < p.tail_rows) {
            put(2 * t4, dt * 16 + g, o[dt][0] * i0);
            put(2 * t4, dt * 16 + g + 8, o[dt][2] * i0);
          }
          if (2 * t4 + 1 < p.tail_rows) {
            put(2 * t4 + 1, dt * 16 + g, o[dt][1] * i1);
            put(2 * t4 + 1, dt * 16 + g + 8, o[dt][3] * i1);
          }
        }
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
    uint32_t s_par = 0, o_cnt = 0, pv_cnt = 0, it = 0;
    // The leader hands a Q/O staging buffer back to the producer once the TMA store has finished READING it.
    // Waiting for that right after issuing the store would stall the leader's warp ~1000 cycles per item (and the
    // whole group with it, through p_ready); it is done one sub-block into the next item instead, when it is free.
    // The two groups share each SM sub-partition's MUFU pipe.  A sub-block is ~650 cycles of non-MUFU work (TMEM
    // load, row maximum, TMEM store, hand-over) followed by ~512 cycles of ex2; started together, both groups sit
    // in the same phase and a sub-block pair costs 650 + 2*512 cycles, started half a period apart one group's ex2
    // phase hides the other's overhead (650 + 512).  The lag is neutral-stable (both groups do identical work), so
    // it is set once here.
    if (t == 1 && p.stagger_cycles > 0) {
      const long long t0 = clock64();
      while (clock64() - t0 < p.stagger_cycles) { }
    }
    int pending_qs = -1;
    auto release_stage = [&]() {
      if (pending_qs >= 0) {
        tma_store_wait_read<0>();
        mbar_arrive(&q_empty[pending_qs]);
        pending_qs = -1;
      }
    };
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      int pair, seq, head;
      decode(item, pair, seq, head);
      const int qs = it & 1;
      const int tile = 2 * pair + t;
      if (tile >= p.n_tiles) {  // no second tile in this item: only keep the Q-stage handshake going
        if (leader) {
          release_stage();
          // paced by the producer (q_full of THIS item), or two early arrivals could complete one phase
          mbar_wait_lean(&q_full[qs], (it >> 1) & 1);
          mbar_arrive(&q_empty[qs]);
        }
        continue;
      }
      const bool warp_live = tile * 128 + quad * 32 < T;  // warp-uniform: any valid query row in this warp
      float m_used = -INFINITY;                           // reference maximum (log2 domain)
      if (!warp_live) {  // no valid query row in this warp: only keep the hand-shakes going (off the common path's loop)
        for (int i = 0; i < nsub; ++i) {
          const int b = i & 1;
          mbar_wait_lean(&s_full[2 * t + b], (s_par >> b) & 1);
          s_par ^= 1u << b;
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&p_ready[2 * t + b]);
        }
      } else
      for (int i = 0; i < nsub; ++i) {
        const int b = i & 1;
        trace(0x20);  // waiting for S(i)
        mbar_wait_lean(&s_full[2 * t + b], (s_par >> b) & 1);
        s_par ^= 1u << b;
        tc_fence_after();
        trace(0x21);  // S(i) ready
        {
          const int rem = T - i * 64;                     // valid keys in this sub-block (>= 1)
          const uint32_t tSb = tS + b * 64;
          // `W` score columns: reference update (lazy), P = 2^(s - m) as packed fp16 over the consumed scores
          auto chunk = [&](auto wtag, auto etag) {
            constexpr int W = decltype(wtag)::value;
            constexpr bool kEdge = decltype(etag)::value;   // the last sub-block of the sequence: keys beyond T are masked
            float v[W];
            {
              uint32_t r[W];  // both 32-column loads in flight before the single wait
#pragma unroll
              for (int g = 0; g < W / 32; ++g) tmem_ld32(tSb + g * 32, reinterpret_cast<uint32_t(&)[32]>(r[g * 32]));
              tmem_wait_ld();
#pragma unroll
              for (int c = 0; c < W; ++c) v[c] = __uint_as_float(r[c]);
            }
            if (kEdge && rem < W) {                       // sequence edge: keys >= T do not exist
#pragma unroll
              for (int c = 0; c < W; ++c)
                if (c >= rem) v[c] = -INFINITY;
            }
            float mx4[4] = {v[0], v[1], v[2], v[3]};
#pragma unroll
            for (int c = 4; c < W; ++c) mx4[c & 3] = fmaxf(mx4[c & 3], v[c]);
            const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])) * kLog2e;
            if (__any_sync(0xffffffffu, mx > m_used + kFaRescaleThreshold)) {
              const float m_new = fmaxf(m_used, mx);
              if (i > 0) {
                // O is quiescent once PV(t, i-1) -- completion number pv_cnt + i of pv_done[t] -- has retired
                mbar_wait_lean(&pv_done[t], (pv_cnt + i - 1) & 1);
                tc_fence_after();
                fa_rescale_o(tO, fa_ex2(m_used - m_new));
                tmem_wait_st();
              }
              m_used = m_new;
            }
            uint32_t pk[W / 2];
            {
              // x = s log2e - m on packed fp32 pairs (fma.rn.f32x2 -> FFMA2): the FMA pipe spends the same time per result,
              // but a pair costs one issue slot, and the softmax warps' issue slots are the kernel's time (-2.4 / -3.3 /
              // -3.5 % at T = 258 / 514 / 1024)
              uint64_t sc, nm;
              asm("mov.b64 %0, {%1, %1};" : "=l"(sc) : "f"(kLog2e));
              asm("mov.b64 %0, {%1, %1};" : "=l"(nm) : "f"(-m_used));
#pragma unroll
              for (int c = 0; c < W; c += 2) {
                uint64_t x;
                asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(v[c]), "f"(v[c + 1]));
                asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x) : "l"(sc), "l"(nm));
                float x0, x1;
                asm("mov.b64 {%0, %1}, %2;" : "=f"(x0), "=f"(x1) : "l"(x));
                pk[c >> 1] = f2h2_rn(fa_ex2(x0), fa_ex2(x1));
              }
            }
            if constexpr (W == 64) tmem_st32(tSb, pk); else tmem_st16(tSb, pk);
          };
          // every sub-block but the last is full: no edge test, no masking code, no width choice on the common path
          if (i + 1 < nsub) chunk(std::integral_constant<int, 64>{}, std::false_type{});
          else if (rem > 32) chunk(std::integral_constant<int, 64>{}, std::true_type{});
          else chunk(std::integral_constant<int, 32>{}, std::true_type{});
          tmem_wait_st();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_ready[2 * t + b]);
        trace(0x22);  // P(i) handed over
        if (leader && i == 0) release_stage();
      }
      pv_cnt += nsub;
      // ---- output: O / l -> fp16 -> staging (this tile's Q buffer: every S MMA of the item has retired) -> TMA store
      mbar_wait_lean(&o_full[t], o_cnt & 1);
      ++o_cnt;
      tc_fence_after();
      trace(0x23);  // O complete
      uint8_t* stage = sQ + (qs * 2 + t) * kFaTile;
      if (warp_live) {
        uint32_t ls[16], o[64];  // row sum and both halves of O in flight before the single wait
        tmem_ld16(tO + 64, ls);
        tmem_ld32(tO, reinterpret_cast<uint32_t(&)[32]>(o[0]));
        tmem_ld32(tO + 32, reinterpret_cast<uint32_t(&)[32]>(o[32]));
        tmem_wait_ld();
        const float inv = 1.0f / __uint_as_float(ls[0]);
        const int r = quad * 32 + lane;
        uint8_t* rowp = stage + r * 128;
        // split-operand mode: the rounding residuals go straight to global memory (one 128-byte line per thread)
        const bool lo_row = p.ctx_lo_off != 0 && tile * 128 + r < T;
        uint4* lo_dst = reinterpret_cast<uint4*>(p.ctx + (static_cast<long long>(seq) * T + tile * 128 + r) * p.ctx_ld +
                                                 p.ctx_lo_off + head * 64);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float f[8];
#pragma unroll
          for (int j = 0; j < 8; j += 2) {   // packed fp32 pairs: one issue slot per two values
            uint64_t x, iv;
            asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "r"(o[8 * q + j]), "r"(o[8 * q + j + 1]));
            asm("mov.b64 %0, {%1, %1};" : "=l"(iv) : "f"(inv));
            asm("mul.rn.f32x2 %0, %0, %1;" : "+l"(x) : "l"(iv));
            asm("mov.b64 {%0, %1}, %2;" : "=f"(f[j]), "=f"(f[j + 1]) : "l"(x));
          }
          const uint4 w = make_uint4(f2h2_sat(f[0], f[1]), f2h2_sat(f[2], f[3]), f2h2_sat(f[4], f[5]), f2h2_sat(f[6], f[7]));
          *reinterpret_cast<uint4*>(rowp + ((q ^ (r & 7)) << 4)) = w;  // 128B swizzle, matches tmCtx
          if (lo_row)
            lo_dst[q] = make_uint4(f2h2_residual(f[0], f[1], w.x), f2h2_residual(f[2], f[3], w.y),
                                   f2h2_residual(f[4], f[5], w.z), f2h2_residual(f[6], f[7], w.w));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_free[t]);
      fence_proxy_async();            // generic-proxy smem writes -> visible to the TMA (async proxy)
      trace(0x24);  // O staged
      fa_bar_sync(1 + t, 128);
      trace(0x25);  // group barrier passed
      if (leader) {
        release_stage();  // (only when the item had a single sub-block)
        tma_store_3d(&tmCtx, stage, head * 64, tile * 128, seq);  // rows >= T are clipped by the tensor map
        tma_store_commit();
        pending_qs = qs;  // released during the next item (see release_stage)
      }
    }
    if (leader) {
      release_stage();
      tma_store_wait<0>();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace pg

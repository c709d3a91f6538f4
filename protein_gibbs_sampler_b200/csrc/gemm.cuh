// Persistent warp-specialised tcgen05 GEMM for sm_100a.
//
//   C[M,N] = epilogue( A[M,K] (fp16, K-major) x B[N,K]^T (fp16, K-major = nn.Linear weight layout) )
//
// One CTA per SM, static tile schedule (tile = cluster + i*num_clusters, n fastest so the CTAs of a wave
// share A row-blocks in L2).  CG = 1: one CTA computes a 128 x BN tile.  CG = 2: the two CTAs of a (2,1,1)
// cluster compute a 256 x BN tile with tcgen05.mma.cta_group::2 -- each CTA stages its own 128 rows of A and
// BN/2 rows of B, so the operand bytes crossing each SM's shared memory per FLOP are halved (the single-CTA
// 128 x 256 tile is shared-memory-bandwidth bound: 96 B/clk of UMMA operand reads + 96 B/clk of TMA writes).
// Roles (per CTA):
//   warp 0      TMA producer: cp.async.bulk.tensor 128B-swizzled A/B k-blocks into a kStages smem ring
//   warp 1      MMA issuer  : one thread issues tcgen05.mma (M=128, N=BN, K=16) into TMEM, fp32 accumulate
//   warp 2      TMEM allocator (512 columns = two BN-wide accumulator stages)
//   warps 4-11  epilogue    : tcgen05.ld TMEM -> registers, fused bias / q-scale / RoPE / erf-GELU / residual,
//                             vectorised global stores; overlaps the next tile's MMAs (double-buffered TMEM)
//
// These replace the cuBLAS sgemm calls issued from fair-esm's nn.Linear layers on the reference's path
// (call site /root/reference/src/pgen/esm_sampler.py:223).
#pragma once
#include "ptx.cuh"

namespace pg {

enum Epilogue : int {
  EPI_BIAS_F16 = 0,   // out16 = acc + bias
  EPI_GELU_F16 = 1,   // out16 = gelu_erf(acc + bias)
  EPI_RESID_F32 = 2,  // x32  += acc + bias            (in-place residual stream)
  EPI_QKV_F16 = 3,    // out16 = (acc + bias) * (col < q_cols ? q_scale : 1)  [+ rotary on q and k]
  EPI_GELU_F32 = 4,   // out32 = gelu_erf(acc + bias)
  EPI_BIAS_F32 = 5,   // out32 = acc + bias
};

struct GemmParams {
  int M, N, K;          // K a multiple of 64; N a multiple of 16
  const float* bias;    // [N] or nullptr
  void* out;            // fp16 or fp32 [M, ldo]
  int ldo;              // elements
  // EPI_QKV_F16
  int q_cols;           // columns [0,q_cols) are scaled by q_scale
  float q_scale;
  int rope_cols;        // columns [0,rope_cols) get rotary embedding (0 = none); heads are head_dim wide
  int head_dim;
  int seq_len;          // token position t = row % seq_len
  const float* rope;    // [seq_len][cos_0 .. cos_{H-1} | sin_0 .. sin_{H-1}], H = head_dim / 2
  // EPI_RESID_F32 only: K-split of the partly-filled last wave (see gemm_work_unit).  split <= 1: off.
  int reverse;          // walk the tiles from the last row block down (see pgibbs_engine::zigzag)
  int split;            // parts each last-wave tile is cut into along K
  int32_t* flags;       // [last-wave tile][part][CTA rank][epilogue warp]: 1 once that warp's reduce-adds have landed;
                        // zero between launches (the one waiter of a flag clears it), so a launch captured in a
                        // CUDA graph can be replayed as it is
  // Split-operand ("precise") mode.  Operands are fp16 hi + lo pairs stored side by side: A rows are [a_hi (K) | a_lo (K)],
  // B rows [w_hi (K) | w_lo (K)].  The main loop runs k_segs passes over K into the SAME accumulator:
  //   1: a_hi w_hi          2: + a_hi w_lo (weights exact to 2^-22)          3: + a_lo w_hi (activations too)
  // (a_lo w_lo ~ 2^-24 is dropped.)  lo parts are plain fp16: |lo| <= 2^-12 |hi| reaches the subnormal range, whose
  // absolute resolution 2^-24 is ample.  0 is treated as 1.
  int k_segs;
  int lo_off;           // fp16-output epilogues: also store the rounding residual of every output at column + lo_off
                        // (the a_lo half of the next GEMM's A operand); 0 = off
  // Output path of the plain-store epilogues (everything but EPI_RESID_F32).  0: every 32-row x 128-byte chunk is staged
  // in shared memory and leaves through the TMA engine -- the engine and the shared-memory port the operand ring
  // lives on.  1: the chunk goes out straight from registers, one row per lane, as four 32-byte st.global.v8; needs a
  // 32-byte aligned output with a 32-byte multiple pitch.  (FC1 -2.7 % per launch at config 2.  The residual epilogue
  // stays on the TMA reduce-add: red.global.add.v4.f32 from registers and a register read-modify-write with
  // 32-byte loads issued a chunk ahead both measured 17-19 % SLOWER on the out-projection,
  // profiles/r02zzz10_epilogue_direct_ab.txt.)
  int direct;
  // Sub-range of the launch's tile sequence (position = order in which the static schedule hands tiles out, before the
  // `reverse` mapping): this launch computes positions [tile_begin, tile_begin + tile_count).  tile_count == 0: all.
  // The engine cuts a residual GEMM into its full waves and its partly-filled last wave with it (engine.cu:
  // run_gemm_resid_ln), so that the LayerNorm behind it can start on the finished row blocks meanwhile.
  int tile_begin, tile_count;
};

constexpr int kBM = 128;
constexpr int kBK = 64;  // 64 fp16 = one 128-byte swizzle row
constexpr int kGemmThreads = 384;
constexpr int kEpiWarps = 8;

// BN = tile width of the CTA (CG=1) or CTA pair (CG=2); each CTA stages BN/CG rows of B per k-block.
__host__ __device__ constexpr int gemm_stage_bytes(int BN, int CG) { return kBM * kBK * 2 + (BN / CG) * kBK * 2; }
__host__ __device__ constexpr int gemm_stages(int BN, int CG) {
  int s = (227 * 1024 - 2048 - 8 * 2 * 4096) / gemm_stage_bytes(BN, CG);  // minus barriers/alignment and epilogue staging
  return s > 8 ? 8 : s;
}
__host__ __device__ constexpr int gemm_smem_bytes(int BN, int CG) {
  return gemm_stages(BN, CG) * gemm_stage_bytes(BN, CG) + 8 * 2 * 4096 + 2048;
}

// erf-GELU without a branch (tools/fit_gelu.py):  e = exp2(a P(a)) ~= erfc(a / sqrt2) with a = min(|v|, 5.75), and
//   gelu(v) = v Phi(v) = relu(v) - |v| e / 2        (v > 0: v (1 - e/2);  v <= 0: v e/2).
// Max abs error 6.2e-7 -- the same as 0.5 v (1 + erff(v/sqrt2)) evaluated in fp32 (6.8e-7).  12 instructions per
// element (|v| is a free operand modifier, ex2.approx.ftz needs no range fix-up: e <= 1 and a flushed e only
// matters where |v| e / 2 < 1e-37); the FC1 epilogue competes with the MMA / TMA issue warps for issue slots.
__device__ __forceinline__ float gelu_erf(float v) {
  const float a = fminf(fabsf(v), 5.75f);
  float q = 4.278695997e-06f;
  q = fmaf(q, a, -1.279769367e-05f);
  q = fmaf(q, a, -5.757883773e-04f);
  q = fmaf(q, a, 7.670788094e-03f);
  q = fmaf(q, a, -5.294856802e-02f);
  q = fmaf(q, a, -4.590439200e-01f);
  q = fmaf(q, a, -1.151126981e+00f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(q * a));
  return fmaf(-0.5f * fabsf(v), e, fmaxf(v, 0.f));
}

// The same for two values at once on packed fp32 pairs (fma.rn.f32x2 / mul.rn.f32x2 -> FFMA2 / FMUL2: the FMA pipe spends
// the same time per result, but the pair costs ONE issue slot -- the epilogue warps share their schedulers with the TMA
// producer and the MMA issuer, and the erf-GELU is what fills those slots in FC1).  Same operations in the same order per
// value as gelu_erf: bit-identical results.  Measured on one box, A B A B (profiles/r02zu_gelu_packed_ab.txt): FC1 196-206
// -> 183-184 us per launch, config 2 41.98 -> 42.4 iterations/s; the same for the bias add and the q scale of the other
// epilogues changes nothing (they are a few instructions per element to begin with).
#ifndef PGIBBS_GELU_PACKED
#define PGIBBS_GELU_PACKED 1
#endif
__device__ __forceinline__ uint64_t f32x2(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r;
}
__device__ __forceinline__ void gelu_erf2(float& v0, float& v1) {
  const float a0 = fminf(fabsf(v0), 5.75f), a1 = fminf(fabsf(v1), 5.75f);
  const uint64_t a = f32x2(a0, a1);
  uint64_t q = f32x2(4.278695997e-06f, 4.278695997e-06f);
  q = fma2(q, a, f32x2(-1.279769367e-05f, -1.279769367e-05f));
  q = fma2(q, a, f32x2(-5.757883773e-04f, -5.757883773e-04f));
  q = fma2(q, a, f32x2(7.670788094e-03f, 7.670788094e-03f));
  q = fma2(q, a, f32x2(-5.294856802e-02f, -5.294856802e-02f));
  q = fma2(q, a, f32x2(-4.590439200e-01f, -4.590439200e-01f));
  q = fma2(q, a, f32x2(-1.151126981e+00f, -1.151126981e+00f));
  q = mul2(q, a);
  float q0, q1, e0, e1;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(q0), "=f"(q1) : "l"(q));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(q0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(q1));
  const uint64_t h = mul2(f32x2(-0.5f, -0.5f), f32x2(fabsf(v0), fabsf(v1)));
  const uint64_t r = fma2(h, f32x2(e0, e1), f32x2(fmaxf(v0, 0.f), fmaxf(v1, 0.f)));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(v0), "=f"(v1) : "l"(r));
}

#ifndef PGIBBS_BIAS_EARLY
#define PGIBBS_BIAS_EARLY 1
#endif
__device__ __forceinline__ uint32_t pack_half2(float a, float b) { return f2h2_sat(a, b); }

__host__ __device__ constexpr bool epi_out_f16(int epi) { return epi == EPI_BIAS_F16 || epi == EPI_GELU_F16 || epi == EPI_QKV_F16; }
constexpr int kEpiStageBytes = 4096;                           // 32 rows x 128 B, one warp's chunk
constexpr int kEpiStagingTotal = kEpiWarps * 2 * kEpiStageBytes;  // double-buffered per warp

// rotary embedding on one 64-column chunk (chunk start is head-aligned): x*cos + rotate_half(x)*sin.
// Table row of token t: [cos_0 .. cos_{H-1} | sin_0 .. sin_{H-1}], H = head_dim / 2, so that four consecutive cosines (sines)
// come in with one 16-byte load and sit in adjacent registers: the rotation then runs on packed fp32 pairs (FMUL2 / FFMA2,
// one issue slot per two values -- the epilogue warps share their schedulers with the TMA producer and the MMA issuer, and
// at ESM-2's shapes this epilogue is what fills them).
__device__ __forceinline__ uint64_t f32x2_of(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
template <int HD>
__device__ __forceinline__ void rope_chunk(float (&v)[64], const float* __restrict__ cs) {
  constexpr int H = HD / 2;
  static_assert(H % 4 == 0, "head_dim is a multiple of 8");
#pragma unroll
  for (int h0 = 0; h0 < 64; h0 += HD) {
#pragma unroll
    for (int j = 0; j < H; j += 4) {
      const float4 c = __ldg(reinterpret_cast<const float4*>(cs + j));
      const float4 sn = __ldg(reinterpret_cast<const float4*>(cs + H + j));
#pragma unroll
      for (int u = 0; u < 4; u += 2) {
        const uint64_t a = f32x2_of(v[h0 + j + u], v[h0 + j + u + 1]), b = f32x2_of(v[h0 + j + u + H], v[h0 + j + u + 1 + H]);
        const uint64_t c2 = u ? f32x2_of(c.z, c.w) : f32x2_of(c.x, c.y), s2 = u ? f32x2_of(sn.z, sn.w) : f32x2_of(sn.x, sn.y);
        uint64_t ac, bc, bs, na, nb;
        asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(ac) : "l"(a), "l"(c2));
        asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(bc) : "l"(b), "l"(c2));
        asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(bs) : "l"(b), "l"(s2));
        asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(na) : "l"(ac), "l"(bs));                    // first half: -x2*sin (FADD2, negated operand)
        asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(nb) : "l"(a), "l"(s2), "l"(bc));        // second half: +x1*sin
        asm("mov.b64 {%0, %1}, %2;" : "=f"(v[h0 + j + u]), "=f"(v[h0 + j + u + 1]) : "l"(na));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(v[h0 + j + u + H]), "=f"(v[h0 + j + u + 1 + H]) : "l"(nb));
      }
    }
  }
}

// Static schedule of one CTA group (a CTA, or a CTA pair): tiles group, group + G, group + 2G, ...  When the tile count
// is not a multiple of G the last wave leaves G - rem groups idle for a whole tile time (out-projection / FC2 at
// config 2: 325 tiles on 74 pairs = 4.39 waves run as 5).  For the residual epilogue -- whose output is ADDED to the
// fp32 stream, so partial sums over K can be added one after the other -- each last-wave tile is cut into `split`
// parts along K, run by `split` different groups at the same time.  The parts' reduce-adds are ordered (part j's
// epilogue warp waits for a flag from the same warp of part j-1, which owns the same output chunks), so every
// element is still x + p0 + p1 + ... in one fixed order: the result does not depend on timing.
struct WorkUnit {
  int tile, kb0, kb1;
  int part;   // -1: whole tile; otherwise which K part
  int slot;   // index of the tile inside the last wave (flag addressing)
};
__device__ __forceinline__ bool gemm_work_unit(int it, int group, int num_groups, int num_tiles, int num_kb, int split,
                                               WorkUnit& u) {
  u.part = -1; u.slot = 0; u.kb0 = 0; u.kb1 = num_kb;
  u.tile = group + it * num_groups;
  if (split <= 1) return u.tile < num_tiles;
  const int full_waves = num_tiles / num_groups;
  if (it < full_waves) return true;
  if (it > full_waves) return false;
  const int rem = num_tiles - full_waves * num_groups;
  if (group >= rem * split) return false;
  u.slot = group / split;
  u.part = group - u.slot * split;
  u.tile = full_waves * num_groups + u.slot;
  u.kb0 = u.part * num_kb / split;
  u.kb1 = (u.part + 1) * num_kb / split;
  return true;
}

__device__ __forceinline__ int ld_acquire_gpu(const int32_t* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// Bounded spin on a flag in global memory: a protocol bug traps (launch failure) instead of hanging the GPU.
__device__ __forceinline__ void spin_until_equal(const int32_t* p, int want) {
  uint32_t spins = 0;
  while (ld_acquire_gpu(p) != want) {
    __nanosleep(40);
    if (++spins > (1u << 25)) {
      printf("pgibbs: GEMM split flag wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void st_release_gpu(int32_t* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Drain one accumulator tile row-slice.  This warp owns TMEM lanes [quad*32, quad*32+32) (= 32 output rows, one per
// thread) and every second chunk (`par`) of 128 output bytes per row (64 fp16 / 32 fp32 columns).  Plain outputs leave
// straight from the registers as four 32-byte stores per lane (GemmParams::direct); the residual stream -- and plain
// outputs whose buffer is not 32-byte aligned -- are staged chunk by chunk in this warp's own 128B-swizzled
// shared-memory buffer and handed to one TMA reduce-add (store): full lines, asynchronous, clipped at the M / N edges.
//   t_row = TMEM address of the slice's column 0; row0 = first output row of the slice; n0 = first output column.
template <int BN, int EPI>
__device__ __forceinline__ void gemm_epilogue_tile(const GemmParams& p, const CUtensorMap* tmC, uint32_t t_row,
                                                   int row0, int lane, int n0, int par, uint8_t* staging, int& sbuf,
                                                   bool add_bias = true) {
  constexpr bool kOut16 = epi_out_f16(EPI);
  constexpr int kCW = kOut16 ? 64 : 32;  // columns per chunk
  static_assert(BN % kCW == 0, "tile width must be a whole number of epilogue chunks");
  constexpr int kChunks = BN / kCW;
  for (int c = par; c < kChunks; c += 2) {
    const int g = n0 + c * kCW;
    if (g >= p.N) break;  // warp-uniform: chunk entirely beyond the N edge
    float v[kCW];
    // fp32 epilogues (32-column chunks; the residual one is the tile period of the short-K out-projection): fetch the
    // chunk's bias BEFORE the TMEM load so that the two latencies overlap (the volatile tcgen05.ld / wait pair keeps
    // the compiler from hoisting the loads itself; ncu: the first bias add was the epilogue's second-largest stall).
    constexpr bool kBiasEarly = !kOut16 && PGIBBS_BIAS_EARLY;
    float4 bq[kBiasEarly ? kCW / 4 : 1];
    const bool bias_early = kBiasEarly && p.bias && add_bias && g + kCW <= p.N;
    if constexpr (kBiasEarly) {
      if (bias_early) {
#pragma unroll
        for (int j4 = 0; j4 < kCW / 4; ++j4) bq[j4] = __ldg(reinterpret_cast<const float4*>(p.bias + g) + j4);
      }
    }
#pragma unroll
    for (int hlf = 0; hlf < kCW / 32; ++hlf) {
      uint32_t r[32];
      tmem_ld32(t_row + c * kCW + hlf * 32, r);
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 32; ++j) v[hlf * 32 + j] = __uint_as_float(r[j]);
    }
    if (bias_early) {
      if constexpr (kBiasEarly) {
#pragma unroll
        for (int j4 = 0; j4 < kCW / 4; ++j4) {
          v[4 * j4 + 0] += bq[j4].x; v[4 * j4 + 1] += bq[j4].y; v[4 * j4 + 2] += bq[j4].z; v[4 * j4 + 3] += bq[j4].w;
        }
      }
    } else if (p.bias && add_bias) {
      if (g + kCW <= p.N) {
#pragma unroll
        for (int j4 = 0; j4 < kCW / 4; ++j4) {
          const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + g) + j4);
          v[4 * j4 + 0] += b.x; v[4 * j4 + 1] += b.y; v[4 * j4 + 2] += b.z; v[4 * j4 + 3] += b.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < kCW; ++j) v[j] += (g + j < p.N) ? __ldg(p.bias + g + j) : 0.f;
      }
    }
    if constexpr (EPI == EPI_GELU_F16 || EPI == EPI_GELU_F32) {
#if PGIBBS_GELU_PACKED
#pragma unroll
      for (int j = 0; j < kCW; j += 2) gelu_erf2(v[j], v[j + 1]);
#else
#pragma unroll
      for (int j = 0; j < kCW; ++j) v[j] = gelu_erf(v[j]);
#endif
    }
    if constexpr (EPI == EPI_QKV_F16) {
      if (g < p.q_cols) {
#pragma unroll
        for (int j = 0; j < kCW; ++j) v[j] *= p.q_scale;
      }
      if (g < p.rope_cols) {
        const int t = (row0 + lane) % p.seq_len;
        const float* cs = p.rope + static_cast<size_t>(t) * p.head_dim;
        if (p.head_dim == 64) rope_chunk<64>(v, cs);
        else if (p.head_dim == 32) rope_chunk<32>(v, cs);
        else rope_chunk<16>(v, cs);
      }
    }
    // Stage 32 rows x 128 B in this warp's 128B-swizzled buffer and hand it to one TMA store / reduce-add at column
    // `col`.  The store issued from the same buffer two stores ago must have finished reading it.
    auto stage_and_store = [&](const uint4 (&w)[8], int col) {
      if (lane == 0) tma_store_wait_read<1>();
      __syncwarp();
      uint8_t* buf = staging + sbuf * kEpiStageBytes;
      uint8_t* rowp = buf + lane * 128;
#pragma unroll
      for (int q = 0; q < 8; ++q)
        *reinterpret_cast<uint4*>(rowp + ((q ^ (lane & 7)) << 4)) = w[q];  // 128B swizzle: conflict-free, matches tmC
      fence_proxy_async();  // generic-proxy smem writes -> visible to the TMA (async proxy)
      __syncwarp();
      if (lane == 0) {
        if constexpr (EPI == EPI_RESID_F32) tma_reduce_add_2d(tmC, buf, col, row0);
        else tma_store_2d(tmC, buf, col, row0);
        tma_store_commit();
      }
      sbuf ^= 1;
    };
    if (EPI != EPI_RESID_F32 && p.direct) {
      // one output row per lane (row0 + lane), kCW consecutive columns from g: 128 bytes = 4 x 32-byte stores; rows
      // beyond M and column groups beyond N (a multiple of 16) are skipped here -- the TMA path gets that clipping
      // from the tensor map
      const int row = row0 + lane;
      if (row < p.M) {
        if constexpr (kOut16) {
          __half* dst = reinterpret_cast<__half*>(p.out) + static_cast<size_t>(row) * p.ldo + g;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (g + 16 * q + 16 > p.N) break;
            const float* x = v + 16 * q;
            const uint4 a = make_uint4(pack_half2(x[0], x[1]), pack_half2(x[2], x[3]), pack_half2(x[4], x[5]), pack_half2(x[6], x[7]));
            const uint4 b = make_uint4(pack_half2(x[8], x[9]), pack_half2(x[10], x[11]), pack_half2(x[12], x[13]), pack_half2(x[14], x[15]));
            st_global_v8(dst + 16 * q, a, b);
            if constexpr (EPI == EPI_GELU_F16) {
              if (p.lo_off)
                st_global_v8(dst + p.lo_off + 16 * q,
                             make_uint4(f2h2_residual(x[0], x[1], a.x), f2h2_residual(x[2], x[3], a.y),
                                        f2h2_residual(x[4], x[5], a.z), f2h2_residual(x[6], x[7], a.w)),
                             make_uint4(f2h2_residual(x[8], x[9], b.x), f2h2_residual(x[10], x[11], b.y),
                                        f2h2_residual(x[12], x[13], b.z), f2h2_residual(x[14], x[15], b.w)));
            }
          }
        } else {
          float* dst = reinterpret_cast<float*>(p.out) + static_cast<size_t>(row) * p.ldo + g;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (g + 8 * q + 8 > p.N) break;
            const float* x = v + 8 * q;
            st_global_v8(dst + 8 * q,
                         make_uint4(__float_as_uint(x[0]), __float_as_uint(x[1]), __float_as_uint(x[2]), __float_as_uint(x[3])),
                         make_uint4(__float_as_uint(x[4]), __float_as_uint(x[5]), __float_as_uint(x[6]), __float_as_uint(x[7])));
          }
        }
      }
      continue;
    }
    uint4 w[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      if constexpr (kOut16) {
        w[q] = make_uint4(pack_half2(v[8 * q], v[8 * q + 1]), pack_half2(v[8 * q + 2], v[8 * q + 3]),
                          pack_half2(v[8 * q + 4], v[8 * q + 5]), pack_half2(v[8 * q + 6], v[8 * q + 7]));
      } else {
        w[q] = make_uint4(__float_as_uint(v[4 * q]), __float_as_uint(v[4 * q + 1]), __float_as_uint(v[4 * q + 2]),
                          __float_as_uint(v[4 * q + 3]));
      }
    }
    stage_and_store(w, g);
    if constexpr (EPI == EPI_GELU_F16) {
      if (p.lo_off) {  // the rounding residuals of this chunk: the a_lo half of the FC2 operand (split-operand mode)
#pragma unroll
        for (int q = 0; q < 8; ++q)
          w[q] = make_uint4(f2h2_residual(v[8 * q], v[8 * q + 1], w[q].x), f2h2_residual(v[8 * q + 2], v[8 * q + 3], w[q].y),
                            f2h2_residual(v[8 * q + 4], v[8 * q + 5], w[q].z), f2h2_residual(v[8 * q + 6], v[8 * q + 7], w[q].w));
        stage_and_store(w, g + p.lo_off);
      }
    }
  }
}

template <int BN, int EPI, int CG>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmC, const GemmParams p) {
  static_assert(CG == 1 || CG == 2, "CTA group is 1 or 2");
  static_assert(BN % 16 == 0 && BN >= 32 && BN <= 256 && (BN / CG) % 8 == 0, "invalid UMMA N");
  constexpr int kStages = gemm_stages(BN, CG);
  constexpr int kBNL = BN / CG;                 // rows of B staged by this CTA
  constexpr int kABytes = kBM * kBK * 2;
  constexpr int kBBytes = kBNL * kBK * 2;
  constexpr int kStageBytes = kABytes + kBBytes;
  constexpr uint32_t kIdesc = make_idesc_f16(kBM * CG, BN, false, false);

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment is required by the 128B swizzle atoms.  (Both CTAs of a pair compute the same offset:
  // the dynamic shared window starts at the same shared::cta address in every CTA of a launch.)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ring = smem;
  uint8_t* staging = smem + kStages * kStageBytes;  // epilogue: 8 warps x 2 x 4 KB (1024-aligned)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + kEpiStagingTotal);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full = empty_bar + kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = CG == 2 ? static_cast<int>(cluster_ctarank()) : 0;
  const int group = blockIdx.x / CG, num_groups = gridDim.x / CG;
  const int m_tiles = (p.M + kBM * CG - 1) / (kBM * CG);
  const int n_tiles = (p.N + BN - 1) / BN;
  const int total_tiles = m_tiles * n_tiles;
  const int num_tiles = p.tile_count > 0 ? p.tile_count : total_tiles;   // positions this launch works through
  const int kb_per_seg = p.K / kBK;
  const int num_kb = kb_per_seg * (p.k_segs > 1 ? p.k_segs : 1);   // split-operand mode: 2 or 3 passes over K
  const int split = EPI == EPI_RESID_F32 ? p.split : 1;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    prefetch_tmap(&tmC);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);   // the leader's producer (arrive.expect_tx for both CTAs' bytes)
      mbar_init(&empty_bar[s], 1);  // one tcgen05.commit (multicast to both CTAs of a pair)
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], kEpiWarps * CG);  // the epilogue warps of both CTAs release the leader's MMA warp
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if constexpr (CG == 2) { tmem_alloc_pair(tmem_slot, 512); tmem_relinquish_pair(); }
    else { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  }
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // everything above is local set-up and may overlap the previous kernel's tail; operands and outputs are not
  pdl_wait();
  pdl_launch_dependents();

  // Producer and MMA warps: the WHOLE warp walks the tile schedule (barrier waits, address arithmetic) and one
  // elected lane issues the TMA / tcgen05 instructions.  Under a `lane == 0` branch the operands are per-thread
  // values and the compiler wraps every UTCHMMA / UTMALDG in an elect + R2UR waterfall loop (~12 instructions);
  // executed warp-uniformly they live in uniform registers and the instructions issue back to back.
  if (warp == 0) {
    // ------------------------------------------------------------ producer
    const bool issuer = elect_one();
    int stage = 0;
    uint32_t phase = 0;
    WorkUnit u;
    for (int it = 0; gemm_work_unit(it, group, num_groups, num_tiles, num_kb, split, u); ++it) {
      const int tile = p.reverse ? total_tiles - 1 - (u.tile + p.tile_begin) : u.tile + p.tile_begin;
      const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
      const int a_row = (m_blk * CG + rank) * kBM, b_row = n_blk * BN + rank * kBNL;
      for (int kb = u.kb0; kb < u.kb1; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = ring + stage * kStageBytes;
        // pass 0: a_hi w_hi, pass 1: a_hi w_lo, pass 2: a_lo w_hi (lo halves sit K columns to the right)
        const int seg = kb / kb_per_seg, kk = (kb - seg * kb_per_seg) * kBK;
        const int a_col = kk + (seg == 2 ? p.K : 0), b_col = kk + (seg == 1 ? p.K : 0);
        if (issuer) {
          if constexpr (CG == 2) {
            if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * kStageBytes);
            tma_load_2d_pair(sa, &tmA, &full_bar[stage], a_col, a_row);
            tma_load_2d_pair(sa + kABytes, &tmB, &full_bar[stage], b_col, b_row);
          } else {
            mbar_arrive_expect_tx(&full_bar[stage], kStageBytes);
            tma_load_2d(sa, &tmA, &full_bar[stage], a_col, a_row);
            tma_load_2d(sa + kABytes, &tmB, &full_bar[stage], b_col, b_row);
          }
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------- MMA issuer (leader CTA of a pair only)
    if (rank == 0) {
      const bool issuer = elect_one();
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      const uint32_t desc_hi = static_cast<uint32_t>(make_smem_desc_sw128(0, 1024) >> 32);
      WorkUnit u;
      for (int it = 0; gemm_work_unit(it, group, num_groups, num_tiles, num_kb, split, u); ++it) {
        mbar_wait(&tmem_empty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = u.kb0; kb < u.kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(ring + stage * kStageBytes);
          const uint32_t a_lo = static_cast<uint32_t>(make_smem_desc_sw128(sa, 1024));
          const uint32_t b_lo = static_cast<uint32_t>(make_smem_desc_sw128(sa + kABytes, 1024));
          if (issuer) {
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k) {
              // +32 bytes per UMMA_K slice inside the 128B swizzle row (address field is >>4)
              const uint64_t adesc = (static_cast<uint64_t>(desc_hi) << 32) | (a_lo + 2 * k);
              const uint64_t bdesc = (static_cast<uint64_t>(desc_hi) << 32) | (b_lo + 2 * k);
              const uint32_t acc = (kb != u.kb0 || k != 0) ? 1u : 0u;
              if constexpr (CG == 2) umma_f16_ss_pair(d_tmem, adesc, bdesc, kIdesc, acc);
              else umma_f16_ss(d_tmem, adesc, bdesc, kIdesc, acc);
            }
            // frees this smem stage (in both CTAs) when the MMAs above retire
            if constexpr (CG == 2) umma_commit_pair(&empty_bar[stage], 0b11); else umma_commit(&empty_bar[stage]);
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        // accumulator ready for the epilogue warps (of both CTAs)
        if (issuer) {
          if constexpr (CG == 2) umma_commit_pair(&tmem_full[as], 0b11); else umma_commit(&tmem_full[as]);
        }
        __syncwarp();
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue
    const int ew = warp - 4;
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    const int par = ew >> 2;    // which alternating 16-column chunks this warp takes
    int as = 0;
    uint32_t aphase = 0;
    const uint32_t leader_empty = CG == 2 ? mapa_u32(smem_u32(&tmem_empty[0]), 0) : 0;
    uint8_t* my_staging = staging + ew * 2 * kEpiStageBytes;
    int sbuf = 0;
    WorkUnit u;
    for (int it = 0; gemm_work_unit(it, group, num_groups, num_tiles, num_kb, split, u); ++it) {
      const int tile = p.reverse ? total_tiles - 1 - (u.tile + p.tile_begin) : u.tile + p.tile_begin;
      const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
      mbar_wait(&tmem_full[as], aphase);
      tc_fence_after();
      const int row0 = (m_blk * CG + rank) * kBM + quad * 32;
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * BN;
      if constexpr (EPI == EPI_RESID_F32) {
        // K-split tile: this warp's chunks are added after the same warp of the previous part has added its own
        int32_t* flag = p.flags + ((u.slot * split + (u.part > 0 ? u.part : 0)) * CG + rank) * kEpiWarps + ew;
        if (u.part > 0) {
          if (lane == 0) {
            spin_until_equal(flag - CG * kEpiWarps, 1);
            *(flag - CG * kEpiWarps) = 0;   // this warp is the flag's only waiter: leave it clear for the next launch
          }
          __syncwarp();
          fence_proxy_async_all();
        }
        gemm_epilogue_tile<BN, EPI>(p, &tmC, t_row, row0, lane, n_blk * BN, par, my_staging, sbuf, u.part <= 0);
        if (u.part >= 0 && u.part < split - 1) {
          if (lane == 0) {
            tma_store_wait<0>();   // the reduce-adds of this part have been performed
            fence_proxy_async_all();
            __threadfence();
            st_release_gpu(flag, 1);
          }
          __syncwarp();
        }
      } else {
        gemm_epilogue_tile<BN, EPI>(p, &tmC, t_row, row0, lane, n_blk * BN, par, my_staging, sbuf);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CG == 2) mbar_arrive_cluster(leader_empty + as * 8);
        else mbar_arrive(&tmem_empty[as]);
      }
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
    if (lane == 0) tma_store_wait<0>();  // all output tiles of this warp are in global memory
  }

  tc_fence_before();
  // A CTA of a pair must outlive every remote access to it (peer MMAs read its smem, commits arrive on its barriers).
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if constexpr (CG == 2) tmem_dealloc_pair(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace pg

// HBM-bound row-wise kernels of the Gibbs step: <mask> scatter, embedding prologue (token-dropout
// rescale + learned positions + LayerNorm-before), LayerNorm (fp32 residual stream -> fp16 GEMM operand),
// and the fused LM-head tail (LayerNorm + tied-embedding projection + top-k / categorical draw + write-back).
// One warp per row everywhere; rows are 128-bit vectorised.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "ptx.cuh"

namespace pg {

constexpr int kMaxVecPerLane = 20;  // float4 per lane -> embed_dim <= 2560

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Position schedule view: positions[iter*iter_stride + chain*chain_stride + p], p < P.
// "all positions" and in-order schedules share one list (strides 0).
// Chain c owns token row  c*seq_stride + seq_offset  (1/0 normally; R/target_row for generate_single).
struct Schedule {
  const int32_t* positions;
  long long iter_stride, chain_stride;
  int P;
  int seq_stride, seq_offset;
  // Iteration index kept on the device (CUDA-graph replay of one captured iteration): when set, kernels read the
  // iteration from here instead of their `iter` argument and advance_iter_kernel bumps it at the end of the iteration.
  const int32_t* iter_dev;
};

__global__ void set_iter_kernel(int32_t* it, int value) { *it = value; }
__global__ void advance_iter_kernel(int32_t* it) {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  *it += 1;
}

// ---------------------------------------------------------------------------------------------
// <mask> scatter: tokens[chain][pos] = mask_idx for every scheduled position.
// Reference: ESM_sampler.mask_target_indexes (esm_sampler.py:259-262), MSA variant esm_msa_sampler.py:255-259.
// ---------------------------------------------------------------------------------------------
__global__ void mask_scatter_kernel(int32_t* __restrict__ tokens, int n_chains, int T, Schedule s, int iter,
                                    int mask_idx) {
  // programmatic dependent launch (no-ops in a normal launch): wait for the previous kernel, release the next one
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= static_cast<long long>(n_chains) * s.P) return;
  const int chain = static_cast<int>(i / s.P), p = static_cast<int>(i % s.P);
  if (s.iter_dev) iter = *s.iter_dev;
  const int pos = s.positions[iter * s.iter_stride + chain * s.chain_stride + p];
  tokens[(static_cast<long long>(chain) * s.seq_stride + s.seq_offset) * T + pos] = mask_idx;
}

// ---------------------------------------------------------------------------------------------
// Embedding prologue.  One warp per token row.
//   ESM-1b/ESM-2 (fair-esm ProteinBertModel / ESM2):
//       x = E[tok] (0 for <mask>) * 0.88 / (1 - n_mask/T)   [token_dropout]
//       x += P[t + 2]   then LayerNorm_before                [learned positions, ESM-1b]
//   MSA Transformer (App. A.4): x = E[tok] + P[c + 2] + P_row[r]; LayerNorm_before.  No token dropout.
// No <pad> appears on the Gibbs path; the host rejects padded inputs.
// ---------------------------------------------------------------------------------------------
struct EmbedParams {
  const int32_t* tokens;   // [n_seq, T]
  const float* tok_emb;    // [V, d]
  const float* pos_emb;    // [max_pos + 2, d] or nullptr
  const float* row_emb;    // [1024, d] (MSA) or nullptr
  const float* ln_w;       // LayerNorm-before (nullptr = none)
  const float* ln_b;
  float* x;                // [n_seq*T, d]
  int n_seq, T, d, rows_per_msa;
  int mask_idx, token_dropout;
  float eps;
  float scale;             // embedding scale (sqrt(d) for ESM-1, else 1)
};

// VPL = float4 vectors held per lane (>= ceil(d / 128)), sized to the model like the LayerNorm kernel's: with the
// d <= 2560 worst case baked in (80 registers of row data) two CTAs fit an SM and the kernel runs at 0.13 of the HBM
// rate; sized to d = 1280 three do, without spills.
template <int VPL>
__global__ void __launch_bounds__(256, VPL <= 10 ? 3 : 2) embed_kernel(EmbedParams p) {
  // programmatic dependent launch (no-ops in a normal launch): wait for the previous kernel, release the next one
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int warps_per_block = blockDim.x >> 5;
  const int row = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= p.n_seq * p.T) return;
  const int seq = row / p.T, t = row % p.T;
  const int32_t* toks = p.tokens + static_cast<long long>(seq) * p.T;
  const int tok = toks[t];
  float den = 1.0f;
  if (p.token_dropout) {
    int n_mask = 0;
    for (int i = lane; i < p.T; i += 32) n_mask += (toks[i] == p.mask_idx);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n_mask += __shfl_xor_sync(0xffffffffu, n_mask, o);
    // mask_ratio_observed = n_mask / src_length; no <pad> on this path so src_length == T
    den = 1.0f - __fdiv_rn(static_cast<float>(n_mask), static_cast<float>(p.T));
  }
  const bool zero = p.token_dropout && tok == p.mask_idx;
  const float4* e = reinterpret_cast<const float4*>(p.tok_emb + static_cast<long long>(tok) * p.d);
  const float4* pe = p.pos_emb ? reinterpret_cast<const float4*>(p.pos_emb + static_cast<long long>(t + 2) * p.d)
                               : nullptr;
  const float4* re = p.row_emb ? reinterpret_cast<const float4*>(
                                     p.row_emb + static_cast<long long>(seq % p.rows_per_msa) * p.d)
                               : nullptr;
  const int nvec = p.d >> 2;
  float4 v[VPL];
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    const int i = lane + k * 32;
    if (i < nvec) {
      float4 a = zero ? make_float4(0, 0, 0, 0) : __ldg(e + i);
      if (p.scale != 1.0f) { a.x *= p.scale; a.y *= p.scale; a.z *= p.scale; a.w *= p.scale; }
      if (p.token_dropout) {
        // x * (1 - 0.15*0.8) / (1 - ratio): multiply then divide, in fair-esm's order
        const float num = 0.88f;
        a.x = __fdiv_rn(a.x * num, den); a.y = __fdiv_rn(a.y * num, den);
        a.z = __fdiv_rn(a.z * num, den); a.w = __fdiv_rn(a.w * num, den);
      }
      if (pe) { const float4 b = __ldg(pe + i); a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
      if (re) { const float4 b = __ldg(re + i); a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
      v[k] = a;
      sum += a.x + a.y + a.z + a.w;
    }
  }
  float4* out = reinterpret_cast<float4*>(p.x + static_cast<long long>(row) * p.d);
  if (p.ln_w) {
    const float mean = warp_sum(sum) / p.d;
    float sq = 0.f;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int i = lane + k * 32;
      if (i < nvec) {
        const float4 a = v[k];
        sq += (a.x - mean) * (a.x - mean) + (a.y - mean) * (a.y - mean) + (a.z - mean) * (a.z - mean) +
              (a.w - mean) * (a.w - mean);
      }
    }
    const float rstd = 1.0f / sqrtf(warp_sum(sq) / p.d + p.eps);
    const float4* w = reinterpret_cast<const float4*>(p.ln_w);
    const float4* b = reinterpret_cast<const float4*>(p.ln_b);
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int i = lane + k * 32;
      if (i < nvec) {
        const float4 a = v[k], g = __ldg(w + i), h = __ldg(b + i);
        out[i] = make_float4((a.x - mean) * rstd * g.x + h.x, (a.y - mean) * rstd * g.y + h.y,
                             (a.z - mean) * rstd * g.z + h.z, (a.w - mean) * rstd * g.w + h.w);
      }
    }
  } else {
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int i = lane + k * 32;
      if (i < nvec) out[i] = v[k];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// LayerNorm over rows of the fp32 residual stream -> fp16 (GEMM A operand) or fp32.
// Optional gather: output row i reads source row  chain*T + positions[...]  (LM head on sampled rows only).
// ---------------------------------------------------------------------------------------------
struct LnParams {
  const float* x;  // [rows_src, d]
  const float* w;
  const float* b;
  void* out;       // [rows_out, d] fp16 or fp32
  int rows_out, d;
  float eps;
  // gather (positions == nullptr -> identity)
  Schedule sched;
  int iter, T;
  int identity;    // copy the (gathered) rows without normalising (ESM-1 has no final LayerNorm)
  int reverse;     // blocks take the rows from the last one down (see pgibbs_engine::zigzag)
  int ld_out;      // output row pitch in elements (0 = d)
  int lo_off;      // fp16 output: also store the rounding residual y - fp16(y) at column + lo_off (split-operand
                   // mode: the a_lo half of the GEMM operand); 0 = off
};

// VPL = float4 vectors held per lane (>= ceil(d / 128)): sized to the model so that the row fits in few registers
// and four CTAs stay resident per SM (the kernel is latency-bound otherwise: ~3.6 TB/s with the d <= 2560 worst case
// baked in, measured at d = 1280).
template <bool OUT_F16, int VPL>
__global__ void __launch_bounds__(256, VPL <= 10 ? 4 : 2) layernorm_kernel(LnParams p) {
  // programmatic dependent launch (no-ops in a normal launch): wait for the kernel that wrote x, then let the next
  // kernel's blocks be scheduled as this one's finish
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int warps_per_block = blockDim.x >> 5;
  int orow = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (orow >= p.rows_out) return;
  if (p.reverse) orow = p.rows_out - 1 - orow;
  long long srow = orow;
  if (p.sched.positions) {
    const int chain = orow / p.sched.P, q = orow % p.sched.P;
    srow = (static_cast<long long>(chain) * p.sched.seq_stride + p.sched.seq_offset) * p.T +
           p.sched.positions[(p.sched.iter_dev ? *p.sched.iter_dev : p.iter) * p.sched.iter_stride +
                             chain * p.sched.chain_stride + q];
  }
  const float4* in = reinterpret_cast<const float4*>(p.x + srow * p.d);
  const int nvec = p.d >> 2;
  float4 v[VPL];
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    const int i = lane + k * 32;
    if (i < nvec) {
      v[k] = in[i];
      sum += v[k].x + v[k].y + v[k].z + v[k].w;
    }
  }
  float mean = warp_sum(sum) / p.d;
  float sq = 0.f;
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    const int i = lane + k * 32;
    if (i < nvec) {
      const float4 a = v[k];
      sq += (a.x - mean) * (a.x - mean) + (a.y - mean) * (a.y - mean) + (a.z - mean) * (a.z - mean) +
            (a.w - mean) * (a.w - mean);
    }
  }
  float rstd = 1.0f / sqrtf(warp_sum(sq) / p.d + p.eps);
  if (p.identity) { mean = 0.f; rstd = 1.f; }
  const float4* w = reinterpret_cast<const float4*>(p.w);
  const float4* b = reinterpret_cast<const float4*>(p.b);
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    const int i = lane + k * 32;
    if (i < nvec) {
      const float4 a = v[k];
      const float4 g = p.identity ? make_float4(1.f, 1.f, 1.f, 1.f) : __ldg(w + i);
      const float4 h = p.identity ? make_float4(0.f, 0.f, 0.f, 0.f) : __ldg(b + i);
      const float y0 = (a.x - mean) * rstd * g.x + h.x, y1 = (a.y - mean) * rstd * g.y + h.y;
      const float y2 = (a.z - mean) * rstd * g.z + h.z, y3 = (a.w - mean) * rstd * g.w + h.w;
      const long long obase = static_cast<long long>(orow) * (p.ld_out ? p.ld_out : p.d);
      if constexpr (OUT_F16) {
        const uint2 u = make_uint2(f2h2_sat(y0, y1), f2h2_sat(y2, y3));
        reinterpret_cast<uint2*>(reinterpret_cast<__half*>(p.out) + obase)[i] = u;
        if (p.lo_off)
          reinterpret_cast<uint2*>(reinterpret_cast<__half*>(p.out) + obase + p.lo_off)[i] =
              make_uint2(f2h2_residual(y0, y1, u.x), f2h2_residual(y2, y3, u.y));
      } else {
        reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + obase)[i] = make_float4(y0, y1, y2, y3);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// ESM-1 `add_bias_kv`: every sequence has one extra key/value slot (the row after its last token) whose k and v are
// the layer's learned bias_k / bias_v (fair-esm MultiheadAttention: k = cat([k, bias_k]), v = cat([v, bias_v])).
// ---------------------------------------------------------------------------------------------
__global__ void bias_kv_kernel(__half* __restrict__ qkv, const float* __restrict__ bias_k,
                               const float* __restrict__ bias_v, int n_seq, int T, int d) {
  // programmatic dependent launch (no-ops in a normal launch): wait for the previous kernel, release the next one
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_seq * d) return;
  const int seq = i / d, c = i % d;
  __half* row = qkv + (static_cast<long long>(seq) * T + (T - 1)) * 3 * d;
  row[d + c] = __float2half_rn(bias_k[c]);
  row[2 * d + c] = __float2half_rn(bias_v[c]);
}

// ---------------------------------------------------------------------------------------------
// Counter-based device RNG (Philox4x32-10) for the exponential race when no replay noise is supplied.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}

// ---------------------------------------------------------------------------------------------
// generate_step (reference esm_sampler.py:8-45) for one row, executed by one warp.
// `mine`/`extra`: lane v holds logit v (v < 32) / logit v+32.  Returns the sampled token id (all lanes).
//   l = logits / temperature ; sub = l[valid_idx] ; (vals, idx) = topk(sub, k)
//   Categorical(logits=vals).sample() == argmax_j softmax(vals - logsumexp(vals))_j / q_j,  q ~ Exp(1)
//   (torch.multinomial's n_sample == 1 path); noise slot j pairs with the j-th largest candidate.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int generate_step_warp(float mine, float extra, int lane, int row, int iter,
                                                  const int32_t* __restrict__ valid_ids, int n, int k,
                                                  float temperature, const float* __restrict__ noise,
                                                  int noise_stride, unsigned long long seed,
                                                  long long rng_row_offset = 0) {
  // Up to 64 candidates: lane holds candidate slots `lane` (a) and `lane + 32` (b).
  const bool has_a = lane < n, has_b = lane + 32 < n;
  const int id_a = has_a ? valid_ids[lane] : 0, id_b = has_b ? valid_ids[lane + 32] : 0;
  auto fetch = [&](int id) {
    const float lo = __shfl_sync(0xffffffffu, mine, id & 31);
    const float hi = __shfl_sync(0xffffffffu, extra, id & 31);
    return id >= 32 ? hi : lo;
  };
  float la = fetch(id_a), lb = fetch(id_b);
  if (temperature == temperature) { la = __fdiv_rn(la, temperature); lb = __fdiv_rn(lb, temperature); }  // NaN = None
  if (!has_a) la = -INFINITY;
  if (!has_b) lb = -INFINITY;
  // rank = position in the descending top-k order (ties: lower candidate slot first)
  int rank_a = 0, rank_b = 0;
  for (int j = 0; j < n; ++j) {
    const float oa = __shfl_sync(0xffffffffu, la, j & 31), ob = __shfl_sync(0xffffffffu, lb, j & 31);
    const float o = j < 32 ? oa : ob;
    rank_a += (o > la) || (o == la && j < lane);
    rank_b += (o > lb) || (o == lb && j < lane + 32);
  }
  const bool top_a = has_a && rank_a < k, top_b = has_b && rank_b < k;
  // Categorical.__init__: logits - logsumexp(logits); .probs = softmax(that)
  float mx = fmaxf(top_a ? la : -INFINITY, top_b ? lb : -INFINITY);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  const float lse = logf(warp_sum((top_a ? expf(la - mx) : 0.f) + (top_b ? expf(lb - mx) : 0.f))) + mx;
  const float na = la - lse, nb = lb - lse;  // normalised logits
  float mx2 = fmaxf(top_a ? na : -INFINITY, top_b ? nb : -INFINITY);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx2 = fmaxf(mx2, __shfl_xor_sync(0xffffffffu, mx2, o));
  const float ea = top_a ? expf(na - mx2) : 0.f, eb = top_b ? expf(nb - mx2) : 0.f;
  const float denom = warp_sum(ea + eb);
  auto draw = [&](int rank) {
    if (noise) return noise[static_cast<long long>(row) * noise_stride + rank];
    // the counter uses the GLOBAL row (this shard's first row + row) so that a sharded run draws what one GPU would
    const uint4 r = philox4x32_10(make_uint4(static_cast<uint32_t>(iter), static_cast<uint32_t>(row + rng_row_offset),
                                             static_cast<uint32_t>(rank), 0x5eedu),
                                  make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32)));
    const float u = (static_cast<float>(r.x >> 8) + 0.5f) * (1.0f / 16777216.0f);  // (0,1)
    return -logf(u);
  };
  const float sa = top_a ? __fdiv_rn(__fdiv_rn(ea, denom), draw(rank_a)) : -INFINITY;
  const float sb = top_b ? __fdiv_rn(__fdiv_rn(eb, denom), draw(rank_b)) : -INFINITY;
  // argmax over slots; ties -> lowest rank (torch.argmax returns the first maximum)
  const int ra = top_a ? rank_a : 0x7fffffff, rb = top_b ? rank_b : 0x7fffffff;
  const bool pick_b = sb > sa || (sb == sa && rb < ra);
  float score = pick_b ? sb : sa;
  int best_rank = pick_b ? rb : ra;
  int best_id = pick_b ? id_b : id_a;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float os = __shfl_xor_sync(0xffffffffu, score, o);
    const int orank = __shfl_xor_sync(0xffffffffu, best_rank, o);
    const int oid = __shfl_xor_sync(0xffffffffu, best_id, o);
    if (os > score || (os == score && orank < best_rank)) { score = os; best_rank = orank; best_id = oid; }
  }
  return best_id;
}

// generate_step on precomputed logits rows (operator-level parity tests of the sampler tail).
__global__ void __launch_bounds__(256) sample_rows_kernel(const float* __restrict__ logits, int rows, int V,
                                                          const int32_t* __restrict__ valid_ids, int n, int k,
                                                          float temperature, const float* __restrict__ noise,
                                                          int noise_stride, unsigned long long seed,
                                                          int32_t* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* lr = logits + static_cast<long long>(row) * V;
  const float mine = lane < V ? lr[lane] : 0.f;
  const float extra = lane + 32 < V ? lr[lane + 32] : 0.f;
  const int id = generate_step_warp(mine, extra, lane, row, 0, valid_ids, n, k, temperature, noise, noise_stride, seed);
  if (lane == 0) out[row] = id;
}

// ---------------------------------------------------------------------------------------------
// Fused LM-head tail.  One warp per sampled row i (compact order: chain-major, then position slot):
//   y = LayerNorm(g[i])                      (RobertaLMHead.layer_norm, fp32)
//   logits[v] = <y, E[v]> + b[v], v < V      (tied embedding projection, fp32 on CUDA cores; V = 33)
//   generate_step (reference esm_sampler.py:8-45):
//       l = logits / temperature ; sub = l[valid_idx] ; k = (sample || top_k<=0 || top_k>n) ? n : top_k
//       (vals, idx) = topk(sub, k) ; draw ~ Categorical(logits=vals) == argmax_j softmax(vals)_j / q_j, q~Exp(1)
//       token = valid_idx[idx[draw]]
//   tokens[chain][pos] = token                (write-back, esm_sampler.py:234)
// E lives in shared memory (fp32, V*d*4 bytes) when it fits, else is read through L1/L2.
// ---------------------------------------------------------------------------------------------
struct HeadParams {
  const float* g;        // [rows, d] fp32: gelu(dense(LN_after(x))) for the sampled rows
  const float* ln_w;
  const float* ln_b;
  const float* emb;      // [V, d] tied projection weight
  const float* out_bias; // [V]
  float* logits_out;     // [rows, V] or nullptr
  int32_t* tokens;       // [n_chains, T] or nullptr (no sampling: forward_logits)
  int rows, d, V, T;
  float eps;
  Schedule sched;
  int iter;
  // sampler
  const int32_t* valid_ids;  // [n_valid] (device)
  int n_valid;
  int top_k;                 // effective k for this iteration (already resolved against burn-in), <= n_valid
  int top_k_raw;             // with sched.iter_dev: the caller's top_k and burn-in, resolved per iteration on the device
  long long burnin;
  float temperature;         // NaN : none (any other value divides the logits, as the reference does)
  const float* noise;        // replay: [rows, noise_stride] Exp(1) draws for this iteration, slot j <-> j-th largest
  int noise_stride;
  unsigned long long seed;   // device RNG otherwise
  long long rng_row_offset;  // global index of this engine's first sampled row (chains sharded across GPUs)
  int emb_in_smem;
  int no_ln;                 // project g as it is (ESM-1: logits = embed_out . x + bias, no LM-head LayerNorm)
  int skip_dup_writes;       // schedule may contain duplicate positions: last slot wins, like the reference loop
  // scoring (pseudo-log-likelihood, reference esm_sampler.py:288-363 / esm_msa_sampler.py:319-432)
  const int32_t* targets;    // [rows] true token of every scheduled slot (< 0: padding slot), or nullptr
  float* logp_out;           // [rows] log_softmax(logits)[target] over the whole vocabulary
};

// R rows per warp, VPL float4 per lane and row (sized to the model like layernorm_kernel).  One pass over the projection
// table serves the R rows of a warp: the table is read from shared memory once per ROW GROUP, not once per row (at
// R = 1 every sampled row streams V x d x 4 = 169 KB through the shared-memory port -- the kernel's bound).
template <int R, int VPL>
__global__ void __launch_bounds__(384) head_sample_kernel(HeadParams p) {
  extern __shared__ float s_emb[];
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  if (p.emb_in_smem) {
    const float4* src = reinterpret_cast<const float4*>(p.emb);
    float4* dst = reinterpret_cast<float4*>(s_emb);
    for (int i = threadIdx.x; i < (p.V * p.d) >> 2; i += blockDim.x) dst[i] = __ldg(src + i);
    __syncthreads();
  }
  // programmatic dependent launch: staging the (constant) projection table above overlaps the previous kernel
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const float* E = p.emb_in_smem ? s_emb : p.emb;
  const int nvec = p.d >> 2;
  for (int row0 = (blockIdx.x * warps_per_block + (threadIdx.x >> 5)) * R; row0 < p.rows;
       row0 += gridDim.x * warps_per_block * R) {
    float4 v[R][VPL];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const bool have = row0 + r < p.rows;
      const float4* in = reinterpret_cast<const float4*>(p.g + static_cast<long long>(have ? row0 + r : row0) * p.d);
      float sum = 0.f;
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        const int i = lane + k * 32;
        v[r][k] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < nvec) {
          v[r][k] = in[i];
          sum += v[r][k].x + v[r][k].y + v[r][k].z + v[r][k].w;
        }
      }
      const float mean = warp_sum(sum) / p.d;
      float sq = 0.f;
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        const int i = lane + k * 32;
        if (i < nvec) {
          const float4 a = v[r][k];
          sq += (a.x - mean) * (a.x - mean) + (a.y - mean) * (a.y - mean) + (a.z - mean) * (a.z - mean) +
                (a.w - mean) * (a.w - mean);
        }
      }
      const float rstd = 1.0f / sqrtf(warp_sum(sq) / p.d + p.eps);
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        const int i = lane + k * 32;
        if (i < nvec && !p.no_ln) {
          const float4 g = __ldg(reinterpret_cast<const float4*>(p.ln_w) + i);
          const float4 h = __ldg(reinterpret_cast<const float4*>(p.ln_b) + i);
          float4 a = v[r][k];
          a.x = (a.x - mean) * rstd * g.x + h.x; a.y = (a.y - mean) * rstd * g.y + h.y;
          a.z = (a.z - mean) * rstd * g.z + h.z; a.w = (a.w - mean) * rstd * g.w + h.w;
          v[r][k] = a;
        }
      }
    }
    // vocabulary projection: lane v ends up holding logit v (v < 32); logit 32.. kept in `extra` on lane v-32.
    // Four tokens at a time: their dot products and the butterfly reductions are independent chains, which is
    // what hides the shared-memory and shuffle latency with only 12 warps per SM (per-token arithmetic unchanged:
    // the same products are accumulated in the same order whatever R is).
    float mine[R], extra[R];
#pragma unroll
    for (int r = 0; r < R; ++r) mine[r] = extra[r] = 0.f;
    for (int tok0 = 0; tok0 < p.V; tok0 += 4) {
      float acc[R][4];
#pragma unroll
      for (int r = 0; r < R; ++r) acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.f;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int tok = tok0 + u < p.V ? tok0 + u : p.V - 1;   // clamp: the surplus lanes of the last group are ignored
        const float4* e = reinterpret_cast<const float4*>(E + static_cast<long long>(tok) * p.d);
#pragma unroll
        for (int k = 0; k < VPL; ++k) {
          const int i = lane + k * 32;
          if (i < nvec) {
            const float4 w = e[i];
#pragma unroll
            for (int r = 0; r < R; ++r) {
              acc[r][u] = fmaf(v[r][k].x, w.x, acc[r][u]); acc[r][u] = fmaf(v[r][k].y, w.y, acc[r][u]);
              acc[r][u] = fmaf(v[r][k].z, w.z, acc[r][u]); acc[r][u] = fmaf(v[r][k].w, w.w, acc[r][u]);
            }
          }
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
          for (int u = 0; u < 4; ++u) acc[r][u] += __shfl_xor_sync(0xffffffffu, acc[r][u], o);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int tok = tok0 + u;
        if (tok < p.V) {
          const float bias = __ldg(p.out_bias + tok);
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const float a = acc[r][u] + bias;
            if ((tok & 31) == lane) { if (tok < 32) mine[r] = a; else extra[r] = a; }
          }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int row = row0 + r;
      if (row >= p.rows) break;
      const float mine_r = mine[r], extra_r = extra[r];
      if (p.logits_out) {
        float* lo = p.logits_out + static_cast<long long>(row) * p.V;
        if (lane < p.V) lo[lane] = mine_r;
        if (lane + 32 < p.V) lo[lane + 32] = extra_r;
      }
      if (p.logp_out) {
        const float a = lane < p.V ? mine_r : -INFINITY, b = lane + 32 < p.V ? extra_r : -INFINITY;
        float mx = fmaxf(a, b);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        const float lse = logf(warp_sum(expf(a - mx) + expf(b - mx))) + mx;
        const int tgt = __ldg(p.targets + row);
        const float lt = __shfl_sync(0xffffffffu, tgt >= 32 ? extra_r : mine_r, tgt & 31);
        if (lane == 0) p.logp_out[row] = tgt >= 0 && tgt < p.V ? lt - lse : 0.f;
      }
      if (!p.tokens) continue;

      // device-resident iteration (graph replay): the effective k and the noise slice follow from it
      int iter = p.iter, top_k = p.top_k;
      const float* noise = p.noise;
      if (p.sched.iter_dev) {
        iter = *p.sched.iter_dev;
        top_k = (iter < p.burnin || p.top_k_raw <= 0 || p.top_k_raw > p.n_valid) ? p.n_valid : p.top_k_raw;
        if (noise) noise += static_cast<long long>(iter) * p.rows * p.noise_stride;
      }
      const int best_id = generate_step_warp(mine_r, extra_r, lane, row, iter, p.valid_ids, p.n_valid, top_k,
                                             p.temperature, noise, p.noise_stride, p.seed, p.rng_row_offset);
      if (lane == 0) {
        const int chain = row / p.sched.P, slot = row % p.sched.P;
        const int32_t* plist = p.sched.positions + iter * p.sched.iter_stride + chain * p.sched.chain_stride;
        const int pos = plist[slot];
        bool write = true;
        if (p.skip_dup_writes) {
          for (int s2 = slot + 1; s2 < p.sched.P; ++s2) if (plist[s2] == pos) { write = false; break; }
        }
        if (write) p.tokens[(static_cast<long long>(chain) * p.sched.seq_stride + p.sched.seq_offset) * p.T + pos] = best_id;
      }
    }
  }
}

}  // namespace pg

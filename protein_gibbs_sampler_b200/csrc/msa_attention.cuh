// MSA Transformer tied row attention (fair-esm esm/axial_attention.py RowSelfAttention; reference call site
// /root/reference/src/pgen/esm_msa_sampler.py:136,236).  One attention map per (MSA, head) is shared by all
// R rows:
//     s[b,h,i,j] = sum_r sum_d q[b,r,i,h,d] k[b,r,j,h,d]        (q pre-scaled by Dh^-1/2 / sqrt(R) in the QKV GEMM)
//     p = softmax_j(s) ;  ctx[b,r,i,h,:] = sum_j p[b,h,i,j] v[b,r,j,h,:]
// Two kernels: (1) score tiles accumulated over the R rows (a [C, R*Dh] x [R*Dh, C] contraction) written as
// fp32, (2) softmax of 64 query rows into shared memory (fp16) followed by P.V for a group of rows.
// Column attention reuses the flash kernel in attention.cuh with a strided token mapping.
// Activation row of (b, r, c) is (b*R + r)*C + c; q|k|v are the three d-wide column blocks of `qkv`.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "attention.cuh"

namespace pg {

// ---------------------------------------------------------------------------------- (1) tied scores
template <int DH>
__global__ void __launch_bounds__(128) msa_row_scores_kernel(const __half* __restrict__ qkv, float* __restrict__ scores,
                                                             int R, int C, int H, int ld, int k_off) {
  // programmatic dependent launch (no-ops in a normal launch): wait for the previous kernel, release the next one
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  constexpr int LDS = DH + 8, CH = DH / 8, KS = DH / 16;
  __shared__ __align__(16) __half sQ[2][64 * LDS];
  __shared__ __align__(16) __half sK[2][64 * LDS];
  const int i0 = blockIdx.x * 64, j0 = blockIdx.y * 64;
  const int b = blockIdx.z / H, head = blockIdx.z % H;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;

  auto load = [&](int r, int buf) {
    const __half* base = qkv + (static_cast<long long>(b) * R + r) * C * ld + head * DH;
    for (int i = tid; i < 64 * CH; i += 128) {
      const int row = i / CH, c = i % CH;
      const bool qi = i0 + row < C, kj = j0 + row < C;
      cp_async16(&sQ[buf][row * LDS + c * 8], base + static_cast<long long>(qi ? i0 + row : 0) * ld + c * 8, qi);
      cp_async16(&sK[buf][row * LDS + c * 8], base + static_cast<long long>(kj ? j0 + row : 0) * ld + k_off + c * 8, kj);
    }
  };
  float s[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) { s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f; }
  load(0, 0);
  cp_async_commit();
  for (int r = 0; r < R; ++r) {
    const int buf = r & 1;
    if (r + 1 < R) load(r + 1, buf ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    uint32_t qf[KS][4];
#pragma unroll
    for (int ks = 0; ks < KS; ++ks)
      ldsm_x4(qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3],
              &sQ[buf][(warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LDS + ks * 16 + (lane >> 4) * 8]);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if constexpr (KS >= 2) {
#pragma unroll
        for (int kp = 0; kp < KS / 2; ++kp) {
          uint32_t b0, b1, b2, b3;
          ldsm_x4(b0, b1, b2, b3, &sK[buf][(j * 8 + (lane & 7)) * LDS + kp * 32 + (lane >> 3) * 8]);
          mma_16816(s[j], qf[2 * kp], b0, b1);
          mma_16816(s[j], qf[2 * kp + 1], b2, b3);
        }
      } else {
        uint32_t b0, b1;
        ldsm_x2(b0, b1, &sK[buf][(j * 8 + (lane & 7)) * LDS + ((lane >> 3) & 1) * 8]);
        mma_16816(s[j], qf[0], b0, b1);
      }
    }
    __syncthreads();
  }
  cp_async_wait<0>();
  float* out = scores + static_cast<long long>(blockIdx.z) * C * C;
  const int r0 = i0 + warp * 16 + g, r1 = r0 + 8;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = j0 + j * 8 + 2 * t4;
    if (r0 < C) {
      if (c < C) out[static_cast<long long>(r0) * C + c] = s[j][0];
      if (c + 1 < C) out[static_cast<long long>(r0) * C + c + 1] = s[j][1];
    }
    if (r1 < C) {
      if (c < C) out[static_cast<long long>(r1) * C + c] = s[j][2];
      if (c + 1 < C) out[static_cast<long long>(r1) * C + c + 1] = s[j][3];
    }
  }
}

// ------------------------------------------------------------------------- (2) softmax + P.V per row
// grid: (ceil(C/64) query tiles, row groups, B*H).  Dynamic smem: P[64][Cpad+8] fp16 + 2 V buffers.
template <int DH>
__global__ void __launch_bounds__(128) msa_row_pv_kernel(const __half* __restrict__ qkv, const float* __restrict__ scores,
                                                         __half* __restrict__ ctx, int R, int C, int H, int ld,
                                                         int ldc, int v_off, int rows_per_group, int Cpad) {
  // programmatic dependent launch (no-ops in a normal launch): wait for the previous kernel, release the next one
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  constexpr int LDS = DH + 8, CH = DH / 8, NT = DH / 8;
  extern __shared__ __align__(16) unsigned char msa_smem[];
  const int ldp = Cpad + 8;
  __half* sP = reinterpret_cast<__half*>(msa_smem);
  __half* sV = sP + 64 * ldp;  // [2][64*LDS]
  const int i0 = blockIdx.x * 64;
  const int b = blockIdx.z / H, head = blockIdx.z % H;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;

  // softmax of this warp's 16 query rows (fp32), stored as fp16 probabilities; columns >= C are zero
  const float* sc = scores + static_cast<long long>(blockIdx.z) * C * C;
  for (int rr = 0; rr < 16; ++rr) {
    const int row = warp * 16 + rr, gi = i0 + row;
    __half* prow = sP + row * ldp;
    if (gi >= C) {
      for (int j = lane; j < Cpad; j += 32) prow[j] = __float2half(0.f);
      continue;
    }
    const float* srow = sc + static_cast<long long>(gi) * C;
    float mx = -INFINITY;
    for (int j = lane; j < C; j += 32) mx = fmaxf(mx, srow[j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    for (int j = lane; j < C; j += 32) sum += expf(srow[j] - mx);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = 1.0f / sum;
    for (int j = lane; j < Cpad; j += 32) prow[j] = __float2half_rn(j < C ? expf(srow[j] - mx) * inv : 0.f);
  }
  __syncthreads();

  const int n_chunks = Cpad / 64;
  const int r_begin = blockIdx.y * rows_per_group, r_end = min(R, r_begin + rows_per_group);
  const int total = (r_end - r_begin) * n_chunks;  // flattened (row, key chunk) pipeline
  auto load_v = [&](int it, int buf) {
    const int r = r_begin + it / n_chunks, k0 = (it % n_chunks) * 64;
    const __half* base = qkv + (static_cast<long long>(b) * R + r) * C * ld + v_off + head * DH;
    for (int i = tid; i < 64 * CH; i += 128) {
      const int row = i / CH, c = i % CH;
      const bool ok = k0 + row < C;
      cp_async16(&sV[buf * 64 * LDS + row * LDS + c * 8], base + static_cast<long long>(ok ? k0 + row : 0) * ld + c * 8, ok);
    }
  };
  if (total > 0) load_v(0, 0);
  cp_async_commit();
  float o[NT][4];
  for (int it = 0; it < total; ++it) {
    const int buf = it & 1, chunk = it % n_chunks;
    if (it + 1 < total) load_v(it + 1, buf ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    if (chunk == 0) {
#pragma unroll
      for (int j = 0; j < NT; ++j) { o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f; }
    }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t pf[4];
      ldsm_x4(pf[0], pf[1], pf[2], pf[3],
              &sP[(warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * ldp + chunk * 64 + kk * 16 + (lane >> 4) * 8]);
#pragma unroll
      for (int np = 0; np < NT / 2; ++np) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4_t(b0, b1, b2, b3,
                  &sV[buf * 64 * LDS + (kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LDS + np * 16 + (lane >> 4) * 8]);
        mma_16816(o[2 * np], pf, b0, b1);
        mma_16816(o[2 * np + 1], pf, b2, b3);
      }
    }
    if (chunk == n_chunks - 1) {
      const int r = r_begin + it / n_chunks;
      __half* out = ctx + (static_cast<long long>(b) * R + r) * C * ldc + head * DH;
      const int q0 = i0 + warp * 16 + g, q1 = q0 + 8;
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        const int c = j * 8 + 2 * t4;
        if (q0 < C) *reinterpret_cast<uint32_t*>(out + static_cast<long long>(q0) * ldc + c) = pack2(o[j][0], o[j][1]);
        if (q1 < C) *reinterpret_cast<uint32_t*>(out + static_cast<long long>(q1) * ldc + c) = pack2(o[j][2], o[j][3]);
      }
    }
    __syncthreads();
  }
  cp_async_wait<0>();
}

// ---------------------------------------------------------------------------- column attention, R <= 32
// MSA column attention (fair-esm ColumnSelfAttention): for every alignment column c of MSA b the R rows attend to
// each other, per head.  The generic flash kernel (attention.cuh) runs it as B*C*H tiny CTAs that each read 128-byte
// row fragments 4.6 KB apart; here one CTA takes HG adjacent heads of one column, so every row contributes HG*128
// contiguous bytes of q, of k and of v, and the whole R x R problem of a head sits in two warps' registers (no key
// loop: R <= 32).  Token r of group s = (b, c) lives in activation row (b*R + r)*C + c.
template <int HG>
__global__ void __launch_bounds__(64 * HG) msa_col_attention_kernel(AttnParams p) {
  // programmatic dependent launch (no-ops in a normal launch): wait for the previous kernel, release the next one
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  constexpr int DH = 64, LDS = DH + 8, KS = DH / 16, NT = DH / 8;
  extern __shared__ __align__(16) unsigned char col_smem[];
  __half* sQ = reinterpret_cast<__half*>(col_smem);  // [HG][32 * LDS]
  __half* sK = sQ + HG * 32 * LDS;
  __half* sV = sK + HG * 32 * LDS;
  const int s = p.reverse ? gridDim.x - 1 - blockIdx.x : blockIdx.x, head0 = blockIdx.y * HG;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
  const long long row0 = (s / p.inner) * p.outer_stride + static_cast<long long>(s % p.inner) * p.inner_stride;
  const long long rs = static_cast<long long>(p.row_step) * p.ld;
  const __half* base = p.qkv + row0 * p.ld + head0 * DH;
  for (int i = tid; i < 32 * HG * 8; i += 64 * HG) {
    const int r = i / (HG * 8), ch = i % (HG * 8), hl = ch >> 3, c8 = ch & 7;
    const bool ok = r < p.T;
    const __half* src = base + static_cast<long long>(ok ? r : 0) * rs + hl * DH + c8 * 8;
    const int dst = hl * 32 * LDS + r * LDS + c8 * 8;
    cp_async16(&sQ[dst], src, ok);
    cp_async16(&sK[dst], src + p.k_off, ok);
    cp_async16(&sV[dst], src + p.v_off, ok);
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();

  const int hl = warp >> 1, q0 = (warp & 1) * 16;  // this warp: head head0 + hl, query rows q0 .. q0 + 15
  if (q0 >= p.T) return;
  const __half* q = sQ + hl * 32 * LDS;
  const __half* k = sK + hl * 32 * LDS;
  const __half* v = sV + hl * 32 * LDS;
  uint32_t qf[KS][4];
#pragma unroll
  for (int ks = 0; ks < KS; ++ks)
    ldsm_x4(qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3],
            &q[(q0 + (lane & 7) + ((lane >> 3) & 1) * 8) * LDS + ks * 16 + (lane >> 4) * 8]);
  float sc[4][4];  // 32 keys = 4 n-tiles
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    sc[j][0] = sc[j][1] = sc[j][2] = sc[j][3] = 0.f;
#pragma unroll
    for (int kp = 0; kp < KS / 2; ++kp) {
      uint32_t b0, b1, b2, b3;
      ldsm_x4(b0, b1, b2, b3, &k[(j * 8 + (lane & 7)) * LDS + kp * 32 + (lane >> 3) * 8]);
      mma_16816(sc[j], qf[2 * kp], b0, b1);
      mma_16816(sc[j], qf[2 * kp + 1], b2, b3);
    }
  }
  float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = j * 8 + 2 * t4;
    if (c >= p.T) { sc[j][0] = -INFINITY; sc[j][2] = -INFINITY; }
    if (c + 1 >= p.T) { sc[j][1] = -INFINITY; sc[j][3] = -INFINITY; }
    mx0 = fmaxf(mx0, fmaxf(sc[j][0], sc[j][1]));
    mx1 = fmaxf(mx1, fmaxf(sc[j][2], sc[j][3]));
  }
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
  constexpr float kLog2e = 1.4426950408889634f;
  float l0 = 0.f, l1 = 0.f;
  uint32_t pf[2][4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float p0 = exp2f((sc[j][0] - mx0) * kLog2e), p1 = exp2f((sc[j][1] - mx0) * kLog2e);
    const float p2 = exp2f((sc[j][2] - mx1) * kLog2e), p3 = exp2f((sc[j][3] - mx1) * kLog2e);
    l0 += p0 + p1; l1 += p2 + p3;
    pf[j >> 1][(j & 1) * 2 + 0] = pack2(p0, p1);
    pf[j >> 1][(j & 1) * 2 + 1] = pack2(p2, p3);
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  float o[NT][4];
#pragma unroll
  for (int j = 0; j < NT; ++j) { o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f; }
#pragma unroll
  for (int kk = 0; kk < 2; ++kk) {
#pragma unroll
    for (int np = 0; np < NT / 2; ++np) {
      uint32_t b0, b1, b2, b3;
      ldsm_x4_t(b0, b1, b2, b3, &v[(kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LDS + np * 16 + (lane >> 4) * 8]);
      mma_16816(o[2 * np], pf[kk], b0, b1);
      mma_16816(o[2 * np + 1], pf[kk], b2, b3);
    }
  }
  const float i0 = 1.0f / l0, i1 = 1.0f / l1;
  const int r0 = q0 + g, r1 = r0 + 8;
  __half* out = p.ctx + row0 * p.ldc + (head0 + hl) * DH;
  const long long os = static_cast<long long>(p.row_step) * p.ldc;
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    const int c = j * 8 + 2 * t4;
    if (r0 < p.T) *reinterpret_cast<uint32_t*>(out + static_cast<long long>(r0) * os + c) = pack2(o[j][0] * i0, o[j][1] * i0);
    if (r1 < p.T) *reinterpret_cast<uint32_t*>(out + static_cast<long long>(r1) * os + c) = pack2(o[j][2] * i1, o[j][3] * i1);
  }
}

// Column attention for head_dim 64, at most 32 rows and a head count divisible by 4 or 2; nullptr on success,
// "" when the shape is not covered (the caller falls back to the generic kernel), else an error string.
static const char* launch_msa_col_attention(const AttnParams& p, int groups, int H, cudaStream_t st) {
  if (p.T > 32) return "";
  const int hg = H % 4 == 0 ? 4 : (H % 2 == 0 ? 2 : 0);
  if (!hg) return "";
  const size_t smem = static_cast<size_t>(3) * hg * 32 * (64 + 8) * sizeof(__half);
  if (ensure_dynamic_smem(msa_col_attention_kernel<4>, 3 * 4 * 32 * 72 * 2) != cudaSuccess)
    return "cudaFuncSetAttribute(msa_col_attention_kernel) failed";
  dim3 grid(groups, H / hg);
  if (hg == 4) msa_col_attention_kernel<4><<<grid, 256, smem, st>>>(p);
  else msa_col_attention_kernel<2><<<grid, 128, smem, st>>>(p);
  if (cudaGetLastError() != cudaSuccess) return "msa_col_attention_kernel launch failed";
  return nullptr;
}

// return nullptr on success, else a static error string
template <int DH>
static const char* launch_msa_row_attention_t(const __half* qkv, __half* ctx, float* scores, int B, int R, int C,
                                              int H, cudaStream_t st, int ldc) {
  const int d = H * DH, ld = 3 * d;
  const int tiles = (C + 63) / 64;
  msa_row_scores_kernel<DH><<<dim3(tiles, tiles, B * H), 128, 0, st>>>(qkv, scores, R, C, H, ld, d);
  if (cudaGetLastError() != cudaSuccess) return "msa_row_scores_kernel launch failed";
  const int Cpad = tiles * 64;
  const size_t smem = (static_cast<size_t>(64) * (Cpad + 8) + 2 * 64 * (DH + 8)) * sizeof(__half);
  if (smem > 227 * 1024) return "MSA too wide for the row-attention kernel's shared-memory softmax tile";
  if (ensure_dynamic_smem(msa_row_pv_kernel<DH>, static_cast<int>(smem)) != cudaSuccess)
    return "cudaFuncSetAttribute(msa_row_pv_kernel) failed";
  const int rows_per_group = R >= 8 ? 4 : 1;
  const int groups = (R + rows_per_group - 1) / rows_per_group;
  msa_row_pv_kernel<DH><<<dim3(tiles, groups, B * H), 128, smem, st>>>(qkv, scores, ctx, R, C, H, ld, ldc, 2 * d,
                                                                         rows_per_group, Cpad);
  if (cudaGetLastError() != cudaSuccess) return "msa_row_pv_kernel launch failed";
  return nullptr;
}

static const char* launch_msa_row_attention(const __half* qkv, __half* ctx, float* scores, int B, int R, int C, int H,
                                            int hd, cudaStream_t st, int ldc = 0) {
  if (!ldc) ldc = H * hd;   // ctx row pitch ([hi | lo] rows in split-operand mode)
  switch (hd) {
    case 16: return launch_msa_row_attention_t<16>(qkv, ctx, scores, B, R, C, H, st, ldc);
    case 32: return launch_msa_row_attention_t<32>(qkv, ctx, scores, B, R, C, H, st, ldc);
    case 64: return launch_msa_row_attention_t<64>(qkv, ctx, scores, B, R, C, H, st, ldc);
  }
  return "unsupported head_dim for MSA row attention (16, 32, 64)";
}

}  // namespace pg

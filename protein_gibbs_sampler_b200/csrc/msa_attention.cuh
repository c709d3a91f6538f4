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

// return nullptr on success, else a static error string
template <int DH>
static const char* launch_msa_row_attention_t(const __half* qkv, __half* ctx, float* scores, int B, int R, int C,
                                              int H, cudaStream_t st) {
  const int d = H * DH, ld = 3 * d;
  const int tiles = (C + 63) / 64;
  msa_row_scores_kernel<DH><<<dim3(tiles, tiles, B * H), 128, 0, st>>>(qkv, scores, R, C, H, ld, d);
  if (cudaGetLastError() != cudaSuccess) return "msa_row_scores_kernel launch failed";
  const int Cpad = tiles * 64;
  const size_t smem = (static_cast<size_t>(64) * (Cpad + 8) + 2 * 64 * (DH + 8)) * sizeof(__half);
  if (smem > 227 * 1024) return "MSA too wide for the row-attention kernel's shared-memory softmax tile";
  static size_t configured = 0;
  if (smem > configured) {
    if (cudaFuncSetAttribute(msa_row_pv_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             static_cast<int>(smem)) != cudaSuccess)
      return "cudaFuncSetAttribute(msa_row_pv_kernel) failed";
    configured = smem;
  }
  const int rows_per_group = R >= 8 ? 4 : 1;
  const int groups = (R + rows_per_group - 1) / rows_per_group;
  msa_row_pv_kernel<DH><<<dim3(tiles, groups, B * H), 128, smem, st>>>(qkv, scores, ctx, R, C, H, ld, d, 2 * d,
                                                                         rows_per_group, Cpad);
  if (cudaGetLastError() != cudaSuccess) return "msa_row_pv_kernel launch failed";
  return nullptr;
}

static const char* launch_msa_row_attention(const __half* qkv, __half* ctx, float* scores, int B, int R, int C, int H,
                                            int hd, cudaStream_t st) {
  switch (hd) {
    case 16: return launch_msa_row_attention_t<16>(qkv, ctx, scores, B, R, C, H, st);
    case 32: return launch_msa_row_attention_t<32>(qkv, ctx, scores, B, R, C, H, st);
    case 64: return launch_msa_row_attention_t<64>(qkv, ctx, scores, B, R, C, H, st);
  }
  return "unsupported head_dim for MSA row attention (16, 32, 64)";
}

}  // namespace pg

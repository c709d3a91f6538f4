// Self-attention over the fused QKV activation (fp16) for the single-sequence models (ESM-1b / ESM-2).
//
//   ctx[s, i, h, :] = softmax_j( q[s,i,h,:] . k[s,j,h,:] ) v[s,j,h,:]        (q is pre-scaled by Dh^-1/2,
//                                                                             RoPE already applied by the QKV GEMM)
// Flash-style: one CTA = 16*NWARPS queries of one (sequence, head); keys/values stream through shared memory in
// 64-row chunks (cp.async double buffer); scores and the running softmax stay in registers in fp32.
// No padding mask: the Gibbs path never contains <pad>.
// Replaces fair-esm MultiheadAttention's bmm/softmax/bmm (call site /root/reference/src/pgen/esm_sampler.py:223).
//
// Warp-level mma.sync.m16n8k16 path.  The production path for head_dim 64 sequence attention is the tcgen05 kernel in
// attention_fa.cuh; this kernel serves head_dim 16 / 32 (ESM-2 8M / 150M), MSA column attention (strided token
// mapping, at most a few dozen keys per group) and 9..16 trailing query rows left over by the tcgen05 kernel.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace pg {

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  const int sz = valid ? 16 : 0;  // src-size 0 -> zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, const void* p) {
  const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, const void* p) {
  const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(a));
}
__device__ __forceinline__ void ldsm_x2(uint32_t& r0, uint32_t& r1, const void* p) {
  const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(a));
}
__device__ __forceinline__ void ldsm_x2_t(uint32_t& r0, uint32_t& r1, const void* p) {
  const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(a));
}
__device__ __forceinline__ void mma_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// Token t of attention group s lives in activation row
//     (s / inner) * outer_stride + (s % inner) * inner_stride + t * row_step.
// Sequence attention: inner=1, outer_stride=T, row_step=1.  MSA column attention (attend over the R rows of
// one alignment column c of MSA b): s = b*C + c, inner=C, outer_stride=R*C, inner_stride=1, row_step=C, T=R.
struct AttnParams {
  const __half* qkv;  // [rows, ld]; q at col 0, k at col k_off, v at col v_off; head h at +h*DH
  __half* ctx;        // [rows, ldc]
  int T, ld, ldc, k_off, v_off;
  int inner, inner_stride, row_step;
  long long outer_stride;
  int q_begin = 0;    // first query row handled by this launch (blockIdx.x counts tiles from here)
  int reverse = 0;    // msa_col_attention_kernel: walk the columns from the last one down (pgibbs_engine::zigzag)
};

constexpr int kAttnBK = 64;   // keys per smem chunk

template <int DH, int NWARPS>
__global__ void __launch_bounds__(NWARPS * 32) attention_kernel(AttnParams p) {
  // programmatic dependent launch (no-ops in a normal launch): wait for the previous kernel, release the next one
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  constexpr int kAttnBQ = 16 * NWARPS;  // queries per CTA
  constexpr int NT_ = NWARPS * 32;
  constexpr int LDS = DH + 8;          // padded row (halfs): 16-byte pad keeps ldmatrix conflict-free
  constexpr int CH = DH / 8;           // 16-byte chunks per row
  constexpr int KS = DH / 16;          // k-steps of QK^T
  constexpr int NT = DH / 8;           // n-tiles of the output
  __shared__ __align__(16) __half sQ[kAttnBQ * LDS];
  __shared__ __align__(16) __half sK[2][kAttnBK * LDS];
  __shared__ __align__(16) __half sV[2][kAttnBK * LDS];

  const int seq = blockIdx.z, head = blockIdx.y, q0 = p.q_begin + blockIdx.x * kAttnBQ;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const long long row0 = (seq / p.inner) * p.outer_stride + static_cast<long long>(seq % p.inner) * p.inner_stride;
  const long long rs = static_cast<long long>(p.row_step) * p.ld;  // elements between consecutive tokens
  const __half* base = p.qkv + row0 * p.ld + head * DH;
  const int n_chunks = (p.T + kAttnBK - 1) / kAttnBK;

  // Q tile
  for (int i = tid; i < kAttnBQ * CH; i += NT_) {
    const int r = i / CH, c = i % CH;
    const bool ok = q0 + r < p.T;
    cp_async16(&sQ[r * LDS + c * 8], base + static_cast<long long>(ok ? q0 + r : 0) * rs + c * 8, ok);
  }
  auto load_kv = [&](int chunk, int buf) {
    const int k0 = chunk * kAttnBK;
    for (int i = tid; i < kAttnBK * CH; i += NT_) {
      const int r = i / CH, c = i % CH;
      const bool ok = k0 + r < p.T;
      const __half* src = base + static_cast<long long>(ok ? k0 + r : 0) * rs + c * 8;
      cp_async16(&sK[buf][r * LDS + c * 8], src + p.k_off, ok);
      cp_async16(&sV[buf][r * LDS + c * 8], src + p.v_off, ok);
    }
  };
  load_kv(0, 0);
  cp_async_commit();

  uint32_t qf[KS][4];
  float o[NT][4];
#pragma unroll
  for (int j = 0; j < NT; ++j) { o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f; }
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  constexpr float kLog2e = 1.4426950408889634f;

  for (int chunk = 0; chunk < n_chunks; ++chunk) {
    const int buf = chunk & 1;
    if (chunk + 1 < n_chunks) load_kv(chunk + 1, buf ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    if (chunk == 0) {
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        const __half* a = &sQ[(warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LDS + ks * 16 + (lane >> 4) * 8];
        ldsm_x4(qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3], a);
      }
    }
    // S = Q K^T for 64 keys: 8 n-tiles
    float s[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) { s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f; }
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
    // mask keys beyond the sequence
    const int kbase = chunk * kAttnBK;
    if (kbase + kAttnBK > p.T) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = kbase + j * 8 + 2 * t4;
        if (c >= p.T) { s[j][0] = -INFINITY; s[j][2] = -INFINITY; }
        if (c + 1 >= p.T) { s[j][1] = -INFINITY; s[j][3] = -INFINITY; }
      }
    }
    // online softmax (rows g and g+8 of this warp's 16)
    float mx0 = m0, mx1 = m1;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
      mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float sc0 = exp2f((m0 - mx0) * kLog2e), sc1 = exp2f((m1 - mx1) * kLog2e);
    m0 = mx0; m1 = mx1;
    float rs0 = 0.f, rs1 = 0.f;
    uint32_t pf[4][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float p0 = exp2f((s[j][0] - mx0) * kLog2e), p1 = exp2f((s[j][1] - mx0) * kLog2e);
      const float p2 = exp2f((s[j][2] - mx1) * kLog2e), p3 = exp2f((s[j][3] - mx1) * kLog2e);
      rs0 += p0 + p1; rs1 += p2 + p3;
      pf[j >> 1][(j & 1) * 2 + 0] = pack2(p0, p1);
      pf[j >> 1][(j & 1) * 2 + 1] = pack2(p2, p3);
    }
    l0 = l0 * sc0 + rs0; l1 = l1 * sc1 + rs1;
#pragma unroll
    for (int j = 0; j < NT; ++j) { o[j][0] *= sc0; o[j][1] *= sc0; o[j][2] *= sc1; o[j][3] *= sc1; }
    // O += P V
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int np = 0; np < NT / 2; ++np) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4_t(b0, b1, b2, b3,
                  &sV[buf][(kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LDS + np * 16 + (lane >> 4) * 8]);
        mma_16816(o[2 * np], pf[kk], b0, b1);
        mma_16816(o[2 * np + 1], pf[kk], b2, b3);
      }
    }
    __syncthreads();  // all warps done with buf before it is refilled two iterations later
  }
  cp_async_wait<0>();
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.0f / l0, i1 = 1.0f / l1;
  const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;
  __half* out = p.ctx + row0 * p.ldc + head * DH;
  const long long os = static_cast<long long>(p.row_step) * p.ldc;
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    const int c = j * 8 + 2 * t4;
    if (r0 < p.T) *reinterpret_cast<uint32_t*>(out + static_cast<long long>(r0) * os + c) = pack2(o[j][0] * i0, o[j][1] * i0);
    if (r1 < p.T) *reinterpret_cast<uint32_t*>(out + static_cast<long long>(r1) * os + c) = pack2(o[j][2] * i1, o[j][3] * i1);
  }
}

}  // namespace pg

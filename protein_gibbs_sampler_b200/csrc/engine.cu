// pgibbs engine: host-side orchestration of the on-device Gibbs step + the C ABI (include/pgibbs.h).
// One engine per GPU.  All work is queued on one CUDA stream (plus a private side stream that run_gemm_resid_ln forks
// to and joins back from inside a layer); nothing between iterations touches the host.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/pgibbs.h"
#include "attention.cuh"
#include "attention_fa.cuh"
#include "gemm.cuh"
#include "msa_attention.cuh"
#include "msa_row_tc.cuh"
#include "rowwise.cuh"

namespace pg {

// ------------------------------------------------------------------------------------------ errors
static thread_local char g_err[1024] = "";
static int fail(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return 1;
}
#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t _e = (call);                                                                       \
    if (_e != cudaSuccess) return fail("%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)
#define TRY(call)               \
  do {                          \
    int _r = (call);            \
    if (_r) return _r;          \
  } while (0)

// ------------------------------------------------------------------------- tensor maps (driver API)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// fp16 row-major [rows, cols] (ld elements between rows); box = 64 columns x box_rows, 128B swizzle.
static int make_tmap_2d(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld,
                        uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail("cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * sizeof(__half)};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled failed with CUresult %d (rows=%llu cols=%llu ld=%llu box=%u)",
                                     (int)r, (unsigned long long)rows, (unsigned long long)cols,
                                     (unsigned long long)ld, box_rows);
  return 0;
}

// GEMM output [rows, cols] (ld elements between rows), fp16 or fp32: box = 128 bytes x 32 rows, 128B swizzle
// (one epilogue warp's chunk).
static int make_tmap_out(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, bool f16) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail("cuTensorMapEncodeTiled entry point not available");
  const uint64_t esz = f16 ? 2 : 4;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * esz};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(128 / esz), 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                  const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled (output) failed with CUresult %d (rows=%llu cols=%llu ld=%llu)",
                                     (int)r, (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld);
  return 0;
}

// Fused qkv activation viewed as [n_seq][T][3d] fp16 for the tcgen05 attention kernel: box = 64 columns (one head of
// q, k or v) x 128 tokens of one sequence, 128B swizzle; tokens t >= T are zero-filled (never the next sequence).
static int make_tmap_qkv3(CUtensorMap* m, const void* ptr, uint64_t n_seq, uint64_t T, uint64_t ld,
                          uint32_t box_rows = 128, uint64_t cols = 0) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail("cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[3] = {cols ? cols : ld, T, n_seq};   // cols < ld: only the hi half of a [hi | lo] row is addressable
  cuuint64_t strides[2] = {ld * sizeof(__half), T * ld * sizeof(__half)};
  cuuint32_t box[3] = {64, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled (qkv 3-D) failed with CUresult %d", (int)r);
  return 0;
}

// ------------------------------------------------------------------------------------- GEMM launch
// SM count of the CURRENT device (cached per device: engines on different GPUs may share a process)
static int num_sms() {
  static int cache[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev < 0 || dev >= 64) { int n = 148; cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); return n; }
  if (!cache[dev]) cudaDeviceGetAttribute(&cache[dev], cudaDevAttrMultiProcessorCount, dev);
  return cache[dev] ? cache[dev] : 148;
}

// A GEMM's tiling: tile width and whether a CTA pair (cta_group::2, 256-row tiles) computes it.
// The B tensor map's box holds bn / cg rows.
constexpr int kSplitFlagInts = 4096;  // K-split flags: >= 148 groups x 2 ranks x 8 epilogue warps
static int g_gemm_split = 0;              // PGIBBS_GEMM_SPLIT=1: last-wave K-split of the residual GEMMs (off: it makes a chain's low-order bits depend on the batch it runs in)
static int g_epi_direct = 6;          // PGIBBS_EPI_DIRECT bit mask: plain-store GEMM epilogues write straight from registers (st.global.v8)
                                      // instead of through shared-memory staging and the TMA engine -- 2: fp16 outputs, 4: fp32 outputs; 0: all TMA
static int g_tail_overlap = 1;        // PGIBBS_TAIL_OVERLAP=0: no fork -- by default the LayerNorm of the finished row blocks runs next to a residual GEMM's partly-filled last wave
static int g_pdl = 1;                 // PGIBBS_PDL=0: plain stream-ordered launches
static int g_graph = 1;               // PGIBBS_GRAPH=0: every iteration is launched kernel by kernel
static int g_zigzag = 1;              // PGIBBS_ZIGZAG=0: every kernel walks its rows in ascending order
constexpr int kGraphMinIters = 16;    // shorter runs do not pay for capture + instantiation (~1 ms)

static thread_local bool t_no_pdl = false;   // set around a launch whose predecessor is an event of another stream
// Launch with programmatic dependent launch allowed: the kernel must call pdl_wait() before it touches global memory.
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = (g_pdl && !t_no_pdl) ? 1 : 0;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}          // PGIBBS_GEMM_SPLIT=0 turns the last-wave K-split off (A/B measurements)

struct GemmPlan {
  int bn = 256, cg = 2;
  int b_box() const { return bn / cg; }
};

template <int BN, int EPI, int CG>
static int launch_gemm_inst(const CUtensorMap& a, const CUtensorMap& b, const GemmParams& p, cudaStream_t st) {
  if ((reinterpret_cast<uintptr_t>(p.out) & 15) || (p.ldo * (epi_out_f16(EPI) ? 2 : 4)) % 16)
    return fail("GEMM output must be 16-byte aligned with a 16-byte multiple row pitch");
  CUtensorMap c;
  // (split-operand mode: the rounding residuals are stored lo_off columns to the right of the N outputs)
  TRY(make_tmap_out(&c, p.out, p.M, p.lo_off ? p.lo_off + p.N : p.N, p.ldo, epi_out_f16(EPI)));
  if (p.lo_off && (p.N % 64 || p.lo_off < p.N)) return fail("split-operand output needs N %% 64 == 0 and lo_off >= N");
  CK(ensure_dynamic_smem(gemm_tcgen05_kernel<BN, EPI, CG>, gemm_smem_bytes(BN, CG)));
  const int m_tiles = (p.M + kBM * CG - 1) / (kBM * CG), n_tiles = (p.N + BN - 1) / BN;
  if (p.tile_count < 0 || p.tile_begin < 0 || p.tile_begin + p.tile_count > m_tiles * n_tiles)
    return fail("GEMM tile range [%d, +%d) outside the %d tiles", p.tile_begin, p.tile_count, m_tiles * n_tiles);
  const int groups = std::min(num_sms() / CG, p.tile_count > 0 ? p.tile_count : m_tiles * n_tiles);
  GemmParams q = p;
  q.split = 1;
  if (EPI == EPI_RESID_F32 && p.flags && g_gemm_split && p.tile_count == 0) {
    // partly-filled last wave: cut its tiles along K so that the idle groups take a share (gemm_work_unit)
    const int tiles = m_tiles * n_tiles, rem = tiles % groups;
    if (tiles > groups && rem) q.split = std::min({groups / rem, 4, p.K / kBK / 32, kSplitFlagInts / (rem * CG * kEpiWarps)});
    if (q.split < 2) q.split = 1;
  }
  {
    const int bit = epi_out_f16(EPI) ? 2 : 4;
    const size_t pitch = static_cast<size_t>(p.ldo) * (epi_out_f16(EPI) ? 2 : 4);
    q.direct = EPI != EPI_RESID_F32 && (g_epi_direct & bit) && !(reinterpret_cast<uintptr_t>(p.out) & 31) &&
               pitch % 32 == 0 && p.N % 16 == 0 && p.lo_off % 16 == 0;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(groups * CG);
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = gemm_smem_bytes(BN, CG);
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = g_pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  CK(cudaLaunchKernelEx(&cfg, gemm_tcgen05_kernel<BN, EPI, CG>, a, b, c, q));
  return 0;
}

template <int EPI>
static int launch_gemm_bn(GemmPlan g, const CUtensorMap& a, const CUtensorMap& b, const GemmParams& p,
                          cudaStream_t st) {
  if (g.cg == 2) {
    switch (g.bn) {
      case 128: return launch_gemm_inst<128, EPI, 2>(a, b, p, st);
      case 192: return launch_gemm_inst<192, EPI, 2>(a, b, p, st);
      case 256: return launch_gemm_inst<256, EPI, 2>(a, b, p, st);
    }
  } else if (g.cg == 1) {
    switch (g.bn) {
      case 64: return launch_gemm_inst<64, EPI, 1>(a, b, p, st);
      case 128: return launch_gemm_inst<128, EPI, 1>(a, b, p, st);
      case 192: return launch_gemm_inst<192, EPI, 1>(a, b, p, st);
      case 256: return launch_gemm_inst<256, EPI, 1>(a, b, p, st);
    }
  }
  return fail("unsupported GEMM tiling block_n=%d cta_group=%d", g.bn, g.cg);
}

static int launch_gemm(int epi, GemmPlan g, const CUtensorMap& a, const CUtensorMap& b, const GemmParams& p,
                       cudaStream_t st) {
  if (p.K % kBK) return fail("GEMM K=%d must be a multiple of %d", p.K, kBK);
  if (p.N % 8) return fail("GEMM N=%d must be a multiple of 8", p.N);
  switch (epi) {
    case EPI_BIAS_F16: return launch_gemm_bn<EPI_BIAS_F16>(g, a, b, p, st);
    case EPI_GELU_F16: return launch_gemm_bn<EPI_GELU_F16>(g, a, b, p, st);
    case EPI_RESID_F32: return launch_gemm_bn<EPI_RESID_F32>(g, a, b, p, st);
    case EPI_QKV_F16: return launch_gemm_bn<EPI_QKV_F16>(g, a, b, p, st);
    case EPI_GELU_F32: return launch_gemm_bn<EPI_GELU_F32>(g, a, b, p, st);
    case EPI_BIAS_F32: return launch_gemm_bn<EPI_BIAS_F32>(g, a, b, p, st);
  }
  return fail("unsupported GEMM epilogue %d", epi);
}

// Tiling choice: least estimated time on the persistent grid = waves x BN / efficiency.  Narrow tiles re-read the
// A operand from shared memory more often per FLOP, so they only win when they remove a whole wave.  CTA pairs
// (256-row tiles, half the operand bytes per SM) are used whenever the problem has more than one 128-row block;
// `g_force_cg` (PGIBBS_GEMM_CG=1|2) pins the choice for A/B measurements.
// (Tried and dropped: running the partly-filled last wave as a second launch with narrower tiles -- 0.6 % on FC2
// at M=16512, nothing on the out-projection once the extra launch is paid for.)
static int g_force_cg = 0;
static GemmPlan pick_gemm_plan(int M, int N, int multiple_of) {
  static const int cands[] = {256, 192, 128, 64};
  static const double eff1[] = {1.0, 0.95, 0.77, 0.45};   // measured, single CTA (B200, M=16512, K=1280)
  static const double eff2[] = {1.08, 1.03, 0.80, 0.0};   // CTA pair, relative to one CTA's 128x256 (same sweep)
  GemmPlan best;
  double best_cost = -1;
  for (int cg = 2; cg >= 1; --cg) {
    if (g_force_cg && cg != g_force_cg) continue;
    if (!g_force_cg && cg == 2 && M <= kBM) continue;
    const long m_tiles = (M + kBM * cg - 1) / (kBM * cg), slots = num_sms() / cg;
    for (int i = 0; i < 4; ++i) {
      const int bn = cands[i];
      if (bn % multiple_of) continue;
      const double eff = cg == 2 ? eff2[i] : eff1[i];
      if (eff <= 0) continue;
      const long tiles = m_tiles * ((N + bn - 1) / bn);
      // a wave of CTA pairs (256 x bn per pair) takes as long as a wave of single CTAs (128 x bn each)
      const double cost = static_cast<double>((tiles + slots - 1) / slots) * bn / eff;
      if (best_cost < 0 || cost < best_cost) { best_cost = cost; best.bn = bn; best.cg = cg; }
    }
  }
  return best;
}

// ------------------------------------------------------------------------------- small helper kernels
__global__ void f32_to_f16_kernel(const float* __restrict__ in, __half* __restrict__ out, long long n) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i < n) out[i] = __float2half_rn(in[i]);
}
// fp32 [rows, K] -> fp16 [rows, 2K] = [hi | lo]: hi = rn(w), lo = rn(w - hi)  (split-operand mode)
__global__ void f32_to_f16_split_kernel(const float* __restrict__ in, __half* __restrict__ out, long long rows, int K) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= rows * K) return;
  const long long r = i / K;
  const int c = static_cast<int>(i - r * K);
  const float w = in[i];
  const __half hi = __float2half_rn(w);
  out[r * 2 * K + c] = hi;
  out[r * 2 * K + K + c] = __float2half_rn(w - __half2float(hi));
}
__global__ void f16_to_f32_kernel(const __half* __restrict__ in, float* __restrict__ out, long long n) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i < n) out[i] = __half2float(in[i]);
}
// row t: [cos(t f_0) .. cos(t f_{half-1}) | sin(t f_0) .. sin(t f_{half-1})]  (the layout rope_chunk loads, gemm.cuh)
__global__ void rope_table_kernel(float* tab, int T, int half) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= T * half) return;
  const int t = i / half, j = i % half;
  // inv_freq = 1 / 10000^(2j/Dh), angle = t * inv_freq  (fair-esm rotary_embedding.py), in fp32 like torch
  const float inv_freq = 1.0f / powf(10000.0f, static_cast<float>(2 * j) / static_cast<float>(2 * half));
  const float ang = static_cast<float>(t) * inv_freq;
  tab[static_cast<size_t>(t) * 2 * half + j] = cosf(ang);
  tab[static_cast<size_t>(t) * 2 * half + half + j] = sinf(ang);
}

static int to_f16(const float* in, __half* out, long long n, cudaStream_t st) {
  f32_to_f16_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(in, out, n);
  CK(cudaGetLastError());
  return 0;
}
// weight [rows, K] -> fp16 [rows, K] (split = false) or [rows, 2K] = [hi | lo]
static int weight_to_f16(const float* in, __half* out, long long rows, int K, bool split, cudaStream_t st) {
  if (!split) return to_f16(in, out, rows * K, st);
  f32_to_f16_split_kernel<<<static_cast<unsigned>((rows * K + 255) / 256), 256, 0, st>>>(in, out, rows, K);
  CK(cudaGetLastError());
  return 0;
}

// --------------------------------------------------------------------------------------- the engine
struct LayerW {
  // single-sequence models: attn = self_attn.  MSA: row = row_self_attention, col = column_self_attention
  __half *wqkv = nullptr, *wo = nullptr, *w1 = nullptr, *w2 = nullptr;
  float *bqkv = nullptr, *bo = nullptr, *b1 = nullptr, *b2 = nullptr;
  const float *ln1w = nullptr, *ln1b = nullptr, *ln2w = nullptr, *ln2b = nullptr;
  const float *bias_k = nullptr, *bias_v = nullptr;  // ESM-1 add_bias_kv
  // MSA column attention
  __half *c_wqkv = nullptr, *c_wo = nullptr;
  float *c_bqkv = nullptr, *c_bo = nullptr;
  const float *lncw = nullptr, *lncb = nullptr;
  CUtensorMap m_wqkv, m_wo, m_w1, m_w2, m_cwqkv, m_cwo;
};

struct ProfEntry { double ms = 0; int launches = 0; };

}  // namespace pg

using namespace pg;

struct pgibbs_engine {
  pgibbs_model_config cfg;
  int device = 0;
  cudaStream_t stream = nullptr, own_stream = nullptr;
  cudaStream_t side_stream = nullptr;              // fork / join partner of `stream` (run_gemm_resid_ln)
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  std::map<std::string, std::pair<float*, int64_t>> raw;  // fp32 device copies by state-dict key
  std::vector<void*> owned;                               // packed weight allocations
  std::vector<LayerW> L;
  bool finalized = false;
  // 0 fast: one pass of fp16 operands.  1: weights carried as fp16 hi + lo (two passes).  2: activations too (three).
  int precision = 0;
  int wk() const { return precision >= 1 ? 2 : 1; }   // weight row = wk() x K halves
  int ak() const { return precision >= 2 ? 2 : 1; }   // GEMM-input activation row = ak() x K halves
  int k_segs() const { return precision + 1; }
  __half* w_dense = nullptr;
  float* b_dense = nullptr;
  CUtensorMap m_wdense;
  float* rope = nullptr;
  int rope_T = 0;
  // activations
  // T = token rows per sequence on the device = Tu (the caller's tokens per sequence) + 1 for ESM-1, whose extra
  // last row is the bias key/value slot of fair-esm's add_bias_kv (its q / residual are computed and ignored)
  int B = 0, R = 0, T = 0, Tu = 0, n_seq = 0, M = 0;
  float ln_eps = 1e-5f, embed_scale = 1.0f;
  int32_t* tokens = nullptr;
  float* x = nullptr;
  __half *h = nullptr, *qkv = nullptr, *ctx = nullptr, *ffn = nullptr, *hs = nullptr;
  float *g = nullptr, *logits = nullptr, *scores = nullptr;
  CUtensorMap m_h, m_ctx, m_ffn, m_hs, m_qkv3, m_ctx3, m_qkv3_keys;  // _keys: box of all (<= 256) key rows
  GemmPlan g_qkv, g_o, g_fc1, g_fc2, g_dense;
  // schedule / rng
  int32_t* positions = nullptr;
  int64_t positions_numel = 0, positions_used = 0;
  int n_iters = 0, P = 0, has_dup = 0;
  int64_t iter_stride = 0, chain_stride = 0;
  float* noise = nullptr;
  int64_t noise_numel = 0;
  int noise_stride = 0;
  uint64_t seed = 0x243F6A8885A308D3ull;
  int64_t rng_chain_offset = 0;  // global index of this engine's first chain (pgibbs_set_chain_offset)
  int32_t* valid_dev = nullptr;
  int32_t* identity_pos = nullptr;  // 0..T-1 (forward_logits)
  // last-wave K-split of the residual GEMMs (gemm.cuh: gemm_work_unit): ordering flags, launch counter
  int32_t* split_flags = nullptr;
  // Zigzag: consecutive kernels of the forward walk the token rows in opposite directions, so that each one starts
  // on the rows its producer wrote last -- the part of a 40-170 MB activation that is still in the 126 MB L2 --
  // instead of on the rows that have just been evicted (ascending order everywhere is the LRU worst case).
  int zigzag = 0;
  int next_dir() { if (!g_zigzag) return 0; zigzag ^= 1; return zigzag; }
  // CUDA-graph replay of the iteration loop: the iteration index lives on the device (Schedule::iter_dev)
  int32_t* iter_dev = nullptr;
  bool dev_iter = false;      // forward() is being issued in device-iteration mode
  int run_top_k = 0;          // the caller's top_k / burn-in for that mode
  int64_t run_burnin = 0;
  // scoring pass (pgibbs_score): per-slot target ids in, log-probabilities out; live only during that call
  int32_t* sc_targets = nullptr;
  float* sc_logp = nullptr;
  int64_t sc_cap = 0;
  bool sc_active = false;
  int layer_limit = -1;
  // profiling
  bool prof = false;
  std::map<std::string, ProfEntry> prof_acc;
  std::vector<std::tuple<std::string, cudaEvent_t, cudaEvent_t>> prof_pending;
  std::vector<cudaEvent_t> event_pool;
  int64_t launches = 0;
};

namespace pg {

struct ProfScope {
  pgibbs_engine* e;
  cudaEvent_t a = nullptr, b = nullptr;
  const char* name;
  ProfScope(pgibbs_engine* e_, const char* n) : e(e_), name(n) {
    e->launches++;
    if (e->prof) {
      a = take_event(e);
      b = take_event(e);
      cudaEventRecord(a, e->stream);
    }
  }
  static cudaEvent_t take_event(pgibbs_engine* e) {
    if (e->event_pool.empty()) {
      for (int i = 0; i < 256; ++i) {
        cudaEvent_t ev;
        cudaEventCreate(&ev);
        e->event_pool.push_back(ev);
      }
    }
    cudaEvent_t ev = e->event_pool.back();
    e->event_pool.pop_back();
    return ev;
  }
  ~ProfScope() {
    if (e->prof) {
      cudaEventRecord(b, e->stream);
      e->prof_pending.emplace_back(name, a, b);
    }
  }
};

static int prof_flush(pgibbs_engine* e) {
  if (e->prof_pending.empty()) return 0;
  CK(cudaStreamSynchronize(e->stream));
  for (auto& t : e->prof_pending) {
    float ms = 0;
    cudaEventElapsedTime(&ms, std::get<1>(t), std::get<2>(t));
    auto& acc = e->prof_acc[std::get<0>(t)];
    acc.ms += ms;
    acc.launches++;
    e->event_pool.push_back(std::get<1>(t));
    e->event_pool.push_back(std::get<2>(t));
  }
  e->prof_pending.clear();
  return 0;
}

template <typename T>
static int dev_alloc(T** p, size_t n) {
  *p = nullptr;
  CK(cudaMalloc(reinterpret_cast<void**>(p), std::max<size_t>(n, 1) * sizeof(T)));
  return 0;
}

static const float* raw_get(pgibbs_engine* e, const std::string& k, int64_t expect) {
  auto it = e->raw.find(k);
  if (it == e->raw.end()) { fail("missing weight tensor '%s'", k.c_str()); return nullptr; }
  if (expect >= 0 && it->second.second != expect) {
    fail("weight '%s' has %lld elements, expected %lld", k.c_str(), (long long)it->second.second, (long long)expect);
    return nullptr;
  }
  return it->second.first;
}

// concat [q;k;v] weights -> fp16 [3d, d], biases -> fp32 [3d]
static int pack_qkv(pgibbs_engine* e, const std::string& prefix, __half** w, float** b) {
  const int d = e->cfg.embed_dim;
  const size_t wrow = static_cast<size_t>(e->wk()) * d;   // halves per weight row ([hi | lo] in split mode)
  TRY(dev_alloc(w, static_cast<size_t>(3) * d * wrow));
  TRY(dev_alloc(b, static_cast<size_t>(3) * d));
  e->owned.push_back(*w);
  e->owned.push_back(*b);
  const char* names[3] = {"q_proj", "k_proj", "v_proj"};
  for (int i = 0; i < 3; ++i) {
    const float* ws = raw_get(e, prefix + names[i] + ".weight", static_cast<int64_t>(d) * d);
    const float* bs = raw_get(e, prefix + names[i] + ".bias", d);
    if (!ws || !bs) return 1;
    TRY(weight_to_f16(ws, *w + static_cast<size_t>(i) * d * wrow, d, d, e->precision >= 1, e->stream));
    CK(cudaMemcpyAsync(*b + static_cast<size_t>(i) * d, bs, d * sizeof(float), cudaMemcpyDeviceToDevice, e->stream));
  }
  return 0;
}

static int pack_linear(pgibbs_engine* e, const std::string& prefix, int n_out, int n_in, __half** w, float** b) {
  const float* ws = raw_get(e, prefix + ".weight", static_cast<int64_t>(n_out) * n_in);
  const float* bs = raw_get(e, prefix + ".bias", n_out);
  if (!ws || !bs) return 1;
  TRY(dev_alloc(w, static_cast<size_t>(n_out) * n_in * e->wk()));
  e->owned.push_back(*w);
  TRY(weight_to_f16(ws, *w, n_out, n_in, e->precision >= 1, e->stream));
  *b = const_cast<float*>(bs);
  return 0;
}

static void drop_raw(pgibbs_engine* e, const std::string& k) {
  auto it = e->raw.find(k);
  if (it != e->raw.end()) { cudaFree(it->second.first); e->raw.erase(it); }
}

static int free_activations(pgibbs_engine* e) {
  void* bufs[] = {e->tokens, e->x, e->h, e->qkv, e->ctx, e->ffn, e->hs, e->g, e->logits, e->scores, e->identity_pos};
  for (void* b : bufs) if (b) cudaFree(b);
  e->tokens = nullptr; e->x = nullptr; e->h = e->qkv = e->ctx = e->ffn = e->hs = nullptr;
  e->g = e->logits = e->scores = nullptr; e->identity_pos = nullptr;
  return 0;
}

static int build_weight_maps(pgibbs_engine* e) {
  const int d = e->cfg.embed_dim, F = e->cfg.ffn_dim;
  for (auto& l : e->L) {
    const int w = e->wk();   // weight rows are [hi | lo] in split-operand mode
    TRY(make_tmap_2d(&l.m_wqkv, l.wqkv, 3 * d, w * d, w * d, e->g_qkv.b_box()));
    TRY(make_tmap_2d(&l.m_wo, l.wo, d, w * d, w * d, e->g_o.b_box()));
    TRY(make_tmap_2d(&l.m_w1, l.w1, F, w * d, w * d, e->g_fc1.b_box()));
    TRY(make_tmap_2d(&l.m_w2, l.w2, d, w * F, w * F, e->g_fc2.b_box()));
    if (e->cfg.arch == PGIBBS_ARCH_MSA) {
      TRY(make_tmap_2d(&l.m_cwqkv, l.c_wqkv, 3 * d, w * d, w * d, e->g_qkv.b_box()));
      TRY(make_tmap_2d(&l.m_cwo, l.c_wo, d, w * d, w * d, e->g_o.b_box()));
    }
  }
  if (e->w_dense) TRY(make_tmap_2d(&e->m_wdense, e->w_dense, d, e->wk() * d, e->wk() * d, e->g_dense.b_box()));
  return 0;
}

static int allocate_shape(pgibbs_engine* e, int B, int R, int T);
static int ensure_shape(pgibbs_engine* e, int B, int R, int T) {
  if (B <= 0 || R <= 0 || T <= 0) return fail("invalid token shape (%d,%d,%d)", B, R, T);
  if (e->cfg.arch != PGIBBS_ARCH_MSA && R != 1) return fail("single-sequence model needs R == 1, got %d", R);
  if (e->cfg.arch != PGIBBS_ARCH_ESM2 && T > e->cfg.max_positions)
    return fail("sequence of %d tokens exceeds the learned position table (%d)", T, e->cfg.max_positions);
  if (e->cfg.arch == PGIBBS_ARCH_MSA && R > 1024) return fail("MSA depth %d exceeds 1024", R);
  if (B == e->B && R == e->R && T == e->Tu) return 0;
  CK(cudaStreamSynchronize(e->stream));
  free_activations(e);
  // A failed allocation (e.g. out of memory on a large batch) must not leave the new shape recorded over null or
  // partial buffers: the next call with the same shape would take the early return above and launch on them.
  const int rc = allocate_shape(e, B, R, T);
  if (rc) {
    free_activations(e);
    e->B = e->R = e->T = e->Tu = e->n_seq = e->M = 0;
  }
  return rc;
}
static int allocate_shape(pgibbs_engine* e, int B, int R, int T) {
  const int d = e->cfg.embed_dim, F = e->cfg.ffn_dim, V = e->cfg.vocab;
  e->Tu = T;
  if (e->cfg.arch == PGIBBS_ARCH_ESM1) T += 1;  // the bias key/value slot
  e->B = B; e->R = R; e->T = T; e->n_seq = B * R;
  const size_t M = static_cast<size_t>(B) * R * T;
  e->M = static_cast<int>(M);
  TRY(dev_alloc(&e->tokens, M));
  TRY(dev_alloc(&e->x, M * d));
  const size_t a = e->ak();   // GEMM-input activations are [hi | lo] rows in split-operand mode
  TRY(dev_alloc(&e->h, M * d * a));
  TRY(dev_alloc(&e->qkv, M * 3 * d));
  TRY(dev_alloc(&e->ctx, M * d * a));
  TRY(dev_alloc(&e->ffn, M * F * a));
  TRY(dev_alloc(&e->hs, M * d * a));
  // attention kernels without a residual output (mma.sync paths, MSA kernels) leave the lo half of ctx at zero
  if (a > 1) CK(cudaMemsetAsync(e->ctx, 0, M * d * a * sizeof(__half), e->stream));
  TRY(dev_alloc(&e->g, M * d));
  TRY(dev_alloc(&e->logits, M * V));
  if (e->cfg.arch == PGIBBS_ARCH_MSA)
    TRY(dev_alloc(&e->scores, static_cast<size_t>(B) * e->cfg.heads * T * T));
  TRY(dev_alloc(&e->identity_pos, static_cast<size_t>(T)));
  std::vector<int32_t> idp(T);
  for (int i = 0; i < T; ++i) idp[i] = i;
  CK(cudaMemcpy(e->identity_pos, idp.data(), T * sizeof(int32_t), cudaMemcpyHostToDevice));
  TRY(make_tmap_2d(&e->m_h, e->h, M, a * d, a * d, kBM));
  TRY(make_tmap_2d(&e->m_ctx, e->ctx, M, a * d, a * d, kBM));
  TRY(make_tmap_2d(&e->m_ffn, e->ffn, M, a * F, a * F, kBM));
  TRY(make_tmap_2d(&e->m_hs, e->hs, M, a * d, a * d, kBM));
  TRY(make_tmap_qkv3(&e->m_qkv3, e->qkv, static_cast<uint64_t>(B) * R, T, 3 * d));
  TRY(make_tmap_qkv3(&e->m_ctx3, e->ctx, static_cast<uint64_t>(B) * R, T, a * d, 128, d));
  if (T <= 256) TRY(make_tmap_qkv3(&e->m_qkv3_keys, e->qkv, static_cast<uint64_t>(B) * R, T, 3 * d, (T + 15) & ~15));
  const int hd = d / e->cfg.heads;
  e->g_qkv = pick_gemm_plan(e->M, 3 * d, hd >= 64 ? 64 : 32);
  e->g_o = pick_gemm_plan(e->M, d, 16);
  e->g_fc1 = pick_gemm_plan(e->M, F, 16);
  e->g_fc2 = pick_gemm_plan(e->M, d, 16);
  e->g_dense = pick_gemm_plan(e->M, d, 16);
  TRY(build_weight_maps(e));
  if (e->cfg.arch == PGIBBS_ARCH_ESM2 && e->rope_T < T) {
    if (e->rope) cudaFree(e->rope);
    TRY(dev_alloc(&e->rope, static_cast<size_t>(T) * hd));
    rope_table_kernel<<<(T * (hd / 2) + 255) / 256, 256, 0, e->stream>>>(e->rope, T, hd / 2);
    CK(cudaGetLastError());
    e->rope_T = T;
  }
  return 0;
}

// ----------------------------------------------------------------------------- the forward pass
// LayerNorm of rows [row0, row0 + rows) of x into the same rows of `out` (ungathered), or of the scheduled rows
// (`gather`).  `dir` < 0: take the next zigzag direction; `st`: stream (default the engine's), `pdl`: programmatic
// dependent launch allowed (not for a kernel whose predecessor is an event of another stream).
static int run_ln(pgibbs_engine* e, const float* x, const float* w, const float* b, __half* out, int rows,
                  const Schedule* gather, int iter, int row0 = 0, int dir = -1, cudaStream_t st = nullptr,
                  bool pdl = true) {
  LnParams p{};
  p.d = e->cfg.embed_dim;
  if (e->precision >= 2) { p.ld_out = 2 * p.d; p.lo_off = p.d; }   // [hi | lo] rows (split-operand mode)
  p.x = x + static_cast<size_t>(row0) * p.d; p.w = w; p.b = b;
  p.out = out + static_cast<size_t>(row0) * (p.ld_out ? p.ld_out : p.d);
  p.rows_out = rows; p.eps = e->ln_eps;
  if (gather) p.sched = *gather; else p.sched.positions = nullptr;
  p.iter = iter; p.T = e->T;
  p.reverse = gather ? 0 : dir >= 0 ? dir : e->next_dir();
  if (rows <= 0) return 0;
  ProfScope ps(e, "layernorm");
  const dim3 grid((rows + 7) / 8);
  const int vpl = (p.d / 4 + 31) / 32;  // float4 vectors per lane
  if (!st) st = e->stream;
  t_no_pdl = !pdl;
  cudaError_t ce;
  if (vpl <= 3) ce = launch_pdl(layernorm_kernel<true, 3>, grid, dim3(256), 0, st, p);
  else if (vpl <= 6) ce = launch_pdl(layernorm_kernel<true, 6>, grid, dim3(256), 0, st, p);
  else if (vpl <= 10) ce = launch_pdl(layernorm_kernel<true, 10>, grid, dim3(256), 0, st, p);
  else ce = launch_pdl(layernorm_kernel<true, kMaxVecPerLane>, grid, dim3(256), 0, st, p);
  t_no_pdl = false;
  CK(ce);
  return 0;
}

// g[i] = x[row of scheduled position i] (fp32 copy, no LayerNorm): the LM-head input of ESM-1.
static int run_gather_f32(pgibbs_engine* e, const float* x, float* out, int rows, const Schedule& gather, int iter) {
  LnParams p{};
  p.x = x; p.w = nullptr; p.b = nullptr; p.out = out; p.rows_out = rows; p.d = e->cfg.embed_dim; p.eps = e->ln_eps;
  p.sched = gather; p.iter = iter; p.T = e->T; p.identity = 1;
  ProfScope ps(e, "layernorm");
  CK(launch_pdl(layernorm_kernel<false, kMaxVecPerLane>, dim3((rows + 7) / 8), dim3(256), 0, e->stream, p));
  return 0;
}

static int run_gemm(pgibbs_engine* e, const char* name, int epi, GemmPlan g, const CUtensorMap& a,
                    const CUtensorMap& b, GemmParams p) {
  if (epi == EPI_RESID_F32) p.flags = e->split_flags;
  p.reverse = e->next_dir();
  p.k_segs = e->k_segs();
  ProfScope ps(e, name);
  return launch_gemm(epi, g, a, b, p, e->stream);
}

// Where a GEMM of M x N outputs is cut: `full` tiles (whole waves of `groups` CTA groups) + `rem` tiles, and the rows
// [lo, hi) of the row blocks whose EVERY column tile lies in the full waves (tile sequence: n fastest, walked from the
// last tile down when `reverse`).  Pure arithmetic (pgibbs_debug_tail_plan exposes it to the CPU tests).
struct TailPlan { int full, rem, lo, hi; };
static TailPlan tail_plan(int M, int N, int bn, int cg, int sms, int reverse) {
  const int rows_per_blk = kBM * cg;
  const int m_tiles = (M + rows_per_blk - 1) / rows_per_blk, n_tiles = (N + bn - 1) / bn;
  const int tiles = m_tiles * n_tiles, groups = std::max(1, std::min(sms / cg, tiles));
  TailPlan t{};
  t.full = tiles / groups * groups;
  t.rem = tiles - t.full;
  const int done_blks = t.full / n_tiles;     // row blocks whose every column tile is in the full waves
  // ascending walk: blocks 0 .. done_blks-1 = rows [0, done_blks * rows_per_blk); descending walk: the LAST done_blks blocks
  t.lo = reverse ? std::min(M, (m_tiles - done_blks) * rows_per_blk) : 0;
  t.hi = reverse ? M : std::min(M, done_blks * rows_per_blk);
  if (done_blks == 0) t.lo = t.hi = 0;
  return t;
}

// A residual GEMM (x += A W^T + b) followed by the LayerNorm of x into h.  Plain form: the two launches one after the
// other (PGIBBS_TAIL_OVERLAP=0, profiling, split precision).  Otherwise, with a partly-filled last wave (config 2: 325 tiles on 74 CTA pairs = 4 full waves +
// 29 tiles, i.e. 60 % of the SMs idle for a fifth of the GEMM) the GEMM is cut at the wave boundary: after the full waves
// a fork hands the row blocks that are COMPLETE (all their column tiles lie in the full waves) to a LayerNorm on the side
// stream, which runs on the SMs the last wave leaves idle; the rest of the rows are normalised after the join.  Same
// kernels, same per-element arithmetic, rows are independent: results are bit-identical to the plain form.  Inside a
// stream capture the fork / join become parallel branches of the graph.
static int run_gemm_resid_ln(pgibbs_engine* e, const char* name, GemmPlan g, const CUtensorMap& a, const CUtensorMap& b,
                             GemmParams p, const float* ln_w, const float* ln_b) {
  const int M = p.M;
  const TailPlan probe = tail_plan(M, p.N, g.bn, g.cg, num_sms(), 0);
  const bool overlap = g_tail_overlap && ln_w && !e->prof && !g_gemm_split && probe.rem > 0 && probe.hi > probe.lo &&
                       e->precision == 0;
  if (!overlap) {
    TRY(run_gemm(e, name, EPI_RESID_F32, g, a, b, p));
    if (ln_w) TRY(run_ln(e, e->x, ln_w, ln_b, e->h, M, nullptr, 0));
    return 0;
  }
  if (!e->side_stream) {
    CK(cudaStreamCreateWithFlags(&e->side_stream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&e->ev_join, cudaEventDisableTiming));
  }
  const int gemm_dir = e->next_dir(), ln_dir = e->next_dir();
  p.flags = e->split_flags; p.reverse = gemm_dir; p.k_segs = e->k_segs();
  const TailPlan t = tail_plan(M, p.N, g.bn, g.cg, num_sms(), gemm_dir);
  GemmParams pa = p, pb = p;
  pa.tile_begin = 0; pa.tile_count = t.full;
  pb.tile_begin = t.full; pb.tile_count = t.rem;
  const int lo = t.lo, hi = t.hi;   // rows of the complete blocks
  { ProfScope ps(e, name); TRY(launch_gemm(EPI_RESID_F32, g, a, b, pa, e->stream)); }
  CK(cudaEventRecord(e->ev_fork, e->stream));
  CK(cudaStreamWaitEvent(e->side_stream, e->ev_fork, 0));
  TRY(run_ln(e, e->x, ln_w, ln_b, e->h, hi - lo, nullptr, 0, lo, ln_dir, e->side_stream, false));
  CK(cudaEventRecord(e->ev_join, e->side_stream));
  { ProfScope ps(e, name); TRY(launch_gemm(EPI_RESID_F32, g, a, b, pb, e->stream)); }
  CK(cudaStreamWaitEvent(e->stream, e->ev_join, 0));
  // the rows the last wave was still working on
  if (gemm_dir) TRY(run_ln(e, e->x, ln_w, ln_b, e->h, lo, nullptr, 0, 0, ln_dir, e->stream, false));
  else TRY(run_ln(e, e->x, ln_w, ln_b, e->h, M - hi, nullptr, 0, hi, ln_dir, e->stream, false));
  return 0;
}

static int launch_attention(const AttnParams& p, int groups, int H, int hd, cudaStream_t st) {
  const int nq = p.T - p.q_begin;  // query rows handled by this launch
  if (nq <= 16 && hd == 64) {      // a few trailing rows (tail of the tcgen05 kernel): one warp per (group, head)
    dim3 grid(1, H, groups);
    CK(launch_pdl(attention_kernel<64, 1>, grid, dim3(32), 0, st, p));
  } else if (nq <= 32) {  // short groups (MSA column attention over R <= 32 rows): 2 warps = 32 queries per CTA
    dim3 grid((nq + 31) / 32, H, groups);
    switch (hd) {
      case 16: CK(launch_pdl(attention_kernel<16, 2>, grid, dim3(64), 0, st, p)); break;
      case 32: CK(launch_pdl(attention_kernel<32, 2>, grid, dim3(64), 0, st, p)); break;
      case 64: CK(launch_pdl(attention_kernel<64, 2>, grid, dim3(64), 0, st, p)); break;
      default: return fail("unsupported head_dim %d (16, 32, 64)", hd);
    }
  } else {
    dim3 grid((nq + 63) / 64, H, groups);
    switch (hd) {
      case 16: CK(launch_pdl(attention_kernel<16, 4>, grid, dim3(128), 0, st, p)); break;
      case 32: CK(launch_pdl(attention_kernel<32, 4>, grid, dim3(128), 0, st, p)); break;
      case 64: CK(launch_pdl(attention_kernel<64, 4>, grid, dim3(128), 0, st, p)); break;
      default: return fail("unsupported head_dim %d (16, 32, 64)", hd);
    }
  }
  CK(cudaGetLastError());
  return 0;
}

// head_dim 64 sequence attention runs on tcgen05 (attention_fa.cuh); everything else keeps the mma.sync kernel.
// PGIBBS_ATTN=legacy selects the mma.sync kernel for A/B measurements.  PGIBBS_ATTN_TAIL: 1 (default) trailing rows
// (T % 128 <= 8) on the kernel's tail warp, 2 on a second mma.sync launch, 0 as a partly filled tile.
static unsigned long long* g_fa_trace = nullptr;  // device buffer for the attention timeline (debug)
static int g_attn_mode = -1;  // 0 legacy, 2 fa
static int g_attn_tail = 1;
static int g_attn_stagger = 600;  // PGIBBS_ATTN_STAGGER (cycles)
static int attn_mode(int hd) {
  if (g_attn_mode < 0) {
    const char* v = getenv("PGIBBS_ATTN");
    g_attn_mode = (v && !strcmp(v, "legacy")) ? 0 : 2;
    if (const char* t = getenv("PGIBBS_ATTN_TAIL")) g_attn_tail = atoi(t);
    if (const char* t = getenv("PGIBBS_ATTN_STAGGER")) g_attn_stagger = atoi(t);
  }
  return hd == 64 ? g_attn_mode : 0;
}
// qkv: fused activation [n_seq*T, 3*H*64]; ctx: [n_seq*T, H*64].
static int launch_attention_fa(const CUtensorMap& qkv3, const CUtensorMap& ctx3, const __half* qkv, __half* ctx,
                               int n_seq, int T, int H, cudaStream_t st, int reverse = 0, int ctx_split = 0) {
  CK(ensure_dynamic_smem(attention_fa_kernel, kFaSmemBytes));
  // Trailing rows (T = 128 k + r): r <= 8 -> the kernel's tail warp; r <= 16 -> the mma.sync kernel (second launch);
  // otherwise a normal partly filled tile.  PGIBBS_ATTN_TAIL=0 forces the partly filled tile.
  const int tail = T % 128;
  const bool any_tail = g_attn_tail && T > 128 && tail > 0 && tail <= 16;
  const bool in_kernel = any_tail && tail <= 8 && g_attn_tail != 2;
  AttnFaParams p{T, H, n_seq, any_tail ? T / 128 : (T + 127) / 128, g_fa_trace, g_attn_stagger,
                 in_kernel ? tail : 0, qkv, ctx, reverse, (ctx_split ? 2 : 1) * H * 64, ctx_split ? H * 64 : 0};
  const int n_items = n_seq * H * ((p.n_tiles + 1) / 2);
  CK(launch_pdl(attention_fa_kernel, dim3(std::min(num_sms(), n_items)), dim3(kFaThreads), kFaSmemBytes, st, qkv3, ctx3, p));
  if (any_tail && !in_kernel) {
    const int d = H * 64;
    AttnParams tp{qkv, ctx, T, 3 * d, (ctx_split ? 2 : 1) * d, d, 2 * d, 1, 0, 1, T, T - tail};
    TRY(launch_attention(tp, n_seq, H, 64, st));
  }
  return 0;
}

// Tied row attention of the MSA Transformer: tcgen05 kernel for head_dim 64 and alignments of up to 256 columns
// (msa_row_tc.cuh), the mma.sync kernels otherwise (PGIBBS_MSA_ROW=legacy forces them).
static int run_msa_row_attention(pgibbs_engine* e) {
  const auto& c = e->cfg;
  const int hd = c.embed_dim / c.heads;
  static int legacy = -1;
  if (legacy < 0) { const char* v = getenv("PGIBBS_MSA_ROW"); legacy = (v && !strcmp(v, "legacy")) ? 1 : 0; }
  ProfScope ps(e, "msa_row_attention");
  if (hd == 64 && e->T <= 256 && !legacy) {
    MsaRowParams p{e->R, e->T, c.heads, (e->T + 15) & ~15, 0, e->next_dir()};
    p.stages = std::min(8, (227 * 1024 - 2 * kMrQBytes - 2048) / mr_stage_bytes(p.NK));
    const int smem = mr_smem_bytes(p.NK, p.stages);
    CK(ensure_dynamic_smem(msa_row_attention_tc_kernel, smem));
    dim3 grid((e->T + 127) / 128, c.heads, e->B);
    CK(launch_pdl(msa_row_attention_tc_kernel, grid, dim3(kMrThreads), smem, e->stream, e->m_qkv3, e->m_qkv3_keys, e->m_ctx3, p));
    return 0;
  }
  if (const char* m = launch_msa_row_attention(e->qkv, e->ctx, e->scores, e->B, e->R, e->T, c.heads, hd, e->stream,
                                               e->ak() * c.embed_dim))
    return fail("%s", m);
  return 0;
}

static int run_attention(pgibbs_engine* e) {
  const int d = e->cfg.embed_dim, H = e->cfg.heads, hd = d / H;
  ProfScope ps(e, "attention");
  const int mode = attn_mode(hd);
  if (mode == 2)
    return launch_attention_fa(e->m_qkv3, e->m_ctx3, e->qkv, e->ctx, e->n_seq, e->T, H, e->stream, e->next_dir(),
                               e->precision >= 2);
  AttnParams p{e->qkv, e->ctx, e->T, 3 * d, e->ak() * d, d, 2 * d, 1, 0, 1, e->T};
  return launch_attention(p, e->n_seq, H, hd, e->stream);
}

static GemmParams gp(int M, int N, int K, const float* bias, void* out, int ldo) {
  GemmParams p{};
  p.M = M; p.N = N; p.K = K; p.bias = bias; p.out = out; p.ldo = ldo;
  return p;
}

// One transformer forward over the resident tokens.  `sched`: rows the LM head is evaluated on (chain-major);
// sampling writes tokens when `sample` is set, logits rows are stored when `logits_out` is non-null.
static int forward(pgibbs_engine* e, const Schedule& sched_in, int n_chains, int iter, bool sample, int k_eff,
                   float temperature, int n_valid, float* logits_out) {
  Schedule sched = sched_in;
  if (e->dev_iter) sched.iter_dev = e->iter_dev;   // `iter` / `k_eff` are ignored by the kernels in this mode
  const auto& c = e->cfg;
  const int d = c.embed_dim, F = c.ffn_dim, M = e->M, hd = d / c.heads;
  cudaStream_t st = e->stream;
  {
    EmbedParams p{};
    p.tokens = e->tokens;
    p.tok_emb = raw_get(e, "embed_tokens.weight", -1);
    p.pos_emb = c.arch == PGIBBS_ARCH_ESM2 ? nullptr : raw_get(e, "embed_positions.weight", -1);
    p.row_emb = c.arch == PGIBBS_ARCH_MSA ? raw_get(e, "msa_position_embedding", -1) : nullptr;
    p.scale = e->embed_scale;
    if (c.arch != PGIBBS_ARCH_ESM2 && c.arch != PGIBBS_ARCH_ESM1) {
      p.ln_w = raw_get(e, "emb_layer_norm_before.weight", -1);
      p.ln_b = raw_get(e, "emb_layer_norm_before.bias", -1);
    }
    p.x = e->x; p.n_seq = e->n_seq; p.T = e->T; p.d = d; p.rows_per_msa = e->R;
    p.mask_idx = c.mask_idx; p.token_dropout = c.token_dropout; p.eps = e->ln_eps;
    ProfScope ps(e, "embed");
    const dim3 grid((M + 7) / 8);
    const int vpl = (d / 4 + 31) / 32;  // float4 vectors per lane
    if (vpl <= 3) CK(launch_pdl(embed_kernel<3>, grid, dim3(256), 0, st, p));
    else if (vpl <= 6) CK(launch_pdl(embed_kernel<6>, grid, dim3(256), 0, st, p));
    else if (vpl <= 10) CK(launch_pdl(embed_kernel<10>, grid, dim3(256), 0, st, p));
    else CK(launch_pdl(embed_kernel<kMaxVecPerLane>, grid, dim3(256), 0, st, p));
  }
  const int n_layers = e->layer_limit >= 0 ? std::min(e->layer_limit, c.layers) : c.layers;
  for (int li = 0; li < n_layers; ++li) {
    LayerW& l = e->L[li];
    if (c.arch == PGIBBS_ARCH_MSA) {
      const float row_scale = (1.0f / sqrtf(static_cast<float>(hd))) / sqrtf(static_cast<float>(e->R));
      // tied row attention (the first LayerNorm of layers > 0 was issued behind the previous layer's FC2)
      if (li == 0) TRY(run_ln(e, e->x, l.ln1w, l.ln1b, e->h, M, nullptr, 0));
      GemmParams q = gp(M, 3 * d, d, l.bqkv, e->qkv, 3 * d);
      q.q_cols = d; q.q_scale = row_scale; q.rope_cols = 0; q.head_dim = hd; q.seq_len = e->T;
      TRY(run_gemm(e, "gemm_qkv", EPI_QKV_F16, e->g_qkv, e->m_h, l.m_wqkv, q));
      TRY(run_msa_row_attention(e));
      // out-projection, then the column attention's LayerNorm
      TRY(run_gemm_resid_ln(e, "gemm_out", e->g_o, e->m_ctx, l.m_wo, gp(M, d, d, l.bo, e->x, d), l.lncw, l.lncb));
      // (R == 1: one key per column, softmax = 1, ctx = v -- fair-esm short-cuts this case to Wo(Wv x), same result)
      GemmParams qc = gp(M, 3 * d, d, l.c_bqkv, e->qkv, 3 * d);
      qc.q_cols = d; qc.q_scale = 1.0f / sqrtf(static_cast<float>(hd)); qc.rope_cols = 0; qc.head_dim = hd;
      qc.seq_len = e->T;
      TRY(run_gemm(e, "gemm_qkv", EPI_QKV_F16, e->g_qkv, e->m_h, l.m_cwqkv, qc));
      {
        ProfScope ps(e, "msa_col_attention");
        AttnParams ap{e->qkv, e->ctx, e->R, 3 * d, e->ak() * d, d, 2 * d, e->T, 1, e->T, static_cast<long long>(e->R) * e->T};
        ap.reverse = e->next_dir();   // (only the dedicated column kernel honours it)
        const char* m = hd == 64 ? launch_msa_col_attention(ap, e->B * e->T, c.heads, st) : "";
        if (m && *m) return fail("%s", m);
        if (m) TRY(launch_attention(ap, e->B * e->T, c.heads, hd, st));  // shape not covered: generic kernel
      }
      TRY(run_gemm_resid_ln(e, "gemm_out", e->g_o, e->m_ctx, l.m_cwo, gp(M, d, d, l.c_bo, e->x, d), l.ln2w, l.ln2b));
    } else {
      if (li == 0) TRY(run_ln(e, e->x, l.ln1w, l.ln1b, e->h, M, nullptr, 0));
      GemmParams q = gp(M, 3 * d, d, l.bqkv, e->qkv, 3 * d);
      q.q_cols = d; q.q_scale = 1.0f / sqrtf(static_cast<float>(hd));
      q.rope_cols = c.arch == PGIBBS_ARCH_ESM2 ? 2 * d : 0;
      q.head_dim = hd; q.seq_len = e->T; q.rope = e->rope;
      TRY(run_gemm(e, "gemm_qkv", EPI_QKV_F16, e->g_qkv, e->m_h, l.m_wqkv, q));
      if (c.arch == PGIBBS_ARCH_ESM1) {  // the extra row of every sequence becomes the learned bias key / value
        ProfScope ps(e, "bias_kv");
        CK(launch_pdl(bias_kv_kernel, dim3((e->n_seq * d + 255) / 256), dim3(256), 0, st, e->qkv, l.bias_k, l.bias_v,
                      e->n_seq, e->T, d));
      }
      TRY(run_attention(e));
      TRY(run_gemm_resid_ln(e, "gemm_out", e->g_o, e->m_ctx, l.m_wo, gp(M, d, d, l.bo, e->x, d), l.ln2w, l.ln2b));
    }
    {
      GemmParams f1 = gp(M, F, d, l.b1, e->ffn, e->ak() * F);
      if (e->precision >= 2) f1.lo_off = F;   // ffn rows are [hi | lo]
      TRY(run_gemm(e, "gemm_fc1", EPI_GELU_F16, e->g_fc1, e->m_h, l.m_w1, f1));
    }
    // FC2, then the next layer's first LayerNorm (none behind the last layer: the LM head normalises its own rows)
    const bool more = li + 1 < n_layers;
    TRY(run_gemm_resid_ln(e, "gemm_fc2", e->g_fc2, e->m_ffn, l.m_w2, gp(M, d, F, l.b2, e->x, d),
                          more ? e->L[li + 1].ln1w : nullptr, more ? e->L[li + 1].ln1b : nullptr));
  }
  // LM head on the scheduled rows only
  const int rows = n_chains * sched.P;
  const bool esm1 = c.arch == PGIBBS_ARCH_ESM1;
  if (esm1) {  // logits = embed_out . x + embed_out_bias: no final LayerNorm, no dense / GELU / LayerNorm head
    TRY(run_gather_f32(e, e->x, e->g, rows, sched, iter));
  } else {
    TRY(run_ln(e, e->x, raw_get(e, "emb_layer_norm_after.weight", -1), raw_get(e, "emb_layer_norm_after.bias", -1),
               e->hs, rows, &sched, iter));
    TRY(run_gemm(e, "gemm_head", EPI_GELU_F32, e->g_dense, e->m_hs, e->m_wdense, gp(rows, d, d, e->b_dense, e->g, d)));
  }
  {
    HeadParams p{};
    p.g = e->g;
    p.no_ln = esm1;
    p.ln_w = esm1 ? nullptr : raw_get(e, "lm_head.layer_norm.weight", -1);
    p.ln_b = esm1 ? nullptr : raw_get(e, "lm_head.layer_norm.bias", -1);
    p.emb = raw_get(e, esm1 ? "embed_out" : "embed_tokens.weight", -1);
    p.out_bias = raw_get(e, esm1 ? "embed_out_bias" : "lm_head.bias", -1);
    p.logits_out = logits_out;
    p.tokens = sample ? e->tokens : nullptr;
    p.rows = rows; p.d = d; p.V = c.vocab; p.T = e->T; p.eps = e->ln_eps;
    p.sched = sched; p.iter = iter;
    p.valid_ids = e->valid_dev; p.n_valid = n_valid; p.top_k = k_eff; p.temperature = temperature;
    p.noise = (sample && e->noise) ? e->noise + (e->dev_iter ? 0 : static_cast<int64_t>(iter) * rows * e->noise_stride)
                                   : nullptr;
    p.top_k_raw = e->run_top_k; p.burnin = e->run_burnin;
    p.noise_stride = e->noise_stride;
    p.seed = e->seed;
    p.rng_row_offset = e->rng_chain_offset * sched.P;
    p.skip_dup_writes = e->has_dup;
    if (e->sc_active) { p.targets = e->sc_targets; p.logp_out = e->sc_logp; }
    const size_t emb_bytes = static_cast<size_t>(c.vocab) * d * sizeof(float);
    p.emb_in_smem = emb_bytes <= 200 * 1024;
    // Rows per warp: as many (4, 2, 1) as still give every warp of the grid work -- the projection table is streamed
    // from shared memory once per row group; registers per row sized to the model as for the LayerNorm kernel.
    const int vpl = (d / 4 + 31) / 32;
    const size_t smem = p.emb_in_smem ? emb_bytes : 0;
    ProfScope ps(e, "head_sample");
    auto go = [&](auto kernel, int R, int threads) -> int {
      const int wpb = threads / 32, units = (rows + R - 1) / R;
      CK(ensure_dynamic_smem(kernel, 200 * 1024));
      CK(launch_pdl(kernel, dim3(std::min(num_sms(), (units + wpb - 1) / wpb)), dim3(threads), smem, st, p));
      return 0;
    };
    const char* cap = getenv("PGIBBS_HEAD_ROWS");   // caps the rows per warp (A/B, and the R-invariance test)
    const int max_r = cap ? atoi(cap) : 4;
    // measured (profiles/r02zq_head_sample_rows_per_warp.txt): 4 rows pay off for d <= 768 once every warp gets several
    // groups; at d = 1280 four rows need 255 registers (256 threads) and are no faster than two
    const auto enough = [&](int R, int waves) { return R <= max_r && rows >= R * 12 * num_sms() * waves; };
    if (vpl <= 6 && enough(4, 2)) TRY(go(head_sample_kernel<4, 6>, 4, 384));
    else if (vpl <= 10 && enough(2, 1)) TRY(go(head_sample_kernel<2, 10>, 2, 384));
    else if (vpl <= 10) TRY(go(head_sample_kernel<1, 10>, 1, 384));
    else TRY(go(head_sample_kernel<1, kMaxVecPerLane>, 1, 384));
  }
  return 0;
}

static int check_ready(pgibbs_engine* e) {
  if (!e) return fail("null engine");
  CK(cudaSetDevice(e->device));
  if (!e->finalized) return fail("weights not finalized (call pgibbs_finalize_weights)");
  return 0;
}

static int upload_valid(pgibbs_engine* e, const int32_t* valid_ids, int n_valid) {
  if (n_valid <= 0 || n_valid > 64) return fail("n_valid=%d out of range (1..64)", n_valid);
  std::vector<int32_t> v(n_valid);
  CK(cudaMemcpy(v.data(), valid_ids, n_valid * sizeof(int32_t), cudaMemcpyDefault));
  for (int i = 0; i < n_valid; ++i)
    if (v[i] < 0 || v[i] >= e->cfg.vocab) return fail("valid id %d outside vocabulary", v[i]);
  if (!e->valid_dev) TRY(dev_alloc(&e->valid_dev, 64));
  CK(cudaMemcpyAsync(e->valid_dev, v.data(), n_valid * sizeof(int32_t), cudaMemcpyHostToDevice, e->stream));
  CK(cudaStreamSynchronize(e->stream));  // v is a stack-backed vector
  return 0;
}

static int run_iters(pgibbs_engine* e, int first_iter, int num_iters, int64_t burnin, int top_k, float temperature,
                     int mask_flag, int mask_row, int target_row, bool single, const int32_t* valid_ids,
                     int n_valid) {
  TRY(check_ready(e));
  if (!e->tokens) return fail("no tokens resident (call pgibbs_set_tokens)");
  if (!e->positions) return fail("no schedule set (call pgibbs_set_schedule)");
  if (first_iter < 0 || first_iter + num_iters > e->n_iters)
    return fail("iterations [%d,%d) outside the schedule (%d)", first_iter, first_iter + num_iters, e->n_iters);
  TRY(upload_valid(e, valid_ids, n_valid));
  const int n_chains = single ? e->B : e->n_seq;
  const int64_t rows = static_cast<int64_t>(n_chains) * e->P;
  if (static_cast<int64_t>(e->n_iters - 1) * e->iter_stride + static_cast<int64_t>(n_chains - 1) * e->chain_stride +
          e->P > e->positions_used)
    return fail("schedule buffer too small for %d chains", n_chains);
  if (e->noise && e->noise_numel < static_cast<int64_t>(first_iter + num_iters) * rows * e->noise_stride)
    return fail("replay noise too short for the requested iterations");
  Schedule s{e->positions, e->iter_stride, e->chain_stride, e->P, single ? e->R : 1, single ? target_row : 0};
  auto one_iteration = [&](int it) -> int {
    if (mask_flag) {
      Schedule ms = s;
      if (single) ms.seq_offset = mask_row;
      if (e->dev_iter) ms.iter_dev = e->iter_dev;
      ProfScope ps(e, "mask_scatter");
      CK(launch_pdl(mask_scatter_kernel, dim3(static_cast<unsigned>((rows + 255) / 256)), dim3(256), 0, e->stream,
                    e->tokens, n_chains, e->T, ms, it, e->cfg.mask_idx));
    }
    const int k_eff = (it < burnin || top_k <= 0 || top_k > n_valid) ? n_valid : top_k;
    TRY(forward(e, s, n_chains, it, true, k_eff, temperature, n_valid, nullptr));
    if (e->dev_iter) CK(launch_pdl(advance_iter_kernel, dim3(1), dim3(1), 0, e->stream, e->iter_dev));
    return 0;
  };
  // Long runs: the first iteration is launched kernel by kernel with the iteration index on the device, the same
  // launches are then captured once into a CUDA graph, and the graph is replayed for the remaining iterations (one
  // driver call per iteration instead of ~270; the kernels read the iteration from Schedule::iter_dev).
  if (g_graph && !e->prof && num_iters >= kGraphMinIters) {
    e->dev_iter = true; e->run_top_k = top_k; e->run_burnin = burnin;
    struct Reset { pgibbs_engine* e; ~Reset() { e->dev_iter = false; } } reset{e};
    set_iter_kernel<<<1, 1, 0, e->stream>>>(e->iter_dev, first_iter);
    CK(cudaGetLastError());
    TRY(one_iteration(first_iter));   // also finishes every lazy one-time configuration outside the capture
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    bool ok = cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
    const int64_t counted = e->launches;
    int64_t per_iter = 0;
    if (ok) {
      const int rc = one_iteration(first_iter + 1);   // recorded, not run
      per_iter = e->launches - counted;
      e->launches = counted;
      const cudaError_t ce = cudaStreamEndCapture(e->stream, &graph);
      ok = rc == 0 && ce == cudaSuccess && graph != nullptr;
      if (ok) ok = cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess;
    }
    int done = 1;
    if (ok) {
      for (; done < num_iters; ++done) {
        if (cudaGraphLaunch(exec, e->stream) != cudaSuccess) { ok = false; break; }
        e->launches += per_iter;
      }
    }
    if (exec) cudaGraphExecDestroy(exec);
    if (graph) cudaGraphDestroy(graph);
    if (!ok) {
      (void)cudaGetLastError();
      static bool warned = false;
      if (!warned) { fprintf(stderr, "pgibbs: CUDA graph capture unavailable, launching kernel by kernel\n"); warned = true; }
      for (; done < num_iters; ++done) TRY(one_iteration(first_iter + done));
    }
    return 0;
  }
  for (int it = first_iter; it < first_iter + num_iters; ++it) TRY(one_iteration(it));
  return 0;
}

}  // namespace pg

// =========================================================================================== C ABI
extern "C" {

const char* pgibbs_last_error(void) { return g_err; }
const char* pgibbs_version(void) { return "pgibbs 0.1 sm_100a (tcgen05/TMA GEMM, fp16 operands, fp32 accumulate)"; }

int pgibbs_create(const pgibbs_model_config* cfg, int32_t device_id, pgibbs_engine** out) {
  if (!cfg || !out) return fail("null argument");
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) return fail("no CUDA device found");
  if (device_id < 0 || device_id >= n_dev) return fail("invalid cuda device number: %d", device_id);
  CK(cudaSetDevice(device_id));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device_id));
  if (prop.major != 10) return fail("device %d is sm_%d%d; this engine only runs on sm_100 (B200)", device_id, prop.major, prop.minor);
  if (const char* f = getenv("PGIBBS_GEMM_CG")) g_force_cg = atoi(f);
  if (const char* f = getenv("PGIBBS_GEMM_SPLIT")) g_gemm_split = atoi(f);
  if (const char* f = getenv("PGIBBS_EPI_DIRECT")) g_epi_direct = atoi(f);
  if (const char* f = getenv("PGIBBS_TAIL_OVERLAP")) g_tail_overlap = atoi(f);
  if (const char* f = getenv("PGIBBS_PDL")) g_pdl = atoi(f);
  if (const char* f = getenv("PGIBBS_GRAPH")) g_graph = atoi(f);
  if (const char* f = getenv("PGIBBS_ZIGZAG")) g_zigzag = atoi(f);
  if (cfg->embed_dim % cfg->heads) return fail("embed_dim %% heads != 0");
  if (cfg->embed_dim % 64 || cfg->ffn_dim % 64) return fail("embed_dim and ffn_dim must be multiples of 64");
  if (cfg->embed_dim > kMaxVecPerLane * 128) return fail("embed_dim %d too large (max %d)", cfg->embed_dim, kMaxVecPerLane * 128);
  if (cfg->vocab > 64) return fail("vocab %d too large (max 64)", cfg->vocab);
  pgibbs_engine* e = new pgibbs_engine();
  e->cfg = *cfg;
  e->device = device_id;
  if (cfg->arch == PGIBBS_ARCH_ESM1) {  // ESM1LayerNorm eps, embed_scale = sqrt(embed_dim)
    e->ln_eps = 1e-12f;
    e->embed_scale = sqrtf(static_cast<float>(cfg->embed_dim));
  }
  if (cudaStreamCreateWithFlags(&e->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete e;
    return fail("cudaStreamCreate failed");
  }
  e->stream = e->own_stream;
  if (cudaMalloc(&e->split_flags, kSplitFlagInts * sizeof(int32_t)) != cudaSuccess ||
      cudaMemset(e->split_flags, 0, kSplitFlagInts * sizeof(int32_t)) != cudaSuccess ||
      cudaMalloc(&e->iter_dev, sizeof(int32_t)) != cudaSuccess) {
    pgibbs_destroy(e);
    return fail("cudaMalloc of the GEMM split flags failed");
  }
  *out = e;
  return 0;
}

int pgibbs_destroy(pgibbs_engine* e) {
  if (!e) return 0;
  cudaSetDevice(e->device);
  cudaStreamSynchronize(e->stream);
  free_activations(e);
  for (auto& kv : e->raw) cudaFree(kv.second.first);
  for (void* p : e->owned) cudaFree(p);
  if (e->rope) cudaFree(e->rope);
  if (e->positions) cudaFree(e->positions);
  if (e->noise) cudaFree(e->noise);
  if (e->valid_dev) cudaFree(e->valid_dev);
  cudaFree(e->sc_targets);
  cudaFree(e->sc_logp);
  cudaFree(e->split_flags);
  cudaFree(e->iter_dev);
  for (auto& t : e->prof_pending) { cudaEventDestroy(std::get<1>(t)); cudaEventDestroy(std::get<2>(t)); }
  for (cudaEvent_t ev : e->event_pool) cudaEventDestroy(ev);
  if (e->ev_fork) cudaEventDestroy(e->ev_fork);
  if (e->ev_join) cudaEventDestroy(e->ev_join);
  if (e->side_stream) cudaStreamDestroy(e->side_stream);
  if (e->own_stream) cudaStreamDestroy(e->own_stream);
  delete e;
  return 0;
}

int pgibbs_set_stream(pgibbs_engine* e, void* cuda_stream, int32_t external) {
  if (!e) return fail("null engine");
  CK(cudaSetDevice(e->device));
  CK(cudaStreamSynchronize(e->stream));
  e->stream = external ? static_cast<cudaStream_t>(cuda_stream) : e->own_stream;
  return 0;
}

int pgibbs_load_weight(pgibbs_engine* e, const char* name, const float* data, int64_t numel) {
  if (!e || !name || !data || numel <= 0) return fail("invalid argument to pgibbs_load_weight");
  CK(cudaSetDevice(e->device));
  if (e->finalized) return fail("weights already finalized");
  float* dptr = nullptr;
  TRY(dev_alloc(&dptr, static_cast<size_t>(numel)));
  if (cudaMemcpy(dptr, data, numel * sizeof(float), cudaMemcpyDefault) != cudaSuccess) {
    cudaFree(dptr);
    return fail("copy of weight '%s' failed", name);
  }
  auto it = e->raw.find(name);
  if (it != e->raw.end()) cudaFree(it->second.first);
  e->raw[name] = {dptr, numel};
  return 0;
}

int pgibbs_set_precision(pgibbs_engine* e, int32_t level) {
  if (!e) return fail("null engine");
  if (level < 0 || level > 2) return fail("precision level %d out of range (0 fast, 1 split weights, 2 split weights and activations)", level);
  if (e->finalized) return fail("precision must be chosen before pgibbs_finalize_weights (the weight packing depends on it)");
  e->precision = level;
  return 0;
}

int pgibbs_finalize_weights(pgibbs_engine* e) {
  if (!e) return fail("null engine");
  CK(cudaSetDevice(e->device));
  if (e->finalized) return 0;
  const auto& c = e->cfg;
  const int64_t d = c.embed_dim, F = c.ffn_dim, V = c.vocab;
  if (!raw_get(e, "embed_tokens.weight", V * d)) return 1;
  const bool esm1 = c.arch == PGIBBS_ARCH_ESM1;
  if (esm1) {
    // sinusoidal table (rows 2.. = positions 0..) with one spare row for the bias key/value slot; untied output
    if (!raw_get(e, "embed_positions.weight", (c.max_positions + 3) * d)) return 1;
    if (!raw_get(e, "embed_out", V * d) || !raw_get(e, "embed_out_bias", V)) return 1;
  } else {
    if (c.arch != PGIBBS_ARCH_ESM2) {
      if (!raw_get(e, "embed_positions.weight", (c.max_positions + 2) * d)) return 1;
      if (!raw_get(e, "emb_layer_norm_before.weight", d) || !raw_get(e, "emb_layer_norm_before.bias", d)) return 1;
    }
    if (c.arch == PGIBBS_ARCH_MSA && !raw_get(e, "msa_position_embedding", 1024 * d)) return 1;
    if (!raw_get(e, "emb_layer_norm_after.weight", d) || !raw_get(e, "emb_layer_norm_after.bias", d)) return 1;
    if (!raw_get(e, "lm_head.layer_norm.weight", d) || !raw_get(e, "lm_head.layer_norm.bias", d)) return 1;
    if (!raw_get(e, "lm_head.bias", V)) return 1;
  }
  e->L.resize(c.layers);
  for (int i = 0; i < c.layers; ++i) {
    LayerW& l = e->L[i];
    const std::string p = "layers." + std::to_string(i) + ".";
    std::string attn, ln1, ffn, ln2;
    if (c.arch == PGIBBS_ARCH_MSA) {
      attn = p + "row_self_attention.layer."; ln1 = p + "row_self_attention.layer_norm";
      ffn = p + "feed_forward_layer.layer."; ln2 = p + "feed_forward_layer.layer_norm";
    } else {
      attn = p + "self_attn."; ln1 = p + "self_attn_layer_norm"; ffn = p; ln2 = p + "final_layer_norm";
    }
    TRY(pack_qkv(e, attn, &l.wqkv, &l.bqkv));
    TRY(pack_linear(e, attn + "out_proj", d, d, &l.wo, &l.bo));
    TRY(pack_linear(e, ffn + "fc1", F, d, &l.w1, &l.b1));
    TRY(pack_linear(e, ffn + "fc2", d, F, &l.w2, &l.b2));
    if (!(l.ln1w = raw_get(e, ln1 + ".weight", d)) || !(l.ln1b = raw_get(e, ln1 + ".bias", d))) return 1;
    if (!(l.ln2w = raw_get(e, ln2 + ".weight", d)) || !(l.ln2b = raw_get(e, ln2 + ".bias", d))) return 1;
    if (esm1 && (!(l.bias_k = raw_get(e, attn + "bias_k", d)) || !(l.bias_v = raw_get(e, attn + "bias_v", d)))) return 1;
    if (c.arch == PGIBBS_ARCH_MSA) {
      const std::string ca = p + "column_self_attention.layer.", cl = p + "column_self_attention.layer_norm";
      TRY(pack_qkv(e, ca, &l.c_wqkv, &l.c_bqkv));
      TRY(pack_linear(e, ca + "out_proj", d, d, &l.c_wo, &l.c_bo));
      if (!(l.lncw = raw_get(e, cl + ".weight", d)) || !(l.lncb = raw_get(e, cl + ".bias", d))) return 1;
    }
    CK(cudaStreamSynchronize(e->stream));
    // the fp32 copies of GEMM weights are no longer needed (biases / LN params stay)
    for (const char* nm : {"q_proj", "k_proj", "v_proj", "out_proj"}) {
      drop_raw(e, attn + nm + ".weight");
      if (c.arch == PGIBBS_ARCH_MSA) drop_raw(e, p + "column_self_attention.layer." + nm + ".weight");
    }
    drop_raw(e, ffn + "fc1.weight");
    drop_raw(e, ffn + "fc2.weight");
  }
  if (!esm1) {
    TRY(pack_linear(e, "lm_head.dense", d, d, &e->w_dense, &e->b_dense));
    CK(cudaStreamSynchronize(e->stream));
    drop_raw(e, "lm_head.dense.weight");
  }
  e->finalized = true;
  return 0;
}

int pgibbs_set_tokens(pgibbs_engine* e, const int32_t* tokens, int32_t B, int32_t R, int32_t T) {
  TRY(check_ready(e));
  if (!tokens) return fail("null tokens");
  const size_t n = static_cast<size_t>(B) * R * T;
  std::vector<int32_t> host(n);
  CK(cudaMemcpy(host.data(), tokens, n * sizeof(int32_t), cudaMemcpyDefault));
  for (size_t i = 0; i < n; ++i) {
    if (host[i] < 0 || host[i] >= e->cfg.vocab) return fail("token id %d at %zu outside vocabulary", host[i], i);
    if (host[i] == e->cfg.padding_idx)
      return fail("<pad> token at flat index %zu: padded batches are not on the Gibbs path and are not supported", i);
  }
  TRY(ensure_shape(e, B, R, T));
  if (e->T != T) {  // ESM-1: one extra row per sequence (bias key/value slot); its token only feeds ignored rows
    std::vector<int32_t> padded(static_cast<size_t>(e->M), e->cfg.cls_idx);
    for (int sq = 0; sq < B * R; ++sq)
      std::copy(host.begin() + static_cast<size_t>(sq) * T, host.begin() + static_cast<size_t>(sq + 1) * T,
                padded.begin() + static_cast<size_t>(sq) * e->T);
    host.swap(padded);
  }
  CK(cudaMemcpyAsync(e->tokens, host.data(), host.size() * sizeof(int32_t), cudaMemcpyHostToDevice, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  return 0;
}

int pgibbs_get_tokens(pgibbs_engine* e, int32_t* tokens_out) {
  TRY(check_ready(e));
  if (!e->tokens || !tokens_out) return fail("no tokens resident or null output");
  // [n_seq, Tu] out of the device's [n_seq, T] rows (T = Tu except for ESM-1's extra slot)
  CK(cudaMemcpy2DAsync(tokens_out, static_cast<size_t>(e->Tu) * sizeof(int32_t), e->tokens,
                       static_cast<size_t>(e->T) * sizeof(int32_t), static_cast<size_t>(e->Tu) * sizeof(int32_t),
                       static_cast<size_t>(e->n_seq), cudaMemcpyDefault, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  TRY(prof_flush(e));
  return 0;
}

int pgibbs_set_schedule(pgibbs_engine* e, const int32_t* positions, int64_t numel, int32_t n_iters, int32_t P,
                        int64_t iter_stride, int64_t chain_stride, int32_t has_duplicates) {
  TRY(check_ready(e));
  if (!e->tokens) return fail("set tokens before the schedule");
  if (!positions || numel <= 0 || n_iters <= 0 || P <= 0) return fail("invalid schedule");
  if (P > e->Tu) return fail("P=%d exceeds the sequence length %d", P, e->Tu);
  if (chain_stride < 0 || iter_stride < 0) return fail("negative schedule stride");
  if (static_cast<int64_t>(n_iters - 1) * iter_stride + P > numel) return fail("schedule buffer too small");
  std::vector<int32_t> host(numel);
  CK(cudaMemcpy(host.data(), positions, numel * sizeof(int32_t), cudaMemcpyDefault));
  for (int64_t i = 0; i < numel; ++i)
    if (host[i] < 0 || host[i] >= e->Tu) return fail("scheduled position %d outside [0,%d)", host[i], e->Tu);
  CK(cudaStreamSynchronize(e->stream));
  if (e->positions_numel < numel) {
    if (e->positions) cudaFree(e->positions);
    TRY(dev_alloc(&e->positions, static_cast<size_t>(numel)));
    e->positions_numel = numel;
  }
  CK(cudaMemcpy(e->positions, host.data(), numel * sizeof(int32_t), cudaMemcpyHostToDevice));
  e->positions_used = numel;
  e->n_iters = n_iters; e->P = P; e->iter_stride = iter_stride; e->chain_stride = chain_stride;
  e->has_dup = has_duplicates;
  return 0;
}

int pgibbs_set_noise(pgibbs_engine* e, const float* exp_noise, int64_t numel, int32_t stride) {
  TRY(check_ready(e));
  CK(cudaStreamSynchronize(e->stream));
  if (e->noise) { cudaFree(e->noise); e->noise = nullptr; e->noise_numel = 0; }
  if (!exp_noise || numel <= 0) return 0;
  if (stride <= 0 || stride > 64) return fail("noise stride %d out of range", stride);
  TRY(dev_alloc(&e->noise, static_cast<size_t>(numel)));
  CK(cudaMemcpy(e->noise, exp_noise, numel * sizeof(float), cudaMemcpyDefault));
  e->noise_numel = numel;
  e->noise_stride = stride;
  return 0;
}

int pgibbs_set_device_rng(pgibbs_engine* e, uint64_t seed) {
  if (!e) return fail("null engine");
  e->seed = seed;
  return 0;
}

int pgibbs_set_chain_offset(pgibbs_engine* e, int64_t first_chain) {
  if (!e) return fail("null engine");
  if (first_chain < 0) return fail("negative chain offset");
  e->rng_chain_offset = first_chain;
  return 0;
}

int pgibbs_run(pgibbs_engine* e, int32_t first_iter, int32_t num_iters, int64_t burnin, int32_t top_k,
               float temperature, int32_t mask_flag, const int32_t* valid_ids, int32_t n_valid) {
  return run_iters(e, first_iter, num_iters, burnin, top_k, temperature, mask_flag, 0, 0, false, valid_ids, n_valid);
}

int pgibbs_run_single(pgibbs_engine* e, int32_t first_iter, int32_t num_iters, int64_t burnin, int32_t top_k,
                      float temperature, int32_t mask_row, int32_t target_row, const int32_t* valid_ids,
                      int32_t n_valid) {
  if (!e) return fail("null engine");
  if (e->cfg.arch != PGIBBS_ARCH_MSA) return fail("pgibbs_run_single needs an MSA model");
  if (mask_row < 0) mask_row += e->R;
  if (target_row < 0) target_row += e->R;
  if (mask_row < 0 || mask_row >= e->R || target_row < 0 || target_row >= e->R) return fail("row index out of range");
  return run_iters(e, first_iter, num_iters, burnin, top_k, temperature, 1, mask_row, target_row, true, valid_ids,
                   n_valid);
}

int pgibbs_forward_logits(pgibbs_engine* e, const int32_t* tokens, int32_t B, int32_t R, int32_t T,
                          float* logits_out) {
  TRY(pgibbs_set_tokens(e, tokens, B, R, T));
  if (!logits_out) return fail("null logits_out");
  Schedule s{e->identity_pos, 0, 0, T, 1, 0};
  TRY(forward(e, s, e->n_seq, 0, false, 0, NAN, 0, e->logits));
  CK(cudaMemcpyAsync(logits_out, e->logits, static_cast<size_t>(e->n_seq) * T * e->cfg.vocab * sizeof(float),
                     cudaMemcpyDefault, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  TRY(prof_flush(e));
  return 0;
}

int pgibbs_score(pgibbs_engine* e, const int32_t* targets, int32_t mask, int32_t row, float* logp_out) {
  TRY(check_ready(e));
  if (!e->tokens) return fail("no tokens resident (call pgibbs_set_tokens)");
  if (!e->positions) return fail("no schedule set (call pgibbs_set_schedule)");
  if (!targets || !logp_out) return fail("null argument");
  const bool single = row >= 0;
  if (single && row >= e->R) return fail("row index out of range");
  const int n_chains = single ? e->B : e->n_seq;
  const int64_t rows = static_cast<int64_t>(n_chains) * e->P;
  if (static_cast<int64_t>(n_chains - 1) * e->chain_stride + e->P > e->positions_used)
    return fail("schedule buffer too small for %d chains", n_chains);
  if (rows > e->sc_cap) {
    cudaFree(e->sc_targets); cudaFree(e->sc_logp);
    e->sc_targets = nullptr; e->sc_logp = nullptr; e->sc_cap = 0;
    TRY(dev_alloc(&e->sc_targets, static_cast<size_t>(rows)));
    TRY(dev_alloc(&e->sc_logp, static_cast<size_t>(rows)));
    e->sc_cap = rows;
  }
  CK(cudaMemcpyAsync(e->sc_targets, targets, rows * sizeof(int32_t), cudaMemcpyDefault, e->stream));
  Schedule s{e->positions, e->iter_stride, e->chain_stride, e->P, single ? e->R : 1, single ? row : 0};
  if (mask) {  // the strided <mask> copies are built on the device from the resident (unmasked) tokens
    ProfScope ps(e, "mask_scatter");
    mask_scatter_kernel<<<static_cast<unsigned>((rows + 255) / 256), 256, 0, e->stream>>>(e->tokens, n_chains, e->T, s,
                                                                                        0, e->cfg.mask_idx);
    CK(cudaGetLastError());
  }
  e->sc_active = true;
  const int rc = forward(e, s, n_chains, 0, false, 0, NAN, 0, nullptr);
  e->sc_active = false;
  if (rc) return rc;
  CK(cudaMemcpyAsync(logp_out, e->sc_logp, rows * sizeof(float), cudaMemcpyDefault, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  TRY(prof_flush(e));
  return 0;
}

int pgibbs_sync(pgibbs_engine* e) {
  if (!e) return fail("null engine");
  CK(cudaSetDevice(e->device));
  CK(cudaStreamSynchronize(e->stream));
  TRY(prof_flush(e));
  return 0;
}

int pgibbs_debug_read(pgibbs_engine* e, const char* which, float* out, int64_t numel) {
  TRY(check_ready(e));
  CK(cudaStreamSynchronize(e->stream));
  const int64_t M = e->M, d = e->cfg.embed_dim, F = e->cfg.ffn_dim;
  const std::string w = which ? which : "";
  const __half* src16 = nullptr;
  const float* src32 = nullptr;
  int64_t n = 0;
  if (w == "x") { src32 = e->x; n = M * d; }
  else if (w == "g") { src32 = e->g; n = M * d; }
  else if (w == "h") { src16 = e->h; n = M * d * e->ak(); }     // split-operand mode: rows are [hi | lo]
  else if (w == "hs") { src16 = e->hs; n = M * d * e->ak(); }
  else if (w == "qkv") { src16 = e->qkv; n = M * 3 * d; }
  else if (w == "ctx") { src16 = e->ctx; n = M * d * e->ak(); }
  else if (w == "ffn") { src16 = e->ffn; n = M * F * e->ak(); }
  else return fail("unknown debug buffer '%s'", w.c_str());
  if (numel > n) numel = n;
  if (src32) {
    CK(cudaMemcpy(out, src32, numel * sizeof(float), cudaMemcpyDefault));
  } else {
    float* tmp = nullptr;
    TRY(dev_alloc(&tmp, static_cast<size_t>(numel)));
    f16_to_f32_kernel<<<static_cast<unsigned>((numel + 255) / 256), 256, 0, e->stream>>>(src16, tmp, numel);
    cudaError_t er = cudaStreamSynchronize(e->stream);
    if (er == cudaSuccess) er = cudaMemcpy(out, tmp, numel * sizeof(float), cudaMemcpyDefault);
    cudaFree(tmp);
    if (er != cudaSuccess) return fail("debug read failed: %s", cudaGetErrorString(er));
  }
  return 0;
}

int pgibbs_debug_tail_plan(int32_t M, int32_t N, int32_t block_n, int32_t cta_group, int32_t sms, int32_t reverse,
                           int32_t* out4) {
  if (M <= 0 || N <= 0 || block_n <= 0 || (cta_group != 1 && cta_group != 2) || sms < cta_group || !out4)
    return fail("pgibbs_debug_tail_plan: invalid arguments");
  const TailPlan t = tail_plan(M, N, block_n, cta_group, sms, reverse);
  out4[0] = t.full; out4[1] = t.rem; out4[2] = t.lo; out4[3] = t.hi;
  return 0;
}

int pgibbs_debug_layer_limit(pgibbs_engine* e, int32_t n_layers) {
  if (!e) return fail("null engine");
  e->layer_limit = n_layers;
  return 0;
}

int pgibbs_profile_enable(pgibbs_engine* e, int32_t on) {
  if (!e) return fail("null engine");
  TRY(prof_flush(e));
  e->prof = on != 0;
  if (on) e->prof_acc.clear();
  return 0;
}

int pgibbs_profile_read(pgibbs_engine* e, char (*names)[32], float* total_ms, int32_t* launches, int32_t cap,
                        int32_t* n) {
  if (!e || !n) return fail("null argument");
  TRY(prof_flush(e));
  int i = 0;
  for (auto& kv : e->prof_acc) {
    if (i >= cap) break;
    snprintf(names[i], 32, "%s", kv.first.c_str());
    total_ms[i] = static_cast<float>(kv.second.ms);
    launches[i] = kv.second.launches;
    ++i;
  }
  *n = i;
  return 0;
}

int64_t pgibbs_launch_count(pgibbs_engine* e) { return e ? e->launches : 0; }

int pgibbs_debug_attention_trace(uint64_t* out, int32_t enable) {
  if (enable) {
    if (!g_fa_trace) CK(cudaMalloc(&g_fa_trace, 4 * kFaTraceCap * sizeof(unsigned long long)));
    CK(cudaMemset(g_fa_trace, 0, 4 * kFaTraceCap * sizeof(unsigned long long)));
    return 0;
  }
  if (!g_fa_trace) return fail("attention trace was not enabled");
  CK(cudaDeviceSynchronize());
  if (out) CK(cudaMemcpy(out, g_fa_trace, 4 * kFaTraceCap * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  CK(cudaFree(g_fa_trace));
  g_fa_trace = nullptr;
  return 0;
}

// ------------------------------------------------------------------------ stand-alone operator entry points
static int op_device(int device_id) {
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) return fail("no CUDA device found");
  if (device_id < 0 || device_id >= n_dev) return fail("invalid cuda device number: %d", device_id);
  CK(cudaSetDevice(device_id));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device_id));
  if (prop.major != 10) return fail("device %d is not sm_100", device_id);
  return 0;
}

int pgibbs_op_gemm(int32_t device_id, const float* A, const float* B, const float* bias, float* C, int32_t M,
                   int32_t N, int32_t K, int32_t epilogue, int32_t block_n, int32_t cta_group, float* elapsed_ms,
                   int32_t reps) {
  TRY(op_device(device_id));
  if (epilogue == EPI_QKV_F16) return fail("use the engine for the QKV epilogue");
  const size_t na = static_cast<size_t>(M) * K, nb = static_cast<size_t>(N) * K, nc = static_cast<size_t>(M) * N;
  float *dA = nullptr, *dB = nullptr, *dbias = nullptr, *dC32 = nullptr;
  __half *hA = nullptr, *hB = nullptr, *dC16 = nullptr;
  int32_t* dflags = nullptr;
  int rc = 0;
  cudaStream_t st = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (const char* f = getenv("PGIBBS_GEMM_SPLIT")) g_gemm_split = atoi(f);
  if (const char* f = getenv("PGIBBS_EPI_DIRECT")) g_epi_direct = atoi(f);
  auto body = [&]() -> int {
    TRY(dev_alloc(&dflags, static_cast<size_t>(kSplitFlagInts)));
    CK(cudaMemset(dflags, 0, kSplitFlagInts * sizeof(int32_t)));
    TRY(dev_alloc(&dA, na)); TRY(dev_alloc(&dB, nb)); TRY(dev_alloc(&hA, na)); TRY(dev_alloc(&hB, nb));
    TRY(dev_alloc(&dC32, nc)); TRY(dev_alloc(&dC16, nc));
    CK(cudaMemcpy(dA, A, na * sizeof(float), cudaMemcpyDefault));
    CK(cudaMemcpy(dB, B, nb * sizeof(float), cudaMemcpyDefault));
    if (bias) { TRY(dev_alloc(&dbias, static_cast<size_t>(N))); CK(cudaMemcpy(dbias, bias, N * sizeof(float), cudaMemcpyDefault)); }
    CK(cudaStreamCreate(&st));
    TRY(to_f16(dA, hA, na, st)); TRY(to_f16(dB, hB, nb, st));
    const bool out16 = epilogue == EPI_BIAS_F16 || epilogue == EPI_GELU_F16;
    if (epilogue == EPI_RESID_F32) CK(cudaMemcpyAsync(dC32, C, nc * sizeof(float), cudaMemcpyDefault, st));
    GemmPlan plan = pick_gemm_plan(M, N, 16);
    if (block_n > 0) plan.bn = block_n;
    if (cta_group > 0) plan.cg = cta_group;
    CUtensorMap ma, mb;
    TRY(make_tmap_2d(&ma, hA, M, K, K, kBM));
    TRY(make_tmap_2d(&mb, hB, N, K, K, plan.b_box()));
    GemmParams p = gp(M, N, K, dbias, out16 ? static_cast<void*>(dC16) : static_cast<void*>(dC32), N);
    p.flags = dflags;
    auto launch_plan = [&]() -> int { return launch_gemm(epilogue, plan, ma, mb, p, st); };
    TRY(launch_plan());
    CK(cudaStreamSynchronize(st));
    if (out16) {
      f16_to_f32_kernel<<<static_cast<unsigned>((nc + 255) / 256), 256, 0, st>>>(dC16, dC32, nc);
      CK(cudaGetLastError());
    }
    CK(cudaMemcpyAsync(C, dC32, nc * sizeof(float), cudaMemcpyDefault, st));
    CK(cudaStreamSynchronize(st));
    if (elapsed_ms && reps > 0) {
      CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
      for (int i = 0; i < 3; ++i) TRY(launch_plan());
      CK(cudaEventRecord(e0, st));
      for (int i = 0; i < reps; ++i) TRY(launch_plan());
      CK(cudaEventRecord(e1, st));
      CK(cudaStreamSynchronize(st));
      float ms = 0;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      *elapsed_ms = ms / reps;
    }
    return 0;
  };
  rc = body();
  for (void* p : {static_cast<void*>(dA), static_cast<void*>(dB), static_cast<void*>(dbias), static_cast<void*>(dC32),
                  static_cast<void*>(hA), static_cast<void*>(hB), static_cast<void*>(dC16), static_cast<void*>(dflags)})
    if (p) cudaFree(p);
  if (e0) cudaEventDestroy(e0);
  if (e1) cudaEventDestroy(e1);
  if (st) cudaStreamDestroy(st);
  return rc;
}

int pgibbs_op_attention(int32_t device_id, const float* qkv, float* ctx, int32_t n_seq, int32_t T, int32_t heads,
                        int32_t head_dim, float* elapsed_ms, int32_t reps) {
  TRY(op_device(device_id));
  const int d = heads * head_dim;
  const size_t nq = static_cast<size_t>(n_seq) * T * 3 * d, nc = static_cast<size_t>(n_seq) * T * d;
  float *d32 = nullptr, *c32 = nullptr;
  __half *d16 = nullptr, *c16 = nullptr;
  auto body = [&]() -> int {
    TRY(dev_alloc(&d32, nq)); TRY(dev_alloc(&d16, nq)); TRY(dev_alloc(&c32, nc)); TRY(dev_alloc(&c16, nc));
    CK(cudaMemcpy(d32, qkv, nq * sizeof(float), cudaMemcpyDefault));
    TRY(to_f16(d32, d16, nq, nullptr));
    CUtensorMap m3, c3;
    const int mode = attn_mode(head_dim);
    if (mode) {
      TRY(make_tmap_qkv3(&m3, d16, n_seq, T, 3 * d));
      TRY(make_tmap_qkv3(&c3, c16, n_seq, T, d));
    }
    auto launch = [&]() -> int {
      if (mode == 2) return launch_attention_fa(m3, c3, d16, c16, n_seq, T, heads, nullptr);
      AttnParams p{d16, c16, T, 3 * d, d, d, 2 * d, 1, 0, 1, T};
      return launch_attention(p, n_seq, heads, head_dim, nullptr);
    };
    TRY(launch());
    if (elapsed_ms && reps > 0) {
      cudaEvent_t e0, e1;
      CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
      for (int i = 0; i < 3; ++i) TRY(launch());
      CK(cudaEventRecord(e0, nullptr));
      for (int i = 0; i < reps; ++i) TRY(launch());
      CK(cudaEventRecord(e1, nullptr));
      CK(cudaDeviceSynchronize());
      float ms = 0;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      *elapsed_ms = ms / reps;
      cudaEventDestroy(e0); cudaEventDestroy(e1);
    }
    f16_to_f32_kernel<<<static_cast<unsigned>((nc + 255) / 256), 256>>>(c16, c32, nc);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(ctx, c32, nc * sizeof(float), cudaMemcpyDefault));
    return 0;
  };
  const int rc = body();
  for (void* p : {static_cast<void*>(d32), static_cast<void*>(c32), static_cast<void*>(d16), static_cast<void*>(c16)})
    if (p) cudaFree(p);
  return rc;
}

int pgibbs_op_sample(int32_t device_id, const float* logits, const float* noise, int32_t rows, int32_t vocab,
                     const int32_t* valid_ids, int32_t n_valid, int32_t top_k, float temperature,
                     int32_t* tokens_out) {
  TRY(op_device(device_id));
  if (vocab > 64 || n_valid > 64 || n_valid <= 0) return fail("vocab <= 64 and 0 < n_valid <= 64 required");
  float *dl = nullptr, *dn = nullptr;
  int32_t *dv = nullptr, *dout = nullptr;
  auto body = [&]() -> int {
    TRY(dev_alloc(&dl, static_cast<size_t>(rows) * vocab));
    TRY(dev_alloc(&dv, static_cast<size_t>(n_valid)));
    TRY(dev_alloc(&dout, static_cast<size_t>(rows)));
    CK(cudaMemcpy(dl, logits, static_cast<size_t>(rows) * vocab * sizeof(float), cudaMemcpyDefault));
    CK(cudaMemcpy(dv, valid_ids, n_valid * sizeof(int32_t), cudaMemcpyDefault));
    if (noise) {
      TRY(dev_alloc(&dn, static_cast<size_t>(rows) * n_valid));
      CK(cudaMemcpy(dn, noise, static_cast<size_t>(rows) * n_valid * sizeof(float), cudaMemcpyDefault));
    }
    const int k = (top_k <= 0 || top_k > n_valid) ? n_valid : top_k;
    sample_rows_kernel<<<(rows + 7) / 8, 256>>>(dl, rows, vocab, dv, n_valid, k, temperature, dn, n_valid, 1234ull, dout);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(tokens_out, dout, rows * sizeof(int32_t), cudaMemcpyDefault));
    return 0;
  };
  const int rc = body();
  for (void* p : {static_cast<void*>(dl), static_cast<void*>(dn), static_cast<void*>(dv), static_cast<void*>(dout)})
    if (p) cudaFree(p);
  return rc;
}

}  // extern "C"

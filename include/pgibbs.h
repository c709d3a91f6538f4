/* pgibbs -- C ABI of the B200-native Gibbs-sampling engine for masked protein language models.
 *
 * The reference (seanrjohnson/protein_gibbs_sampler) has no FFI of its own: its hot path is the Python loop
 * body of ESM_sampler.generate (src/pgen/esm_sampler.py:209-234) and ESM_MSA_sampler.generate /
 * generate_single (src/pgen/esm_msa_sampler.py:221-248, 126-145), which calls a fair-esm nn.Module.  These
 * entry points are what a binding for that loop body binds; each cites the reference lines it replaces.
 * Plain pointers and sizes only; the caller owns every buffer it passes; pointers may be host or device
 * addresses (copies use cudaMemcpyDefault).  One engine per GPU; calls on one engine must be serialised by
 * the caller.  Every function returns 0 on success, non-zero on failure with pgibbs_last_error() set.
 */
#ifndef PGIBBS_H
#define PGIBBS_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pgibbs_engine pgibbs_engine;

/* ESM1B: ESM-1b / ESM-1v (learned positions); ESM2: rotary; MSA: MSA Transformer; ESM1: esm1_t6/t12/t34 (sinusoidal
 * positions supplied as the `embed_positions.weight` table, sqrt(d) embedding scale, bias key/value, untied output). */
enum { PGIBBS_ARCH_ESM1B = 0, PGIBBS_ARCH_ESM2 = 1, PGIBBS_ARCH_MSA = 2, PGIBBS_ARCH_ESM1 = 3 };

/* Geometry of the model behind `model.model(batch)["logits"]` (esm_sampler.py:223, esm_msa_sampler.py:236). */
typedef struct {
  int32_t arch;          /* PGIBBS_ARCH_* */
  int32_t layers, embed_dim, heads, ffn_dim, vocab, max_positions;
  int32_t token_dropout; /* fair-esm token-dropout rescale (ESM-1b, ESM-2) */
  int32_t padding_idx, mask_idx, cls_idx, eos_idx; /* model.alphabet.* (esm_sampler.py:82,262) */
} pgibbs_model_config;

/* Thread-local description of the last failure on the calling thread. */
const char* pgibbs_last_error(void);
/* Library / build identification ("pgibbs <ver> sm_100a ..."). */
const char* pgibbs_version(void);

/* ESM_sampler.__init__ / ESM_MSA_sampler.__init__ device placement: `model.to(device)`
 * (esm_sampler.py:68-80, esm_msa_sampler.py:51-63).  Fails if no sm_100 GPU is present. */
int pgibbs_create(const pgibbs_model_config* cfg, int32_t device_id, pgibbs_engine** out);
int pgibbs_destroy(pgibbs_engine* e);
/* external != 0: launch on the caller's CUDA stream `cuda_stream` (a cudaStream_t; NULL is the CUDA default
 * stream).  external == 0: back to the engine's own non-blocking stream.  Inside a forward the engine may fork part of
 * the work to a private second stream (LayerNorm next to a GEMM's last wave); every fork is joined back on THIS stream
 * before the next dependent kernel, so stream order as seen by the caller is unchanged. */
int pgibbs_set_stream(pgibbs_engine* e, void* cuda_stream, int32_t external);

/* One fp32 tensor of the fair-esm state dict by key name (models.py:61-86 bind the loaders). */
int pgibbs_load_weight(pgibbs_engine* e, const char* name, const float* data, int64_t numel);
/* Numerics of the forward that replaces `self.model.model(batch)` (esm_sampler.py:223; fp32 in the reference).
 *   0 (default) every GEMM is one tensor-core pass over fp16 operands, fp32 accumulate;
 *   1 weights are carried as fp16 hi + lo pairs (exact to 2^-22), two passes;
 *   2 the GEMM input activations too (three passes: a_hi w_hi + a_hi w_lo + a_lo w_hi) -- logits within 1e-3 of fp32
 *     per token row at 33 layers, at about 2.5x the step time.
 * Must be called before pgibbs_finalize_weights (the weight packing depends on it). */
int pgibbs_set_precision(pgibbs_engine* e, int32_t level);
/* Check that every tensor the architecture needs was supplied; pack GEMM operands to fp16. */
int pgibbs_finalize_weights(pgibbs_engine* e);

/* `batch = get_init_seq(...)` / `get_init_msa(...)` then `batch.cuda()` (esm_sampler.py:201-202,
 * esm_msa_sampler.py:212-213): int32 tokens [B, R, T]; R = 1 for single-sequence models. */
int pgibbs_set_tokens(pgibbs_engine* e, const int32_t* tokens, int32_t B, int32_t R, int32_t T);
/* Final `batch` read back for untokenize_batch (esm_sampler.py:236-239). */
int pgibbs_get_tokens(pgibbs_engine* e, int32_t* tokens_out);

/* Target positions for every iteration, pre-drawn by the host in the reference's RNG order
 * (get_random_target_index / get_target_index_in_order / all positions: esm_sampler.py:210-218,242-257):
 * position of slot p of chain c at iteration i is positions[i*iter_stride + c*chain_stride + p], p < P.
 * A chain is one sequence row (B*R chains).  Strides of 0 share one list.  `has_duplicates` != 0 keeps the
 * reference's last-write-wins order for user `indexes` lists with repeats. */
int pgibbs_set_schedule(pgibbs_engine* e, const int32_t* positions, int64_t numel, int32_t n_iters, int32_t P,
                        int64_t iter_stride, int64_t chain_stride, int32_t has_duplicates);
/* Replay mode: Exp(1) draws consumed by Categorical.sample() inside generate_step (esm_sampler.py:41-43),
 * laid out [n_iters][n_chains*P][stride]; slot j < stride pairs with the j-th largest candidate logit
 * (stride >= the largest effective k of any iteration).  NULL clears. */
int pgibbs_set_noise(pgibbs_engine* e, const float* exp_noise, int64_t numel, int32_t stride);
/* Device RNG (Philox4x32-10) used when no replay noise is set. */
int pgibbs_set_device_rng(pgibbs_engine* e, uint64_t seed);
/* Chains sharded across GPUs: global index of this engine's first chain, so that the device RNG draws for chain c of
 * this shard what a single engine holding all chains would draw for chain first_chain + c. */
int pgibbs_set_chain_offset(pgibbs_engine* e, int64_t first_chain);

/* The loop body esm_sampler.py:209-234 (esm_msa_sampler.py:221-248) for iterations
 * [first_iter, first_iter + num_iters): <mask> scatter (if mask_flag) -> forward -> generate_step at every
 * scheduled position -> write-back, entirely on device, no host synchronisation between iterations.
 * burnin: iterations with index < burnin sample the full candidate set (generate_step `sample=True`);
 * top_k <= 0 or > n_valid also means the full set.  temperature = NaN means None; any other value divides the logits (esm_sampler.py:24-25: a negative one inverts the ranking; 0 is rejected).  Asynchronous. */
int pgibbs_run(pgibbs_engine* e, int32_t first_iter, int32_t num_iters, int64_t burnin, int32_t top_k,
               float temperature, int32_t mask_flag, const int32_t* valid_ids, int32_t n_valid);
/* ESM_MSA_sampler.generate_single's step (esm_msa_sampler.py:132-145): mask `mask_row` of every MSA at the
 * scheduled positions, forward, then sample/write `target_row` only.  Schedule chains are MSAs here. */
int pgibbs_run_single(pgibbs_engine* e, int32_t first_iter, int32_t num_iters, int64_t burnin, int32_t top_k,
                      float temperature, int32_t mask_row, int32_t target_row, const int32_t* valid_ids,
                      int32_t n_valid);
/* `model.model(tokens)["logits"]` for parity checks and log-likelihood scoring: fp32 [B, R, T, vocab].
 * Leaves `tokens` resident. */
int pgibbs_forward_logits(pgibbs_engine* e, const int32_t* tokens, int32_t B, int32_t R, int32_t T,
                          float* logits_out);
/* Pseudo-log-likelihood pass: replaces the body of log_likelihood_batch (/root/reference/src/pgen/esm_sampler.py:
 * 316-352, esm_msa_sampler.py:371-424: clone + strided <mask> fill, forward, log_softmax, gather at the true token).
 * With the UNMASKED tokens and a one-iteration schedule resident (chain c scores positions[c*chain_stride .. +P)),
 * optionally writes <mask> at the scheduled positions on the device, runs the forward with the LM head on the
 * scheduled rows only, and returns log_softmax(logits over the whole vocabulary)[targets[slot]] per slot.
 * targets / logp_out: [n_chains * P]; a target < 0 marks a padding slot (its position must still be valid; result 0).
 * row >= 0: chains are MSAs and only row `row` of each is masked and scored; row < 0: every sequence is a chain.
 * The resident tokens are left masked. */
int pgibbs_score(pgibbs_engine* e, const int32_t* targets, int32_t mask, int32_t row, float* logp_out);
/* Block until all queued work of this engine has finished; reports asynchronous kernel failures. */
int pgibbs_sync(pgibbs_engine* e);

/* Debug/parity taps: copy an internal activation buffer ("x", "h", "qkv", "ctx", "ffn", "g") as fp32. */
int pgibbs_debug_read(pgibbs_engine* e, const char* which, float* out, int64_t numel);
/* Host arithmetic only (no GPU): where the engine cuts a residual GEMM of M x N outputs tiled (128 * cta_group) x block_n
 * on `sms` SMs so that the following LayerNorm can start next to the partly-filled last wave.
 * out4 = {tiles in the full waves, tiles in the last wave, first row, one-past-last row of the finished row blocks}. */
int pgibbs_debug_tail_plan(int32_t M, int32_t N, int32_t block_n, int32_t cta_group, int32_t sms, int32_t reverse,
                           int32_t* out4);
/* Stop the forward after `n_layers` transformer layers (negative = all); parity bisecting only. */
int pgibbs_debug_layer_limit(pgibbs_engine* e, int32_t n_layers);

/* Per-kernel-class device timing (CUDA events on the engine's stream) for the roofline report.
 * pgibbs_profile_read fills up to `cap` entries; returns the number of classes via *n. */
int pgibbs_profile_enable(pgibbs_engine* e, int32_t on);
int pgibbs_profile_read(pgibbs_engine* e, char (*names)[32], float* total_ms, int32_t* launches, int32_t cap,
                        int32_t* n);
/* Kernels launched by this engine since creation (bench `gpu_launches`). */
int64_t pgibbs_launch_count(pgibbs_engine* e);

/* Debug: timeline of the tcgen05 attention kernel's CTA 0.  enable != 0 arms it for the following attention launches;
 * enable == 0 copies 4 x 2048 words of (clock64 << 8 | event code) to `out` (0 = unused slot) and disarms. */
int pgibbs_debug_attention_trace(uint64_t* out, int32_t enable);

/* Stand-alone operator entry points (unit parity tests; same kernels the engine launches).
 * gemm: C[M,N] = epi(A[M,K] . B[N,K]^T + bias) with fp32 host/device inputs rounded to fp16 operands.
 * epilogue: 0 bias->fp16, 1 gelu->fp16, 2 residual add into C (fp32), 4 gelu->fp32, 5 bias->fp32. */
int pgibbs_op_gemm(int32_t device_id, const float* A, const float* B, const float* bias, float* C, int32_t M,
                   int32_t N, int32_t K, int32_t epilogue, int32_t block_n, int32_t cta_group, float* elapsed_ms,
                   int32_t reps);
/* attention over fused qkv [n_seq*T, 3*heads*head_dim] (fp32 in, rounded to fp16) -> ctx [n_seq*T, heads*head_dim]. */
int pgibbs_op_attention(int32_t device_id, const float* qkv, float* ctx, int32_t n_seq, int32_t T, int32_t heads,
                        int32_t head_dim, float* elapsed_ms, int32_t reps);
/* generate_step on given logits rows [rows, vocab] with given Exp(1) noise [rows, n_valid] (NULL: device RNG)
 * -> token ids. */
int pgibbs_op_sample(int32_t device_id, const float* logits, const float* noise, int32_t rows, int32_t vocab,
                     const int32_t* valid_ids, int32_t n_valid, int32_t top_k, float temperature,
                     int32_t* tokens_out);

#ifdef __cplusplus
}
#endif
#endif /* PGIBBS_H */

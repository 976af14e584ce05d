/* sduss_b200 -- C ABI of the B200-native (sm_100a) denoising-step kernels.
 *
 * This is the drop-in boundary for the hot path of MiRaCLeXeoN/sduss ("Mixfusion"): the
 * batched mixed-resolution denoising step. It supersedes the reference's only native
 * interface, the pybind11 module `esymred_mp`
 * (sduss/model_executor/modules/kernels/norm_silu_concat.cpp:66-106), and the third-party
 * kernels its Python hot path reaches through torch / xformers / diffusers.
 *
 * Conventions (all entry points):
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless stated;
 *   - `stream` is a cudaStream_t passed as void*; nothing here synchronises or allocates;
 *   - the return value is 0 on success, a cudaError_t value on a CUDA failure, or one of the
 *     B200_ERR_* codes below for rejected arguments. The reference only printed CUDA errors
 *     (norm_silu_concat.cu:434-437); callers of this ABI must turn non-zero into an exception.
 *   - activations are bf16 row-major "packed" buffers: rows = tokens / pixels of all requests
 *     back to back, columns = channels. Integer descriptor tables are int32.
 *   - scratch memory is always the caller's: the four entry points that need any come with a
 *     *_bytes query (b200_attn_workspace_bytes, b200_conv3x3_maps_bytes,
 *     b200_groupnorm_workspace_bytes, b200_patch_mask_workspace_bytes); every other entry point
 *     needs none.
 */
#ifndef SDUSS_B200_H_
#define SDUSS_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_OK 0
#define B200_ERR_INVALID 10001
#define B200_ERR_DRIVER 10002
#define B200_ERR_UNSUPPORTED 10003

/* ABI version; bumped on any signature change. */
int b200_version(void);
/* Number of SMs of the current device (148 on B200); 0 when no device is usable. */
int b200_sm_count(void);

/* ------------------------------------------------------------------------------------------
 * Linear layers: C = epilogue(A[M,K] * W[N,K]^T), bf16 in, fp32 accumulate (tcgen05 + TMA).
 * Replaces F.linear at sduss/model_executor/modules/resnet.py:163 (SplitLinear) and the
 * diffusers Attention / FeedForward / AdaLayerNorm linears called from
 * modules/attention.py:73-96,259-274,411 and modules/transformer.py:185-279,317-386.
 * ---------------------------------------------------------------------------------------- */
enum B200EpilogueMode {
  B200_EPI_BIAS = 0,       /* C = acc + bias                                              */
  B200_EPI_GELU_TANH = 1,  /* C = gelu_tanh(acc + bias)            (SD3 FeedForward)      */
  B200_EPI_GATE_RESID = 2, /* C = resid + gate[row_group[row]] * (acc + bias)             */
  B200_EPI_QK_RMSNORM = 3, /* per 64-wide head: RMSNorm(q)*q_scale, RMSNorm(k), v as is   */
  B200_EPI_GEGLU = 4,      /* W rows interleaved [32 hidden | 32 gate]; C is [M, N/2]     */
  B200_EPI_ROWVEC = 5      /* C = acc + bias + rowvec[row_group[row]]                     */
};

typedef struct B200EpilogueDesc {
  void* C;              /* [M, ldc] bf16, or fp32 when out_fp32 != 0                      */
  int32_t ldc;
  int32_t out_fp32;
  const void* bias;     /* [N] bf16 or NULL                                               */
  const void* resid;    /* [M, ldr] bf16 or NULL (may alias C)                            */
  int32_t ldr;
  const void* gate;     /* [G, ldg] bf16 or NULL (NULL = 1)                               */
  int32_t ldg;
  const int32_t* row_group; /* [M] request id per row; needed with gate / rowvec          */
  const void* rowvec;   /* [G, ldv] bf16                                                  */
  int32_t ldv;
  const void* rms_wq;   /* [64] bf16 RMSNorm weight for q heads                           */
  const void* rms_wk;   /* [64] bf16 RMSNorm weight for k heads                           */
  int32_t rms_q_cols;   /* columns [0, q_cols) are q heads                                */
  int32_t rms_k_cols;   /* columns [q_cols, q_cols + k_cols) are k heads                  */
  float rms_eps;
  float q_scale;        /* folded into normalised q (softmax scale * log2(e))             */
  int32_t act;          /* activation of B200_EPI_GELU_TANH / gate of B200_EPI_GEGLU:     */
                        /* 0 = the mode's default (tanh-GELU / erf-GELU), 1 tanh-GELU,    */
                        /* 2 erf-GELU, 3 quick-GELU x*sigmoid(1.702x) (text encoders)     */
  const int32_t* row_mask; /* patch cache, may be NULL: device int32 per chunk   */
  int32_t row_mask_shift;  /* of 2^shift rows (shift >= 8); M tiles of a chunk whose entry is 0   */
                           /* are skipped: no loads, no MMA, their rows of C stay as they are     */
  float* stats_out;        /* b200_conv3x3_bf16 only, may be NULL: [n_mtiles * 2][Cout][2] fp32,  */
                           /* per (16x8 output tile, row half, channel) the (sum, sum of squares) */
                           /* of the stored bf16 outputs: the statistics pass of the GroupNorm    */
                           /* that follows (b200_groupnorm_from_conv_stats), fused                */
  const float* ln_stats;   /* b200_gemm_bf16 only, may be NULL: [M][2] fp32 (mean, rstd) per row of A */
  const float* ln_colsum;  /* [N] fp32. LayerNorm FOLDED into the GEMM: A is the un-normalised x, W    */
                           /* carries gamma (W' = W o gamma), colsum[n] = sum_k W'[n, k], bias carries  */
                           /* beta W^T: acc <- rstd * (acc - mean * colsum) before the epilogue mode,   */
                           /* so C = epilogue(LN(x) W^T + b) without a LayerNorm pass (b200_row_stats)  */
  const float* ln_rowpart; /* ... or, instead of ln_stats, the partial sums the PRODUCER of A left:    */
  int32_t ln_nparts;       /* [ln_nparts][M][2] fp32 (sum, sum of squares) per 64-column chunk of a    */
  float ln_eps;            /* row of A, ln_nparts = K / 64; (mean, rstd) are reduced once per tile      */
  float* rowpart_out;      /* producer side, may be NULL: [ceil(N/64)][M][2] partial sums of the rows   */
                           /* of C (not with B200_EPI_GEGLU / fp32 outputs): no statistics kernel at all */
  int32_t w_static;        /* 1: nothing on `stream` writes W (weights): the kernel may request its first */
                           /* W tiles before waiting for the previous kernel (programmatic dependent       */
                           /* launch); 0 (default): W is loaded after the wait, like A                     */
  int32_t row_mask_scale;  /* b200_conv3x3_bf16 only: log2(rows of the level row_mask describes / output   */
                           /* rows): 0 same level, +2 a stride-2 downsampler, -2 the convolution after a    */
                           /* 2x upsample. A 16 x 8 pixel block is skipped when every patch its 16 pixel    */
                           /* rows touch is clean (reference: masked conv1 / conv2 / samplers,              */
                           /* modules/resnet.py:328-339,365-377,414-454)                                    */
} B200EpilogueDesc;

/* A: [M, lda] bf16, W: [N, ldw] bf16 (K contiguous in both). N, K, lda, ldw, ldc multiples
 * of 8; base pointers 16-byte aligned. */
int b200_gemm_bf16(const void* A, int lda, const void* W, int ldw, int M, int N, int K,
                   int epi_mode, const B200EpilogueDesc* ep, void* stream);

/* ------------------------------------------------------------------------------------------
 * Packed variable-length attention, head_dim 64, non-causal, bf16 (tcgen05 + TMEM + TMA).
 * One launch for all requests of the batch. Each sequence's queries and keys/values are the
 * concatenation of up to two row segments ("A" then "B") that live in two source buffers:
 *   SD3 joint attention : A = image tokens, B = context tokens (Q, K and V from both)
 *   SD3 attn2, SDXL self: A only
 *   SDXL cross          : Q from A (image tokens), K/V from B (text tokens)
 * Replaces xformers.ops.memory_efficient_attention / F.scaled_dot_product_attention and the
 * per-resolution Python loops at sduss/model_executor/modules/attention.py:86,155-203,297-368.
 * ---------------------------------------------------------------------------------------- */
typedef struct B200AttnSource {
  const void* q;  int32_t ldq;  int32_t q_col;   /* [q_rows, ldq] bf16; head h at q_col + 64 h  */
  int32_t q_rows;
  const void* k;  int32_t ldk;  int32_t k_col;   /* [kv_rows, ldk]                              */
  const void* v;  int32_t ldv;  int32_t v_col;   /* [kv_rows, ldv]                              */
  int32_t kv_rows;
  void* out;      int32_t ldo;  int32_t o_col;   /* [q_rows, ldo] attention output of this side */
} B200AttnSource;

/* seq_table: int32 [n_seq][8] = {qa_row, qa_len, qb_row, qb_len, ka_row, ka_len, kb_row, kb_len}
 *            (row offsets into the A / B source buffers; a length of 0 disables the segment).
 * The kernel is persistent (one CTA per SM, or max_ctas of them if max_ctas > 0). Its work list
 * comes from b200_attn_build_schedule, a pure host function (no GPU needed) run once per batch
 * composition:
 *   work_units:  int32 [n_units][4] = {seq, q_segment (0 = A, 1 = B), row offset of the query
 *                block inside that segment, head}, longest first; a query block covers
 *                b200_attn_rows_per_item() consecutive rows (a build constant).
 * Call it with work_units == NULL to get n_units, allocate, call again (host memory; seq_table
 * there is the HOST copy), then copy the table to the device.
 * sched_state: int32 [2] in device memory, zeroed once by the caller; the CTAs draw units from
 * it and the last one re-arms it, so the same buffer serves every later launch (CUDA-graph
 * replays included). Launches that can run concurrently need separate buffers.
 * Every row of the K / V source buffers must hold finite values. src_b may be NULL.
 * NULL q / k / v pointers inside a source mean that side contributes no such segment. */
int b200_attn_rows_per_item(void);
long long b200_attn_workspace_bytes(void);  /* size of sched_state (device, zeroed once by the caller) */
int b200_attn_build_schedule(const int32_t* seq_table, int n_seq, int n_heads,
                             int32_t* work_units, int* n_units_out);
int b200_attn_varlen_bf16(const B200AttnSource* src_a, const B200AttnSource* src_b,
                          const int32_t* seq_table, const int32_t* work_units, int n_units,
                          int32_t* sched_state, int max_ctas, float softmax_scale, void* stream);

/* The same kernel with the two variants the PREPARE stage's text encoders need (SURVEY.md row f-4;
 * transformers CLIPAttention / T5Attention reached from encode_prompt,
 * pipeline_stable_diffusion_3_esymred.py:119-141): a causal mask (CLIP: key position <= query
 * position inside the sequence) and an additive relative position bias (T5):
 * logits[q, k] += rel_bias[head * rel_ld + (k - q) + rel_len - 1], fp32, pre-divided by softmax_scale,
 * valid for sequences of up to rel_len tokens (rel_ld >= 2 rel_len - 1). extra == NULL: plain. */
typedef struct B200AttnExtra {
  int32_t causal;
  int32_t rel_len;
  const float* rel_bias;
  int32_t rel_ld;
  int32_t q_mask_shift;    /* patch cache: q_mask[row >> shift] (shift >= 7) == 0 */
  const int32_t* q_mask;   /* marks segment-A query rows to skip (output rows stay untouched);     */
                           /* keys / values of skipped rows are still read. NULL = all computed.   */
  int32_t bounded_logits;  /* 1: the caller guarantees |logit * softmax_scale * log2(e)| <= 64 for  */
                           /* every pair (SD3.5: q and k are RMS-normalised per head, the model     */
                           /* checks its learned norm weights at load). The kernel then runs the   */
                           /* softmax without a reference maximum (same result, one pass less in   */
                           /* the step's dependent chain). Ignored with causal / rel_bias.          */
} B200AttnExtra;
int b200_attn_varlen_ex(const B200AttnSource* src_a, const B200AttnSource* src_b,
                        const int32_t* seq_table, const int32_t* work_units, int n_units,
                        int32_t* sched_state, int max_ctas, float softmax_scale,
                        const B200AttnExtra* extra, void* stream);

/* SDXL cross attention with a SHORT key sequence (queries: segment A rows of q_src, keys / values:
 * segment B rows of kv_src, at most b200_attn_cross_short_max_keys() = 80 per sequence: the 77 text
 * tokens; reference: modules/attention.py:59-110). Same seq_table as above (only qa_row, qa_len, kb_row,
 * kb_len are read), no work list and no scratch memory: one CTA per 128 query rows of a (sequence,
 * head), grid sized from max_q_len = the longest qa_len. A launch of this shape is ~1 GFLOP -- latency,
 * not throughput -- and takes a third of the time of the persistent kernel above, which remains the
 * fallback for longer key sequences (B200_ERR_UNSUPPORTED when max_kv_len is too large).
 * q_mask (patch cache, may be NULL): as in B200AttnExtra, one int32 per 2^q_mask_shift rows (>= 7). */
int b200_attn_cross_short_max_keys(void);
int b200_attn_cross_short_bf16(const B200AttnSource* q_src, const B200AttnSource* kv_src,
                               const int32_t* seq_table, int n_seq, int n_heads, int max_q_len,
                               int max_kv_len, float softmax_scale, const int32_t* q_mask,
                               int q_mask_shift, void* stream);

/* ------------------------------------------------------------------------------------------
 * HBM-bound kernels (coalesced 16-byte vectors, warp-shuffle reductions).
 * ---------------------------------------------------------------------------------------- */

/* y = LN(x) [*gamma + beta] [*(1 + mod[g, scale_col:]) + mod[g, shift_col:]], g = row_group[row]
 * (g = 0 when row_group is NULL); optional second output y2 with (scale2_col, shift2_col).
 * Covers diffusers LayerNorm, AdaLayerNormZero, SD35AdaLayerNormZeroX and
 * AdaLayerNormContinuous as called from sduss modules/transformer.py:185-279,317-386 and
 * modules/SD3Transformer.py:238. x, y: [T, ld] bf16; D <= 2048, D % 8 == 0; mod: [G, ldm] bf16.
 * row_mask (patch cache, may be NULL): device int32 per chunk of 2^row_mask_shift rows; rows of a
 * chunk whose entry is 0 are left untouched in y / y2. */
int b200_layernorm_mod_bf16(const void* x, int ldx, int T, int D, float eps, const void* gamma,
                            const void* beta, const void* mod, int ldm, const int32_t* row_group,
                            int shift_col, int scale_col, void* y, int ldy, int shift2_col,
                            int scale2_col, void* y2, int ldy2, const int32_t* row_mask,
                            int row_mask_shift, void* stream);

/* stats[row] = (mean, rstd = 1/sqrt(var + eps)) of x[row, :D] (fp32 pairs): the row statistics of
 * a LayerNorm whose affine part is folded into the consuming GEMM (B200EpilogueDesc.ln_stats).
 * D <= 2048, D % 8 == 0. */
int b200_row_stats_bf16(const void* x, int ldx, int T, int D, float eps, float* stats, void* stream);

/* y = x * sigmoid(x), n % 8 == 0 elements of bf16. */
int b200_silu_bf16(const void* x, void* y, long long n, void* stream);

/* Sinusoidal embedding, flip_sin_to_cos=True, freq_shift=0 (diffusers Timesteps; call sites
 * sduss modules/unet.py:314-334, SD3Transformer.py:81): out[i] = [cos(t_i f) | sin(t_i f)],
 * t: [n] fp32, out: [n, ldo] bf16. */
int b200_timestep_embedding(const float* t, int n, int dim, void* out, int ldo, void* stream);

/* SD3 pack: latents [C, h, w] bf16 (one device pointer per latent) -> packed token buffer
 * tokens[tok_off_i + ty*wt + tx, c*p*p + py*p + px] (im2col of the k=p,s=p PatchEmbed conv,
 * SD3Transformer.py:82-83). The packed buffer is, row for row, the reference's
 * split_sample_sd3 chunk stack viewed as [T*256, D] (modules/utils.py:86-122).
 * desc: int32 [n][4] = {token offset, tokens per row (w/p), token rows (h/p), 0}. */
int b200_sd3_patchify(const uint64_t* lat_ptr, const int32_t* desc, int n_latents, int max_tokens,
                      int C, int p, void* tokens, int ldt, void* stream);

/* SD3 scatter: tokens [T, p*p*C] -> latents [C, h, w] (concat_sample + einsum nhwpqc->nchpwq,
 * modules/utils.py:124-136, SD3Transformer.py:244-259). */
int b200_sd3_unpatchify(const void* tokens, int ldt, const int32_t* desc, int n_latents,
                        int max_tokens, int C, int p, const uint64_t* out_ptr, void* stream);

/* ------------------------------------------------------------------------------------------
 * Request-state kernels. Their per-request descriptors are HOST arrays that travel inside the
 * kernel parameter block (32 requests per launch, longer lists are split): a denoising step issues
 * no host->device copy and never drains the stream. The reference builds `torch.tensor(sigmas)` on
 * the host and `.to(device)`s it 2-3 times per resolution and step
 * (scheduling_euler_discrete.py:176,213,254; scheduling_flow_match_euler_discrete.py:183-189).
 * ---------------------------------------------------------------------------------------- */
enum B200DType { B200_DT_BF16 = 0, B200_DT_F16 = 1, B200_DT_F32 = 2 };

typedef struct B200LatentRef {
  const void* src;    /* DEVICE: current latent of the request, `elems` contiguous elements   */
  void* dst;          /* DEVICE: where the updated latent goes (step only; may equal src)     */
  int64_t elems;
  int64_t off_a;      /* gather: element offset of the (uncond) copy in the staging buffer;   */
                      /* step: element offset of the uncond prediction in eps (cfg only)      */
  int64_t off_b;      /* gather: offset of the CFG duplicate, or -1; step: cond prediction    */
  float sigma;        /* sigmas[_step_index]                                                  */
  float sigma_next;   /* sigmas[_step_index + 1] (step only)                                  */
} B200LatentRef;

/* Gathers the latents of all requests (dtype: B200DType of the latents) into the model's bf16
 * input staging buffer, writing each twice under CFG, and -- scale_input != 0 -- applies
 * EulerDiscreteScheduler.batch_scale_model_input, x / sqrt(sigma^2 + 1), with the reference's
 * rounding (sigma and every op in the latent dtype; scheduling_euler_discrete.py:161-184).
 * Replaces the per-resolution torch.cat / torch.cat([x] * 2) of
 * pipeline_stable_diffusion_3_esymred.py:270-291 and pipeline_stable_diffusion_xl_esymred.py:300-330. */
int b200_gather_latents(const B200LatentRef* reqs_host, int n_requests, int dtype,
                        int scale_input, void* staging_bf16, void* stream);

/* Fused CFG combine + scheduler update, one pass over the latents of all requests.
 *   mode 0: flow-match Euler (scheduling_flow_match_euler_discrete.py:159-203)
 *   mode 1: Euler, epsilon prediction; mode 2: Euler, v_prediction
 *           (scheduling_euler_discrete.py:187-274)
 * cfg != 0: eps = u + guidance * (c - u) (pipeline_stable_diffusion_xl_esymred.py:382-385) in the
 * arithmetic of eps_dtype (B200_DT_BF16 or B200_DT_F32), each op rounded like the torch tensor op.
 * The update runs in fp32 on the upcast sample in the reference's op order; the result is stored
 * in the latent's dtype (the reference stores the model-output dtype: identical when they agree). */
int b200_cfg_scheduler_step(const void* eps, int eps_dtype, const B200LatentRef* reqs_host,
                            int n_requests, int latent_dtype, float guidance, int cfg, int mode,
                            void* stream);

/* dst[0:n] = vals_host[0:n] (fp32): the per-latent timesteps of the step. */
int b200_write_f32(void* dst, const float* vals_host, int n, void* stream);

/* dst + i * dst_stride_bytes <- src_host[i][0 : bytes_each), i < n: gathers per-request
 * conditioning (cached projected text context, pooled embeddings) into the packed per-step
 * buffers. src_host: HOST array of device pointers; sizes multiples of 4 bytes. */
int b200_gather_rows(void* dst, long long dst_stride_bytes, const void* const* src_host, int n,
                     long long bytes_each, void* stream);

/* ------------------------------------------------------------------------------------------
 * SDXL UNet path, packed NHWC layout [sum_i H_i*W_i, C] bf16.
 * ---------------------------------------------------------------------------------------- */

/* 3x3 convolution, padding 1, stride 1 or 2, as an implicit GEMM on tcgen05 (TMA box loads
 * with out-of-bounds zero fill provide the padding: no halo exchange). Replaces F.conv2d on
 * haloed patches + the halo kernels: sduss modules/resnet.py:102-133,262-278,350-378 and
 * kernels/norm_silu_concat.cu:248 (MockNormSiluConcat / get_adjacency).
 * Step 1 (once per batch composition, host): encode one input tensor map per latent.
 *   in_desc_host: HOST int32 [n][4] = {input row offset, Hin, Win, 0}; maps_host: HOST buffer
 *   of n * 128 bytes, to be uploaded to a 64-byte-aligned device buffer. */
long long b200_conv3x3_maps_bytes(int n_latents);  /* size of maps_host / its device copy */
int b200_conv3x3_encode_maps(const void* x, int ldx, int Cin, const int32_t* in_desc_host,
                             int n_latents, int stride, void* maps_host);
/* Step 2 (per call): tiles_dev int32 [n_mtiles][4] = {latent, y0, x0, 0}: 16x8 (rows x cols)
 * OUTPUT pixel blocks; out_lat_dev int32 [n][4] = {output row offset, Hout, Wout, 0};
 * Wt: [Cout, 9*Cin] bf16 with K index = (ky*3 + kx)*Cin + c; M_total = rows of the output
 * buffer. Epilogue modes: B200_EPI_BIAS, B200_EPI_ROWVEC (+ time-embedding vector of the
 * request), B200_EPI_GATE_RESID (+ residual). Cin % 64 == 0, Cout % 8 == 0. */
int b200_conv3x3_bf16(const void* in_maps_dev, const void* out_maps_dev,
                      const void* resid_maps_dev, const int32_t* tiles_dev, int n_mtiles,
                      const int32_t* out_lat_dev, int Cin, int Cout, int stride, const void* Wt,
                      int M_total, int epi_mode, const B200EpilogueDesc* ep, void* stream);

/* Per-request GroupNorm (+ optional SiLU) with exact whole-latent statistics. Supersedes the
 * reference's native kernels RowwiseMoments / GetFullMeanAndRstd / NormSiluConcat
 * (kernels/norm_silu_concat.cu:41,361,87; pybind `groupnorm`, norm_silu_concat.cpp:66-86)
 * without their approximated variance (D1), in-place race and per-call device syncs.
 * Every latent's pixel count must be a multiple of 64. lat_chunks: int32 [n][4] = {first
 * 64-row chunk, number of chunks, 0, 0}; workspace: b200_groupnorm_workspace_bytes() bytes, ZEROED once
 * by the caller before its first use and not written by anything else (it carries the epoch of the
 * grid barrier of the single-launch kernel across calls); one workspace per stream. */
long long b200_groupnorm_workspace_bytes(long long total_rows, int n_latents);
int b200_groupnorm_nhwc_bf16(const void* x, int ldx, long long T, int C, int groups, float eps,
                             const void* gamma, const void* beta, const int32_t* row_group,
                             const int32_t* lat_chunks, int n_latents, int silu, void* y, int ldy,
                             void* workspace, void* stream);

/* The same GroupNorm when the tensor was written by b200_conv3x3_bf16 with ep->stats_out set: the
 * per-tile partial sums replace the statistics pass, so x is read once (finalize + apply).
 * lat_tiles: int32 [n][4] = {first tile of the latent in the conv's tile list, its tiles, its
 * pixels, 0}. Same workspace as b200_groupnorm_nhwc_bf16. */
int b200_groupnorm_from_conv_stats(const void* x, int ldx, long long T, int C, int groups, float eps,
                                   const void* gamma, const void* beta, const int32_t* row_group,
                                   const float* conv_stats, const int32_t* lat_tiles, int n_latents,
                                   int silu, void* y, int ldy, void* workspace, void* stream);

/* SDXL pack: NCHW latents -> im2col rows of the 3x3/pad-1 conv_in (K = 9*C padded to ldo with
 * zeros): out[row_i + y*W + x, c*9 + ky*3 + kx]. Replaces split_sample's haloed windows
 * (modules/unet.py:104-184). desc int32 [n][4] = {row offset, H, W, 0}. */
int b200_pack_im2col3x3(const uint64_t* lat_ptr, const int32_t* desc, int n_latents,
                        int max_pixels, int C, void* out, int ldo, void* stream);
/* SDXL scatter: NHWC rows (first C columns of x) -> NCHW latents (concat_sample,
 * modules/unet.py:187-202). */
int b200_scatter_nchw(const void* x, int ldx, const int32_t* desc, int n_latents, int max_pixels,
                      int C, const uint64_t* out_ptr, void* stream);
/* Nearest 2x upsample on the packed layout (modules/resnet.py:316). */
int b200_upsample2x_nhwc(const void* x, int ldx, const int32_t* in_desc, const int32_t* out_desc,
                         int n_latents, int max_out_pixels, int C, void* y, int ldy, void* stream);
/* dst[:, :cols] = src[:, :cols] with independent row strides (skip-connection concat). */
int b200_copy_cols_bf16(const void* src, int lds, void* dst, int ldd, long long T, int cols,
                        void* stream);

/* The reference's own patch format, index-exact (interoperability / parity fixtures):
 * split_sample (modules/unet.py:104-184): NCHW latents -> [P, C, ps+2, ps+2] haloed windows;
 * concat_sample (modules/unet.py:187-202): [P, C, ps, ps] -> NCHW latents.
 * ldesc int32 [n][4] = {0, H, W, 0}; pdesc int32 [P][4] = {latent, patch row, patch col, 0}. */
int b200_split_patches(const uint64_t* lat_ptr, const int32_t* ldesc, const int32_t* pdesc,
                       int n_patches, int C, int ps, void* out, void* stream);
int b200_concat_patches(const void* patches, const int32_t* ldesc, const int32_t* pdesc,
                        int n_patches, int C, int ps, const uint64_t* out_ptr, void* stream);

/* ---- Patch cache ("block skip"; CacheManager.get_sd3_mask,
 * sduss/model_executor/modules/cache_manager.py:161-191). One launch per transformer block decides,
 * on the device and inside the captured graph, which 256-token patches are recomputed: it streams the
 * block input x once -- per-patch MSE against `prev` (the block input at the previous step), which
 * it refreshes in the same pass -- evaluates the flattened RandomForest on [block_index, timestep,
 * MSE] and applies the refresh rule. mask[p] = 1: recompute. The mask is what the row_mask /
 * q_mask arguments of the GEMM, LayerNorm and attention entry points take.
 * A patch of a latent with latent_valid == 0 (its kept copies belong to another request or an older
 * step) has MSE = float(sys.maxsize) as in the reference and is always recomputed. */
typedef struct B200Forest {      /* all arrays in DEVICE memory                                   */
  const int32_t* feature;        /* [nodes] feature index (0 block, 1 timestep, 2 MSE); < 0 = leaf */
  const float* threshold;        /* [nodes] go left when x[feature] <= threshold (sklearn)         */
  const int32_t* left;           /* [nodes]                                                        */
  const int32_t* right;          /* [nodes]                                                        */
  const float* value;            /* [nodes] leaf: probability of class 1 (= recompute)             */
  const int32_t* roots;          /* [n_trees] root node of each tree                               */
  int32_t n_trees;
} B200Forest;
long long b200_patch_mask_workspace_bytes(int n_patches);  /* zero it once; the kernel leaves it zeroed */
int b200_patch_mask_bf16(const void* x, int ldx, void* prev, int ldp, int n_patches,
                         int rows_per_patch, int D, const int32_t* patch_latent,
                         const float* latent_t, const float* latent_valid, int32_t* skipped,
                         int32_t* mask, float* mse, const B200Forest* forest, int block_index,
                         int refresh, void* workspace, void* stream);
/* The SDXL UNet variant (CacheManager.get_mask, cache_manager.py:101-159; called per down / mid / up
 * block, modules/unet_2d_blocks.py:40,102,180,250,345). Down and mid blocks use the three features
 * above (refresh = 4, :147). An up block's feature row continues with the MSE of every skip tensor it
 * consumes (is_upsample, :106-121): launch this once per skip tensor with forest == NULL (MSE-only:
 * `mse` [n_patches] is written, `prev` refreshed, mask / skipped may be NULL) and then once on the block
 * input with extra_mse = those results, [n_extra][n_patches], n_extra <= 3 (features 3 .. 2 + n_extra). */
int b200_patch_mask_ex(const void* x, int ldx, void* prev, int ldp, int n_patches, int rows_per_patch,
                       int D, const int32_t* patch_latent, const float* latent_t,
                       const float* latent_valid, int32_t* skipped, int32_t* mask, float* mse,
                       const B200Forest* forest, int block_index, int refresh, const float* extra_mse,
                       int n_extra, void* workspace, void* stream);

/* ---- Prepare stage (text encoders behind encode_prompt,
 * pipeline_stable_diffusion_3_esymred.py:119-141, pipeline_stable_diffusion_xl_esymred.py:120-160).
 * Linear layers, CLIP's LayerNorm and the attention run on the kernels above; these two are added. */
/* out[i, :] = table[ids[i], :] (+ pos[i % seq_len, :] when pos != NULL): CLIPTextEmbeddings
 * (token + learned position embedding) / T5 embed_tokens. ids: DEVICE int32 [n]. D % 8 == 0. */
int b200_embed_rows_bf16(const int32_t* ids, int n, const void* table, int vocab, int D,
                         const void* pos, int seq_len, void* out, int ldo, void* stream);
/* y = weight * bf16(x * rsqrt(mean(x^2) + eps)): T5LayerNorm (no mean subtraction, no bias). */
int b200_rmsnorm_bf16(const void* x, int ldx, int T, int D, float eps, const void* weight, void* y,
                      int ldy, void* stream);

/* ---- VAE decode stage (post_inference, pipeline_stable_diffusion_xl_esymred.py:
 * 406-462 and pipeline_stable_diffusion_3_esymred.py:391-415). The decoder's convolutions, GroupNorms,
 * upsamples and linears run on the kernels above; these two are what the stage adds. */
/* out_l[co, p] = bias[co] + sum_ci weight[co, ci] * in_l[ci, p] on NCHW bf16 latents (fp32 math):
 * `latents / scaling_factor (+ shift_factor)` and the 1x1 post_quant_conv folded into one per-pixel
 * affine map. weight fp32 [c_out, c_in], bias fp32 [c_out], c_in, c_out <= 16;
 * desc int32 [n][4] = {row offset, H, W, 0}; in_ptr / out_ptr: device arrays of latent pointers. */
int b200_latent_affine(const uint64_t* in_ptr, const uint64_t* out_ptr, const int32_t* desc,
                       int n_latents, int max_pixels, int c_in, int c_out, const float* weight,
                       const float* bias, void* stream);
/* p[r, :] = softmax(scale * s[r, :]): fp32 logits (GEMM output) -> bf16 probabilities, the softmax of
 * the decoder's single-head mid-block attention (diffusers Attention, head_dim = channels = 512).
 * cols % 4 == 0, cols <= 16384; lds / ldp in elements. */
int b200_softmax_rows(const float* s, long long lds, int rows, int cols, float scale, void* p,
                      long long ldp, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SDUSS_B200_H_ */

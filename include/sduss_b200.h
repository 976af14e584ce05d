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
 * work_items: int32 [n_items][4] = {seq, q_segment (0 = A, 1 = B), row offset of the 128-row
 *            query tile inside that segment, 0}; one CTA per (item, head).
 * Every row of the K / V source buffers must hold finite values. src_b may be NULL.
 * NULL q / k / v pointers inside a source mean that side contributes no such segment. */
int b200_attn_varlen_bf16(const B200AttnSource* src_a, const B200AttnSource* src_b,
                          const int32_t* seq_table, const int32_t* work_items, int n_items,
                          int n_heads, float softmax_scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SDUSS_B200_H_ */

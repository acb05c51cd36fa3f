/*
 * mebt_b200 C ABI — hand-written sm_100a kernels for the MeBT latent-bottleneck transformer hot path.
 *
 * The reference (Ugness/MeBT) is pure PyTorch: it has no FFI layer, so each entry point below names the
 * reference call site (file:line under /root/reference) whose ATen/cuBLAS library calls it replaces.
 * INTEGRATION.md shows the ctypes binding a maintainer would add on the reference side.
 *
 * Conventions
 *   - every function returns 0 on success, a non-zero MEBT_ERR_* code otherwise; mebt_last_error()
 *     returns a thread-local message.  No C++ exception crosses this boundary.
 *   - all pointers are DEVICE pointers into caller-owned (torch-owned) storage unless the name says
 *     `host`; inputs are const, outputs pre-allocated by the caller.
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous and never synchronise.
 *   - library-owned device memory (allocated on first use, per device, never on the data path's steady state):
 *     the split-K scratch of mebt_gemm_bf16 (96 MiB + tile counters per stream, at most four streams; only shapes
 *     with few tiles and K >= 2048 use it), 4 KiB of reduction tickets per (device, stream) for the column-sum /
 *     LayerNorm-parameter reductions (zeroed on that stream), a 4-byte error flag, and the three side streams +
 *     events of the training engine.  Everything else is the caller's.
 *   - bf16 = __nv_bfloat16 bits, row-major; `ld*` are row strides in ELEMENTS.
 *   - there is no CPU fallback: without an sm_100 device every compute entry point fails.
 */
#ifndef MEBT_B200_H_
#define MEBT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
  MEBT_OK = 0,
  MEBT_ERR_SHAPE = 1,
  MEBT_ERR_DTYPE = 2,
  MEBT_ERR_WORKSPACE = 3,
  MEBT_ERR_CUDA = 4,
  MEBT_ERR_UNSUPPORTED = 5,
  MEBT_ERR_DEVICE = 6
};

/* ---- runtime ------------------------------------------------------------------------------- */
const char* mebt_version(void);
const char* mebt_last_error(void);
/* 0 iff the current CUDA device is sm_100-class (B200). */
int mebt_device_check(void);
/* Number of mebt_b200 kernels launched by this process so far (bench.py's `gpu_launches`). */
unsigned long long mebt_launch_count(void);
/* Per-kernel-family timing with CUDA events recorded on each launch's own stream.  Families (index into the
 * report arrays, length 10): 0 gemm (work = flops), 1 attention (flops), 2 layernorm, 3 embed, 4 sample, 5 ce,
 * 6 remask, 7 scatter, 8 vq, 9 other (work = algorithmic bytes).  mebt_profile_report synchronises the device,
 * fills the sums since the last report and clears them.  Profiling serialises nothing but adds two event records
 * per launch: keep it off inside timed throughput regions. */
void mebt_profile_enable(int on);
int mebt_profile_report(double* time_ms, double* work, long long* launches);

/* ---- K2 / K4 : dense contractions ---------------------------------------------------------- */
enum {
  MEBT_GEMM_GELU = 1,         /* exact erf GELU after bias (gpt.py:152 nn.GELU) */
  MEBT_GEMM_OUT_FP32 = 2,     /* C is float (logits, weight gradients); default bf16 */
  MEBT_GEMM_ACCUMULATE = 4,   /* C += result (fp32 C only; gradient accumulation) */
  MEBT_GEMM_DGELU = 8,        /* result *= gelu'(aux): the backward of the fused fc1+GELU epilogue */
  MEBT_GEMM_FORCE_BN256 = 16, /* tile-width overrides, for tests and tuning */
  MEBT_GEMM_FORCE_BN128 = 32,
  MEBT_GEMM_FORCE_BN64 = 64,
  MEBT_GEMM_NO_SPLITK = 128,  /* never split the reduction across CTAs */
  MEBT_GEMM_NO_PAIR = 256,    /* never use the 2-CTA cluster variant (B tile shared by TMA multicast) */
  MEBT_GEMM_FORCE_PAIR = 512,
  MEBT_GEMM_DUAL = 1024       /* 128-wide tiles: two MMA-issuing threads with an accumulator half each (opt-in) */
};
/*
 * C[M,N] = act( A * B^T + bias ) + residual, tcgen05/TMEM/TMA.
 *   a_mn_major = 0: A is [M,K] row-major (lda >= K).   1: A is stored [K,M] row-major (lda >= M).
 *   b_mn_major = 0: B is [N,K] row-major — a torch nn.Linear weight.   1: B is stored [K,N].
 * Replaces: nn.Linear forward at mebt/modules/gpt.py:126-128 (query/key/value), :140 (proj),
 * :150-155 (mlp fc1+GELU, fc2), :248 (head, no bias); with MN-major operands the dgrad
 * (dX = dY * W) and wgrad (dW = dY^T * X) GEMMs autograd runs for the same modules.
 * bias: fp32 [N] or NULL.  residual: bf16 [M, ldres] or NULL (gpt.py:184-185 residual adds).
 */
int mebt_gemm_bf16(const void* A, int lda, int a_mn_major, const void* B, int ldb, int b_mn_major, void* C, int ldc,
                   int M, int N, int K, const float* bias, const void* residual, int ldres, int flags, void* stream);
/* Same with an auxiliary bf16 [M, ldaux] tensor: with MEBT_GEMM_GELU it RECEIVES the pre-activation (saved for
 * backward); with MEBT_GEMM_DGELU it SUPPLIES the pre-activation and the result is multiplied by gelu'(aux), i.e.
 * d(pre-activation) = (dY * W) .* gelu'(a) for the MLP of mebt/modules/gpt.py:150-155. */
int mebt_gemm_bf16_aux(const void* A, int lda, int a_mn_major, const void* B, int ldb, int b_mn_major, void* C, int ldc,
                       int M, int N, int K, const float* bias, const void* residual, int ldres, void* aux, int ldaux,
                       int flags, void* stream);

/* Up to 6 independent weight-gradient problems  dW[n_out, k_in] (+)= dY[rows, n_out]^T X[rows, k_in]  (bf16 operands,
 * fp32 dW) in ONE persistent launch: the nn.Linear weight gradients autograd computes for one Block
 * (mebt/modules/gpt.py:126-128,140,150-155 in the backward of :159-195).  Each alone is a fraction of a wave of tiles at
 * the training shapes; together they fill the machine.  The outputs must not overlap (the problems run concurrently).
 * Shapes the grouped tile does not fit (k_in % 256 != 0) are run as separate mebt_gemm_bf16 launches. */
typedef struct mebt_wgrad_desc {
  const void* dY; int ld_dy;      /* bf16 [rows, n_out], row stride in elements */
  const void* X;  int ldx;        /* bf16 [rows, k_in] */
  float* dW;      int ldw;        /* fp32 [n_out, k_in] */
  int n_out, k_in, rows;
  int accumulate;                 /* dW += result instead of dW = result */
} mebt_wgrad_desc_t;
int mebt_gemm_grouped_wgrad(const mebt_wgrad_desc_t* problems, int n_problems, void* stream);

/* ---- dtypes --------------------------------------------------------------------------------- */
enum { MEBT_DTYPE_BF16 = 0, MEBT_DTYPE_FP32 = 1 };

/* ---- K1 : embedding stem -------------------------------------------------------------------- */
/*
 * contexts[b,i,:] = tok_emb[x[b, ctx_idx[b,i]]] + pos_emb[ctx_idx[b,i]]
 * targets [b,i,:] = mask_emb + pos_emb[tgt_idx[b,i]]        latents[b,l,:] = sos_emb[l]
 * Replaces the gather / nn.Embedding / pos_emb.repeat + gather / mask_emb.repeat / sos_emb.repeat chain at
 * mebt/transformer.py:298-317 (reconstruct_mask) and :255-277 (forward).
 * x_indices/ctx_idx/tgt_idx: int64, row strides in elements (index tensors may be views of a permutation).
 * tok_emb [V,D], pos_emb [n_pos,D], mask_emb [D], sos_emb [L,D]: fp32 parameters.  Outputs: [B*NC,D], [B*NT,D],
 * [B*L,D] in out_dtype.  Out-of-range ids/positions are skipped and flagged (mebt_check_index_errors).
 */
int mebt_embed_gather(const int64_t* x_indices, int x_stride, const int64_t* ctx_idx, int ctx_stride,
                      const int64_t* tgt_idx, int tgt_stride, const float* tok_emb, const float* pos_emb,
                      const float* mask_emb, const float* sos_emb, void* contexts, void* targets, void* latents, int B,
                      int NC, int NT, int L, int D, int V, int n_pos, int out_dtype, void* stream);

/* nn.LayerNorm(D), eps inside the sqrt (mebt/modules/gpt.py:147-148 ln1/ln2, :216 ln_f).  gamma/beta fp32.
 * mean_out / rstd_out: optional fp32 [rows] (saved for backward). */
int mebt_layernorm(const void* x, int ldx, int in_dtype, const float* gamma, const float* beta, void* y, int ldy,
                   int out_dtype, int rows, int D, float eps, float* mean_out, float* rstd_out, void* stream);

/* ---- K8 : write sampled ids back -------------------------------------------------------------- */
/* x[b, tgt_idx[b,i]] = ids[b,i]; replaces the two sparse_coo_tensor(...).to_dense() + where at
 * mebt/transformer.py:413-439, :571-585, :615-629. */
int mebt_scatter_ids(int64_t* x, int x_stride, const int64_t* tgt_idx, int tgt_stride, const int64_t* ids, int B, int NT,
                     int N, void* stream);

/* ---- K10 : codebook row gather ---------------------------------------------------------------- */
/* out = E[enc]; channel_first=0 -> [batch*S, C] (Codebook.dictionary_lookup, modules/codebook.py:99-101);
 * channel_first=1 -> [batch, C, S], fusing the shift_dim of mebt/vqgan.py:91-92 / codebook.py:61-62. */
int mebt_row_gather(const int64_t* enc, const float* E, float* out, int batch, int S, int C, int K, int channel_first,
                    void* stream);

/* fp32 -> bf16 (master weights -> tensor-core operands). n % 4 == 0. */
int mebt_cast_f32_to_bf16(const float* in, void* out, long long n, void* stream);

/* Reads and clears the device-side index-error flag. SYNCHRONISES the stream: tests / debug only. */
int mebt_check_index_errors(void* stream);

/* ---- K5 : masked cross-entropy ---------------------------------------------------------------- */
/*
 * Per row r of logits [rows, V] (row stride ld): row_loss[r] = (1-eps) * (lse - x_t) + eps * (lse - mean x),
 * row_rank[r] = #{v : x_v > x_t}  (top-1 hit <=> 0, top-5 hit <=> < 5), and optionally
 * dlogits = grad_scale * (softmax - (1-eps) onehot - eps/V) in the logits dtype (may alias logits).
 * Replaces F.cross_entropy(reduction='sum', label_smoothing) at mebt/transformer.py:726 and the topk(5) of
 * accuracy() (mebt/utils.py:80-94).  V % 4 == 0, V <= 16384.
 */
int mebt_masked_ce(const void* logits, long long ld, int dtype, const int64_t* targets, int rows, int V,
                   float label_smoothing, float* row_loss, int* row_rank, void* dlogits, long long ld_d,
                   float grad_scale, void* stream);
/* out3 = {sum row_loss, #rank==0, #rank<5}; fixed-order (deterministic) reduction. row_rank may be NULL. */
int mebt_ce_reduce(const float* row_loss, const int* row_rank, int rows, float* out3, void* stream);

/* ---- K6 : sampling from logits ---------------------------------------------------------------- */
/*
 * ids[r] = argmax_v (p_v / sum p) / q_v with p = softmax(top_k_filter(logits / (temperature + 1e-8))),
 * scores[r] = p[ids[r]].  q: Exp(1) noise [rows, V] supplied by the caller (parity mode, = the reference's
 * exponential_ draw).  When noise == NULL (fast mode) the id is drawn from the same categorical distribution p by
 * inverse-CDF with one in-kernel Philox4x32-10(seed, offset, row) uniform per row, so no noise tensor exists.
 * probs (optional, fp32 [rows, V]) receives the softmax the reference returns with return_probs=True.
 * Replaces sample_from_logits + gumbel_sort + top_k_logits, mebt/transformer.py:843-895 (a full 16384-way sort
 * per row and ~10 passes over [B,NT,V] become one pass).  top_k <= 0 and top_p outside (0,1) disable the filters.
 * top_p (nucleus, transformer.py:898-910) keeps, in descending order, every token up to and including the one whose
 * cumulative mass first reaches top_p, then renormalises; tokens tying with the boundary value are all kept.
 */
int mebt_sample_logits(const void* logits, long long ld, int dtype, int rows, int V, float temperature, int top_k,
                       float top_p, const float* noise, unsigned long long seed, unsigned long long offset,
                       int64_t* ids, float* scores, float* probs, void* stream);

/* ---- K7 : confidence re-masking ---------------------------------------------------------------- */
/*
 * order = argsort_desc((score / sum score) / q^ctemp); next_ctx = cat[ctx, tgt[order[:n_new]]],
 * next_tgt = tgt[order[n_new:]].  Replaces MaskGen.gumbel_top_k + the gathers of generate_next_mask
 * (mebt/mask_sampler.py:178-187, :226-234).  noise q [B,NT] Exp(1) or NULL (Philox).  Ties: lower index first.
 * Outputs are optional (NULL to skip); order_out int64 [B,NT].  NT <= 16384.
 */
int mebt_remask_sort(const float* score, const float* noise, float ctemp, const int64_t* ctx_idx, int ctx_stride,
                     const int64_t* tgt_idx, int tgt_stride, int B, int NC, int NT, int n_new, unsigned long long seed,
                     unsigned long long offset, int64_t* next_ctx, int64_t* next_tgt, int64_t* order_out, void* stream);

/* ---- K9 : codebook nearest neighbour ---------------------------------------------------------- */
/* out[k] = sum_c E[k,c]^2 (the |E|^2 term of modules/codebook.py:55; constant while the VQGAN is frozen). */
int mebt_row_sqnorm(const float* E, int K, int C, float* out, void* stream);
size_t mebt_vq_argmin_workspace_bytes(long long M);
/*
 * out_idx[b*S + s] = argmin_k |z[b,:,s] - E[k]|^2, fp32-accurate, lowest index on exact ties.
 * z is channel-first [batch, C, S] exactly as VQGAN.encode hands it to the codebook; replaces
 * shift_dim/flatten + distance matrix + argmin at mebt/modules/codebook.py:52-57.
 */
int mebt_vq_argmin(const float* z_channel_first, int batch, int C, int S, const float* E, const float* e_sqnorm, int K,
                   int64_t* out_idx, void* workspace, size_t workspace_bytes, void* stream);
/*
 * The same search on the tensor cores (the default of Codebook.forward): z.E^T as ONE fp16 tcgen05 GEMM over operands split
 * into fp16 (hi, lo) halves, reduction dimension 3C ([z_hi|z_lo|z_hi].[e_hi|e_hi|e_lo]^T, fp32 accumulation, relative
 * 2^-21 per product), with the distance (|z|^2 - 2 z.e) + |e|^2 and the running argmin as the GEMM epilogue; lowest index
 * on exact ties.  e_split: the codebook in that form, fp16 [K, 3C], made once per codebook by mebt_vq_split_codebook
 * (mebt_vq_codebook_split_bytes bytes).  C and K multiples of 64.
 */
size_t mebt_vq_codebook_split_bytes(int K, int C);
int mebt_vq_split_codebook(const float* E, int K, int C, void* e_split, void* stream);
size_t mebt_vq_argmin_tc_workspace_bytes(long long M, int C);
int mebt_vq_argmin_tc(const float* z_channel_first, int batch, int C, int S, const void* e_split, const float* e_sqnorm,
                      int K, int64_t* out_idx, void* workspace, size_t workspace_bytes, void* stream);

/* ---- K3 : latent attention --------------------------------------------------------------------- */
/*
 * O[b,q,h,:] = softmax_k(Q[b,q,h,:].K[b,k,h,:] / sqrt(head_dim)) V[b,k,h,:], bf16 in/out, fp32 softmax, flash-style
 * (no [B,h,NQ,NK] matrix), tcgen05 QK^T and PV.  Replaces the bmm/softmax/bmm of CrossAttention.forward,
 * mebt/modules/gpt.py:131-137.  Keys/values come from up to two sources that are walked back to back, which is
 * how lt2l's torch.cat([sos_emb, targets]) (gpt.py:175) is consumed without materialising it:
 *   Q  : rows b*NQ+q  of a [B*NQ,  ldq] buffer, head h at columns q_col0 + 64h
 *   KV1: rows b*NK1+k of a [B*NK1, ld1] buffer, K at k1_col0 + 64h, V at v1_col0 + 64h   (NK1 may be 0)
 *   KV2: same with NK2 (0 = absent).            O: [B*NQ, ldo], head h at columns 64h.
 * NK1 + NK2 == 0 gives O = 0 (the reference's empty softmax).  lse: optional fp32 [B,H,NQ] (for backward).
 */
int mebt_latent_attention_fwd(const void* Q, int ldq, int q_col0, const void* KV1, int ld1, int k1_col0, int v1_col0,
                              int NK1, const void* KV2, int ld2, int k2_col0, int v2_col0, int NK2, void* O, int ldo,
                              float* lse, int B, int H, int NQ, int head_dim, void* stream);
/* The same with caller-provided scratch (mebt_latent_attention_fwd_workspace_bytes): launches with few work items and
 * long key lists (small-batch 128-frame sampling: B*H*ceil(NQ/256) <= 74, >= 1024 keys) split each item's keys over
 * up to 8 CTAs and merge the partial rows in a second kernel; otherwise identical. */
size_t mebt_latent_attention_fwd_workspace_bytes(int B, int H, int NQ);
int mebt_latent_attention_fwd_ws(const void* Q, int ldq, int q_col0, const void* KV1, int ld1, int k1_col0, int v1_col0,
                                 int NK1, const void* KV2, int ld2, int k2_col0, int v2_col0, int NK2, void* O, int ldo,
                                 float* lse, int B, int H, int NQ, int head_dim, void* workspace, size_t workspace_bytes,
                                 void* stream);

/* ---- backward of the memory-bound ops (training step, config #2) -------------------------------- */
/* out[n] (+)= sum_r X[r,n], X bf16 [rows, ld]: the bias gradients autograd computes for every nn.Linear of
 * mebt/modules/gpt.py:102-116,150-155.  Two-stage fixed-order reduction (deterministic). */
size_t mebt_colsum_workspace_bytes(int N);
int mebt_colsum(const void* X, int ld, int rows, int N, float* out, int accumulate, void* workspace,
                size_t workspace_bytes, void* stream);
/* nn.LayerNorm backward (gpt.py:147-148,216).  dy, x, dx: bf16 [rows, D] contiguous; mean/rstd as saved by
 * mebt_layernorm.  dx (+)= ...; dgamma/dbeta (+)= ... (fp32 [D]). */
size_t mebt_layernorm_bwd_workspace_bytes(int D);
int mebt_layernorm_bwd(const void* dy, const void* x, const float* mean, const float* rstd, const float* gamma, void* dx,
                       int accumulate_dx, float* dgamma, float* dbeta, int accumulate_params, int rows, int D,
                       void* workspace, size_t workspace_bytes, void* stream);
/* Backward of mebt_embed_gather: scatter-adds bf16 stream gradients into the fp32 gradients of tok_emb [V,D],
 * pos_emb [n_pos,D], mask_emb [D], sos_emb [L,D] (all ACCUMULATED into).  Workspace: mebt_colsum_workspace_bytes(D). */
int mebt_embed_backward(const int64_t* x_indices, int x_stride, const int64_t* ctx_idx, int ctx_stride,
                        const int64_t* tgt_idx, int tgt_stride, const void* d_contexts, const void* d_targets,
                        const void* d_latents, float* d_tok_emb, float* d_pos_emb, float* d_mask_emb, float* d_sos_emb,
                        int B, int NC, int NT, int L, int D, void* workspace, size_t workspace_bytes, void* stream);

/* Backward of mebt_latent_attention_fwd (what autograd runs for mebt/modules/gpt.py:131-137).  Q/KV1/KV2/O and
 * their layouts exactly as in the forward call; dO [B*NQ, lddo] (head h at columns 64h); lse from the forward.
 * Gradients are written with the forward's geometry: dQ into a [B*NQ, lddq] buffer at column dq_col0 + 64h, dK/dV of
 * source s into a [B*NKs, ldds] buffer at dks_col0 / dvs_col0 + 64h (buffers may alias, e.g. one [rows,3D] dQKV).
 * Workspace: mebt_latent_attention_bwd_workspace_bytes (the per-row delta = rowsum(dO .* O)). */
size_t mebt_latent_attention_bwd_workspace_bytes(int B, int H, int NQ);
int mebt_latent_attention_bwd(const void* Q, int ldq, int q_col0, const void* KV1, int ld1, int k1_col0, int v1_col0,
                              int NK1, const void* KV2, int ld2, int k2_col0, int v2_col0, int NK2, const void* O,
                              int ldo, const void* dO, int lddo, const float* lse, void* dQ, int lddq, int dq_col0,
                              void* dKV1, int ldd1, int dk1_col0, int dv1_col0, void* dKV2, int ldd2, int dk2_col0,
                              int dv2_col0, int B, int H, int NQ, int head_dim, void* workspace, size_t workspace_bytes,
                              void* stream);

/* ---- fp32-accurate mode (logits within 1e-4 of the fp32 reference; configs[0]) ---------------------------- */
/* x (fp32 [rows, K], row stride ld) -> bf16 [rows, 3K] = hi | lo | hi (weight_side = 0) or hi | hi | lo (weight_side = 1),
 * hi = bf16(x), lo = bf16(x - hi).  mebt_gemm_bf16 on the two splits (K' = 3K, fp32 output) then evaluates the
 * nn.Linear of mebt/modules/gpt.py:126-128,140,150-155,248 with a relative error of 2^-16 per product. */
int mebt_split_f32_bf16x3(const float* x, int ld, int rows, int K, void* out, int weight_side, void* stream);
/* mebt_latent_attention_fwd on fp32 buffers in plain fp32 arithmetic (same geometry; no LSE, no dropout). */
int mebt_latent_attention_fwd_f32(const float* Q, int ldq, int q_col0, const float* KV1, int ld1, int k1_col0, int v1_col0,
                                  int NK1, const float* KV2, int ld2, int k2_col0, int v2_col0, int NK2, float* O, int ldo,
                                  int B, int H, int NQ, int head_dim, void* stream);

/* ---- optimizer step (mebt/transformer.py:749-798 configure_optimizers -> torch.optim.AdamW, betas (0.9, 0.95)) ---- */
/* One AdamW step over flat fp32 buffers p / g / m / v [n] with torch's fused-AdamW arithmetic, plus p_bf16 = bf16(p) (the
 * tensor-core operand copy).  decay_blocks[i >> block_shift] != 0 marks elements of weight-decayed tensors; `step` counts
 * from 1 (bias correction). */
int mebt_adamw_flat(float* p, const float* g, float* m, float* v, void* p_bf16, const unsigned char* decay_blocks,
                    int block_shift, long long n, float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                    void* stream);
/* The same update as a background kernel of at most `max_ctas` CTAs (0 = the full grid): it trickles through HBM at a
 * fraction of the bandwidth, for a caller that overlaps the update of finished parameter ranges with the rest of the
 * backward pass on another stream (the latency-bound kernels there keep their SMs and their L2). */
int mebt_adamw_flat_bg(float* p, const float* g, float* m, float* v, void* p_bf16, const unsigned char* decay_blocks,
                       int block_shift, long long n, float lr, float beta1, float beta2, float eps, float weight_decay,
                       int step, int max_ctas, void* stream);

/* ---- dropout (training mode; nn.Dropout at mebt/modules/gpt.py:112-113,140,150-155,216,239-242) ------------- */
/* Keep decisions are counter-based: a pure function of (seed, site, row, column), regenerated by the backward
 * kernels instead of being stored.  p is quantised to 1/65536 and kept elements are scaled by 65536/(65536-round(p*65536)).
 * y[r,:] = resid[r,:] + x[r,:] .* keep / (1-p)   (bf16 [rows, D]; resid may be NULL; y may alias x).
 * Calling it on dy with the same (p, seed, site) and resid = NULL is the backward. */
int mebt_dropout_rows(const void* x, int ldx, const void* resid, int ldres, void* y, int ldy, int rows, int D, float p,
                      unsigned long long seed, unsigned long long site, void* stream);
/* mebt_latent_attention_fwd / _bwd with attn_drop applied to the softmax output (gpt.py:136): O = (P .* keep/(1-p)) V. */
int mebt_latent_attention_fwd_dropout(const void* Q, int ldq, int q_col0, const void* KV1, int ld1, int k1_col0,
                                      int v1_col0, int NK1, const void* KV2, int ld2, int k2_col0, int v2_col0, int NK2,
                                      void* O, int ldo, float* lse, int B, int H, int NQ, int head_dim, float p,
                                      unsigned long long seed, void* stream);
int mebt_latent_attention_bwd_dropout(const void* Q, int ldq, int q_col0, const void* KV1, int ld1, int k1_col0,
                                      int v1_col0, int NK1, const void* KV2, int ld2, int k2_col0, int v2_col0, int NK2,
                                      const void* O, int ldo, const void* dO, int lddo, const float* lse, void* dQ,
                                      int lddq, int dq_col0, void* dKV1, int ldd1, int dk1_col0, int dv1_col0, void* dKV2,
                                      int ldd2, int dk2_col0, int dv2_col0, int B, int H, int NQ, int head_dim, float p,
                                      unsigned long long seed, void* workspace, size_t workspace_bytes, void* stream);
/* The keep factors (0 or 1/(1-p)) those two kernels apply, as fp32 [B, H, NQ, NK1+NK2] (test support). */
int mebt_attention_dropout_mask(float* out, int B, int H, int NQ, int NK1, int NK2, float p, unsigned long long seed,
                                void* stream);

/* ---- the layer stack in one call ---------------------------------------------------------------- */
enum {
  MEBT_MODE_LATENT_ENC = 0,  /* q = latents, kv = contexts            -> latents   (gpt.py:167-169) */
  MEBT_MODE_LATENT_SELF = 1, /* q = kv = latents                      -> latents   (gpt.py:164-166) */
  MEBT_MODE_LATENT_DEC = 2,  /* q = targets, kv = latents             -> targets   (gpt.py:170-172) */
  MEBT_MODE_LT2L = 3,        /* q = latents, kv = cat[latents,targets]-> latents   (gpt.py:173-175) */
  MEBT_MODE_MASKGIT = 4      /* q = kv = cat[contexts,targets]        -> both      (gpt.py:176-178) */
};
/* One Block's parameters. Device pointers; *_w matrices are bf16 [out,in] row-major (nn.Linear layout),
 * w_qkv = rows (query | key | value) stacked to [3D, D]; LayerNorm and bias vectors are fp32. */
typedef struct mebt_layer {
  int mode;
  const float* ln1_w; const float* ln1_b; const float* ln2_w; const float* ln2_b;
  const void* w_qkv;  const float* b_qkv;
  const void* w_proj; const float* b_proj;
  const void* w_fc1;  const float* b_fc1;
  const void* w_fc2;  const float* b_fc2;
} mebt_layer_t;

/* Optional inference-time hoist of the latent_enc K|V projections (contexts are constant through the stack and
 * ln1's statistics are block-independent): w_enc_kv bf16 [n_enc*2D, D] = per-block (key|value) weights with ln1's
 * gamma folded in, b_enc_kv fp32 [n_enc*2D] = bias + W.beta, ones/zeros fp32 [D].  Blocks are matched in order of
 * appearance of MEBT_MODE_LATENT_ENC. */
typedef struct mebt_enc_hoist {
  int n_enc;
  const void* w_enc_kv;
  const float* b_enc_kv;
  const float* ones;
  const float* zeros;
} mebt_enc_hoist_t;

size_t mebt_stack_forward_workspace_bytes(int B, int L, int NC, int NT, int D);
size_t mebt_stack_forward_hoisted_workspace_bytes(int B, int L, int NC, int NT, int D, int n_enc);
/*
 * GPT.forward (mebt/modules/gpt.py:234-253) in eval mode: n_layers Blocks threaded over the three streams, then
 * logits = head(ln_f(targets)).  `layers` is a HOST array.  lat [B*L,D], ctx [B*NC,D], tgt [B*NT,D]: bf16 streams
 * as produced by mebt_embed_gather; lat and tgt are updated in place, ctx is read-only for the latent modes.
 * logits: [B*NT, V] in logits_dtype, or NULL to stop after the blocks.  Blocks that cannot reach the logits
 * (after the last latent_dec) are skipped.  All launches go to `stream`; nothing synchronises.
 */
int mebt_stack_forward(const mebt_layer_t* layers, int n_layers, const float* lnf_w, const float* lnf_b,
                       const void* w_head, int B, int L, int NC, int NT, int D, int H, int V, void* lat, void* ctx,
                       void* tgt, void* logits, int logits_dtype, void* workspace, size_t workspace_bytes, void* stream);
/* Same with the latent_enc K|V hoist (hoist may be NULL). Workspace: mebt_stack_forward_hoisted_workspace_bytes. */
int mebt_stack_forward_hoisted(const mebt_layer_t* layers, int n_layers, const float* lnf_w, const float* lnf_b,
                               const void* w_head, const mebt_enc_hoist_t* hoist, int B, int L, int NC, int NT, int D,
                               int H, int V, void* lat, void* ctx, void* tgt, void* logits, int logits_dtype,
                               void* workspace, size_t workspace_bytes, void* stream);
/* The same forward with the sampling step fused into the head GEMM (K6): when sample_ids != NULL, ids[b*NT + i] is one
 * categorical draw from softmax(logits / temperature) by the Gumbel-max rule - argmax_v(logit_v / T + G_v), G_v from a counter
 * hash of (seed, offset, row, v) - taken in the GEMM epilogue from the fp32 accumulators, so the [B*NT, V] logits are never
 * written (replaces mebt/modules/gpt.py:248 + sample_from_logits, mebt/transformer.py:843-889, on the draft / revise passes;
 * no top-k / top-p, no scores).  logits may be NULL. */
int mebt_stack_forward_sample(const mebt_layer_t* layers, int n_layers, const float* lnf_w, const float* lnf_b,
                              const void* w_head, const mebt_enc_hoist_t* hoist, int B, int L, int NC, int NT, int D,
                              int H, int V, void* lat, void* ctx, void* tgt, void* logits, int logits_dtype,
                              int64_t* sample_ids, float temperature, unsigned long long seed, unsigned long long offset,
                              void* workspace, size_t workspace_bytes, void* stream);
/* The head + sampling step alone: ids[r] ~ softmax(x[r] . w_head^T / temperature), x bf16 [rows, D] (= ln_f(targets)),
 * w_head bf16 [V, D]; workspace mebt_head_sample_workspace_bytes(rows). */
size_t mebt_head_sample_workspace_bytes(long long rows);
int mebt_head_sample(const void* x, int ldx, const void* w_head, int ldw, int rows, int V, int D, float temperature,
                     unsigned long long seed, unsigned long long offset, int64_t* ids, void* workspace,
                     size_t workspace_bytes, void* stream);

/* ---- training step: forward that saves activations + full backward -------------------------------------- */
/* fp32 gradient destinations of one Block, same geometry as mebt_layer_t (w_qkv = [3D,D] query|key|value rows). */
typedef struct mebt_layer_grads {
  float* ln1_w; float* ln1_b; float* ln2_w; float* ln2_b;
  float* w_qkv; float* b_qkv; float* w_proj; float* b_proj; float* w_fc1; float* b_fc1; float* w_fc2; float* b_fc2;
} mebt_layer_grads_t;

size_t mebt_stack_train_saved_bytes(const mebt_layer_t* layers, int n_layers, int B, int L, int NC, int NT, int D, int H);
size_t mebt_stack_backward_workspace_bytes(int B, int L, int NC, int NT, int D, int H);
/*
 * GPT.forward (mebt/modules/gpt.py:234-253) as run inside training_step (mebt/transformer.py:734): like
 * mebt_stack_forward, but the input streams are left untouched and every tensor backward needs (LayerNorm outputs
 * and statistics, projections, attention output + log-sum-exp, pre-GELU activations, each new stream version) is
 * kept in the caller's `saved` arena.  Only the four latent modes are supported (dropout p = 0).
 */
int mebt_stack_forward_train(const mebt_layer_t* layers, int n_layers, const float* lnf_w, const float* lnf_b,
                             const void* w_head, int B, int L, int NC, int NT, int D, int H, int V, const void* lat0,
                             const void* ctx, const void* tgt0, void* logits, int logits_dtype, void* saved,
                             size_t saved_bytes, void* stream);
/*
 * Backward of the above — what loss.backward() runs through the reference's stack.  dlogits: bf16 [B*NT, V]
 * (e.g. from mebt_masked_ce).  Weight/bias/LayerNorm gradients are written (grad_accumulate = 0) or added
 * (grad_accumulate = 1) to the fp32 destinations in `grads`, d_lnf_*, d_w_head.  d_lat/d_ctx/d_tgt: bf16 stream
 * gradients [B*L,D], [B*NC,D], [B*NT,D]; on return from the call that includes block 0 they hold the gradients
 * w.r.t. the stem outputs (feed them to mebt_embed_backward).  Blocks are processed in reverse over
 * [layer_begin, layer_end); the head is processed when layer_end == n_layers, so a data-parallel caller can issue
 * the backward in chunks and start the gradient all-reduce of a finished chunk while the next one runs.
 */
int mebt_stack_backward(const mebt_layer_t* layers, const mebt_layer_grads_t* grads, int n_layers, const float* lnf_w,
                        float* d_lnf_w, float* d_lnf_b, const void* w_head, float* d_w_head, int B, int L, int NC, int NT,
                        int D, int H, int V, const void* lat0, const void* ctx, const void* tgt0, const void* dlogits,
                        void* saved, size_t saved_bytes, void* d_lat, void* d_ctx, void* d_tgt, int layer_begin,
                        int layer_end, int grad_accumulate, void* workspace, size_t workspace_bytes, void* stream);

/* The same two calls in training mode with dropout.  attn_p: on the attention probabilities (attn_drop, gpt.py:136);
 * resid_p: on the proj and MLP outputs before their residual adds (resid_drop gpt.py:140, mlp[3] gpt.py:154).  The
 * stem dropout (embd_pdrop, gpt.py:239-242) is applied by the caller on the input streams (mebt_dropout_rows).
 * Block i draws its masks from sites (seed, 4*i + {0: attention, 1: proj, 2: mlp}); a NULL `drop` or p = 0 is the
 * deterministic path above.  Forward and backward must be given the same values. */
typedef struct mebt_dropout {
  float attn_p;
  float resid_p;
  unsigned long long seed;
} mebt_dropout_t;
int mebt_stack_forward_train_dropout(const mebt_layer_t* layers, int n_layers, const float* lnf_w, const float* lnf_b,
                                     const void* w_head, int B, int L, int NC, int NT, int D, int H, int V,
                                     const void* lat0, const void* ctx, const void* tgt0, void* logits, int logits_dtype,
                                     void* saved, size_t saved_bytes, const mebt_dropout_t* drop, void* stream);
int mebt_stack_backward_dropout(const mebt_layer_t* layers, const mebt_layer_grads_t* grads, int n_layers,
                                const float* lnf_w, float* d_lnf_w, float* d_lnf_b, const void* w_head, float* d_w_head,
                                int B, int L, int NC, int NT, int D, int H, int V, const void* lat0, const void* ctx,
                                const void* tgt0, const void* dlogits, void* saved, size_t saved_bytes, void* d_lat,
                                void* d_ctx, void* d_tgt, int layer_begin, int layer_end, int grad_accumulate,
                                const mebt_dropout_t* drop, void* workspace, size_t workspace_bytes, void* stream);

/* ---- VQGAN encoder / decoder convolutions (mebt/vqgan.py:263-405) ------------------------------------
 * Activations are channels-last bf16 [B, T, H, W, ld] (ld >= C, both multiples of 8).
 *
 * mebt_pad_norm_act: y = replicate_pad(act(norm(x))) - F.pad(..., mode='replicate') of SamePadConv3d /
 * SamePadConvTranspose3d (vqgan.py:381,404) fused with the Normalize + SiLU in front of it (ResBlock.forward :344-349,
 * Encoder / Decoder.final_block).  pad6 = (t_before, t_after, h_before, h_after, w_before, w_after); norm: 0 none,
 * 1 GroupNorm(groups, eps) with weight gamma / bias beta, 2 per-channel affine x * gamma + beta (eval-mode BatchNorm folded);
 * act: 0 none, 1 SiLU.  y has C real channels per position, [C, ldy) zeroed.  workspace: mebt_groupnorm_workspace_bytes
 * (partial sums per slab of positions + the per-(batch element, channel) scale / shift; C <= 2048, groups <= 256).
 *
 * mebt_conv3d_ndhwc: implicit-GEMM convolution on the tensor cores over an ALREADY PADDED input,
 *   y[b, t*ystep+yorigin, ..., co] = bias[co] + resid[b,t,h,w,co] + sum_{taps, c} xp[b, t*step + origin + dt, ..., c] * w[co][(dt,dh,dw)][c]
 * w: bf16 [ceil64(cout)][taps * ceil64(cin)] (zero padded per tap; zero rows behind cout).  nn.Conv3d: taps = kernel, step = stride.  nn.ConvTranspose3d
 * (kernel 4, stride 2): one launch per output parity with 2 taps, origin = ystep-origin = parity, ystep = 2 (the caller
 * packs the matching kernel slices).  odims3 = positions computed per dimension (must tile into 128-position patches).
 * Window mode (cin = 64 > ldx, one tap and unit step along w): the K slice of a position is the 64 consecutive ELEMENTS from
 * its first channel, i.e. the 64 / ldx positions from it on - the taps along w of a few-channel input (the RGB video) packed
 * into one k-block; w: [cout][taps_t * taps_h][(dw, c)], zero where dw >= the kernel width; rows padded by 64 / ldx - 1. */
size_t mebt_groupnorm_workspace_bytes(int B, int groups);
int mebt_pad_norm_act(const void* x, int ldx, void* y, int ldy, int B, int T, int H, int W, int C, const int* pad6, int norm,
                      int act, int groups, float eps, const float* gamma, const float* beta, void* workspace,
                      size_t workspace_bytes, void* stream);
int mebt_conv3d_ndhwc(const void* xp, int ldx, const int* xdims4, const void* w, int cin, const float* bias,
                      const void* resid, int ldr, void* y, int ldy, const int* ydims4, int cout, const int* taps3,
                      const int* step3, const int* origin3, const int* ystep3, const int* yorigin3, const int* odims3,
                      void* stream);

/* Backward with the optimizer step of the blocks' Linear weights FUSED into their weight-gradient GEMMs (single GPU, no
 * gradient accumulation): the epilogue of the grouped weight-gradient kernel applies torch's fused-AdamW arithmetic
 * (mebt_adamw_flat) to its fp32 accumulator tile - reads p / m / v, writes p / m / v and the bf16 operand copy - so the
 * gradients of those weights are never written and the optimizer's HBM traffic runs under the tensor work.  What the
 * reference does as loss.backward() followed by optimizer.step() (mebt/transformer.py:665-681, :749-798) for these tensors.
 * grad_base: the flat gradient buffer the dW pointers of mebt_layer_grads_t point into; p, m, v, p_bf16: flat buffers of the
 * same layout.  decay_blocks / block_shift: as mebt_adamw_flat.  step >= 1 is the step being taken.  Afterwards
 * mebt_adamw_flat must still run over every other parameter: its decay_blocks table takes bit 1 (value 2) = "skip,
 * already updated".  fuse == NULL: identical to mebt_stack_backward_dropout. */
typedef struct {
  const float* grad_base;
  float* p;
  float* m;
  float* v;
  void* p_bf16;
  const unsigned char* decay_blocks;
  int block_shift;
  float lr, beta1, beta2, eps, weight_decay;
  int step;
} mebt_fused_adamw_t;
int mebt_stack_backward_fused(const mebt_layer_t* layers, const mebt_layer_grads_t* grads, int n_layers,
                              const float* lnf_w, float* d_lnf_w, float* d_lnf_b, const void* w_head, float* d_w_head,
                              int B, int L, int NC, int NT, int D, int H, int V, const void* lat0, const void* ctx,
                              const void* tgt0, const void* dlogits, void* saved, size_t saved_bytes, void* d_lat,
                              void* d_ctx, void* d_tgt, int layer_begin, int layer_end, int grad_accumulate,
                              const mebt_dropout_t* drop, const mebt_fused_adamw_t* fuse, void* workspace,
                              size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MEBT_B200_H_ */

// Training-step engine: forward that keeps what backward needs, and the full backward of the layer stack, each as
// one C-ABI call (the per-op host cost is a kernel launch).  This is what torch autograd does for the reference's
// GPT.forward (mebt/modules/gpt.py:234-253) inside Net2NetTransformer.training_step (mebt/transformer.py:734-739).
//
// Per block (gpt.py:159-195), with q the stream the block rewrites and k its key source(s):
//   forward : qn = ln1(q), kn = ln1(k);  Q|K|V projections;  att = attention;  x = qn + proj(att);
//             h = ln2(x);  a = fc1 h + b;  u = gelu(a);  out = x + fc2 u
//   backward: da = (d_out W2) .* gelu'(a);  dh = da W1;  dx = d_out + ln2'(dh);  datt = dx Wp;
//             (dQ, dK, dV) = attention'(datt);  dqn = dx + dQKV Wqkv;  dkn = dKV Wkv;
//             d_q = ln1'(dqn) (assigned),  d_k += ln1'(dkn) (accumulated: a stream version may feed several blocks);
//             weight grads = (dY)^T X through the MN-major GEMM modes, bias grads = column sums.
#include <vector>

#include "common.cuh"

namespace mebt {

int gemm_bf16_aux(const void* A, int lda, int a_mn, const void* B, int ldb, int b_mn, void* C, int ldc, int M, int N,
                  int K, const float* bias, const void* residual, int ldres, void* aux, int ldaux, int flags,
                  cudaStream_t stream);
int layernorm(const void* x, int ldx, int in_dtype, const float* gamma, const float* beta, void* y, int ldy,
              int out_dtype, int rows, int D, float eps, float* mean, float* rstd, cudaStream_t st);
int layernorm_bwd(const void* dy, const void* x, const float* mean, const float* rstd, const float* gamma, void* dx,
                  int accumulate_dx, float* dgamma, float* dbeta, int accumulate_params, int rows, int D,
                  float* workspace, size_t ws_bytes, cudaStream_t st);
int layernorm_bwd_resid(const void* dy, const void* x, const float* mean, const float* rstd, const float* gamma,
                        void* dx, const void* resid, float* dgamma, float* dbeta, int accumulate_params, int rows, int D,
                        float* workspace, size_t ws_bytes, cudaStream_t st);
size_t layernorm_bwd_workspace_bytes(int D);
int colsum(const void* X, int ld, int rows, int N, float* out, int accumulate, float* workspace, size_t ws_bytes,
           cudaStream_t st);
int dropout_rows(const void* x, int ldx, const void* resid, int ldres, void* y, int ldy, int rows, int D, float p,
                 unsigned long long seed, unsigned long long site, cudaStream_t st);

namespace {

inline size_t al(size_t b) { return (b + 255) & ~size_t(255); }

struct LayerSaved {          // byte offsets into the saved-activation arena (SIZE_MAX = absent)
  size_t q_mean, q_rstd, k_mean, k_rstd, qn, kn, qkv, kv, att, lse, x, x_mean, x_rstd, h, a, u, out;
  int rq, rk;                // rows of the query stream / of the separately projected key source
  int nq;                    // queries per batch element
};

struct Plan {
  std::vector<LayerSaved> layers;
  size_t xf, f_mean, f_rstd;  // ln_f output and stats
  size_t total;
  int last;                   // last block that can reach the logits
};

Plan make_plan(const mebt_layer_t* layers, int n_layers, int B, int L, int NC, int NT, int D, int H) {
  Plan p;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += al(bytes); return o; };
  p.last = -1;
  for (int i = 0; i < n_layers; ++i)
    if (layers[i].mode == MEBT_MODE_LATENT_DEC) p.last = i;
  p.layers.resize(n_layers);
  for (int i = 0; i <= p.last; ++i) {
    LayerSaved s;
    const int mode = layers[i].mode;
    s.nq = mode == MEBT_MODE_LATENT_DEC ? NT : L;
    s.rq = B * s.nq;
    s.rk = mode == MEBT_MODE_LATENT_ENC ? B * NC : mode == MEBT_MODE_LATENT_DEC ? B * L : mode == MEBT_MODE_LT2L ? B * NT : 0;
    const bool fused_qkv = mode == MEBT_MODE_LATENT_SELF || mode == MEBT_MODE_LT2L;
    s.q_mean = take(size_t(s.rq) * 4); s.q_rstd = take(size_t(s.rq) * 4);
    s.k_mean = take(size_t(s.rk) * 4); s.k_rstd = take(size_t(s.rk) * 4);
    s.qn = take(size_t(s.rq) * D * 2);
    s.kn = take(size_t(s.rk) * D * 2);
    s.qkv = take(size_t(s.rq) * (fused_qkv ? 3 : 1) * D * 2);
    s.kv = take(size_t(s.rk) * 2 * D * 2);
    s.att = take(size_t(s.rq) * D * 2);
    s.lse = take(size_t(B) * H * s.nq * 4);
    s.x = take(size_t(s.rq) * D * 2);
    s.x_mean = take(size_t(s.rq) * 4); s.x_rstd = take(size_t(s.rq) * 4);
    s.h = take(size_t(s.rq) * D * 2);
    s.a = take(size_t(s.rq) * 4 * D * 2);
    s.u = take(size_t(s.rq) * 4 * D * 2);
    s.out = take(size_t(s.rq) * D * 2);
    p.layers[i] = s;
  }
  p.xf = take(size_t(B) * NT * D * 2);
  p.f_mean = take(size_t(B) * NT * 4);
  p.f_rstd = take(size_t(B) * NT * 4);
  p.total = off + 256;
  return p;
}

// Weight-gradient GEMMs and bias column sums only consume tensors the data-gradient chain has already produced, so
// they run on a library-owned side stream, concurrently with the next links of the chain (every GEMM of the 16-frame
// training step is at most one wave of tiles: two of them fit on the 148 SMs side by side).
struct SideStream {
  cudaStream_t stream = nullptr;
  cudaEvent_t fork[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t join = nullptr;
  bool ok = false;
};
SideStream& side_stream() {
  static SideStream s;
  if (!s.ok && s.stream == nullptr) {
    bool good = cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) == cudaSuccess;
    for (int i = 0; i < 4 && good; ++i) good = cudaEventCreateWithFlags(&s.fork[i], cudaEventDisableTiming) == cudaSuccess;
    good = good && cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming) == cudaSuccess;
    s.ok = good;
  }
  return s;
}

size_t backward_workspace_bytes(int B, int L, int NC, int NT, int D, int H) {
  const size_t rmax = size_t(B) * size_t(L > NT ? L : NT);
  const size_t rk = size_t(B) * size_t(NC > NT ? (NC > L ? NC : L) : (NT > L ? NT : L));
  size_t t = 0;
  t += al(rmax * 4 * D * 2);       // da
  t += al(rmax * D * 2);           // dh / datt
  t += al(rmax * 3 * D * 2);       // dqkv
  t += al(rk * 2 * D * 2);         // dkv
  t += al(rmax * D * 2);           // dqn
  t += al(rk * D * 2);             // dkn
  t += al(size_t(B) * H * (L > NT ? L : NT) * 4);   // attention delta
  t += al(rmax * D * 2);           // dx (kept separate from the stream gradient: the side stream still reads d_out)
  t += 2 * al(rmax * D * 2);       // dropout: gradients w.r.t. the pre-dropout MLP / proj outputs
  size_t red = layernorm_bwd_workspace_bytes(D);
  const size_t cs = size_t(64) * 16384 * 4;         // column-sum partials up to N = 16384
  t += 2 * al(red > cs ? red : cs);                 // one reduction scratch per stream
  return t + 4096;
}

}  // namespace
}  // namespace mebt

extern "C" {

size_t mebt_stack_train_saved_bytes(const mebt_layer_t* layers, int n_layers, int B, int L, int NC, int NT, int D, int H) {
  return mebt::make_plan(layers, n_layers, B, L, NC, NT, D, H).total;
}

size_t mebt_stack_backward_workspace_bytes(int B, int L, int NC, int NT, int D, int H) {
  return mebt::backward_workspace_bytes(B, L, NC, NT, D, H);
}

#define TRY(expr) do { int _rc = (expr); if (_rc != MEBT_OK) return _rc; } while (0)

int mebt_stack_forward_train(const mebt_layer_t* layers, int n_layers, const float* lnf_w, const float* lnf_b,
                             const void* w_head, int B, int L, int NC, int NT, int D, int H, int V, const void* lat0,
                             const void* ctx, const void* tgt0, void* logits, int logits_dtype, void* saved,
                             size_t saved_bytes, void* stream) {
  return mebt_stack_forward_train_dropout(layers, n_layers, lnf_w, lnf_b, w_head, B, L, NC, NT, D, H, V, lat0, ctx, tgt0,
                                          logits, logits_dtype, saved, saved_bytes, nullptr, stream);
}

int mebt_stack_forward_train_dropout(const mebt_layer_t* layers, int n_layers, const float* lnf_w, const float* lnf_b,
                                     const void* w_head, int B, int L, int NC, int NT, int D, int H, int V,
                                     const void* lat0, const void* ctx, const void* tgt0, void* logits, int logits_dtype,
                                     void* saved, size_t saved_bytes, const mebt_dropout_t* drop, void* stream) {
  using namespace mebt;
  const float attn_p = drop != nullptr ? drop->attn_p : 0.f, resid_p = drop != nullptr ? drop->resid_p : 0.f;
  const unsigned long long seed = drop != nullptr ? drop->seed : 0ull;
  MEBT_REQUIRE(attn_p >= 0.f && attn_p < 1.f && resid_p >= 0.f && resid_p < 1.f, MEBT_ERR_SHAPE, "forward_train: bad dropout p");
  MEBT_REQUIRE(B > 0 && L > 0 && NC >= 0 && NT > 0 && D == H * 64, MEBT_ERR_SHAPE, "forward_train: bad shape");
  for (int i = 0; i < n_layers; ++i)
    MEBT_REQUIRE(layers[i].mode >= MEBT_MODE_LATENT_ENC && layers[i].mode <= MEBT_MODE_LT2L, MEBT_ERR_UNSUPPORTED,
                 "forward_train: block %d has mode %d; only the four latent modes are trainable on this path", i,
                 layers[i].mode);
  const Plan plan = make_plan(layers, n_layers, B, L, NC, NT, D, H);
  MEBT_REQUIRE(saved != nullptr && saved_bytes >= plan.total, MEBT_ERR_WORKSPACE,
               "forward_train: saved-activation arena too small (%zu < %zu)", saved_bytes, plan.total);
  MEBT_REQUIRE(plan.last >= 0, MEBT_ERR_UNSUPPORTED, "forward_train: no latent_dec block, the logits do not depend on the stack");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  char* S = static_cast<char*>(saved);
  const void* lat = lat0;
  const void* tgt = tgt0;
  for (int i = 0; i <= plan.last; ++i) {
    const mebt_layer_t& w = layers[i];
    const LayerSaved& s = plan.layers[i];
    const __nv_bfloat16* wqkv = static_cast<const __nv_bfloat16*>(w.w_qkv);
    const __nv_bfloat16* w_kv = wqkv + size_t(D) * D;
    const float* b_kv = w.b_qkv + D;
    const void* q_in = w.mode == MEBT_MODE_LATENT_DEC ? tgt : lat;
    const void* k_in = w.mode == MEBT_MODE_LATENT_ENC ? ctx : w.mode == MEBT_MODE_LATENT_DEC ? lat : w.mode == MEBT_MODE_LT2L ? tgt : nullptr;
    void *qn = S + s.qn, *kn = S + s.kn, *qkv = S + s.qkv, *kv = S + s.kv, *att = S + s.att, *x = S + s.x, *h = S + s.h,
         *a = S + s.a, *u = S + s.u, *out = S + s.out;
    float* lse = reinterpret_cast<float*>(S + s.lse);
    const bool fused = w.mode == MEBT_MODE_LATENT_SELF || w.mode == MEBT_MODE_LT2L;
    TRY(layernorm(q_in, D, MEBT_DTYPE_BF16, w.ln1_w, w.ln1_b, qn, D, MEBT_DTYPE_BF16, s.rq, D, 1e-5f,
                  reinterpret_cast<float*>(S + s.q_mean), reinterpret_cast<float*>(S + s.q_rstd), st));
    const int qw = fused ? 3 * D : D;
    TRY(gemm_bf16_aux(qn, D, 0, wqkv, D, 0, qkv, qw, s.rq, qw, D, w.b_qkv, nullptr, 0, nullptr, 0, 0, st));
    if (s.rk > 0) {
      TRY(layernorm(k_in, D, MEBT_DTYPE_BF16, w.ln1_w, w.ln1_b, kn, D, MEBT_DTYPE_BF16, s.rk, D, 1e-5f,
                    reinterpret_cast<float*>(S + s.k_mean), reinterpret_cast<float*>(S + s.k_rstd), st));
      TRY(gemm_bf16_aux(kn, D, 0, w_kv, D, 0, kv, 2 * D, s.rk, 2 * D, D, b_kv, nullptr, 0, nullptr, 0, 0, st));
    }
    const int nk_sep = s.rk / B;
    const unsigned long long site = 4ull * i;      // + 0: attention, + 1: proj, + 2: mlp
    if (fused)
      TRY(mebt_latent_attention_fwd_dropout(qkv, 3 * D, 0, qkv, 3 * D, D, 2 * D, L, s.rk > 0 ? kv : nullptr, 2 * D, 0, D,
                                            nk_sep, att, D, lse, B, H, s.nq, 64, attn_p, seed + site, stream));
    else
      TRY(mebt_latent_attention_fwd_dropout(qkv, D, 0, s.rk > 0 ? kv : nullptr, 2 * D, 0, D, nk_sep, nullptr, 0, 0, 0, 0,
                                            att, D, lse, B, H, s.nq, 64, attn_p, seed + site, stream));
    if (resid_p > 0.f) {     // x = qn + drop(att Wp + b): the GEMM leaves the pre-dropout value, the dropout kernel adds qn
      TRY(gemm_bf16_aux(att, D, 0, w.w_proj, D, 0, x, D, s.rq, D, D, w.b_proj, nullptr, 0, nullptr, 0, 0, st));
      TRY(dropout_rows(x, D, qn, D, x, D, s.rq, D, resid_p, seed, site + 1, st));
    } else {
      TRY(gemm_bf16_aux(att, D, 0, w.w_proj, D, 0, x, D, s.rq, D, D, w.b_proj, qn, D, nullptr, 0, 0, st));
    }
    TRY(layernorm(x, D, MEBT_DTYPE_BF16, w.ln2_w, w.ln2_b, h, D, MEBT_DTYPE_BF16, s.rq, D, 1e-5f,
                  reinterpret_cast<float*>(S + s.x_mean), reinterpret_cast<float*>(S + s.x_rstd), st));
    TRY(gemm_bf16_aux(h, D, 0, w.w_fc1, D, 0, u, 4 * D, s.rq, 4 * D, D, w.b_fc1, nullptr, 0, a, 4 * D, MEBT_GEMM_GELU, st));
    if (resid_p > 0.f) {     // out = x + drop(u W2 + b)
      TRY(gemm_bf16_aux(u, 4 * D, 0, w.w_fc2, 4 * D, 0, out, D, s.rq, D, 4 * D, w.b_fc2, nullptr, 0, nullptr, 0, 0, st));
      TRY(dropout_rows(out, D, x, D, out, D, s.rq, D, resid_p, seed, site + 2, st));
    } else {
      TRY(gemm_bf16_aux(u, 4 * D, 0, w.w_fc2, 4 * D, 0, out, D, s.rq, D, 4 * D, w.b_fc2, x, D, nullptr, 0, 0, st));
    }
    if (w.mode == MEBT_MODE_LATENT_DEC) tgt = out; else lat = out;
  }
  TRY(layernorm(tgt, D, MEBT_DTYPE_BF16, lnf_w, lnf_b, S + plan.xf, D, MEBT_DTYPE_BF16, B * NT, D, 1e-5f,
                reinterpret_cast<float*>(S + plan.f_mean), reinterpret_cast<float*>(S + plan.f_rstd), st));
  TRY(gemm_bf16_aux(S + plan.xf, D, 0, w_head, D, 0, logits, V, B * NT, V, D, nullptr, nullptr, 0, nullptr, 0,
                    logits_dtype == MEBT_DTYPE_FP32 ? MEBT_GEMM_OUT_FP32 : 0, st));
  return MEBT_OK;
}

int mebt_stack_backward(const mebt_layer_t* layers, const mebt_layer_grads_t* grads, int n_layers, const float* lnf_w,
                        float* d_lnf_w, float* d_lnf_b, const void* w_head, float* d_w_head, int B, int L, int NC, int NT,
                        int D, int H, int V, const void* lat0, const void* ctx, const void* tgt0, const void* dlogits,
                        void* saved, size_t saved_bytes, void* d_lat, void* d_ctx, void* d_tgt, int layer_begin,
                        int layer_end, int grad_accumulate, void* workspace, size_t workspace_bytes, void* stream) {
  return mebt_stack_backward_dropout(layers, grads, n_layers, lnf_w, d_lnf_w, d_lnf_b, w_head, d_w_head, B, L, NC, NT, D, H, V,
                                     lat0, ctx, tgt0, dlogits, saved, saved_bytes, d_lat, d_ctx, d_tgt, layer_begin,
                                     layer_end, grad_accumulate, nullptr, workspace, workspace_bytes, stream);
}

int mebt_stack_backward_dropout(const mebt_layer_t* layers, const mebt_layer_grads_t* grads, int n_layers,
                                const float* lnf_w, float* d_lnf_w, float* d_lnf_b, const void* w_head, float* d_w_head,
                                int B, int L, int NC, int NT, int D, int H, int V, const void* lat0, const void* ctx,
                                const void* tgt0, const void* dlogits, void* saved, size_t saved_bytes, void* d_lat,
                                void* d_ctx, void* d_tgt, int layer_begin, int layer_end, int grad_accumulate,
                                const mebt_dropout_t* drop, void* workspace, size_t workspace_bytes, void* stream) {
  using namespace mebt;
  const float attn_p = drop != nullptr ? drop->attn_p : 0.f, resid_p = drop != nullptr ? drop->resid_p : 0.f;
  const unsigned long long seed = drop != nullptr ? drop->seed : 0ull;
  MEBT_REQUIRE(attn_p >= 0.f && attn_p < 1.f && resid_p >= 0.f && resid_p < 1.f, MEBT_ERR_SHAPE, "backward: bad dropout p");
  MEBT_REQUIRE(B > 0 && L > 0 && NC >= 0 && NT > 0 && D == H * 64, MEBT_ERR_SHAPE, "backward: bad shape");
  MEBT_REQUIRE(0 <= layer_begin && layer_begin <= layer_end && layer_end <= n_layers, MEBT_ERR_SHAPE, "backward: bad layer range");
  const Plan plan = make_plan(layers, n_layers, B, L, NC, NT, D, H);
  MEBT_REQUIRE(saved != nullptr && saved_bytes >= plan.total, MEBT_ERR_WORKSPACE, "backward: saved arena too small");
  MEBT_REQUIRE(workspace != nullptr && workspace_bytes >= backward_workspace_bytes(B, L, NC, NT, D, H), MEBT_ERR_WORKSPACE,
               "backward: workspace too small (%zu < %zu)", workspace_bytes, backward_workspace_bytes(B, L, NC, NT, D, H));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  char* S = static_cast<char*>(saved);
  char* W = static_cast<char*>(workspace);
  const size_t rmax = size_t(B) * size_t(L > NT ? L : NT);
  const size_t rkmax = size_t(B) * size_t(NC > NT ? (NC > L ? NC : L) : (NT > L ? NT : L));
  size_t off = 0;
  auto take = [&](size_t bytes) { char* p = W + off; off += al(bytes); return p; };
  void* da = take(rmax * 4 * D * 2);
  void* dh = take(rmax * D * 2);
  void* dqkv = take(rmax * 3 * D * 2);
  void* dkv = take(rkmax * 2 * D * 2);
  void* dqn = take(rmax * D * 2);
  void* dkn = take(rkmax * D * 2);
  const size_t delta_bytes = size_t(B) * H * (L > NT ? L : NT) * 4;
  void* delta = take(delta_bytes);
  void* dxb = take(rmax * D * 2);
  void* dy_mlp = take(rmax * D * 2);     // d_out .* keep/(1-p): gradient w.r.t. the pre-dropout MLP output
  void* dy_proj = take(rmax * D * 2);    // dx .* keep/(1-p): gradient w.r.t. the pre-dropout proj output
  const size_t red_each = ((workspace_bytes - off - 512) / 2) & ~size_t(255);
  const size_t red_bytes = red_each;
  float* red = reinterpret_cast<float*>(W + off);
  float* red_side = reinterpret_cast<float*>(W + off + red_each);
  const int acc = grad_accumulate ? 1 : 0;
  SideStream& side = side_stream();
  MEBT_REQUIRE(side.ok, MEBT_ERR_CUDA, "backward: cannot create the weight-gradient side stream");
  cudaStream_t sst = side.stream;

  // weight gradient dW[N_out, K_in] (+)= dY^T X : A = dY stored [rows, N_out] (MN-major), B = X stored [rows, K_in] (MN-major)
  auto WGRAD_ON = [&](cudaStream_t on, const void* dY, int ld_dy, const void* X, int ldx, float* dWt, int ldw, int n_out,
                      int k_in, int rows, int accumulate) {
    return gemm_bf16_aux(dY, ld_dy, 1, X, ldx, 1, dWt, ldw, n_out, k_in, rows, nullptr, nullptr, 0, nullptr, 0,
                         MEBT_GEMM_OUT_FP32 | (accumulate ? MEBT_GEMM_ACCUMULATE : 0), on);
  };
  auto WGRAD = [&](const void* dY, int ld_dy, const void* X, int ldx, float* dWt, int ldw, int n_out, int k_in, int rows,
                   int accumulate) { return WGRAD_ON(st, dY, ld_dy, X, ldx, dWt, ldw, n_out, k_in, rows, accumulate); };
  // weight + bias gradient of one nn.Linear on the side stream, after everything recorded so far on the main stream
  auto SIDE_LINEAR = [&](int slot, const void* dY, int ld_dy, const void* X, int ldx, float* dWt, int ldw, float* db,
                         int n_out, int k_in, int rows, int accumulate) -> int {
    MEBT_CUDA_OK(cudaEventRecord(side.fork[slot], st));
    MEBT_CUDA_OK(cudaStreamWaitEvent(sst, side.fork[slot], 0));
    int rc2 = WGRAD_ON(sst, dY, ld_dy, X, ldx, dWt, ldw, n_out, k_in, rows, accumulate);
    if (rc2) return rc2;
    return colsum(dY, ld_dy, rows, n_out, db, accumulate, red_side, red_bytes, sst);
  };
  // data gradient dX[rows, K_in] = dY[rows, N_out] W[N_out, K_in] : B = W stored [K_red = N_out, N = K_in] (MN-major)
  auto DGRAD = [&](const void* dY, int ld_dy, const void* Wt, int ldw, void* dX, int rows, int k_in, int n_out,
                   const void* residual, void* aux, int ldaux, int flags) {
    return gemm_bf16_aux(dY, ld_dy, 0, Wt, ldw, 1, dX, k_in, rows, k_in, n_out, nullptr, residual, k_in, aux, ldaux, flags, st);
  };

  if (layer_end == n_layers) {
    // ---- head + ln_f (gpt.py:247-248) ----
    const int rows = B * NT;
    void* d_xf = dh;
    TRY(DGRAD(dlogits, V, w_head, D, d_xf, rows, D, V, nullptr, nullptr, 0, 0));
    TRY(WGRAD(dlogits, V, S + plan.xf, D, d_w_head, D, V, D, rows, acc));
    // the final targets stream is the `out` of the last latent_dec block
    const void* tgt_final = S + plan.layers[plan.last].out;
    TRY(layernorm_bwd(d_xf, tgt_final, reinterpret_cast<float*>(S + plan.f_mean), reinterpret_cast<float*>(S + plan.f_rstd),
                      lnf_w, d_tgt, 0, d_lnf_w, d_lnf_b, acc, rows, D, red, red_bytes, st));
    MEBT_CUDA_OK(cudaMemsetAsync(d_lat, 0, size_t(B) * L * D * 2, st));
    if (NC > 0) MEBT_CUDA_OK(cudaMemsetAsync(d_ctx, 0, size_t(B) * NC * D * 2, st));
    if (!acc) {
      // blocks that cannot reach the logits get exact-zero gradients
      for (int i = plan.last + 1; i < n_layers; ++i) {
        const mebt_layer_grads_t& g = grads[i];
        MEBT_CUDA_OK(cudaMemsetAsync(g.w_qkv, 0, size_t(3) * D * D * 4, st));
        MEBT_CUDA_OK(cudaMemsetAsync(g.b_qkv, 0, size_t(3) * D * 4, st));
        MEBT_CUDA_OK(cudaMemsetAsync(g.w_proj, 0, size_t(D) * D * 4, st));
        MEBT_CUDA_OK(cudaMemsetAsync(g.b_proj, 0, size_t(D) * 4, st));
        MEBT_CUDA_OK(cudaMemsetAsync(g.w_fc1, 0, size_t(4) * D * D * 4, st));
        MEBT_CUDA_OK(cudaMemsetAsync(g.b_fc1, 0, size_t(4) * D * 4, st));
        MEBT_CUDA_OK(cudaMemsetAsync(g.w_fc2, 0, size_t(4) * D * D * 4, st));
        MEBT_CUDA_OK(cudaMemsetAsync(g.b_fc2, 0, size_t(D) * 4, st));
        for (float* v : {g.ln1_w, g.ln1_b, g.ln2_w, g.ln2_b}) MEBT_CUDA_OK(cudaMemsetAsync(v, 0, size_t(D) * 4, st));
      }
    }
  }

  // stream versions feeding each block: recompute the forward chain of `out` pointers
  std::vector<const void*> lat_in(n_layers, nullptr), tgt_in(n_layers, nullptr);
  {
    const void* lat = lat0;
    const void* tgt = tgt0;
    for (int i = 0; i <= plan.last; ++i) {
      lat_in[i] = lat; tgt_in[i] = tgt;
      if (layers[i].mode == MEBT_MODE_LATENT_DEC) tgt = S + plan.layers[i].out; else lat = S + plan.layers[i].out;
    }
  }

  for (int i = (layer_end - 1 < plan.last ? layer_end - 1 : plan.last); i >= layer_begin; --i) {
    const mebt_layer_t& w = layers[i];
    const mebt_layer_grads_t& g = grads[i];
    const LayerSaved& s = plan.layers[i];
    const int mode = w.mode;
    const bool fused = mode == MEBT_MODE_LATENT_SELF || mode == MEBT_MODE_LT2L;
    const __nv_bfloat16* wqkv = static_cast<const __nv_bfloat16*>(w.w_qkv);
    const __nv_bfloat16* w_kv = wqkv + size_t(D) * D;
    void* d_out = mode == MEBT_MODE_LATENT_DEC ? d_tgt : d_lat;      // gradient w.r.t. this block's output stream
    const void* q_in = mode == MEBT_MODE_LATENT_DEC ? tgt_in[i] : lat_in[i];
    const void* k_in = mode == MEBT_MODE_LATENT_ENC ? ctx : mode == MEBT_MODE_LATENT_DEC ? lat_in[i] : mode == MEBT_MODE_LT2L ? tgt_in[i] : nullptr;
    void* d_k_stream = mode == MEBT_MODE_LATENT_ENC ? d_ctx : mode == MEBT_MODE_LATENT_DEC ? d_lat : mode == MEBT_MODE_LT2L ? d_tgt : nullptr;
    const int rq = s.rq, rk = s.rk;

    const unsigned long long site = 4ull * i;
    // ---- MLP ----  (main stream: the data-gradient chain; side stream: weight and bias gradients)
    const void* d_mlp = d_out;                      // gradient w.r.t. (u W2 + b): d_out through the dropout mask
    if (resid_p > 0.f) {
      TRY(dropout_rows(d_out, D, nullptr, 0, dy_mlp, D, rq, D, resid_p, seed, site + 2, st));
      d_mlp = dy_mlp;
    }
    TRY(SIDE_LINEAR(0, d_mlp, D, S + s.u, 4 * D, g.w_fc2, 4 * D, g.b_fc2, D, 4 * D, rq, acc));
    TRY(DGRAD(d_mlp, D, w.w_fc2, 4 * D, da, rq, 4 * D, D, nullptr, S + s.a, 4 * D, MEBT_GEMM_DGELU));     // da
    TRY(SIDE_LINEAR(1, da, 4 * D, S + s.h, D, g.w_fc1, D, g.b_fc1, 4 * D, D, rq, acc));
    TRY(DGRAD(da, 4 * D, w.w_fc1, D, dh, rq, D, 4 * D, nullptr, nullptr, 0, 0));                            // dh
    // dx = d_out + ln2'(dh), written to its own buffer (d_out is still being read by the side stream)
    void* dx = dxb;
    TRY(layernorm_bwd_resid(dh, S + s.x, reinterpret_cast<float*>(S + s.x_mean), reinterpret_cast<float*>(S + s.x_rstd),
                            w.ln2_w, dx, d_out, g.ln2_w, g.ln2_b, acc, rq, D, red, red_bytes, st));
    // ---- attention output projection ----
    void* datt = dh;
    const void* d_proj = dx;                        // gradient w.r.t. (att Wp + b)
    if (resid_p > 0.f) {
      TRY(dropout_rows(dx, D, nullptr, 0, dy_proj, D, rq, D, resid_p, seed, site + 1, st));
      d_proj = dy_proj;
    }
    TRY(SIDE_LINEAR(2, d_proj, D, S + s.att, D, g.w_proj, D, g.b_proj, D, D, rq, acc));
    TRY(DGRAD(d_proj, D, w.w_proj, D, datt, rq, D, D, nullptr, nullptr, 0, 0));
    // ---- attention ----
    const int nk_sep = rk / B;
    const float* lse = reinterpret_cast<float*>(S + s.lse);
    if (fused)
      TRY(mebt_latent_attention_bwd_dropout(S + s.qkv, 3 * D, 0, S + s.qkv, 3 * D, D, 2 * D, L, rk > 0 ? S + s.kv : nullptr,
                                            2 * D, 0, D, nk_sep, S + s.att, D, datt, D, lse, dqkv, 3 * D, 0, dqkv, 3 * D, D,
                                            2 * D, rk > 0 ? dkv : nullptr, 2 * D, 0, D, B, H, s.nq, 64, attn_p, seed + site,
                                            delta, delta_bytes, stream));
    else
      TRY(mebt_latent_attention_bwd_dropout(S + s.qkv, D, 0, rk > 0 ? S + s.kv : nullptr, 2 * D, 0, D, nk_sep, nullptr, 0, 0,
                                            0, 0, S + s.att, D, datt, D, lse, dqkv, D, 0, rk > 0 ? dkv : nullptr, 2 * D, 0,
                                            D, nullptr, 0, 0, 0, B, H, s.nq, 64, attn_p, seed + site, delta, delta_bytes,
                                            stream));
    // ---- q/k/v projections ----
    const int qw = fused ? 3 * D : D;
    TRY(SIDE_LINEAR(3, dqkv, qw, S + s.qn, D, g.w_qkv, D, g.b_qkv, qw, D, rq, acc));
    if (rk > 0) {   // same side stream, after the q-side gradients: they may accumulate into the same rows (lt2l)
      TRY(WGRAD_ON(sst, dkv, 2 * D, S + s.kn, D, g.w_qkv + size_t(D) * D, D, 2 * D, D, rk, fused ? 1 : acc));
      TRY(colsum(dkv, 2 * D, rk, 2 * D, g.b_qkv + D, fused ? 1 : acc, red_side, red_bytes, sst));
    } else if (!fused && !acc) {
      // latent_enc with no context: key/value projections receive exact-zero gradients
      MEBT_CUDA_OK(cudaMemsetAsync(g.w_qkv + size_t(D) * D, 0, size_t(2) * D * D * 4, sst));
      MEBT_CUDA_OK(cudaMemsetAsync(g.b_qkv + D, 0, size_t(2) * D * 4, sst));
    }
    TRY(DGRAD(dqkv, qw, wqkv, D, dqn, rq, D, qw, dx, nullptr, 0, 0));                                       // dqn = dx + dQKV Wqkv
    if (rk > 0) TRY(DGRAD(dkv, 2 * D, w_kv, D, dkn, rk, D, 2 * D, nullptr, nullptr, 0, 0));
    // the side stream must be done with d_out / da / dx / dqkv / dkv before they are overwritten
    MEBT_CUDA_OK(cudaEventRecord(side.join, sst));
    MEBT_CUDA_OK(cudaStreamWaitEvent(st, side.join, 0));
    // ---- ln1 on both streams ----
    TRY(layernorm_bwd(dqn, q_in, reinterpret_cast<float*>(S + s.q_mean), reinterpret_cast<float*>(S + s.q_rstd), w.ln1_w,
                      d_out, 0, g.ln1_w, g.ln1_b, acc, rq, D, red, red_bytes, st));                        // assigns d(q stream)
    if (rk > 0)
      TRY(layernorm_bwd(dkn, k_in, reinterpret_cast<float*>(S + s.k_mean), reinterpret_cast<float*>(S + s.k_rstd), w.ln1_w,
                        d_k_stream, 1, g.ln1_w, g.ln1_b, 1, rk, D, red, red_bytes, st));                   // accumulates
  }
  return MEBT_OK;
}

#undef TRY

}  // extern "C"

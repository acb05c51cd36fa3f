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
#include <cstdlib>
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
int gemm_bf16_drop(const void* A, int lda, int a_mn, const void* B, int ldb, int b_mn, void* C, int ldc, int M, int N,
                   int K, const float* bias, const void* residual, int ldres, void* aux, int ldaux, int flags,
                   const DropKey* drop, cudaStream_t stream);
int gemm_bf16_ex(const void* A, int lda, int a_mn, const void* B, int ldb, int b_mn, void* C, int ldc, int M, int N,
                 int K, const float* bias, const void* residual, int ldres, void* aux, int ldaux, int flags,
                 const DropKey* drop, float* delta_out, int delta_nq, int delta_h, cudaStream_t stream);
int latent_attention_bwd_launch(const void* Q, int ldq, int q_col0, const void* KV1, int ld1, int k1_col0,
                                int v1_col0, int NK1, const void* KV2, int ld2, int k2_col0, int v2_col0, int NK2,
                                const void* O, int ldo, const void* dO, int lddo, const float* lse, void* dQ,
                                int lddq, int dq_col0, void* dKV1, int ldd1, int dk1_col0, int dv1_col0, void* dKV2,
                                int ldd2, int dk2_col0, int dv2_col0, int B, int H, int NQ, int head_dim, float drop_p,
                                unsigned long long drop_seed, void* workspace, size_t workspace_bytes, int delta_ready,
                                void* stream);
struct WgradDesc {           // csrc/gemm_grouped.cu
  const void* dY; int ld_dy;
  const void* X; int ldx;
  float* dW; int ldw;
  int n_out, k_in, rows, accumulate;
};
int gemm_grouped_wgrad(const WgradDesc* d, int n, cudaStream_t stream);
struct WgradDescEx {         // csrc/gemm_grouped.cu: optional second reduction segment dW (+)= dY^T X + dY2^T X2
  const void* dY; int ld_dy;
  const void* X; int ldx;
  float* dW; int ldw;
  int n_out, k_in, rows, accumulate;
  const void* dY2; int ld_dy2;
  const void* X2; int ldx2;
  int rows2;
};
struct AdamFuseHost {        // csrc/gemm_grouped.cu: the optimizer step in the weight-gradient epilogue
  const float* grad_base; float* p; float* m; float* v; void* p16; const unsigned char* decay; int shift;
  float lr, beta1, beta2, eps, wd; int step;
};
int gemm_grouped_wgrad_ex(const WgradDescEx* d, int n, const AdamFuseHost* fuse, cudaStream_t stream);
int layernorm_bwd_params(const void* dy, const void* x, const float* mean, const float* rstd, float* dgamma, float* dbeta,
                         int accumulate_params, int rows, int D, float* workspace, size_t ws_bytes, cudaStream_t st);
int layernorm_bwd_dx(const void* dy, const void* x, const float* mean, const float* rstd, const float* gamma, void* dx,
                     const void* resid, int rows, int D, void* dx_drop, const DropKey* drop, cudaStream_t st);

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

// Weight-gradient GEMMs, bias column sums and LayerNorm parameter gradients only consume tensors the data-gradient
// chain has already produced, so they run on two library-owned side streams (A: the tcgen05 weight-gradient GEMMs,
// B: the small reductions), concurrently with the next links of the chain: every kernel of the 16-frame training step
// is at most about one wave of CTAs, so the chain is bound by per-kernel latency, not by SM time.
constexpr int OWED_RING = 4;          // depth of the masked-copy ring (dy_mlp): writers never wait for recent readers
struct SideStreams {
  cudaStream_t a = nullptr, b = nullptr, c = nullptr;
  cudaEvent_t fork[8] = {};
  cudaEvent_t join_a = nullptr, join_b = nullptr, join_c = nullptr;
  // fc2_*[ring slot]: the readers of a block's masked stream gradient are done (A: the grouped weight gradients,
  // B: the mlp.2 bias gradient).  end_*[block parity]: all of the block's side work is done (it reads the scratch
  // buffers the block after next reuses).  kdgrad_c[parity]: the key-side data gradient (stream C) exists.
  cudaEvent_t fc2_a[OWED_RING] = {}, fc2_b[OWED_RING] = {}, end_a[2] = {}, end_b[2] = {}, end_c[2] = {}, kdgrad_c[2] = {};
  cudaEvent_t qdgrad[2] = {};           // fused optimizer: the block's last reader of its bf16 weights on the main stream is done
  std::vector<cudaEvent_t> fwd_k;       // forward: block i's key/value projection (stream C) is ready
  bool ok = false;
};
SideStreams& side_streams() {
  static SideStreams s;
  if (!s.ok && s.a == nullptr) {
    bool good = cudaStreamCreateWithFlags(&s.a, cudaStreamNonBlocking) == cudaSuccess &&
                cudaStreamCreateWithFlags(&s.b, cudaStreamNonBlocking) == cudaSuccess &&
                cudaStreamCreateWithFlags(&s.c, cudaStreamNonBlocking) == cudaSuccess;
    auto make = [&](cudaEvent_t* e, int n) {
      for (int i = 0; i < n && good; ++i) good = cudaEventCreateWithFlags(&e[i], cudaEventDisableTiming) == cudaSuccess;
    };
    make(s.fork, 8);
    make(&s.join_a, 1); make(&s.join_b, 1); make(&s.join_c, 1);
    make(s.fc2_a, OWED_RING); make(s.fc2_b, OWED_RING);
    make(s.end_a, 2); make(s.end_b, 2); make(s.end_c, 2); make(s.kdgrad_c, 2); make(s.qdgrad, 2);
    s.ok = good;
  }
  return s;
}
bool ensure_fwd_events(SideStreams& s, int n) {
  while (int(s.fwd_k.size()) < n) {
    cudaEvent_t e;
    if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return false;
    s.fwd_k.push_back(e);
  }
  return true;
}

// Which LayerNorm backward finishes the gradient w.r.t. block i's output stream (the last accumulation before block i
// reads it): the first later block that reads that stream, through its query side (0) or key side (1), or ln_f (2)
// for the last latent_dec block; -1 when nothing reads it.  That kernel also writes the gradient multiplied by block
// i's mlp-dropout mask, which is what block i's first two GEMMs consume.
struct FinalWriter { int block; int side; };
FinalWriter final_writer(const mebt_layer_t* layers, int last, int i) {
  const bool tgt_stream = layers[i].mode == MEBT_MODE_LATENT_DEC;
  for (int j = i + 1; j <= last; ++j) {
    const int m = layers[j].mode;
    if (tgt_stream) {
      if (m == MEBT_MODE_LATENT_DEC) return {j, 0};
      if (m == MEBT_MODE_LT2L) return {j, 1};
    } else {
      if (m == MEBT_MODE_LATENT_DEC) return {j, 1};
      return {j, 0};                       // latent_enc / latent_self / lt2l read the latents as queries
    }
  }
  if (i == last) return {last + 1, 2};
  return {-1, -1};
}

// Every scratch tensor of a block's backward exists twice (by block parity): the side streams may still be reading
// block i's buffers while the main stream runs block i-1, and block i-2 only starts after block i's side work.
struct BwdWorkspace {
  void *da[2], *dh[2], *datt[2], *dqkv[2], *dkv[2], *dqn[2], *dkn[2], *dxb[2], *dy_proj[2], *delta, *dy_mlp[2][OWED_RING];
  float *red, *red_b;
  size_t red_bytes, delta_bytes, total;
};
BwdWorkspace carve_backward_workspace(char* W, int B, int L, int NC, int NT, int D, int H) {
  const size_t rmax = size_t(B) * size_t(L > NT ? L : NT);
  const size_t rk = size_t(B) * size_t(NC > NT ? (NC > L ? NC : L) : (NT > L ? NT : L));
  size_t off = 0;
  auto take = [&](size_t bytes) { char* p = W + off; off += al(bytes); return static_cast<void*>(p); };
  BwdWorkspace w;
  for (int i = 0; i < 2; ++i) {
    w.da[i] = take(rmax * 4 * D * 2);
    w.dh[i] = take(rmax * D * 2);
    w.datt[i] = take(rmax * D * 2);        // separate from dh: side stream B still reads dh (ln2 parameter gradients)
    w.dqkv[i] = take(rmax * 3 * D * 2);
    w.dkv[i] = take(rk * 2 * D * 2);
    w.dqn[i] = take(rmax * D * 2);
    w.dkn[i] = take(rk * D * 2);
    w.dxb[i] = take(rmax * D * 2);         // dx, separate from the stream gradient (the side streams still read d_out)
    w.dy_proj[i] = take(rmax * D * 2);     // dx .* keep/(1-p): gradient w.r.t. the pre-dropout proj output
  }
  w.delta_bytes = size_t(B) * H * (L > NT ? L : NT) * 4;
  w.delta = take(w.delta_bytes);
  for (int i = 0; i < OWED_RING; ++i) {          // [stream][ring slot of the consuming block]
    w.dy_mlp[0][i] = take(size_t(B) * L * D * 2);     // d(latents) .* keep/(1-p): gradient w.r.t. the pre-dropout MLP output
    w.dy_mlp[1][i] = take(size_t(B) * NT * D * 2);    // the same for the targets stream
  }
  size_t red = layernorm_bwd_workspace_bytes(D);
  const size_t cs = size_t(64) * 16384 * 4;         // column-sum partials up to N = 16384
  w.red_bytes = al(red > cs ? red : cs);
  w.red = static_cast<float*>(take(w.red_bytes));
  w.red_b = static_cast<float*>(take(w.red_bytes));
  w.total = off + 4096;
  return w;
}

size_t backward_workspace_bytes(int B, int L, int NC, int NT, int D, int H) {
  return carve_backward_workspace(nullptr, B, L, NC, NT, D, H).total;
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
  // The key side of a block (ln1 of the key stream + its K|V projection) does not depend on the query side: it runs
  // on library stream C, next to the query side on the caller's stream, and the attention kernel waits for both.  The
  // contexts never change through the stack (gpt.py:243-245), so the key sides of ALL latent_enc blocks are issued
  // up front and are off the critical path altogether.
  SideStreams& side = side_streams();
  MEBT_REQUIRE(side.ok && ensure_fwd_events(side, n_layers), MEBT_ERR_CUDA, "forward_train: cannot create the side streams");
  cudaStream_t sc = side.c;
  int fork_slot = 0;
  auto FORK_C = [&]() -> int {
    cudaEvent_t e = side.fork[fork_slot];
    fork_slot = (fork_slot + 1) & 7;
    MEBT_CUDA_OK(cudaEventRecord(e, st));
    MEBT_CUDA_OK(cudaStreamWaitEvent(sc, e, 0));
    return MEBT_OK;
  };
  auto KEY_SIDE = [&](int i, const void* k_in) -> int {
    const mebt_layer_t& w = layers[i];
    const LayerSaved& s = plan.layers[i];
    const __nv_bfloat16* w_kv = static_cast<const __nv_bfloat16*>(w.w_qkv) + size_t(D) * D;
    int rc2 = layernorm(k_in, D, MEBT_DTYPE_BF16, w.ln1_w, w.ln1_b, S + s.kn, D, MEBT_DTYPE_BF16, s.rk, D, 1e-5f,
                        reinterpret_cast<float*>(S + s.k_mean), reinterpret_cast<float*>(S + s.k_rstd), sc);
    if (rc2) return rc2;
    rc2 = gemm_bf16_aux(S + s.kn, D, 0, w_kv, D, 0, S + s.kv, 2 * D, s.rk, 2 * D, D, w.b_qkv + D, nullptr, 0, nullptr, 0, 0, sc);
    if (rc2) return rc2;
    MEBT_CUDA_OK(cudaEventRecord(side.fwd_k[i], sc));
    return MEBT_OK;
  };
  if (NC > 0) {
    TRY(FORK_C());
    for (int i = 0; i <= plan.last; ++i)
      if (layers[i].mode == MEBT_MODE_LATENT_ENC) TRY(KEY_SIDE(i, ctx));
  }
  const void* lat = lat0;
  const void* tgt = tgt0;
  for (int i = 0; i <= plan.last; ++i) {
    const mebt_layer_t& w = layers[i];
    const LayerSaved& s = plan.layers[i];
    const __nv_bfloat16* wqkv = static_cast<const __nv_bfloat16*>(w.w_qkv);
    const void* q_in = w.mode == MEBT_MODE_LATENT_DEC ? tgt : lat;
    const void* k_in = w.mode == MEBT_MODE_LATENT_ENC ? ctx : w.mode == MEBT_MODE_LATENT_DEC ? lat : w.mode == MEBT_MODE_LT2L ? tgt : nullptr;
    void *qn = S + s.qn, *qkv = S + s.qkv, *kv = S + s.kv, *att = S + s.att, *x = S + s.x, *h = S + s.h,
         *a = S + s.a, *u = S + s.u, *out = S + s.out;
    float* lse = reinterpret_cast<float*>(S + s.lse);
    const bool fused = w.mode == MEBT_MODE_LATENT_SELF || w.mode == MEBT_MODE_LT2L;
    if (s.rk > 0 && w.mode != MEBT_MODE_LATENT_ENC) {
      TRY(FORK_C());
      TRY(KEY_SIDE(i, k_in));
    }
    TRY(layernorm(q_in, D, MEBT_DTYPE_BF16, w.ln1_w, w.ln1_b, qn, D, MEBT_DTYPE_BF16, s.rq, D, 1e-5f,
                  reinterpret_cast<float*>(S + s.q_mean), reinterpret_cast<float*>(S + s.q_rstd), st));
    const int qw = fused ? 3 * D : D;
    TRY(gemm_bf16_aux(qn, D, 0, wqkv, D, 0, qkv, qw, s.rq, qw, D, w.b_qkv, nullptr, 0, nullptr, 0, 0, st));
    if (s.rk > 0) MEBT_CUDA_OK(cudaStreamWaitEvent(st, side.fwd_k[i], 0));
    const int nk_sep = s.rk / B;
    const unsigned long long site = 4ull * i;      // + 0: attention, + 1: proj, + 2: mlp
    if (fused)
      TRY(mebt_latent_attention_fwd_dropout(qkv, 3 * D, 0, qkv, 3 * D, D, 2 * D, L, s.rk > 0 ? kv : nullptr, 2 * D, 0, D,
                                            nk_sep, att, D, lse, B, H, s.nq, 64, attn_p, seed + site, stream));
    else
      TRY(mebt_latent_attention_fwd_dropout(qkv, D, 0, s.rk > 0 ? kv : nullptr, 2 * D, 0, D, nk_sep, nullptr, 0, 0, 0, 0,
                                            att, D, lse, B, H, s.nq, 64, attn_p, seed + site, stream));
    // x = qn + drop(att Wp + b): the keep factors are applied in the GEMM epilogue, between the bias and the residual
    const DropKey k_proj = make_drop_key(resid_p, seed, site + 1), k_mlp = make_drop_key(resid_p, seed, site + 2);
    TRY(gemm_bf16_drop(att, D, 0, w.w_proj, D, 0, x, D, s.rq, D, D, w.b_proj, qn, D, nullptr, 0, 0,
                       resid_p > 0.f ? &k_proj : nullptr, st));
    TRY(layernorm(x, D, MEBT_DTYPE_BF16, w.ln2_w, w.ln2_b, h, D, MEBT_DTYPE_BF16, s.rq, D, 1e-5f,
                  reinterpret_cast<float*>(S + s.x_mean), reinterpret_cast<float*>(S + s.x_rstd), st));
    TRY(gemm_bf16_aux(h, D, 0, w.w_fc1, D, 0, u, 4 * D, s.rq, 4 * D, D, w.b_fc1, nullptr, 0, a, 4 * D, MEBT_GEMM_GELU, st));
    TRY(gemm_bf16_drop(u, 4 * D, 0, w.w_fc2, 4 * D, 0, out, D, s.rq, D, 4 * D, w.b_fc2, x, D, nullptr, 0, 0,
                       resid_p > 0.f ? &k_mlp : nullptr, st));               // out = x + drop(u W2 + b)
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
  return mebt_stack_backward_fused(layers, grads, n_layers, lnf_w, d_lnf_w, d_lnf_b, w_head, d_w_head, B, L, NC, NT, D, H, V,
                                   lat0, ctx, tgt0, dlogits, saved, saved_bytes, d_lat, d_ctx, d_tgt, layer_begin, layer_end,
                                   grad_accumulate, drop, nullptr, workspace, workspace_bytes, stream);
}

int mebt_stack_backward_fused(const mebt_layer_t* layers, const mebt_layer_grads_t* grads, int n_layers,
                              const float* lnf_w, float* d_lnf_w, float* d_lnf_b, const void* w_head, float* d_w_head,
                              int B, int L, int NC, int NT, int D, int H, int V, const void* lat0, const void* ctx,
                              const void* tgt0, const void* dlogits, void* saved, size_t saved_bytes, void* d_lat,
                              void* d_ctx, void* d_tgt, int layer_begin, int layer_end, int grad_accumulate,
                              const mebt_dropout_t* drop, const mebt_fused_adamw_t* fuse, void* workspace,
                              size_t workspace_bytes, void* stream) {
  using namespace mebt;
  // experiment knob: SMs the persistent kernels of the data-gradient chain may take (the rest stays with the side-stream
  // weight-gradient launch, whose grid MEBT_WGRAD_CTAS caps)
  static const int chain_cap = getenv("MEBT_BWD_CHAIN_SMS") != nullptr ? atoi(getenv("MEBT_BWD_CHAIN_SMS")) : 0;
  GridCapScope cap_scope(chain_cap);
  // The optimizer step of the blocks' Linear weights inside their weight-gradient GEMMs: every block must then produce
  // each weight's whole gradient in ONE accumulator (no accumulation into earlier gradients, no second launch adding
  // the key|value rows of lt2l: a two-segment problem instead) and must not overwrite bf16 weights a data-gradient GEMM
  // of the same block still reads (the grouped launch waits for them).
  MEBT_REQUIRE(fuse == nullptr || (!grad_accumulate && NC > 0 && fuse->step >= 1 && fuse->grad_base != nullptr &&
                                   fuse->p != nullptr && fuse->m != nullptr && fuse->v != nullptr && fuse->p_bf16 != nullptr &&
                                   fuse->decay_blocks != nullptr && D % 256 == 0),
               MEBT_ERR_UNSUPPORTED, "backward: the fused optimizer step needs NC > 0, D %% 256 == 0, no gradient accumulation and all flat buffers");
  AdamFuseHost fuse_h;
  if (fuse != nullptr)
    fuse_h = AdamFuseHost{fuse->grad_base, fuse->p, fuse->m, fuse->v, fuse->p_bf16, fuse->decay_blocks, fuse->block_shift,
                          fuse->lr, fuse->beta1, fuse->beta2, fuse->eps, fuse->weight_decay, fuse->step};
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
  const BwdWorkspace ws = carve_backward_workspace(static_cast<char*>(workspace), B, L, NC, NT, D, H);
  void* delta = ws.delta;
  const size_t delta_bytes = ws.delta_bytes, red_bytes = ws.red_bytes;
  float* red_b = ws.red_b;            // reduction scratch of side stream B (ws.red: main stream, unused by this engine now)
  const int acc = grad_accumulate ? 1 : 0;
  SideStreams& side = side_streams();
  MEBT_REQUIRE(side.ok, MEBT_ERR_CUDA, "backward: cannot create the weight-gradient side streams");
  cudaStream_t sa = side.a, sb = side.b, sc = side.c;
  int fork_slot = 0;
  static const bool fence_forks = getenv("MEBT_FORK_NOPDL") != nullptr;
  // everything issued so far on the main stream happens before whatever is issued next on the two side streams
  auto FORK = [&]() -> int {
    cudaEvent_t e = side.fork[fork_slot];
    fork_slot = (fork_slot + 1) & 7;
    MEBT_CUDA_OK(cudaEventRecord(e, st));
    if (fence_forks) pdl_fence(st);
    MEBT_CUDA_OK(cudaStreamWaitEvent(sa, e, 0));
    MEBT_CUDA_OK(cudaStreamWaitEvent(sb, e, 0));
    return MEBT_OK;
  };
  auto FORK_C = [&]() -> int {
    cudaEvent_t e = side.fork[fork_slot];
    fork_slot = (fork_slot + 1) & 7;
    MEBT_CUDA_OK(cudaEventRecord(e, st));
    MEBT_CUDA_OK(cudaStreamWaitEvent(sc, e, 0));
    return MEBT_OK;
  };
  // the side streams' work so far is marked by (ea on A, eb on B) ... / ... and the main stream waits for such a mark
  auto MARK = [&](cudaEvent_t ea, cudaEvent_t eb) -> int {
    MEBT_CUDA_OK(cudaEventRecord(ea, sa));
    MEBT_CUDA_OK(cudaEventRecord(eb, sb));
    return MEBT_OK;
  };
  auto AWAIT = [&](cudaEvent_t ea, cudaEvent_t eb) -> int {
    MEBT_CUDA_OK(cudaStreamWaitEvent(st, ea, 0));
    MEBT_CUDA_OK(cudaStreamWaitEvent(st, eb, 0));
    return MEBT_OK;
  };
  // ... and everything issued so far on the side streams happens before whatever is issued next on the main stream
  auto JOIN = [&]() -> int {
    MEBT_CUDA_OK(cudaEventRecord(side.join_a, sa));
    MEBT_CUDA_OK(cudaEventRecord(side.join_b, sb));
    MEBT_CUDA_OK(cudaEventRecord(side.join_c, sc));
    MEBT_CUDA_OK(cudaStreamWaitEvent(st, side.join_a, 0));
    MEBT_CUDA_OK(cudaStreamWaitEvent(st, side.join_b, 0));
    MEBT_CUDA_OK(cudaStreamWaitEvent(st, side.join_c, 0));
    return MEBT_OK;
  };

  // weight gradient dW[N_out, K_in] (+)= dY^T X : A = dY stored [rows, N_out] (MN-major), B = X stored [rows, K_in] (MN-major)
  auto WGRAD_ON = [&](cudaStream_t on, const void* dY, int ld_dy, const void* X, int ldx, float* dWt, int ldw, int n_out,
                      int k_in, int rows, int accumulate) {
    return gemm_bf16_aux(dY, ld_dy, 1, X, ldx, 1, dWt, ldw, n_out, k_in, rows, nullptr, nullptr, 0, nullptr, 0,
                         MEBT_GEMM_OUT_FP32 | (accumulate ? MEBT_GEMM_ACCUMULATE : 0), on);
  };
  // data gradient dX[rows, K_in] = dY[rows, N_out] W[N_out, K_in] : B = W stored [K_red = N_out, N = K_in] (MN-major)
  auto DGRAD = [&](const void* dY, int ld_dy, const void* Wt, int ldw, void* dX, int rows, int k_in, int n_out,
                   const void* residual, void* aux, int ldaux, int flags) {
    return gemm_bf16_aux(dY, ld_dy, 0, Wt, ldw, 1, dX, k_in, rows, k_in, n_out, nullptr, residual, k_in, aux, ldaux, flags, st);
  };
  // the masked copy the LayerNorm backward `side` (0 query / 1 key / 2 ln_f) of block `j` owes to an earlier block
  // (with resid_p = 0 the copy is unmasked: block i's GEMMs then read a buffer nobody overwrites behind their back)
  struct Owed { void* dst; DropKey key; int consumer; };
  auto owed_by = [&](int j, int which) {
    Owed o{nullptr, make_drop_key(0.f, 0ull, 0ull), -1};
    for (int i = 0; i <= plan.last && i < j; ++i) {
      const FinalWriter fw = final_writer(layers, plan.last, i);
      if (fw.block == j && fw.side == which) {
        o.dst = ws.dy_mlp[layers[i].mode == MEBT_MODE_LATENT_DEC ? 1 : 0][i % OWED_RING];
        o.key = make_drop_key(resid_p, seed, 4ull * i + 2);
        o.consumer = i;
      }
    }
    return o;
  };
  // before the copy owed to block `consumer` is written (on stream `on`): the previous readers of that ring slot
  // (side work of block consumer + OWED_RING, issued long ago) must be done
  auto AWAIT_OWED = [&](const Owed& o, cudaStream_t on) -> int {
    if (o.consumer < 0) return MEBT_OK;
    MEBT_CUDA_OK(cudaStreamWaitEvent(on, side.fc2_a[o.consumer % OWED_RING], 0));
    MEBT_CUDA_OK(cudaStreamWaitEvent(on, side.fc2_b[o.consumer % OWED_RING], 0));
    return MEBT_OK;
  };

  if (layer_end == n_layers) {
    // ---- head + ln_f (gpt.py:247-248) ----
    const int rows = B * NT;
    const int hp = (plan.last + 1) & 1;           // the head takes the scratch parity of a block behind the last one
    void* d_xf = ws.dh[hp];
    TRY(DGRAD(dlogits, V, w_head, D, d_xf, rows, D, V, nullptr, nullptr, 0, 0));
    TRY(FORK());
    TRY(WGRAD_ON(sa, dlogits, V, S + plan.xf, D, d_w_head, D, V, D, rows, acc));
    // the final targets stream is the `out` of the last latent_dec block
    const void* tgt_final = S + plan.layers[plan.last].out;
    TRY(layernorm_bwd_params(d_xf, tgt_final, reinterpret_cast<float*>(S + plan.f_mean), reinterpret_cast<float*>(S + plan.f_rstd),
                             d_lnf_w, d_lnf_b, acc, rows, D, red_b, red_bytes, sb));
    const Owed o = owed_by(plan.last + 1, 2);
    TRY(AWAIT_OWED(o, st));
    TRY(layernorm_bwd_dx(d_xf, tgt_final, reinterpret_cast<float*>(S + plan.f_mean), reinterpret_cast<float*>(S + plan.f_rstd),
                         lnf_w, d_tgt, nullptr, rows, D, o.dst, &o.key, st));
    TRY(MARK(side.end_a[hp], side.end_b[hp]));
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
    const int par = i & 1;
    void *da = ws.da[par], *dh = ws.dh[par], *datt = ws.datt[par], *dqkv = ws.dqkv[par], *dkv = ws.dkv[par],
         *dqn = ws.dqn[par], *dkn = ws.dkn[par], *dxb = ws.dxb[par], *dy_proj = ws.dy_proj[par];
    const int ring = i % OWED_RING;
    // this block reuses the scratch buffers of block i + 2: that block's side work (issued two blocks ago) must be done
    TRY(AWAIT(side.end_a[par], side.end_b[par]));
    MEBT_CUDA_OK(cudaStreamWaitEvent(st, side.end_c[par], 0));

    const unsigned long long site = 4ull * i;
    // ---- MLP ----  (main stream: the data-gradient chain; side stream A: ONE grouped launch of the block's weight
    // gradients once its last operand exists; side stream B: bias and LayerNorm parameter gradients as they become ready)
    // d_mlp = d_out through the mlp-dropout mask = gradient w.r.t. (u W2 + b): a copy written by the LayerNorm backward
    // that finished d_out (see final_writer), so that this block's last kernels may overwrite d_out while the side
    // streams still read d_mlp; made here when no such kernel exists
    void* d_mlp = ws.dy_mlp[mode == MEBT_MODE_LATENT_DEC ? 1 : 0][ring];
    if (final_writer(layers, plan.last, i).block < 0) {
      TRY(AWAIT(side.fc2_a[ring], side.fc2_b[ring]));
      TRY(dropout_rows(d_out, D, nullptr, 0, d_mlp, D, rq, D, resid_p, seed, site + 2, st));
    }
    auto SIDE_COLSUM = [&](const void* dY, int ld_dy, int rows, int n_out, float* db, int accumulate) -> int {
      int rc2 = FORK();
      if (rc2) return rc2;
      return colsum(dY, ld_dy, rows, n_out, db, accumulate, red_b, red_bytes, sb);
    };
    TRY(SIDE_COLSUM(d_mlp, D, rq, D, g.b_fc2, acc));
    MEBT_CUDA_OK(cudaEventRecord(side.fc2_b[ring], sb));
    TRY(DGRAD(d_mlp, D, w.w_fc2, 4 * D, da, rq, 4 * D, D, nullptr, S + s.a, 4 * D, MEBT_GEMM_DGELU));     // da
    TRY(SIDE_COLSUM(da, 4 * D, rq, 4 * D, g.b_fc1, acc));
    TRY(DGRAD(da, 4 * D, w.w_fc1, D, dh, rq, D, 4 * D, nullptr, nullptr, 0, 0));                            // dh
    // dx = d_out + ln2'(dh), in its own buffer; the same kernel writes dx through the proj-dropout mask = the gradient
    // w.r.t. (att Wp + b)
    void* dx = dxb;
    TRY(FORK());
    TRY(layernorm_bwd_params(dh, S + s.x, reinterpret_cast<float*>(S + s.x_mean), reinterpret_cast<float*>(S + s.x_rstd),
                             g.ln2_w, g.ln2_b, acc, rq, D, red_b, red_bytes, sb));
    const DropKey k_proj = make_drop_key(resid_p, seed, site + 1);
    TRY(layernorm_bwd_dx(dh, S + s.x, reinterpret_cast<float*>(S + s.x_mean), reinterpret_cast<float*>(S + s.x_rstd),
                         w.ln2_w, dx, d_out, rq, D, resid_p > 0.f ? dy_proj : nullptr, &k_proj, st));
    // ---- attention output projection ----
    const void* d_proj = resid_p > 0.f ? dy_proj : dx;
    TRY(SIDE_COLSUM(d_proj, D, rq, D, g.b_proj, acc));
    // datt = d_proj Wp; the same epilogue emits delta = rowsum_head(datt .* att) for the attention backward
    TRY(gemm_bf16_ex(d_proj, D, 0, w.w_proj, D, 1, datt, D, rq, D, D, nullptr, nullptr, 0, S + s.att, D, 0, nullptr,
                     static_cast<float*>(delta), s.nq, H, st));
    // ---- attention ----
    const int nk_sep = rk / B;
    const float* lse = reinterpret_cast<float*>(S + s.lse);
    if (fused)
      TRY(latent_attention_bwd_launch(S + s.qkv, 3 * D, 0, S + s.qkv, 3 * D, D, 2 * D, L, rk > 0 ? S + s.kv : nullptr,
                                      2 * D, 0, D, nk_sep, S + s.att, D, datt, D, lse, dqkv, 3 * D, 0, dqkv, 3 * D, D,
                                      2 * D, rk > 0 ? dkv : nullptr, 2 * D, 0, D, B, H, s.nq, 64, attn_p, seed + site,
                                      delta, delta_bytes, 1, stream));
    else
      TRY(latent_attention_bwd_launch(S + s.qkv, D, 0, rk > 0 ? S + s.kv : nullptr, 2 * D, 0, D, nk_sep, nullptr, 0, 0,
                                      0, 0, S + s.att, D, datt, D, lse, dqkv, D, 0, rk > 0 ? dkv : nullptr, 2 * D, 0,
                                      D, nullptr, 0, 0, 0, B, H, s.nq, 64, attn_p, seed + site, delta, delta_bytes, 1,
                                      stream));
    // ---- weight gradients of the whole block: one grouped launch on side stream A ----
    const int qw = fused ? 3 * D : D;
    WgradDescEx wx[5];
    int nx = 0;
    if (fuse == nullptr) {
      WgradDesc wd[5];
      int n = 0;
      wd[n++] = WgradDesc{d_mlp, D, S + s.u, 4 * D, g.w_fc2, 4 * D, D, 4 * D, rq, acc};
      wd[n++] = WgradDesc{da, 4 * D, S + s.h, D, g.w_fc1, D, 4 * D, D, rq, acc};
      wd[n++] = WgradDesc{d_proj, D, S + s.att, D, g.w_proj, D, D, D, rq, acc};
      wd[n++] = WgradDesc{dqkv, qw, S + s.qn, D, g.w_qkv, D, qw, D, rq, acc};
      // key|value rows of a separately projected key source: disjoint from the query rows unless the block also
      // projected keys / values from its own stream (lt2l), in which case they are added by a second launch
      if (rk > 0 && !fused) wd[n++] = WgradDesc{dkv, 2 * D, S + s.kn, D, g.w_qkv + size_t(D) * D, D, 2 * D, D, rk, acc};
      TRY(FORK());
      if (rk > 0) MEBT_CUDA_OK(cudaStreamWaitEvent(sc, side.fork[(fork_slot + 7) & 7], 0));   // stream C forks here too
      TRY(gemm_grouped_wgrad(wd, n, sa));
      if (rk > 0 && fused) {
        const WgradDesc kv{dkv, 2 * D, S + s.kn, D, g.w_qkv + size_t(D) * D, D, 2 * D, D, rk, 1};
        TRY(gemm_grouped_wgrad(&kv, 1, sa));
      }
      MEBT_CUDA_OK(cudaEventRecord(side.fc2_a[ring], sa));
    } else {
      // fused optimizer step: launched further down, once the qkv data gradients (the last readers of this block's bf16
      // weights) are done.  lt2l: the key|value rows take both of their sources as one two-segment reduction.
      const __nv_bfloat16* dqkv_b = static_cast<const __nv_bfloat16*>(dqkv);
      wx[nx++] = WgradDescEx{d_mlp, D, S + s.u, 4 * D, g.w_fc2, 4 * D, D, 4 * D, rq, 0, nullptr, 0, nullptr, 0, 0};
      wx[nx++] = WgradDescEx{da, 4 * D, S + s.h, D, g.w_fc1, D, 4 * D, D, rq, 0, nullptr, 0, nullptr, 0, 0};
      wx[nx++] = WgradDescEx{d_proj, D, S + s.att, D, g.w_proj, D, D, D, rq, 0, nullptr, 0, nullptr, 0, 0};
      if (fused && rk > 0) {
        wx[nx++] = WgradDescEx{dqkv, qw, S + s.qn, D, g.w_qkv, D, D, D, rq, 0, nullptr, 0, nullptr, 0, 0};
        wx[nx++] = WgradDescEx{dqkv_b + D, qw, S + s.qn, D, g.w_qkv + size_t(D) * D, D, 2 * D, D, rq, 0, dkv, 2 * D, S + s.kn, D, rk};
      } else {
        wx[nx++] = WgradDescEx{dqkv, qw, S + s.qn, D, g.w_qkv, D, qw, D, rq, 0, nullptr, 0, nullptr, 0, 0};
        if (rk > 0) wx[nx++] = WgradDescEx{dkv, 2 * D, S + s.kn, D, g.w_qkv + size_t(D) * D, D, 2 * D, D, rk, 0, nullptr, 0, nullptr, 0, 0};
      }
      TRY(FORK());
      if (rk > 0) MEBT_CUDA_OK(cudaStreamWaitEvent(sc, side.fork[(fork_slot + 7) & 7], 0));   // stream C forks here too
    }
    {
      TRY(colsum(dqkv, qw, rq, qw, g.b_qkv, acc, red_b, red_bytes, sb));
      if (rk > 0) {
        TRY(colsum(dkv, 2 * D, rk, 2 * D, g.b_qkv + D, fused ? 1 : acc, red_b, red_bytes, sb));
      } else if (!fused && !acc) {
        // latent_enc with no context: key/value projections receive exact-zero gradients
        MEBT_CUDA_OK(cudaMemsetAsync(g.w_qkv + size_t(D) * D, 0, size_t(2) * D * D * 4, sa));
        MEBT_CUDA_OK(cudaMemsetAsync(g.b_qkv + D, 0, size_t(2) * D * 4, sb));
      }
    }
    // ---- key side, on stream C next to the query side: data gradient of the K|V projection, then ln1 backward into
    // the key stream's gradient (and the masked copy an earlier block's GEMMs will read).  For latent_enc the key
    // stream is the contexts, whose gradient only the embedding backward reads: never joined before the end.
    if (rk > 0) {
      TRY(gemm_bf16_aux(dkv, 2 * D, 0, w_kv, D, 1, dkn, D, rk, D, 2 * D, nullptr, nullptr, 0, nullptr, 0, 0, sc));   // dkn
      MEBT_CUDA_OK(cudaEventRecord(side.kdgrad_c[par], sc));
      const Owed o = owed_by(i, 1);
      TRY(AWAIT_OWED(o, sc));
      TRY(layernorm_bwd_dx(dkn, k_in, reinterpret_cast<float*>(S + s.k_mean), reinterpret_cast<float*>(S + s.k_rstd), w.ln1_w,
                           d_k_stream, d_k_stream, rk, D, o.dst, &o.key, sc));                             // accumulates
      MEBT_CUDA_OK(cudaEventRecord(side.end_c[par], sc));
    }
    TRY(DGRAD(dqkv, qw, wqkv, D, dqn, rq, D, qw, dx, nullptr, 0, 0));                                       // dqn = dx + dQKV Wqkv
    if (fuse != nullptr) {
      MEBT_CUDA_OK(cudaEventRecord(side.qdgrad[par], st));
      MEBT_CUDA_OK(cudaStreamWaitEvent(sa, side.qdgrad[par], 0));
      if (rk > 0) MEBT_CUDA_OK(cudaStreamWaitEvent(sa, side.kdgrad_c[par], 0));
      TRY(gemm_grouped_wgrad_ex(wx, nx, &fuse_h, sa));
      MEBT_CUDA_OK(cudaEventRecord(side.fc2_a[ring], sa));
    }
    // ---- ln1, query side ----  parameter gradients on side stream B (query side, then key side: they add into the
    // same vectors); the dx kernel also writes the masked copy an earlier block's GEMMs will read
    TRY(FORK());
    TRY(layernorm_bwd_params(dqn, q_in, reinterpret_cast<float*>(S + s.q_mean), reinterpret_cast<float*>(S + s.q_rstd),
                             g.ln1_w, g.ln1_b, acc, rq, D, red_b, red_bytes, sb));
    if (rk > 0) {
      MEBT_CUDA_OK(cudaStreamWaitEvent(sb, side.kdgrad_c[par], 0));
      TRY(layernorm_bwd_params(dkn, k_in, reinterpret_cast<float*>(S + s.k_mean), reinterpret_cast<float*>(S + s.k_rstd),
                               g.ln1_w, g.ln1_b, 1, rk, D, red_b, red_bytes, sb));
    }
    {
      const Owed o = owed_by(i, 0);
      TRY(AWAIT_OWED(o, st));
      TRY(layernorm_bwd_dx(dqn, q_in, reinterpret_cast<float*>(S + s.q_mean), reinterpret_cast<float*>(S + s.q_rstd), w.ln1_w,
                           d_out, nullptr, rq, D, o.dst, &o.key, st));                                     // assigns d(q stream)
    }
    // the next block (backward order) consumes the key stream's gradient unless that stream is the contexts
    if (rk > 0 && mode != MEBT_MODE_LATENT_ENC) MEBT_CUDA_OK(cudaStreamWaitEvent(st, side.end_c[par], 0));
    TRY(MARK(side.end_a[par], side.end_b[par]));
  }
  // the caller's stream sees every gradient of this call (the all-reduce / optimizer is ordered after it)
  TRY(JOIN());
  return MEBT_OK;
}

#undef TRY

}  // extern "C"

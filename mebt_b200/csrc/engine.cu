// Whole-stack forward in one C-ABI call: the 24-block MeBT layer stack launched back to back from C++ so that
// the per-op host cost is a kernel launch, not a Python round trip (and the sequence can be captured in a CUDA
// graph by the caller, since every launch goes to the caller's stream and nothing synchronises).
//
// reference: GPT.forward, mebt/modules/gpt.py:234-253, and Block.forward, :159-195:
//   q^ = ln1(q), k^ = ln1(k) (same ln1), x = q^ + proj(attn(q^, k^)), out = x + fc2(gelu(fc1(ln2(x)))).
#include "common.cuh"

namespace mebt {

int gemm_bf16(const void* A, int lda, int a_mn, const void* B, int ldb, int b_mn, void* C, int ldc, int M, int N,
              int K, const float* bias, const void* residual, int ldres, int flags, cudaStream_t stream);
int gemm_bf16_sample(const void* A, int lda, const void* B, int ldb, int M, int N, int K, float temperature,
                     unsigned long long seed, unsigned long long offset, unsigned long long* packed, cudaStream_t stream);
size_t latent_attention_fwd_workspace_bytes(int B, int H, int NQ);
int latent_attention_fwd(const void* Q, int ldq, int q_col0, const void* KV1, int ld1, int k1_col0, int v1_col0, int NK1,
                         const void* KV2, int ld2, int k2_col0, int v2_col0, int NK2, void* O, int ldo, float* lse, int B,
                         int H, int NQ, int head_dim, float drop_p, unsigned long long drop_seed, void* workspace,
                         size_t workspace_bytes, void* stream);
int layernorm(const void* x, int ldx, int in_dtype, const float* gamma, const float* beta, void* y, int ldy,
              int out_dtype, int rows, int D, float eps, float* mean, float* rstd, cudaStream_t st);

namespace {

// fused head + sampling: the GEMM's packed (key << 32 | id) minima before / after
__global__ void packed_init_kernel(unsigned long long* __restrict__ packed, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) packed[i] = ~0ull;
}
__global__ void packed_ids_kernel(const unsigned long long* __restrict__ packed, int64_t* __restrict__ ids, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) ids[i] = int64_t(packed[i] & 0xFFFFFFFFull);
}

struct Workspace {
  char* base;
  size_t used, cap;
  void* take(size_t bytes) {
    bytes = (bytes + 255) & ~size_t(255);
    void* p = base + used;
    used += bytes;
    return used <= cap ? p : nullptr;
  }
};

// Scratch of the split-KV attention (csrc/attention.cu): only launches with at most half an SM count of work items
// split, i.e. small batches; sized for the largest query count such a launch of this stack can have.
size_t attention_split_bytes(int B, int L, int NC, int NT, int D) {
  const int H = D / 64;
  size_t need = 0;
  for (int nq : {L, NT, NC + NT}) {
    const int pairs = ((nq + 127) / 128 + 1) / 2;
    if (nq > 0 && 2 * B * H * pairs <= sm_count()) {
      const size_t b = latent_attention_fwd_workspace_bytes(B, H, nq);
      need = b > need ? b : need;
    }
  }
  return (need + 255) & ~size_t(255);
}

size_t stack_workspace_bytes(int B, int L, int NC, int NT, int D, int n_enc_hoisted = 0) {
  const size_t q_rows = size_t(B) * size_t(L > NC + NT ? L : NC + NT);
  const size_t k_rows = q_rows;
  size_t total = 0;
  auto add = [&](size_t rows, size_t cols) { total += ((rows * cols * 2) + 255) & ~size_t(255); };
  add(q_rows, D);        // qn
  add(k_rows, D);        // kn
  add(q_rows, 3 * D);    // q / qkv
  add(k_rows, 2 * D);    // kv
  add(q_rows, D);        // att
  add(q_rows, D);        // x
  add(q_rows, D);        // h
  add(q_rows, 4 * D);    // u
  add(q_rows, D);        // maskgit concat stream
  add(size_t(B) * NC, size_t(n_enc_hoisted) * 2 * D);   // hoisted K|V of every latent_enc block
  total += attention_split_bytes(B, L, NC, NT, D);     // split-KV partials (small batches only)
  total += (size_t(B) * NT * 8 + 255) & ~size_t(255);  // packed draws of the fused head + sampling form
  return total + 4096;
}

}  // namespace
}  // namespace mebt

extern "C" {

size_t mebt_stack_forward_workspace_bytes(int B, int L, int NC, int NT, int D) {
  return mebt::stack_workspace_bytes(B, L, NC, NT, D);
}

size_t mebt_stack_forward_hoisted_workspace_bytes(int B, int L, int NC, int NT, int D, int n_enc) {
  return mebt::stack_workspace_bytes(B, L, NC, NT, D, n_enc);
}

int mebt_stack_forward(const mebt_layer_t* layers, int n_layers, const float* lnf_w, const float* lnf_b,
                       const void* w_head, int B, int L, int NC, int NT, int D, int H, int V, void* lat, void* ctx,
                       void* tgt, void* logits, int logits_dtype, void* workspace, size_t workspace_bytes,
                       void* stream) {
  return mebt_stack_forward_hoisted(layers, n_layers, lnf_w, lnf_b, w_head, nullptr, B, L, NC, NT, D, H, V, lat, ctx, tgt,
                                    logits, logits_dtype, workspace, workspace_bytes, stream);
}

int mebt_stack_forward_hoisted(const mebt_layer_t* layers, int n_layers, const float* lnf_w, const float* lnf_b,
                               const void* w_head, const mebt_enc_hoist_t* hoist, int B, int L, int NC, int NT, int D,
                               int H, int V, void* lat, void* ctx, void* tgt, void* logits, int logits_dtype,
                               void* workspace, size_t workspace_bytes, void* stream) {
  return mebt_stack_forward_sample(layers, n_layers, lnf_w, lnf_b, w_head, hoist, B, L, NC, NT, D, H, V, lat, ctx, tgt, logits,
                                   logits_dtype, nullptr, 0.f, 0ull, 0ull, workspace, workspace_bytes, stream);
}

int mebt_stack_forward_sample(const mebt_layer_t* layers, int n_layers, const float* lnf_w, const float* lnf_b,
                              const void* w_head, const mebt_enc_hoist_t* hoist, int B, int L, int NC, int NT, int D,
                              int H, int V, void* lat, void* ctx, void* tgt, void* logits, int logits_dtype,
                              int64_t* sample_ids, float temperature, unsigned long long seed, unsigned long long offset,
                              void* workspace, size_t workspace_bytes, void* stream) {
  using namespace mebt;
  const int n_hoist = (hoist != nullptr && NC > 0) ? hoist->n_enc : 0;
  MEBT_REQUIRE(B > 0 && L > 0 && NC >= 0 && NT > 0 && D > 0 && H > 0 && D == H * 64, MEBT_ERR_SHAPE,
               "stack_forward: bad shape B=%d L=%d NC=%d NT=%d D=%d H=%d (head_dim must be 64)", B, L, NC, NT, D, H);
  MEBT_REQUIRE(workspace != nullptr && workspace_bytes >= stack_workspace_bytes(B, L, NC, NT, D, n_hoist), MEBT_ERR_WORKSPACE,
               "stack_forward: workspace too small (%zu < %zu)", workspace_bytes,
               stack_workspace_bytes(B, L, NC, NT, D, n_hoist));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Workspace ws{static_cast<char*>(workspace), 0, workspace_bytes};
  const int maxq = L > NC + NT ? L : NC + NT;
  const size_t qr = size_t(B) * maxq, kr = qr;
  void* qn = ws.take(qr * D * 2);
  void* kn = ws.take(kr * D * 2);
  void* qkv = ws.take(qr * 3 * D * 2);
  void* kv = ws.take(kr * 2 * D * 2);
  void* att = ws.take(qr * D * 2);
  void* x = ws.take(qr * D * 2);
  void* h = ws.take(qr * D * 2);
  void* u = ws.take(qr * 4 * D * 2);
  void* cat = ws.take(qr * D * 2);
  void* kv_all = n_hoist > 0 ? ws.take(size_t(B) * NC * size_t(n_hoist) * 2 * D * 2) : nullptr;
  const size_t attn_ws_bytes = attention_split_bytes(B, L, NC, NT, D);
  void* attn_ws = attn_ws_bytes > 0 ? ws.take(attn_ws_bytes) : nullptr;
  unsigned long long* packed = static_cast<unsigned long long*>(ws.take(size_t(B) * NT * 8));
  MEBT_REQUIRE(cat != nullptr && (n_hoist == 0 || kv_all != nullptr), MEBT_ERR_WORKSPACE,
               "stack_forward: workspace exhausted");

  // blocks after the last one that writes `targets` cannot influence the logits (gpt.py:247)
  int last = -1;
  for (int i = 0; i < n_layers; ++i)
    if (layers[i].mode == MEBT_MODE_LATENT_DEC || layers[i].mode == MEBT_MODE_MASKGIT) last = i;

  auto LN = [&](const void* in, const float* g, const float* b, void* out, int rows) {
    return layernorm(in, D, MEBT_DTYPE_BF16, g, b, out, D, MEBT_DTYPE_BF16, rows, D, 1e-5f, nullptr, nullptr, st);
  };
  auto GEMM = [&](const void* A, const void* W, int ldw, void* C, int ldc, int M, int N, int K, const float* bias,
                  const void* res, int flags) {
    return gemm_bf16(A, K, 0, W, ldw, 0, C, ldc, M, N, K, bias, res, N, flags, st);
  };
  int rc;
#define TRY(expr) do { rc = (expr); if (rc != MEBT_OK) return rc; } while (0)

  // Contexts never change through the stack (gpt.py:186-195) and ln1's statistics do not depend on the block, so the
  // K|V projections of ALL latent_enc blocks are one GEMM over the normalised contexts: xhat (no affine) times the
  // per-block weights with ln1's gamma folded in (bias absorbs W.beta).  Replaces n_enc LayerNorm passes over
  // [B*NC, D] and n_enc [B*NC,2D,D] GEMMs by one pass and one [B*NC, n_enc*2D, D] GEMM.
  int enc_seen = 0;
  const int ld_all = n_hoist * 2 * D;
  if (n_hoist > 0) {
    TRY(LN(ctx, hoist->ones, hoist->zeros, kn, B * NC));
    TRY(gemm_bf16(kn, D, 0, hoist->w_enc_kv, D, 0, kv_all, ld_all, B * NC, ld_all, D, hoist->b_enc_kv, nullptr, 0, 0, st));
  }

  for (int i = 0; i <= last; ++i) {
    const mebt_layer_t& w = layers[i];
    const __nv_bfloat16* wqkv = static_cast<const __nv_bfloat16*>(w.w_qkv);
    const __nv_bfloat16* w_kv = wqkv + size_t(D) * D;       // rows D..3D = (key, value)
    const float* b_kv = w.b_qkv + D;
    void* q_stream;      // the stream this block rewrites
    int nq;
    const void *Qb, *KV1 = nullptr, *KV2 = nullptr;
    int ldq, ld1 = 0, ld2 = 0, k1c = 0, v1c = 0, nk1 = 0, nk2 = 0;
    switch (w.mode) {
      case MEBT_MODE_LATENT_ENC:
        q_stream = lat; nq = L;
        TRY(LN(lat, w.ln1_w, w.ln1_b, qn, B * L));
        TRY(GEMM(qn, wqkv, D, qkv, D, B * L, D, D, w.b_qkv, nullptr, 0));
        Qb = qkv; ldq = D;
        if (n_hoist > 0) {
          MEBT_REQUIRE(enc_seen < n_hoist, MEBT_ERR_SHAPE, "stack_forward: more latent_enc blocks than hoisted weights");
          KV1 = kv_all; ld1 = ld_all; k1c = enc_seen * 2 * D; v1c = k1c + D; nk1 = NC;
          ++enc_seen;
        } else if (NC > 0) {
          TRY(LN(ctx, w.ln1_w, w.ln1_b, kn, B * NC));
          TRY(GEMM(kn, w_kv, D, kv, 2 * D, B * NC, 2 * D, D, b_kv, nullptr, 0));
          KV1 = kv; ld1 = 2 * D; k1c = 0; v1c = D; nk1 = NC;
        }
        break;
      case MEBT_MODE_LATENT_SELF:
        q_stream = lat; nq = L;
        TRY(LN(lat, w.ln1_w, w.ln1_b, qn, B * L));
        TRY(GEMM(qn, wqkv, D, qkv, 3 * D, B * L, 3 * D, D, w.b_qkv, nullptr, 0));
        Qb = qkv; ldq = 3 * D; KV1 = qkv; ld1 = 3 * D; k1c = D; v1c = 2 * D; nk1 = L;
        break;
      case MEBT_MODE_LATENT_DEC:
        q_stream = tgt; nq = NT;
        TRY(LN(tgt, w.ln1_w, w.ln1_b, qn, B * NT));
        TRY(GEMM(qn, wqkv, D, qkv, D, B * NT, D, D, w.b_qkv, nullptr, 0));
        Qb = qkv; ldq = D;
        TRY(LN(lat, w.ln1_w, w.ln1_b, kn, B * L));
        TRY(GEMM(kn, w_kv, D, kv, 2 * D, B * L, 2 * D, D, b_kv, nullptr, 0));
        KV1 = kv; ld1 = 2 * D; k1c = 0; v1c = D; nk1 = L;
        break;
      case MEBT_MODE_LT2L:
        q_stream = lat; nq = L;
        TRY(LN(lat, w.ln1_w, w.ln1_b, qn, B * L));
        TRY(GEMM(qn, wqkv, D, qkv, 3 * D, B * L, 3 * D, D, w.b_qkv, nullptr, 0));
        Qb = qkv; ldq = 3 * D; KV1 = qkv; ld1 = 3 * D; k1c = D; v1c = 2 * D; nk1 = L;
        TRY(LN(tgt, w.ln1_w, w.ln1_b, kn, B * NT));
        TRY(GEMM(kn, w_kv, D, kv, 2 * D, B * NT, 2 * D, D, b_kv, nullptr, 0));
        KV2 = kv; ld2 = 2 * D; nk2 = NT;
        break;
      case MEBT_MODE_MASKGIT: {
        // q = k = cat[contexts, targets] per batch element (gpt.py:176-178)
        const int n = NC + NT;
        for (int b = 0; b < B; ++b) {
          char* dst = static_cast<char*>(cat) + size_t(b) * n * D * 2;
          if (NC > 0)
            MEBT_CUDA_OK(cudaMemcpyAsync(dst, static_cast<char*>(ctx) + size_t(b) * NC * D * 2, size_t(NC) * D * 2,
                                         cudaMemcpyDeviceToDevice, st));
          MEBT_CUDA_OK(cudaMemcpyAsync(dst + size_t(NC) * D * 2, static_cast<char*>(tgt) + size_t(b) * NT * D * 2,
                                       size_t(NT) * D * 2, cudaMemcpyDeviceToDevice, st));
        }
        q_stream = cat; nq = n;
        TRY(LN(cat, w.ln1_w, w.ln1_b, qn, B * n));
        TRY(GEMM(qn, wqkv, D, qkv, 3 * D, B * n, 3 * D, D, w.b_qkv, nullptr, 0));
        Qb = qkv; ldq = 3 * D; KV1 = qkv; ld1 = 3 * D; k1c = D; v1c = 2 * D; nk1 = n;
        break;
      }
      default:
        MEBT_REQUIRE(false, MEBT_ERR_UNSUPPORTED, "stack_forward: unknown block mode %d", w.mode);
    }
    TRY(latent_attention_fwd(Qb, ldq, 0, KV1, ld1, k1c, v1c, nk1, KV2, ld2, 0, D, nk2, att, D, nullptr, B, H, nq, 64, 0.f, 0ull,
                             attn_ws, attn_ws_bytes, stream));
    const int rows = B * nq;
    TRY(GEMM(att, w.w_proj, D, x, D, rows, D, D, w.b_proj, qn, 0));                    // x = ln1(q) + proj(att)
    TRY(LN(x, w.ln2_w, w.ln2_b, h, rows));
    TRY(GEMM(h, w.w_fc1, D, u, 4 * D, rows, 4 * D, D, w.b_fc1, nullptr, MEBT_GEMM_GELU));
    TRY(GEMM(u, w.w_fc2, 4 * D, q_stream, D, rows, D, 4 * D, w.b_fc2, x, 0));          // stream = x + mlp(ln2(x))
    if (w.mode == MEBT_MODE_MASKGIT) {
      const int n = NC + NT;
      for (int b = 0; b < B; ++b) {
        const char* src = static_cast<const char*>(cat) + size_t(b) * n * D * 2;
        if (NC > 0)
          MEBT_CUDA_OK(cudaMemcpyAsync(static_cast<char*>(ctx) + size_t(b) * NC * D * 2, src, size_t(NC) * D * 2,
                                       cudaMemcpyDeviceToDevice, st));
        MEBT_CUDA_OK(cudaMemcpyAsync(static_cast<char*>(tgt) + size_t(b) * NT * D * 2, src + size_t(NC) * D * 2,
                                     size_t(NT) * D * 2, cudaMemcpyDeviceToDevice, st));
      }
    }
  }
  if (logits != nullptr || sample_ids != nullptr) TRY(LN(tgt, lnf_w, lnf_b, h, B * NT));
  if (logits != nullptr)
    TRY(gemm_bf16(h, D, 0, w_head, D, 0, logits, V, B * NT, V, D, nullptr, nullptr, 0,
                  logits_dtype == MEBT_DTYPE_FP32 ? MEBT_GEMM_OUT_FP32 : 0, st));
  if (sample_ids != nullptr) {
    // head GEMM whose epilogue draws one token per row (Gumbel-max over the 16384 logits): no logits in HBM
    const long long rows = (long long)B * NT;
    packed_init_kernel<<<int((rows + 255) / 256), 256, 0, st>>>(packed, rows);
    MEBT_LAUNCH_OK("packed_init_kernel");
    TRY(gemm_bf16_sample(h, D, w_head, D, int(rows), V, D, temperature, seed, offset, packed, st));
    packed_ids_kernel<<<int((rows + 255) / 256), 256, 0, st>>>(packed, sample_ids, rows);
    MEBT_LAUNCH_OK("packed_ids_kernel");
  }
#undef TRY
  return MEBT_OK;
}


size_t mebt_head_sample_workspace_bytes(long long rows) { return size_t(rows) * 8; }

int mebt_head_sample(const void* x, int ldx, const void* w_head, int ldw, int rows, int V, int D, float temperature,
                     unsigned long long seed, unsigned long long offset, int64_t* ids, void* workspace,
                     size_t workspace_bytes, void* stream) {
  using namespace mebt;
  MEBT_REQUIRE(rows >= 0 && V > 0 && D > 0, MEBT_ERR_SHAPE, "head_sample: bad shape");
  if (rows == 0) return MEBT_OK;
  MEBT_REQUIRE(workspace != nullptr && workspace_bytes >= size_t(rows) * 8, MEBT_ERR_WORKSPACE, "head_sample: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  unsigned long long* packed = static_cast<unsigned long long*>(workspace);
  packed_init_kernel<<<(rows + 255) / 256, 256, 0, st>>>(packed, rows);
  MEBT_LAUNCH_OK("packed_init_kernel");
  int rc = gemm_bf16_sample(x, ldx, w_head, ldw, rows, V, D, temperature, seed, offset, packed, st);
  if (rc) return rc;
  packed_ids_kernel<<<(rows + 255) / 256, 256, 0, st>>>(packed, ids, rows);
  MEBT_LAUNCH_OK("packed_ids_kernel");
  return MEBT_OK;
}

}  // extern "C"

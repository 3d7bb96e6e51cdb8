// Encoder tower forward / backward: the launch sequence over the kernels of this directory, plus the buffer layouts
// of the C ABI (include/simx_b200.h).  Replaces HFBertEncoder.forward / loss.backward() through one tower
// (SimANS/model/models.py:77-82, co_training_marco_train.py:222; math: LEAD/modeling_bert.py:201-660,
// LEAD/modeling_distilbert.py:93-376 — the same post-LN block in both architectures).
//
// per layer, forward (x = layer input, all activations [T, .] as 1 or 3 bf16 planes):
//   qkv = x Wqkv^T + b            tcgen05 GEMM, BIAS epilogue            [T,3H]
//   ctx = softmax(q k^T/8 + m) v  attention kernel, saves lse            [T,H]
//   z1  = ctx Wo^T + b + x        GEMM, BIAS_ADD epilogue                [T,H]
//   x1  = LN(z1)                  row kernel                             [T,H]
//   g'  = gelu'(x1 W1^T + b) ; a = gelu(.)  GEMM, BIAS_GELU epilogue (two outputs; g' is all backward needs) [T,I] x2
//   z2  = a W2^T + b + x1         GEMM, BIAS_ADD epilogue                [T,H]
//   out = LN(z2)                                                         [T,H]
// backward mirrors it: LN-bwd (also yields the bias gradient of the GEMM before it), dgrad GEMMs against the
// pre-transposed weights with the gelu' / residual-add fused into their epilogues, split-K MN-major wgrad GEMMs
// accumulating with fp32 atomics straight into the flat gradient buffer.
#include "common.cuh"

namespace simx {

// kernels from the other translation units
int embed_ln(int np, const void* ids, const float* embeds, const void* pos_ids, int id_bytes, int max_pos, int T, int S, int H,
             int vocab, const float* word, const float* pos, const float* type0, const float* gamma, const float* beta,
             float eps, bf16* z0, bf16* x, long ps, const DropSpec& drop, cudaStream_t st);
int ln_fwd(int np, const bf16* z, int T, int H, const float* gamma, const float* beta, float eps, bf16* x, long ps,
           cudaStream_t st);
int ln_bwd(int np, const bf16* dy, const bf16* z, int T, int H, const float* gamma, float eps, bf16* dz, long ps,
           float* dgamma, float* dbeta, float* dbias, bf16* dzm, int drop_mode, const DropSpec& drop, cudaStream_t st);
int colsum(int np, const bf16* x, long rows, int N, long row_stride, long ps, float* out, cudaStream_t st);
int cls_gather(int np, const bf16* x, int B, int S, int H, long ps, float* out, long out_stride_b, cudaStream_t st);
int cls_scatter_add(int np, bf16* dx, int B, int S, int H, long ps, const float* d, long d_stride_b, cudaStream_t st);
int f32_to_planes(int np, const float* src, long n, bf16* dst, long ps, cudaStream_t st);
int planes_to_f32(int np, const bf16* src, long n, long ps, float* dst, cudaStream_t st);
int planes_to_f32_strided(int np, const bf16* src, int B, long SH, long ps, float* dst, long dst_stride_b, cudaStream_t st);
int planes_add_f32(int np, bf16* dx, int B, long SH, long ps, const float* d, long d_stride_b, cudaStream_t st);
int embed_scatter(int np, const void* ids, int id_bytes, const bf16* dz0, int T, int H, int vocab, int pad_id, long ps,
                  float* dword, cudaStream_t st);
int convert_weight(int np, const float* w, int R, int C, bf16* dst, long ld_dst, long ps_dst, bf16* dst_t, long ld_t,
                   long ps_t, cudaStream_t st);
// batched form: every job queued between begin and end goes out in one launch per 56 jobs (C == 0: fp32 copy of R values)
int conv_begin(void** handle);
int conv_add(void* handle, int np, const float* w, int R, int C, bf16* dst, long ld_dst, long ps_dst, bf16* dst_t, long ld_t,
             long ps_t, cudaStream_t st);
int conv_end(void* handle, int np, cudaStream_t st);
int attention_forward(int np, const bf16* qkv, long ps_qkv, bf16* ctx, long ps_ctx, const void* mask, int mask_w,
                      float* lse2, int B, int S, int H, int heads, const DropSpec& drop, cudaStream_t st);
int attention_backward(int np, const bf16* qkv, long ps_qkv, const bf16* ctx, const bf16* dctx, long ps_ctx, bf16* dqkv,
                       const void* mask, int mask_w, const float* lse2, float* delta, int B, int S, int H, int heads,
                       float* dbias, const DropSpec& drop, cudaStream_t st);

static inline size_t al(size_t x) { return (x + 255) & ~size_t(255); }

static int check_cfg(const simx_encoder_cfg* c) {
  SIMX_CHECK(c != nullptr, SIMX_ERR_INVALID, "null encoder cfg");
  SIMX_CHECK(c->arch == SIMX_ARCH_BERT || c->arch == SIMX_ARCH_DISTILBERT, SIMX_ERR_INVALID, "unknown arch %d", c->arch);
  SIMX_CHECK(c->num_layers >= 1 && c->num_layers <= 64, SIMX_ERR_INVALID, "num_layers=%d out of range", c->num_layers);
  SIMX_CHECK(c->hidden >= 64 && c->hidden % 64 == 0 && c->hidden <= 1024, SIMX_ERR_INVALID,
             "hidden=%d must be a multiple of 64 in [64,1024]", c->hidden);
  SIMX_CHECK(c->heads * 64 == c->hidden, SIMX_ERR_INVALID, "head_dim must be 64 (hidden=%d heads=%d)", c->hidden, c->heads);
  SIMX_CHECK(c->intermediate >= 64 && c->intermediate % 64 == 0, SIMX_ERR_INVALID, "intermediate=%d must be a multiple of 64",
             c->intermediate);
  SIMX_CHECK(c->vocab > 0 && c->max_pos > 0, SIMX_ERR_INVALID, "vocab / max_pos must be positive");
  SIMX_CHECK(c->arch != SIMX_ARCH_BERT || c->type_vocab >= 1, SIMX_ERR_INVALID, "BERT needs type_vocab >= 1");
  SIMX_CHECK(c->precision == SIMX_PREC_BF16 || c->precision == SIMX_PREC_BF16X3, SIMX_ERR_INVALID, "unknown precision %d",
             c->precision);
  SIMX_CHECK(c->hidden_dropout >= 0.f && c->hidden_dropout < 1.f && c->attn_dropout >= 0.f && c->attn_dropout < 1.f,
             SIMX_ERR_INVALID, "dropout probabilities must lie in [0, 1): hidden %g, attention %g", c->hidden_dropout,
             c->attn_dropout);
  SIMX_CHECK(c->save_policy == SIMX_SAVE_ALL || c->save_policy == SIMX_SAVE_LAYER_INPUTS, SIMX_ERR_INVALID,
             "unknown save_policy %d", c->save_policy);
  SIMX_CHECK(c->id_bytes == 0 || c->id_bytes == 4 || c->id_bytes == 8, SIMX_ERR_INVALID, "id_bytes must be 4 or 8 (0 = 8), got %d",
             c->id_bytes);
  return SIMX_OK;
}
static inline int id_bytes(const simx_encoder_cfg* c) { return c->id_bytes == 4 ? 4 : 8; }
static inline bool ckpt(const simx_encoder_cfg* c) { return c->save_policy == SIMX_SAVE_LAYER_INPUTS; }
// dropout masks of this call: sites as in LEAD/modeling_bert.py (:239, :358, :386, :464); DistilBERT has no dropout after
// the attention output projection (LEAD/modeling_distilbert.py:260-319)
static inline DropSpec drop_at(const simx_encoder_cfg* c, uint32_t site, int layer) {
  float p = site == DROP_SITE_ATTN ? c->attn_dropout : c->hidden_dropout;
  if (site == DROP_SITE_ATTN_OUT && c->arch == SIMX_ARCH_DISTILBERT) p = 0.f;
  return p > 0.f ? make_drop(p, c->dropout_seed, site, layer) : no_drop();
}

struct ParamTable {
  int base;      // index of layer 0's first parameter
  int n;         // number of parameters
  bool bert;
};
static ParamTable ptable(const simx_encoder_cfg* c) {
  ParamTable t;
  t.bert = c->arch == SIMX_ARCH_BERT;
  t.base = t.bert ? 5 : 4;
  t.n = t.base + 16 * c->num_layers + (t.bert ? 2 : 0);
  return t;
}
static int64_t param_numel(const simx_encoder_cfg* c, int idx) {
  const ParamTable t = ptable(c);
  const int64_t H = c->hidden, I = c->intermediate;
  if (idx < 0 || idx >= t.n) return -1;
  if (idx == 0) return (int64_t)c->vocab * H;
  if (idx == 1) return (int64_t)c->max_pos * H;
  if (t.bert && idx == 2) return (int64_t)c->type_vocab * H;
  if (idx < t.base) return H;
  if (idx >= t.base + 16 * c->num_layers) return idx == t.base + 16 * c->num_layers ? H * H : H;
  const int k = (idx - t.base) % 16;
  switch (k) {
    case 0: case 2: case 4: case 6: return H * H;
    case 10: case 12: return I * H;
    case 11: return I;
    default: return H;
  }
}
// flat gradient layout: [word, pos, (type), embLN.w, embLN.b] then per layer
// [Wqkv(3H*H) bqkv(3H) Wo bo ln1w ln1b W1 b1 W2 b2 ln2w ln2b] then (pooler.w, pooler.b)
static int64_t layer_grad_numel(const simx_encoder_cfg* c) {
  const int64_t H = c->hidden, I = c->intermediate;
  return 3 * H * H + 3 * H + H * H + H + 2 * H + I * H + I + H * I + H + 2 * H;
}
static int64_t emb_grad_numel(const simx_encoder_cfg* c) {
  const int64_t H = c->hidden;
  return ((int64_t)c->vocab + c->max_pos + (c->arch == SIMX_ARCH_BERT ? c->type_vocab : 0)) * H + 2 * H;
}
static int64_t grad_offset(const simx_encoder_cfg* c, int idx) {
  const ParamTable t = ptable(c);
  const int64_t H = c->hidden, I = c->intermediate;
  if (idx < 0 || idx >= t.n) return -1;
  if (idx < t.base) {
    int64_t off = 0;
    for (int i = 0; i < idx; ++i) off += param_numel(c, i);
    return off;
  }
  const int64_t e = emb_grad_numel(c);
  if (idx >= t.base + 16 * c->num_layers)
    return e + layer_grad_numel(c) * c->num_layers + (idx == t.base + 16 * c->num_layers ? 0 : H * H);
  const int l = (idx - t.base) / 16, k = (idx - t.base) % 16;
  const int64_t lo = e + layer_grad_numel(c) * l;
  const int64_t o_qkvb = 3 * H * H, o_wo = o_qkvb + 3 * H, o_bo = o_wo + H * H, o_ln1 = o_bo + H, o_w1 = o_ln1 + 2 * H,
                o_b1 = o_w1 + I * H, o_w2 = o_b1 + I, o_b2 = o_w2 + H * I, o_ln2 = o_b2 + H;
  switch (k) {
    case 0: return lo;
    case 2: return lo + H * H;
    case 4: return lo + 2 * H * H;
    case 1: return lo + o_qkvb;
    case 3: return lo + o_qkvb + H;
    case 5: return lo + o_qkvb + 2 * H;
    case 6: return lo + o_wo;
    case 7: return lo + o_bo;
    case 8: return lo + o_ln1;
    case 9: return lo + o_ln1 + H;
    case 10: return lo + o_w1;
    case 11: return lo + o_b1;
    case 12: return lo + o_w2;
    case 13: return lo + o_b2;
    case 14: return lo + o_ln2;
    default: return lo + o_ln2 + H;
  }
}

// ---- weight buffer ------------------------------------------------------------------------------
struct LayerW {
  size_t wqkv, wqkvT, wo, woT, w1, w1T, w2, w2T, bqkv, total;
};
static LayerW layer_w(const simx_encoder_cfg* c) {
  const size_t H = c->hidden, I = c->intermediate, np = c->precision;
  LayerW w;
  size_t off = 0;
  w.wqkv = off;  off += al(np * 3 * H * H * 2);
  w.wqkvT = off; off += al(np * 3 * H * H * 2);
  w.wo = off;    off += al(np * H * H * 2);
  w.woT = off;   off += al(np * H * H * 2);
  w.w1 = off;    off += al(np * I * H * 2);
  w.w1T = off;   off += al(np * I * H * 2);
  w.w2 = off;    off += al(np * I * H * 2);
  w.w2T = off;   off += al(np * I * H * 2);
  w.bqkv = off;  off += al(3 * H * 4);
  w.total = off;
  return w;
}

// ---- saved activations --------------------------------------------------------------------------
struct SavedL {
  size_t z0, x, layer0, l_qkv, l_ctx, l_z1, l_x1, l_h, l_a, l_z2, l_lse, l_total, total;
  size_t x_stride;
};
static SavedL saved_layout(const simx_encoder_cfg* c, int B, int S) {
  const size_t T = (size_t)B * S, H = c->hidden, I = c->intermediate, np = c->precision, L = c->num_layers;
  SavedL s;
  size_t off = 0;
  s.z0 = off; off += al(np * T * H * 2);
  s.x_stride = al(np * T * H * 2);
  s.x = off;  off += s.x_stride * (L + 1);
  size_t lo = 0;
  s.l_qkv = lo; lo += al(np * T * 3 * H * 2);
  s.l_ctx = lo; lo += al(np * T * H * 2);
  s.l_z1 = lo;  lo += al(np * T * H * 2);
  s.l_x1 = lo;  lo += al(np * T * H * 2);
  s.l_h = lo;   lo += al(np * T * I * 2);
  s.l_a = lo;   lo += al(np * T * I * 2);
  s.l_z2 = lo;  lo += al(np * T * H * 2);
  s.l_lse = lo; lo += al((size_t)B * c->heads * S * 4);
  s.l_total = lo;
  s.layer0 = off;
  // SIMX_SAVE_LAYER_INPUTS (the reference's --gradient_checkpointing): only z0 and the layer inputs x[0..L] are kept;
  // backward recomputes one layer at a time into the workspace
  if (!ckpt(c)) off += lo * L;
  s.total = off + 256;
  return s;
}

// ---- workspace: backward temporaries, and one layer of activations for saved == NULL (inference) ----
struct WsL {
  size_t dx, dza, dxb, dzm, dh, dqkv, delta, scratch, infer, total;
};
static WsL ws_layout(const simx_encoder_cfg* c, int B, int S) {
  const size_t T = (size_t)B * S, H = c->hidden, I = c->intermediate, np = c->precision;
  WsL w;
  size_t off = 0;
  w.dx = off;    off += al(np * T * H * 2);
  w.dza = off;   off += al(np * T * H * 2);
  w.dxb = off;   off += al(np * T * H * 2);
  w.dzm = off;   off += c->hidden_dropout > 0.f ? al(np * T * H * 2) : 0;  // dropout-masked copy of the LayerNorm-backward output
  w.dh = off;    off += al(np * T * I * 2);
  w.dqkv = off;  off += al(np * T * 3 * H * 2);
  w.delta = off; off += al((size_t)B * c->heads * S * 4);
  const SavedL s = saved_layout(c, B, S);
  // layer-input checkpointing: one layer of activations, written by the training forward and re-written by backward
  w.scratch = off;
  if (ckpt(c)) off += s.l_total;
  // inference scratch aliases the backward temporaries (never live together): z0 + 2 x + one layer
  w.infer = 0;
  const size_t infer_bytes = al(np * T * H * 2) * 3 + s.l_total;
  w.total = (off > infer_bytes ? off : infer_bytes) + 256;
  return w;
}

static inline uint8_t* align_ptr(const void* p) {
  uintptr_t u = reinterpret_cast<uintptr_t>(p);
  return reinterpret_cast<uint8_t*>((u + 255) & ~uintptr_t(255));
}

static int wgrad_ksplit(int M, int N, int K) {
  const int bn = N <= 64 ? 64 : (N <= 128 ? 128 : 256);
  const int tiles = ((M + 127) / 128) * ((N + bn - 1) / bn);
  const int kb = (K + 63) / 64;
  int ks = (3 * num_sms() + tiles - 1) / tiles;
  ks = min(ks, max(1, kb / 4));  // at least 4 k-blocks per split
  return max(1, ks);
}

}  // namespace simx

using namespace simx;

extern "C" int simx_encoder_num_params(const simx_encoder_cfg* cfg) { return check_cfg(cfg) ? -1 : ptable(cfg).n; }
extern "C" int64_t simx_encoder_param_numel(const simx_encoder_cfg* cfg, int idx) {
  return check_cfg(cfg) ? -1 : param_numel(cfg, idx);
}
extern "C" int64_t simx_encoder_grad_offset(const simx_encoder_cfg* cfg, int idx) {
  return check_cfg(cfg) ? -1 : grad_offset(cfg, idx);
}
extern "C" int64_t simx_encoder_grad_numel(const simx_encoder_cfg* cfg) {
  if (check_cfg(cfg)) return -1;
  const int64_t H = cfg->hidden;
  return emb_grad_numel(cfg) + layer_grad_numel(cfg) * cfg->num_layers + (cfg->arch == SIMX_ARCH_BERT ? H * H + H : 0);
}
extern "C" size_t simx_encoder_weight_bytes(const simx_encoder_cfg* cfg) {
  return check_cfg(cfg) ? 0 : layer_w(cfg).total * cfg->num_layers + 256;
}
extern "C" size_t simx_encoder_saved_bytes(const simx_encoder_cfg* cfg, int B, int S) {
  return (check_cfg(cfg) || B <= 0 || S <= 0) ? 0 : saved_layout(cfg, B, S).total;
}
extern "C" size_t simx_encoder_workspace_bytes(const simx_encoder_cfg* cfg, int B, int S) {
  return (check_cfg(cfg) || B <= 0 || S <= 0) ? 0 : ws_layout(cfg, B, S).total;
}

extern "C" int simx_encoder_prepare_weights(const simx_encoder_cfg* cfg, const float* const* params, void* wbuf,
                                            void* stream) {
  SIMX_TRY(check_cfg(cfg));
  SIMX_CHECK(params && wbuf, SIMX_ERR_INVALID, "prepare_weights: null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int np = cfg->precision, H = cfg->hidden, I = cfg->intermediate;
  const ParamTable t = ptable(cfg);
  const LayerW lw = layer_w(cfg);
  uint8_t* base = align_ptr(wbuf);
  void* batch = nullptr;
  SIMX_TRY(conv_begin(&batch));
  int rc = SIMX_OK;
  for (int l = 0; l < cfg->num_layers && rc == SIMX_OK; ++l) {
    const float* const* P = params + t.base + 16 * l;
    uint8_t* w = base + lw.total * l;
    bf16* wqkv = reinterpret_cast<bf16*>(w + lw.wqkv);
    bf16* wqkvT = reinterpret_cast<bf16*>(w + lw.wqkvT);
    const long ps3 = 3L * H * H;
    for (int j = 0; j < 3 && rc == SIMX_OK; ++j)  // q, k, v rows stacked; transposed copy [H, 3H] has them as column blocks
      rc = conv_add(batch, np, P[2 * j], H, H, wqkv + (long)j * H * H, H, ps3, wqkvT + (long)j * H, 3L * H, ps3, st);
    if (rc == SIMX_OK)
      rc = conv_add(batch, np, P[6], H, H, reinterpret_cast<bf16*>(w + lw.wo), H, (long)H * H,
                    reinterpret_cast<bf16*>(w + lw.woT), H, (long)H * H, st);
    if (rc == SIMX_OK)
      rc = conv_add(batch, np, P[10], I, H, reinterpret_cast<bf16*>(w + lw.w1), H, (long)I * H,
                    reinterpret_cast<bf16*>(w + lw.w1T), I, (long)I * H, st);
    if (rc == SIMX_OK)
      rc = conv_add(batch, np, P[12], H, I, reinterpret_cast<bf16*>(w + lw.w2), I, (long)I * H,
                    reinterpret_cast<bf16*>(w + lw.w2T), H, (long)I * H, st);
    float* bqkv = reinterpret_cast<float*>(w + lw.bqkv);
    for (int j = 0; j < 3 && rc == SIMX_OK; ++j)  // fused [3H] bias: fp32 copies ride in the same launch
      rc = conv_add(batch, np, P[2 * j + 1], H, 0, reinterpret_cast<bf16*>(bqkv + j * H), 0, 0, nullptr, 0, 0, st);
  }
  const int rc_end = conv_end(batch, np, st);
  SIMX_TRY(rc);
  SIMX_TRY(rc_end);
  return SIMX_OK;
}

// one layer's activation buffers (in `saved`, or in the workspace scratch when they are recomputed / not kept)
struct LayerBufs {
  bf16 *qkv, *ctx, *z1, *x1, *h, *a, *z2;
  float* lse;
};
static LayerBufs layer_bufs(uint8_t* base, const SavedL& sl) {
  LayerBufs b;
  b.qkv = reinterpret_cast<bf16*>(base + sl.l_qkv);
  b.ctx = reinterpret_cast<bf16*>(base + sl.l_ctx);
  b.z1 = reinterpret_cast<bf16*>(base + sl.l_z1);
  b.x1 = reinterpret_cast<bf16*>(base + sl.l_x1);
  b.h = reinterpret_cast<bf16*>(base + sl.l_h);
  b.a = reinterpret_cast<bf16*>(base + sl.l_a);
  b.z2 = reinterpret_cast<bf16*>(base + sl.l_z2);
  b.lse = reinterpret_cast<float*>(base + sl.l_lse);
  return b;
}

// forward of layer l: x -> (lb.*) -> xo = LN(z2) (xo == NULL: stop at z2 — the recomputation inside backward).
// train: the FFN-up epilogue also stores gelu' for backward.
static int layer_forward(const simx_encoder_cfg* cfg, const float* const* P, const uint8_t* w, const LayerW& lw, int l,
                         const bf16* x, bf16* xo, const LayerBufs& lb, const void* mask, int B, int S, bool train,
                         cudaStream_t st) {
  const int np = cfg->precision, H = cfg->hidden, I = cfg->intermediate;
  const int T = B * S;
  const long psH = (long)T * H, ps3H = (long)T * 3 * H, psI = (long)T * I;
  const bf16* wqkv = reinterpret_cast<const bf16*>(w + lw.wqkv);
  const bf16* wo = reinterpret_cast<const bf16*>(w + lw.wo);
  const bf16* w1 = reinterpret_cast<const bf16*>(w + lw.w1);
  const bf16* w2 = reinterpret_cast<const bf16*>(w + lw.w2);
  const float* bqkv = reinterpret_cast<const float*>(w + lw.bqkv);
  // dropout follows the cfg probabilities (the caller zeroes them in eval mode), whether or not activations are kept:
  // the reference's model.train() + torch.no_grad() still drops
  const DropSpec d_attn = drop_at(cfg, DROP_SITE_ATTN, l);
  const DropSpec d_ao = drop_at(cfg, DROP_SITE_ATTN_OUT, l);
  const DropSpec d_fo = drop_at(cfg, DROP_SITE_FFN_OUT, l);

  prof_label("qkv_fwd");
  SIMX_TRY(gemm(T, 3 * H, H, x, H, psH, false, wqkv, H, 3L * H * H, false, np, SIMX_EPI_BIAS, bqkv, nullptr, 0, 0, lb.qkv,
                nullptr, 3 * H, ps3H, nullptr, 0, 1, st));
  prof_label(nullptr);
  SIMX_TRY(attention_forward(np, lb.qkv, ps3H, lb.ctx, psH, mask, id_bytes(cfg), lb.lse, B, S, H, cfg->heads, d_attn, st));
  prof_label("attnout_fwd");
  SIMX_TRY(gemm(T, H, H, lb.ctx, H, psH, false, wo, H, (long)H * H, false, np, SIMX_EPI_BIAS_ADD, P[7], x, H, psH, lb.z1,
                nullptr, H, psH, nullptr, 0, 1, st, &d_ao));
  prof_label(nullptr);
  SIMX_TRY(ln_fwd(np, lb.z1, T, H, P[8], P[9], cfg->ln_eps, lb.x1, psH, st));
  prof_label("ffn1_fwd");
  if (train)
    SIMX_TRY(gemm(T, I, H, lb.x1, H, psH, false, w1, H, (long)I * H, false, np, SIMX_EPI_BIAS_GELU, P[11], nullptr, 0, 0, lb.h,
                  lb.a, I, psI, nullptr, 0, 1, st));
  else  // inference: gelu only, one [T, I] output instead of two
    SIMX_TRY(gemm(T, I, H, lb.x1, H, psH, false, w1, H, (long)I * H, false, np, SIMX_EPI_BIAS_GELU_FWD, P[11], nullptr, 0, 0,
                  lb.a, nullptr, I, psI, nullptr, 0, 1, st));
  prof_label("ffn2_fwd");
  SIMX_TRY(gemm(T, H, I, lb.a, I, psI, false, w2, I, (long)I * H, false, np, SIMX_EPI_BIAS_ADD, P[13], lb.x1, H, psH, lb.z2,
                nullptr, H, psH, nullptr, 0, 1, st, &d_fo));
  prof_label(nullptr);
  if (xo != nullptr) SIMX_TRY(ln_fwd(np, lb.z2, T, H, P[14], P[15], cfg->ln_eps, xo, psH, st));
  return SIMX_OK;
}

static int encoder_forward_impl(const simx_encoder_cfg* cfg, const float* const* params, const void* wbuf,
                                const void* ids, const float* embeds, const void* mask, const void* pos_ids, int B, int S,
                                void* saved, void* workspace, float* cls_out, float* seq_out, void* stream) {
  SIMX_TRY(check_cfg(cfg));
  SIMX_CHECK(params && wbuf && (ids || embeds), SIMX_ERR_INVALID, "encoder_forward: null argument");
  SIMX_CHECK(B > 0 && S > 0 && S <= cfg->max_pos && S <= 512, SIMX_ERR_INVALID,
             "encoder_forward: need 0 < S <= min(max_pos=%d, 512), got B=%d S=%d", cfg->max_pos, B, S);
  SIMX_CHECK(saved != nullptr || workspace != nullptr, SIMX_ERR_INVALID, "encoder_forward: need saved or workspace");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int np = cfg->precision, H = cfg->hidden, L = cfg->num_layers;
  const int T = B * S;
  const long psH = (long)T * H;
  const ParamTable t = ptable(cfg);
  const LayerW lw = layer_w(cfg);
  const SavedL sl = saved_layout(cfg, B, S);
  const WsL wl = ws_layout(cfg, B, S);
  const uint8_t* wb = align_ptr(wbuf);
  const bool keep = saved != nullptr;   // a backward pass will follow: activations (or layer inputs) are kept
  SIMX_CHECK(!(keep && ckpt(cfg)) || workspace != nullptr, SIMX_ERR_INVALID,
             "encoder_forward: SIMX_SAVE_LAYER_INPUTS needs the workspace for the layer scratch");
  uint8_t* sb = keep ? align_ptr(saved) : align_ptr(workspace);
  // inference: [z0][xa][xb][one layer] inside the workspace
  auto xbuf = [&](int l) -> bf16* {
    if (keep) return reinterpret_cast<bf16*>(sb + sl.x + sl.x_stride * l);
    return reinterpret_cast<bf16*>(sb + sl.x_stride * (1 + (l & 1)));
  };
  auto lbase = [&](int l) -> uint8_t* {
    if (keep && ckpt(cfg)) return align_ptr(workspace) + wl.scratch;
    if (keep) return sb + sl.layer0 + sl.l_total * l;
    return sb + sl.x_stride * 3;
  };
  bf16* z0 = keep ? reinterpret_cast<bf16*>(sb + sl.z0) : nullptr;
  const float* type0 = t.bert ? params[2] : nullptr;
  const float* eg = params[t.bert ? 3 : 2];
  const float* eb = params[t.bert ? 4 : 3];
  SIMX_TRY(embed_ln(np, ids, embeds, pos_ids, id_bytes(cfg), cfg->max_pos, T, S, H, cfg->vocab, params[0], params[1], type0, eg,
                    eb, cfg->ln_eps, z0, xbuf(0), psH, drop_at(cfg, DROP_SITE_EMBED, 0), st));
  const long cls_stride_b = (long)(L + 1) * H;
  if (cls_out) SIMX_TRY(cls_gather(np, xbuf(0), B, S, H, psH, cls_out, cls_stride_b, st));

  for (int l = 0; l < L; ++l) {
    const LayerBufs lb = layer_bufs(lbase(l), sl);
    SIMX_TRY(layer_forward(cfg, params + t.base + 16 * l, wb + lw.total * l, lw, l, xbuf(l), xbuf(l + 1), lb, mask, B, S, keep,
                           st));
    if (cls_out) SIMX_TRY(cls_gather(np, xbuf(l + 1), B, S, H, psH, cls_out + (long)(l + 1) * H, cls_stride_b, st));
  }
  if (seq_out) SIMX_TRY(planes_to_f32(np, xbuf(L), (long)T * H, psH, seq_out, st));
  return SIMX_OK;
}

extern "C" int simx_encoder_forward(const simx_encoder_cfg* cfg, const float* const* params, const void* wbuf,
                                    const void* ids, const void* mask, int B, int S, void* saved, void* workspace,
                                    float* cls_out, float* seq_out, void* stream) {
  return encoder_forward_impl(cfg, params, wbuf, ids, nullptr, mask, nullptr, B, S, saved, workspace, cls_out, seq_out, stream);
}
extern "C" int simx_encoder_forward_pos(const simx_encoder_cfg* cfg, const float* const* params, const void* wbuf,
                                        const void* ids, const void* mask, const void* pos_ids, int B, int S,
                                        void* saved, void* workspace, float* cls_out, float* seq_out, void* stream) {
  SIMX_CHECK(pos_ids != nullptr, SIMX_ERR_INVALID, "encoder_forward_pos: null position ids");
  return encoder_forward_impl(cfg, params, wbuf, ids, nullptr, mask, pos_ids, B, S, saved, workspace, cls_out, seq_out, stream);
}
extern "C" int simx_encoder_forward_embeds(const simx_encoder_cfg* cfg, const float* const* params, const void* wbuf,
                                           const float* inputs_embeds, const void* mask, const void* pos_ids, int B, int S,
                                           void* saved, void* workspace, float* cls_out, float* seq_out, void* stream) {
  SIMX_CHECK(inputs_embeds != nullptr, SIMX_ERR_INVALID, "encoder_forward_embeds: null inputs_embeds");
  return encoder_forward_impl(cfg, params, wbuf, nullptr, inputs_embeds, mask, pos_ids, B, S, saved, workspace, cls_out, seq_out,
                              stream);
}

// Backward through layers [layer_lo, layer_hi) in descending order.  layer_hi == L starts the pass (the running cotangent
// dx in the workspace is initialised from d_seq / d_cls[:, L]); layer_lo == -1 finishes it with the embedding backward.
// Between calls the cotangent lives in the workspace, so a caller that splits the pass (one call per layer: gradients of
// layer l are complete — and can be handed to DDP's reducer — while layer l-1 is still being computed) must pass the
// same workspace to every call.
static int encoder_backward_impl(const simx_encoder_cfg* cfg, const float* const* params, const void* wbuf,
                                 const void* ids, const void* mask, const void* pos_ids, int pad_id, int B, int S,
                                 const void* saved, void* workspace, const float* d_cls, const float* d_seq, float* grads,
                                 int layer_lo, int layer_hi, void* stream, float* d_embeds = nullptr) {
  SIMX_TRY(check_cfg(cfg));
  SIMX_CHECK(params && wbuf && (ids || d_embeds) && saved && workspace && grads, SIMX_ERR_INVALID,
             "encoder_backward: null argument");
  SIMX_CHECK(B > 0 && S > 0 && S <= cfg->max_pos && S <= 512, SIMX_ERR_INVALID, "encoder_backward: bad B=%d S=%d", B, S);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int np = cfg->precision, H = cfg->hidden, I = cfg->intermediate, L = cfg->num_layers;
  SIMX_CHECK(layer_lo >= -1 && layer_lo < layer_hi && layer_hi <= L, SIMX_ERR_INVALID,
             "encoder_backward: bad layer range [%d, %d) for %d layers", layer_lo, layer_hi, L);
  SIMX_CHECK(layer_hi < L || d_cls != nullptr || d_seq != nullptr, SIMX_ERR_INVALID, "encoder_backward: no cotangent given");
  const int T = B * S;
  const long psH = (long)T * H, ps3H = (long)T * 3 * H, psI = (long)T * I;
  const ParamTable t = ptable(cfg);
  const LayerW lw = layer_w(cfg);
  const SavedL sl = saved_layout(cfg, B, S);
  const WsL wl = ws_layout(cfg, B, S);
  const uint8_t* wb = align_ptr(wbuf);
  uint8_t* sb = align_ptr(saved);
  uint8_t* ws = align_ptr(workspace);
  bf16* dx = reinterpret_cast<bf16*>(ws + wl.dx);
  bf16* dza = reinterpret_cast<bf16*>(ws + wl.dza);
  bf16* dxb = reinterpret_cast<bf16*>(ws + wl.dxb);
  bf16* dzm = reinterpret_cast<bf16*>(ws + wl.dzm);
  bf16* dh = reinterpret_cast<bf16*>(ws + wl.dh);
  bf16* dqkv = reinterpret_cast<bf16*>(ws + wl.dqkv);
  float* delta = reinterpret_cast<float*>(ws + wl.delta);
  const long cls_stride_b = (long)(L + 1) * H;
  const int idw = id_bytes(cfg);

  auto G = [&](int idx) -> float* { return grads + grad_offset(cfg, idx); };
  // the gradient slices this call produces are accumulated into (split-K wgrads, atomics): zero them first
  {
    const int l0 = layer_lo < 0 ? 0 : layer_lo;
    SIMX_CUDA(cudaMemsetAsync(grads + emb_grad_numel(cfg) + layer_grad_numel(cfg) * l0, 0,
                              (size_t)layer_grad_numel(cfg) * (layer_hi - l0) * sizeof(float), st));
    if (layer_lo < 0) SIMX_CUDA(cudaMemsetAsync(grads, 0, (size_t)emb_grad_numel(cfg) * sizeof(float), st));
    if (layer_hi == L && t.bert)  // pooler: dead math in the reference (x 0), dense zero gradients
      SIMX_CUDA(cudaMemsetAsync(grads + emb_grad_numel(cfg) + layer_grad_numel(cfg) * L, 0, ((size_t)H * H + H) * sizeof(float),
                                st));
  }

  if (layer_hi == L) {  // cotangent of the last hidden state
    SIMX_TRY(f32_to_planes(np, d_seq, (long)T * H, dx, psH, st));
    if (d_cls) SIMX_TRY(cls_scatter_add(np, dx, B, S, H, psH, d_cls + (long)L * H, cls_stride_b, st));
  }

  for (int l = layer_hi - 1; l >= 0 && l >= layer_lo; --l) {
    const float* const* P = params + t.base + 16 * l;
    const int pi = t.base + 16 * l;
    const uint8_t* w = wb + lw.total * l;
    const bf16* wqkvT = reinterpret_cast<const bf16*>(w + lw.wqkvT);
    const bf16* woT = reinterpret_cast<const bf16*>(w + lw.woT);
    const bf16* w1T = reinterpret_cast<const bf16*>(w + lw.w1T);
    const bf16* w2T = reinterpret_cast<const bf16*>(w + lw.w2T);
    const bf16* x = reinterpret_cast<const bf16*>(sb + sl.x + sl.x_stride * l);
    LayerBufs lb;
    if (ckpt(cfg)) {  // recompute this layer's activations from its saved input (same dropout masks: counter-based)
      lb = layer_bufs(ws + wl.scratch, sl);
      SIMX_TRY(layer_forward(cfg, P, w, lw, l, x, nullptr, lb, mask, B, S, true, st));
    } else {
      lb = layer_bufs(sb + sl.layer0 + sl.l_total * l, sl);
    }
    const DropSpec d_attn = drop_at(cfg, DROP_SITE_ATTN, l);
    const DropSpec d_ao = drop_at(cfg, DROP_SITE_ATTN_OUT, l);
    const DropSpec d_fo = drop_at(cfg, DROP_SITE_FFN_OUT, l);

    // output LayerNorm -> dz2 (= d residual x1); dz2m = dz2 o mask flows into the FFN-down dense (dz2m == dz2 without
    // dropout); dgamma/dbeta, and db2 = colsum(dz2m)
    const bf16* dz2m = d_fo.thr_hi ? dzm : dza;
    SIMX_TRY(ln_bwd(np, dx, lb.z2, T, H, P[14], cfg->ln_eps, dza, psH, G(pi + 14), G(pi + 15), G(pi + 13), dzm, 2, d_fo, st));
    // dW2 [H,I] += dz2m^T a
    prof_label("ffn2_wgrad");
    SIMX_TRY(gemm(H, I, T, dz2m, H, psH, true, lb.a, I, psI, true, np, SIMX_EPI_ATOMIC_F32, nullptr, nullptr, 0, 0, nullptr,
                  nullptr, 0, 0, G(pi + 12), I, wgrad_ksplit(H, I, T), st));
    // dh = (dz2m W2) o gelu'(pre-activation)   (the forward epilogue saved gelu' in the `h` slot); db1 = colsum(dh)
    // comes out of the same epilogue (a separate column-sum pass only where the staged epilogue does not run)
    prof_label("ffn2_dgrad");
    SIMX_TRY(gemm(T, I, H, dz2m, H, psH, false, w2T, H, (long)I * H, false, np, SIMX_EPI_MUL, nullptr, lb.h, I, psI, dh,
                  nullptr, I, psI, G(pi + 11), 0, 1, st));
    // dW1 [I,H] += dh^T x1
    prof_label("ffn1_wgrad");
    SIMX_TRY(gemm(I, H, T, dh, I, psI, true, lb.x1, H, psH, true, np, SIMX_EPI_ATOMIC_F32, nullptr, nullptr, 0, 0, nullptr,
                  nullptr, 0, 0, G(pi + 10), H, wgrad_ksplit(I, H, T), st));
    // dx1 = dh W1 + dz2 (the residual path carries the un-masked dz2)
    prof_label("ffn1_dgrad");
    SIMX_TRY(gemm(T, H, I, dh, I, psI, false, w1T, I, (long)I * H, false, np, SIMX_EPI_ADD, nullptr, dza, H, psH, dxb,
                  nullptr, H, psH, nullptr, 0, 1, st));
    prof_label(nullptr);
    // attention-output LayerNorm -> dz1 (+ masked copy for the dense), and dbo = colsum(dz1m)
    const bf16* dz1m = d_ao.thr_hi ? dzm : dza;
    SIMX_TRY(ln_bwd(np, dxb, lb.z1, T, H, P[8], cfg->ln_eps, dza, psH, G(pi + 8), G(pi + 9), G(pi + 7), dzm, 2, d_ao, st));
    // dWo [H,H] += dz1m^T ctx
    prof_label("attnout_wgrad");
    SIMX_TRY(gemm(H, H, T, dz1m, H, psH, true, lb.ctx, H, psH, true, np, SIMX_EPI_ATOMIC_F32, nullptr, nullptr, 0, 0, nullptr,
                  nullptr, 0, 0, G(pi + 6), H, wgrad_ksplit(H, H, T), st));
    // dctx = dz1m Wo
    prof_label("attnout_dgrad");
    SIMX_TRY(gemm(T, H, H, dz1m, H, psH, false, woT, H, (long)H * H, false, np, SIMX_EPI_PLAIN, nullptr, nullptr, 0, 0, dxb,
                  nullptr, H, psH, nullptr, 0, 1, st));
    prof_label(nullptr);
    // dqkv, and the fused q/k/v bias gradient (column sums of dqkv) from the same kernel
    SIMX_TRY(attention_backward(np, lb.qkv, ps3H, lb.ctx, dxb, psH, dqkv, mask, idw, lb.lse, delta, B, S, H, cfg->heads,
                                G(pi + 1), d_attn, st));
    // fused q/k/v weight gradient ([3H,H] block of the flat gradient buffer)
    prof_label("qkv_wgrad");
    SIMX_TRY(gemm(3 * H, H, T, dqkv, 3 * H, ps3H, true, x, H, psH, true, np, SIMX_EPI_ATOMIC_F32, nullptr, nullptr, 0, 0,
                  nullptr, nullptr, 0, 0, G(pi + 0), H, wgrad_ksplit(3 * H, H, T), st));
    // dx(layer input) = dqkv Wqkv + dz1   (+ the [CLS] cotangent LEAD injects at this depth)
    prof_label("qkv_dgrad");
    SIMX_TRY(gemm(T, H, 3 * H, dqkv, 3 * H, ps3H, false, wqkvT, 3 * H, 3L * H * H, false, np, SIMX_EPI_ADD, nullptr, dza, H,
                  psH, dx, nullptr, H, psH, nullptr, 0, 1, st));
    prof_label(nullptr);
    if (d_cls) SIMX_TRY(cls_scatter_add(np, dx, B, S, H, psH, d_cls + (long)l * H, cls_stride_b, st));
  }
  if (layer_lo >= 0) return SIMX_OK;
  // embeddings: x0 = dropout(LN(z0)); the mask is re-applied to the cotangent on load
  const bf16* z0 = reinterpret_cast<const bf16*>(sb + sl.z0);
  const int ig = t.bert ? 3 : 2;
  SIMX_TRY(ln_bwd(np, dx, z0, T, H, params[ig], cfg->ln_eps, dza, psH, G(ig), G(ig + 1), t.bert ? G(2) : nullptr, nullptr, 1,
                  drop_at(cfg, DROP_SITE_EMBED, 0), st));
  if (pos_ids == nullptr) {
    // position rows: dPos[s] = sum_b dz0[b*S + s]: a column sum of the [B, S*H] view
    SIMX_TRY(colsum(np, dza, B, S * H, (long)S * H, psH, G(1), st));
  } else {
    // caller-supplied position ids (RoBERTa): scatter-add like the word rows; the pad position row gets nothing
    SIMX_TRY(embed_scatter(np, pos_ids, idw, dza, T, H, cfg->max_pos, pad_id, psH, G(1), st));
  }
  if (ids != nullptr) SIMX_TRY(embed_scatter(np, ids, idw, dza, T, H, cfg->vocab, pad_id, psH, G(0), st));
  // caller-supplied input embeddings: their cotangent is dz0 itself (the word table took no part: its gradient stays zero)
  if (d_embeds != nullptr) SIMX_TRY(planes_to_f32(np, dza, (long)T * H, psH, d_embeds, st));
  return SIMX_OK;
}

extern "C" int simx_encoder_hidden_state(const simx_encoder_cfg* cfg, const void* saved, int B, int S, int index, float* out,
                                         int64_t out_stride_b, void* stream) {
  SIMX_TRY(check_cfg(cfg));
  SIMX_CHECK(saved && out && B > 0 && S > 0 && S <= 512, SIMX_ERR_INVALID, "encoder_hidden_state: bad argument");
  SIMX_CHECK(index >= 0 && index <= cfg->num_layers, SIMX_ERR_INVALID, "encoder_hidden_state: index %d outside [0, %d]", index,
             cfg->num_layers);
  SIMX_CHECK(out_stride_b >= (long)S * cfg->hidden, SIMX_ERR_INVALID, "encoder_hidden_state: out_stride_b %ld < S*H", out_stride_b);
  const SavedL sl = saved_layout(cfg, B, S);
  const uint8_t* sb = align_ptr(saved);
  const long T = (long)B * S;
  return planes_to_f32_strided(cfg->precision, reinterpret_cast<const bf16*>(sb + sl.x + sl.x_stride * index), B,
                               (long)S * cfg->hidden, T * cfg->hidden, out, out_stride_b, static_cast<cudaStream_t>(stream));
}

extern "C" int simx_encoder_add_cotangent(const simx_encoder_cfg* cfg, void* workspace, int B, int S, const float* d_hidden,
                                          int64_t d_stride_b, void* stream) {
  SIMX_TRY(check_cfg(cfg));
  SIMX_CHECK(workspace && d_hidden && B > 0 && S > 0 && S <= 512, SIMX_ERR_INVALID, "encoder_add_cotangent: bad argument");
  SIMX_CHECK(d_stride_b >= (long)S * cfg->hidden, SIMX_ERR_INVALID, "encoder_add_cotangent: d_stride_b %ld < S*H", d_stride_b);
  const WsL wl = ws_layout(cfg, B, S);
  uint8_t* ws = align_ptr(workspace);
  const long T = (long)B * S;
  return planes_add_f32(cfg->precision, reinterpret_cast<bf16*>(ws + wl.dx), B, (long)S * cfg->hidden, T * cfg->hidden, d_hidden,
                        d_stride_b, static_cast<cudaStream_t>(stream));
}

extern "C" int simx_encoder_backward(const simx_encoder_cfg* cfg, const float* const* params, const void* wbuf,
                                     const void* ids, const void* mask, int B, int S, const void* saved,
                                     void* workspace, const float* d_cls, const float* d_seq, float* grads,
                                     void* stream) {
  SIMX_TRY(check_cfg(cfg));
  return encoder_backward_impl(cfg, params, wbuf, ids, mask, nullptr, 0, B, S, saved, workspace, d_cls, d_seq, grads, -1,
                               cfg->num_layers, stream);
}
extern "C" int simx_encoder_backward_pos(const simx_encoder_cfg* cfg, const float* const* params, const void* wbuf,
                                         const void* ids, const void* mask, const void* pos_ids, int pad_token_id,
                                         int B, int S, const void* saved, void* workspace, const float* d_cls,
                                         const float* d_seq, float* grads, void* stream) {
  SIMX_TRY(check_cfg(cfg));
  SIMX_CHECK(pos_ids != nullptr, SIMX_ERR_INVALID, "encoder_backward_pos: null position ids");
  return encoder_backward_impl(cfg, params, wbuf, ids, mask, pos_ids, pad_token_id, B, S, saved, workspace, d_cls, d_seq,
                               grads, -1, cfg->num_layers, stream);
}
extern "C" int simx_encoder_backward_embeds(const simx_encoder_cfg* cfg, const float* const* params, const void* wbuf,
                                            const void* mask, const void* pos_ids, int pad_token_id, int B, int S,
                                            const void* saved, void* workspace, const float* d_cls, const float* d_seq,
                                            float* grads, float* d_inputs_embeds, void* stream) {
  SIMX_TRY(check_cfg(cfg));
  SIMX_CHECK(d_inputs_embeds != nullptr, SIMX_ERR_INVALID, "encoder_backward_embeds: null d_inputs_embeds");
  return encoder_backward_impl(cfg, params, wbuf, nullptr, mask, pos_ids, pad_token_id, B, S, saved, workspace, d_cls, d_seq,
                               grads, -1, cfg->num_layers, stream, d_inputs_embeds);
}
extern "C" int simx_encoder_backward_layers(const simx_encoder_cfg* cfg, const float* const* params, const void* wbuf,
                                            const void* ids, const void* mask, const void* pos_ids, int pad_token_id,
                                            int B, int S, const void* saved, void* workspace, const float* d_cls,
                                            const float* d_seq, float* grads, int layer_lo, int layer_hi, void* stream) {
  return encoder_backward_impl(cfg, params, wbuf, ids, mask, pos_ids, pad_token_id, B, S, saved, workspace, d_cls, d_seq,
                               grads, layer_lo, layer_hi, stream);
}

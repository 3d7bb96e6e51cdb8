/* libsimx_b200 — C ABI of the B200-native dual-encoder training step.
 *
 * The reference (microsoft/SimXNS) has no FFI: its hot path is reached through a Python module import
 * (`from model.models import BiBertEncoder, HFBertEncoder, Reranker`, SimANS/co_training/co_training_marco_train.py:18).
 * This header is the boundary a replacement `models.py` binds (simxns_b200/_lib.py does it with ctypes;
 * INTEGRATION.md shows the stub).  Each entry point names the reference code it replaces.
 *
 * Conventions
 *   - plain pointers and sizes only; every device buffer is allocated and owned by the caller (PyTorch);
 *     the library never allocates or frees device memory and keeps no pointer past a call;
 *   - every function returns SIMX_OK (0) or a SIMX_ERR_* code; `simx_last_error()` returns the thread-local
 *     message of the last failure; no C++ exception crosses the ABI;
 *   - all work is enqueued on the caller's `stream` (a cudaStream_t passed as void*); calls are re-entrant;
 *   - matrices are row-major; "planes" is 1 (bf16) or 3 (split-bf16, fp32-class: value = hi + mid + lo, plane p
 *     at element offset p * plane_stride).
 */
#ifndef SIMX_B200_H_
#define SIMX_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SIMX_OK 0
#define SIMX_ERR_INVALID 1     /* bad argument / unsupported shape */
#define SIMX_ERR_CUDA 2        /* a CUDA runtime / driver call failed */
#define SIMX_ERR_UNSUPPORTED 3 /* feature not available (e.g. not an sm_100 device) */

#define SIMX_ARCH_BERT 0       /* HF BertModel: word+position+token_type embeddings (LEAD/modeling_bert.py:181-660) */
#define SIMX_ARCH_DISTILBERT 1 /* HF DistilBertModel: word+position embeddings (LEAD/modeling_distilbert.py:93-376) */

#define SIMX_PREC_BF16 1   /* bf16 operands, fp32 accumulate (throughput mode) */
#define SIMX_PREC_BF16X3 3 /* 3 bf16 planes per activation/weight, 6 MMA passes (fp32-class parity mode) */

typedef struct simx_encoder_cfg {
  int32_t arch;         /* SIMX_ARCH_* */
  int32_t num_layers;   /* L */
  int32_t hidden;       /* H, multiple of 64 */
  int32_t heads;        /* H / heads must be 64 */
  int32_t intermediate; /* I, multiple of 64 */
  int32_t vocab;
  int32_t max_pos;
  int32_t type_vocab;   /* 0 for DistilBERT */
  int32_t precision;    /* SIMX_PREC_* */
  float ln_eps;
} simx_encoder_cfg;

/* ------------------------------------------------------------------------------------------------
 * Library
 * ---------------------------------------------------------------------------------------------- */
const char* simx_last_error(void);
int simx_version(void);
/* 1 when the current CUDA device can run the kernels (compute capability 10.x), else 0. */
int simx_device_supported(void);

/* Per-launch device timing for bench.py's roofline lines: between begin and end every kernel wrapper records CUDA
 * events on its launching stream; end synchronises and writes a JSON array aggregated by kernel label:
 * [{"name","launches","ms","flops","bytes"}] (algorithmic FLOPs / bytes as stated in DESIGN.md). */
int simx_profile_begin(void);
int simx_profile_end(char* json_buf, size_t capacity);

/* ------------------------------------------------------------------------------------------------
 * Parameter table.  `params` / offsets use this canonical order (fp32 tensors, HF state_dict names):
 *   BERT:  0 word_embeddings [V,H]  1 position_embeddings [P,H]  2 token_type_embeddings [Tt,H]
 *          3 embeddings.LayerNorm.weight  4 .bias
 *          5+16*l + {0 q.w [H,H], 1 q.b, 2 k.w, 3 k.b, 4 v.w, 5 v.b, 6 attn.out.w [H,H], 7 attn.out.b,
 *                    8 attn.LN.w, 9 attn.LN.b, 10 intermediate.w [I,H], 11 intermediate.b,
 *                    12 output.w [H,I], 13 output.b, 14 output.LN.w, 15 output.LN.b}
 *          5+16*L pooler.w [H,H]   6+16*L pooler.b      (pooler is dead math in the reference:
 *                                                         SimANS/model/models.py:80 multiplies it by 0)
 *   DistilBERT: same table without entry 2 and without the pooler (indices shift down accordingly).
 * ---------------------------------------------------------------------------------------------- */
int simx_encoder_num_params(const simx_encoder_cfg* cfg);
/* number of elements of parameter `idx` */
int64_t simx_encoder_param_numel(const simx_encoder_cfg* cfg, int idx);
/* element offset of parameter `idx`'s gradient inside the flat fp32 gradient buffer written by
 * simx_encoder_backward (q/k/v weights and biases are adjacent there: the fused QKV wgrad writes them as one
 * [3H,H] block).  Total size: simx_encoder_grad_numel(). */
int64_t simx_encoder_grad_offset(const simx_encoder_cfg* cfg, int idx);
int64_t simx_encoder_grad_numel(const simx_encoder_cfg* cfg);

/* Buffer sizes in bytes for a batch of B sequences padded to S tokens. */
size_t simx_encoder_weight_bytes(const simx_encoder_cfg* cfg);
size_t simx_encoder_saved_bytes(const simx_encoder_cfg* cfg, int B, int S);
size_t simx_encoder_workspace_bytes(const simx_encoder_cfg* cfg, int B, int S);

/* fp32 master weights -> the bf16 (or split-bf16) GEMM operands in `wbuf`: fused [3H,H] QKV weight, and the
 * transposed copies the dgrad GEMMs read.  Call after every optimizer step (or parameter load). */
int simx_encoder_prepare_weights(const simx_encoder_cfg* cfg, const float* const* params, void* wbuf, void* stream);

/* Replaces HFBertEncoder.forward -> transformers BertModel.forward
 * (SimANS/model/models.py:77-82, LEAD/models.py:184-196; math restated in LEAD/modeling_bert.py:201-660).
 *   ids  [B,S] int64 token ids;  mask [B,S] int64 attention mask (1 = attend, HF semantics; may be NULL = all ones)
 *   cls_out  [B, L+1, H] fp32: the [CLS] (position 0) row of the embedding output (index 0) and of every
 *            layer's output (index l+1) — index L is `pooled_output` of the reference wrapper, the others are
 *            what LEAD reads from its all-layer hidden states (LEAD/util.py:605-614).
 *   seq_out  [B,S,H] fp32 last hidden state, or NULL to skip materialising it.
 *   saved    activations kept for backward (simx_encoder_saved_bytes); NULL = inference (nothing kept).
 * Dropout is not applied (the reference's parity/eval configuration). */
int simx_encoder_forward(const simx_encoder_cfg* cfg, const float* const* params, const void* wbuf,
                         const int64_t* ids, const int64_t* mask, int B, int S, void* saved, void* workspace,
                         float* cls_out, float* seq_out, void* stream);

/* Replaces loss.backward() through the tower (co_training_marco_train.py:222).
 *   d_cls  [B, L+1, H] fp32 cotangent of cls_out (NULL rows are not allowed; pass zeros);
 *   d_seq  [B,S,H] fp32 cotangent of seq_out or NULL.
 *   grads  flat fp32 buffer of simx_encoder_grad_numel() elements, OVERWRITTEN with d(loss)/d(param)
 *          (dense for every parameter: pooler and unused token-type rows get zeros). */
int simx_encoder_backward(const simx_encoder_cfg* cfg, const float* const* params, const void* wbuf,
                          const int64_t* ids, const int64_t* mask, int B, int S, const void* saved,
                          void* workspace, const float* d_cls, const float* d_seq, float* grads, void* stream);

/* RoBERTa towers (SimANS MS-Doc: HFRobertaEncoder SimANS/model/models.py:168-192, RobertaDot :333-352 over HF
 * RobertaModel) are BERT towers (type_vocab 1, ln_eps 1e-5, max_pos 514) whose position rows are chosen by the caller:
 * pos_ids [B,S] int64 = pad + cumsum(ids != pad) for real tokens, pad for padding (HF create_position_ids_from_input_ids).
 * Backward scatters the position gradient by pos_ids, and row `pad_token_id` of BOTH embedding tables receives none
 * (nn.Embedding(padding_idx = pad) on both; the plain entry points above use pad id 0 for the word table, BERT's). */
int simx_encoder_forward_pos(const simx_encoder_cfg* cfg, const float* const* params, const void* wbuf,
                             const int64_t* ids, const int64_t* mask, const int64_t* pos_ids, int B, int S, void* saved,
                             void* workspace, float* cls_out, float* seq_out, void* stream);
int simx_encoder_backward_pos(const simx_encoder_cfg* cfg, const float* const* params, const void* wbuf,
                              const int64_t* ids, const int64_t* mask, const int64_t* pos_ids, int pad_token_id, int B,
                              int S, const void* saved, void* workspace, const float* d_cls, const float* d_seq,
                              float* grads, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Fused score / loss / gradient step (replaces the tail of the train step:
 * co_training_marco_train.py:199-214, SimANS/model/models.py:468-505, PROD/ProD_KD/model/models.py:669-781,
 * MASTER/finetune/MS/co_training_model.py:249-269).  See simx_loss_cfg below.
 * ---------------------------------------------------------------------------------------------- */
#define SIMX_LOSS_NLL_GLOBAL 1   /* a-5  InfoNCE over the global [GQ,GC] matrix, positives at pos_idx[q] */
#define SIMX_LOSS_KL_LOCAL 2     /* a-4  SimANS: KLDiv_batchmean(log(softmax(s)+eps), softmax(teacher/T)) on [B,1+n] */
#define SIMX_LOSS_CE_LOCAL 4     /* a-4' / a-5': -sum p_t log(p_s+eps) / B  on [B,1+n] */
#define SIMX_LOSS_KD_LOCAL 8     /* a-6  PROD KD_softmax: T^2 * mean_b KL(softmax(t/T) || softmax(s/T)) on [B,1+n] */
#define SIMX_LOSS_NLL_LOCAL 16   /* a-6  hard loss of CrossBERTKDLoss: NLL of column 0 on [B,1+n] */
#define SIMX_LOSS_ADV_LOCAL 32   /* a-4' adversarial term of the wiki driver: sum_{b,j} reward_bj log(p_s+eps),
                                    reward_bj = log(sigmoid(t_b0 - t_bj) + eps) on the raw teacher logits
                                    (SimANS/wiki/co_training_wiki_train.py:212-225); a SUM, not a batch mean */

#define SIMX_LOSS_MSE_LOCAL 64   /* a-6  PROD KD_logit: 0.5 * mean_{b,j} (s - t)^2 on the raw logits (models.py:778-780) */
#define SIMX_LOSS_DKD_LOCAL 128  /* a-6  PROD DKD: T^2/B * sum_b [alpha TCKD + beta NCKD], target = column 0
                                    (PROD/ProD_KD/model/models.py:783-812) */

typedef struct simx_loss_cfg {
  int32_t terms;        /* OR of SIMX_LOSS_* */
  int32_t GQ, GC, H;    /* global query / passage counts, embedding width */
  int32_t q_begin, q_count; /* rows [q_begin, q_begin+q_count) of Q are this rank's queries (get dQ) */
  int32_t c_begin, c_count; /* rows of C that are this rank's passages (get dC) */
  int32_t group;        /* 1+n passages per local query: local passage block of query i = c_begin + i*group */
  float w_nll_global, w_kl_local, w_ce_local, w_kd_local, w_nll_local; /* term weights */
  float temperature;    /* T of the teacher softmax (KL/CE: softmax(t/T); KD: both sides /T, times T^2) */
  float eps;            /* 1e-7 inside the log of KL/CE (co_training_marco_train.py:213) */
  float score_scale;    /* multiplies student local scores (1 or 1/sqrt(H) with --scale_simmila) */
  float grad_scale;     /* multiplies every gradient and nothing else (loss / gradient_accumulation_steps) */
  float w_adv_local;    /* weight of SIMX_LOSS_ADV_LOCAL (adv_lambda) */
  float w_mse_local, w_dkd_local; /* weights of SIMX_LOSS_MSE_LOCAL / SIMX_LOSS_DKD_LOCAL */
  float dkd_alpha, dkd_beta;      /* args.DKD_alpha / args.DKD_beta */
} simx_loss_cfg;

/*   Q [GQ,H], C [GC,H] fp32 gathered embeddings; pos_idx [GQ] int32 positive column of each global query
 *   (NLL_GLOBAL); teacher [q_count, group] fp32 teacher logits (KL/CE/KD terms), or NULL.
 *   out_losses[12] fp32: {total, nll_global, kl_local, ce_local, kd_local, nll_local, n_correct_global,
 *                       n_correct_local (local rows whose column 0 has the largest score), mse_local, dkd_local, 0, 0}
 *   dQ [q_count,H], dC [c_count,H] fp32 gradients of `total * grad_scale` w.r.t. the local embeddings.
 *   workspace: simx_score_loss_workspace_bytes(cfg). */
size_t simx_score_loss_workspace_bytes(const simx_loss_cfg* cfg);
int simx_score_loss(const simx_loss_cfg* cfg, const float* Q, const float* C, const int32_t* pos_idx,
                    const float* teacher, float* out_losses, float* dQ, float* dC, void* workspace, void* stream);

/* LEAD mutual distillation between two dual encoders on the global score matrices A = Qa Ca^T (student) and
 * B = Qb Cb^T (teacher)  — replaces the DE->DB branch of distill_loss (LEAD/util.py:594-644, 737-760) and, called once per
 * selected layer pair on the per-layer [CLS] rows, layer_score_dis_loss (LEAD/util.py:539-559):
 *   total = w_nll_a NLL(A) + w_nll_b NLL(B) + w_kl_b2a KL(softmax(B/T) || softmax(A)) + w_kl_a2b KL(softmax(A/T) || softmax(B))
 * KL is `batchmean` over GQ rows; targets are NOT detached (as in the reference): both matrices receive gradients from
 * every term.  out_losses[8] = {total, nll_a, nll_b, kl_b2a, kl_a2b, kd_b2a, n_correct_a, mse_b2a + dkd_b2a}.  dQb/dCb may be NULL (frozen teacher).
 * Local row/column ranges as in simx_loss_cfg; gathered embeddings fp32. */
typedef struct simx_pair_cfg {
  int32_t GQ, GC, H;
  int32_t q_begin, q_count, c_begin, c_count;
  float w_nll_a, w_nll_b, w_kl_b2a, w_kl_a2b;
  float temperature;
  float grad_scale;
  float w_kd_b2a; /* PROD BiEncoderKDLoss KD_softmax on the global matrix: T^2 KL(softmax(B/T) || softmax(A/T))
                     (PROD/ProD_KD/model/models.py:971-1059) */
  float w_mse_b2a; /* PROD KD_logit on the global matrix: 0.5 mean (A - B)^2                 (models.py:1057-1059) */
  float w_dkd_b2a; /* PROD DKD on the global matrix, target = pos_idx, with dkd_alpha / dkd_beta (models.py:1070-1094).
                      B is a constant for these two terms (PROD runs its teacher under no_grad): no dQb / dCb share. */
  float dkd_alpha, dkd_beta;
} simx_pair_cfg;
size_t simx_pair_distill_workspace_bytes(int GQ, int GC, int H);
int simx_pair_distill(const simx_pair_cfg* cfg, const float* Qa, const float* Ca, const float* Qb, const float* Cb,
                      const int32_t* pos_idx, float* out_losses, float* dQa, float* dCa, float* dQb, float* dCb,
                      void* workspace, void* stream);

/* LEAD cross-encoder-teacher variant (--distill_ce: LEAD/util.py:669-717, 801-829): the same four terms on each query's
 * own [1+n] block (student scores q_i . c_ij vs. the teacher's relevance logits [q_count, 1+n], positive = column 0).
 * Uses cfg->q_count, c_count (= q_count * group), H, the four weights, temperature, grad_scale.  dT [q_count, 1+n] =
 * d(total)/d(teacher logits) or NULL. */
int simx_local_pair_distill(const simx_pair_cfg* cfg, const float* Q_local, const float* C_local,
                            const float* teacher_logits, float* out_losses, float* dQ, float* dC, float* dT, void* stream);

/* One half of PROD's Bi_logit (BiEncoderKDLoss.bi_logit, PROD/ProD_KD/model/models.py:1062-1067) on raw embeddings:
 * *loss += weight * 0.5 * mean((a - b)^2) and grad[i] = weight * (a[i] - b[i]) / n   (a, b, grad: n fp32 elements). */
int simx_embed_mse(const float* a, const float* b, int64_t n, float weight, float* loss, float* grad, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Exact inner-product top-k over a resident embedding shard — replaces, for the ANN refresh of every
 * SimANS round, faiss.IndexFlatIP + index_cpu_to_all_gpus(shard=True).add(P) and .search(Q, 200 | 1000)
 * (SimANS/co_training/co_training_generate.py:359-384, 412-427).
 *   corpus_planes: caller-owned device buffer of simx_topk_corpus_bytes(N, H), 256-byte aligned, filled slab by slab with
 *   simx_topk_add (fp32 rows [n_rows, H] at row_offset; the rows are kept as 3 bf16 planes = fp32-class operands of the
 *   tcgen05 score GEMM).  simx_topk_search: Q [nq, H] fp32 -> out_scores [nq, k] fp32 (descending), out_idx [nq, k]
 *   int64 row numbers inside this shard (ties: lowest row first; -1 / -inf padding when N < k).  k <= 1024, nq <= 65535,
 *   N < 2^31, H % 64 == 0.
 * ---------------------------------------------------------------------------------------------- */
size_t simx_topk_corpus_bytes(int64_t N, int H);
int simx_topk_add(const float* P, int64_t n_rows, int64_t row_offset, int64_t N_total, int H, void* corpus_planes,
                  void* stream);
size_t simx_topk_workspace_bytes(int nq, int64_t N, int H, int k);
int simx_topk_search(const float* Q, int nq, const void* corpus_planes, int64_t N, int H, int k, float* out_scores,
                     int64_t* out_idx, void* workspace, void* stream);

/* SimANS ambiguous-negative sampling (SimANS/utils/MARCO_until_new.py:165-202, SimANS/utils/util_wiki.py:613-639):
 * for each of NQ queries choose k of its first n_valid[q] (<= NC <= 1024) candidates, without duplicates, with probability
 * proportional to  exp(-p0 |s - s+|)  (mode 0, p0 = tau)  or  exp(-p0 (s - s+ + p1)^2)  (mode 1, p0 = a, p1 = b); queries
 * whose pos_score is 0 take the last k candidates.  scores [NQ,NC] fp32, pos_score [NQ], out_idx [NQ,k] int32 candidate
 * indices.  Same rounds as the reference (k draws with replacement, keep distinct, drop them from the pool, repeat);
 * counter-based RNG from `seed` — distribution-level parity only, the reference's stream is CPython's. */
int simx_simans_sample(const float* scores, const int32_t* n_valid, const float* pos_score, int NQ, int NC, int k,
                       int mode, float p0, float p1, uint64_t seed, int32_t* out_idx, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Optimizer tail: torch.nn.utils.clip_grad_norm_ + AdamW.step over every parameter tensor of a model in
 * two launches (co_training_marco_train.py:57-69,246-263; LEAD/util.py:901,  LEAD/run_LEAD.py:232-238).
 *   SIMX_ADAMW_TORCH  torch.optim.AdamW (LEAD): p *= 1 - lr wd;  p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
 *   SIMX_ADAMW_HF     transformers.AdamW (SimANS, PROD; transformers v4 optimization.py, not vendored):
 *                     p -= lr sqrt(1-b2^t)/(1-b1^t) * m / (sqrt(v) + eps)   [plain lr if !correct_bias];  then p -= lr wd p
 * The tensor table lives in DEVICE memory (the caller copies it; the library keeps nothing).  Build it on the host,
 * call simx_opt_plan() to fill `chunk_begin` and get the grid size, then copy it to the device.
 * ---------------------------------------------------------------------------------------------- */
#define SIMX_ADAMW_TORCH 0
#define SIMX_ADAMW_HF 1
typedef struct simx_opt_tensor {
  float* p;            /* parameter (fp32 master), updated in place */
  float* g;            /* gradient, or NULL: the tensor is skipped (a parameter that received no gradient) */
  float* m;            /* exp_avg */
  float* v;            /* exp_avg_sq */
  int64_t numel;
  int32_t chunk_begin; /* filled by simx_opt_plan */
  float lr;            /* the tensor's param-group learning rate at this step (the scheduler's value) */
  float weight_decay;  /* 0 for the bias / LayerNorm.weight group */
  int32_t reserved;
} simx_opt_tensor;
/* host helper: fills chunk_begin of tensors_host[0..n) and returns the number of 8192-element chunks (= grid size), or
 * -1 on a bad table. */
int simx_opt_plan(simx_opt_tensor* tensors_host, int n);
/* *out_sq_dev (device, fp64) = sum over all tensors of sum(g^2): total_norm^2 of clip_grad_norm_ (L2). */
int simx_grad_sq_norm(const simx_opt_tensor* tensors_dev, int n, int total_chunks, double* out_sq_dev, void* stream);
/* g *= min(1, max_norm / (sqrt(*sq_norm_dev) + 1e-6)) in place — what clip_grad_norm_ leaves in .grad. */
int simx_clip_grads(const simx_opt_tensor* tensors_dev, int n, int total_chunks, const double* sq_norm_dev,
                    float max_norm, void* stream);
/* One AdamW step (`step` counts from 1).  sq_norm_dev != NULL: the gradient is clipped on the fly by the coefficient
 * above (no host round trip, no extra pass); write_clipped_grads stores the clipped gradient back as well. */
int simx_adamw_step(const simx_opt_tensor* tensors_dev, int n, int total_chunks, int mode, double beta1, double beta2,
                    float eps, int step, int correct_bias, const double* sq_norm_dev, float max_norm,
                    int write_clipped_grads, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Operator-level entry points (used by the unit tests and by bench.py's kernel timing; the encoder calls the
 * same kernels internally).
 * ---------------------------------------------------------------------------------------------- */
#define SIMX_EPI_PLAIN 0     /* out = acc */
#define SIMX_EPI_BIAS 1      /* out = acc + bias[n]                                   (QKV projection) */
#define SIMX_EPI_BIAS_GELU 2 /* h = acc + bias[n]; out = gelu_erf'(h); out2 = gelu_erf(h) (BertIntermediate; out is what
                                backward multiplies by, the pre-activation itself is never needed again) */
#define SIMX_EPI_BIAS_ADD 3  /* out = acc + bias[n] + aux[m,n]                        (BertSelfOutput/BertOutput pre-LN) */
#define SIMX_EPI_MUL 4       /* out = acc * aux[m,n]                                  (dgrad through GELU: aux = saved gelu');
                                if outf != NULL: outf[n] += sum_m out[m,n] (fp32, [N]) — the bias gradient of the layer
                                whose dgrad this is, reduced in the epilogue instead of by a pass over `out` */
#define SIMX_EPI_ADD 5       /* out = acc + aux[m,n]                                  (dgrad + residual grad) */
#define SIMX_EPI_ATOMIC_F32 6 /* outf[m,n] += acc   (fp32 atomics; split-K wgrad)     */
#define SIMX_EPI_BIAS_GELU_FWD 7 /* out = gelu_erf(acc + bias[n])                     (BertIntermediate at inference: nothing is
                                    kept for a backward pass, so the derivative is neither computed nor stored) */

/* D[M,N] = sum_k A[m,k] * B[n,k]   on tcgen05 tensor cores (bf16 operands, fp32 accumulate in TMEM).
 *   a_mn_major == 0: A is [M,K] row-major (lda >= K);  == 1: A is stored [K,M] row-major (lda >= M).  Same for B/N.
 *   planes: operands, aux and bf16 outputs all carry `planes` planes; with 3 planes the 6 significant plane
 *   pairs are accumulated.  ksplit > 1 only with SIMX_EPI_ATOMIC_F32.
 *   out/out2/aux: bf16 [M,N] (ld*), outf: fp32 [M,N] (ATOMIC), bias: fp32 [N]. */
int simx_gemm(int M, int N, int K, const void* A, int64_t lda, int64_t a_plane_stride, int a_mn_major,
              const void* B, int64_t ldb, int64_t b_plane_stride, int b_mn_major, int planes, int epilogue,
              const float* bias, const void* aux, int64_t ldaux, int64_t aux_plane_stride, void* out, void* out2,
              int64_t ldo, int64_t out_plane_stride, float* outf, int64_t ldof, int ksplit, void* stream);

/* Self-attention of the encoder on the fused [T, 3H] projection (q | k | v column blocks, head h at columns h*64):
 *   ctx = softmax(q k^T / 8 + (1 - mask) * -inf) v       (LEAD/modeling_bert.py:324-368), head_dim 64, S <= 512.
 * lse2 [B, heads, S] (log2-domain log-sum-exp) is produced by forward and consumed by backward; delta [B, heads, S]
 * is backward scratch.  planes / plane strides as for simx_gemm.  dbias_qkv (optional, [3H] fp32): += column sums of dqkv,
 * i.e. the gradient of the fused q/k/v bias (reduced from the staged output tiles by the tcgen05 kernel). */
int simx_attention_forward(const void* qkv, void* ctx, const int64_t* mask, float* lse2, int B, int S, int H, int heads,
                           int planes, int64_t qkv_plane_stride, int64_t ctx_plane_stride, void* stream);
int simx_attention_backward(const void* qkv, const void* ctx, const void* dctx, void* dqkv, const int64_t* mask,
                            const float* lse2, float* delta, int B, int S, int H, int heads, int planes,
                            int64_t qkv_plane_stride, int64_t ctx_plane_stride, float* dbias_qkv, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SIMX_B200_H_ */

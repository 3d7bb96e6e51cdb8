// Score / loss / gradient step on the gathered [CLS] embeddings (C ABI: simx_score_loss).
//
//  * local terms (SimANS KL, CE-to-teacher, PROD KD_softmax, local NLL) live on the query's own [1+n] block:
//    ONE fused kernel, a warp per query — scores, softmax, every requested loss term, dL/ds, dQ and dC
//    (co_training_marco_train.py:199-214; PROD/ProD_KD/model/models.py:669-781).
//  * the global in-batch InfoNCE (BiEncoderNllLoss.calc on all-gathered embeddings, SimANS/model/models.py:468-505)
//    runs S = Q C^T, dQ = dS C and dC = dS^T Q on the tcgen05 GEMM in fp32-class (3-plane) precision, with one
//    row kernel in between doing log-sum-exp, NLL, arg-max and dS = (softmax - onehot) / GQ in place.
// Gradients reach only the local rows/columns; remote slices are constants (LEAD/util.py:200-204).
#include "common.cuh"
#include "ptx.cuh"

namespace simx {

int gemm(int M, int N, int K, const bf16* A, long lda, long a_ps, bool a_mn, const bf16* B, long ldb, long b_ps,
         bool b_mn, int planes, int epi, const float* bias, const bf16* aux, long ldaux, long aux_ps, bf16* out,
         bf16* out2, long ldo, long out_ps, float* outf, long ldof, int ksplit, cudaStream_t stream);
int convert_weight(int np, const float* w, int R, int C, bf16* dst, long ld_dst, long ps_dst, bf16* dst_t, long ld_t,
                   long ps_t, cudaStream_t st);

constexpr int MAX_GROUP = 64;

// ------------------------------------------------------------------------------------------------
// fused local-block kernel: one warp per local query
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
local_loss_kernel(const simx_loss_cfg cfg, const float* __restrict__ Q, const float* __restrict__ C,
                  const float* __restrict__ teacher, float* out_losses, float* dQ, float* dC, int accumulate) {
  __shared__ float sh_ds[4][MAX_GROUP];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * 4 + warp;
  if (i >= cfg.q_count) return;
  const int H = cfg.H, G = cfg.group;
  const float* q = Q + (long)(cfg.q_begin + i) * H;
  const float* c = C + (long)(cfg.c_begin + (long)i * G) * H;
  float* ds = sh_ds[warp];
  // scores
  for (int j = 0; j < G; ++j) {
    float acc = 0.f;
    for (int h = lane * 4; h < H; h += 128) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(q + h));
      const float4 b = __ldg(reinterpret_cast<const float4*>(c + (long)j * H + h));
      acc += a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
    }
    acc = warp_sum(acc) * cfg.score_scale;
    if (lane == 0) ds[j] = acc;
  }
  __syncwarp();
  // every lane redundantly walks the (tiny) row; lane 0 publishes
  const float T = cfg.temperature, invB = 1.f / cfg.q_count;
  float smax = -INFINITY, tmax = -INFINITY;
  for (int j = 0; j < G; ++j) {
    smax = fmaxf(smax, ds[j]);
    if (teacher) tmax = fmaxf(tmax, teacher[(long)i * G + j] / T);
  }
  float ssum = 0.f, ssumT = 0.f, tsum = 0.f;
  for (int j = 0; j < G; ++j) {
    ssum += expf(ds[j] - smax);
    ssumT += expf((ds[j] - smax) / T);
    if (teacher) tsum += expf(teacher[(long)i * G + j] / T - tmax);
  }
  const float lse = smax + logf(ssum);
  const float lseT = smax / T + logf(ssumT);
  const float tlse = teacher ? tmax + logf(tsum) : 0.f;
  float l_kl = 0.f, l_ce = 0.f, l_kd = 0.f, l_nll = 0.f, rsum = 0.f;
  const bool want_klce = (cfg.terms & (SIMX_LOSS_KL_LOCAL | SIMX_LOSS_CE_LOCAL)) != 0;
  for (int j = 0; j < G; ++j) {
    const float p = expf(ds[j] - lse);
    if (teacher) {
      const float tl = teacher[(long)i * G + j] / T - tlse;  // log t
      const float t = expf(tl);
      if (want_klce) {
        const float lp = logf(p + cfg.eps);
        l_ce += -t * lp;
        l_kl += (t > 0.f ? t * tl : 0.f) - t * lp;
        rsum += t * p / (p + cfg.eps);
      }
      if (cfg.terms & SIMX_LOSS_KD_LOCAL) l_kd += t > 0.f ? t * (tl - (ds[j] / T - lseT)) : 0.f;
    }
  }
  if (cfg.terms & SIMX_LOSS_NLL_LOCAL) l_nll = -(ds[0] - lse);
  const float w_klce = ((cfg.terms & SIMX_LOSS_KL_LOCAL) ? cfg.w_kl_local : 0.f) +
                       ((cfg.terms & SIMX_LOSS_CE_LOCAL) ? cfg.w_ce_local : 0.f);
  __syncwarp();
  // lane (j % 32) is the only reader and writer of ds[j] from here on: overwrite the score with dL/ds_j
  for (int j = lane; j < G; j += 32) {
    const float s = ds[j];
    const float p = expf(s - lse);
    float g = 0.f;
    if (teacher) {
      const float t = expf(teacher[(long)i * G + j] / T - tlse);
      if (want_klce) g += w_klce * (-(t * p / (p + cfg.eps) - p * rsum) * invB);
      if (cfg.terms & SIMX_LOSS_KD_LOCAL) g += cfg.w_kd_local * (T * invB) * (expf(s / T - lseT) - t);
    }
    if (cfg.terms & SIMX_LOSS_NLL_LOCAL) g += cfg.w_nll_local * invB * (p - (j == 0 ? 1.f : 0.f));
    ds[j] = g * cfg.score_scale * cfg.grad_scale;
  }
  __syncwarp();
  if (lane == 0) {
    const float T2 = T * T;
    if (cfg.terms & SIMX_LOSS_KL_LOCAL) atomicAdd(out_losses + 2, l_kl * invB);
    if (cfg.terms & SIMX_LOSS_CE_LOCAL) atomicAdd(out_losses + 3, l_ce * invB);
    if (cfg.terms & SIMX_LOSS_KD_LOCAL) atomicAdd(out_losses + 4, l_kd * invB * T2);
    if (cfg.terms & SIMX_LOSS_NLL_LOCAL) atomicAdd(out_losses + 5, l_nll * invB);
    float tot = 0.f;
    if (cfg.terms & SIMX_LOSS_KL_LOCAL) tot += cfg.w_kl_local * l_kl * invB;
    if (cfg.terms & SIMX_LOSS_CE_LOCAL) tot += cfg.w_ce_local * l_ce * invB;
    if (cfg.terms & SIMX_LOSS_KD_LOCAL) tot += cfg.w_kd_local * l_kd * invB * T2;
    if (cfg.terms & SIMX_LOSS_NLL_LOCAL) tot += cfg.w_nll_local * l_nll * invB;
    atomicAdd(out_losses + 0, tot);
  }
  // gradients: dQ_i = sum_j ds_j c_ij ; dC_ij = ds_j q_i      (this warp is the only writer of these rows)
  float* dq = dQ + (long)i * H;
  float* dc = dC + (long)i * G * H;
  for (int h = lane * 4; h < H; h += 128) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(q + h));
    float4 accq = accumulate ? *reinterpret_cast<float4*>(dq + h) : make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = 0; j < G; ++j) {
      const float g = ds[j];
      const float4 b = __ldg(reinterpret_cast<const float4*>(c + (long)j * H + h));
      accq.x += g * b.x; accq.y += g * b.y; accq.z += g * b.z; accq.w += g * b.w;
      float4 o = accumulate ? *reinterpret_cast<float4*>(dc + (long)j * H + h) : make_float4(0.f, 0.f, 0.f, 0.f);
      o.x += g * a.x; o.y += g * a.y; o.z += g * a.z; o.w += g * a.w;
      *reinterpret_cast<float4*>(dc + (long)j * H + h) = o;
    }
    *reinterpret_cast<float4*>(dq + h) = accq;
  }
}

// ------------------------------------------------------------------------------------------------
// global InfoNCE row kernel: one block per global query row of S [GQ, GC] (fp32).
// Writes dS = w * grad_scale * (softmax - onehot) / GQ as 3 bf16 planes, accumulates loss and arg-max hits.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
nll_row_kernel(const float* __restrict__ S, int GQ, int GC, int ldS, const int32_t* __restrict__ pos_idx, float w,
               float grad_scale, bf16* dS, long ds_ps, float* out_losses) {
  __shared__ float red_m[8], red_s[8];
  __shared__ int red_i[8];
  __shared__ float sh_lse;
  const int row = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* s = S + (long)row * ldS;
  float m = -INFINITY, sum = 0.f;
  int am = 0;
  for (int c = threadIdx.x; c < GC; c += 256) {
    const float v = s[c];
    if (v > m) { sum = sum * expf(m - v) + 1.f; m = v; am = c; }
    else sum += expf(v - m);
  }
  // warp combine (max, sum, argmax with lowest index on ties — torch.max returns the first maximal index)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o);
    const float s2 = __shfl_xor_sync(0xffffffffu, sum, o);
    const int a2 = __shfl_xor_sync(0xffffffffu, am, o);
    if (m2 > -INFINITY) {  // lanes that saw no column carry (-inf, 0): nothing to merge
      const float mn = fmaxf(m, m2);
      sum = (m > -INFINITY ? sum * expf(m - mn) : 0.f) + s2 * expf(m2 - mn);
      if (m2 > m || (m2 == m && a2 < am)) am = a2;
      m = mn;
    }
  }
  if (lane == 0) { red_m[warp] = m; red_s[warp] = sum; red_i[warp] = am; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float M = red_m[0], Ssum = red_s[0];
    int A = red_i[0];
    for (int k = 1; k < 8; ++k) {
      if (!(red_m[k] > -INFINITY)) continue;
      const float mn = fmaxf(M, red_m[k]);
      Ssum = Ssum * expf(M - mn) + red_s[k] * expf(red_m[k] - mn);
      if (red_m[k] > M || (red_m[k] == M && red_i[k] < A)) A = red_i[k];
      M = mn;
    }
    const float lse = M + logf(Ssum);
    sh_lse = lse;
    const int pos = pos_idx[row];
    const float nll = lse - s[pos];
    atomicAdd(out_losses + 1, nll / GQ);
    atomicAdd(out_losses + 0, w * nll / GQ);
    if (A == pos) atomicAdd(out_losses + 6, 1.f);
  }
  __syncthreads();
  const float lse = sh_lse;
  const int pos = pos_idx[row];
  const float k = w * grad_scale / GQ;
  for (int c = threadIdx.x; c < GC; c += 256) {
    const float g = k * (expf(s[c] - lse) - (c == pos ? 1.f : 0.f));
    split_store<3>(dS + (long)row * ldS + c, ds_ps, g);
  }
}

static inline size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

struct LossWs {
  size_t qp, cp, ctp, s, ds, total;
};
static LossWs loss_ws(const simx_loss_cfg* c) {
  LossWs w{};
  size_t off = 0;
  if (c->terms & SIMX_LOSS_NLL_GLOBAL) {
    const size_t GCp = (c->GC + 7) / 8 * 8;  // score-matrix row pitch (TMA needs 16-byte row strides)
    w.qp = off;  off += align256(3ull * c->GQ * c->H * 2);
    w.cp = off;  off += align256(3ull * GCp * c->H * 2);
    w.ctp = off; off += align256(3ull * GCp * c->H * 2);
    w.s = off;   off += align256(4ull * c->GQ * GCp);
    w.ds = off;  off += align256(3ull * c->GQ * GCp * 2);
  }
  w.total = off + 256;
  return w;
}

}  // namespace simx

extern "C" size_t simx_score_loss_workspace_bytes(const simx_loss_cfg* cfg) { return simx::loss_ws(cfg).total; }

extern "C" int simx_score_loss(const simx_loss_cfg* cfg, const float* Q, const float* C, const int32_t* pos_idx,
                               const float* teacher, float* out_losses, float* dQ, float* dC, void* workspace,
                               void* stream) {
  using namespace simx;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SIMX_CHECK(cfg != nullptr && Q && C && out_losses && dQ && dC, SIMX_ERR_INVALID, "score_loss: null argument");
  const int H = cfg->H;
  SIMX_CHECK(H > 0 && H % 8 == 0, SIMX_ERR_INVALID, "score_loss: H=%d must be a positive multiple of 8", H);
  SIMX_CHECK(cfg->q_count > 0 && cfg->q_begin >= 0 && cfg->q_begin + cfg->q_count <= cfg->GQ, SIMX_ERR_INVALID,
             "score_loss: local query range [%d,+%d) outside GQ=%d", cfg->q_begin, cfg->q_count, cfg->GQ);
  SIMX_CHECK(cfg->c_count > 0 && cfg->c_begin >= 0 && cfg->c_begin + cfg->c_count <= cfg->GC, SIMX_ERR_INVALID,
             "score_loss: local passage range [%d,+%d) outside GC=%d", cfg->c_begin, cfg->c_count, cfg->GC);
  const int local_terms = cfg->terms & (SIMX_LOSS_KL_LOCAL | SIMX_LOSS_CE_LOCAL | SIMX_LOSS_KD_LOCAL | SIMX_LOSS_NLL_LOCAL);
  SIMX_CHECK(cfg->terms != 0 && (cfg->terms & ~31) == 0, SIMX_ERR_INVALID, "score_loss: bad terms mask %d", cfg->terms);
  if (local_terms) {
    SIMX_CHECK(cfg->group >= 1 && cfg->group <= MAX_GROUP && cfg->c_count == cfg->q_count * cfg->group,
               SIMX_ERR_INVALID, "score_loss: local terms need c_count == q_count * group, group <= %d", MAX_GROUP);
    if (cfg->terms & (SIMX_LOSS_KL_LOCAL | SIMX_LOSS_CE_LOCAL | SIMX_LOSS_KD_LOCAL))
      SIMX_CHECK(teacher != nullptr && cfg->temperature > 0.f, SIMX_ERR_INVALID,
                 "score_loss: distillation terms need teacher logits and a positive temperature");
  }
  SIMX_CUDA(cudaMemsetAsync(out_losses, 0, 8 * sizeof(float), st));
  bool have_grads = false;
  if (cfg->terms & SIMX_LOSS_NLL_GLOBAL) {
    SIMX_CHECK(pos_idx != nullptr && workspace != nullptr, SIMX_ERR_INVALID, "score_loss: NLL_GLOBAL needs pos_idx and workspace");
    SIMX_CHECK(cfg->c_begin % 8 == 0, SIMX_ERR_INVALID,
               "score_loss: c_begin=%d must be a multiple of 8 (ranks contribute multiples of 8 passages)", cfg->c_begin);
    const LossWs w = loss_ws(cfg);
    uint8_t* ws = static_cast<uint8_t*>(workspace);
    ws += (256 - (reinterpret_cast<uintptr_t>(ws) & 255)) & 255;
    bf16* Qp = reinterpret_cast<bf16*>(ws + w.qp);
    bf16* Cp = reinterpret_cast<bf16*>(ws + w.cp);
    bf16* CTp = reinterpret_cast<bf16*>(ws + w.ctp);
    float* S = reinterpret_cast<float*>(ws + w.s);
    bf16* dS = reinterpret_cast<bf16*>(ws + w.ds);
    const int GQ = cfg->GQ, GC = cfg->GC, GCp = (GC + 7) / 8 * 8;
    const long q_ps = (long)GQ * H, c_ps = (long)GCp * H, s_ps = (long)GQ * GCp;
    if (GCp != GC) {  // zero the padding rows / columns once: they contribute nothing to any contraction
      SIMX_CUDA(cudaMemsetAsync(Cp, 0, 3ull * GCp * H * 2, st));
      SIMX_CUDA(cudaMemsetAsync(CTp, 0, 3ull * GCp * H * 2, st));
      SIMX_CUDA(cudaMemsetAsync(dS, 0, 3ull * GQ * GCp * 2, st));
    }
    SIMX_TRY(convert_weight(3, Q, GQ, H, Qp, H, q_ps, nullptr, 0, 0, st));
    SIMX_TRY(convert_weight(3, C, GC, H, Cp, H, c_ps, CTp, GCp, c_ps, st));
    SIMX_CUDA(cudaMemsetAsync(S, 0, 4ull * GQ * GCp, st));
    SIMX_TRY(gemm(GQ, GCp, H, Qp, H, q_ps, false, Cp, H, c_ps, false, 3, SIMX_EPI_ATOMIC_F32, nullptr, nullptr, 0, 0,
                  nullptr, nullptr, 0, 0, S, GCp, 1, st));
    nll_row_kernel<<<GQ, 256, 0, st>>>(S, GQ, GC, GCp, pos_idx, cfg->w_nll_global, cfg->grad_scale, dS, s_ps, out_losses);
    SIMX_LAUNCH_CHECK();
    SIMX_CUDA(cudaMemsetAsync(dQ, 0, 4ull * cfg->q_count * H, st));
    SIMX_CUDA(cudaMemsetAsync(dC, 0, 4ull * cfg->c_count * H, st));
    const int sms = num_sms();
    // dQ_local = dS[local rows, :] C          (K-major x K-major against C^T)
    {
      const int tiles = ((cfg->q_count + 127) / 128) * ((H + 255) / 256);
      const int ksplit = max(1, min((GCp + 63) / 64, (sms + tiles - 1) / tiles));
      SIMX_TRY(gemm(cfg->q_count, H, GCp, dS + (long)cfg->q_begin * GCp, GCp, s_ps, false, CTp, GCp, c_ps, false, 3,
                    SIMX_EPI_ATOMIC_F32, nullptr, nullptr, 0, 0, nullptr, nullptr, 0, 0, dQ, H, ksplit, st));
    }
    // dC_local = dS[:, local cols]^T Q        (MN-major x MN-major, contraction over the global queries)
    {
      const int tiles = ((cfg->c_count + 127) / 128) * ((H + 255) / 256);
      const int ksplit = max(1, min((GQ + 63) / 64, (sms + tiles - 1) / tiles));
      SIMX_TRY(gemm(cfg->c_count, H, GQ, dS + cfg->c_begin, GCp, s_ps, true, Qp, H, q_ps, true, 3, SIMX_EPI_ATOMIC_F32,
                    nullptr, nullptr, 0, 0, nullptr, nullptr, 0, 0, dC, H, ksplit, st));
    }
    have_grads = true;
  }
  if (local_terms) {
    const int grid = (cfg->q_count + 3) / 4;
    local_loss_kernel<<<grid, 128, 0, st>>>(*cfg, Q, C, teacher, out_losses, dQ, dC, have_grads ? 1 : 0);
    SIMX_LAUNCH_CHECK();
  }
  return SIMX_OK;
}

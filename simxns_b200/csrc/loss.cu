// Score / loss / gradient step on the gathered [CLS] embeddings (C ABI: simx_score_loss).
//
//  * local terms (SimANS KL, CE-to-teacher, PROD KD_softmax, local NLL) live on the query's own [1+n] block:
//    ONE fused kernel, a warp per query — scores, softmax, every requested loss term, dL/ds, dQ and dC
//    (co_training_marco_train.py:199-214; PROD/ProD_KD/model/models.py:669-781).
//  * the global in-batch InfoNCE (BiEncoderNllLoss.calc on all-gathered embeddings, SimANS/model/models.py:468-505)
//    runs S = Q C^T, dQ = dS C and dC = dS^T Q on the tcgen05 GEMM in fp32-class (3-plane) precision, with one
//    row kernel in between doing log-sum-exp, NLL, arg-max and dS = (softmax - onehot) / GQ in place.
// Gradients reach only the local rows/columns; remote slices are constants (LEAD/util.py:200-204).
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace simx {

int convert_weight(int np, const float* w, int R, int C, bf16* dst, long ld_dst, long ps_dst, bf16* dst_t, long ld_t,
                   long ps_t, cudaStream_t st);
// loss_fused.cu: the global InfoNCE as one cooperative launch (no score matrix in memory)
size_t nll_fused_workspace_bytes(int GQ, int GC);
int nll_global_fused(const simx_loss_cfg* cfg, const float* Q, const float* C, const int32_t* pos_idx, float* out_losses,
                     float* dQ, float* dC, void* workspace, cudaStream_t st);
// SIMX_LOSS_UNFUSED=1: the round-1 path (tcgen05 score GEMM in split-bf16 planes, S and dS in memory, two gradient GEMMs) —
// kept as the checker of the fused kernel
static bool use_unfused_global_loss() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SIMX_LOSS_UNFUSED");
    v = (e != nullptr && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

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
  const bool pos_first = ds[0] >= smax;  // positive (column 0) ranked first; torch.max returns the first maximal index
  float ssum = 0.f, ssumT = 0.f, tsum = 0.f;
  for (int j = 0; j < G; ++j) {
    ssum += expf(ds[j] - smax);
    ssumT += expf((ds[j] - smax) / T);
    if (teacher) tsum += expf(teacher[(long)i * G + j] / T - tmax);
  }
  const float lse = smax + logf(ssum);
  const float lseT = smax / T + logf(ssumT);
  const float tlse = teacher ? tmax + logf(tsum) : 0.f;
  float l_kl = 0.f, l_ce = 0.f, l_kd = 0.f, l_nll = 0.f, l_adv = 0.f, rsum = 0.f, rho_sum = 0.f;
  const bool want_klce = (cfg.terms & (SIMX_LOSS_KL_LOCAL | SIMX_LOSS_CE_LOCAL)) != 0;
  const bool want_adv = (cfg.terms & SIMX_LOSS_ADV_LOCAL) != 0;
  const float t_raw0 = teacher ? teacher[(long)i * G] : 0.f;
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
      if (want_adv) {
        // reward_j = log(softmax([t_0, t_j])_0 + eps) on the RAW teacher logits (co_training_wiki_train.py:212-217)
        const float reward = logf(1.f / (1.f + expf(teacher[(long)i * G + j] - t_raw0)) + cfg.eps);
        l_adv += reward * logf(p + cfg.eps);
        rho_sum += reward * p / (p + cfg.eps);
      }
    }
  }
  if (cfg.terms & SIMX_LOSS_NLL_LOCAL) l_nll = -(ds[0] - lse);
  // PROD KD_logit (0.5 MSE on raw logits) and DKD (decoupled KD at temperature T, target = column 0):
  //   TCKD = KL((pt_t, pt_o) || (ps_t, ps_o)),  NCKD = KL(pt / pt_o || ps / ps_o) over the non-target columns
  const bool want_mse = (cfg.terms & SIMX_LOSS_MSE_LOCAL) != 0 && teacher != nullptr;
  const bool want_dkd = (cfg.terms & SIMX_LOSS_DKD_LOCAL) != 0 && teacher != nullptr && G > 1;
  float l_mse = 0.f, l_dkd = 0.f, ps_t = 0.f, pt_t = 0.f, ps_o = 0.f, pt_o = 0.f;
  if (want_mse)
    for (int j = 0; j < G; ++j) {
      const float d = ds[j] - teacher[(long)i * G + j];
      l_mse += 0.5f * d * d;
    }
  if (want_dkd) {
    ps_t = expf(ds[0] / T - lseT);
    pt_t = expf(teacher[(long)i * G] / T - tlse);
    for (int j = 1; j < G; ++j) {
      ps_o += expf(ds[j] / T - lseT);
      pt_o += expf(teacher[(long)i * G + j] / T - tlse);
    }
    float tck = (pt_t > 0.f ? pt_t * (logf(pt_t) - logf(ps_t)) : 0.f) + (pt_o > 0.f ? pt_o * (logf(pt_o) - logf(ps_o)) : 0.f);
    float nck = 0.f;
    for (int j = 1; j < G; ++j) {
      const float ltj = teacher[(long)i * G + j] / T - tlse, lsj = ds[j] / T - lseT;
      const float t2 = expf(ltj) / pt_o;
      if (t2 > 0.f) nck += t2 * ((ltj - logf(pt_o)) - (lsj - logf(ps_o)));
    }
    l_dkd = cfg.dkd_alpha * tck + cfg.dkd_beta * nck;
  }
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
      if (want_adv) {
        const float reward = logf(1.f / (1.f + expf(teacher[(long)i * G + j] - t_raw0)) + cfg.eps);
        g += cfg.w_adv_local * (reward * p / (p + cfg.eps) - p * rho_sum);  // adv_loss is a SUM over the batch
      }
    }
    if (cfg.terms & SIMX_LOSS_NLL_LOCAL) g += cfg.w_nll_local * invB * (p - (j == 0 ? 1.f : 0.f));
    if (want_mse) g += cfg.w_mse_local * (s - teacher[(long)i * G + j]) * invB / G;
    if (want_dkd) {
      const float psj = expf(s / T - lseT), ptj = expf(teacher[(long)i * G + j] / T - tlse);
      const float g_t = j == 0 ? ps_t - pt_t : psj * (pt_t - pt_o * ps_t / ps_o);
      const float g_n = j == 0 ? 0.f : psj / ps_o - ptj / pt_o;
      g += cfg.w_dkd_local * (T * invB) * (cfg.dkd_alpha * g_t + cfg.dkd_beta * g_n);
    }
    ds[j] = g * cfg.score_scale * cfg.grad_scale;
  }
  __syncwarp();
  if (lane == 0) {
    const float T2 = T * T;
    if (pos_first) atomicAdd(out_losses + 7, 1.f);
    if (cfg.terms & SIMX_LOSS_KL_LOCAL) atomicAdd(out_losses + 2, l_kl * invB);
    if (cfg.terms & SIMX_LOSS_CE_LOCAL) atomicAdd(out_losses + 3, l_ce * invB);
    if (cfg.terms & SIMX_LOSS_KD_LOCAL) atomicAdd(out_losses + 4, l_kd * invB * T2);
    if (cfg.terms & SIMX_LOSS_NLL_LOCAL) atomicAdd(out_losses + 5, l_nll * invB);
    float tot = 0.f;
    if (cfg.terms & SIMX_LOSS_KL_LOCAL) tot += cfg.w_kl_local * l_kl * invB;
    if (cfg.terms & SIMX_LOSS_CE_LOCAL) tot += cfg.w_ce_local * l_ce * invB;
    if (cfg.terms & SIMX_LOSS_KD_LOCAL) tot += cfg.w_kd_local * l_kd * invB * T2;
    if (cfg.terms & SIMX_LOSS_NLL_LOCAL) tot += cfg.w_nll_local * l_nll * invB;
    if (want_adv) tot += cfg.w_adv_local * l_adv;
    if (want_mse) {
      atomicAdd(out_losses + 8, l_mse * invB / G);
      tot += cfg.w_mse_local * l_mse * invB / G;
    }
    if (want_dkd) {
      atomicAdd(out_losses + 9, l_dkd * invB * T2);
      tot += cfg.w_dkd_local * l_dkd * invB * T2;
    }
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
    if (!use_unfused_global_loss() && !(c->terms & SIMX_LOSS_GLOBAL_UNFUSED))
      off = nll_fused_workspace_bytes(c->GQ, c->GC);  // the fused kernel's small tables only
  }
  w.total = off + 256;
  return w;
}

// ------------------------------------------------------------------------------------------------
// LEAD-style mutual distillation between two score matrices A (student) and B (teacher), both [GQ, GC] fp32
// (LEAD/util.py:594-644 scores, :737-760 loss assembly, no .detach() anywhere — targets carry gradient):
//   nll_a = NLL(log_softmax(A), pos)               nll_b = NLL(log_softmax(B), pos)
//   kl_b2a = KL_batchmean(softmax(B/T) || softmax(A))    (teacher targets, student log-probs at temperature 1)
//   kl_a2b = KL_batchmean(softmax(A/T) || softmax(B))
// One block per row; three sweeps over the two L2-resident rows: log-sum-exps, KL row sums, gradients.
// dA/dB are written as 3 bf16 planes for the gradient GEMMs.
// ------------------------------------------------------------------------------------------------
struct PairW {
  float nll_a, nll_b, kl_b2a, kl_a2b, T, grad_scale, kd_b2a;
  float mse_b2a, dkd_b2a, alpha, beta;
};

__device__ __forceinline__ void block_lse(const float* s, int n, float invT, float* sh, float& lse_out) {
  // log-sum-exp of s[0..n) * invT over the block (256 threads); sh: >= 16 floats
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float m = -INFINITY, sum = 0.f;
  for (int c = threadIdx.x; c < n; c += 256) {
    const float v = s[c] * invT;
    if (v > m) { sum = sum * __expf(m - v) + 1.f; m = v; }
    else sum += __expf(v - m);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, sum, o);
    if (m2 > -INFINITY) {
      const float mn = fmaxf(m, m2);
      sum = (m > -INFINITY ? sum * __expf(m - mn) : 0.f) + s2 * __expf(m2 - mn);
      m = mn;
    }
  }
  __syncthreads();
  if (lane == 0) { sh[warp] = m; sh[8 + warp] = sum; }
  __syncthreads();
  float M = sh[0], Ssum = sh[8];
  for (int k = 1; k < 8; ++k) {
    if (!(sh[k] > -INFINITY)) continue;
    const float mn = fmaxf(M, sh[k]);
    Ssum = (M > -INFINITY ? Ssum * __expf(M - mn) : 0.f) + sh[8 + k] * __expf(sh[k] - mn);
    M = mn;
  }
  lse_out = M + logf(Ssum);
}
__device__ __forceinline__ float block_sum(float v, float* sh) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  for (int k = 0; k < 8; ++k) t += sh[k];
  return t;
}

__global__ void __launch_bounds__(256)
pair_row_kernel(const float* __restrict__ A, const float* __restrict__ B, int GQ, int GC, int ld,
                const int32_t* __restrict__ pos_idx, PairW w, bf16* dA, bf16* dB, long ps, float* out_losses) {
  __shared__ float sh[16];
  const int row = blockIdx.x;
  const float* a = A + (long)row * ld;
  const float* b = B + (long)row * ld;
  const float invT = 1.f / w.T;
  float lse_a, lse_b, lse_aT, lse_bT;
  block_lse(a, GC, 1.f, sh, lse_a);
  block_lse(b, GC, 1.f, sh, lse_b);
  block_lse(a, GC, invT, sh, lse_aT);
  block_lse(b, GC, invT, sh, lse_bT);
  // KL row sums
  float k_b2a = 0.f, k_a2b = 0.f, k_kd = 0.f;
  for (int c = threadIdx.x; c < GC; c += 256) {
    const float lta = a[c] * invT - lse_aT, ltb = b[c] * invT - lse_bT;  // log softmax(./T)
    const float lsa = a[c] - lse_a, lsb = b[c] - lse_b;
    const float tb = __expf(ltb);
    k_b2a += tb * (ltb - lsa);
    k_a2b += __expf(lta) * (lta - lsb);
    k_kd += tb * (ltb - lta);  // PROD KD_softmax: both sides at temperature T
  }
  k_b2a = block_sum(k_b2a, sh);
  k_a2b = block_sum(k_a2b, sh);
  k_kd = block_sum(k_kd, sh);
  const int pos = pos_idx[row];
  // PROD KD_logit / DKD on the global row (B is a constant for these: no share of gb)
  float r_mse = 0.f, ps_t = 0.f, pt_t = 0.f, ps_o = 1.f, pt_o = 1.f, l_dkd = 0.f;
  if (w.mse_b2a != 0.f) {
    for (int c = threadIdx.x; c < GC; c += 256) {
      const float d = a[c] - b[c];
      r_mse += 0.5f * d * d;
    }
    r_mse = block_sum(r_mse, sh);
  }
  if (w.dkd_b2a != 0.f) {
    float so = 0.f, to = 0.f, kx = 0.f;
    for (int c = threadIdx.x; c < GC; c += 256) {
      if (c == pos) continue;
      const float lta = a[c] * invT - lse_aT, ltb = b[c] * invT - lse_bT;
      const float tb = __expf(ltb);
      so += __expf(lta);
      to += tb;
      kx += tb * (ltb - lta);
    }
    ps_o = block_sum(so, sh);
    pt_o = block_sum(to, sh);
    kx = block_sum(kx, sh);
    ps_t = __expf(a[pos] * invT - lse_aT);
    pt_t = __expf(b[pos] * invT - lse_bT);
    const float tck = (pt_t > 0.f ? pt_t * (logf(pt_t) - logf(ps_t)) : 0.f) + (pt_o > 0.f ? pt_o * (logf(pt_o) - logf(ps_o)) : 0.f);
    const float nck = pt_o > 0.f ? kx / pt_o - logf(pt_o) + logf(ps_o) : 0.f;
    l_dkd = w.alpha * tck + w.beta * nck;
  }
  if (threadIdx.x == 0) {
    const float na = lse_a - a[pos], nb = lse_b - b[pos], r = 1.f / GQ;
    if (w.mse_b2a != 0.f || w.dkd_b2a != 0.f) {
      const float mse = r_mse * r / GC, dk = l_dkd * w.T * w.T * r;
      atomicAdd(out_losses + 7, (w.mse_b2a != 0.f ? mse : 0.f) + (w.dkd_b2a != 0.f ? dk : 0.f));
      atomicAdd(out_losses + 0, w.mse_b2a * mse + w.dkd_b2a * dk);
    }
    atomicAdd(out_losses + 1, na * r);
    atomicAdd(out_losses + 2, nb * r);
    atomicAdd(out_losses + 3, k_b2a * r);
    atomicAdd(out_losses + 4, k_a2b * r);
    atomicAdd(out_losses + 5, k_kd * w.T * w.T * r);
    atomicAdd(out_losses + 0, (w.nll_a * na + w.nll_b * nb + w.kl_b2a * k_b2a + w.kl_a2b * k_a2b +
                               w.kd_b2a * w.T * w.T * k_kd) * r);
  }
  {  // arg-max of row A == positive ?  (correct_predictions_count of the reference loss classes)
    float bv = -INFINITY;
    int bi = 0;
    for (int c = threadIdx.x; c < GC; c += 256)
      if (a[c] > bv) { bv = a[c]; bi = c; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float v2 = __shfl_xor_sync(0xffffffffu, bv, o);
      const int i2 = __shfl_xor_sync(0xffffffffu, bi, o);
      if (v2 > bv || (v2 == bv && i2 < bi)) { bv = v2; bi = i2; }
    }
    __shared__ float sv[8];
    __shared__ int si[8];
    __syncthreads();
    if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = bv; si[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int k = 1; k < 8; ++k)
        if (sv[k] > sv[0] || (sv[k] == sv[0] && si[k] < si[0])) { sv[0] = sv[k]; si[0] = si[k]; }
      if (si[0] == pos) atomicAdd(out_losses + 6, 1.f);
    }
  }
  const float r = w.grad_scale / GQ;
  for (int c = threadIdx.x; c < GC; c += 256) {
    const float lta = a[c] * invT - lse_aT, ltb = b[c] * invT - lse_bT;
    const float lsa = a[c] - lse_a, lsb = b[c] - lse_b;
    const float pa = __expf(lsa), pb = __expf(lsb), ta = __expf(lta), tb = __expf(ltb);
    const float oh = c == pos ? 1.f : 0.f;
    float ga = r * (w.nll_a * (pa - oh) + w.kl_b2a * (pa - tb) + w.kl_a2b * invT * ta * ((lta - lsb) - k_a2b) +
                    w.kd_b2a * w.T * (ta - tb));
    if (w.mse_b2a != 0.f) ga += r * w.mse_b2a * (a[c] - b[c]) / GC;
    if (w.dkd_b2a != 0.f) {
      const float g_t = c == pos ? ps_t - pt_t : ta * (pt_t - pt_o * ps_t / ps_o);
      const float g_n = c == pos ? 0.f : ta / ps_o - tb / pt_o;
      ga += r * w.dkd_b2a * w.T * (w.alpha * g_t + w.beta * g_n);
    }
    const float gb = r * (w.nll_b * (pb - oh) + w.kl_a2b * (pb - ta) + w.kl_b2a * invT * tb * ((ltb - lsa) - k_b2a) +
                          w.kd_b2a * w.T * tb * ((ltb - lta) - k_kd));
    split_store<3>(dA + (long)row * ld + c, ps, ga);
    if (dB != nullptr) split_store<3>(dB + (long)row * ld + c, ps, gb);
  }
}

// ------------------------------------------------------------------------------------------------
// LEAD cross-encoder teacher variant (--distill_ce, LEAD/util.py:669-717, 801-829): the same four terms as
// pair_row_kernel but on the query's own [1+n] block (cross-batch negatives removed, util.py:704), a = student scores
// q_i . c_ij, b = the teacher's relevance logits [B, 1+n], positive = column 0.  One warp per query; also returns
// d(total)/d(teacher logits) for --train_ce.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
local_pair_kernel(int q_count, int G, int H, const float* __restrict__ Q, const float* __restrict__ C,
                  const float* __restrict__ teacher, PairW w, float* out_losses, float* dQ, float* dC, float* dT) {
  __shared__ float sa[4][MAX_GROUP], sg[4][MAX_GROUP];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * 4 + warp;
  if (i >= q_count) return;
  const float* q = Q + (long)i * H;
  const float* c = C + (long)i * G * H;
  const float* b = teacher + (long)i * G;
  float* a = sa[warp];
  float* g = sg[warp];
  for (int j = 0; j < G; ++j) {
    float acc = 0.f;
    for (int h = lane * 4; h < H; h += 128) {
      const float4 x = __ldg(reinterpret_cast<const float4*>(q + h));
      const float4 y = __ldg(reinterpret_cast<const float4*>(c + (long)j * H + h));
      acc += x.x * y.x + x.y * y.y + x.z * y.z + x.w * y.w;
    }
    acc = warp_sum(acc);
    if (lane == 0) a[j] = acc;
  }
  __syncwarp();
  const float invT = 1.f / w.T, r = 1.f / q_count;
  float ma = -INFINITY, mb = -INFINITY;
  for (int j = 0; j < G; ++j) { ma = fmaxf(ma, a[j]); mb = fmaxf(mb, b[j]); }
  float sa1 = 0.f, sb1 = 0.f, saT = 0.f, sbT = 0.f;
  for (int j = 0; j < G; ++j) {
    sa1 += expf(a[j] - ma); sb1 += expf(b[j] - mb);
    saT += expf((a[j] - ma) * invT); sbT += expf((b[j] - mb) * invT);
  }
  const float lse_a = ma + logf(sa1), lse_b = mb + logf(sb1), lse_aT = ma * invT + logf(saT), lse_bT = mb * invT + logf(sbT);
  float k_b2a = 0.f, k_a2b = 0.f;
  for (int j = 0; j < G; ++j) {
    const float lta = a[j] * invT - lse_aT, ltb = b[j] * invT - lse_bT;
    k_b2a += expf(ltb) * (ltb - (a[j] - lse_a));
    k_a2b += expf(lta) * (lta - (b[j] - lse_b));
  }
  if (lane == 0) {
    const float na = lse_a - a[0], nb = lse_b - b[0];
    atomicAdd(out_losses + 1, na * r);
    atomicAdd(out_losses + 2, nb * r);
    atomicAdd(out_losses + 3, k_b2a * r);
    atomicAdd(out_losses + 4, k_a2b * r);
    atomicAdd(out_losses + 0, (w.nll_a * na + w.nll_b * nb + w.kl_b2a * k_b2a + w.kl_a2b * k_a2b) * r);
    if (a[0] >= ma) atomicAdd(out_losses + 6, 1.f);
  }
  const float rs = r * w.grad_scale;
  for (int j = lane; j < G; j += 32) {
    const float lta = a[j] * invT - lse_aT, ltb = b[j] * invT - lse_bT, lsa = a[j] - lse_a, lsb = b[j] - lse_b;
    const float pa = expf(lsa), pb = expf(lsb), ta = expf(lta), tb = expf(ltb), oh = j == 0 ? 1.f : 0.f;
    g[j] = rs * (w.nll_a * (pa - oh) + w.kl_b2a * (pa - tb) + w.kl_a2b * invT * ta * ((lta - lsb) - k_a2b));
    if (dT != nullptr)
      dT[(long)i * G + j] = rs * (w.nll_b * (pb - oh) + w.kl_a2b * (pb - ta) + w.kl_b2a * invT * tb * ((ltb - lsa) - k_b2a));
  }
  __syncwarp();
  float* dq = dQ + (long)i * H;
  float* dc = dC + (long)i * G * H;
  for (int h = lane * 4; h < H; h += 128) {
    const float4 x = __ldg(reinterpret_cast<const float4*>(q + h));
    float4 accq = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = 0; j < G; ++j) {
      const float gj = g[j];
      const float4 y = __ldg(reinterpret_cast<const float4*>(c + (long)j * H + h));
      accq.x += gj * y.x; accq.y += gj * y.y; accq.z += gj * y.z; accq.w += gj * y.w;
      *reinterpret_cast<float4*>(dc + (long)j * H + h) = make_float4(gj * x.x, gj * x.y, gj * x.z, gj * x.w);
    }
    *reinterpret_cast<float4*>(dq + h) = accq;
  }
}

// one half of PROD's Bi_logit: weight * 0.5 * mean((a - b)^2) and its gradient w.r.t. a
__global__ void __launch_bounds__(256) embed_mse_kernel(const float* __restrict__ a, const float* __restrict__ b, long n,
                                                        float weight, float* loss, float* grad) {
  __shared__ float sh[16];
  float acc = 0.f;
  const float k = weight / (float)n;
  for (long i = blockIdx.x * 256L + threadIdx.x; i < n; i += (long)gridDim.x * 256L) {
    const float d = a[i] - b[i];
    acc += d * d;
    grad[i] = k * d;
  }
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0 && acc != 0.f) atomicAdd(loss, 0.5f * k * acc);
}

// shared pieces of the global-matrix paths ---------------------------------------------------------
struct ScoreWs {  // per embedding pair (Q, C): planes, transposed planes, scores, score gradients
  size_t qp, cp, ctp, s, ds, total;
};
static ScoreWs score_ws(int GQ, int GC, int H) {
  ScoreWs w{};
  const size_t GCp = (GC + 7) / 8 * 8;
  size_t off = 0;
  w.qp = off;  off += align256(3ull * GQ * H * 2);
  w.cp = off;  off += align256(3ull * GCp * H * 2);
  w.ctp = off; off += align256(3ull * GCp * H * 2);
  w.s = off;   off += align256(4ull * GQ * GCp);
  w.ds = off;  off += align256(3ull * GQ * GCp * 2);
  w.total = off;
  return w;
}
// S = Q C^T in fp32-class precision (planes of Q, C and C^T stay in the workspace for the gradient GEMMs)
static int score_forward(const float* Q, const float* C, int GQ, int GC, int H, uint8_t* ws, const ScoreWs& w,
                         cudaStream_t st) {
  const int GCp = (GC + 7) / 8 * 8;
  bf16* Qp = reinterpret_cast<bf16*>(ws + w.qp);
  bf16* Cp = reinterpret_cast<bf16*>(ws + w.cp);
  bf16* CTp = reinterpret_cast<bf16*>(ws + w.ctp);
  float* S = reinterpret_cast<float*>(ws + w.s);
  const long q_ps = (long)GQ * H, c_ps = (long)GCp * H;
  if (GCp != GC) {
    SIMX_CUDA(cudaMemsetAsync(Cp, 0, 3ull * GCp * H * 2, st));
    SIMX_CUDA(cudaMemsetAsync(CTp, 0, 3ull * GCp * H * 2, st));
    SIMX_CUDA(cudaMemsetAsync(ws + w.ds, 0, 3ull * GQ * GCp * 2, st));
  }
  SIMX_TRY(convert_weight(3, Q, GQ, H, Qp, H, q_ps, nullptr, 0, 0, st));
  SIMX_TRY(convert_weight(3, C, GC, H, Cp, H, c_ps, CTp, GCp, c_ps, st));
  SIMX_CUDA(cudaMemsetAsync(S, 0, 4ull * GQ * GCp, st));
  return gemm(GQ, GCp, H, Qp, H, q_ps, false, Cp, H, c_ps, false, 3, SIMX_EPI_ATOMIC_F32, nullptr, nullptr, 0, 0, nullptr,
              nullptr, 0, 0, S, GCp, 1, st);
}
// dQ_local = dS[local rows] C ;  dC_local = dS[:, local cols]^T Q     (both overwritten)
static int score_backward(int GQ, int GC, int H, int q_begin, int q_count, int c_begin, int c_count, uint8_t* ws,
                          const ScoreWs& w, float* dQ, float* dC, cudaStream_t st) {
  const int GCp = (GC + 7) / 8 * 8;
  bf16* Qp = reinterpret_cast<bf16*>(ws + w.qp);
  bf16* CTp = reinterpret_cast<bf16*>(ws + w.ctp);
  bf16* dS = reinterpret_cast<bf16*>(ws + w.ds);
  const long q_ps = (long)GQ * H, c_ps = (long)GCp * H, s_ps = (long)GQ * GCp;
  SIMX_CUDA(cudaMemsetAsync(dQ, 0, 4ull * q_count * H, st));
  SIMX_CUDA(cudaMemsetAsync(dC, 0, 4ull * c_count * H, st));
  const int sms = num_sms();
  {
    const int tiles = ((q_count + 127) / 128) * ((H + 255) / 256);
    const int ksplit = max(1, min((GCp + 63) / 64, (sms + tiles - 1) / tiles));
    SIMX_TRY(gemm(q_count, H, GCp, dS + (long)q_begin * GCp, GCp, s_ps, false, CTp, GCp, c_ps, false, 3,
                  SIMX_EPI_ATOMIC_F32, nullptr, nullptr, 0, 0, nullptr, nullptr, 0, 0, dQ, H, ksplit, st));
  }
  {
    const int tiles = ((c_count + 127) / 128) * ((H + 255) / 256);
    const int ksplit = max(1, min((GQ + 63) / 64, (sms + tiles - 1) / tiles));
    SIMX_TRY(gemm(c_count, H, GQ, dS + c_begin, GCp, s_ps, true, Qp, H, q_ps, true, 3, SIMX_EPI_ATOMIC_F32, nullptr,
                  nullptr, 0, 0, nullptr, nullptr, 0, 0, dC, H, ksplit, st));
  }
  return SIMX_OK;
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
  const int local_terms = cfg->terms & (SIMX_LOSS_KL_LOCAL | SIMX_LOSS_CE_LOCAL | SIMX_LOSS_KD_LOCAL | SIMX_LOSS_NLL_LOCAL |
                                        SIMX_LOSS_ADV_LOCAL | SIMX_LOSS_MSE_LOCAL | SIMX_LOSS_DKD_LOCAL);
  SIMX_CHECK((cfg->terms & 255) != 0 && (cfg->terms & ~511) == 0, SIMX_ERR_INVALID, "score_loss: bad terms mask %d", cfg->terms);
  const bool unfused = use_unfused_global_loss() || (cfg->terms & SIMX_LOSS_GLOBAL_UNFUSED) != 0;
  if (local_terms) {
    SIMX_CHECK(cfg->group >= 1 && cfg->group <= MAX_GROUP && cfg->c_count == cfg->q_count * cfg->group,
               SIMX_ERR_INVALID, "score_loss: local terms need c_count == q_count * group, group <= %d", MAX_GROUP);
    if (cfg->terms & (SIMX_LOSS_KL_LOCAL | SIMX_LOSS_CE_LOCAL | SIMX_LOSS_KD_LOCAL | SIMX_LOSS_ADV_LOCAL | SIMX_LOSS_MSE_LOCAL |
                      SIMX_LOSS_DKD_LOCAL))
      SIMX_CHECK(teacher != nullptr && cfg->temperature > 0.f, SIMX_ERR_INVALID,
                 "score_loss: distillation terms need teacher logits and a positive temperature");
  }
  SIMX_CUDA(cudaMemsetAsync(out_losses, 0, 12 * sizeof(float), st));
  bool have_grads = false;
  if (cfg->terms & SIMX_LOSS_NLL_GLOBAL) {
    SIMX_CHECK(pos_idx != nullptr && workspace != nullptr, SIMX_ERR_INVALID, "score_loss: NLL_GLOBAL needs pos_idx and workspace");
  }
  if ((cfg->terms & SIMX_LOSS_NLL_GLOBAL) && !unfused) {
    SIMX_TRY(nll_global_fused(cfg, Q, C, pos_idx, out_losses, dQ, dC, workspace, st));
    have_grads = true;
  } else if (cfg->terms & SIMX_LOSS_NLL_GLOBAL) {
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

extern "C" size_t simx_pair_distill_workspace_bytes(int GQ, int GC, int H) {
  return 2 * simx::score_ws(GQ, GC, H).total + 512;
}

extern "C" int simx_pair_distill(const simx_pair_cfg* cfg, const float* Qa, const float* Ca, const float* Qb,
                                 const float* Cb, const int32_t* pos_idx, float* out_losses, float* dQa, float* dCa,
                                 float* dQb, float* dCb, void* workspace, void* stream) {
  using namespace simx;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SIMX_CHECK(cfg && Qa && Ca && Qb && Cb && pos_idx && out_losses && dQa && dCa && workspace, SIMX_ERR_INVALID,
             "pair_distill: null argument");
  const int GQ = cfg->GQ, GC = cfg->GC, H = cfg->H;
  SIMX_CHECK(GQ > 0 && GC > 0 && H > 0 && H % 8 == 0, SIMX_ERR_INVALID, "pair_distill: bad sizes GQ=%d GC=%d H=%d", GQ, GC, H);
  SIMX_CHECK(cfg->q_count > 0 && cfg->q_begin >= 0 && cfg->q_begin + cfg->q_count <= GQ && cfg->c_count > 0 &&
                 cfg->c_begin >= 0 && cfg->c_begin + cfg->c_count <= GC && cfg->c_begin % 8 == 0,
             SIMX_ERR_INVALID, "pair_distill: bad local ranges");
  SIMX_CHECK(cfg->temperature > 0.f, SIMX_ERR_INVALID, "pair_distill: temperature must be positive");
  const bool train_b = dQb != nullptr && dCb != nullptr;
  const ScoreWs w = score_ws(GQ, GC, H);
  uint8_t* wa = static_cast<uint8_t*>(workspace);
  wa += (256 - (reinterpret_cast<uintptr_t>(wa) & 255)) & 255;
  uint8_t* wb = wa + w.total;
  const int GCp = (GC + 7) / 8 * 8;
  SIMX_CUDA(cudaMemsetAsync(out_losses, 0, 8 * sizeof(float), st));
  SIMX_TRY(score_forward(Qa, Ca, GQ, GC, H, wa, w, st));
  SIMX_TRY(score_forward(Qb, Cb, GQ, GC, H, wb, w, st));
  PairW pw{cfg->w_nll_a, cfg->w_nll_b, cfg->w_kl_b2a, cfg->w_kl_a2b, cfg->temperature, cfg->grad_scale, cfg->w_kd_b2a,
           cfg->w_mse_b2a, cfg->w_dkd_b2a, cfg->dkd_alpha, cfg->dkd_beta};
  pair_row_kernel<<<GQ, 256, 0, st>>>(reinterpret_cast<const float*>(wa + w.s), reinterpret_cast<const float*>(wb + w.s), GQ,
                                      GC, GCp, pos_idx, pw, reinterpret_cast<bf16*>(wa + w.ds),
                                      train_b ? reinterpret_cast<bf16*>(wb + w.ds) : nullptr, (long)GQ * GCp, out_losses);
  SIMX_LAUNCH_CHECK();
  SIMX_TRY(score_backward(GQ, GC, H, cfg->q_begin, cfg->q_count, cfg->c_begin, cfg->c_count, wa, w, dQa, dCa, st));
  if (train_b) SIMX_TRY(score_backward(GQ, GC, H, cfg->q_begin, cfg->q_count, cfg->c_begin, cfg->c_count, wb, w, dQb, dCb, st));
  return SIMX_OK;
}

extern "C" int simx_embed_mse(const float* a, const float* b, int64_t n, float weight, float* loss, float* grad,
                              void* stream) {
  using namespace simx;
  SIMX_CHECK(a && b && loss && grad && n > 0, SIMX_ERR_INVALID, "embed_mse: null argument or empty tensor");
  const int grid = (int)min((n + 255) / 256, (int64_t)num_sms() * 4);
  embed_mse_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(a, b, (long)n, weight, loss, grad);
  SIMX_LAUNCH_CHECK();
  return SIMX_OK;
}

extern "C" int simx_local_pair_distill(const simx_pair_cfg* cfg, const float* Q_local, const float* C_local,
                                       const float* teacher_logits, float* out_losses, float* dQ, float* dC, float* dT,
                                       void* stream) {
  using namespace simx;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SIMX_CHECK(cfg && Q_local && C_local && teacher_logits && out_losses && dQ && dC, SIMX_ERR_INVALID,
             "local_pair_distill: null argument");
  SIMX_CHECK(cfg->q_count > 0 && cfg->c_count > 0 && cfg->c_count % cfg->q_count == 0 &&
                 cfg->c_count / cfg->q_count <= MAX_GROUP, SIMX_ERR_INVALID,
             "local_pair_distill: c_count must be q_count * group with group <= %d", MAX_GROUP);
  SIMX_CHECK(cfg->H > 0 && cfg->H % 4 == 0 && cfg->temperature > 0.f, SIMX_ERR_INVALID, "local_pair_distill: bad H / temperature");
  SIMX_CUDA(cudaMemsetAsync(out_losses, 0, 8 * sizeof(float), st));
  PairW pw{cfg->w_nll_a, cfg->w_nll_b, cfg->w_kl_b2a, cfg->w_kl_a2b, cfg->temperature, cfg->grad_scale, 0.f, 0.f, 0.f, 0.f, 0.f};
  local_pair_kernel<<<(cfg->q_count + 3) / 4, 128, 0, st>>>(cfg->q_count, cfg->c_count / cfg->q_count, cfg->H, Q_local, C_local,
                                                            teacher_logits, pw, out_losses, dQ, dC, dT);
  SIMX_LAUNCH_CHECK();
  return SIMX_OK;
}


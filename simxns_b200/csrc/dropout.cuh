// Counter-based dropout masks (Philox4x32-7) shared by every kernel that applies or re-applies dropout.
//
// The reference trains every tower with nn.Dropout(0.1) at three kinds of sites (SimANS/model/models.py:66-72 sets
// hidden_dropout_prob / attention_probs_dropout_prob; LEAD/modeling_bert.py:239 embeddings, :358 attention probabilities,
// :386 / :464 the hidden state after BertSelfOutput.dense / BertOutput.dense; DistilBERT: LEAD/modeling_distilbert.py:133,
// :221, :254).  torch's Philox stream cannot be reproduced (it depends on torch's kernel tiling), so parity is by
// distribution, and by exact agreement with oracle/dropout_numpy.py, which restates THIS generator.
//
// A mask bit is a pure function of (seed, site, layer, row, column): no mask is ever stored.  Backward regenerates it, and
// kernels with different tilings (tcgen05 / mma.sync / FFMA attention, GEMM epilogue / row kernels) agree by construction.
//   one Philox call -> 128 bits -> 8 x 16-bit uniforms for the 8 consecutive columns [8*c0, 8*c0+8) of one row:
//   counter = (c0 = column >> 3, c1 = row, c2 = plane (attention: b*heads + h, else 0), c3 = site | layer << 8),
//   key = (seed_lo, seed_hi); column e of the group uses word e >> 1, low half for even e, high half for odd e;
//   an element is DROPPED iff its 16-bit uniform < thr = round(p * 65536); survivors are scaled by 65536 / (65536 - thr)
//   (the exact inverse keep probability of the quantised threshold, so the mask is unbiased).
#pragma once
#include <stdint.h>

namespace simx {

enum DropSite : uint32_t { DROP_SITE_EMBED = 0, DROP_SITE_ATTN = 1, DROP_SITE_ATTN_OUT = 2, DROP_SITE_FFN_OUT = 3 };

struct DropSpec {
  uint32_t thr_hi;  // thr << 16 (0 = dropout off): an element is kept iff (its 16-bit uniform << 16) >= thr_hi
  float scale;      // 1 / keep probability
  uint32_t k0, k1;  // seed
  uint32_t tag;     // site | layer << 8
};

static inline DropSpec make_drop(float p, uint64_t seed, uint32_t site, int layer) {
  DropSpec d;
  long thr = (long)(p * 65536.0 + 0.5);
  if (thr < 0) thr = 0;
  if (thr > 65535) thr = 65535;
  d.thr_hi = (uint32_t)thr << 16;
  d.scale = (float)(65536.0 / (65536.0 - (double)thr));
  d.k0 = (uint32_t)(seed & 0xffffffffu);
  d.k1 = (uint32_t)(seed >> 32);
  d.tag = site | ((uint32_t)layer << 8);
  return d;
}
static inline DropSpec no_drop() { return DropSpec{0u, 1.f, 0u, 0u, 0u}; }

#ifdef __CUDACC__
__device__ __forceinline__ uint4 philox7(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 7; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0;
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k1;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}
// the 8 uniforms of columns [8*cgroup, 8*cgroup + 8) of `row` in plane `plane`
__device__ __forceinline__ uint4 drop_bits(const DropSpec& d, uint32_t cgroup, uint32_t row, uint32_t plane) {
  return philox7(cgroup, row, plane, d.tag, d.k0, d.k1);
}
// keep-multiplier (scale or 0) of element e (0..7) of the group
__device__ __forceinline__ float drop_mult(const uint4& r, int e, uint32_t thr_hi, float scale) {
  const uint32_t w = (e >> 1) == 0 ? r.x : ((e >> 1) == 1 ? r.y : ((e >> 1) == 2 ? r.z : r.w));
  const uint32_t u = (e & 1) ? w : (w << 16);
  return u >= thr_hi ? scale : 0.f;
}
__device__ __forceinline__ void drop_apply8(float (&v)[8], const uint4& r, uint32_t thr_hi, float scale) {
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = (w[i] << 16) >= thr_hi ? v[2 * i] * scale : 0.f;
    v[2 * i + 1] = w[i] >= thr_hi ? v[2 * i + 1] * scale : 0.f;
  }
}
#endif

}  // namespace simx

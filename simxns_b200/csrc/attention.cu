// Self-attention of the encoder, head_dim = 64, any S <= 512, arbitrary HF key mask.
//   softmax(Q K^T / 8 + (1 - mask) * -inf) V          (LEAD/modeling_bert.py:324-368; DistilBERT :196-226)
//
// The attention GEMMs are ~3 % of the encoder FLOPs; what matters is that Q/K/V/ctx cross HBM once.
// One CTA owns a 128-row tile of one (sequence, head): 8 warps x 16 rows, the whole "column" operand pair of the
// sequence resident in swizzled shared memory (cp.async), 64-column chunks, mma.sync m16n8k16 bf16 -> fp32 with the
// accumulator-to-A-fragment repack, online softmax via quad shuffles and exp2f.  Three instances of the same shape:
//   forward        rows = queries, columns = keys   : O = softmax(S) V, saves log2-sum-exp per row
//   backward dQ    rows = queries, columns = keys   : dQ = (P o (dO V^T - delta)) K / 8, computes delta = rowsum(dO o O)
//   backward dK/dV rows = keys,    columns = queries: dV = P^T dO, dK = dS^T Q / 8   (S^T recomputed as K Q^T)
// (legacy tensor path: the tcgen05 version of this kernel is the next step).
// The fp32-class mode (3 planes) uses the plain FFMA kernels at the bottom of this file.
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace simx {

constexpr int ATT_D = 64;
constexpr int ATT_ROWS = 128;   // rows per CTA
constexpr int ATT_THREADS = 256;
constexpr float LOG2E = 1.4426950408889634f;
constexpr float NEG_BIG = -1.0e30f;

__device__ __forceinline__ uint32_t swz(int row, int chunk) { return row * 128 + (((chunk ^ row) & 7) << 4); }

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                         uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// [nrows_total x 64] bf16 tile <- global rows (row stride ld elements); rows >= nvalid are zero-filled
__device__ __forceinline__ void load_tile(uint32_t tile, const bf16* g, long ld, int nvalid, int nrows_total) {
  for (int idx = threadIdx.x; idx < nrows_total * 8; idx += ATT_THREADS) {
    const int row = idx >> 3, chunk = idx & 7;
    const uint32_t dst = tile + swz(row, chunk);
    if (row < nvalid) cp_async16(dst, g + (long)row * ld + chunk * 8);
    else asm volatile("st.shared.v4.u32 [%0], {%1,%1,%1,%1};" ::"r"(dst), "r"(0u) : "memory");
  }
}

// acc[j] (8 n-tiles of 8 columns) = A(rows of this warp, 64-deep) * Bcols(chunk c0.. c0+64)^T
__device__ __forceinline__ void mma_rows_cols(float (&acc)[8][4], uint32_t a_tile, uint32_t b_tile, int warp, int lane,
                                              int c0) {
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[j][i] = 0.f;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    uint32_t a0, a1, a2, a3;
    ldsm_x4(a_tile + swz(warp * 16 + (lane & 15), ks * 2 + (lane >> 4)), a0, a1, a2, a3);
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t b0, b1, b2, b3;
      ldsm_x4(b_tile + swz(c0 + np * 16 + (lane & 7) + ((lane >> 4) << 3), ks * 2 + ((lane >> 3) & 1)), b0, b1, b2, b3);
      mma16816(acc[2 * np], a0, a1, a2, a3, b0, b1);
      mma16816(acc[2 * np + 1], a0, a1, a2, a3, b2, b3);
    }
  }
}

// out[jd] (8 n-tiles over d) += P(16 x 64, fp32 accum layout, converted to bf16 A frags) * Z(rows c0..c0+64 of z_tile)
__device__ __forceinline__ void mma_p_z(float (&out)[8][4], const float (&p)[8][4], uint32_t z_tile, int lane, int c0) {
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    const uint32_t a0 = pack_bf16x2(p[2 * kk][0], p[2 * kk][1]);
    const uint32_t a1 = pack_bf16x2(p[2 * kk][2], p[2 * kk][3]);
    const uint32_t a2 = pack_bf16x2(p[2 * kk + 1][0], p[2 * kk + 1][1]);
    const uint32_t a3 = pack_bf16x2(p[2 * kk + 1][2], p[2 * kk + 1][3]);
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t b0, b1, b2, b3;
      ldsm_x4_t(z_tile + swz(c0 + kk * 16 + (lane & 7) + (((lane >> 3) & 1) << 3), np * 2 + (lane >> 4)), b0, b1, b2, b3);
      mma16816(out[2 * np], a0, a1, a2, a3, b0, b1);
      mma16816(out[2 * np + 1], a0, a1, a2, a3, b2, b3);
    }
  }
}

// write this warp's 16 x 64 fp32 fragment (scaled per row) as bf16 through its own rows of `stage_tile`, then
// coalesced 16-byte stores to global rows [row0 + warp*16, ...)
__device__ __forceinline__ void store_rows(const float (&o)[8][4], float s_lo, float s_hi, uint32_t stage_tile,
                                           uint8_t* stage_ptr, int warp, int lane, bf16* g, long ld, int row0,
                                           int nvalid) {
  const int g_ = lane >> 2, t = lane & 3;
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int r_lo = warp * 16 + g_, r_hi = r_lo + 8;
    const uint32_t lo = pack_bf16x2(o[j][0] * s_lo, o[j][1] * s_lo);
    const uint32_t hi = pack_bf16x2(o[j][2] * s_hi, o[j][3] * s_hi);
    *reinterpret_cast<uint32_t*>(stage_ptr + swz(r_lo, j) + t * 4) = lo;
    *reinterpret_cast<uint32_t*>(stage_ptr + swz(r_hi, j) + t * 4) = hi;
  }
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = warp * 16 + i * 4 + (lane >> 3), chunk = lane & 7;
    if (row0 + r < nvalid) {
      const uint4 v = *reinterpret_cast<const uint4*>(stage_ptr + swz(r, chunk));
      *reinterpret_cast<uint4*>(g + (long)(row0 + r) * ld + chunk * 8) = v;
    }
  }
}

struct AttnParams {
  const bf16* qkv;  // [T, 3H]
  bf16* ctx;        // [T, H]       (forward out / backward in: O)
  const bf16* dctx; // [T, H]
  bf16* dqkv;       // [T, 3H]
  const void* mask;  // [B, S] int64 / int32 (mask_w bytes per entry) or NULL
  float* lse2;      // [B, heads, S]  log2-domain log-sum-exp of the scaled, masked scores
  float* delta;     // [B, heads, S]
  int B, S, H, heads, s_pad;
  int mask_w;
  DropSpec drop;    // attention-probability dropout (forward kernel only; the mma.sync backward is a legacy path)
};
__device__ __forceinline__ uint32_t u4_word_rt(const uint4& r, int i) { return i == 0 ? r.x : (i == 1 ? r.y : (i == 2 ? r.z : r.w)); }

// -------------------------------------------------------------------------------------------------
// forward
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ATT_THREADS)
attn_fwd_kernel(const AttnParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int b = blockIdx.z, h = blockIdx.y, row0 = blockIdx.x * ATT_ROWS;
  const int S = p.S, SP = p.s_pad;
  uint8_t* q_ptr = smem;
  const uint32_t q_tile = smem_u32(smem);
  const uint32_t k_tile = q_tile + ATT_ROWS * 128;
  const uint32_t v_tile = k_tile + SP * 128;
  float* bias = reinterpret_cast<float*>(smem + ATT_ROWS * 128 + 2 * SP * 128);

  const long ld = 3L * p.H;
  const bf16* base = p.qkv + (long)b * S * ld + h * ATT_D;
  load_tile(q_tile, base + (long)row0 * ld, ld, min(ATT_ROWS, S - row0), ATT_ROWS);
  load_tile(k_tile, base + p.H, ld, S, SP);
  load_tile(v_tile, base + 2 * p.H, ld, S, SP);
  for (int c = threadIdx.x; c < SP; c += ATT_THREADS)
    bias[c] = (c < S && (p.mask == nullptr || ld_id(p.mask, (long)b * S + c, p.mask_w) != 0)) ? 0.f : NEG_BIG;
  cp_async_wait_all();
  __syncthreads();

  const float sl2 = 0.125f * LOG2E;
  float o[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) o[j][i] = 0.f;
  float m_lo = NEG_BIG, m_hi = NEG_BIG, l_lo = 0.f, l_hi = 0.f;

  if (row0 + warp * 16 < S) {  // warps whose rows are all padding skip the math (they still hit the barriers above)
    for (int c0 = 0; c0 < SP; c0 += 64) {
      float s[8][4];
      mma_rows_cols(s, q_tile, k_tile, warp, lane, c0);
      float mx_lo = m_lo, mx_hi = m_hi;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float b0 = bias[c0 + j * 8 + 2 * t], b1 = bias[c0 + j * 8 + 2 * t + 1];
        s[j][0] = s[j][0] * sl2 + b0; s[j][1] = s[j][1] * sl2 + b1;
        s[j][2] = s[j][2] * sl2 + b0; s[j][3] = s[j][3] * sl2 + b1;
        mx_lo = fmaxf(mx_lo, fmaxf(s[j][0], s[j][1]));
        mx_hi = fmaxf(mx_hi, fmaxf(s[j][2], s[j][3]));
      }
      mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 1));
      mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 2));
      mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 1));
      mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 2));
      const float corr_lo = exp2f(m_lo - mx_lo), corr_hi = exp2f(m_hi - mx_hi);
      m_lo = mx_lo; m_hi = mx_hi;
      float sum_lo = 0.f, sum_hi = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s[j][0] = exp2f(s[j][0] - m_lo); s[j][1] = exp2f(s[j][1] - m_lo);
        s[j][2] = exp2f(s[j][2] - m_hi); s[j][3] = exp2f(s[j][3] - m_hi);
        sum_lo += s[j][0] + s[j][1];
        sum_hi += s[j][2] + s[j][3];
      }
      l_lo = l_lo * corr_lo + sum_lo;
      l_hi = l_hi * corr_hi + sum_hi;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        o[j][0] *= corr_lo; o[j][1] *= corr_lo; o[j][2] *= corr_hi; o[j][3] *= corr_hi;
      }
      if (p.drop.thr_hi != 0u) {
        // this thread holds keys c0 + 8j + 2t, +1 of rows r_lo / r_hi: word t of the 8-key group's Philox block
        const uint32_t plane = b * p.heads + h, q_lo = row0 + warp * 16 + g;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t w_lo = u4_word_rt(drop_bits(p.drop, (c0 >> 3) + j, q_lo, plane), t);
          const uint32_t w_hi = u4_word_rt(drop_bits(p.drop, (c0 >> 3) + j, q_lo + 8, plane), t);
          if ((w_lo << 16) < p.drop.thr_hi) s[j][0] = 0.f;
          if (w_lo < p.drop.thr_hi) s[j][1] = 0.f;
          if ((w_hi << 16) < p.drop.thr_hi) s[j][2] = 0.f;
          if (w_hi < p.drop.thr_hi) s[j][3] = 0.f;
        }
      }
      mma_p_z(o, s, v_tile, lane, c0);
    }
    l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 1);
    l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 2);
    l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 1);
    l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 2);
    const int r_lo = row0 + warp * 16 + g, r_hi = r_lo + 8;
    if (p.lse2 != nullptr && t == 0) {
      float* L = p.lse2 + ((long)b * p.heads + h) * S;
      if (r_lo < S) L[r_lo] = m_lo + log2f(l_lo);
      if (r_hi < S) L[r_hi] = m_hi + log2f(l_hi);
    }
    const float ksc = p.drop.thr_hi != 0u ? p.drop.scale : 1.f;
    store_rows(o, ksc / l_lo, ksc / l_hi, q_tile, q_ptr, warp, lane, p.ctx + (long)b * S * p.H + h * ATT_D, p.H, row0, S);
  }
}

// -------------------------------------------------------------------------------------------------
// backward, dQ
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ATT_THREADS)
attn_bwd_dq_kernel(const AttnParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int b = blockIdx.z, h = blockIdx.y, row0 = blockIdx.x * ATT_ROWS;
  const int S = p.S, SP = p.s_pad;
  uint8_t* q_ptr = smem;
  const uint32_t q_tile = smem_u32(smem);
  const uint32_t do_tile = q_tile + ATT_ROWS * 128;
  const uint32_t k_tile = do_tile + ATT_ROWS * 128;
  const uint32_t v_tile = k_tile + SP * 128;
  float* bias = reinterpret_cast<float*>(smem + 2 * ATT_ROWS * 128 + 2 * SP * 128);
  float* delta_s = bias + SP;

  const long ld = 3L * p.H;
  const bf16* base = p.qkv + (long)b * S * ld + h * ATT_D;
  const bf16* dO = p.dctx + (long)b * S * p.H + h * ATT_D;
  const bf16* O = p.ctx + (long)b * S * p.H + h * ATT_D;
  const int nrows = min(ATT_ROWS, S - row0);
  load_tile(q_tile, base + (long)row0 * ld, ld, nrows, ATT_ROWS);
  load_tile(do_tile, dO + (long)row0 * p.H, p.H, nrows, ATT_ROWS);
  load_tile(k_tile, base + p.H, ld, S, SP);
  load_tile(v_tile, base + 2 * p.H, ld, S, SP);
  for (int c = threadIdx.x; c < SP; c += ATT_THREADS)
    bias[c] = (c < S && (p.mask == nullptr || ld_id(p.mask, (long)b * S + c, p.mask_w) != 0)) ? 0.f : NEG_BIG;
  {
    // delta[row] = sum_d dO[row,d] * O[row,d]: two lanes per row, 32 elements each
    const int r = warp * 16 + (lane >> 1), half = lane & 1;
    float acc = 0.f;
    if (row0 + r < S) {
      const uint4* a = reinterpret_cast<const uint4*>(dO + (long)(row0 + r) * p.H + half * 32);
      const uint4* c = reinterpret_cast<const uint4*>(O + (long)(row0 + r) * p.H + half * 32);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint4 x = a[i], y = c[i];
        const uint32_t xs[4] = {x.x, x.y, x.z, x.w}, ys[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 fx = unpack_bf16x2(xs[k]), fy = unpack_bf16x2(ys[k]);
          acc += fx.x * fy.x + fx.y * fy.y;
        }
      }
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    if (half == 0) {
      delta_s[r] = acc;
      if (row0 + r < S) p.delta[((long)b * p.heads + h) * S + row0 + r] = acc;
    }
  }
  cp_async_wait_all();
  __syncthreads();

  if (row0 + warp * 16 >= S) return;
  const float sl2 = 0.125f * LOG2E;
  const int r_lo = warp * 16 + g, r_hi = r_lo + 8;
  const float* L = p.lse2 + ((long)b * p.heads + h) * S;
  const float lse_lo = (row0 + r_lo < S) ? L[row0 + r_lo] : 0.f;
  const float lse_hi = (row0 + r_hi < S) ? L[row0 + r_hi] : 0.f;
  const float d_lo = delta_s[r_lo], d_hi = delta_s[r_hi];
  float dq[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) dq[j][i] = 0.f;

  for (int c0 = 0; c0 < SP; c0 += 64) {
    float s[8][4], dp[8][4];
    mma_rows_cols(s, q_tile, k_tile, warp, lane, c0);
    mma_rows_cols(dp, do_tile, v_tile, warp, lane, c0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float b0 = bias[c0 + j * 8 + 2 * t], b1 = bias[c0 + j * 8 + 2 * t + 1];
      const float p0 = exp2f(s[j][0] * sl2 + b0 - lse_lo), p1 = exp2f(s[j][1] * sl2 + b1 - lse_lo);
      const float p2 = exp2f(s[j][2] * sl2 + b0 - lse_hi), p3 = exp2f(s[j][3] * sl2 + b1 - lse_hi);
      s[j][0] = p0 * (dp[j][0] - d_lo); s[j][1] = p1 * (dp[j][1] - d_lo);
      s[j][2] = p2 * (dp[j][2] - d_hi); s[j][3] = p3 * (dp[j][3] - d_hi);
    }
    mma_p_z(dq, s, k_tile, lane, c0);
  }
  store_rows(dq, 0.125f, 0.125f, q_tile, q_ptr, warp, lane, p.dqkv + (long)b * S * ld + h * ATT_D, ld, row0, S);
}

// -------------------------------------------------------------------------------------------------
// backward, dK and dV   (rows = keys, columns = queries)
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ATT_THREADS)
attn_bwd_dkv_kernel(const AttnParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int b = blockIdx.z, h = blockIdx.y, row0 = blockIdx.x * ATT_ROWS;
  const int S = p.S, SP = p.s_pad;
  uint8_t* k_ptr = smem;
  uint8_t* v_ptr = smem + ATT_ROWS * 128;
  const uint32_t k_tile = smem_u32(smem);
  const uint32_t v_tile = k_tile + ATT_ROWS * 128;
  const uint32_t q_tile = v_tile + ATT_ROWS * 128;
  const uint32_t do_tile = q_tile + SP * 128;
  float* lse_s = reinterpret_cast<float*>(smem + 2 * ATT_ROWS * 128 + 2 * SP * 128);
  float* delta_s = lse_s + SP;

  const long ld = 3L * p.H;
  const bf16* base = p.qkv + (long)b * S * ld + h * ATT_D;
  const bf16* dO = p.dctx + (long)b * S * p.H + h * ATT_D;
  const int nrows = min(ATT_ROWS, S - row0);
  load_tile(k_tile, base + p.H + (long)row0 * ld, ld, nrows, ATT_ROWS);
  load_tile(v_tile, base + 2 * p.H + (long)row0 * ld, ld, nrows, ATT_ROWS);
  load_tile(q_tile, base, ld, S, SP);
  load_tile(do_tile, dO, p.H, S, SP);
  for (int c = threadIdx.x; c < SP; c += ATT_THREADS) {
    lse_s[c] = c < S ? p.lse2[((long)b * p.heads + h) * S + c] : 1.0e30f;  // padded query columns -> P = 0
    delta_s[c] = c < S ? p.delta[((long)b * p.heads + h) * S + c] : 0.f;
  }
  cp_async_wait_all();
  __syncthreads();

  if (row0 + warp * 16 >= S) return;
  const float sl2 = 0.125f * LOG2E;
  const int r_lo = row0 + warp * 16 + g, r_hi = r_lo + 8;
  const float kb_lo = (r_lo < S && (p.mask == nullptr || ld_id(p.mask, (long)b * S + r_lo, p.mask_w) != 0)) ? 0.f : NEG_BIG;
  const float kb_hi = (r_hi < S && (p.mask == nullptr || ld_id(p.mask, (long)b * S + r_hi, p.mask_w) != 0)) ? 0.f : NEG_BIG;
  float dk[8][4], dv[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) dk[j][i] = dv[j][i] = 0.f;

  for (int c0 = 0; c0 < SP; c0 += 64) {
    float s[8][4], dp[8][4];
    mma_rows_cols(s, k_tile, q_tile, warp, lane, c0);    // S^T  = K Q^T
    mma_rows_cols(dp, v_tile, do_tile, warp, lane, c0);  // dP^T = V dO^T
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = c0 + j * 8 + 2 * t;
      const float l0 = lse_s[c], l1 = lse_s[c + 1], e0 = delta_s[c], e1 = delta_s[c + 1];
      const float p0 = exp2f(s[j][0] * sl2 + kb_lo - l0), p1 = exp2f(s[j][1] * sl2 + kb_lo - l1);
      const float p2 = exp2f(s[j][2] * sl2 + kb_hi - l0), p3 = exp2f(s[j][3] * sl2 + kb_hi - l1);
      s[j][0] = p0; s[j][1] = p1; s[j][2] = p2; s[j][3] = p3;
      dp[j][0] = p0 * (dp[j][0] - e0); dp[j][1] = p1 * (dp[j][1] - e1);
      dp[j][2] = p2 * (dp[j][2] - e0); dp[j][3] = p3 * (dp[j][3] - e1);
    }
    mma_p_z(dv, s, do_tile, lane, c0);  // dV += P^T dO
    mma_p_z(dk, dp, q_tile, lane, c0);  // dK += dS^T Q
  }
  bf16* out = p.dqkv + (long)b * S * ld + h * ATT_D;
  store_rows(dk, 0.125f, 0.125f, k_tile, k_ptr, warp, lane, out + p.H, ld, row0, S);
  store_rows(dv, 1.f, 1.f, v_tile, v_ptr, warp, lane, out + 2 * p.H, ld, row0, S);
}

// -------------------------------------------------------------------------------------------------
// fp32-class mode (3 planes): plain FFMA kernels, one thread per row, columns streamed through shared memory.
// Same math in fp32; used by the parity tests only (small shapes).
// -------------------------------------------------------------------------------------------------
constexpr int REF_CHUNK = 64;

struct AttnRefParams {
  const bf16* qkv; bf16* ctx; const bf16* dctx; bf16* dqkv; const void* mask; float* lse2; float* delta;
  long ps_qkv, ps_ctx;
  int B, S, H, heads;
  int mask_w;
  DropSpec drop;
};

// mode 0: forward; 1: dQ (+delta); 2: dK/dV
template <int MODE>
__global__ void __launch_bounds__(128)
attn_ref_kernel(const AttnRefParams p) {
  __shared__ float cA[REF_CHUNK][ATT_D + 1];
  __shared__ float cB[REF_CHUNK][ATT_D + 1];
  __shared__ float cbias[REF_CHUNK], clse[REF_CHUNK], cdelta[REF_CHUNK];
  const int b = blockIdx.z, h = blockIdx.y, S = p.S;
  const int r = blockIdx.x * 128 + threadIdx.x;
  const bool valid = r < S;
  const long ld = 3L * p.H;
  const bf16* qkv = p.qkv + (long)b * S * ld + h * ATT_D;
  const bf16* ctx = p.ctx + (long)b * S * p.H + h * ATT_D;
  const bf16* dctx = p.dctx ? p.dctx + (long)b * S * p.H + h * ATT_D : nullptr;
  const long lrow = ((long)b * p.heads + h) * S;

  float x[ATT_D], y[ATT_D], acc0[ATT_D], acc1[ATT_D];
#pragma unroll
  for (int d = 0; d < ATT_D; ++d) x[d] = y[d] = acc0[d] = acc1[d] = 0.f;
  float m = NEG_BIG, l = 0.f, lse = 0.f, delta = 0.f, kbias = 0.f;
  if (valid) {
    if (MODE == 0 || MODE == 1) {
#pragma unroll
      for (int d = 0; d < ATT_D; ++d) x[d] = plane_load<3>(qkv + (long)r * ld + d, p.ps_qkv);  // q
    }
    if (MODE == 1) {
#pragma unroll
      for (int d = 0; d < ATT_D; ++d) {
        y[d] = plane_load<3>(dctx + (long)r * p.H + d, p.ps_ctx);  // dO
        delta += y[d] * plane_load<3>(ctx + (long)r * p.H + d, p.ps_ctx);
      }
      p.delta[lrow + r] = delta;
      lse = p.lse2[lrow + r];
    }
    if (MODE == 2) {
#pragma unroll
      for (int d = 0; d < ATT_D; ++d) {
        x[d] = plane_load<3>(qkv + p.H + (long)r * ld + d, p.ps_qkv);      // k
        y[d] = plane_load<3>(qkv + 2 * p.H + (long)r * ld + d, p.ps_qkv);  // v
      }
      kbias = (p.mask == nullptr || ld_id(p.mask, (long)b * S + r, p.mask_w) != 0) ? 0.f : NEG_BIG;
    }
  }
  const float sl2 = 0.125f * LOG2E;
  for (int c0 = 0; c0 < S; c0 += REF_CHUNK) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < REF_CHUNK * ATT_D; idx += 128) {
      const int c = idx / ATT_D, d = idx % ATT_D, cc = c0 + c;
      float a = 0.f, bb = 0.f;
      if (cc < S) {
        if (MODE == 2) {
          a = plane_load<3>(qkv + (long)cc * ld + d, p.ps_qkv);          // q
          bb = plane_load<3>(dctx + (long)cc * p.H + d, p.ps_ctx);       // dO
        } else {
          a = plane_load<3>(qkv + p.H + (long)cc * ld + d, p.ps_qkv);      // k
          bb = plane_load<3>(qkv + 2 * p.H + (long)cc * ld + d, p.ps_qkv); // v
        }
      }
      cA[c][d] = a;
      cB[c][d] = bb;
    }
    for (int c = threadIdx.x; c < REF_CHUNK; c += 128) {
      const int cc = c0 + c;
      cbias[c] = (cc < S && (p.mask == nullptr || ld_id(p.mask, (long)b * S + cc, p.mask_w) != 0)) ? 0.f : NEG_BIG;
      if (MODE == 2) {
        clse[c] = cc < S ? p.lse2[lrow + cc] : 1.0e30f;
        cdelta[c] = cc < S ? p.delta[lrow + cc] : 0.f;
      }
    }
    __syncthreads();
    if (!valid) continue;
    const bool dropping = p.drop.thr_hi != 0u;
    const uint32_t plane = b * p.heads + h;
    uint4 rb = make_uint4(0u, 0u, 0u, 0u);
    for (int c = 0; c < REF_CHUNK; ++c) {
      float s = 0.f;
#pragma unroll
      for (int d = 0; d < ATT_D; ++d) s += x[d] * cA[c][d];
      // keep-multiplier of P[query, key]: MODE 0 / 1 walk the keys of one query row (one Philox block per 8 keys),
      // MODE 2 walks the queries of one key (the block of (query cc, key group r >> 3), element r & 7)
      float km = 1.f;
      if (dropping) {
        if (MODE == 2) {
          km = drop_mult(drop_bits(p.drop, static_cast<uint32_t>(r) >> 3, c0 + c, plane), r & 7, p.drop.thr_hi, p.drop.scale);
        } else {
          if ((c & 7) == 0) rb = drop_bits(p.drop, static_cast<uint32_t>(c0 + c) >> 3, r, plane);
          km = drop_mult(rb, c & 7, p.drop.thr_hi, p.drop.scale);
        }
      }
      if (MODE == 0) {
        s = s * sl2 + cbias[c];
        const float mn = fmaxf(m, s);
        const float corr = exp2f(m - mn), pe = exp2f(s - mn);
        m = mn;
        l = l * corr + pe;
        const float pd = pe * km;
#pragma unroll
        for (int d = 0; d < ATT_D; ++d) acc0[d] = acc0[d] * corr + pd * cB[c][d];
      } else if (MODE == 1) {
        const float pe = exp2f(s * sl2 + cbias[c] - lse);
        float dp = 0.f;
#pragma unroll
        for (int d = 0; d < ATT_D; ++d) dp += y[d] * cB[c][d];
        const float ds = pe * (km * dp - delta);
#pragma unroll
        for (int d = 0; d < ATT_D; ++d) acc0[d] += ds * cA[c][d];
      } else {
        const float pe = exp2f(s * sl2 + kbias - clse[c]);
        float dp = 0.f;
#pragma unroll
        for (int d = 0; d < ATT_D; ++d) dp += y[d] * cB[c][d];
        const float ds = pe * (km * dp - cdelta[c]);
        const float pd = pe * km;
#pragma unroll
        for (int d = 0; d < ATT_D; ++d) {
          acc0[d] += ds * cA[c][d];  // dK
          acc1[d] += pd * cB[c][d];  // dV
        }
      }
    }
  }
  if (!valid) return;
  if (MODE == 0) {
    const float inv = 1.f / l;
    bf16* o = p.ctx + (long)b * S * p.H + h * ATT_D + (long)r * p.H;
#pragma unroll
    for (int d = 0; d < ATT_D; ++d) split_store<3>(o + d, p.ps_ctx, acc0[d] * inv);
    if (p.lse2) p.lse2[lrow + r] = m + log2f(l);
  } else if (MODE == 1) {
    bf16* o = p.dqkv + (long)b * S * ld + h * ATT_D + (long)r * ld;
#pragma unroll
    for (int d = 0; d < ATT_D; ++d) split_store<3>(o + d, p.ps_qkv, acc0[d] * 0.125f);
  } else {
    bf16* o = p.dqkv + (long)b * S * ld + h * ATT_D + (long)r * ld;
#pragma unroll
    for (int d = 0; d < ATT_D; ++d) {
      split_store<3>(o + p.H + d, p.ps_qkv, acc0[d] * 0.125f);
      split_store<3>(o + 2 * p.H + d, p.ps_qkv, acc1[d]);
    }
  }
}

// -------------------------------------------------------------------------------------------------
// host wrappers
// -------------------------------------------------------------------------------------------------
int attention_tc_forward(const bf16* qkv, bf16* ctx, const void* mask, int mask_w, float* lse2, int B, int S, int H,
                         int heads, const DropSpec& drop, cudaStream_t st);
int attention_tc_backward(const bf16* qkv, const bf16* ctx, const bf16* dctx, bf16* dqkv, const void* mask, int mask_w,
                          const float* lse2, int B, int S, int H, int heads, float* dbias, const DropSpec& drop,
                          cudaStream_t st);
int colsum(int np, const bf16* x, long rows, int N, long row_stride, long ps, float* out, cudaStream_t st);

// single-tile sequences go to the tcgen05 kernels (attention_tc.cu); SIMX_ATTN_LEGACY=1 forces the mma.sync path
static bool use_tc_attention(int np, int S) {
  static int legacy = -1;
  if (legacy < 0) {
    const char* e = getenv("SIMX_ATTN_LEGACY");
    legacy = (e != nullptr && e[0] == '1') ? 1 : 0;
  }
  return np == 1 && S <= 128 && legacy == 0;
}

static int attn_check(int B, int S, int H, int heads) {
  SIMX_CHECK(B > 0 && S > 0, SIMX_ERR_INVALID, "attention: empty batch");
  SIMX_CHECK(heads > 0 && H == heads * ATT_D, SIMX_ERR_INVALID, "attention: head_dim must be 64 (H=%d heads=%d)", H, heads);
  SIMX_CHECK(S <= 512, SIMX_ERR_INVALID, "attention: S=%d > 512 unsupported", S);
  return SIMX_OK;
}

int attention_forward(int np, const bf16* qkv, long ps_qkv, bf16* ctx, long ps_ctx, const void* mask, int mask_w,
                      float* lse2, int B, int S, int H, int heads, const DropSpec& drop, cudaStream_t st) {
  SIMX_TRY(attn_check(B, S, H, heads));
  ProfScope prof("attn_fwd", 4.0 * B * heads * (double)S * S * ATT_D, 2.0 * np * B * (double)S * 4 * H, st);
  // 128 < S <= 256: the two-key-tile tcgen05 kernel (2048 sequences x 12 heads: S = 160 0.97 vs 1.16 ms for the mma.sync
  // kernel, S = 256 1.31 vs 1.71 ms)
  if (use_tc_attention(np, S) || use_tc_attention(np, S - 128)) {
    return attention_tc_forward(qkv, ctx, mask, mask_w, lse2, B, S, H, heads, drop, st);
  }
  if (np == 3) {
    AttnRefParams p{qkv, ctx, nullptr, nullptr, mask, lse2, nullptr, ps_qkv, ps_ctx, B, S, H, heads, mask_w, drop};
    dim3 grid((S + 127) / 128, heads, B);
    attn_ref_kernel<0><<<grid, 128, 0, st>>>(p);
    SIMX_LAUNCH_CHECK();
    return SIMX_OK;
  }
  AttnParams p{qkv, ctx, nullptr, nullptr, mask, lse2, nullptr, B, S, H, heads, (S + 63) / 64 * 64, mask_w, drop};
  const size_t smem = ATT_ROWS * 128 + 2 * (size_t)p.s_pad * 128 + p.s_pad * sizeof(float);
  static bool attr = false;
  if (!attr) {
    SIMX_CUDA(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr = true;
  }
  dim3 grid((S + ATT_ROWS - 1) / ATT_ROWS, heads, B);
  attn_fwd_kernel<<<grid, ATT_THREADS, smem, st>>>(p);
  SIMX_LAUNCH_CHECK();
  return SIMX_OK;
}

// dbias (optional): [3H] fp32, += column sums of dqkv — the fused q/k/v bias gradient.  The tcgen05 kernel reduces
// its staged output tiles in shared memory; the other paths run the column-sum kernel over dqkv afterwards.
static int attention_backward_core(int np, const bf16* qkv, long ps_qkv, const bf16* ctx, const bf16* dctx, long ps_ctx,
                                   bf16* dqkv, const void* mask, int mask_w, const float* lse2, float* delta, int B, int S,
                                   int H, int heads, const DropSpec& drop, cudaStream_t st);

int attention_backward(int np, const bf16* qkv, long ps_qkv, const bf16* ctx, const bf16* dctx, long ps_ctx, bf16* dqkv,
                       const void* mask, int mask_w, const float* lse2, float* delta, int B, int S, int H, int heads,
                       float* dbias, const DropSpec& drop, cudaStream_t st) {
  SIMX_TRY(attn_check(B, S, H, heads));
  if (use_tc_attention(np, min(S, 128))) {  // bf16 mode, any S <= 512: ceil(S/128)^2 independent tcgen05 blocks
    ProfScope prof("attn_bwd", 10.0 * B * heads * (double)S * S * ATT_D, 2.0 * np * B * (double)S * 8 * H, st);
    return attention_tc_backward(qkv, ctx, dctx, dqkv, mask, mask_w, lse2, B, S, H, heads, dbias, drop, st);
  }
  SIMX_TRY(attention_backward_core(np, qkv, ps_qkv, ctx, dctx, ps_ctx, dqkv, mask, mask_w, lse2, delta, B, S, H, heads, drop,
                                   st));
  if (dbias != nullptr) SIMX_TRY(colsum(np, dqkv, (long)B * S, 3 * H, 3L * H, ps_qkv, dbias, st));
  return SIMX_OK;
}

static int attention_backward_core(int np, const bf16* qkv, long ps_qkv, const bf16* ctx, const bf16* dctx, long ps_ctx,
                                   bf16* dqkv, const void* mask, int mask_w, const float* lse2, float* delta, int B, int S,
                                   int H, int heads, const DropSpec& drop, cudaStream_t st) {
  ProfScope prof("attn_bwd", 10.0 * B * heads * (double)S * S * ATT_D, 2.0 * np * B * (double)S * 8 * H, st);
  if (np == 3) {
    AttnRefParams p{qkv, const_cast<bf16*>(ctx), dctx, dqkv, mask, const_cast<float*>(lse2), delta, ps_qkv, ps_ctx,
                    B, S, H, heads, mask_w, drop};
    dim3 grid((S + 127) / 128, heads, B);
    attn_ref_kernel<1><<<grid, 128, 0, st>>>(p);
    SIMX_LAUNCH_CHECK();
    attn_ref_kernel<2><<<grid, 128, 0, st>>>(p);
    SIMX_LAUNCH_CHECK();
    return SIMX_OK;
  }
  SIMX_CHECK(drop.thr_hi == 0u, SIMX_ERR_UNSUPPORTED,
             "attention dropout is not implemented in the legacy mma.sync backward (unset SIMX_ATTN_LEGACY)");
  AttnParams p{qkv, const_cast<bf16*>(ctx), dctx, dqkv, mask, const_cast<float*>(lse2), delta, B, S, H, heads,
               (S + 63) / 64 * 64, mask_w, no_drop()};
  const size_t smem = 2 * ATT_ROWS * 128 + 2 * (size_t)p.s_pad * 128 + (2 * p.s_pad + ATT_ROWS) * sizeof(float);
  static bool attr = false;
  if (!attr) {
    SIMX_CUDA(cudaFuncSetAttribute(attn_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    SIMX_CUDA(cudaFuncSetAttribute(attn_bwd_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr = true;
  }
  dim3 grid((S + ATT_ROWS - 1) / ATT_ROWS, heads, B);
  attn_bwd_dq_kernel<<<grid, ATT_THREADS, smem, st>>>(p);
  SIMX_LAUNCH_CHECK();
  attn_bwd_dkv_kernel<<<grid, ATT_THREADS, smem, st>>>(p);
  SIMX_LAUNCH_CHECK();
  return SIMX_OK;
}

}  // namespace simx

extern "C" int simx_attention_forward(const void* qkv, void* ctx, const int64_t* mask, float* lse2, int B, int S, int H,
                                      int heads, int planes, int64_t qkv_plane_stride, int64_t ctx_plane_stride,
                                      void* stream) {
  using namespace simx;
  SIMX_CHECK(qkv && ctx && lse2, SIMX_ERR_INVALID, "attention_forward: null argument");
  SIMX_CHECK(planes == 1 || planes == 3, SIMX_ERR_INVALID, "attention_forward: planes must be 1 or 3");
  return attention_forward(planes, static_cast<const bf16*>(qkv), qkv_plane_stride, static_cast<bf16*>(ctx),
                           ctx_plane_stride, mask, 8, lse2, B, S, H, heads, no_drop(), static_cast<cudaStream_t>(stream));
}

extern "C" int simx_attention_backward(const void* qkv, const void* ctx, const void* dctx, void* dqkv,
                                       const int64_t* mask, const float* lse2, float* delta, int B, int S, int H,
                                       int heads, int planes, int64_t qkv_plane_stride, int64_t ctx_plane_stride,
                                       float* dbias_qkv, void* stream) {
  using namespace simx;
  SIMX_CHECK(qkv && ctx && dctx && dqkv && lse2 && delta, SIMX_ERR_INVALID, "attention_backward: null argument");
  SIMX_CHECK(planes == 1 || planes == 3, SIMX_ERR_INVALID, "attention_backward: planes must be 1 or 3");
  return attention_backward(planes, static_cast<const bf16*>(qkv), qkv_plane_stride, static_cast<const bf16*>(ctx),
                            static_cast<const bf16*>(dctx), ctx_plane_stride, static_cast<bf16*>(dqkv), mask, 8, lse2, delta,
                            B, S, H, heads, dbias_qkv, no_drop(), static_cast<cudaStream_t>(stream));
}


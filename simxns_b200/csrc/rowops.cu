// HBM-bound row kernels of the encoder: embedding gather + LayerNorm, LayerNorm forward / backward (warp-shuffle row
// reductions), column sums (bias gradients), [CLS] gather / cotangent scatter, embedding-gradient scatter, and the
// fp32 -> (split-)bf16 weight preparation.  One warp owns one row; a lane owns 8-element (16 B) chunks strided by
// 256 elements, so every global access of a warp is a contiguous 512 B segment.
//
// Reference math: LEAD/modeling_bert.py:201-240 (BertEmbeddings), :384-388 / :462-466 (residual + LayerNorm),
// torch.nn.LayerNorm (biased variance, eps inside the sqrt).
#include "common.cuh"
#include "ptx.cuh"

namespace simx {

constexpr int ROW_THREADS = 256;
constexpr int ROW_WARPS = ROW_THREADS / 32;
constexpr int MAXC = 4;  // chunks per lane: supports H <= 1024

__device__ __forceinline__ void ld8f(const float* p, float (&v)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p + 4));
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

// four packed pairs seen as the eight floats the load / store / dropout helpers take
__device__ __forceinline__ float (&as8(float2 (&v)[4]))[8] { return reinterpret_cast<float (&)[8]>(v); }

// normalise the row held in v (NC chunks of 8 per lane), write y = xhat * g + b
template <int NP>
__device__ __forceinline__ void ln_row_store(float (&v)[MAXC][8], int H, int lane, const float* gamma,
                                             const float* beta, float eps, bf16* out, long ps, const DropSpec& drop,
                                             int row) {
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < MAXC; ++c)
    if (c * 256 + lane * 8 < H)
#pragma unroll
      for (int i = 0; i < 8; ++i) s += v[c][i];
  const float mean = warp_sum(s) / H;
  float q = 0.f;
#pragma unroll
  for (int c = 0; c < MAXC; ++c)
    if (c * 256 + lane * 8 < H)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float d = v[c][i] - mean;
        q += d * d;
      }
  const float rstd = rsqrtf(warp_sum(q) / H + eps);
#pragma unroll
  for (int c = 0; c < MAXC; ++c) {
    const int col = c * 256 + lane * 8;
    if (col < H) {
      float g[8], b[8], y[8];
      ld8f(gamma + col, g);
      ld8f(beta + col, b);
#pragma unroll
      for (int i = 0; i < 8; ++i) y[i] = (v[c][i] - mean) * rstd * g[i] + b[i];
      if (drop.thr_hi != 0u) drop_apply8(y, drop_bits(drop, col >> 3, row, 0u), drop.thr_hi, drop.scale);  // embeddings dropout
      split_store8<NP>(out + col, ps, y);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// embeddings: z0 = word[id] + pos[s] (+ type[0]);  x = LN(z0).  z0 is kept for backward.
// ------------------------------------------------------------------------------------------------
template <int NP>
__global__ void __launch_bounds__(ROW_THREADS)
embed_ln_kernel(const void* __restrict__ ids, const float* __restrict__ embeds, const void* __restrict__ pos_ids, int id_bytes,
                int max_pos, int T, int S, int H, int vocab, const float* __restrict__ word, const float* __restrict__ pos,
                const float* __restrict__ type0, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                bf16* z0, bf16* x, long ps, const DropSpec drop) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * ROW_WARPS + warp;
  if (row >= T) return;
  // the word row of this token — or, with caller-supplied `inputs_embeds` (HF BertEmbeddings.forward(inputs_embeds=...),
  // LEAD/modeling_bert.py:222-227), row `row` of that [T, H] fp32 tensor
  long id = embeds != nullptr ? row : ld_id(ids, row, id_bytes);
  if (embeds == nullptr && (id < 0 || id >= vocab)) id = 0;  // out-of-range ids are a caller bug; stay in bounds
  const float* wsrc = embeds != nullptr ? embeds : word;
  // position row: the token's index in its sequence (BERT), or the caller's position ids (RoBERTa: pad + running count
  // of real tokens)
  long s = row % S;
  if (pos_ids != nullptr) {
    s = ld_id(pos_ids, row, id_bytes);
    if (s < 0 || s >= max_pos) s = 0;
  }
  float v[MAXC][8];
#pragma unroll
  for (int c = 0; c < MAXC; ++c) {
    const int col = c * 256 + lane * 8;
    if (col < H) {
      float w[8], p[8];
      ld8f(wsrc + id * H + col, w);
      ld8f(pos + s * H + col, p);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[c][i] = w[i] + p[i];
      if (type0 != nullptr) {
        ld8f(type0 + col, p);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[c][i] += p[i];
      }
      if (z0 != nullptr) split_store8<NP>(z0 + (long)row * H + col, ps, v[c]);
      if constexpr (NP == 1) {
        // the LayerNorm input must be what backward will re-read: round through the stored representation
        if (z0 != nullptr) {
#pragma unroll
          for (int i = 0; i < 8; ++i) v[c][i] = __bfloat162float(__float2bfloat16_rn(v[c][i]));
        }
      }
    }
  }
  ln_row_store<NP>(v, H, lane, gamma, beta, eps, x + (long)row * H, ps, drop, row);
}

template <int NP>
__device__ __forceinline__ void unpack8(const uint4 (&raw)[NP], float (&v)[8]) {
#pragma unroll
  for (int q = 0; q < NP; ++q) {
    const uint32_t w[4] = {raw[q].x, raw[q].y, raw[q].z, raw[q].w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = unpack_bf16x2(w[i]);
      if (q == 0) { v[2 * i] = f.x; v[2 * i + 1] = f.y; }
      else { v[2 * i] += f.x; v[2 * i + 1] += f.y; }
    }
  }
}

// A warp owns `rpw` consecutive rows: gamma / beta are read once into registers (re-reading them per row made the
// kernel L1-bound: ncu l1tex 86 %, DRAM 55 %), and the next row's 16-byte loads are in flight while the current row is
// reduced.  NC = ceil(H / 256) chunks per lane.
template <int NP, int NC>
__global__ void __launch_bounds__(ROW_THREADS)
ln_fwd_kernel(const bf16* __restrict__ z, int T, int H, const float* __restrict__ gamma,
              const float* __restrict__ beta, float eps, bf16* x, long ps, int rpw) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float2 g[NC][4], b[NC][4];
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const int col = c * 256 + lane * 8;
    if (col < H) {
      ld8f(gamma + col, as8(g[c]));
      ld8f(beta + col, as8(b[c]));
    }
  }
  uint4 nxt[NC][NP];
  auto fetch = [&](long r) {
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int col = c * 256 + lane * 8;
      if (col < H) {
#pragma unroll
        for (int q = 0; q < NP; ++q) nxt[c][q] = *reinterpret_cast<const uint4*>(z + q * ps + r * H + col);
      }
    }
  };
  const float invH = 1.f / H;
  // persistent over groups of `rpw` consecutive rows (grid = a couple of blocks per SM): gamma / beta are loaded once per
  // warp for the whole launch, no wave tail, and the next group's first row is already in flight when a group ends
  const long stride = (long)gridDim.x * ROW_WARPS * rpw;
  long r0 = ((long)blockIdx.x * ROW_WARPS + warp) * rpw;
  if (r0 < T) fetch(r0);
  for (; r0 < T; r0 += stride) {
    const long r1 = min((long)T, r0 + rpw);
    for (long r = r0; r < r1; ++r) {
      float2 v[NC][4];                       // packed fp32 (FADD2 / FFMA2 / FMUL2): half the issue slots of the scalar form
      float2 sp = make_float2(0.f, 0.f);
#pragma unroll
      for (int c = 0; c < NC; ++c)
        if (c * 256 + lane * 8 < H) {
          unpack8<NP>(nxt[c], as8(v[c]));
#pragma unroll
          for (int i = 0; i < 4; ++i) sp = __fadd2_rn(sp, v[c][i]);
        }
      if (r + 1 < r1) fetch(r + 1);
      else if (r0 + stride < T) fetch(r0 + stride);
      const float mean = warp_sum(sp.x + sp.y) * invH;
      const float2 nmean = make_float2(-mean, -mean);
      float2 qp = make_float2(0.f, 0.f);
#pragma unroll
      for (int c = 0; c < NC; ++c)
        if (c * 256 + lane * 8 < H)
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            v[c][i] = __fadd2_rn(v[c][i], nmean);
            qp = __ffma2_rn(v[c][i], v[c][i], qp);
          }
      const float rstd = rsqrtf(warp_sum(qp.x + qp.y) * invH + eps);
      const float2 rstd2 = make_float2(rstd, rstd);
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const int col = c * 256 + lane * 8;
        if (col < H) {
          float2 y[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) y[i] = __ffma2_rn(__fmul2_rn(v[c][i], rstd2), g[c][i], b[c][i]);
          split_store8<NP>(x + r * H + col, ps, as8(y));
        }
      }
    }
  }
}

// bf16 mode: the same row loop fed by the bulk-copy engine.  The register-prefetch version above keeps ONE row per warp
// in flight (ncu r02j: long-scoreboard stall 3.6 per issue, 24 % warps active, DRAM 62 %) and more rows cost registers the
// gamma / beta copies already use.  Here each warp owns a ring of LNF_SLOTS shared-memory slots of LNF_G consecutive rows;
// lane 0 keeps LNF_SLOTS - 1 bulk copies (cp.async.bulk, completion on an mbarrier per slot) ahead of the row being
// normalised, so bytes in flight no longer depend on registers: 16 warps x 2 slots x 3 KB = 96 KB per SM.
// (slots, rows per slot) swept on one box at 4 layers: (3,2) 5.5-5.6 TB/s, (4,2) 5.4, (2,4) 5.3, (4,1) 4.9, (3,4) 4.2 —
// tools/sweep_ln_fwd.sh builds the variants with -DSIMX_LNF_SLOTS / -DSIMX_LNF_G
#ifndef SIMX_LNF_SLOTS
#define SIMX_LNF_SLOTS 3
#endif
#ifndef SIMX_LNF_G
#define SIMX_LNF_G 2
#endif
constexpr int LNF_SLOTS = SIMX_LNF_SLOTS, LNF_G = SIMX_LNF_G;
template <int NC>
__global__ void __launch_bounds__(ROW_THREADS, 2)
ln_fwd_bulk_kernel(const bf16* __restrict__ z, int T, int H, const float* __restrict__ gamma, const float* __restrict__ beta,
                   float eps, bf16* x) {
  extern __shared__ __align__(128) uint8_t lnf_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t row_bytes = (uint32_t)H * 2u, slot_bytes = row_bytes * LNF_G;
  uint64_t* bars = reinterpret_cast<uint64_t*>(lnf_smem) + warp * LNF_SLOTS;
  const uint32_t ring = smem_u32(lnf_smem) + 256 + warp * LNF_SLOTS * slot_bytes;
  float2 g[NC][4], b[NC][4];
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const int col = c * 256 + lane * 8;
    if (col < H) {
      ld8f(gamma + col, as8(g[c]));
      ld8f(beta + col, as8(b[c]));
    }
  }
  // unit u = rows [u * LNF_G, (u + 1) * LNF_G); this warp takes units gw, gw + W, gw + 2 W, ...
  const long n_units = ((long)T + LNF_G - 1) / LNF_G;
  const long gw = (long)blockIdx.x * ROW_WARPS + warp, W = (long)gridDim.x * ROW_WARPS;
  auto issue = [&](long u, int slot) {  // lane 0 only
    const long r = u * LNF_G;
    const uint32_t bytes = (uint32_t)min((long)LNF_G, (long)T - r) * row_bytes;
    mbar_arrive_expect_tx(&bars[slot], bytes);
    bulk_load_1d(ring + slot * slot_bytes, z + r * H, bytes, &bars[slot]);
  };
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < LNF_SLOTS; ++s) mbar_init(&bars[s], 1);
    fence_mbar_init();
#pragma unroll
    for (int s = 0; s < LNF_SLOTS - 1; ++s)
      if (gw + s * W < n_units) issue(gw + s * W, s);
  }
  __syncwarp();
  const float invH = 1.f / H;
  int slot = 0;
  uint32_t phase = 0;
  for (long u = gw; u < n_units; u += W) {
    // refill the slot that was consumed in the previous iteration (every lane is past its shared-memory reads: the
    // __syncwarp at the end of the loop body), then wait for this unit
    if (lane == 0) {
      const long un = u + (long)(LNF_SLOTS - 1) * W;
      if (un < n_units) issue(un, slot == 0 ? LNF_SLOTS - 1 : slot - 1);
    }
    mbar_wait(&bars[slot], phase);
    const long r0 = u * LNF_G;
#pragma unroll
    for (int rr = 0; rr < LNF_G; ++rr) {
      const long r = r0 + rr;
      if (r >= T) break;
      float2 v[NC][4];
      float2 sp = make_float2(0.f, 0.f);
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const int col = c * 256 + lane * 8;
        if (col < H) {
          const uint4 raw[1] = {ld_shared_v4(ring + slot * slot_bytes + rr * row_bytes + col * 2)};
          unpack8<1>(raw, as8(v[c]));
#pragma unroll
          for (int i = 0; i < 4; ++i) sp = __fadd2_rn(sp, v[c][i]);
        }
      }
      const float mean = warp_sum(sp.x + sp.y) * invH;
      const float2 nmean = make_float2(-mean, -mean);
      float2 qp = make_float2(0.f, 0.f);
#pragma unroll
      for (int c = 0; c < NC; ++c)
        if (c * 256 + lane * 8 < H)
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            v[c][i] = __fadd2_rn(v[c][i], nmean);
            qp = __ffma2_rn(v[c][i], v[c][i], qp);
          }
      const float rstd = rsqrtf(warp_sum(qp.x + qp.y) * invH + eps);
      const float2 rstd2 = make_float2(rstd, rstd);
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const int col = c * 256 + lane * 8;
        if (col < H) {
          float2 y[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) y[i] = __ffma2_rn(__fmul2_rn(v[c][i], rstd2), g[c][i], b[c][i]);
          split_store8<1>(x + r * H + col, 0, as8(y));
        }
      }
    }
    __syncwarp();
    if (++slot == LNF_SLOTS) { slot = 0; phase ^= 1u; }
  }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm backward.  dz = rstd * (g*dy - mean(g*dy) - xhat * mean(g*dy*xhat));
// dgamma += sum dy*xhat, dbeta += sum dy, and (optionally) dbias += sum dz  — the bias gradient of the GEMM that
// produced z (BertSelfOutput.dense / BertOutput.dense), so that GEMM needs no separate column-sum pass.
// Each block walks rows with a grid stride and keeps its partial column sums in registers.
// ------------------------------------------------------------------------------------------------
// NC = ceil(H / 256) chunks per lane.  The next row's z and dy are prefetched (raw 16-byte loads) while the current
// row is reduced, so each warp keeps two rows of HBM traffic in flight.
// (A cp.async.bulk-staged variant like ln_fwd_bulk_kernel was measured and dropped: one row of z and dy per copy 9.5 ms per
// step, two rows per copy 9.0 ms, against 7.3 ms for this register-prefetch form — with ~13 packed fp32 operations per
// element pair and three running column sums the kernel is bound by issue slots, not by bytes in flight.)
constexpr int LNB_THREADS = 384;  // 12 warps x ~168 registers: the whole register file of an SM (8 warps left it half idle:
constexpr int LNB_WARPS = LNB_THREADS / 32;  // ncu warps_active 12.5 %, DRAM 53 %)
// DM: 0 no dropout; 1 the cotangent dy is masked on load (the LayerNorm output was dropped: embeddings); 2 a second,
// masked copy dzm = dz o mask is written (what flows into the dense layer whose dropped output was added to the
// residual: dz itself continues down the residual path) and dbias sums the masked values.
template <int NP, int NC, int DM>
__global__ void __launch_bounds__(LNB_THREADS)
ln_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ z, int T, int H, const float* __restrict__ gamma,
              float eps, bf16* dz, long ps, float* dgamma, float* dbeta, float* dbias, bf16* dzm, const DropSpec drop) {
  __shared__ float red[LNB_WARPS][32 * 8 + 8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float2 ag[NC][4], ab[NC][4], ad[NC][4];
#pragma unroll
  for (int c = 0; c < NC; ++c)
#pragma unroll
    for (int i = 0; i < 4; ++i) ag[c][i] = ab[c][i] = ad[c][i] = make_float2(0.f, 0.f);
  const float invH = 1.f / H;
  const int stride = gridDim.x * LNB_WARPS;
  int row = blockIdx.x * LNB_WARPS + warp;
  uint4 zr[NC][NP], dr[NC][NP];
  auto fetch = [&](int r) {
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int col = c * 256 + lane * 8;
      if (col < H) {
#pragma unroll
        for (int q = 0; q < NP; ++q) {
          zr[c][q] = *reinterpret_cast<const uint4*>(z + q * ps + (long)r * H + col);
          dr[c][q] = *reinterpret_cast<const uint4*>(dy + q * ps + (long)r * H + col);
        }
      }
    }
  };
  if (row < T) fetch(row);
  // The arithmetic below is packed: sm_100's FFMA2 / FADD2 / FMUL2 (__ffma2_rn ...) work on two fp32 values per
  // instruction.  With scalar FFMA this kernel sat on the fma pipe (ncu r02j: issue active 62 % at 64 % of DRAM, ~13 fp32
  // operations per element against a 3-register FFMA rate of one warp instruction per two cycles per scheduler).
  for (; row < T; row += stride) {
    float2 v[NC][4], g[NC][4];
    float2 sp = make_float2(0.f, 0.f);
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      if (c * 256 + lane * 8 < H) {
        unpack8<NP>(zr[c], as8(v[c]));
        unpack8<NP>(dr[c], as8(g[c]));
        if constexpr (DM == 1)
          drop_apply8(as8(g[c]), drop_bits(drop, (c * 256 + lane * 8) >> 3, row, 0u), drop.thr_hi, drop.scale);
#pragma unroll
        for (int i = 0; i < 4; ++i) sp = __fadd2_rn(sp, v[c][i]);
      }
    }
    if (row + stride < T) fetch(row + stride);
    const float mean = warp_sum(sp.x + sp.y) * invH;
    const float2 nmean = make_float2(-mean, -mean);
    float2 qp = make_float2(0.f, 0.f);
#pragma unroll
    for (int c = 0; c < NC; ++c)
      if (c * 256 + lane * 8 < H)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          v[c][i] = __fadd2_rn(v[c][i], nmean);
          qp = __ffma2_rn(v[c][i], v[c][i], qp);
        }
    const float rstd = rsqrtf(warp_sum(qp.x + qp.y) * invH + eps);
    const float2 rstd2 = make_float2(rstd, rstd);
    float2 s1p = make_float2(0.f, 0.f), s2p = make_float2(0.f, 0.f);
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int col = c * 256 + lane * 8;
      if (col < H) {
        float2 gm[4];
        ld8f(gamma + col, as8(gm));
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 xh = __fmul2_rn(v[c][i], rstd2);
          const float2 d = g[c][i];
          v[c][i] = xh;
          ag[c][i] = __ffma2_rn(d, xh, ag[c][i]);
          ab[c][i] = __fadd2_rn(ab[c][i], d);
          g[c][i] = __fmul2_rn(d, gm[i]);
          s1p = __fadd2_rn(s1p, g[c][i]);
          s2p = __ffma2_rn(g[c][i], xh, s2p);
        }
      }
    }
    const float s1 = warp_sum(s1p.x + s1p.y) * invH;
    const float s2 = warp_sum(s2p.x + s2p.y) * invH;
    // dz = rstd * (g - s1 - xhat * s2) = g * rstd + (xhat * (-rstd s2) + (-rstd s1))
    const float2 kb = make_float2(-rstd * s1, -rstd * s1), kc = make_float2(-rstd * s2, -rstd * s2);
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int col = c * 256 + lane * 8;
      if (col < H) {
        float2 o[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) o[i] = __ffma2_rn(g[c][i], rstd2, __ffma2_rn(v[c][i], kc, kb));
        split_store8<NP>(dz + (long)row * H + col, ps, as8(o));
        if constexpr (DM == 2) {
          drop_apply8(as8(o), drop_bits(drop, col >> 3, row, 0u), drop.thr_hi, drop.scale);
          split_store8<NP>(dzm + (long)row * H + col, ps, as8(o));
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) ad[c][i] = __fadd2_rn(ad[c][i], o[i]);
      }
    }
  }
  // block reduction of the three column-sum vectors, then one atomic per column per block
  float* outs[3] = {dgamma, dbeta, dbias};
#pragma unroll
  for (int which = 0; which < 3; ++which) {
    if (outs[which] == nullptr) continue;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      if (c * 256 >= H) break;
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 8; ++i)
        red[warp][lane * 8 + i] = which == 0 ? as8(ag[c])[i] : (which == 1 ? as8(ab[c])[i] : as8(ad[c])[i]);
      __syncthreads();
      const int col = c * 256 + threadIdx.x;
      if (threadIdx.x < 256 && col < H) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < LNB_WARPS; ++w) t += red[w][threadIdx.x];
        atomicAdd(outs[which] + col, t);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// out[col] += sum_{r in rows} x[r*row_stride + col]   (bias gradients; position-embedding gradient with
// row_stride = S*H and a batch loop).  grid = (ceil(N/256), row_splits)
// ------------------------------------------------------------------------------------------------
template <int NP>
__global__ void __launch_bounds__(ROW_THREADS)
colsum_kernel(const bf16* __restrict__ x, long rows, int N, long row_stride, long ps, float* out) {
  __shared__ float red[ROW_WARPS][32 * 8 + 8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int col = blockIdx.x * 256 + lane * 8;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  if (col < N) {
    for (long r = (long)blockIdx.y * ROW_WARPS + warp; r < rows; r += (long)gridDim.y * ROW_WARPS) {
      float v[8];
      plane_load8<NP>(x + r * row_stride + col, ps, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += v[i];
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) red[warp][lane * 8 + i] = acc[i];
  __syncthreads();
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c < N) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < ROW_WARPS; ++w) t += red[w][threadIdx.x];
    atomicAdd(out + c, t);
  }
}

// ------------------------------------------------------------------------------------------------
// [CLS] rows: cls_out[b, l, :] = x[b*S, :]  and the reverse (cotangent injection into the gradient stream)
// ------------------------------------------------------------------------------------------------
template <int NP>
__global__ void cls_gather_kernel(const bf16* __restrict__ x, int B, int S, int H, long ps, float* out,
                                  long out_stride_b) {
  const int b = blockIdx.x;
  for (int col = threadIdx.x * 8; col < H; col += blockDim.x * 8) {
    float v[8];
    plane_load8<NP>(x + (long)b * S * H + col, ps, v);
#pragma unroll
    for (int i = 0; i < 8; ++i) out[(long)b * out_stride_b + col + i] = v[i];
  }
}
// dx[b*S, :] += d_cls[b, l, :]
template <int NP>
__global__ void cls_scatter_add_kernel(bf16* dx, int B, int S, int H, long ps, const float* __restrict__ d,
                                       long d_stride_b) {
  const int b = blockIdx.x;
  for (int col = threadIdx.x * 8; col < H; col += blockDim.x * 8) {
    float v[8];
    plane_load8<NP>(dx + (long)b * S * H + col, ps, v);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] += d[(long)b * d_stride_b + col + i];
    split_store8<NP>(dx + (long)b * S * H + col, ps, v);
  }
}

// fp32 [n] -> planes (n % 8 == 0), src may be NULL (zeros); and planes -> fp32
template <int NP>
__global__ void f32_to_planes_kernel(const float* __restrict__ src, long n, bf16* dst, long ps) {
  for (long i = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 8; i < n; i += (long)gridDim.x * blockDim.x * 8) {
    float v[8];
    if (src != nullptr) ld8f(src + i, v);
    else
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = 0.f;
    split_store8<NP>(dst + i, ps, v);
  }
}
template <int NP>
__global__ void planes_to_f32_kernel(const bf16* __restrict__ src, long n, long ps, float* dst) {
  for (long i = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 8; i < n; i += (long)gridDim.x * blockDim.x * 8) {
    float v[8];
    plane_load8<NP>(src + i, ps, v);
    *reinterpret_cast<float4*>(dst + i) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(dst + i + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
}

// per-sequence strided variants: planes [B][SH] <-> fp32 with a batch stride (one layer of a [B, L, S, H] tensor)
template <int NP>
__global__ void planes_to_f32_strided_kernel(const bf16* __restrict__ src, long SH, long ps, float* dst, long dst_stride_b) {
  const long b = blockIdx.y;
  for (long i = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 8; i < SH; i += (long)gridDim.x * blockDim.x * 8) {
    float v[8];
    plane_load8<NP>(src + b * SH + i, ps, v);
    float* o = dst + b * dst_stride_b + i;
    *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
}
// dx[b, :] += d[b, :]  (the running cotangent takes an extra term at a layer boundary)
template <int NP>
__global__ void planes_add_f32_kernel(bf16* dx, long SH, long ps, const float* __restrict__ d, long d_stride_b) {
  const long b = blockIdx.y;
  for (long i = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 8; i < SH; i += (long)gridDim.x * blockDim.x * 8) {
    float v[8], a[8];
    plane_load8<NP>(dx + b * SH + i, ps, v);
    ld8f(d + b * d_stride_b + i, a);
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] += a[k];
    split_store8<NP>(dx + b * SH + i, ps, v);
  }
}

// ------------------------------------------------------------------------------------------------
// embedding gradient: dWord[id] += dz0[row] (skipping the padding row 0: nn.Embedding(padding_idx=0))
// ------------------------------------------------------------------------------------------------
template <int NP>
__global__ void __launch_bounds__(ROW_THREADS)
embed_scatter_kernel(const void* __restrict__ ids, int id_bytes, const bf16* __restrict__ dz0, int T, int H, int vocab,
                     int pad_id, long ps, float* dword) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * ROW_WARPS + warp;
  if (row >= T) return;
  const long id = ld_id(ids, row, id_bytes);
  if (id < 0 || id >= vocab || id == pad_id) return;  // nn.Embedding(padding_idx): the pad row never receives gradient
  for (int col = lane * 8; col < H; col += 256) {
    float v[8];
    plane_load8<NP>(dz0 + (long)row * H + col, ps, v);
    float* o = dword + id * H + col;
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3])
                 : "memory");
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + 4), "f"(v[4]), "f"(v[5]), "f"(v[6]),
                 "f"(v[7])
                 : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// weight preparation: fp32 [R,C] -> planes [R,C] at dst (+ optional transposed copy [C,R] at dst_t)
// ------------------------------------------------------------------------------------------------
template <int NP>
__global__ void convert_weight_kernel(const float* __restrict__ w, int R, int C, bf16* dst, long ld_dst, long ps_dst,
                                      bf16* dst_t, long ld_t, long ps_t) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  for (int j = ty; j < 32; j += 8) {
    const int r = r0 + j, c = c0 + tx;
    float v = 0.f;
    if (r < R && c < C) {
      v = w[(long)r * C + c];
      if (dst != nullptr) split_store<NP>(dst + (long)r * ld_dst + c, ps_dst, v);
    }
    tile[j][tx] = v;
  }
  if (dst_t == nullptr) return;
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int c = c0 + j, r = r0 + tx;
    if (r < R && c < C) split_store<NP>(dst_t + (long)c * ld_t + r, ps_t, tile[tx][j]);
  }
}

// All weight matrices of a tower in one launch (blockIdx.z = job): the per-matrix launches of prepare_weights were
// 84 launches of ~7 us per tower after every optimizer step.  A job with C == 0 copies R fp32 values (the fused qkv bias).
struct ConvJob {
  const float* w;
  bf16* dst;
  bf16* dst_t;
  int R, C;
  long ld_dst, ps_dst, ld_t, ps_t;
};
constexpr int CONV_MAX_JOBS = 56;  // 56 x 64 B of kernel parameters
struct ConvJobs {
  ConvJob j[CONV_MAX_JOBS];
};
template <int NP>
__global__ void convert_weights_batched_kernel(const ConvJobs jobs) {
  __shared__ float tile[32][33];
  const ConvJob jb = jobs.j[blockIdx.z];
  const int tx = threadIdx.x, ty = threadIdx.y;
  if (jb.C == 0) {  // plain fp32 copy of R values into `dst` (reinterpreted)
    if (blockIdx.y != 0) return;
    float* d = reinterpret_cast<float*>(jb.dst);
    for (int i = blockIdx.x * 256 + ty * 32 + tx; i < jb.R; i += gridDim.x * 256) d[i] = jb.w[i];
    return;
  }
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  if (c0 >= jb.C || r0 >= jb.R) return;
  for (int j = ty; j < 32; j += 8) {
    const int r = r0 + j, c = c0 + tx;
    float v = 0.f;
    if (r < jb.R && c < jb.C) {
      v = jb.w[(long)r * jb.C + c];
      if (jb.dst != nullptr) split_store<NP>(jb.dst + (long)r * jb.ld_dst + c, jb.ps_dst, v);
    }
    tile[j][tx] = v;
  }
  if (jb.dst_t == nullptr) return;
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int c = c0 + j, r = r0 + tx;
    if (r < jb.R && c < jb.C) split_store<NP>(jb.dst_t + (long)c * jb.ld_t + r, jb.ps_t, tile[tx][j]);
  }
}

// ------------------------------------------------------------------------------------------------
// host wrappers
// ------------------------------------------------------------------------------------------------
#define SIMX_NP_DISPATCH(NP, CALL)                                      \
  do {                                                                  \
    if ((NP) == 1) { constexpr int kNP = 1; CALL; }                     \
    else { constexpr int kNP = 3; CALL; }                               \
  } while (0)

int embed_ln(int np, const void* ids, const float* embeds, const void* pos_ids, int id_bytes, int max_pos, int T, int S, int H,
             int vocab, const float* word, const float* pos, const float* type0, const float* gamma, const float* beta,
             float eps, bf16* z0, bf16* x, long ps, const DropSpec& drop, cudaStream_t st) {
  ProfScope prof("embed_ln", 0, (double)T * H * (8.0 + 4.0 * np), st);
  SIMX_CHECK(H % 8 == 0 && H <= 256 * MAXC, SIMX_ERR_INVALID, "hidden size %d unsupported (multiple of 8, <= %d)", H,
             256 * MAXC);
  const int grid = (T + ROW_WARPS - 1) / ROW_WARPS;
  SIMX_NP_DISPATCH(np, (embed_ln_kernel<kNP><<<grid, ROW_THREADS, 0, st>>>(ids, embeds, pos_ids, id_bytes, max_pos, T, S, H, vocab,
                                                                           word, pos, type0, gamma, beta, eps, z0, x, ps, drop)));
  SIMX_LAUNCH_CHECK();
  return SIMX_OK;
}

int ln_fwd(int np, const bf16* z, int T, int H, const float* gamma, const float* beta, float eps, bf16* x, long ps,
           cudaStream_t st) {
  ProfScope prof("ln_fwd", 0, (double)T * H * 4.0 * np, st);
  SIMX_CHECK(H % 8 == 0 && H <= 256 * MAXC, SIMX_ERR_INVALID, "hidden size %d unsupported", H);
  // rows per warp: up to 8 (amortises the gamma / beta loads) while keeping >= ~8 blocks per SM in the grid
  const int rpw = (int)max(1L, min(8L, (long)T / ((long)num_sms() * 8 * ROW_WARPS)));
  const long groups = ((long)T + (long)ROW_WARPS * rpw - 1) / ((long)ROW_WARPS * rpw);
  const int grid = (int)min(groups, (long)num_sms() * 2);  // persistent: two 256-thread blocks per SM (111 registers per thread)
  const int nc = (H + 255) / 256;
  if (np == 1 && H % 64 == 0 && ((uintptr_t)z & 15) == 0) {
    // bulk-copy staged variant: persistent, two blocks per SM
    const size_t smem = 256 + (size_t)ROW_WARPS * LNF_SLOTS * LNF_G * H * 2;
    const long units = ((long)T + LNF_G - 1) / LNF_G;
    const int grid_b = (int)max(1L, min((units + ROW_WARPS - 1) / ROW_WARPS, (long)num_sms() * 2));
#define SIMX_LN_FWD_B(NC_)                                                                                              \
  do {                                                                                                                  \
    static bool attr = false;                                                                                           \
    if (!attr) {                                                                                                        \
      SIMX_CUDA(cudaFuncSetAttribute(ln_fwd_bulk_kernel<NC_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 + ROW_WARPS * LNF_SLOTS * LNF_G * 256 * NC_ * 2)); \
      attr = true;                                                                                                      \
    }                                                                                                                   \
    ln_fwd_bulk_kernel<NC_><<<grid_b, ROW_THREADS, smem, st>>>(z, T, H, gamma, beta, eps, x);                             \
  } while (0)
    if (nc == 1) SIMX_LN_FWD_B(1); else if (nc == 2) SIMX_LN_FWD_B(2); else if (nc == 3) SIMX_LN_FWD_B(3); else SIMX_LN_FWD_B(4);
#undef SIMX_LN_FWD_B
    SIMX_LAUNCH_CHECK();
    return SIMX_OK;
  }
#define SIMX_LN_FWD(NP_, NC_) ln_fwd_kernel<NP_, NC_><<<grid, ROW_THREADS, 0, st>>>(z, T, H, gamma, beta, eps, x, ps, rpw)
  if (np == 1) {
    if (nc == 1) SIMX_LN_FWD(1, 1); else if (nc == 2) SIMX_LN_FWD(1, 2); else if (nc == 3) SIMX_LN_FWD(1, 3); else SIMX_LN_FWD(1, 4);
  } else {
    if (nc == 1) SIMX_LN_FWD(3, 1); else if (nc == 2) SIMX_LN_FWD(3, 2); else if (nc == 3) SIMX_LN_FWD(3, 3); else SIMX_LN_FWD(3, 4);
  }
#undef SIMX_LN_FWD
  SIMX_LAUNCH_CHECK();
  return SIMX_OK;
}

// drop_mode 0: plain; 1: dy is masked on load (`drop` = the mask that was applied to this LayerNorm's output);
// 2: dzm (same shape as dz) receives dz o mask and dbias sums the masked values (`drop` = the mask of the dense output
// that was added to the residual to form z)
int ln_bwd(int np, const bf16* dy, const bf16* z, int T, int H, const float* gamma, float eps, bf16* dz, long ps,
           float* dgamma, float* dbeta, float* dbias, bf16* dzm, int drop_mode, const DropSpec& drop, cudaStream_t st) {
  ProfScope prof("ln_bwd", 0, (double)T * H * (drop_mode == 2 ? 8.0 : 6.0) * np, st);
  if (drop.thr_hi == 0u) drop_mode = 0;
  SIMX_CHECK(drop_mode != 2 || dzm != nullptr, SIMX_ERR_INVALID, "ln_bwd: masked output buffer missing");
  SIMX_CHECK(H % 8 == 0 && H <= 256 * MAXC, SIMX_ERR_INVALID, "hidden size %d unsupported", H);
  int grid = (T + LNB_WARPS - 1) / LNB_WARPS;
  grid = min(grid, num_sms());  // ~168 registers/thread: one resident block per SM, persistent over rows
  const int nc = (H + 255) / 256;
#define SIMX_LN_BWD_DM(NP_, NC_, DM_) \
  ln_bwd_kernel<NP_, NC_, DM_><<<grid, LNB_THREADS, 0, st>>>(dy, z, T, H, gamma, eps, dz, ps, dgamma, dbeta, dbias, dzm, drop)
#define SIMX_LN_BWD(NP_, NC_)                                       \
  do {                                                              \
    if (drop_mode == 0) SIMX_LN_BWD_DM(NP_, NC_, 0);                \
    else if (drop_mode == 1) SIMX_LN_BWD_DM(NP_, NC_, 1);           \
    else SIMX_LN_BWD_DM(NP_, NC_, 2);                               \
  } while (0)
  if (np == 1) {
    if (nc == 1) SIMX_LN_BWD(1, 1); else if (nc == 2) SIMX_LN_BWD(1, 2); else if (nc == 3) SIMX_LN_BWD(1, 3); else SIMX_LN_BWD(1, 4);
  } else {
    if (nc == 1) SIMX_LN_BWD(3, 1); else if (nc == 2) SIMX_LN_BWD(3, 2); else if (nc == 3) SIMX_LN_BWD(3, 3); else SIMX_LN_BWD(3, 4);
  }
#undef SIMX_LN_BWD
#undef SIMX_LN_BWD_DM
  SIMX_LAUNCH_CHECK();
  return SIMX_OK;
}

int colsum(int np, const bf16* x, long rows, int N, long row_stride, long ps, float* out, cudaStream_t st) {
  ProfScope prof("colsum", 0, (double)rows * N * 2.0 * np, st);
  SIMX_CHECK(N % 8 == 0, SIMX_ERR_INVALID, "colsum: N %% 8 != 0");
  dim3 grid((N + 255) / 256, (unsigned)max(1L, min((rows + ROW_WARPS * 8 - 1) / (ROW_WARPS * 8), (long)num_sms() * 2)));
  SIMX_NP_DISPATCH(np, (colsum_kernel<kNP><<<grid, ROW_THREADS, 0, st>>>(x, rows, N, row_stride, ps, out)));
  SIMX_LAUNCH_CHECK();
  return SIMX_OK;
}

int cls_gather(int np, const bf16* x, int B, int S, int H, long ps, float* out, long out_stride_b, cudaStream_t st) {
  SIMX_NP_DISPATCH(np, (cls_gather_kernel<kNP><<<B, 128, 0, st>>>(x, B, S, H, ps, out, out_stride_b)));
  SIMX_LAUNCH_CHECK();
  return SIMX_OK;
}
int cls_scatter_add(int np, bf16* dx, int B, int S, int H, long ps, const float* d, long d_stride_b,
                    cudaStream_t st) {
  SIMX_NP_DISPATCH(np, (cls_scatter_add_kernel<kNP><<<B, 128, 0, st>>>(dx, B, S, H, ps, d, d_stride_b)));
  SIMX_LAUNCH_CHECK();
  return SIMX_OK;
}
int f32_to_planes(int np, const float* src, long n, bf16* dst, long ps, cudaStream_t st) {
  ProfScope prof("f32_to_planes", 0, (double)n * (4.0 + 2.0 * np), st);
  SIMX_CHECK(n % 8 == 0, SIMX_ERR_INVALID, "f32_to_planes: n %% 8 != 0");
  const int grid = (int)min((n / 8 + 255) / 256, (long)num_sms() * 8);
  SIMX_NP_DISPATCH(np, (f32_to_planes_kernel<kNP><<<grid, 256, 0, st>>>(src, n, dst, ps)));
  SIMX_LAUNCH_CHECK();
  return SIMX_OK;
}
int planes_to_f32(int np, const bf16* src, long n, long ps, float* dst, cudaStream_t st) {
  ProfScope prof("planes_to_f32", 0, (double)n * (4.0 + 2.0 * np), st);
  SIMX_CHECK(n % 8 == 0, SIMX_ERR_INVALID, "planes_to_f32: n %% 8 != 0");
  const int grid = (int)min((n / 8 + 255) / 256, (long)num_sms() * 8);
  SIMX_NP_DISPATCH(np, (planes_to_f32_kernel<kNP><<<grid, 256, 0, st>>>(src, n, ps, dst)));
  SIMX_LAUNCH_CHECK();
  return SIMX_OK;
}
int planes_to_f32_strided(int np, const bf16* src, int B, long SH, long ps, float* dst, long dst_stride_b, cudaStream_t st) {
  ProfScope prof("planes_to_f32", 0, (double)B * SH * (4.0 + 2.0 * np), st);
  SIMX_CHECK(SH % 8 == 0 && dst_stride_b % 4 == 0, SIMX_ERR_INVALID, "planes_to_f32_strided: S*H %% 8 or stride %% 4 != 0");
  const dim3 grid((unsigned)min((SH / 8 + 255) / 256, 64L), (unsigned)B);
  SIMX_NP_DISPATCH(np, (planes_to_f32_strided_kernel<kNP><<<grid, 256, 0, st>>>(src, SH, ps, dst, dst_stride_b)));
  SIMX_LAUNCH_CHECK();
  return SIMX_OK;
}
int planes_add_f32(int np, bf16* dx, int B, long SH, long ps, const float* d, long d_stride_b, cudaStream_t st) {
  ProfScope prof("planes_add_f32", 0, (double)B * SH * (4.0 + 4.0 * np), st);
  SIMX_CHECK(SH % 8 == 0 && d_stride_b % 4 == 0, SIMX_ERR_INVALID, "planes_add_f32: S*H %% 8 or stride %% 4 != 0");
  const dim3 grid((unsigned)min((SH / 8 + 255) / 256, 64L), (unsigned)B);
  SIMX_NP_DISPATCH(np, (planes_add_f32_kernel<kNP><<<grid, 256, 0, st>>>(dx, SH, ps, d, d_stride_b)));
  SIMX_LAUNCH_CHECK();
  return SIMX_OK;
}
int embed_scatter(int np, const void* ids, int id_bytes, const bf16* dz0, int T, int H, int vocab, int pad_id, long ps,
                  float* dword, cudaStream_t st) {
  ProfScope prof("embed_scatter", 0, (double)T * H * (2.0 * np + 8.0), st);
  const int grid = (T + ROW_WARPS - 1) / ROW_WARPS;
  SIMX_NP_DISPATCH(np, (embed_scatter_kernel<kNP><<<grid, ROW_THREADS, 0, st>>>(ids, id_bytes, dz0, T, H, vocab, pad_id, ps,
                                                                                dword)));
  SIMX_LAUNCH_CHECK();
  return SIMX_OK;
}
// queue of conversion jobs, flushed as one launch per CONV_MAX_JOBS
struct ConvBatch {
  ConvJobs jobs;
  int n = 0, max_r = 0, max_c = 0;
  double bytes = 0;
};
static int conv_flush(int np, ConvBatch& b, cudaStream_t st) {
  if (b.n == 0) return SIMX_OK;
  ProfScope prof("convert_weight", 0, b.bytes, st);
  dim3 grid((b.max_c + 31) / 32, (b.max_r + 31) / 32, b.n), block(32, 8);
  SIMX_NP_DISPATCH(np, (convert_weights_batched_kernel<kNP><<<grid, block, 0, st>>>(b.jobs)));
  SIMX_LAUNCH_CHECK();
  b.n = 0; b.max_r = 0; b.max_c = 0; b.bytes = 0;
  return SIMX_OK;
}
int conv_begin(void** handle) {
  *handle = new ConvBatch();
  return SIMX_OK;
}
int conv_add(void* handle, int np, const float* w, int R, int C, bf16* dst, long ld_dst, long ps_dst, bf16* dst_t, long ld_t,
             long ps_t, cudaStream_t st) {
  ConvBatch& b = *static_cast<ConvBatch*>(handle);
  if (b.n == CONV_MAX_JOBS) SIMX_TRY(conv_flush(np, b, st));
  b.jobs.j[b.n++] = ConvJob{w, dst, dst_t, R, C, ld_dst, ps_dst, ld_t, ps_t};
  if (C == 0) {
    b.max_c = max(b.max_c, 32 * 4);  // a few blocks in x for the copy loop
    b.max_r = max(b.max_r, 1);
    b.bytes += 8.0 * R;
  } else {
    b.max_r = max(b.max_r, R);
    b.max_c = max(b.max_c, C);
    b.bytes += (double)R * C * (4.0 + 2.0 * np * (dst_t ? 2 : 1));
  }
  return SIMX_OK;
}
int conv_end(void* handle, int np, cudaStream_t st) {
  ConvBatch* b = static_cast<ConvBatch*>(handle);
  const int rc = conv_flush(np, *b, st);
  delete b;
  return rc;
}

int convert_weight(int np, const float* w, int R, int C, bf16* dst, long ld_dst, long ps_dst, bf16* dst_t, long ld_t,
                   long ps_t, cudaStream_t st) {
  ProfScope prof("convert_weight", 0, (double)R * C * (4.0 + 2.0 * np * (dst_t ? 2 : 1)), st);
  dim3 grid((C + 31) / 32, (R + 31) / 32), block(32, 8);
  SIMX_NP_DISPATCH(np, (convert_weight_kernel<kNP><<<grid, block, 0, st>>>(w, R, C, dst, ld_dst, ps_dst, dst_t, ld_t,
                                                                           ps_t)));
  SIMX_LAUNCH_CHECK();
  return SIMX_OK;
}

}  // namespace simx

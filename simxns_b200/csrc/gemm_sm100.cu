// Persistent warp-specialised tcgen05 GEMM for sm_100a:   D[M,N] = sum_k A[m,k] * B[n,k]
//
//   warp 0      : TMA producer   (cp.async.bulk.tensor, 128B-swizzled tiles, 4-stage mbarrier ring)
//   warp 1      : MMA issuer     (one thread, tcgen05.mma kind::f16 128xBNx16, fp32 accumulator in TMEM)
//   warp 2      : TMEM allocator
//   warps 4..19 : epilogue       (tcgen05.ld -> registers -> bias / GELU+GELU' / residual / multiply -> global)
//   TMEM holds two accumulators, so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// Both operands may be K-major (activations x nn.Linear weights, the "TN" case of every forward and dgrad GEMM:
// LEAD/modeling_bert.py:285-310,384-388,449-466) or MN-major (the wgrad GEMMs dW = dY^T X, whose contraction runs
// over tokens; split-K with fp32 atomics into the gradient buffer).  With 3 planes per operand (fp32-class mode)
// the producer streams the 6 significant plane pairs through the same accumulator.
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace simx {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_THREADS = 640;  // 4 control warps + 16 epilogue warps

struct GemmParams {
  int M, N, K;
  int m_tiles, n_tiles, ksplit, kb_per_split, kb_total;
  int npass;
  int pa[6], pb[6];
  bf16* out;
  bf16* out2;
  long out_ps;
  long ldo;
  const bf16* aux;
  long aux_ps;
  long ldaux;
  const float* bias;
  float* outf;
  long ldof;
  float* colsum;  // EPI_MUL, staged epilogue: [N] fp32, += column sums of the bf16 output (bias gradient of the layer
                  // whose dgrad this GEMM computes), or NULL
  DropSpec drop;  // EPI_BIAS_DROP_ADD: the dropout mask of (acc + bias), row = m, column = n
};

// TS: the epilogue stages its bf16 output tile in shared memory and writes it with TMA (full 128-byte lines) instead
// of 16-byte-per-row scattered stores — those capped every K=768 GEMM at ~1.2 TB/s of half-sector L2 writes.
// CG = 2: a CTA pair (cta_group::2) shares one 256 x BN tile — each CTA stages its own 128 rows of A and HALF of B,
// the leader's single tcgen05.mma covers M = 256 and reads both halves: operand bytes per flop drop by a third.
template <int BN, bool TS, int CG>
struct GemmSmem {
  static constexpr int STAGES = (BN == 256) ? (CG == 2 ? (TS ? 4 : 6) : (TS ? 3 : 4)) : (BN == 128 ? 4 : 6);
  static constexpr uint32_t A_BYTES = GEMM_BM * GEMM_BK * 2;
  static constexpr uint32_t B_BYTES = (BN / CG) * GEMM_BK * 2;
  static constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr uint32_t STG_OFF = STAGES * STAGE_BYTES;
  static constexpr uint32_t STG_BYTES = TS ? 16 * 4096 : 0;  // one [32 rows x 64 cols] bf16 box per epilogue warp
  static constexpr uint32_t BAR_OFF = STG_OFF + STG_BYTES;
  static constexpr uint32_t BIAS_OFF = BAR_OFF + 512;  // ring + TMEM barriers, then one barrier per epilogue warp (aux loads)
  static constexpr uint32_t BIAS_FLOATS = 4096;  // the whole bias vector lives in smem when N fits (else global loads)
  static constexpr uint32_t TOTAL = BIAS_OFF + BIAS_FLOATS * 4 + 1024;  // + alignment slack
};
template <int BN, bool MN, int EPI, int NP>
constexpr bool kTmaStore = (BN == 256) && (NP == 1) && !MN && (EPI != SIMX_EPI_ATOMIC_F32);

// operands the epilogue fetches from global memory for one 32-column chunk; issued BEFORE the TMEM wait so their
// latency overlaps the tcgen05.ld in flight
template <int EPI, int NP>
struct EpiPre {
  static constexpr bool kBias = EPI == SIMX_EPI_BIAS || EPI == SIMX_EPI_BIAS_GELU || EPI == SIMX_EPI_BIAS_ADD ||
                                EPI == SIMX_EPI_BIAS_GELU_FWD || EPI == EPI_BIAS_DROP_ADD;
  static constexpr bool kAux = EPI == SIMX_EPI_BIAS_ADD || EPI == SIMX_EPI_ADD || EPI == SIMX_EPI_MUL ||
                               EPI == EPI_BIAS_DROP_ADD;
  uint4 aux[kAux ? NP : 1][kAux ? 4 : 1];
};

template <int EPI, int NP, int CW>
__device__ __forceinline__ void epi_prefetch(const GemmParams& p, int row, int col0, EpiPre<EPI, NP>& pre) {
  if constexpr (EpiPre<EPI, NP>::kAux) {
    const bool rv = row < p.M;
#pragma unroll
    for (int q = 0; q < NP; ++q)
#pragma unroll
      for (int g = 0; g < CW / 8; ++g)
        pre.aux[q][g] = (rv && col0 + g * 8 < p.N)
                            ? *reinterpret_cast<const uint4*>(p.aux + q * p.aux_ps + (long)row * p.ldaux + col0 + g * 8)
                            : make_uint4(0u, 0u, 0u, 0u);
  }
}

// GELU (erf form, ACT2FN["gelu"] of BERT) and its derivative for the bf16 throughput mode, two elements per
// instruction in packed f16 arithmetic.  With K = 768 the tile's MMAs take ~6k cycles, i.e. a budget of ~24 issue
// slots per output element across the SM; the fp32 erf / exp formulation (22 instr + 2 MUFU per element) made the
// FFN-up GEMM epilogue-bound at 0.3x tensor peak.  Here:
//   Phi(x) = 0.5 + 0.5 tanh(u),  u = x (c0 + c1 x^2 + c2 x^4)   (minimax fit: |gelu err| < 4e-5, |gelu' err| < 1e-4)
//   gelu = x Phi,   gelu' = Phi + x * 0.5 (1 - tanh^2) u'
// one MUFU (tanh.approx.f16x2) per PAIR.  f16 round-off (~5e-4 rms) stays well under the bf16 rounding of the stored
// results (3e-3 rms); the fp32-class mode keeps the exact erf path below.
__device__ __forceinline__ void gelu_pair_f16(float x0, float x1, uint32_t& g_out, uint32_t& dg_out) {
  const __half2 x = __floats2half2_rn(x0, x1);
  // the odd polynomial u(x) is only monotone on [-8, 8]: clamp x^2 (one instruction instead of a two-sided clamp of x);
  // beyond 8 the polynomial factor freezes, u = x p(64) keeps growing linearly and tanh saturates: Phi -> 0 / 1, gelu' -> Phi
  const __half2 x2 = __hmin2(__hmul2(x, x), __float2half2_rn(64.0f));
  __half2 p = __hfma2(x2, __float2half2_rn(-3.47550486e-04f), __float2half2_rn(3.70038725e-02f));
  p = __hfma2(x2, p, __float2half2_rn(7.97422816e-01f));
  const __half2 u = __hmul2(x, p);
  uint32_t tu;
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(tu) : "r"(*reinterpret_cast<const uint32_t*>(&u)));
  const __half2 t = *reinterpret_cast<const __half2*>(&tu);
  const __half2 half = __float2half2_rn(0.5f);
  const __half2 phi = __hfma2(t, half, half);
  const __half2 g = __hmul2(x, phi);
  const __half2 w = __hfma2(__hneg2(t), t, __float2half2_rn(1.0f));  // 1 - tanh^2
  __half2 q = __hfma2(x2, __float2half2_rn(2.5f * -3.47550486e-04f), __float2half2_rn(1.5f * 3.70038725e-02f));
  q = __hfma2(x2, q, __float2half2_rn(0.5f * 7.97422816e-01f));      // 0.5 u'
  // |x| > 8: tanh is saturated (t = +-1 exactly in f16), w = 0 and the product vanishes for every finite x
  const __half2 dg = __hfma2(__hmul2(x, w), q, phi);
  const float2 gf = __half22float2(g), df = __half22float2(dg);
  g_out = pack_bf16x2(gf.x, gf.y);
  dg_out = pack_bf16x2(df.x, df.y);
}
// forward-only variant (inference): gelu alone, same tanh-form fit
__device__ __forceinline__ uint32_t gelu_only_pair_f16(float x0, float x1) {
  const __half2 x = __floats2half2_rn(x0, x1);
  const __half2 x2 = __hmin2(__hmul2(x, x), __float2half2_rn(64.0f));  // see gelu_pair_f16
  __half2 p = __hfma2(x2, __float2half2_rn(-3.47550486e-04f), __float2half2_rn(3.70038725e-02f));
  p = __hfma2(x2, p, __float2half2_rn(7.97422816e-01f));
  const __half2 u = __hmul2(x, p);
  uint32_t tu;
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(tu) : "r"(*reinterpret_cast<const uint32_t*>(&u)));
  const __half2 t = *reinterpret_cast<const __half2*>(&tu);
  const __half2 half = __float2half2_rn(0.5f);
  const float2 gf = __half22float2(__hmul2(x, __hfma2(t, half, half)));
  return pack_bf16x2(gf.x, gf.y);
}
__device__ __forceinline__ void gelu_exact(float x, float& g, float& dg) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
  g = x * cdf;
  dg = cdf + x * 0.39894228040143267794f * expf(-0.5f * x * x);
}

// values of D[row, col .. col+8) after the fused epilogue: v (and for BIAS_GELU: v = gelu', v2 = gelu) in fp32
template <int EPI, int NP>
__device__ __forceinline__ void epi_values(const GemmParams& p, int row, int col, const uint32_t* acc8,
                                           const EpiPre<EPI, NP>& pre, int g, const float* sbias, float (&v)[8]) {
  // packed fp32 arithmetic (sm_100 FADD2 / FMUL2: two values per instruction): the epilogue competes with the MMA issue
  // for ~24 issue slots per output element, and a 3-register scalar FADD only issues every other cycle per scheduler
  float2(&v2)[4] = reinterpret_cast<float2(&)[4]>(v);
#pragma unroll
  for (int i = 0; i < 4; ++i) v2[i] = make_float2(__uint_as_float(acc8[2 * i]), __uint_as_float(acc8[2 * i + 1]));
  if constexpr (EpiPre<EPI, NP>::kBias) {
    // bias from shared memory (the epilogue's own store traffic keeps evicting it from the small L1)
    const float* bsrc = sbias != nullptr ? sbias + col : p.bias + col;
    const float4 b0 = *reinterpret_cast<const float4*>(bsrc);
    const float4 b1 = *reinterpret_cast<const float4*>(bsrc + 4);
    v2[0] = __fadd2_rn(v2[0], make_float2(b0.x, b0.y));
    v2[1] = __fadd2_rn(v2[1], make_float2(b0.z, b0.w));
    v2[2] = __fadd2_rn(v2[2], make_float2(b1.x, b1.y));
    v2[3] = __fadd2_rn(v2[3], make_float2(b1.z, b1.w));
  }
  if constexpr (EPI == EPI_BIAS_DROP_ADD)  // dropout sits between the dense layer and the residual add
    drop_apply8(v, drop_bits(p.drop, static_cast<uint32_t>(col) >> 3, static_cast<uint32_t>(row), 0u), p.drop.thr_hi,
                p.drop.scale);
  if constexpr (EpiPre<EPI, NP>::kAux) {
    float2 a[4];
#pragma unroll
    for (int q = 0; q < NP; ++q) {
      const uint32_t w[4] = {pre.aux[q][g].x, pre.aux[q][g].y, pre.aux[q][g].z, pre.aux[q][g].w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = unpack_bf16x2(w[i]);
        a[i] = q == 0 ? f : __fadd2_rn(a[i], f);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if constexpr (EPI == SIMX_EPI_MUL) v2[i] = __fmul2_rn(v2[i], a[i]);
      else v2[i] = __fadd2_rn(v2[i], a[i]);
    }
  }
}

// direct (register -> global) epilogue: this thread owns D[row, col0 .. col0+CW)
template <int EPI, int NP, int CW>
__device__ __forceinline__ void epilogue_chunk(const GemmParams& p, int row, int col0, const uint32_t (&acc)[CW],
                                               const EpiPre<EPI, NP>& pre, const float* sbias) {
  if (row >= p.M) return;
#pragma unroll
  for (int g = 0; g < CW / 8; ++g) {
    const int col = col0 + g * 8;
    if (col >= p.N) break;  // N % 8 == 0 is enforced on the host
    float v[8];
    epi_values<EPI, NP>(p, row, col, &acc[g * 8], pre, g, sbias, v);
    if constexpr (EPI == SIMX_EPI_ATOMIC_F32) {
      float* o = p.outf + (long)row * p.ldof + col;
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3])
                   : "memory");
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + 4), "f"(v[4]), "f"(v[5]), "f"(v[6]),
                   "f"(v[7])
                   : "memory");
    } else if constexpr (EPI == SIMX_EPI_BIAS_GELU) {
      if constexpr (NP == 1) {
        uint32_t ge[4], dg[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) gelu_pair_f16(v[2 * i], v[2 * i + 1], ge[i], dg[i]);
        *reinterpret_cast<uint4*>(p.out + (long)row * p.ldo + col) = make_uint4(dg[0], dg[1], dg[2], dg[3]);
        *reinterpret_cast<uint4*>(p.out2 + (long)row * p.ldo + col) = make_uint4(ge[0], ge[1], ge[2], ge[3]);
      } else {
        float ge[8], dg[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) gelu_exact(v[i], ge[i], dg[i]);
        split_store8<NP>(p.out + (long)row * p.ldo + col, p.out_ps, dg);   // gelu'(pre-activation), kept for backward
        split_store8<NP>(p.out2 + (long)row * p.ldo + col, p.out_ps, ge);  // gelu(pre-activation)
      }
    } else if constexpr (EPI == SIMX_EPI_BIAS_GELU_FWD) {
      if constexpr (NP == 1) {
        *reinterpret_cast<uint4*>(p.out + (long)row * p.ldo + col) =
            make_uint4(gelu_only_pair_f16(v[0], v[1]), gelu_only_pair_f16(v[2], v[3]), gelu_only_pair_f16(v[4], v[5]),
                       gelu_only_pair_f16(v[6], v[7]));
      } else {
        float ge[8], dg;
#pragma unroll
        for (int i = 0; i < 8; ++i) gelu_exact(v[i], ge[i], dg);
        split_store8<NP>(p.out + (long)row * p.ldo + col, p.out_ps, ge);
      }
    } else {
      split_store8<NP>(p.out + (long)row * p.ldo + col, p.out_ps, v);
    }
  }
}

template <int BN, bool MN, int EPI, int NP, int CG>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_kernel(const GemmParams p, const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
            const __grid_constant__ CUtensorMap mapOut, const __grid_constant__ CUtensorMap mapOut2,
            const __grid_constant__ CUtensorMap mapAux) {
  constexpr bool TS = kTmaStore<BN, MN, EPI, NP>;
  using L = GemmSmem<BN, TS, CG>;
  constexpr int TILE_M = GEMM_BM * CG;  // rows of one output tile (both CTAs of a pair)
  const uint32_t cta_rank = CG == 2 ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;
  constexpr int STAGES = L::STAGES;
  extern __shared__ uint8_t smem_raw[];
  // 128B swizzle atoms need 1024-byte aligned tiles
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * L::A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  uint64_t* full = bars;                    // [STAGES]  TMA -> MMA
  uint64_t* empty = bars + STAGES;          // [STAGES]  MMA -> TMA
  uint64_t* tfull = bars + 2 * STAGES;      // [2]       MMA -> epilogue
  uint64_t* tempty = bars + 2 * STAGES + 2; // [2]       epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  // staged epilogue with an aux operand: the warp's [32 x 64] aux box is TMA-loaded into its staging box (a lane reading
  // its own row straight from global touches 32 different lines per instruction: the K = 768 GEMMs with a residual /
  // GELU' operand ran 20 % behind the plain epilogue on L1 wavefronts)
  constexpr bool kAuxTma = TS && EpiPre<EPI, NP>::kAux;
  uint64_t* auxbar = bars + 2 * STAGES + 6;  // [16 warps][2 half boxes]
  float* sbias = nullptr;
  if constexpr (EpiPre<EPI, NP>::kBias) {
    if (p.N <= (int)L::BIAS_FLOATS) {
      sbias = reinterpret_cast<float*>(smem + L::BIAS_OFF);
      for (int i = threadIdx.x; i < p.N; i += GEMM_THREADS) sbias[i] = p.bias[i];
    }
  }

  // column sums of the output (fused bias gradient): per-CTA fp32 accumulator in the (otherwise unused) bias area
  float* scol = nullptr;
  if constexpr (EPI == SIMX_EPI_MUL && TS) {
    if (p.colsum != nullptr) {
      scol = reinterpret_cast<float*>(smem + L::BIAS_OFF);
      for (int i = threadIdx.x; i < p.N; i += GEMM_THREADS) scol[i] = 0.f;
    }
  }

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapA);
    tma_prefetch_desc(&mapB);
    if constexpr (TS) {
      tma_prefetch_desc(&mapOut);
      if constexpr (EPI == SIMX_EPI_BIAS_GELU) tma_prefetch_desc(&mapOut2);
      if constexpr (kAuxTma) tma_prefetch_desc(&mapAux);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], CG);   // leader: arrive.expect_tx; peer: remote arrive (its bytes land on the leader's barrier)
      mbar_init(&empty[i], 1);   // tcgen05.commit, multicast to both CTAs when CG == 2
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 16 * CG);  // every epilogue warp of the pair arrives on the LEADER's barrier
    }
    if constexpr (kAuxTma)
      for (int i = 0; i < 32; ++i) mbar_init(&auxbar[i], 1);
    fence_mbar_init();
  }
  if constexpr (CG == 2) cluster_sync_all();  // both CTAs alive before the pair-wide TMEM allocation
  if (warp == 2) tmem_alloc<CG>(tmem_slot, 2 * BN);
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all();
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int total_tiles = p.m_tiles * p.n_tiles * p.ksplit;
  const int first_tile = blockIdx.x / CG, tile_stride = gridDim.x / CG;  // a CTA pair walks the tiles together

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = first_tile; tile < total_tiles; tile += tile_stride) {
        const int nt = tile % p.n_tiles;
        const int mt = (tile / p.n_tiles) % p.m_tiles;
        const int ks = tile / (p.n_tiles * p.m_tiles);
        const int kb0 = ks * p.kb_per_split;
        const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        // this CTA's 128 rows of A and its BN/CG columns of B
        const int m0 = mt * TILE_M + cta_rank * GEMM_BM, n0 = nt * BN + cta_rank * (BN / CG);
        for (int pass = 0; pass < p.npass; ++pass) {
          const int pla = p.pa[pass], plb = p.pb[pass];
          for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(&empty[stage], phase ^ 1);
            if (leader) mbar_arrive_expect_tx(&full[stage], L::STAGE_BYTES * CG);
            else mbar_arrive_remote(&full[stage], 0);
            uint8_t* a = sA + stage * L::A_BYTES;
            uint8_t* b = sB + stage * L::B_BYTES;
            if constexpr (!MN) {
              tma_load_3d_cg<CG>(&mapA, &full[stage], a, kb * GEMM_BK, m0, pla);
              tma_load_3d_cg<CG>(&mapB, &full[stage], b, kb * GEMM_BK, n0, plb);
            } else {
              // MN-major: one [64 k-rows x 64 mn (128 B)] box per 64-wide slice of the MN extent
#pragma unroll
              for (int i = 0; i < GEMM_BM / 64; ++i)
                tma_load_3d_cg<CG>(&mapA, &full[stage], a + i * (GEMM_BK * 128), m0 + i * 64, kb * GEMM_BK, pla);
#pragma unroll
              for (int i = 0; i < BN / CG / 64; ++i)
                tma_load_3d_cg<CG>(&mapB, &full[stage], b + i * (GEMM_BK * 128), n0 + i * 64, kb * GEMM_BK, plb);
            }
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && leader) {
      // ===================== MMA issuer (leader CTA only when CG == 2) =====================
      constexpr uint32_t idesc = make_idesc_bf16(TILE_M, BN, MN, MN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = first_tile; tile < total_tiles; tile += tile_stride, ++it) {
        const int ks = tile / (p.n_tiles * p.m_tiles);
        const int kb0 = ks * p.kb_per_split;
        const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        const int nk = (kb1 - kb0) * p.npass;
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int i = 0; i < nk; ++i) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + stage * L::A_BYTES);
          const uint32_t b_addr = smem_u32(sB + stage * L::B_BYTES);
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {
            uint64_t da, db;
            if constexpr (!MN) {
              // K-major, 128B swizzle: 8-row groups 1024 B apart; 16 k-elements = 32 B inside the swizzle atom
              da = make_smem_desc_sw128(a_addr + k * 32, 0, 1024);
              db = make_smem_desc_sw128(b_addr + k * 32, 0, 1024);
            } else {
              // MN-major, 128B swizzle: atoms of [8 k-rows x 64 mn]; next 8 k-rows +1024 B (SBO), next 64 mn
              // +BK*128 B (LBO); 16 k-rows per MMA = 2048 B
              da = make_smem_desc_sw128(a_addr + k * 2048, GEMM_BK * 128, 1024);
              db = make_smem_desc_sw128(b_addr + k * 2048, GEMM_BK * 128, 1024);
            }
            umma_bf16<CG>(d_tmem, da, db, idesc, (i | k) != 0);
          }
          umma_commit<CG>(&empty[stage]);  // frees the smem slot (in both CTAs) when these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit<CG>(&tfull[acc]);  // accumulator complete -> epilogue warps of both CTAs
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: 16 warps, four per TMEM lane quarter, each owning a quarter of the columns ======
    // (the epilogue is latency-bound — TMEM load, operand fetch, MUFU — so it wants warps, not wider warps)
    const int quarter = warp & 3;          // the TMEM lanes [32*quarter, +32) this warp may read
    const int part = (warp - 4) >> 2;      // 0..3
    constexpr int CPW = BN / 4;            // columns per warp: 64 / 32 / 16
    constexpr int CW = CPW >= 32 ? 32 : 16;  // columns per tcgen05.ld
    constexpr int NCH = CPW / CW;
    int it = 0;
    for (int tile = first_tile; tile < total_tiles; tile += tile_stride, ++it) {
      const int nt = tile % p.n_tiles;
      const int mt = (tile / p.n_tiles) % p.m_tiles;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int row_base = mt * TILE_M + cta_rank * GEMM_BM + quarter * 32;
      const int row = row_base + lane;
      const int col_base = nt * BN + part * CPW;
      EpiPre<EPI, NP> pre;
      if constexpr (!TS) {
        epi_prefetch<EPI, NP, CW>(p, row, col_base, pre);
        mbar_wait(&tfull[acc], acc_phase);
        tc_fence_after();
        const uint32_t t_addr = tmem_base + acc * BN + part * CPW + (static_cast<uint32_t>(quarter * 32) << 16);
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          uint32_t v[CW];
          tmem_ld_cols<CW>(t_addr + c * CW, v);
          tmem_ld_wait();
          if (c + 1 == NCH) {
            // accumulator fully read by this warp: hand the TMEM buffer back before the math and the global stores
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (leader) mbar_arrive(&tempty[acc]); else mbar_arrive_remote(&tempty[acc], 0); }
          }
          epilogue_chunk<EPI, NP, CW>(p, row, col_base + c * CW, v, pre, sbias);
          if (c + 1 < NCH) epi_prefetch<EPI, NP, CW>(p, row, col_base + (c + 1) * CW, pre);
        }
      } else {
        // ---- staged epilogue (BN = 256, one plane): the warp's [32 rows x 64 cols] block goes out as two half boxes of
        // [32 x 32] bf16 (64-byte rows, 64B swizzle) through TMA stores.  Chunk c of a tile uses half box c, so the store of
        // one half drains while the other half is being computed, and an aux half (residual / GELU' / addend) is TMA-loaded
        // into its box one chunk of work ahead.  (One [32 x 64] box per warp serialised load -> math -> store: ncu r02 showed
        // 16 % of the FFN-up epilogue's samples on cp.async.bulk.wait_group.read and the K = 768 GEMMs at 53-69 % tensor
        // pipe.)  BIAS_GELU has two outputs per chunk (gelu' -> half 0, gelu -> half 1), stored together.
        static_assert(CW == 32 && NCH == 2, "staged epilogue: two 32-column chunks per warp");
        uint8_t* stg = smem + L::STG_OFF + (warp - 4) * 4096;
        uint64_t* abar = auxbar + 2 * (warp - 4);
        constexpr bool kGelu2 = EPI == SIMX_EPI_BIAS_GELU;
        if constexpr (kAuxTma) {
          // aux of chunk 1 -> half 1: that half's last store (previous tile, chunk 1) is the most recent one
          __syncwarp();
          if (lane == 0) {
            if (it == 0) {  // first tile of this CTA: nothing was prefetched yet
              mbar_arrive_expect_tx(&abar[0], 2048);
              tma_load_3d(&mapAux, &abar[0], stg, col_base, row_base, 0);
            }
            tma_store_wait_read<0>();
            mbar_arrive_expect_tx(&abar[1], 2048);
            tma_load_3d(&mapAux, &abar[1], stg + 2048, col_base + 32, row_base, 0);
          }
        }
        mbar_wait(&tfull[acc], acc_phase);
        tc_fence_after();
        const uint32_t t_addr = tmem_base + acc * BN + part * CPW + (static_cast<uint32_t>(quarter * 32) << 16);
        const uint32_t sw = (lane >> 1) & 3;  // 64B swizzle: 16-byte chunk q of row r sits at q ^ ((r >> 1) & 3)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t v[CW];
          tmem_ld_cols<CW>(t_addr + c * CW, v);
          tmem_ld_wait();
          if (c == 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (leader) mbar_arrive(&tempty[acc]); else mbar_arrive_remote(&tempty[acc], 0); }
          }
          uint8_t* box = stg + (kGelu2 ? 0 : c * 2048);
          const uint32_t box_row = smem_u32(box) + lane * 64;
          if constexpr (kAuxTma) mbar_wait(&abar[c], it & 1);  // this chunk's aux half has landed (one phase per tile)
          uint4 o[4], o2[kGelu2 ? 4 : 1];
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            if constexpr (kAuxTma) pre.aux[0][g] = ld_shared_v4(box_row + ((g ^ sw) << 4));
            float f[8];
            epi_values<EPI, NP>(p, row, min(col_base + c * 32 + g * 8, p.N - 8), &v[g * 8], pre, g, sbias, f);
            if constexpr (kGelu2) {
              uint32_t ge[4], dg[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) gelu_pair_f16(f[2 * i], f[2 * i + 1], ge[i], dg[i]);
              o[g] = make_uint4(dg[0], dg[1], dg[2], dg[3]);
              o2[g] = make_uint4(ge[0], ge[1], ge[2], ge[3]);
            } else if constexpr (EPI == SIMX_EPI_BIAS_GELU_FWD) {
              o[g] = make_uint4(gelu_only_pair_f16(f[0], f[1]), gelu_only_pair_f16(f[2], f[3]), gelu_only_pair_f16(f[4], f[5]),
                                gelu_only_pair_f16(f[6], f[7]));
            } else {
              o[g] = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
                                pack_bf16x2(f[6], f[7]));
            }
          }
          // the box must be free: its previous store has finished READING shared memory.  Those stores were issued one
          // chunk of math ago (with an aux operand the wait already happened before the aux load went into the box).
          if constexpr (!kAuxTma) {
            if (lane == 0) {
              if constexpr (kGelu2) tma_store_wait_read<0>();
              else tma_store_wait_read<1>();
            }
            __syncwarp();
          }
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            st_shared_v4(box_row + ((g ^ sw) << 4), o[g]);
            if constexpr (kGelu2) st_shared_v4(box_row + 2048 + ((g ^ sw) << 4), o2[g]);
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_3d(&mapOut, box, col_base + c * 32, row_base, 0);
            if constexpr (kGelu2) tma_store_3d(&mapOut2, box + 2048, col_base + c * 32, row_base, 0);
            tma_store_commit();
          }
          if constexpr (EPI == SIMX_EPI_MUL) {
            if (scol != nullptr) {
              // column sums of the bf16 values being stored: lane l takes column pair l & 15 over 16 of the 32 rows (the two
              // half-warps start one row apart so that their 64-byte rows fall into different bank halves)
              const int pr = lane & 15, hsel = lane >> 4;
              const uint32_t sb = smem_u32(box) + ((pr & 3) << 2);
              float2 sp = make_float2(0.f, 0.f);
#pragma unroll 8
              for (int k = 0; k < 16; ++k) {
                const int r = hsel * 16 + ((k + hsel) & 15);
                sp = __fadd2_rn(sp, unpack_bf16x2(ld_shared_u32(sb + r * 64 + ((((pr >> 2) ^ (r >> 1)) & 3) << 4))));
              }
              float s0 = sp.x, s1 = sp.y;
              s0 += __shfl_xor_sync(0xffffffffu, s0, 16);
              s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
              const int col = col_base + c * 32 + 2 * pr;
              if (hsel == 0 && col < p.N) {
                atomicAdd(&scol[col], s0);
                atomicAdd(&scol[col + 1], s1);
              }
              __syncwarp();  // every lane is done reading the box before the next aux half may land in it
            }
          }
          if constexpr (kAuxTma) {
            if (c == 1) {
              // prefetch the NEXT tile's chunk-0 aux half into half 0: its store (chunk 0 of this tile) was issued a chunk
              // of math ago; only the chunk-1 store just issued may still be pending
              const int nxt = tile + tile_stride;
              if (nxt < total_tiles && lane == 0) {
                const int nnt = nxt % p.n_tiles, nmt = (nxt / p.n_tiles) % p.m_tiles;
                tma_store_wait_read<1>();
                mbar_arrive_expect_tx(&abar[0], 2048);
                tma_load_3d(&mapAux, &abar[0], stg, nnt * BN + part * CPW, nmt * TILE_M + cta_rank * GEMM_BM + quarter * 32, 0);
              }
            }
          }
        }
      }
    }
    if constexpr (TS) {
      if (lane == 0) tma_store_wait_all<0>();  // smem must outlive the last bulk store
    }
    if constexpr (EPI == SIMX_EPI_MUL && TS) {
      if (scol != nullptr) {
        named_bar_sync(1, GEMM_THREADS - 128);  // the 16 epilogue warps: every shared-memory atomic has landed
        for (int i = threadIdx.x - 128; i < p.N; i += GEMM_THREADS - 128) {
          const float v = scol[i];
          if (v != 0.f) atomicAdd(p.colsum + i, v);
        }
      }
    }
  }

  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all();  // the peer may still be arriving on / reading from this CTA
  else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<CG>(tmem_base, 2 * BN);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(ptr);
  return fn;
}

// bf16 tensor [planes][outer][inner] (inner contiguous), box = [1][box_outer][64]
static int make_tmap_bf16_box(CUtensorMap* map, const void* ptr, uint64_t inner, uint64_t outer, uint64_t planes, uint64_t ld,
                              uint64_t plane_stride, uint32_t box_outer, uint32_t box_inner);
int make_tmap_bf16(CUtensorMap* map, const void* ptr, uint64_t inner, uint64_t outer, uint64_t planes, uint64_t ld,
                   uint64_t plane_stride, uint32_t box_outer) {
  return make_tmap_bf16_box(map, ptr, inner, outer, planes, ld, plane_stride, box_outer, 64);
}
// box_inner = 64 elements (128-byte rows, 128B swizzle: operand tiles) or 32 (64-byte rows, 64B swizzle: the epilogue's
// half boxes)
static int make_tmap_bf16_box(CUtensorMap* map, const void* ptr, uint64_t inner, uint64_t outer, uint64_t planes, uint64_t ld,
                              uint64_t plane_stride, uint32_t box_outer, uint32_t box_inner) {
  EncodeTiledFn fn = get_encode_fn();
  SIMX_CHECK(fn != nullptr, SIMX_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  SIMX_CHECK((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, SIMX_ERR_INVALID, "TMA operand not 16-byte aligned");
  SIMX_CHECK((ld * 2) % 16 == 0, SIMX_ERR_INVALID, "TMA operand row stride (%llu elements) not a multiple of 8",
             (unsigned long long)ld);
  if (planes < 1) planes = 1;
  if (planes == 1) plane_stride = ld * outer;  // any legal value
  SIMX_CHECK((plane_stride * 2) % 16 == 0, SIMX_ERR_INVALID, "TMA plane stride not a multiple of 8 elements");
  cuuint64_t gdim[3] = {inner, outer, planes};
  cuuint64_t gstride[2] = {ld * 2, plane_stride * 2};
  cuuint32_t box[3] = {box_inner, box_outer, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, box_inner == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SIMX_CHECK(r == CUDA_SUCCESS, SIMX_ERR_CUDA,
             "cuTensorMapEncodeTiled failed (%d): inner=%llu outer=%llu planes=%llu ld=%llu box_outer=%u", (int)r,
             (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)planes, (unsigned long long)ld,
             box_outer);
  return SIMX_OK;
}

// SIMX_GEMM_1CTA=1 forces the single-CTA kernels (A/B comparison, fallback)
static bool use_cta_pairs() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SIMX_GEMM_1CTA");
    v = (e != nullptr && e[0] == '1') ? 0 : 1;
  }
  return v == 1;
}

template <int BN, bool MN, int EPI, int NP, int CG>
static int launch_gemm_cg(GemmParams p, const bf16* A, long lda, long a_ps, const bf16* B, long ldb, long b_ps, int planes,
                          cudaStream_t stream) {
  auto kern = gemm_kernel<BN, MN, EPI, NP, CG>;
  constexpr bool TS = kTmaStore<BN, MN, EPI, NP>;
  constexpr uint32_t smem_bytes = GemmSmem<BN, TS, CG>::TOTAL;
  static bool attr_set = false;
  if (!attr_set) {
    SIMX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    attr_set = true;
  }
  p.m_tiles = (p.M + GEMM_BM * CG - 1) / (GEMM_BM * CG);
  CUtensorMap ma, mb;
  if (!MN) {
    SIMX_TRY(make_tmap_bf16(&ma, A, p.K, p.M, planes, lda, a_ps, GEMM_BM));
    SIMX_TRY(make_tmap_bf16(&mb, B, p.K, p.N, planes, ldb, b_ps, BN / CG));
  } else {
    SIMX_TRY(make_tmap_bf16(&ma, A, p.M, p.K, planes, lda, a_ps, GEMM_BK));
    SIMX_TRY(make_tmap_bf16(&mb, B, p.N, p.K, planes, ldb, b_ps, GEMM_BK));
  }
  CUtensorMap mo = ma, mo2 = ma, mx = ma;  // placeholders when the epilogue stores from registers
  if constexpr (TS) {
    SIMX_TRY(make_tmap_bf16_box(&mo, p.out, p.N, p.M, 1, p.ldo, 0, 32, 32));
    if constexpr (EPI == SIMX_EPI_BIAS_GELU) SIMX_TRY(make_tmap_bf16_box(&mo2, p.out2, p.N, p.M, 1, p.ldo, 0, 32, 32));
    if constexpr (EpiPre<EPI, NP>::kAux) SIMX_TRY(make_tmap_bf16_box(&mx, p.aux, p.N, p.M, 1, p.ldaux, 0, 32, 32));
  }
  const int total = p.m_tiles * p.n_tiles * p.ksplit;
  const int grid = CG * min(total, num_sms() / CG);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  SIMX_CUDA(cudaLaunchKernelEx(&cfg, kern, p, ma, mb, mo, mo2, mx));
  SIMX_LAUNCH_CHECK();
  return SIMX_OK;
}

template <int BN, bool MN, int EPI, int NP>
static int launch_gemm(const GemmParams& p, const bf16* A, long lda, long a_ps, const bf16* B, long ldb, long b_ps, int planes,
                       cudaStream_t stream) {
  if constexpr (BN == 256) {
    if (use_cta_pairs()) return launch_gemm_cg<BN, MN, EPI, NP, 2>(p, A, lda, a_ps, B, ldb, b_ps, planes, stream);
  }
  return launch_gemm_cg<BN, MN, EPI, NP, 1>(p, A, lda, a_ps, B, ldb, b_ps, planes, stream);
}

template <int BN, bool MN, int NP>
static int dispatch_epi(int epi, const GemmParams& p, const bf16* A, long lda, long a_ps, const bf16* B, long ldb,
                        long b_ps, int planes, cudaStream_t stream) {
  if constexpr (MN) {
    SIMX_CHECK(epi == SIMX_EPI_ATOMIC_F32 || epi == SIMX_EPI_PLAIN, SIMX_ERR_INVALID,
               "MN-major GEMM supports the PLAIN and ATOMIC_F32 epilogues only");
    if (epi == SIMX_EPI_ATOMIC_F32) return launch_gemm<BN, MN, SIMX_EPI_ATOMIC_F32, NP>(p, A, lda, a_ps, B, ldb, b_ps, planes, stream);
    return launch_gemm<BN, MN, SIMX_EPI_PLAIN, NP>(p, A, lda, a_ps, B, ldb, b_ps, planes, stream);
  } else {
    switch (epi) {
      case SIMX_EPI_PLAIN: return launch_gemm<BN, MN, SIMX_EPI_PLAIN, NP>(p, A, lda, a_ps, B, ldb, b_ps, planes, stream);
      case SIMX_EPI_BIAS: return launch_gemm<BN, MN, SIMX_EPI_BIAS, NP>(p, A, lda, a_ps, B, ldb, b_ps, planes, stream);
      case SIMX_EPI_BIAS_GELU: return launch_gemm<BN, MN, SIMX_EPI_BIAS_GELU, NP>(p, A, lda, a_ps, B, ldb, b_ps, planes, stream);
      case SIMX_EPI_BIAS_ADD: return launch_gemm<BN, MN, SIMX_EPI_BIAS_ADD, NP>(p, A, lda, a_ps, B, ldb, b_ps, planes, stream);
      case SIMX_EPI_MUL: return launch_gemm<BN, MN, SIMX_EPI_MUL, NP>(p, A, lda, a_ps, B, ldb, b_ps, planes, stream);
      case SIMX_EPI_ADD: return launch_gemm<BN, MN, SIMX_EPI_ADD, NP>(p, A, lda, a_ps, B, ldb, b_ps, planes, stream);
      case SIMX_EPI_ATOMIC_F32: return launch_gemm<BN, MN, SIMX_EPI_ATOMIC_F32, NP>(p, A, lda, a_ps, B, ldb, b_ps, planes, stream);
      case SIMX_EPI_BIAS_GELU_FWD: return launch_gemm<BN, MN, SIMX_EPI_BIAS_GELU_FWD, NP>(p, A, lda, a_ps, B, ldb, b_ps, planes, stream);
      case EPI_BIAS_DROP_ADD: return launch_gemm<BN, MN, EPI_BIAS_DROP_ADD, NP>(p, A, lda, a_ps, B, ldb, b_ps, planes, stream);
    }
    SIMX_FAIL(SIMX_ERR_INVALID, "unknown GEMM epilogue %d", epi);
  }
}

static int gemm_launch(int M, int N, int K, const bf16* A, long lda, long a_ps, bool a_mn, const bf16* B, long ldb, long b_ps,
                       bool b_mn, int planes, int epi, const float* bias, const bf16* aux, long ldaux, long aux_ps,
                       bf16* out, bf16* out2, long ldo, long out_ps, float* outf, long ldof, int ksplit,
                       bool* colsum_fused, const DropSpec* drop, cudaStream_t stream) {
  SIMX_CHECK(M > 0 && N > 0 && K > 0, SIMX_ERR_INVALID, "gemm: empty problem %dx%dx%d", M, N, K);
  SIMX_CHECK(N % 8 == 0, SIMX_ERR_INVALID, "gemm: N=%d must be a multiple of 8", N);
  SIMX_CHECK(a_mn == b_mn, SIMX_ERR_INVALID, "gemm: mixed operand majors are not instantiated");
  SIMX_CHECK(planes == 1 || planes == 3, SIMX_ERR_INVALID, "gemm: planes must be 1 or 3");
  SIMX_CHECK(ksplit >= 1, SIMX_ERR_INVALID, "gemm: ksplit must be >= 1");
  SIMX_CHECK(ksplit == 1 || epi == SIMX_EPI_ATOMIC_F32, SIMX_ERR_INVALID, "gemm: split-K needs the atomic epilogue");
  if (epi == SIMX_EPI_ATOMIC_F32) {
    SIMX_CHECK(outf != nullptr && ldof % 4 == 0 && (reinterpret_cast<uintptr_t>(outf) & 15) == 0, SIMX_ERR_INVALID,
               "gemm: fp32 output must be 16-byte aligned with ld %% 4 == 0");
  } else {
    SIMX_CHECK(out != nullptr && ldo % 8 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0, SIMX_ERR_INVALID,
               "gemm: bf16 output must be 16-byte aligned with ld %% 8 == 0");
  }
  if (epi == SIMX_EPI_BIAS_ADD && drop != nullptr && drop->thr_hi != 0u) epi = EPI_BIAS_DROP_ADD;
  if (epi == SIMX_EPI_BIAS || epi == SIMX_EPI_BIAS_GELU || epi == SIMX_EPI_BIAS_ADD || epi == SIMX_EPI_BIAS_GELU_FWD ||
      epi == EPI_BIAS_DROP_ADD)
    SIMX_CHECK(bias != nullptr && (reinterpret_cast<uintptr_t>(bias) & 15) == 0, SIMX_ERR_INVALID,
               "gemm: bias missing or misaligned");
  if (epi == SIMX_EPI_BIAS_ADD || epi == SIMX_EPI_MUL || epi == SIMX_EPI_ADD || epi == EPI_BIAS_DROP_ADD)
    SIMX_CHECK(aux != nullptr && ldaux % 8 == 0 && (reinterpret_cast<uintptr_t>(aux) & 15) == 0, SIMX_ERR_INVALID,
               "gemm: aux missing or misaligned");
  if (epi == SIMX_EPI_BIAS_GELU) SIMX_CHECK(out2 != nullptr, SIMX_ERR_INVALID, "gemm: out2 missing");

  // tile width: 256 when N fills it, else the smaller instantiations (less padding waste)
  int BN = 256;
  if (N <= 64) BN = 64;
  else if (N <= 128) BN = 128;

  GemmParams p;
  p.M = M; p.N = N; p.K = K;
  p.m_tiles = (M + GEMM_BM - 1) / GEMM_BM;
  p.n_tiles = (N + BN - 1) / BN;
  p.kb_total = (K + GEMM_BK - 1) / GEMM_BK;
  if (ksplit > p.kb_total) ksplit = p.kb_total;
  p.kb_per_split = (p.kb_total + ksplit - 1) / ksplit;
  p.ksplit = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;
  if (planes == 1) {
    p.npass = 1; p.pa[0] = 0; p.pb[0] = 0;
  } else {
    // plane pairs ordered small -> large so the big products are added last
    const int ia[6] = {2, 0, 1, 1, 0, 0}, ib[6] = {0, 2, 1, 0, 1, 0};
    p.npass = 6;
    for (int i = 0; i < 6; ++i) { p.pa[i] = ia[i]; p.pb[i] = ib[i]; }
  }
  p.out = out; p.out2 = out2; p.out_ps = out_ps; p.ldo = ldo;
  p.aux = aux; p.aux_ps = aux_ps; p.ldaux = ldaux;
  p.bias = bias; p.outf = outf; p.ldof = ldof;
  // EPI_MUL + outf: outf[N] += column sums of the output.  Fused into the staged epilogue when that instantiation runs
  // (BN 256, one plane, N fits the smem accumulator); otherwise the column-sum kernel runs over `out` afterwards.
  p.drop = (epi == EPI_BIAS_DROP_ADD) ? *drop : no_drop();
  p.colsum = nullptr;
  const bool want_colsum = epi == SIMX_EPI_MUL && outf != nullptr;
  const bool fuse_colsum = want_colsum && BN == 256 && planes == 1 && !a_mn && N <= 4096;
  if (fuse_colsum) p.colsum = outf;
  *colsum_fused = fuse_colsum;

  ProfScope prof(a_mn ? "gemm_mn" : "gemm_kk", 2.0 * M * N * (double)K * p.npass,
                 2.0 * planes * ((double)M * K + (double)N * K) + (epi == SIMX_EPI_ATOMIC_F32 ? 4.0 : 2.0 * planes) * M * N, stream);
#define SIMX_GEMM_DISPATCH(BN_)                                                                  \
  if (BN == BN_) {                                                                               \
    if (a_mn) {                                                                                  \
      if (planes == 1) return dispatch_epi<BN_, true, 1>(epi, p, A, lda, a_ps, B, ldb, b_ps, planes, stream);                \
      return dispatch_epi<BN_, true, 3>(epi, p, A, lda, a_ps, B, ldb, b_ps, planes, stream);                                 \
    } else {                                                                                     \
      if (planes == 1) return dispatch_epi<BN_, false, 1>(epi, p, A, lda, a_ps, B, ldb, b_ps, planes, stream);               \
      return dispatch_epi<BN_, false, 3>(epi, p, A, lda, a_ps, B, ldb, b_ps, planes, stream);                                \
    }                                                                                            \
  }
  SIMX_GEMM_DISPATCH(256)
  SIMX_GEMM_DISPATCH(128)
  SIMX_GEMM_DISPATCH(64)
#undef SIMX_GEMM_DISPATCH
  SIMX_FAIL(SIMX_ERR_INVALID, "gemm: no tile configuration");
}

int colsum(int np, const bf16* x, long rows, int N, long row_stride, long ps, float* out, cudaStream_t st);

// Internal entry used by the encoder and by simx_gemm(): the GEMM, plus the column-sum pass when the EPI_MUL bias
// gradient could not ride in the epilogue.
// drop (optional): turns SIMX_EPI_BIAS_ADD into out = dropout(acc + bias) + aux.
int gemm(int M, int N, int K, const bf16* A, long lda, long a_ps, bool a_mn, const bf16* B, long ldb, long b_ps,
         bool b_mn, int planes, int epi, const float* bias, const bf16* aux, long ldaux, long aux_ps, bf16* out,
         bf16* out2, long ldo, long out_ps, float* outf, long ldof, int ksplit, cudaStream_t stream,
         const DropSpec* drop) {
  bool fused = false;
  SIMX_TRY(gemm_launch(M, N, K, A, lda, a_ps, a_mn, B, ldb, b_ps, b_mn, planes, epi, bias, aux, ldaux, aux_ps, out, out2,
                       ldo, out_ps, outf, ldof, ksplit, &fused, drop, stream));
  if (epi == SIMX_EPI_MUL && outf != nullptr && !fused) SIMX_TRY(colsum(planes, out, M, N, ldo, out_ps, outf, stream));
  return SIMX_OK;
}

}  // namespace simx

extern "C" int simx_gemm(int M, int N, int K, const void* A, int64_t lda, int64_t a_plane_stride, int a_mn_major,
                         const void* B, int64_t ldb, int64_t b_plane_stride, int b_mn_major, int planes, int epilogue,
                         const float* bias, const void* aux, int64_t ldaux, int64_t aux_plane_stride, void* out,
                         void* out2, int64_t ldo, int64_t out_plane_stride, float* outf, int64_t ldof, int ksplit,
                         void* stream) {
  using namespace simx;
  SIMX_CHECK(epilogue >= SIMX_EPI_PLAIN && epilogue <= SIMX_EPI_BIAS_GELU_FWD, SIMX_ERR_INVALID, "unknown GEMM epilogue %d",
             epilogue);
  return gemm(M, N, K, static_cast<const bf16*>(A), lda, a_plane_stride, a_mn_major != 0, static_cast<const bf16*>(B),
              ldb, b_plane_stride, b_mn_major != 0, planes, epilogue, bias, static_cast<const bf16*>(aux), ldaux,
              aux_plane_stride, static_cast<bf16*>(out), static_cast<bf16*>(out2), ldo, out_plane_stride, outf, ldof,
              ksplit, static_cast<cudaStream_t>(stream), nullptr);
}

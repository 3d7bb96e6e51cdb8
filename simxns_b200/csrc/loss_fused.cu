// The global in-batch InfoNCE step as ONE kernel launch (north_star: "the contrastive step ... is a single fused kernel"):
//   S = Q C^T over the all-gathered embeddings, row log-sum-exp, NLL of the positives, arg-max hits, and the gradients
//   dQ_local = dS[local rows, :] C,  dC_local = dS[:, local cols]^T Q  with dS = w (softmax(S) - onehot) / GQ
// (BiEncoderNllLoss.calc, SimANS/model/models.py:468-505, on gathered vectors as in MASTER/finetune/MS/run_de_model.py:180-234
// and LEAD/util.py:416-451) — without ever materialising the [GQ, GC] score matrix.
//
// A cooperative launch (one wave of CTAs, grid-wide barriers between phases):
//   phase 1  every score tile once: fp32 FFMA tile product, per-row (max, sum-exp, first arg-max) of the tile's columns
//            -> a [GQ, n_tiles] table of partials (8 bytes + 4 bytes per row and column tile), and the positive's score
//   phase 2  a thread per row merges its partials (ascending column tiles: the first maximal index wins, as torch.max) ->
//            row log-sum-exp, loss, hit count
//   phase 3  only the tiles that touch this rank's rows or columns are recomputed; dS of the tile goes to shared memory
//            and is contracted on the spot with the C rows (-> dQ) and / or the Q rows (-> dC) of the tile, hidden chunk
//            by hidden chunk, into the outputs with vector reductions (red.global.add.v4.f32)
// Everything is fp32 (the parity mode of the loss is the only mode: SURVEY §8a-5); the 3-GEMM path in loss.cu (tcgen05,
// split-bf16 planes, S and dS in memory) stays as the checker (SIMX_LOSS_UNFUSED=1).
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace simx {

struct FusedNllParams {
  const float* Q;       // [GQ, H]
  const float* C;       // [GC, H]
  const int32_t* pos;   // [GQ]
  int GQ, GC, H;
  int q_begin, q_count, c_begin, c_count;
  float w, grad_scale;
  float* out_losses;    // [12], zeroed by the caller: [0] += w nll, [1] += nll, [6] += hits
  float* dQ;            // [q_count, H]   (zeroed here)
  float* dC;            // [c_count, H]   (zeroed here)
  float2* part;         // [GQ, n_tiles] (max, sum-exp) per row and column tile
  int* part_arg;        // [GQ, n_tiles] first arg-max column inside the tile
  float* spos;          // [GQ] score of the positive
  float* lse;           // [GQ]
};

constexpr int FBK = 16;

template <int TM>
struct FusedTile {
  static constexpr int BT = 16 * TM;       // tile edge (rows and columns)
  static constexpr int LD = BT + 4;        // smem row pitch (floats)
  static constexpr size_t SMEM = (size_t)(4 * FBK * LD + BT * LD) * sizeof(float);  // A/B double buffers + the dS tile
  // index of micro-tile element i (0..TM) of thread coordinate t (0..15) inside the tile: 4-wide groups, 64 apart
  __device__ static __forceinline__ int at(int t, int i) {
    if constexpr (TM == 8) return (i < 4) ? t * 4 + i : 64 + t * 4 + (i - 4);
    else return t * 4 + i;
  }
};

// acc[i][j] += a[i] * b[j] with sm_100's packed FFMA2 (two fp32 FMAs per instruction, same rounding as fmaf): the scalar
// 3-register FFMA issues once per two cycles per scheduler, which held these tiles at ~40 TFLOP/s
template <int TM>
__device__ __forceinline__ void outer_fma(float (&acc)[TM][TM], const float (&a)[TM], const float (&b)[TM]) {
  float2(&acc2)[TM][TM / 2] = reinterpret_cast<float2(&)[TM][TM / 2]>(acc);
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const float2 ai = make_float2(a[i], a[i]);
#pragma unroll
    for (int j = 0; j < TM / 2; ++j) acc2[i][j] = __ffma2_rn(ai, make_float2(b[2 * j], b[2 * j + 1]), acc2[i][j]);
  }
}

// acc[i][j] = sum_k A[at(ty,i)][k] * B[at(tx,j)][k] over k < H; A, B row-major with pitch H, rows >= *_valid read as zero
template <int TM>
__device__ __forceinline__ void score_tile(float (&acc)[TM][TM], const float* __restrict__ A, int a_valid,
                                           const float* __restrict__ B, int b_valid, int H, float* As, float* Bs) {
  using T = FusedTile<TM>;
  constexpr int LD = T::LD, NV = TM / 4;  // float4 loads per thread and operand per k-tile
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TM; ++j) acc[i][j] = 0.f;
  float4 ra[NV], rb[NV];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int s = 0; s < NV; ++s) {
      const int idx = tid + s * 256, r = idx >> 2, k = k0 + (idx & 3) * 4;
      ra[s] = (r < a_valid && k < H) ? __ldg(reinterpret_cast<const float4*>(A + (long)r * H + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
      rb[s] = (r < b_valid && k < H) ? __ldg(reinterpret_cast<const float4*>(B + (long)r * H + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int s = 0; s < NV; ++s) {
      const int idx = tid + s * 256, r = idx >> 2, kq = (idx & 3) * 4;
      float* a = As + buf * FBK * LD + kq * LD + r;
      float* b = Bs + buf * FBK * LD + kq * LD + r;
      a[0] = ra[s].x; a[LD] = ra[s].y; a[2 * LD] = ra[s].z; a[3 * LD] = ra[s].w;
      b[0] = rb[s].x; b[LD] = rb[s].y; b[2 * LD] = rb[s].z; b[3 * LD] = rb[s].w;
    }
  };
  const int nk = (H + FBK - 1) / FBK;
  fetch(0);
  __syncthreads();  // the previous user of As / Bs is done
  stash(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) fetch((kt + 1) * FBK);
    const float* a_s = As + buf * FBK * LD;
    const float* b_s = Bs + buf * FBK * LD;
#pragma unroll
    for (int k = 0; k < FBK; ++k) {
      float a[TM], b[TM];
#pragma unroll
      for (int g = 0; g < TM / 4; ++g) {
        const float4 va = *reinterpret_cast<const float4*>(a_s + k * LD + g * 64 + ty * 4);
        const float4 vb = *reinterpret_cast<const float4*>(b_s + k * LD + g * 64 + tx * 4);
        a[4 * g] = va.x; a[4 * g + 1] = va.y; a[4 * g + 2] = va.z; a[4 * g + 3] = va.w;
        b[4 * g] = vb.x; b[4 * g + 1] = vb.y; b[4 * g + 2] = vb.z; b[4 * g + 3] = vb.w;
      }
      outer_fma<TM>(acc, a, b);
    }
    if (kt + 1 < nk) {
      stash(buf ^ 1);
      __syncthreads();
    }
  }
}

// out[at(ty,i)][h0 + at(tx,j)] += sum_k  Gop(i, k) * X[k][h0 + at(tx,j)]   for k < k_valid (X row-major, pitch H):
//   ROWS_OUT = true : Gop(i, k) = G[at(ty,i)][k]   (output rows = the tile's score rows  -> dQ, X = C rows of the tile)
//   ROWS_OUT = false: Gop(i, k) = G[k][at(ty,i)]   (output rows = the tile's score columns -> dC, X = Q rows of the tile)
template <int TM, bool ROWS_OUT>
__device__ __forceinline__ void grad_chunk(float (&acc)[TM][TM], const float* G, const float* __restrict__ X, int k_valid,
                                           int h0, int H, float* Bs) {
  using T = FusedTile<TM>;
  constexpr int LD = T::LD, BT = T::BT, NV = TM / 4;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TM; ++j) acc[i][j] = 0.f;
  float4 rb[NV];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int s = 0; s < NV; ++s) {
      const int idx = tid + s * 256, k = k0 + idx / (BT / 4), n = h0 + (idx % (BT / 4)) * 4;
      rb[s] = (k < k_valid && n < H) ? __ldg(reinterpret_cast<const float4*>(X + (long)k * H + n)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int s = 0; s < NV; ++s) {
      const int idx = tid + s * 256;
      *reinterpret_cast<float4*>(Bs + buf * FBK * LD + (idx / (BT / 4)) * LD + (idx % (BT / 4)) * 4) = rb[s];
    }
  };
  constexpr int nk = BT / FBK;
  fetch(0);
  __syncthreads();
  stash(0);
  __syncthreads();
#pragma unroll 1
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) fetch((kt + 1) * FBK);
    const float* b_s = Bs + buf * FBK * LD;
#pragma unroll
    for (int k = 0; k < FBK; ++k) {
      float a[TM], b[TM];
      const int kk = kt * FBK + k;
#pragma unroll
      for (int g = 0; g < TM / 4; ++g) {
        const float4 vb = *reinterpret_cast<const float4*>(b_s + k * LD + g * 64 + tx * 4);
        b[4 * g] = vb.x; b[4 * g + 1] = vb.y; b[4 * g + 2] = vb.z; b[4 * g + 3] = vb.w;
        if constexpr (ROWS_OUT) {
#pragma unroll
          for (int i = 0; i < 4; ++i) a[4 * g + i] = G[(g * 64 + ty * 4 + i) * LD + kk];
        } else {
          const float4 va = *reinterpret_cast<const float4*>(G + kk * LD + g * 64 + ty * 4);
          a[4 * g] = va.x; a[4 * g + 1] = va.y; a[4 * g + 2] = va.z; a[4 * g + 3] = va.w;
        }
      }
      outer_fma<TM>(acc, a, b);
    }
    if (kt + 1 < nk) {
      stash(buf ^ 1);
      __syncthreads();
    }
  }
}

__device__ __forceinline__ void red_add4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <int TM>
__global__ void __launch_bounds__(256)  // one CTA per SM at 187 registers (capping at 128 for two CTAs/SM spills: 2.93 vs 2.60 ms)
nll_global_fused_kernel(const FusedNllParams p) {
  using T = FusedTile<TM>;
  constexpr int BT = T::BT, LD = T::LD;
  extern __shared__ __align__(16) float fsm[];
  float* As = fsm;
  float* Bs = As + 2 * FBK * LD;
  float* G = Bs + 2 * FBK * LD;
  cg::grid_group grid = cg::this_grid();
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int GQ = p.GQ, GC = p.GC, H = p.H;
  const int m_tiles = (GQ + BT - 1) / BT, n_tiles = (GC + BT - 1) / BT;

  // ---- phase 0: the outputs this launch accumulates into ----
  for (long i = (long)blockIdx.x * 256 + tid; i < (long)p.q_count * H; i += (long)gridDim.x * 256) p.dQ[i] = 0.f;
  for (long i = (long)blockIdx.x * 256 + tid; i < (long)p.c_count * H; i += (long)gridDim.x * 256) p.dC[i] = 0.f;

  // ---- phase 1: every score tile once -> per-(row, column tile) partials ----
  for (int tile = blockIdx.x; tile < m_tiles * n_tiles; tile += gridDim.x) {
    const int mt = tile / n_tiles, nt = tile % n_tiles;
    const int r0 = mt * BT, c0 = nt * BT;
    float acc[TM][TM];
    score_tile<TM>(acc, p.Q + (long)r0 * H, GQ - r0, p.C + (long)c0 * H, GC - c0, H, As, Bs);
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      const int row = r0 + T::at(ty, i);
      const int pos = row < GQ ? p.pos[row] : -1;
      float m = -INFINITY;
      int am = 0x7fffffff;
#pragma unroll
      for (int j = 0; j < TM; ++j) {  // columns ascend with j inside a thread
        const int col = c0 + T::at(tx, j);
        if (col < GC) {
          if (acc[i][j] > m) { m = acc[i][j]; am = col; }
          if (col == pos) p.spos[row] = acc[i][j];
        }
      }
      float l = 0.f;
#pragma unroll
      for (int j = 0; j < TM; ++j)
        if (c0 + T::at(tx, j) < GC) l += expf(acc[i][j] - m);
      // the 16 threads of one ty hold the tile's columns of this row: butterfly over tx (lanes differ in bits 0..3)
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {
        const float m2 = __shfl_xor_sync(0xffffffffu, m, o), l2 = __shfl_xor_sync(0xffffffffu, l, o);
        const int a2 = __shfl_xor_sync(0xffffffffu, am, o);
        if (m2 > -INFINITY) {
          const float mn = fmaxf(m, m2);
          l = (m > -INFINITY ? l * expf(m - mn) : 0.f) + l2 * expf(m2 - mn);
          if (m2 > m || (m2 == m && a2 < am)) am = a2;
          m = mn;
        }
      }
      if (tx == 0 && row < GQ) {
        p.part[(long)row * n_tiles + nt] = make_float2(m, l);
        p.part_arg[(long)row * n_tiles + nt] = am;
      }
    }
  }
  __threadfence();
  grid.sync();

  // ---- phase 2: a thread per row merges its column-tile partials (ascending: the first maximal index wins) ----
  for (int row = blockIdx.x * 256 + tid; row < GQ; row += gridDim.x * 256) {
    float M = -INFINITY, L = 0.f;
    int A = 0x7fffffff;
    for (int nt = 0; nt < n_tiles; ++nt) {
      const float2 v = p.part[(long)row * n_tiles + nt];
      if (!(v.x > -INFINITY)) continue;
      const float mn = fmaxf(M, v.x);
      L = (M > -INFINITY ? L * expf(M - mn) : 0.f) + v.y * expf(v.x - mn);
      if (v.x > M) A = p.part_arg[(long)row * n_tiles + nt];
      M = mn;
    }
    const float lse = M + logf(L);
    p.lse[row] = lse;
    const float nll = lse - p.spos[row];
    atomicAdd(p.out_losses + 1, nll / GQ);
    atomicAdd(p.out_losses + 0, p.w * nll / GQ);
    if (A == p.pos[row]) atomicAdd(p.out_losses + 6, 1.f);
  }
  __threadfence();
  grid.sync();

  // ---- phase 3: gradients from the tiles that touch this rank's rows or columns ----
  const int q_end = p.q_begin + p.q_count, c_end = p.c_begin + p.c_count;
  const int mt_lo = p.q_begin / BT, mt_hi = (q_end - 1) / BT;  // row tiles holding local queries
  const int nt_lo = p.c_begin / BT, nt_hi = (c_end - 1) / BT;  // column tiles holding local passages
  const int n_rowtiles = (mt_hi - mt_lo + 1) * n_tiles;         // (local rows) x (all columns)
  const int n_coltiles = (m_tiles - (mt_hi - mt_lo + 1)) * (nt_hi - nt_lo + 1);  // (other rows) x (local columns)
  const float kscale = p.w * p.grad_scale / GQ;
  const int h_chunks = (H + BT - 1) / BT;
  for (int wi = blockIdx.x; wi < n_rowtiles + n_coltiles; wi += gridDim.x) {
    int mt, nt;
    if (wi < n_rowtiles) {
      mt = mt_lo + wi / n_tiles;
      nt = wi % n_tiles;
    } else {
      const int u = wi - n_rowtiles, ncl = nt_hi - nt_lo + 1;
      mt = u / ncl;
      if (mt >= mt_lo) mt += mt_hi - mt_lo + 1;  // skip the local row tiles (already covered above)
      nt = nt_lo + u % ncl;
    }
    const int r0 = mt * BT, c0 = nt * BT;
    const bool rows_local = mt >= mt_lo && mt <= mt_hi, cols_local = nt >= nt_lo && nt <= nt_hi;
    float acc[TM][TM];
    score_tile<TM>(acc, p.Q + (long)r0 * H, GQ - r0, p.C + (long)c0 * H, GC - c0, H, As, Bs);
    __syncthreads();  // G may still be read by the previous work item's last chunk
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      const int row = r0 + T::at(ty, i);
      const float lse = row < GQ ? p.lse[row] : 0.f;
      const int pos = row < GQ ? p.pos[row] : -1;
#pragma unroll
      for (int g = 0; g < TM / 4; ++g) {
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int col = c0 + g * 64 + tx * 4 + j;
          v[j] = (row < GQ && col < GC) ? kscale * (expf(acc[i][4 * g + j] - lse) - (col == pos ? 1.f : 0.f)) : 0.f;
        }
        *reinterpret_cast<float4*>(G + T::at(ty, i) * LD + g * 64 + tx * 4) = make_float4(v[0], v[1], v[2], v[3]);
      }
    }
    __syncthreads();
    if (rows_local) {  // dQ[row] += sum_col dS[row, col] C[col]
      for (int hc = 0; hc < h_chunks; ++hc) {
        grad_chunk<TM, true>(acc, G, p.C + (long)c0 * H, GC - c0, hc * BT, H, Bs);
#pragma unroll
        for (int i = 0; i < TM; ++i) {
          const int row = r0 + T::at(ty, i);
          if (row < p.q_begin || row >= q_end) continue;
#pragma unroll
          for (int g = 0; g < TM / 4; ++g) {
            const int h = hc * BT + g * 64 + tx * 4;
            if (h < H) red_add4(p.dQ + (long)(row - p.q_begin) * H + h, acc[i][4 * g], acc[i][4 * g + 1], acc[i][4 * g + 2], acc[i][4 * g + 3]);
          }
        }
      }
    }
    if (cols_local) {  // dC[col] += sum_row dS[row, col] Q[row]
      for (int hc = 0; hc < h_chunks; ++hc) {
        grad_chunk<TM, false>(acc, G, p.Q + (long)r0 * H, GQ - r0, hc * BT, H, Bs);
#pragma unroll
        for (int i = 0; i < TM; ++i) {
          const int col = c0 + T::at(ty, i);
          if (col < p.c_begin || col >= c_end) continue;
#pragma unroll
          for (int g = 0; g < TM / 4; ++g) {
            const int h = hc * BT + g * 64 + tx * 4;
            if (h < H) red_add4(p.dC + (long)(col - p.c_begin) * H + h, acc[i][4 * g], acc[i][4 * g + 1], acc[i][4 * g + 2], acc[i][4 * g + 3]);
          }
        }
      }
    }
  }
}

size_t nll_fused_workspace_bytes(int GQ, int GC) {
  const size_t nt = (GC + 63) / 64;  // the smaller tile edge bounds the table
  return 256 + (size_t)GQ * nt * (sizeof(float2) + sizeof(int)) + 2 * (size_t)GQ * sizeof(float) + 512;
}

template <int TM>
static int launch_fused(FusedNllParams p, cudaStream_t st) {
  auto kern = nll_global_fused_kernel<TM>;
  const size_t smem = FusedTile<TM>::SMEM;
  static bool attr = false;
  if (!attr) {
    SIMX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  int per_sm = 0;
  SIMX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, smem));
  SIMX_CHECK(per_sm >= 1, SIMX_ERR_CUDA, "fused loss kernel does not fit an SM");
  constexpr int BT = FusedTile<TM>::BT;
  const int tiles = ((p.GQ + BT - 1) / BT) * ((p.GC + BT - 1) / BT);
  int grid = min(tiles, per_sm * num_sms());  // one co-resident wave (cooperative launch)
  grid = max(grid, 1);
  void* args[] = {&p};
  SIMX_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(kern), dim3(grid), dim3(256), args, smem, st));
  return SIMX_OK;
}

// out_losses must be zeroed by the caller; dQ / dC are overwritten.
int nll_global_fused(const simx_loss_cfg* cfg, const float* Q, const float* C, const int32_t* pos_idx, float* out_losses,
                     float* dQ, float* dC, void* workspace, cudaStream_t st) {
  SIMX_CHECK(cfg->H % 4 == 0, SIMX_ERR_INVALID, "fused loss: H must be a multiple of 4");
  FusedNllParams p;
  p.Q = Q; p.C = C; p.pos = pos_idx;
  p.GQ = cfg->GQ; p.GC = cfg->GC; p.H = cfg->H;
  p.q_begin = cfg->q_begin; p.q_count = cfg->q_count; p.c_begin = cfg->c_begin; p.c_count = cfg->c_count;
  p.w = cfg->w_nll_global; p.grad_scale = cfg->grad_scale;
  p.out_losses = out_losses; p.dQ = dQ; p.dC = dC;
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  ws += (256 - (reinterpret_cast<uintptr_t>(ws) & 255)) & 255;
  const size_t nt = (cfg->GC + 63) / 64;
  p.part = reinterpret_cast<float2*>(ws);
  p.part_arg = reinterpret_cast<int*>(ws + (size_t)cfg->GQ * nt * sizeof(float2));
  p.spos = reinterpret_cast<float*>(ws + (size_t)cfg->GQ * nt * (sizeof(float2) + sizeof(int)));
  p.lse = p.spos + cfg->GQ;
  ProfScope prof("nll_global_fused", 2.0 * cfg->GQ * (double)cfg->GC * cfg->H, 4.0 * ((double)cfg->GQ + cfg->GC) * cfg->H, st);
  // 128-wide tiles when they fill the machine, 64-wide ones for small score matrices (more CTAs in flight)
  const long tiles128 = (long)((cfg->GQ + 127) / 128) * ((cfg->GC + 127) / 128);
  if (tiles128 >= num_sms()) return launch_fused<8>(p, st);
  return launch_fused<4>(p, st);
}

}  // namespace simx

// Exact inner-product top-k over a resident passage-embedding shard: the GPU side of the ANN refresh of
// every SimANS round — faiss.IndexFlatIP + index_cpu_to_all_gpus(shard=True) + .search(q, 200 | 1000)
// (SimANS/co_training/co_training_generate.py:359-384, 412-427).
//
//   corpus  : kept in HBM as 3 bf16 planes per value (fp32-class, 6 B instead of 4 B per element) — the B operand of the
//             tcgen05 GEMM, K-major as stored, never re-converted;
//   scores  : S[nq, chunk] = Q P_chunk^T on the tensor cores (6 plane-pair passes, fp32 accumulation in TMEM), written
//             once as 3 planes;
//   select  : one CTA per (query, 16384-column segment): the segment is read ONCE from HBM into shared memory as
//             order-preserving 32-bit keys relative to the segment minimum, a radix select over the bits the segment's
//             range actually spans (11 bits per pass, warp-aggregated shared-memory atomics) finds the k-th largest, the winners
//             (ties: lowest index first) go to a candidate list; candidate lists are reduced by the same kernel until one
//             segment is left, which is also sorted (bitonic, (score desc, index asc)) and written as (fp32, int64).
// HBM-bound: 12 B per (query, passage) pair for the score round trip; the select itself runs out of shared memory.
#include <algorithm>

#include "common.cuh"
#include "ptx.cuh"

namespace simx {

int convert_weight(int np, const float* w, int R, int C, bf16* dst, long ld_dst, long ps_dst, bf16* dst_t, long ld_t,
                   long ps_t, cudaStream_t st);

constexpr int TOPK_SEG = 16384;      // elements of one select segment (64 KB of keys in shared memory: two CTAs per SM,
                                     // one loading its segment from HBM while the other selects)
constexpr int TOPK_THREADS = 512;    // 32 elements per thread
constexpr int TOPK_MAX_K = 1024;
constexpr int TOPK_CHUNK = 8 * TOPK_SEG;  // passages per score GEMM
static_assert(TOPK_SEG == 32 * TOPK_THREADS, "one 32-bit flag word per thread covers its elements");

__device__ __forceinline__ uint32_t float_to_key(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // monotone: larger float <=> larger key (-0 < +0)
}
__device__ __forceinline__ float key_to_float(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

struct SelectArgs {
  const bf16* planes;   // phase 1 input: 3-plane scores [rows, ld] (plane stride ps), or NULL
  long ps;
  const float* vals;    // phase 2 input: fp32 candidates [rows, ld]
  const int32_t* idx;   // phase 2: their passage indices
  long ld;              // row stride of the input
  int n;                // valid elements per row
  int idx_base;         // phase 1: passage index of column 0
  int k;
  float* out_vals;      // [rows, out_ld]: segment s writes k entries at s * k
  int32_t* out_idx;
  long out_ld;
  float* final_vals;    // when non-NULL (single segment): sorted results [rows, k]
  int64_t* final_idx;
  long long* timing;    // debug: clock64 stamps of one CTA (simx_debug_topk_timing), else NULL
};
static long long* g_topk_timing = nullptr;
#define TOPK_STAMP(i) do { if (a.timing != nullptr && tid == 0 && blockIdx.x == 1 && blockIdx.y == gridDim.y / 2) a.timing[i] = clock64(); } while (0)

// Block-wide: the bin, counted from the top, in which the `want`-th largest element lies, how many elements lie in the
// bins above it, and its own population -> s_out[0..2].  Thread t owns 4 consecutive bins, top-down.  Ends with a barrier.
__device__ __forceinline__ void find_bin(const uint32_t* hist, uint32_t want, int tid, uint32_t* s_wtot, uint32_t* s_out) {
  static_assert(TOPK_THREADS * 4 == 2048, "4 bins per thread");
  const int lane = tid & 31, warp = tid >> 5;
  const int hi = 2048 - 4 * tid;  // owns bins [hi-4, hi), scanned from hi-1 down
  uint32_t c[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) c[i] = hist[hi - 1 - i];
  const uint32_t mine = c[0] + c[1] + c[2] + c[3];
  uint32_t incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_wtot[warp] = incl;
  __syncthreads();
  uint32_t before = 0;
  for (int wv = 0; wv < warp; ++wv) before += s_wtot[wv];
  uint32_t acc = before + incl - mine;
  if (acc < want && want <= acc + mine) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (acc + c[i] >= want) {
        s_out[0] = (uint32_t)(hi - 1 - i);
        s_out[1] = acc;
        s_out[2] = c[i];
        break;
      }
      acc += c[i];
    }
  }
  __syncthreads();
}

// bitonic sort, descending, of n (power of two) 64-bit words in shared memory by the whole block
__device__ __forceinline__ void bitonic_desc(unsigned long long* a, int n, int tid) {
  for (int size = 2; size <= n; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = tid; i < n; i += TOPK_THREADS) {
        const int partner = i ^ stride;
        if (partner > i) {
          const unsigned long long x = a[i], y = a[partner];
          const bool desc = (i & size) == 0;
          if (desc ? x < y : x > y) { a[i] = y; a[partner] = x; }
        }
      }
      __syncthreads();
    }
  }
}

constexpr int TOPK_CONTENDERS = 2048;  // elements of the threshold bin the fast path sorts directly

__global__ void __launch_bounds__(TOPK_THREADS, 2) topk_select_kernel(const SelectArgs a) {
  extern __shared__ uint32_t sm[];
  uint32_t* keys = sm;                    // [TOPK_SEG]
  uint32_t* hist = sm + TOPK_SEG;         // [2048]
  unsigned long long* cont = reinterpret_cast<unsigned long long*>(sm + TOPK_SEG + 2048);  // [TOPK_CONTENDERS]
  __shared__ uint32_t s_ties_base, s_min, s_max, s_sel[3];
  __shared__ uint32_t s_scan[TOPK_THREADS / 32], s_red[2][TOPK_THREADS / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int seg = blockIdx.x;
  const long row = blockIdx.y;
  const int begin = seg * TOPK_SEG;
  const int m = min(TOPK_SEG, a.n - begin);
  const int k = a.k;
  float* ov = a.out_vals ? a.out_vals + row * a.out_ld + (long)seg * k : nullptr;
  int32_t* oi = a.out_idx ? a.out_idx + row * a.out_ld + (long)seg * k : nullptr;
  const bool fin = a.final_vals != nullptr;

  // ---- load the segment once: HBM -> order-preserving keys in shared memory (min / max on the way) ----
  uint32_t lo = 0xffffffffu, hi = 0u;
  if (a.planes != nullptr) {
    const bf16* p = a.planes + row * a.ld + begin;
    // full 8-element groups: one 16-byte load per plane (ld, ps and begin are multiples of 8, the buffer is 256-byte aligned)
    const int m8 = m & ~7;
#pragma unroll 4
    for (int j = tid * 8; j < m8; j += TOPK_THREADS * 8) {
      const uint4 x0 = *reinterpret_cast<const uint4*>(p + j);
      const uint4 x1 = *reinterpret_cast<const uint4*>(p + a.ps + j);
      const uint4 x2 = *reinterpret_cast<const uint4*>(p + 2 * a.ps + j);
      const uint32_t w0[4] = {x0.x, x0.y, x0.z, x0.w}, w1[4] = {x1.x, x1.y, x1.z, x1.w}, w2[4] = {x2.x, x2.y, x2.z, x2.w};
      uint32_t kk[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f0 = unpack_bf16x2(w0[i]), f1 = unpack_bf16x2(w1[i]), f2 = unpack_bf16x2(w2[i]);
        kk[2 * i] = float_to_key(f0.x + f1.x + f2.x);
        kk[2 * i + 1] = float_to_key(f0.y + f1.y + f2.y);
        lo = min(lo, min(kk[2 * i], kk[2 * i + 1]));
        hi = max(hi, max(kk[2 * i], kk[2 * i + 1]));
      }
      *reinterpret_cast<uint4*>(keys + j) = make_uint4(kk[0], kk[1], kk[2], kk[3]);
      *reinterpret_cast<uint4*>(keys + j + 4) = make_uint4(kk[4], kk[5], kk[6], kk[7]);
    }
    for (int j = m8 + tid * 2; j < m; j += TOPK_THREADS * 2) {  // ragged tail: bf16 pairs
      if (j + 1 < m) {
        const __nv_bfloat162 x0 = *reinterpret_cast<const __nv_bfloat162*>(p + j);
        const __nv_bfloat162 x1 = *reinterpret_cast<const __nv_bfloat162*>(p + a.ps + j);
        const __nv_bfloat162 x2 = *reinterpret_cast<const __nv_bfloat162*>(p + 2 * a.ps + j);
        const uint32_t k0 = float_to_key(__low2float(x0) + __low2float(x1) + __low2float(x2));
        const uint32_t k1 = float_to_key(__high2float(x0) + __high2float(x1) + __high2float(x2));
        *reinterpret_cast<uint2*>(keys + j) = make_uint2(k0, k1);
        lo = min(lo, min(k0, k1));
        hi = max(hi, max(k0, k1));
      } else {
        const uint32_t k0 = float_to_key(__bfloat162float(p[j]) + __bfloat162float(p[a.ps + j]) + __bfloat162float(p[2 * a.ps + j]));
        keys[j] = k0;
        lo = min(lo, k0);
        hi = max(hi, k0);
      }
    }
  } else {
    const float* p = a.vals + row * a.ld + begin;
    for (int j = tid; j < m; j += TOPK_THREADS) {
      const uint32_t k0 = float_to_key(p[j]);
      keys[j] = k0;
      lo = min(lo, k0);
      hi = max(hi, k0);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  TOPK_STAMP(0);
  if (lane == 0) { s_red[0][warp] = lo; s_red[1][warp] = hi; }
  if (tid == 0) s_ties_base = 0;
  __syncthreads();
  if (tid == 0) {
    for (int wv = 1; wv < TOPK_THREADS / 32; ++wv) { lo = min(lo, s_red[0][wv]); hi = max(hi, s_red[1][wv]); }
    s_min = lo;
    s_max = hi;
  }
  __syncthreads();

  auto index_of = [&](int j) -> int32_t {
    return a.idx != nullptr ? a.idx[row * a.ld + begin + j] : a.idx_base + begin + j;
  };

  // winners are staged in `hist` (2048 words: k relative keys + k positions, k <= 1024) once the histogram is dead
  uint32_t* w_key = hist;
  uint32_t* w_pos = hist + TOPK_MAX_K;
  const int take = min(k, m);
  // keys are handled relative to the segment minimum: the radix digits then cover only the bits the scores actually
  // span, so the elements spread over the 2048 bins instead of piling into the few bins a shared exponent would give
  const uint32_t kmin = m > k ? s_min : 0u;

  if (m > k) {
    const uint32_t range = s_max - kmin;
    const int top = 32 - __clz(range);  // significant bits of the relative keys (0: all elements equal)
    const int sh = max(top - 11, 0);
    // ---- pass 1: histogram of the top 11 significant bits ----
    for (int b = tid; b < 2048; b += TOPK_THREADS) hist[b] = 0;
    __syncthreads();
    TOPK_STAMP(1);
    {
      const int m4 = m & ~3;
      for (int j = tid * 4; j < m4; j += TOPK_THREADS * 4) {
        const uint4 q = *reinterpret_cast<const uint4*>(keys + j);
        atomicAdd(&hist[(q.x - kmin) >> sh], 1u);
        atomicAdd(&hist[(q.y - kmin) >> sh], 1u);
        atomicAdd(&hist[(q.z - kmin) >> sh], 1u);
        atomicAdd(&hist[(q.w - kmin) >> sh], 1u);
      }
      for (int j = m4 + tid; j < m; j += TOPK_THREADS) atomicAdd(&hist[(keys[j] - kmin) >> sh], 1u);
    }
    __syncthreads();
    TOPK_STAMP(2);
    find_bin(hist, (uint32_t)k, tid, s_scan, s_sel);
    TOPK_STAMP(3);
    const uint32_t bin = s_sel[0], above = s_sel[1], cb = s_sel[2];
    __syncthreads();  // everyone has read the histogram's verdict: `hist` becomes the winner staging area
    if (cb <= TOPK_CONTENDERS) {
      // ---- fast path: bins above the threshold bin win outright; the threshold bin's few elements are sorted by
      //      (key desc, passage index asc) and the best `k - above` of them complete the list.  Compaction without
      //      atomics: a flag word per thread for its 32 elements, a block scan of the counts, then ordered writes ----
      // sweep 1 (branch-free, 3 instructions per element): flag everything at or above the threshold bin
      uint32_t fmask = 0;
      const uint32_t rel_thr = bin << sh;  // almost every element is below it
      const int m4 = m & ~3;
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const int j = (tid + g * TOPK_THREADS) * 4;
        if (j < m4) {
          const uint4 q = *reinterpret_cast<const uint4*>(keys + j);
          fmask |= (q.x - kmin >= rel_thr ? 1u : 0u) << (g * 4);
          fmask |= (q.y - kmin >= rel_thr ? 1u : 0u) << (g * 4 + 1);
          fmask |= (q.z - kmin >= rel_thr ? 1u : 0u) << (g * 4 + 2);
          fmask |= (q.w - kmin >= rel_thr ? 1u : 0u) << (g * 4 + 3);
        } else if (j < m) {  // the ragged last group of a row
          for (int e = 0; e < m - j; ++e) fmask |= (keys[j + e] - kmin >= rel_thr ? 1u : 0u) << (g * 4 + e);
        }
      }
      // the few flagged ones: winner (bin above) or contender (threshold bin)
      uint32_t wmask = 0, cmask = 0;
      for (uint32_t f = fmask; f; f &= f - 1) {
        const int i = __ffs(f) - 1;
        const int j = (tid + (i >> 2) * TOPK_THREADS) * 4 + (i & 3);
        if (((keys[j] - kmin) >> sh) > bin) wmask |= 1u << i;
        else cmask |= 1u << i;
      }
      const uint32_t mine = (uint32_t)__popc(wmask) | ((uint32_t)__popc(cmask) << 16);  // both totals < 65536
      uint32_t incl = mine;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      if (lane == 31) s_scan[warp] = incl;
      __syncthreads();
      uint32_t off = incl - mine;
      for (int wv = 0; wv < warp; ++wv) off += s_scan[wv];
      TOPK_STAMP(7);
      uint32_t wo = off & 0xffffu, co = off >> 16;
      while (wmask) {
        const int i = __ffs(wmask) - 1;
        wmask &= wmask - 1;
        const int j = (tid + (i >> 2) * TOPK_THREADS) * 4 + (i & 3);
        w_key[wo] = keys[j] - kmin;
        w_pos[wo++] = (uint32_t)j;
      }
      while (cmask) {
        const int i = __ffs(cmask) - 1;
        cmask &= cmask - 1;
        const int j = (tid + (i >> 2) * TOPK_THREADS) * 4 + (i & 3);
        // low word: position for now; replaced by the passage index below (a global load in the reduce rounds)
        cont[co++] = ((unsigned long long)(keys[j] - kmin) << 32) | (uint32_t)j;
      }
      __syncthreads();
      TOPK_STAMP(4);
      int P = 32;
      while (P < (int)cb) P <<= 1;
      for (int i = tid; i < P; i += TOPK_THREADS) {
        if (i < (int)cb) {
          const unsigned long long v = cont[i];
          const uint32_t pos = (uint32_t)(v & 0xffffffffu);
          // low word := inverted passage index, so that equal keys sort by ascending passage index
          const uint32_t inv = 0x7fffffffu - (uint32_t)max(index_of((int)pos), 0);
          cont[i] = (v & 0xffffffff00000000ull) | inv;
        } else {
          cont[i] = 0ull;
        }
      }
      __syncthreads();
      const uint32_t need = (uint32_t)k - above;
      auto place = [&](unsigned long long v, uint32_t r) {
        const int32_t id = (int32_t)(0x7fffffffu - (uint32_t)(v & 0xffffffffu));
        w_key[above + r] = (uint32_t)(v >> 32);
        // position inside the segment, recovered from the passage index in the score rounds; in the reduce rounds the
        // index itself is carried (flagged by the top bit) because candidates are not laid out in index order
        w_pos[above + r] = a.idx != nullptr ? (0x80000000u | (uint32_t)id) : (uint32_t)(id - a.idx_base - begin);
      };
      if (cb <= TOPK_THREADS) {
        // a handful of contenders (the usual case): rank by counting — thread i counts the contenders that beat its own
        // (all reads of cont[j] are broadcasts); the composites are distinct, so ranks are a permutation
        if (tid < (int)cb) {
          const unsigned long long mine_v = cont[tid];
          uint32_t r = 0;
          for (int j = 0; j < (int)cb; ++j) r += cont[j] > mine_v ? 1u : 0u;
          if (r < need) place(mine_v, r);
        }
      } else {
        bitonic_desc(cont, P, tid);
        for (int i = tid; i < (int)need; i += TOPK_THREADS) place(cont[i], (uint32_t)i);
      }
    } else {
      // ---- slow path (more than 2048 elements share the threshold bin: heavy duplication): finish the radix select
      //      on the remaining low bits, then emit with an ordered scan for the ties ----
      for (int j = tid; j < m; j += TOPK_THREADS) keys[j] -= kmin;
      __syncthreads();
      uint32_t prefix = bin << sh, mask = 2047u << sh, want = (uint32_t)k - above;
      int rem = sh;
#pragma unroll 1
      while (rem > 0) {
        const int wbits = min(11, rem), s2 = rem - wbits;
        const uint32_t dmask = (1u << wbits) - 1u;
        for (int b = tid; b < 2048; b += TOPK_THREADS) hist[b] = 0;
        __syncthreads();
        for (int base = 0; base < m; base += TOPK_THREADS) {
          const int j = base + tid;
          const uint32_t key = j < m ? keys[j] : 0u;
          const bool act = j < m && (key & mask) == prefix;
          const uint32_t digit = (key >> s2) & dmask;
          const uint32_t am = __ballot_sync(0xffffffffu, act);
          if (act) {  // one shared-memory atomic per distinct digit per warp
            const uint32_t peers = __match_any_sync(am, digit);
            if (lane == __ffs(peers) - 1) atomicAdd(&hist[digit], (uint32_t)__popc(peers));
          }
        }
        __syncthreads();
        find_bin(hist, want, tid, s_scan, s_sel);
        prefix |= s_sel[0] << s2;
        want -= s_sel[1];
        mask |= dmask << s2;
        rem = s2;
        __syncthreads();
      }
      const uint32_t thr = prefix, n_greater = (uint32_t)k - want;
      if (tid == 0) s_sel[0] = 0;
      __syncthreads();
      for (int j = tid; j < m; j += TOPK_THREADS) {
        const uint32_t key = keys[j];
        if (key > thr) {
          const uint32_t slot = atomicAdd(&s_sel[0], 1u);
          w_key[slot] = key;
          w_pos[slot] = (uint32_t)j;
        }
      }
      __syncthreads();
      // ties: block-wide ordered compaction, lowest position first (= lowest passage index in the score rounds)
      const uint32_t need = want;
      for (int base = 0; base < m; base += TOPK_THREADS) {
        const int j = base + tid;
        const bool tie = j < m && keys[j] == thr;
        const uint32_t bal = __ballot_sync(0xffffffffu, tie);
        if (lane == 0) s_scan[warp] = __popc(bal);
        __syncthreads();
        uint32_t before = s_ties_base;
        for (int wv = 0; wv < warp; ++wv) before += s_scan[wv];
        const uint32_t rank = before + __popc(bal & ((1u << lane) - 1u));
        if (tie && rank < need) {
          w_key[n_greater + rank] = thr;
          w_pos[n_greater + rank] = (uint32_t)j;
        }
        __syncthreads();
        if (tid == 0) {
          uint32_t tot = 0;
          for (int wv = 0; wv < TOPK_THREADS / 32; ++wv) tot += s_scan[wv];
          s_ties_base += tot;
        }
        __syncthreads();
        if (s_ties_base >= need) break;
      }
    }
  } else {
    for (int j = tid; j < m; j += TOPK_THREADS) {
      w_key[j] = keys[j];
      w_pos[j] = (uint32_t)j;
    }
  }
  __syncthreads();

  TOPK_STAMP(5);
  auto winner_index = [&](uint32_t wp) -> int32_t {
    return (wp & 0x80000000u) ? (int32_t)(wp & 0x7fffffffu) : index_of((int)wp);
  };

  if (!fin) {
    for (int t = tid; t < k; t += TOPK_THREADS) {
      if (t < take) {
        ov[t] = key_to_float(w_key[t] + kmin);
        oi[t] = winner_index(w_pos[t]);
      } else {
        ov[t] = -INFINITY;  // fewer than k elements in this segment: padding never wins a later round
        oi[t] = -1;
      }
    }
    TOPK_STAMP(6);
    return;
  }

  // ---- final segment: bitonic sort of (key desc, passage index asc) over 1024 slots, then write fp32 / int64 ----
  unsigned long long* srt = reinterpret_cast<unsigned long long*>(keys);  // the keys are dead: reuse their storage
  int P = 32;
  while (P < take) P <<= 1;
  for (int t = tid; t < P; t += TOPK_THREADS) {
    unsigned long long v = 0ull;  // padding sorts last
    if (t < take) {
      const uint32_t key = w_key[t] + kmin;
      const int32_t id = winner_index(w_pos[t]);
      v = ((unsigned long long)key << 32) | (uint32_t)(0x7fffffff - max(id, 0));
      if (id < 0) v = 0ull;  // padding entries of an earlier round
    }
    srt[t] = v;  // `keys` (the source of nothing read above) is dead: w_key / w_pos live in `hist`
  }
  __syncthreads();
  bitonic_desc(srt, P, tid);
  for (int t = tid; t < k; t += TOPK_THREADS) {
    const unsigned long long v = t < P ? srt[t] : 0ull;
    const bool valid = v != 0ull;
    a.final_vals[row * k + t] = valid ? key_to_float((uint32_t)(v >> 32)) : -INFINITY;
    a.final_idx[row * k + t] = valid ? (int64_t)(0x7fffffff - (uint32_t)(v & 0xffffffffu)) : -1;
  }
}

struct TopkWs {
  size_t qp, s, cand_v[2], cand_i[2], total;
  long cand_ld;
  int chunk;
};
static inline size_t al256(size_t x) { return (x + 255) & ~size_t(255); }
static TopkWs topk_ws(int nq, int64_t N, int H, int k) {
  TopkWs w{};
  w.chunk = (int)std::min<int64_t>(TOPK_CHUNK, (N + 7) / 8 * 8);
  const int64_t n_chunks = (N + w.chunk - 1) / w.chunk;
  const int64_t segs = n_chunks * ((w.chunk + TOPK_SEG - 1) / TOPK_SEG);
  w.cand_ld = segs * k;
  size_t off = 0;
  w.qp = off; off += al256(3ull * nq * H * 2);
  w.s = off;  off += al256(3ull * nq * (size_t)w.chunk * 2);
  for (int i = 0; i < 2; ++i) {
    const size_t ld = i == 0 ? (size_t)w.cand_ld : (size_t)((w.cand_ld + TOPK_SEG - 1) / TOPK_SEG) * k;
    w.cand_v[i] = off; off += al256(4ull * nq * ld);
    w.cand_i[i] = off; off += al256(4ull * nq * ld);
  }
  w.total = off + 256;
  return w;
}

}  // namespace simx

using namespace simx;

// rows are padded to a multiple of 8 per plane: the score GEMM wants N % 8 == 0 and may touch up to 7 rows past the end
static inline int64_t pad8(int64_t n) { return (n + 7) / 8 * 8; }

extern "C" size_t simx_topk_corpus_bytes(int64_t N, int H) {
  return (N <= 0 || H <= 0) ? 0 : 3ull * (size_t)pad8(N) * H * 2;
}

extern "C" int simx_topk_add(const float* P, int64_t n_rows, int64_t row_offset, int64_t N_total, int H,
                             void* corpus_planes, void* stream) {
  SIMX_CHECK(P && corpus_planes && n_rows > 0 && row_offset >= 0 && row_offset + n_rows <= N_total, SIMX_ERR_INVALID,
             "topk_add: bad slab [%lld, +%lld) of %lld rows", (long long)row_offset, (long long)n_rows, (long long)N_total);
  SIMX_CHECK(H > 0 && H % 64 == 0, SIMX_ERR_INVALID, "topk_add: H=%d must be a multiple of 64", H);
  SIMX_CHECK((reinterpret_cast<uintptr_t>(corpus_planes) & 255) == 0, SIMX_ERR_INVALID, "topk_add: corpus buffer must be 256-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  bf16* dst = static_cast<bf16*>(corpus_planes) + row_offset * H;
  const long ps = (long)(pad8(N_total) * H);
  const int64_t slab = 1 << 20;  // grid.y limit of the converter
  for (int64_t r = 0; r < n_rows; r += slab) {
    const int rows = (int)std::min<int64_t>(slab, n_rows - r);
    SIMX_TRY(convert_weight(3, P + r * H, rows, H, dst + r * H, H, ps, nullptr, 0, 0, st));
  }
  if (row_offset + n_rows == N_total && pad8(N_total) != N_total)  // the padding rows score 0 and are never selected
    for (int p = 0; p < 3; ++p)
      SIMX_CUDA(cudaMemsetAsync(static_cast<bf16*>(corpus_planes) + p * ps + N_total * H, 0,
                                (size_t)(pad8(N_total) - N_total) * H * 2, st));
  return SIMX_OK;
}

extern "C" size_t simx_topk_workspace_bytes(int nq, int64_t N, int H, int k) {
  if (nq <= 0 || N <= 0 || H <= 0 || k <= 0) return 0;
  return topk_ws(nq, N, H, k).total;
}

extern "C" int simx_topk_search(const float* Q, int nq, const void* corpus_planes, int64_t N, int H, int k,
                                float* out_scores, int64_t* out_idx, void* workspace, void* stream) {
  SIMX_CHECK(Q && corpus_planes && out_scores && out_idx && workspace, SIMX_ERR_INVALID, "topk_search: null argument");
  SIMX_CHECK(nq > 0 && N > 0 && N < 0x7fffffff, SIMX_ERR_INVALID, "topk_search: need 0 < nq, 0 < N < 2^31");
  SIMX_CHECK(k >= 1 && k <= TOPK_MAX_K, SIMX_ERR_INVALID, "topk_search: k=%d outside [1, %d]", k, TOPK_MAX_K);
  SIMX_CHECK(H > 0 && H % 64 == 0, SIMX_ERR_INVALID, "topk_search: H=%d must be a multiple of 64", H);
  SIMX_CHECK(nq <= 65535, SIMX_ERR_INVALID, "topk_search: at most 65535 queries per call (got %d)", nq);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const TopkWs w = topk_ws(nq, N, H, k);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  ws += (256 - (reinterpret_cast<uintptr_t>(ws) & 255)) & 255;
  bf16* Qp = reinterpret_cast<bf16*>(ws + w.qp);
  bf16* S = reinterpret_cast<bf16*>(ws + w.s);
  float* cv[2] = {reinterpret_cast<float*>(ws + w.cand_v[0]), reinterpret_cast<float*>(ws + w.cand_v[1])};
  int32_t* ci[2] = {reinterpret_cast<int32_t*>(ws + w.cand_i[0]), reinterpret_cast<int32_t*>(ws + w.cand_i[1])};
  const bf16* corpus = static_cast<const bf16*>(corpus_planes);
  const size_t smem = (TOPK_SEG + 2048) * sizeof(uint32_t) + TOPK_CONTENDERS * sizeof(unsigned long long);
  static bool attr = false;
  if (!attr) {
    SIMX_CUDA(cudaFuncSetAttribute(topk_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  SIMX_TRY(convert_weight(3, Q, nq, H, Qp, H, (long)nq * H, nullptr, 0, 0, st));
  const bool single = N <= TOPK_SEG;  // one segment: select, sort and write the results straight from the scores

  const int segs_per_chunk = (w.chunk + TOPK_SEG - 1) / TOPK_SEG;
  int chunk_i = 0;
  for (int64_t c0 = 0; c0 < N; c0 += w.chunk, ++chunk_i) {
    const int cols = (int)std::min<int64_t>(w.chunk, N - c0);
    const int cols8 = (cols + 7) / 8 * 8;  // the GEMM wants N % 8 == 0: the last chunk also scores the (zero) padding
    // rows of the corpus buffer; the select reads only `cols` columns
    {
      ProfScope prof("topk_scores", 12.0 * nq * (double)cols * H, 6.0 * ((double)cols * H + (double)nq * cols), st);
      SIMX_TRY(gemm(nq, cols8, H, Qp, H, (long)nq * H, false, corpus + c0 * H, H, (long)(pad8(N) * H), false, 3, SIMX_EPI_PLAIN,
                    nullptr, nullptr, 0, 0, S, nullptr, w.chunk, (long)nq * w.chunk, nullptr, 0, 1, st));
    }
    SelectArgs a{};
    a.timing = g_topk_timing;
    a.planes = S; a.ps = (long)nq * w.chunk; a.ld = w.chunk; a.n = cols; a.idx_base = (int)c0; a.k = k;
    if (single) {
      a.final_vals = out_scores; a.final_idx = out_idx;
    } else {
      a.out_vals = cv[0] + (long)chunk_i * segs_per_chunk * k;
      a.out_idx = ci[0] + (long)chunk_i * segs_per_chunk * k;
      a.out_ld = w.cand_ld;
    }
    const int segs = (cols + TOPK_SEG - 1) / TOPK_SEG;
    ProfScope prof("topk_select", 0, 6.0 * nq * (double)cols, st);
    topk_select_kernel<<<dim3(segs, nq), TOPK_THREADS, smem, st>>>(a);
    SIMX_LAUNCH_CHECK();
  }
  if (!single) {
    // candidates actually written: full chunks contribute segs_per_chunk * k each, the last chunk its own segments
    const int64_t full_chunks = N / w.chunk;
    const int64_t tail = N - full_chunks * w.chunk;
    long n_cand = (long)(full_chunks * segs_per_chunk + (tail + TOPK_SEG - 1) / TOPK_SEG) * k;
    int cur = 0;
    long ld = w.cand_ld;
    for (;;) {
      const int segs = (int)((n_cand + TOPK_SEG - 1) / TOPK_SEG);
      SelectArgs a{};
      a.vals = cv[cur]; a.idx = ci[cur]; a.ld = ld; a.n = (int)n_cand; a.k = k;
      if (segs == 1) {
        a.final_vals = out_scores; a.final_idx = out_idx;
      } else {
        a.out_vals = cv[cur ^ 1]; a.out_idx = ci[cur ^ 1]; a.out_ld = (long)segs * k;
      }
      ProfScope prof("topk_reduce", 0, 8.0 * nq * (double)n_cand, st);
      topk_select_kernel<<<dim3(segs, nq), TOPK_THREADS, smem, st>>>(a);
      SIMX_LAUNCH_CHECK();
      if (segs == 1) break;
      ld = (long)segs * k;
      n_cand = ld;
      cur ^= 1;
    }
  }
  return SIMX_OK;
}

// debug hook (not part of the public header): device buffer of >= 8 int64 receiving clock64 stamps of one score-round CTA
extern "C" void simx_debug_topk_timing(long long* device_buffer) { simx::g_topk_timing = device_buffer; }

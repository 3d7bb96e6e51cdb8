// SimANS ambiguous-negative sampling on the device (C ABI: simx_simans_sample).
//
// Reference: Rocketqa_v2Dataset.__getitem__ (SimANS/utils/MARCO_until_new.py:165-202; tau form) and TraditionDataset
// (SimANS/utils/util_wiki.py:613-639; a, b form), statement of the method in SimANS/README.md:12-24.  Per query, from
// <= 1024 candidates with retriever scores s_i and the positive's score s+:
//     pos_score == 0            -> the last k candidates (no sampling)
//     else  w_i = exp(-tau |s_i - s+|)      (mode 0)    or    exp(-a (s_i - s+ + b)^2)   (mode 1)
//           repeat: draw k candidates WITH replacement by weight, add the distinct ones to the chosen set, drop them from
//           the pool; until >= k are chosen; keep k of them.
// The reference runs this in DataLoader workers with Python's Mersenne twister, so bit-level RNG parity is impossible by
// construction; parity is at the level of the sampling distribution.  One quirk matters for that
// distribution: a round can overshoot k (up to 2k-1 distinct candidates) and the reference then keeps `list(set)[:k]`.
// For a CPython set of 19..85 small ints the hash table has 128 slots and hash(i) == i, so iteration runs in order of
// (i mod 128): the reference effectively keeps the k chosen candidates with the smallest (i & 127) — for top-200 lists
// sorted by score that is a bias toward the best-ranked ones (measured: 95 % of overshooting draws; the rest are probe
// collisions between i and i+128).  This kernel runs the same rounds with a counter-based generator and applies that rule;
// tests compare inclusion frequencies against the verbatim reference procedure.
// One warp per query: weights and their running sums live in shared memory.
#include "common.cuh"

namespace simx {

constexpr int SAMP_MAX_C = 1024;
constexpr int SAMP_MAX_K = 64;

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
__device__ __forceinline__ float uniform01(uint64_t seed, uint32_t query, uint32_t counter) {
  const uint64_t h = splitmix64(seed ^ splitmix64(((uint64_t)query << 32) | counter));
  return (float)(h >> 40) * (1.0f / 16777216.0f);  // 24 random bits -> [0, 1)
}

__global__ void __launch_bounds__(128)
simans_sample_kernel(const float* __restrict__ scores, const int32_t* __restrict__ n_valid, const float* __restrict__ pos,
                     int NQ, int NC, int k, int mode, float p0, float p1, uint64_t seed, int32_t* out) {
  __shared__ float w[4][SAMP_MAX_C];
  __shared__ float cdf[4][SAMP_MAX_C];
  __shared__ int draws[4][SAMP_MAX_K];
  __shared__ int chosen[4][2 * SAMP_MAX_K];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = blockIdx.x * 4 + warp;
  if (q >= NQ) return;
  const int n = min(n_valid ? n_valid[q] : NC, NC);
  const float* s = scores + (long)q * NC;
  int32_t* o = out + (long)q * k;
  const float sp = pos[q];
  if (sp == 0.f || n <= k) {
    // no retriever score for the positive (or nothing to choose from): the last k candidates, MARCO_until_new.py:179-180
    for (int j = lane; j < k; j += 32) o[j] = max(0, n - k) + min(j, max(n - 1, 0));
    return;
  }
  for (int i = lane; i < n; i += 32) {
    const float d = s[i] - sp;
    w[warp][i] = mode == 0 ? expf(-fabsf(d) * p0) : expf(-p0 * (d + p1) * (d + p1));
  }
  __syncwarp();
  int n_chosen = 0;
  uint32_t counter = 0;
  for (int round = 0; round < 64 && n_chosen < k; ++round) {
    // inclusive running sums of the current pool weights (chunks of 32 with a carried total)
    float carry = 0.f;
    for (int base = 0; base < n; base += 32) {
      float v = base + lane < n ? w[warp][base + lane] : 0.f;
#pragma unroll
      for (int o2 = 1; o2 < 32; o2 <<= 1) {
        const float t = __shfl_up_sync(0xffffffffu, v, o2);
        if (lane >= o2) v += t;
      }
      if (base + lane < n) cdf[warp][base + lane] = v + carry;
      carry += __shfl_sync(0xffffffffu, v, 31);
    }
    __syncwarp();
    const float total = carry;
    if (!(total > 0.f)) break;  // pool exhausted (all weights underflowed): fall through to the deterministic fill
    // k draws with replacement: random.choices(cands, weights, k) == bisect(cum_weights, random() * total)
    for (int j = lane; j < k; j += 32) {
      const float u = uniform01(seed, (uint32_t)q, counter + j) * total;
      int lo = 0, hi = n - 1;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (cdf[warp][mid] > u) hi = mid; else lo = mid + 1;
      }
      draws[warp][j] = lo;
    }
    counter += k;
    __syncwarp();
    if (lane == 0) {
      for (int j = 0; j < k; ++j) {  // the whole round joins the set, even past k (set.union of the k draws)
        const int c = draws[warp][j];
        if (w[warp][c] > 0.f) {  // not chosen yet (chosen candidates are dropped from the pool = weight 0)
          chosen[warp][n_chosen++] = c;
          w[warp][c] = 0.f;
        }
      }
    }
    n_chosen = __shfl_sync(0xffffffffu, n_chosen, 0);
    __syncwarp();
  }
  if (lane == 0) {
    // degenerate pools (fewer than k candidates with non-zero weight): fill with the highest-index unused candidates
    for (int c = n - 1; c >= 0 && n_chosen < k; --c) {
      bool used = false;
      for (int j = 0; j < n_chosen; ++j) used = used || chosen[warp][j] == c;
      if (!used) chosen[warp][n_chosen++] = c;
    }
  }
  if (lane == 0 && n_chosen > k) {
    // list(set)[:k]: keep the k smallest in CPython's iteration order for small ints, (i & 127) then i
    for (int a = 0; a < k; ++a) {
      int best = a;
      for (int b2 = a + 1; b2 < n_chosen; ++b2) {
        const int x = chosen[warp][b2], y = chosen[warp][best];
        if ((x & 127) < (y & 127) || ((x & 127) == (y & 127) && x < y)) best = b2;
      }
      const int t = chosen[warp][a];
      chosen[warp][a] = chosen[warp][best];
      chosen[warp][best] = t;
    }
  }
  __syncwarp();
  for (int j = lane; j < k; j += 32) o[j] = chosen[warp][j];
}

}  // namespace simx

extern "C" int simx_simans_sample(const float* scores, const int32_t* n_valid, const float* pos_score, int NQ, int NC,
                                  int k, int mode, float p0, float p1, uint64_t seed, int32_t* out_idx, void* stream) {
  using namespace simx;
  SIMX_CHECK(scores && pos_score && out_idx, SIMX_ERR_INVALID, "simans_sample: null argument");
  SIMX_CHECK(NQ > 0 && NC > 0 && NC <= SAMP_MAX_C, SIMX_ERR_INVALID, "simans_sample: need 0 < NC <= %d candidates", SAMP_MAX_C);
  SIMX_CHECK(k > 0 && k <= SAMP_MAX_K && k <= NC, SIMX_ERR_INVALID, "simans_sample: need 0 < k <= min(%d, NC)", SAMP_MAX_K);
  SIMX_CHECK(mode == 0 || mode == 1, SIMX_ERR_INVALID, "simans_sample: mode must be 0 (tau |d|) or 1 (a (d+b)^2)");
  simans_sample_kernel<<<(NQ + 3) / 4, 128, 0, static_cast<cudaStream_t>(stream)>>>(scores, n_valid, pos_score, NQ, NC, k,
                                                                                    mode, p0, p1, seed, out_idx);
  SIMX_LAUNCH_CHECK();
  return SIMX_OK;
}

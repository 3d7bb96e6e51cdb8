// Shared host/device helpers for libsimx_b200: status codes, the thread-local error string, the
// "split-bf16 plane" representation used by the fp32-class precision mode, launch checks.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/simx_b200.h"
#include "dropout.cuh"

namespace simx {

typedef __nv_bfloat16 bf16;

// internal GEMM epilogue (not reachable through simx_gemm: it needs a DropSpec):
//   out = dropout(acc + bias[n]) + aux[m,n]   (BertSelfOutput / BertOutput in training: LEAD/modeling_bert.py:384-388, 462-466)
constexpr int EPI_BIAS_DROP_ADD = 8;

// ---------------------------------------------------------------------------------------------
// error plumbing (no C++ exceptions cross the ABI; see include/simx_b200.h)
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
#define SIMX_FAIL(code, ...)     \
  do {                           \
    ::simx::set_error(__VA_ARGS__); \
    return (code);               \
  } while (0)
#define SIMX_CHECK(cond, code, ...) \
  do {                              \
    if (!(cond)) SIMX_FAIL(code, __VA_ARGS__); \
  } while (0)
#define SIMX_CUDA(expr)                                                                          \
  do {                                                                                           \
    cudaError_t e__ = (expr);                                                                    \
    if (e__ != cudaSuccess)                                                                      \
      SIMX_FAIL(SIMX_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)
#define SIMX_LAUNCH_CHECK()                                                                      \
  do {                                                                                           \
    cudaError_t e__ = cudaGetLastError();                                                        \
    if (e__ != cudaSuccess)                                                                      \
      SIMX_FAIL(SIMX_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)
#define SIMX_TRY(expr)          \
  do {                          \
    int s__ = (expr);           \
    if (s__ != SIMX_OK) return s__; \
  } while (0)

int num_sms();  // cached cudaDevAttrMultiProcessorCount of the current device

// the tcgen05 GEMM (gemm_sm100.cu); `drop` turns SIMX_EPI_BIAS_ADD into out = dropout(acc + bias) + aux
int gemm(int M, int N, int K, const __nv_bfloat16* A, long lda, long a_ps, bool a_mn, const __nv_bfloat16* B, long ldb,
         long b_ps, bool b_mn, int planes, int epi, const float* bias, const __nv_bfloat16* aux, long ldaux, long aux_ps,
         __nv_bfloat16* out, __nv_bfloat16* out2, long ldo, long out_ps, float* outf, long ldof, int ksplit,
         cudaStream_t stream, const DropSpec* drop = nullptr);

// ---------------------------------------------------------------------------------------------
// optional per-launch timing (simx_profile_begin / simx_profile_end): CUDA events recorded on the launching stream
// around every kernel wrapper, aggregated by label.  Off by default; costs nothing when off.
// ---------------------------------------------------------------------------------------------
void prof_label(const char* label);  // label for the next wrapper call(s) on this thread (nullptr = wrapper default)
struct ProfScope {
  ProfScope(const char* default_name, double flops, double bytes, cudaStream_t st);
  ~ProfScope();
  int slot;
  cudaStream_t st;
};

// ---------------------------------------------------------------------------------------------
// Activation tensors.  Every activation is NP bf16 "planes" of the same [rows, cols] shape, plane p at
// ptr + p * plane_stride.  NP == 1: plain bf16 (throughput mode).  NP == 3: value = hi + mid + lo with
// hi = bf16(v), mid = bf16(v - hi), lo = bf16(v - hi - mid)  -> 24 significant bits, i.e. fp32-class storage;
// GEMMs then run 6 plane-pair passes on the same tcgen05 kernel (products of order < 2^-24 dropped).
// ---------------------------------------------------------------------------------------------
struct Act {
  bf16* ptr;
  long ps;  // plane stride in elements
  int ld;   // row stride in elements
};

template <int NP>
__device__ __forceinline__ void split_store(bf16* p, long ps, float v) {
  bf16 h = __float2bfloat16_rn(v);
  p[0] = h;
  if constexpr (NP == 3) {
    float r = v - __bfloat162float(h);
    bf16 m = __float2bfloat16_rn(r);
    p[ps] = m;
    p[2 * ps] = __float2bfloat16_rn(r - __bfloat162float(m));
  }
}
template <int NP>
__device__ __forceinline__ float plane_load(const bf16* p, long ps) {
  float v = __bfloat162float(p[0]);
  if constexpr (NP == 3) v += __bfloat162float(p[ps]) + __bfloat162float(p[2 * ps]);
  return v;
}

// 8 consecutive elements (16 bytes per plane)
template <int NP>
__device__ __forceinline__ void plane_load8(const bf16* p, long ps, float (&v)[8]) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
  if constexpr (NP == 3) {
#pragma unroll
    for (int q = 1; q < 3; ++q) {
      uint4 w = *reinterpret_cast<const uint4*>(p + q * ps);
      const __nv_bfloat162* g = reinterpret_cast<const __nv_bfloat162*>(&w);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float2 f = __bfloat1622float2(g[i]);
        v[2 * i] += f.x;
        v[2 * i + 1] += f.y;
      }
    }
  }
}
template <int NP>
__device__ __forceinline__ void split_store8(bf16* p, long ps, const float (&v)[8]) {
  float r[8];
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    float2 f = __bfloat1622float2(h[i]);
    r[2 * i] = v[2 * i] - f.x;
    r[2 * i + 1] = v[2 * i + 1] - f.y;
  }
  *reinterpret_cast<uint4*>(p) = u;
  if constexpr (NP == 3) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      h[i] = __floats2bfloat162_rn(r[2 * i], r[2 * i + 1]);
      float2 f = __bfloat1622float2(h[i]);
      r[2 * i] -= f.x;
      r[2 * i + 1] -= f.y;
    }
    *reinterpret_cast<uint4*>(p + ps) = u;
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(r[2 * i], r[2 * i + 1]);
    *reinterpret_cast<uint4*>(p + 2 * ps) = u;
  }
}

// token ids / attention masks / position ids arrive as int64 (the reference's collate, MARCO_until_new.py:240-260) or
// int32 (simx_encoder_cfg.id_bytes = 4: half the host->device bytes)
__device__ __forceinline__ long ld_id(const void* p, long i, int id_bytes) {
  return id_bytes == 4 ? (long)static_cast<const int32_t*>(p)[i] : (long)static_cast<const int64_t*>(p)[i];
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace simx

// Optimizer tail of the train step on the GPU: global-norm gradient clipping + AdamW over ALL parameter
// tensors of a model in two launches (the reference runs ~3 small kernels per parameter tensor for the clip and ~8 for
// the update: torch.nn.utils.clip_grad_norm_ + AdamW.step, co_training_marco_train.py:57-69,246-263; LEAD/util.py:901).
//
// HBM-bound integer/float streaming: per element the update reads p, g, m, v and writes p, m, v = 28 B; the norm pass
// reads g = 4 B.  A block owns one 8192-element chunk of one tensor (16-byte vector accesses, a binary search over the
// chunk prefix finds the tensor), so the grid is ~27k blocks for the 219 M parameters of the two BERT-base towers.
#include "common.cuh"

namespace simx {

constexpr int OPT_CHUNK = 8192;  // elements per block
constexpr int OPT_THREADS = 256;

__device__ __forceinline__ int find_tensor(const simx_opt_tensor* t, int n, int chunk) {
  int lo = 0, hi = n - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (t[mid].chunk_begin <= chunk) lo = mid;
    else hi = mid - 1;
  }
  return lo;
}

__device__ __forceinline__ double block_sum(double v) {
  __shared__ double red[OPT_THREADS / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
  if (threadIdx.x < OPT_THREADS / 32) s = red[threadIdx.x];
  if (threadIdx.x < 32) {
#pragma unroll
    for (int o = OPT_THREADS / 64; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  }
  return s;  // valid in thread 0
}

// out_sq += sum over all tensors of g^2
__global__ void __launch_bounds__(OPT_THREADS) grad_sq_norm_kernel(const simx_opt_tensor* __restrict__ tensors, int n,
                                                                   double* __restrict__ out_sq) {
  const int ti = find_tensor(tensors, n, blockIdx.x);
  const simx_opt_tensor t = tensors[ti];
  float acc = 0.f;
  if (t.g != nullptr) {
    const int64_t begin = (int64_t)(blockIdx.x - t.chunk_begin) * OPT_CHUNK;
    const int64_t end = min(begin + OPT_CHUNK, t.numel);
    const float* g = t.g;
    if ((reinterpret_cast<uintptr_t>(g) & 15) == 0) {
      const int64_t v_end = begin + ((end - begin) & ~int64_t(3));
      for (int64_t i = begin + threadIdx.x * 4; i < v_end; i += OPT_THREADS * 4) {
        const float4 x = *reinterpret_cast<const float4*>(g + i);
        acc += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
      }
      for (int64_t i = v_end + threadIdx.x; i < end; i += OPT_THREADS) acc += g[i] * g[i];
    } else {
      for (int64_t i = begin + threadIdx.x; i < end; i += OPT_THREADS) acc += g[i] * g[i];
    }
  }
  const double s = block_sum((double)acc);
  if (threadIdx.x == 0 && s != 0.0) atomicAdd(out_sq, s);
}

// torch.nn.utils.clip_grad_norm_: coef = min(1, max_norm / (total_norm + 1e-6))
__device__ __forceinline__ float clip_coef(const double* sq_norm, float max_norm) {
  if (sq_norm == nullptr) return 1.f;
  const float total = (float)sqrt(*sq_norm);
  return fminf(1.f, max_norm / (total + 1e-6f));
}

__global__ void __launch_bounds__(OPT_THREADS) scale_grads_kernel(const simx_opt_tensor* __restrict__ tensors, int n,
                                                                  const double* __restrict__ sq_norm, float max_norm) {
  const int ti = find_tensor(tensors, n, blockIdx.x);
  const simx_opt_tensor t = tensors[ti];
  if (t.g == nullptr) return;
  const float c = clip_coef(sq_norm, max_norm);
  const int64_t begin = (int64_t)(blockIdx.x - t.chunk_begin) * OPT_CHUNK;
  const int64_t end = min(begin + OPT_CHUNK, t.numel);
  for (int64_t i = begin + threadIdx.x; i < end; i += OPT_THREADS) t.g[i] *= c;
}

struct AdamArgs {
  int mode;           // SIMX_ADAMW_TORCH / SIMX_ADAMW_HF
  float beta1, beta2, eps;
  float omb1, omb2;     // 1 - beta, rounded from the double-precision difference (1.f - 0.999f is off by 5e-5 relative)
  float bc1, bc2_sqrt;  // 1 - beta1^t, sqrt(1 - beta2^t)
  int correct_bias;   // HF mode only
  float max_norm;
  int write_grads;    // store the clipped gradient back (what clip_grad_norm_ leaves behind)
};

__device__ __forceinline__ void adam_one(const AdamArgs& a, float lr, float wd, float clip, float& p, float& g, float& m,
                                         float& v) {
  g *= clip;
  if (a.mode == SIMX_ADAMW_TORCH) {
    // torch.optim.AdamW (LEAD/util.py:901): decoupled decay first, bias-corrected second moment inside the denominator
    p *= 1.f - lr * wd;
    m = m + (g - m) * a.omb1;
    v = v * a.beta2 + a.omb2 * g * g;
    const float denom = sqrtf(v) / a.bc2_sqrt + a.eps;
    p -= (lr / a.bc1) * (m / denom);
  } else {
    // transformers.AdamW, the optimizer of the SimANS / PROD drivers (co_training_marco_train.py:22,67): eps is added
    // to the UN-corrected sqrt(v), the bias corrections scale the step size, decay is applied after the update
    m = m * a.beta1 + a.omb1 * g;
    v = v * a.beta2 + a.omb2 * g * g;
    const float denom = sqrtf(v) + a.eps;
    const float step = a.correct_bias ? lr * a.bc2_sqrt / a.bc1 : lr;
    p -= step * (m / denom);
    if (wd > 0.f) p -= lr * wd * p;
  }
}

__global__ void __launch_bounds__(OPT_THREADS) adamw_kernel(const simx_opt_tensor* __restrict__ tensors, int n,
                                                            const AdamArgs a, const double* __restrict__ sq_norm) {
  const int ti = find_tensor(tensors, n, blockIdx.x);
  const simx_opt_tensor t = tensors[ti];
  if (t.g == nullptr) return;  // a parameter without a gradient is skipped, as both reference optimizers do
  const float clip = clip_coef(sq_norm, a.max_norm);
  const int64_t begin = (int64_t)(blockIdx.x - t.chunk_begin) * OPT_CHUNK;
  const int64_t end = min(begin + OPT_CHUNK, t.numel);
  const bool aligned = ((reinterpret_cast<uintptr_t>(t.p) | reinterpret_cast<uintptr_t>(t.g) |
                         reinterpret_cast<uintptr_t>(t.m) | reinterpret_cast<uintptr_t>(t.v)) & 15) == 0;
  int64_t scalar_from = begin;
  if (aligned) {
    const int64_t v_end = begin + ((end - begin) & ~int64_t(3));
    for (int64_t i = begin + threadIdx.x * 4; i < v_end; i += OPT_THREADS * 4) {
      float4 p = *reinterpret_cast<const float4*>(t.p + i);
      float4 g = *reinterpret_cast<const float4*>(t.g + i);
      float4 m = *reinterpret_cast<const float4*>(t.m + i);
      float4 v = *reinterpret_cast<const float4*>(t.v + i);
      adam_one(a, t.lr, t.weight_decay, clip, p.x, g.x, m.x, v.x);
      adam_one(a, t.lr, t.weight_decay, clip, p.y, g.y, m.y, v.y);
      adam_one(a, t.lr, t.weight_decay, clip, p.z, g.z, m.z, v.z);
      adam_one(a, t.lr, t.weight_decay, clip, p.w, g.w, m.w, v.w);
      *reinterpret_cast<float4*>(t.p + i) = p;
      *reinterpret_cast<float4*>(t.m + i) = m;
      *reinterpret_cast<float4*>(t.v + i) = v;
      if (a.write_grads) *reinterpret_cast<float4*>(t.g + i) = g;
    }
    scalar_from = v_end;
  }
  for (int64_t i = scalar_from + threadIdx.x; i < end; i += OPT_THREADS) {
    float p = t.p[i], g = t.g[i], m = t.m[i], v = t.v[i];
    adam_one(a, t.lr, t.weight_decay, clip, p, g, m, v);
    t.p[i] = p;
    t.m[i] = m;
    t.v[i] = v;
    if (a.write_grads) t.g[i] = g;
  }
}

}  // namespace simx

using namespace simx;

extern "C" int simx_opt_plan(simx_opt_tensor* tensors_host, int n) {
  if (tensors_host == nullptr || n <= 0) {
    set_error("opt_plan: empty tensor table");
    return -1;
  }
  int64_t chunk = 0;
  for (int i = 0; i < n; ++i) {
    if (tensors_host[i].numel <= 0 || tensors_host[i].p == nullptr) {
      set_error("opt_plan: tensor %d has no elements or a null parameter pointer", i);
      return -1;
    }
    tensors_host[i].chunk_begin = (int32_t)chunk;
    chunk += (tensors_host[i].numel + OPT_CHUNK - 1) / OPT_CHUNK;
    if (chunk > 0x7fffffff) {
      set_error("opt_plan: too many chunks");
      return -1;
    }
  }
  return (int)chunk;
}

extern "C" int simx_grad_sq_norm(const simx_opt_tensor* tensors_dev, int n, int total_chunks, double* out_sq_dev,
                                 void* stream) {
  SIMX_CHECK(tensors_dev && out_sq_dev && n > 0 && total_chunks > 0, SIMX_ERR_INVALID, "grad_sq_norm: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ProfScope prof("grad_sq_norm", 0, 4.0 * (double)total_chunks * OPT_CHUNK, st);
  SIMX_CUDA(cudaMemsetAsync(out_sq_dev, 0, sizeof(double), st));
  grad_sq_norm_kernel<<<total_chunks, OPT_THREADS, 0, st>>>(tensors_dev, n, out_sq_dev);
  SIMX_LAUNCH_CHECK();
  return SIMX_OK;
}

extern "C" int simx_clip_grads(const simx_opt_tensor* tensors_dev, int n, int total_chunks, const double* sq_norm_dev,
                               float max_norm, void* stream) {
  SIMX_CHECK(tensors_dev && sq_norm_dev && n > 0 && total_chunks > 0, SIMX_ERR_INVALID, "clip_grads: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ProfScope prof("clip_grads", 0, 8.0 * (double)total_chunks * OPT_CHUNK, st);
  scale_grads_kernel<<<total_chunks, OPT_THREADS, 0, st>>>(tensors_dev, n, sq_norm_dev, max_norm);
  SIMX_LAUNCH_CHECK();
  return SIMX_OK;
}

extern "C" int simx_adamw_step(const simx_opt_tensor* tensors_dev, int n, int total_chunks, int mode, double beta1,
                               double beta2, float eps, int step, int correct_bias, const double* sq_norm_dev,
                               float max_norm, int write_clipped_grads, void* stream) {
  SIMX_CHECK(tensors_dev && n > 0 && total_chunks > 0, SIMX_ERR_INVALID, "adamw_step: bad argument");
  SIMX_CHECK(mode == SIMX_ADAMW_TORCH || mode == SIMX_ADAMW_HF, SIMX_ERR_INVALID, "adamw_step: unknown mode %d", mode);
  SIMX_CHECK(step >= 1, SIMX_ERR_INVALID, "adamw_step: step counts from 1 (got %d)", step);
  SIMX_CHECK(beta1 >= 0.0 && beta1 < 1.0 && beta2 >= 0.0 && beta2 < 1.0, SIMX_ERR_INVALID, "adamw_step: betas out of range");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  AdamArgs a;
  a.mode = mode;
  a.beta1 = (float)beta1;
  a.beta2 = (float)beta2;
  a.omb1 = (float)(1.0 - beta1);
  a.omb2 = (float)(1.0 - beta2);
  a.eps = eps;
  a.bc1 = (float)(1.0 - pow(beta1, (double)step));
  a.bc2_sqrt = (float)sqrt(1.0 - pow(beta2, (double)step));
  a.correct_bias = correct_bias;
  a.max_norm = max_norm;
  a.write_grads = write_clipped_grads && sq_norm_dev != nullptr;
  ProfScope prof("adamw_step", 0, 28.0 * (double)total_chunks * OPT_CHUNK, st);
  adamw_kernel<<<total_chunks, OPT_THREADS, 0, st>>>(tensors_dev, n, a, sq_norm_dev);
  SIMX_LAUNCH_CHECK();
  return SIMX_OK;
}

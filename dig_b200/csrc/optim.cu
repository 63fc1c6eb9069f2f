// dig_b200 -- multi-tensor parameter kernels: one launch walks every parameter tensor through a device-side
// pointer table (no flattening of the nn.Parameters is required, so state_dict()/DDP/optimizer see ordinary tensors).
//   * fp32 -> bf16 shadow cast of the online weights (operands of the tcgen05 GEMMs)
//   * EMA momentum update  p_m = m p_m + (1-m) p  (M:428-442) fused with the bf16 shadow refresh of the momentum weights
//   * sum of squares for the global gradient norm (U:507-519)
//   * AdamW step (custom_optim/_functional.py:115-140) fused with gradient unscale / clipping
// Table layout (all device arrays): ptr tables hold raw addresses as int64; blk_tensor[b], blk_chunk[b] give the tensor and
// the 16384-element chunk that thread block b processes.
#include <math.h>

#include "common.cuh"
#include "../../include/dig_b200.h"

namespace dig {

static constexpr int kChunk = 16384;

__global__ void __launch_bounds__(256)
mt_cast_bf16_kernel(const long long* __restrict__ src, const long long* __restrict__ dst, const long long* __restrict__ numel,
                    const int* __restrict__ blk_tensor, const int* __restrict__ blk_chunk) {
  const int t = blk_tensor[blockIdx.x];
  const long long n = numel[t], o0 = (long long)blk_chunk[blockIdx.x] * kChunk;
  const float* s = reinterpret_cast<const float*>(src[t]);
  __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(dst[t]);
  const long long end = min(n, o0 + kChunk);
  for (long long i = o0 + threadIdx.x; i < end; i += 256) d[i] = __float2bfloat16(s[i]);
}

__global__ void __launch_bounds__(256)
mt_copy_f32_kernel(const long long* __restrict__ src, const long long* __restrict__ dst, const long long* __restrict__ numel,
                   const int* __restrict__ blk_tensor, const int* __restrict__ blk_chunk) {
  const int t = blk_tensor[blockIdx.x];
  const long long n = numel[t], o0 = (long long)blk_chunk[blockIdx.x] * kChunk;
  const float* s = reinterpret_cast<const float*>(src[t]);
  float* d = reinterpret_cast<float*>(dst[t]);
  const long long end = min(n, o0 + kChunk);
  for (long long i = o0 + threadIdx.x; i < end; i += 256) d[i] = s[i];
}

__global__ void __launch_bounds__(256)
mt_ema_kernel(const long long* __restrict__ online, const long long* __restrict__ target, const long long* __restrict__ shadow,
              const long long* __restrict__ numel, const int* __restrict__ blk_tensor, const int* __restrict__ blk_chunk, float m) {
  const int t = blk_tensor[blockIdx.x];
  const long long n = numel[t], o0 = (long long)blk_chunk[blockIdx.x] * kChunk;
  const float* p = reinterpret_cast<const float*>(online[t]);
  float* pm = reinterpret_cast<float*>(target[t]);
  __nv_bfloat16* sh = (shadow && shadow[t]) ? reinterpret_cast<__nv_bfloat16*>(shadow[t]) : nullptr;
  const long long end = min(n, o0 + kChunk);
  const float om = 1.0f - m;
  for (long long i = o0 + threadIdx.x; i < end; i += 256) {
    const float v = pm[i] * m + p[i] * om;
    pm[i] = v;
    if (sh) sh[i] = __float2bfloat16(v);
  }
}

__global__ void __launch_bounds__(256)
mt_sumsq_kernel(const long long* __restrict__ src, const long long* __restrict__ numel, const int* __restrict__ blk_tensor,
                const int* __restrict__ blk_chunk, float* __restrict__ out) {
  const int t = blk_tensor[blockIdx.x];
  const long long n = numel[t], o0 = (long long)blk_chunk[blockIdx.x] * kChunk;
  const float* s = reinterpret_cast<const float*>(src[t]);
  const long long end = min(n, o0 + kChunk);
  float acc = 0.f;
  for (long long i = o0 + threadIdx.x; i < end; i += 256) { const float v = s[i]; acc += v * v; }
  acc = warp_sum(acc);
  __shared__ float red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int w = 0; w < 8; ++w) tot += red[w];
    atomicAdd(out, tot);
  }
}

// scalars[0] = sum of squared gradients (for clipping); grad_scale multiplies every gradient first (1/loss_scale).
__global__ void __launch_bounds__(256)
mt_adamw_kernel(const long long* __restrict__ params, const long long* __restrict__ grads, const long long* __restrict__ exp_avg,
                const long long* __restrict__ exp_avg_sq, const long long* __restrict__ shadow, const long long* __restrict__ numel,
                const float* __restrict__ lr_t, const float* __restrict__ wd_t, const int* __restrict__ blk_tensor,
                const int* __restrict__ blk_chunk, float beta1, float beta2, float eps, float bc1, float bc2_sqrt, float grad_scale,
                const float* __restrict__ sumsq, float max_norm) {
  const int t = blk_tensor[blockIdx.x];
  const long long n = numel[t], o0 = (long long)blk_chunk[blockIdx.x] * kChunk;
  float* p = reinterpret_cast<float*>(params[t]);
  const float* g = reinterpret_cast<const float*>(grads[t]);
  float* m1 = reinterpret_cast<float*>(exp_avg[t]);
  float* m2 = reinterpret_cast<float*>(exp_avg_sq[t]);
  __nv_bfloat16* sh = (shadow && shadow[t]) ? reinterpret_cast<__nv_bfloat16*>(shadow[t]) : nullptr;
  const float lr = lr_t[t], wd = wd_t[t];
  float gs = grad_scale;
  if (max_norm > 0.f && sumsq) {  // torch.nn.utils.clip_grad_norm_: coef = max_norm / (norm + 1e-6), applied when < 1
    const float norm = sqrtf(sumsq[0]) * grad_scale;
    const float coef = max_norm / (norm + 1e-6f);
    if (coef < 1.f) gs *= coef;
  }
  const long long end = min(n, o0 + kChunk);
  const float step = lr / bc1;
  for (long long i = o0 + threadIdx.x; i < end; i += 256) {
    const float gr = g[i] * gs;
    float w = p[i] * (1.f - lr * wd);
    const float a = m1[i] * beta1 + gr * (1.f - beta1);
    const float b = m2[i] * beta2 + gr * gr * (1.f - beta2);
    m1[i] = a;
    m2[i] = b;
    w -= step * a / (sqrtf(b) / bc2_sqrt + eps);
    p[i] = w;
    if (sh) sh[i] = __float2bfloat16(w);
  }
}

}  // namespace dig

using namespace dig;

extern "C" int dig_mt_chunk(void) { return kChunk; }

extern "C" int dig_mt_cast_bf16(const int64_t* src, const int64_t* dst, const int64_t* numel, const int32_t* blk_tensor,
                                const int32_t* blk_chunk, int32_t num_blocks, void* stream) {
  DIG_REQUIRE(src && dst && numel && blk_tensor && blk_chunk, "dig_mt_cast_bf16: null table");
  if (num_blocks <= 0) return 0;
  mt_cast_bf16_kernel<<<num_blocks, 256, 0, (cudaStream_t)stream>>>((const long long*)src, (const long long*)dst, (const long long*)numel,
                                                                  blk_tensor, blk_chunk);
  DIG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int dig_mt_copy_f32(const int64_t* src, const int64_t* dst, const int64_t* numel, const int32_t* blk_tensor,
                               const int32_t* blk_chunk, int32_t num_blocks, void* stream) {
  DIG_REQUIRE(src && dst && numel && blk_tensor && blk_chunk, "dig_mt_copy_f32: null table");
  if (num_blocks <= 0) return 0;
  mt_copy_f32_kernel<<<num_blocks, 256, 0, (cudaStream_t)stream>>>((const long long*)src, (const long long*)dst, (const long long*)numel,
                                                                 blk_tensor, blk_chunk);
  DIG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int dig_mt_ema(const int64_t* online, const int64_t* target, const int64_t* shadow, const int64_t* numel,
                          const int32_t* blk_tensor, const int32_t* blk_chunk, int32_t num_blocks, float m, void* stream) {
  DIG_REQUIRE(online && target && numel && blk_tensor && blk_chunk, "dig_mt_ema: null table");
  if (num_blocks <= 0) return 0;
  mt_ema_kernel<<<num_blocks, 256, 0, (cudaStream_t)stream>>>((const long long*)online, (const long long*)target, (const long long*)shadow,
                                                            (const long long*)numel, blk_tensor, blk_chunk, m);
  DIG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int dig_mt_sumsq(const int64_t* src, const int64_t* numel, const int32_t* blk_tensor, const int32_t* blk_chunk,
                            int32_t num_blocks, float* out, void* stream) {
  DIG_REQUIRE(src && numel && blk_tensor && blk_chunk && out, "dig_mt_sumsq: null table");
  if (num_blocks <= 0) return 0;
  mt_sumsq_kernel<<<num_blocks, 256, 0, (cudaStream_t)stream>>>((const long long*)src, (const long long*)numel, blk_tensor, blk_chunk, out);
  DIG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int dig_mt_adamw(const int64_t* params, const int64_t* grads, const int64_t* exp_avg, const int64_t* exp_avg_sq,
                            const int64_t* shadow, const int64_t* numel, const float* lr, const float* weight_decay,
                            const int32_t* blk_tensor, const int32_t* blk_chunk, int32_t num_blocks, float beta1, float beta2, float eps,
                            int64_t step, float grad_scale, const float* sumsq, float max_norm, void* stream) {
  DIG_REQUIRE(params && grads && exp_avg && exp_avg_sq && numel && lr && weight_decay && blk_tensor && blk_chunk,
              "dig_mt_adamw: null table");
  DIG_REQUIRE(step >= 1, "dig_mt_adamw: step must be >= 1");
  if (num_blocks <= 0) return 0;
  const float bc1 = (float)(1.0 - pow((double)beta1, (double)step));
  const float bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
  mt_adamw_kernel<<<num_blocks, 256, 0, (cudaStream_t)stream>>>(
      (const long long*)params, (const long long*)grads, (const long long*)exp_avg, (const long long*)exp_avg_sq, (const long long*)shadow,
      (const long long*)numel, lr, weight_decay, blk_tensor, blk_chunk, beta1, beta2, eps, bc1, bc2_sqrt, grad_scale, sumsq, max_norm);
  DIG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

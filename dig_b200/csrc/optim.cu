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

// Every kernel below walks its 16384-element chunk with 128-bit accesses: a thread owns float4 #(threadIdx + 256 j) of the chunk, four of
// them per loop trip (independent loads in flight before the first use), so a warp touches 512 contiguous bytes per stream and request.
// Tensors whose base addresses are not 16-byte aligned (never the case for the step's own tables: torch allocations are 512-byte
// aligned, gradient views sit at multiples of 4 floats of the flat buffer, shadows at multiples of 8 bf16) and the sub-float4 tail of
// a tensor take the scalar path.
__device__ __forceinline__ bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
__device__ __forceinline__ uint2 pack_bf16x4(float4 v) { return make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w)); }
__device__ __forceinline__ float4 ldg_stream_f4(const float* p) {   // read-once data: do not keep it in L1
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

__global__ void __launch_bounds__(256)
mt_cast_bf16_kernel(const long long* __restrict__ src, const long long* __restrict__ dst, const long long* __restrict__ numel,
                    const int* __restrict__ blk_tensor, const int* __restrict__ blk_chunk) {
  const int t = blk_tensor[blockIdx.x];
  const long long n = numel[t], o0 = (long long)blk_chunk[blockIdx.x] * kChunk;
  const float* s = reinterpret_cast<const float*>(src[t]) + o0;
  __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(dst[t]) + o0;
  const int cnt = (int)min((long long)kChunk, n - o0);
  int done = 0;
  if (aligned16(s) && (reinterpret_cast<uintptr_t>(d) & 7) == 0) {
    const int nv = cnt >> 2;
    for (int i0 = threadIdx.x; i0 < nv; i0 += 1024) {
      float4 v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) if (i0 + 256 * j < nv) v[j] = ldg_stream_f4(s + 4 * (i0 + 256 * j));
#pragma unroll
      for (int j = 0; j < 4; ++j) if (i0 + 256 * j < nv) *reinterpret_cast<uint2*>(d + 4 * (i0 + 256 * j)) = pack_bf16x4(v[j]);
    }
    done = nv << 2;
  }
  for (int i = done + threadIdx.x; i < cnt; i += 256) d[i] = __float2bfloat16(s[i]);
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
  const float* p = reinterpret_cast<const float*>(online[t]) + o0;
  float* pm = reinterpret_cast<float*>(target[t]) + o0;
  __nv_bfloat16* sh = (shadow && shadow[t]) ? reinterpret_cast<__nv_bfloat16*>(shadow[t]) + o0 : nullptr;
  const int cnt = (int)min((long long)kChunk, n - o0);
  const float om = 1.0f - m;
  int done = 0;
  if (aligned16(p) && aligned16(pm) && (reinterpret_cast<uintptr_t>(sh) & 7) == 0) {
    const int nv = cnt >> 2;
    for (int i0 = threadIdx.x; i0 < nv; i0 += 1024) {
      float4 a[4], b[4];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (i0 + 256 * j < nv) {
          a[j] = ldg_stream_f4(p + 4 * (i0 + 256 * j));
          b[j] = *reinterpret_cast<const float4*>(pm + 4 * (i0 + 256 * j));
        }
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (i0 + 256 * j < nv) {
          float4 v;
          v.x = b[j].x * m + a[j].x * om; v.y = b[j].y * m + a[j].y * om;
          v.z = b[j].z * m + a[j].z * om; v.w = b[j].w * m + a[j].w * om;
          *reinterpret_cast<float4*>(pm + 4 * (i0 + 256 * j)) = v;
          if (sh) *reinterpret_cast<uint2*>(sh + 4 * (i0 + 256 * j)) = pack_bf16x4(v);
        }
    }
    done = nv << 2;
  }
  for (int i = done + threadIdx.x; i < cnt; i += 256) {
    const float v = pm[i] * m + p[i] * om;
    pm[i] = v;
    if (sh) sh[i] = __float2bfloat16(v);
  }
}

__device__ __forceinline__ void block_atomic_sum(float acc, float* out) {
  acc = warp_sum(acc);
  __shared__ float red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) tot += red[w];
    atomicAdd(out, tot);
  }
}

__global__ void __launch_bounds__(256)
mt_sumsq_kernel(const long long* __restrict__ src, const long long* __restrict__ numel, const int* __restrict__ blk_tensor,
                const int* __restrict__ blk_chunk, float* __restrict__ out) {
  const int t = blk_tensor[blockIdx.x];
  const long long n = numel[t], o0 = (long long)blk_chunk[blockIdx.x] * kChunk;
  const float* s = reinterpret_cast<const float*>(src[t]) + o0;
  const int cnt = (int)min((long long)kChunk, n - o0);
  float acc = 0.f;
  int done = 0;
  if (aligned16(s)) {
    const int nv = cnt >> 2;
    for (int i0 = threadIdx.x; i0 < nv; i0 += 1024) {
      float4 v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = (i0 + 256 * j < nv) ? *reinterpret_cast<const float4*>(s + 4 * (i0 + 256 * j)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int j = 0; j < 4; ++j) acc += v[j].x * v[j].x + v[j].y * v[j].y + v[j].z * v[j].z + v[j].w * v[j].w;
    }
    done = nv << 2;
  }
  for (int i = done + threadIdx.x; i < cnt; i += 256) { const float v = s[i]; acc += v * v; }
  block_atomic_sum(acc, out);
}

// sumsq[0] = sum of squared (unscaled) gradients, needed up front only for clipping; grad_scale multiplies every gradient first
// (1/loss_scale).  sumsq_out (optional) += sum of squared scaled gradients as this launch reads them: when nothing is clipped the
// gradient norm the engine logs (U:507-519) costs no extra pass over the gradients.
struct AdamwScalars { float beta1, beta2, eps, bc1, bc2_sqrt, gs, lr, wd; };
__device__ __forceinline__ float adamw_one(float& w, float g, float& m1, float& m2, const AdamwScalars& c, float step) {
  const float gr = g * c.gs;
  w *= (1.f - c.lr * c.wd);
  m1 = m1 * c.beta1 + gr * (1.f - c.beta1);
  m2 = m2 * c.beta2 + gr * gr * (1.f - c.beta2);
  w -= step * m1 / (sqrtf(m2) / c.bc2_sqrt + c.eps);
  return gr * gr;
}

__global__ void __launch_bounds__(256)
mt_adamw_kernel(const long long* __restrict__ params, const long long* __restrict__ grads, const long long* __restrict__ exp_avg,
                const long long* __restrict__ exp_avg_sq, const long long* __restrict__ shadow, const long long* __restrict__ numel,
                const float* __restrict__ lr_t, const float* __restrict__ wd_t, const int* __restrict__ blk_tensor,
                const int* __restrict__ blk_chunk, float beta1, float beta2, float eps, float bc1, float bc2_sqrt, float grad_scale,
                const float* __restrict__ sumsq, float max_norm, float* __restrict__ sumsq_out, const float* __restrict__ guard) {
  // guard: the step's loss on the device -- a non-finite loss leaves parameters, moments and shadows untouched (E:148-150 aborts
  // before backward/step; here the host sees the loss one step late, so the skip happens on the device)
  if (guard != nullptr && !isfinite(guard[0])) return;
  const int t = blk_tensor[blockIdx.x];
  const long long n = numel[t], o0 = (long long)blk_chunk[blockIdx.x] * kChunk;
  float* p = reinterpret_cast<float*>(params[t]) + o0;
  const float* g = reinterpret_cast<const float*>(grads[t]) + o0;
  float* m1 = reinterpret_cast<float*>(exp_avg[t]) + o0;
  float* m2 = reinterpret_cast<float*>(exp_avg_sq[t]) + o0;
  __nv_bfloat16* sh = (shadow && shadow[t]) ? reinterpret_cast<__nv_bfloat16*>(shadow[t]) + o0 : nullptr;
  AdamwScalars c;
  c.beta1 = beta1; c.beta2 = beta2; c.eps = eps; c.bc1 = bc1; c.bc2_sqrt = bc2_sqrt; c.lr = lr_t[t]; c.wd = wd_t[t];
  c.gs = grad_scale;
  if (max_norm >= 0.f && sumsq) {  // torch.nn.utils.clip_grad_norm_: coef = max_norm / (norm + 1e-6), applied when < 1 (max_norm < 0: no clipping)
    const float norm = sqrtf(sumsq[0]) * grad_scale;
    const float coef = max_norm / (norm + 1e-6f);
    if (coef < 1.f) c.gs *= coef;
  }
  const int cnt = (int)min((long long)kChunk, n - o0);
  const float step = c.lr / bc1;
  float acc = 0.f;
  int done = 0;
  if (aligned16(p) && aligned16(g) && aligned16(m1) && aligned16(m2) && (reinterpret_cast<uintptr_t>(sh) & 7) == 0) {
    const int nv = cnt >> 2;
    for (int i0 = threadIdx.x; i0 < nv; i0 += 512) {
      float4 w[2], gr[2], a[2], b[2];
#pragma unroll
      for (int j = 0; j < 2; ++j)
        if (i0 + 256 * j < nv) {
          const int o = 4 * (i0 + 256 * j);
          gr[j] = ldg_stream_f4(g + o);
          w[j] = *reinterpret_cast<const float4*>(p + o);
          a[j] = *reinterpret_cast<const float4*>(m1 + o);
          b[j] = *reinterpret_cast<const float4*>(m2 + o);
        }
#pragma unroll
      for (int j = 0; j < 2; ++j)
        if (i0 + 256 * j < nv) {
          const int o = 4 * (i0 + 256 * j);
          acc += adamw_one(w[j].x, gr[j].x, a[j].x, b[j].x, c, step);
          acc += adamw_one(w[j].y, gr[j].y, a[j].y, b[j].y, c, step);
          acc += adamw_one(w[j].z, gr[j].z, a[j].z, b[j].z, c, step);
          acc += adamw_one(w[j].w, gr[j].w, a[j].w, b[j].w, c, step);
          *reinterpret_cast<float4*>(m1 + o) = a[j];
          *reinterpret_cast<float4*>(m2 + o) = b[j];
          *reinterpret_cast<float4*>(p + o) = w[j];
          if (sh) *reinterpret_cast<uint2*>(sh + o) = pack_bf16x4(w[j]);
        }
    }
    done = nv << 2;
  }
  for (int i = done + threadIdx.x; i < cnt; i += 256) {
    float w = p[i], a = m1[i], b = m2[i];
    acc += adamw_one(w, g[i], a, b, c, step);
    m1[i] = a; m2[i] = b; p[i] = w;
    if (sh) sh[i] = __float2bfloat16(w);
  }
  if (sumsq_out != nullptr) block_atomic_sum(acc, sumsq_out);
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
                            int64_t step, float grad_scale, const float* sumsq, float max_norm, float* sumsq_out, const float* guard, void* stream) {
  DIG_REQUIRE(params && grads && exp_avg && exp_avg_sq && numel && lr && weight_decay && blk_tensor && blk_chunk,
              "dig_mt_adamw: null table");
  DIG_REQUIRE(step >= 1, "dig_mt_adamw: step must be >= 1");
  if (num_blocks <= 0) return 0;
  const float bc1 = (float)(1.0 - pow((double)beta1, (double)step));
  const float bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
  mt_adamw_kernel<<<num_blocks, 256, 0, (cudaStream_t)stream>>>(
      (const long long*)params, (const long long*)grads, (const long long*)exp_avg, (const long long*)exp_avg_sq, (const long long*)shadow,
      (const long long*)numel, lr, weight_decay, blk_tensor, blk_chunk, beta1, beta2, eps, bc1, bc2_sqrt, grad_scale, sumsq, max_norm, sumsq_out, guard);
  DIG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

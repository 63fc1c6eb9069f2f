// dig_b200 -- HBM-bound row / column kernels of the pre-training step (coalesced, 64/128-bit vectorised).
//   im2col for the 4x4 patch embed (F:188-195), LayerNorm forward/backward (F:134,140; M:424), window pooling
//   (M:189-193), row gather / scatter-add for the masked-pixel decoder (M:563-570), column sums for bias
//   gradients, BatchNorm1d batch statistics / apply / backward (M:463-482).
#include <string.h>

#include "common.cuh"
#include "peer.cuh"
#include "../../include/dig_b200.h"

namespace dig {

// ------------------------------------------------------------------------------------------------
// im2col: images fp32 [S,3,32,128] -> bf16 [S*256, 48]; column k = c*16 + kh*4 + kw (conv weight order)
// ------------------------------------------------------------------------------------------------
__global__ void im2col_patch4_kernel(const float* __restrict__ img, __nv_bfloat16* __restrict__ out, long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // one thread per (row, c, kh): 4 kw values
  if (idx >= total) return;
  const int grp = (int)(idx % 12);
  const long long row = idx / 12;
  const int c = grp >> 2, kh = grp & 3;
  const int tok = (int)(row & 255);
  const long long s = row >> 8;
  const int ph = tok >> 5, pw = tok & 31;
  const float4 v = __ldg(reinterpret_cast<const float4*>(img + ((s * 3 + c) * 32 + ph * 4 + kh) * 128 + pw * 4));
  uint2 packed;
  packed.x = pack_bf16(v.x, v.y);
  packed.y = pack_bf16(v.z, v.w);
  *reinterpret_cast<uint2*>(out + row * 48 + grp * 4) = packed;
}

// ------------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, d <= 512, d % 64 == 0
// ------------------------------------------------------------------------------------------------
static constexpr int kLnMaxPairs = 8;  // float2 per lane

template <bool GELU>
__global__ void __launch_bounds__(256)
layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                     __nv_bfloat16* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out, long long rows, int d,
                     float eps) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int np = d >> 6;
  const float* xr = x + row * d;
  float2 v[kLnMaxPairs];
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < kLnMaxPairs; ++k)
    if (k < np) {
      v[k] = *reinterpret_cast<const float2*>(xr + k * 64 + lane * 2);
      s += v[k].x + v[k].y;
    }
  const float mean = warp_sum(s) / d;
  float q = 0.f;
#pragma unroll
  for (int k = 0; k < kLnMaxPairs; ++k)
    if (k < np) {
      const float a = v[k].x - mean, b = v[k].y - mean;
      q += a * a + b * b;
    }
  const float rstd = rsqrtf(warp_sum(q) / d + eps);
#pragma unroll
  for (int k = 0; k < kLnMaxPairs; ++k)
    if (k < np) {
      const int c = k * 64 + lane * 2;
      const float2 g = *reinterpret_cast<const float2*>(gamma + c);
      const float2 b = *reinterpret_cast<const float2*>(beta + c);
      float o0 = (v[k].x - mean) * rstd * g.x + b.x;
      float o1 = (v[k].y - mean) * rstd * g.y + b.y;
      if (GELU) { o0 = gelu_erf(o0); o1 = gelu_erf(o1); }
      *reinterpret_cast<uint32_t*>(y + row * d + c) = pack_bf16(o0, o1);
    }
  if (lane == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
}

// dx = dres + LN_bwd(dy);  dgamma += sum_rows dy*xhat;  dbeta += sum_rows dy;  dxsum += sum_rows dx
// (dy is w.r.t. the LN output, or the GELU(LN) output when GELU).  One warp per row, D/32 contiguous-by-4 columns per lane,
// every global access 128-bit (64-bit for the bf16 streams); all loads of a row are issued before the first reduction.
// RBF: the residual gradient `dres` is a bf16 stream (dig_layernorm_bwd_bf16res) instead of fp32.
template <int D, bool GELU, bool RBF>
__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ mean,
                     const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ beta,
                     const void* dres_v, float* dx_f32, __nv_bfloat16* __restrict__ dx_bf16,
                     float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dxsum, long long rows) {
  constexpr int NV = D / 128;  // float4 groups per lane
  __shared__ float part[3 * D];
  pdl_launch_dependents();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int i = threadIdx.x; i < 3 * D; i += blockDim.x) part[i] = 0.f;
  __syncthreads();
  pdl_wait();
  float4 ag[NV], ab[NV], ax[NV], gm[NV], bt[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    ag[k] = ab[k] = ax[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    gm[k] = *reinterpret_cast<const float4*>(gamma + k * 128 + lane * 4);
    bt[k] = GELU ? *reinterpret_cast<const float4*>(beta + k * 128 + lane * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (long long row = (long long)blockIdx.x * nwarps + warp; row < rows; row += (long long)gridDim.x * nwarps) {
    float4 xv[NV], dv[NV], rv[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c = k * 128 + lane * 4;
      xv[k] = *reinterpret_cast<const float4*>(x + row * D + c);
      const uint2 t = *reinterpret_cast<const uint2*>(dy + row * D + c);
      dv[k] = make_float4(bf16_lo(t.x), bf16_hi(t.x), bf16_lo(t.y), bf16_hi(t.y));
      if (RBF) {
        const uint2 r = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(dres_v) + row * D + c);
        rv[k] = make_float4(bf16_lo(r.x), bf16_hi(r.x), bf16_lo(r.y), bf16_hi(r.y));
      } else {
        const float* dres = reinterpret_cast<const float*>(dres_v);
        rv[k] = dres ? *reinterpret_cast<const float4*>(dres + row * D + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    const float mu = mean[row], rs = rstd[row];
    float c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      float* xe = reinterpret_cast<float*>(&xv[k]);
      float* de = reinterpret_cast<float*>(&dv[k]);
      const float* ge = reinterpret_cast<const float*>(&gm[k]);
      const float* be = reinterpret_cast<const float*>(&bt[k]);
      float* age = reinterpret_cast<float*>(&ag[k]);
      float* abe = reinterpret_cast<float*>(&ab[k]);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float xh = (xe[e] - mu) * rs;
        float d0 = de[e];
        if (GELU) d0 *= gelu_erf_grad(xh * ge[e] + be[e]);
        age[e] += d0 * xh;
        abe[e] += d0;
        const float g = d0 * ge[e];
        c1 += g;
        c2 += g * xh;
        xe[e] = xh;   // keep xhat
        de[e] = g;    // keep dy*gamma
      }
    }
    c1 = warp_sum(c1) * (1.0f / D);
    c2 = warp_sum(c2) * (1.0f / D);
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c = k * 128 + lane * 4;
      float4 o;
      o.x = rs * (dv[k].x - c1 - xv[k].x * c2) + rv[k].x;
      o.y = rs * (dv[k].y - c1 - xv[k].y * c2) + rv[k].y;
      o.z = rs * (dv[k].z - c1 - xv[k].z * c2) + rv[k].z;
      o.w = rs * (dv[k].w - c1 - xv[k].w * c2) + rv[k].w;
      ax[k].x += o.x; ax[k].y += o.y; ax[k].z += o.z; ax[k].w += o.w;
      if (dx_f32) *reinterpret_cast<float4*>(dx_f32 + row * D + c) = o;
      if (dx_bf16) *reinterpret_cast<uint2*>(dx_bf16 + row * D + c) = make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
    }
  }
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c = k * 128 + lane * 4;
    const float* age = reinterpret_cast<const float*>(&ag[k]);
    const float* abe = reinterpret_cast<const float*>(&ab[k]);
    const float* axe = reinterpret_cast<const float*>(&ax[k]);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      atomicAdd(&part[c + e], age[e]);
      atomicAdd(&part[D + c + e], abe[e]);
      if (dxsum) atomicAdd(&part[2 * D + c + e], axe[e]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    atomicAdd(dgamma + i, part[i]);
    atomicAdd(dbeta + i, part[D + i]);
    if (dxsum) atomicAdd(dxsum + i, part[2 * D + i]);
  }
}

// generic-width variant (d % 64 == 0, d <= 512), used for the 192-wide pix_decoder LayerNorm
template <bool GELU, bool RBF>
__global__ void __launch_bounds__(256)
layernorm_bwd_generic_kernel(const __nv_bfloat16* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ mean,
                             const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ beta,
                             const void* dres_v, float* dx_f32, __nv_bfloat16* __restrict__ dx_bf16,
                             float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dxsum, long long rows, int d) {
  extern __shared__ float part[];  // [3][d] block partials (dgamma | dbeta | column sums of dx)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int np = d >> 6;
  for (int i = threadIdx.x; i < 3 * d; i += blockDim.x) part[i] = 0.f;
  __syncthreads();
  float2 ag[kLnMaxPairs], ab[kLnMaxPairs], ax[kLnMaxPairs];
#pragma unroll
  for (int k = 0; k < kLnMaxPairs; ++k) { ag[k] = make_float2(0.f, 0.f); ab[k] = make_float2(0.f, 0.f); ax[k] = make_float2(0.f, 0.f); }
  float2 gm[kLnMaxPairs], bt[kLnMaxPairs];
#pragma unroll
  for (int k = 0; k < kLnMaxPairs; ++k)
    if (k < np) {
      gm[k] = *reinterpret_cast<const float2*>(gamma + k * 64 + lane * 2);
      bt[k] = GELU ? *reinterpret_cast<const float2*>(beta + k * 64 + lane * 2) : make_float2(0.f, 0.f);
    }
  for (long long row = (long long)blockIdx.x * nwarps + warp; row < rows; row += (long long)gridDim.x * nwarps) {
    const float mu = mean[row], rs = rstd[row];
    float2 xh[kLnMaxPairs], g[kLnMaxPairs];
    float c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int k = 0; k < kLnMaxPairs; ++k)
      if (k < np) {
        const int c = k * 64 + lane * 2;
        const float2 xv = *reinterpret_cast<const float2*>(x + row * d + c);
        const uint32_t dv = *reinterpret_cast<const uint32_t*>(dy + row * d + c);
        xh[k] = make_float2((xv.x - mu) * rs, (xv.y - mu) * rs);
        float d0 = bf16_lo(dv), d1 = bf16_hi(dv);
        if (GELU) {
          d0 *= gelu_erf_grad(xh[k].x * gm[k].x + bt[k].x);
          d1 *= gelu_erf_grad(xh[k].y * gm[k].y + bt[k].y);
        }
        ag[k].x += d0 * xh[k].x; ag[k].y += d1 * xh[k].y;
        ab[k].x += d0; ab[k].y += d1;
        g[k] = make_float2(d0 * gm[k].x, d1 * gm[k].y);
        c1 += g[k].x + g[k].y;
        c2 += g[k].x * xh[k].x + g[k].y * xh[k].y;
      }
    c1 = warp_sum(c1) / d;
    c2 = warp_sum(c2) / d;
#pragma unroll
    for (int k = 0; k < kLnMaxPairs; ++k)
      if (k < np) {
        const int c = k * 64 + lane * 2;
        float o0 = rs * (g[k].x - c1 - xh[k].x * c2);
        float o1 = rs * (g[k].y - c1 - xh[k].y * c2);
        if (RBF) {
          const uint32_t r = *reinterpret_cast<const uint32_t*>(reinterpret_cast<const __nv_bfloat16*>(dres_v) + row * d + c);
          o0 += bf16_lo(r); o1 += bf16_hi(r);
        } else if (dres_v) {
          const float2 r = *reinterpret_cast<const float2*>(reinterpret_cast<const float*>(dres_v) + row * d + c);
          o0 += r.x; o1 += r.y;
        }
        ax[k].x += o0; ax[k].y += o1;
        if (dx_f32) *reinterpret_cast<float2*>(dx_f32 + row * d + c) = make_float2(o0, o1);
        if (dx_bf16) *reinterpret_cast<uint32_t*>(dx_bf16 + row * d + c) = pack_bf16(o0, o1);
      }
  }
#pragma unroll
  for (int k = 0; k < kLnMaxPairs; ++k)
    if (k < np) {
      const int c = k * 64 + lane * 2;
      atomicAdd(&part[c], ag[k].x); atomicAdd(&part[c + 1], ag[k].y);
      atomicAdd(&part[d + c], ab[k].x); atomicAdd(&part[d + c + 1], ab[k].y);
      if (dxsum) { atomicAdd(&part[2 * d + c], ax[k].x); atomicAdd(&part[2 * d + c + 1], ax[k].y); }
    }
  __syncthreads();
  for (int i = threadIdx.x; i < d; i += blockDim.x) {
    atomicAdd(dgamma + i, part[i]);
    atomicAdd(dbeta + i, part[d + i]);
    if (dxsum) atomicAdd(dxsum + i, part[2 * d + i]);
  }
}

// ------------------------------------------------------------------------------------------------
// window pooling (PatchNet without patch transformer): [S,8,32,d] -> mean over 8 rows x 8 columns -> [S,4,d]
// sequences [0, split) are read from x0, the rest from x1 (masked view after pix_projector | augmented view)
// ------------------------------------------------------------------------------------------------
__global__ void pool_fwd_kernel(const float* __restrict__ x0, const float* __restrict__ x1, long long split, __nv_bfloat16* __restrict__ out,
                                long long num_seqs, int d, int num_windows) {
  const int dv = d >> 2;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= num_seqs * num_windows * dv) return;
  const int c4 = (int)(idx % dv);
  const int w = (int)((idx / dv) % num_windows);
  const long long s = idx / ((long long)dv * num_windows);
  const float* src = (s < split) ? (x0 + s * 256 * d) : (x1 + (s - split) * 256 * d);
  const int wcols = 32 / num_windows;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int h = 0; h < 8; ++h)
    for (int cc = 0; cc < wcols; ++cc) {
      const float4 v = *reinterpret_cast<const float4*>(src + (long long)(h * 32 + w * wcols + cc) * d + c4 * 4);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  const float inv = 1.0f / (8 * wcols);
  uint2 p;
  p.x = pack_bf16(acc.x * inv, acc.y * inv);
  p.y = pack_bf16(acc.z * inv, acc.w * inv);
  *reinterpret_cast<uint2*>(out + (s * num_windows + w) * d + c4 * 4) = p;
}

// dx[s, tok, :] = dpool[s, window(tok), :] / window_size, written to dx0 (s < split) or dx1
__global__ void pool_bwd_kernel(const float* __restrict__ dpool, long long split, float* __restrict__ dx0, float* __restrict__ dx1,
                                long long num_seqs, int d, int num_windows) {
  const int dv = d >> 2;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= num_seqs * 256 * dv) return;
  const int c4 = (int)(idx % dv);
  const int tok = (int)((idx / dv) & 255);
  const long long s = idx / ((long long)dv * 256);
  const int wcols = 32 / num_windows;
  const int w = (tok & 31) / wcols;
  const float inv = 1.0f / (8 * wcols);
  float4 v = *reinterpret_cast<const float4*>(dpool + (s * num_windows + w) * d + c4 * 4);
  v.x *= inv; v.y *= inv; v.z *= inv; v.w *= inv;
  float* dst = (s < split) ? (dx0 + (s * 256 + tok) * d) : (dx1 + ((s - split) * 256 + tok) * d);
  *reinterpret_cast<float4*>(dst + c4 * 4) = v;
}

// ------------------------------------------------------------------------------------------------
// row gather (fp32 -> bf16) and scatter-add (fp32 += fp32) by int32 row index
// ------------------------------------------------------------------------------------------------
__global__ void gather_rows_kernel(const float* __restrict__ x, const int* __restrict__ idx, __nv_bfloat16* __restrict__ out, long long n,
                                   int d) {
  const int dv = d >> 2;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * dv) return;
  const long long r = i / dv;
  const int c4 = (int)(i % dv);
  const float4 v = *reinterpret_cast<const float4*>(x + (long long)idx[r] * d + c4 * 4);
  uint2 p;
  p.x = pack_bf16(v.x, v.y);
  p.y = pack_bf16(v.z, v.w);
  *reinterpret_cast<uint2*>(out + r * d + c4 * 4) = p;
}

__global__ void scatter_add_rows_kernel(const float* __restrict__ src, const int* __restrict__ idx, float* __restrict__ dst, long long n,
                                        int d) {
  const int dv = d >> 2;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * dv) return;
  const long long r = i / dv;
  const int c4 = (int)(i % dv);
  const float4 v = *reinterpret_cast<const float4*>(src + r * d + c4 * 4);
  float4* o = reinterpret_cast<float4*>(dst + (long long)idx[r] * d + c4 * 4);
  float4 t = *o;
  t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
  *o = t;
}

// ------------------------------------------------------------------------------------------------
// column sums: out[n] += sum_m x[m, n]   (bias gradients).  Block = 32 column lanes x 8 row lanes.
// ------------------------------------------------------------------------------------------------
// Each thread owns VEC consecutive columns (one 128-bit load per row), a warp covers 32*VEC columns of a row.
template <typename T>
struct ColVec;
template <>
struct ColVec<float> {
  static constexpr int kVec = 4;
  static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
};
template <>
struct ColVec<__nv_bfloat16> {
  static constexpr int kVec = 8;
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[8]) {
    const uint4 t = *reinterpret_cast<const uint4*>(p);
    v[0] = bf16_lo(t.x); v[1] = bf16_hi(t.x); v[2] = bf16_lo(t.y); v[3] = bf16_hi(t.y);
    v[4] = bf16_lo(t.z); v[5] = bf16_hi(t.z); v[6] = bf16_lo(t.w); v[7] = bf16_hi(t.w);
  }
};

template <typename T, bool SQUARES, bool PEER>
__global__ void __launch_bounds__(256)
colsum_kernel(const T* __restrict__ x, long long ld, float* __restrict__ out, float* __restrict__ out_sq, long long rows, int cols,
              int rows_per_block, PeerReduce pr) {
  constexpr int V = ColVec<T>::kVec;
  __shared__ float sm[SQUARES ? 2 : 1][8][32 * V + 1];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int col = (blockIdx.x * 32 + cx) * V;
  const long long r0 = (long long)blockIdx.y * rows_per_block;
  const long long r1 = min(rows, r0 + rows_per_block);
  float s[V], q[V];
#pragma unroll
  for (int k = 0; k < V; ++k) { s[k] = 0.f; q[k] = 0.f; }
  if (col < cols)
    for (long long r = r0 + ry; r < r1; r += 8) {
      float v[V];
      ColVec<T>::load(x + r * ld + col, v);
#pragma unroll
      for (int k = 0; k < V; ++k) { s[k] += v[k]; if (SQUARES) q[k] += v[k] * v[k]; }
    }
#pragma unroll
  for (int k = 0; k < V; ++k) {
    sm[0][ry][cx * V + k] = s[k];
    if (SQUARES) sm[SQUARES ? 1 : 0][ry][cx * V + k] = q[k];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 32 * V; c += 256) {
    const int gc = blockIdx.x * 32 * V + c;
    if (gc < cols) {
      float ts = 0.f, tq = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) { ts += sm[0][k][c]; if (SQUARES) tq += sm[SQUARES ? 1 : 0][k][c]; }
      atomicAdd(out + gc, ts);
      if (SQUARES) atomicAdd(out_sq + gc, tq);
    }
  }
  if (PEER) peer_allreduce_grid_tail(pr);   // SyncBatchNorm: the grid's last block sums [sum | sumsq] over the ranks through NVLink peer memory
}

// ------------------------------------------------------------------------------------------------
// BatchNorm1d (training mode).  stats = [sum(C) | sumsq(C)] over `count` rows (possibly all-reduced across ranks).
// ------------------------------------------------------------------------------------------------
constexpr int kBnRows = 16;  // rows per thread in the BatchNorm elementwise kernels: per-column constants are computed once per 16 rows

__global__ void bn_apply_kernel(const float* __restrict__ x, const float* __restrict__ stats, float count, const float* __restrict__ gamma,
                                const float* __restrict__ beta, int relu, float eps, __nv_bfloat16* __restrict__ y_bf16,
                                float* __restrict__ y_f32, long long rows, int C) {
  // thread = 4 consecutive columns x kBnRows rows; consecutive threads take consecutive column groups (coalesced 16-byte accesses)
  const int cv = C >> 2;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long chunks = (rows + kBnRows - 1) / kBnRows;
  if (i >= chunks * cv) return;
  const int c = (int)(i % cv) * 4;
  const long long r0 = (i / cv) * kBnRows;
  const float4 s = *reinterpret_cast<const float4*>(stats + c);
  const float4 q = *reinterpret_cast<const float4*>(stats + C + c);
  const float inv = 1.0f / count;
  const float sv[4] = {s.x, s.y, s.z, s.w}, qv[4] = {q.x, q.y, q.z, q.w};
  float sc[4], sh[4];   // y = x * sc + sh
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float mu = sv[k] * inv;
    const float rs = rsqrtf(fmaxf(qv[k] * inv - mu * mu, 0.f) + eps);
    const float g = gamma ? gamma[c + k] : 1.f, bt = gamma ? beta[c + k] : 0.f;
    sc[k] = rs * g;
    sh[k] = bt - mu * rs * g;
  }
  const long long r1 = r0 + kBnRows < rows ? r0 + kBnRows : rows;
  for (long long r = r0; r < r1; ++r) {
    const float4 v = *reinterpret_cast<const float4*>(x + r * C + c);
    float o[4] = {fmaf(v.x, sc[0], sh[0]), fmaf(v.y, sc[1], sh[1]), fmaf(v.z, sc[2], sh[2]), fmaf(v.w, sc[3], sh[3])};
    if (relu) {
#pragma unroll
      for (int k = 0; k < 4; ++k) o[k] = fmaxf(o[k], 0.f);
    }
    if (y_bf16) {
      uint2 p;
      p.x = pack_bf16(o[0], o[1]);
      p.y = pack_bf16(o[2], o[3]);
      *reinterpret_cast<uint2*>(y_bf16 + r * C + c) = p;
    }
    if (y_f32) *reinterpret_cast<float4*>(y_f32 + r * C + c) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// running_mean/var update (momentum 0.1, unbiased variance), num_batches_tracked += 1
__global__ void bn_running_kernel(const float* __restrict__ stats, float count, float momentum, float* __restrict__ running_mean,
                                  float* __restrict__ running_var, long long* __restrict__ num_batches_tracked, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) {
    const float mu = stats[c] / count;
    const float var = fmaxf(stats[C + c] / count - mu * mu, 0.f);
    const float unbiased = var * (count / fmaxf(count - 1.f, 1.f));
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mu;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * unbiased;
  }
  if (c == 0 && num_batches_tracked) *num_batches_tracked += 1;
}

// backward statistics: bstats = [sum dy | sum dy*xhat]  (dy already masked by the ReLU of this layer)
template <bool PEER>
__global__ void __launch_bounds__(256)
bn_bwd_stats_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ stats, float count, float eps,
                    float* __restrict__ bstats, long long rows, int C, int rows_per_block, PeerReduce pr) {
  __shared__ float sm[2][8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + cx;
  const long long r0 = (long long)blockIdx.y * rows_per_block;
  const long long r1 = min(rows, r0 + rows_per_block);
  float s = 0.f, q = 0.f;
  if (col < C) {
    const float mu = stats[col] / count;
    const float rs = rsqrtf(fmaxf(stats[C + col] / count - mu * mu, 0.f) + eps);
    for (long long r = r0 + ry; r < r1; r += 8) {
      const float g = dy[r * C + col];
      s += g;
      q += g * (x[r * C + col] - mu) * rs;
    }
  }
  sm[0][ry][cx] = s;
  sm[1][ry][cx] = q;
  __syncthreads();
  if (ry == 0 && col < C) {
    float ts = 0.f, tq = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) { ts += sm[0][k][cx]; tq += sm[1][k][cx]; }
    atomicAdd(bstats + col, ts);
    atomicAdd(bstats + C + col, tq);
  }
  if (PEER) peer_allreduce_grid_tail(pr);   // SyncBatchNorm backward: [sum dy | sum dy xhat] over the ranks; the rank-local sums go to pr.loc0 / loc1
}

// dx = gamma * rstd * (dy - sum_dy/n - xhat * sum_dy_xhat/n)   (bstats/count possibly all-reduced across ranks)
__global__ void bn_bwd_apply_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ stats,
                                    const float* __restrict__ bstats, float count, const float* __restrict__ gamma, float eps,
                                    __nv_bfloat16* __restrict__ dx_bf16, float* __restrict__ dx_f32, long long rows, int C) {
  // same decomposition as bn_apply_kernel; dx = a * dy + b * x + c0 with per-column a = gamma rstd, b = -a rstd^2 S2/n, c0 = -a S1/n - b mu
  const int cv = C >> 2;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long chunks = (rows + kBnRows - 1) / kBnRows;
  if (i >= chunks * cv) return;
  const int c = (int)(i % cv) * 4;
  const long long r0 = (i / cv) * kBnRows;
  const float inv = 1.0f / count;
  float ca[4], cb[4], cc[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float mu = stats[c + k] * inv;
    const float rs = rsqrtf(fmaxf(stats[C + c + k] * inv - mu * mu, 0.f) + eps);
    const float g = gamma ? gamma[c + k] : 1.f;
    const float m1 = bstats[c + k] * inv, m2 = bstats[C + c + k] * inv;   // mean(dy), mean(dy * xhat)
    ca[k] = g * rs;
    cb[k] = -ca[k] * rs * m2;             // xhat = (x - mu) rs
    cc[k] = -ca[k] * m1 - cb[k] * mu;
  }
  const long long r1 = r0 + kBnRows < rows ? r0 + kBnRows : rows;
  for (long long r = r0; r < r1; ++r) {
    const float4 d4 = *reinterpret_cast<const float4*>(dy + r * C + c);
    const float4 x4 = *reinterpret_cast<const float4*>(x + r * C + c);
    const float o0 = fmaf(ca[0], d4.x, fmaf(cb[0], x4.x, cc[0])), o1 = fmaf(ca[1], d4.y, fmaf(cb[1], x4.y, cc[1]));
    const float o2 = fmaf(ca[2], d4.z, fmaf(cb[2], x4.z, cc[2])), o3 = fmaf(ca[3], d4.w, fmaf(cb[3], x4.w, cc[3]));
    if (dx_bf16) {
      uint2 p;
      p.x = pack_bf16(o0, o1);
      p.y = pack_bf16(o2, o3);
      *reinterpret_cast<uint2*>(dx_bf16 + r * C + c) = p;
    }
    if (dx_f32) *reinterpret_cast<float4*>(dx_f32 + r * C + c) = make_float4(o0, o1, o2, o3);
  }
}

// fp32 -> bf16 elementwise
__global__ void cast_f32_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, long long n4) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 v = reinterpret_cast<const float4*>(x)[i];
  uint2 p;
  p.x = pack_bf16(v.x, v.y);
  p.y = pack_bf16(v.z, v.w);
  reinterpret_cast<uint2*>(y)[i] = p;
}

// y (bf16) = row_mask[r] ? 0 : x  (gradient reaching the patch-embed GEMM: masked tokens were replaced by mask_token, V:95-97)
__global__ void zero_masked_rows_bf16_kernel(const __nv_bfloat16* __restrict__ x, const uint8_t* __restrict__ row_mask,
                                             __nv_bfloat16* __restrict__ y, long long rows, int d) {
  const int dv = d >> 3;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * dv) return;
  uint4 v = reinterpret_cast<const uint4*>(x)[i];
  if (row_mask[i / dv]) v = make_uint4(0u, 0u, 0u, 0u);
  reinterpret_cast<uint4*>(y)[i] = v;
}

__global__ void zero_masked_rows_kernel(const float* __restrict__ x, const uint8_t* __restrict__ row_mask, __nv_bfloat16* __restrict__ y,
                                        long long rows, int d) {
  const int dv = d >> 2;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * dv) return;
  const long long r = i / dv;
  float4 v = reinterpret_cast<const float4*>(x)[i];
  if (row_mask[r]) v = make_float4(0.f, 0.f, 0.f, 0.f);
  uint2 p;
  p.x = pack_bf16(v.x, v.y);
  p.y = pack_bf16(v.z, v.w);
  reinterpret_cast<uint2*>(y)[i] = p;
}

// mask u8 [B,256] -> idx[b*n_per + j] = b*256 + position of the j-th set bit (boolean-mask order, M:569); one warp per sample.
// err[0] is set when a sample does not have exactly n_per set bits.
__global__ void mask_to_index_kernel(const uint8_t* __restrict__ mask, int* __restrict__ idx, int* __restrict__ err, int B, int n_per) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  int base = 0;
  for (int t0 = 0; t0 < 256; t0 += 32) {
    const int set = mask[(long long)b * 256 + t0 + lane] != 0;
    const unsigned bal = __ballot_sync(0xffffffffu, set);
    const int pos = base + __popc(bal & ((1u << lane) - 1u));
    if (set && pos < n_per) idx[(long long)b * n_per + pos] = b * 256 + t0 + lane;
    base += __popc(bal);
  }
  // a sample with too few set bits: the rest of its index slots still get a valid row (its first token), so that the gather / scatter /
  // loss kernels downstream never see uninitialised indices; err tells the host, which raises
  for (int pos = base + lane; pos < n_per; pos += 32) idx[(long long)b * n_per + pos] = b * 256;
  if (lane == 0 && base != n_per) atomicExch(err, 1);
}

static inline int blocks_for(long long n, int per) { return (int)((n + per - 1) / per); }

}  // namespace dig

using namespace dig;

extern "C" int dig_im2col_patch4(const float* images, void* out, int64_t num_images, void* stream) {
  DIG_REQUIRE(images && out && num_images > 0, "dig_im2col_patch4: bad arguments");
  const long long total = (long long)num_images * 256 * 12;
  im2col_patch4_kernel<<<blocks_for(total, 256), 256, 0, (cudaStream_t)stream>>>(images, (__nv_bfloat16*)out, total);
  DIG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

namespace dig {
// LayerNorm forward for d = NC * 128 (the encoder widths 384 / 512): a warp walks rows with stride (grid-wide warp count), keeps gamma
// and beta in registers, reads each row with NC 128-bit loads per lane and writes NC 64-bit bf16x4 stores (the generic kernel above
// re-reads gamma/beta from L1 for every row and moves 8 bytes per load: 66 % issue-active at 4.3 TB/s).
template <int NC>
__global__ void __launch_bounds__(256)
layernorm_fwd_vec_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                         __nv_bfloat16* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out, long long rows, float eps) {
  constexpr int d = NC * 128;
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  float4 g[NC], b[NC];
#pragma unroll
  for (int k = 0; k < NC; ++k) {
    g[k] = *reinterpret_cast<const float4*>(gamma + k * 128 + lane * 4);
    b[k] = *reinterpret_cast<const float4*>(beta + k * 128 + lane * 4);
  }
  for (long long row = warp0; row < rows; row += nwarps) {
    const float* xr = x + row * d;
    float4 v[NC];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < NC; ++k) {
      v[k] = *reinterpret_cast<const float4*>(xr + k * 128 + lane * 4);
      s += (v[k].x + v[k].y) + (v[k].z + v[k].w);
    }
    const float mean = warp_sum(s) * (1.0f / d);
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < NC; ++k) {
      v[k].x -= mean; v[k].y -= mean; v[k].z -= mean; v[k].w -= mean;
      q += (v[k].x * v[k].x + v[k].y * v[k].y) + (v[k].z * v[k].z + v[k].w * v[k].w);
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / d) + eps);
#pragma unroll
    for (int k = 0; k < NC; ++k) {
      uint2 p;
      p.x = pack_bf16(fmaf(v[k].x * rstd, g[k].x, b[k].x), fmaf(v[k].y * rstd, g[k].y, b[k].y));
      p.y = pack_bf16(fmaf(v[k].z * rstd, g[k].z, b[k].z), fmaf(v[k].w * rstd, g[k].w, b[k].w));
      *reinterpret_cast<uint2*>(y + row * d + k * 128 + lane * 4) = p;
    }
    if (lane == 0) {
      if (mean_out) mean_out[row] = mean;
      if (rstd_out) rstd_out[row] = rstd;
    }
  }
}
}  // namespace dig

extern "C" int dig_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd, int64_t rows,
                                 int32_t d, float eps, int32_t gelu, void* stream) {
  DIG_REQUIRE(x && gamma && beta && y && rows > 0, "dig_layernorm_fwd: bad arguments");
  DIG_REQUIRE(d % 64 == 0 && d <= 512, "dig_layernorm_fwd: d must be a multiple of 64 and <= 512 (got %d)", d);
  if (!gelu && (d == 384 || d == 512) && (((uintptr_t)x | (uintptr_t)gamma | (uintptr_t)beta) & 15) == 0 && ((uintptr_t)y & 7) == 0) {
    long long want = (rows + 7) / 8;
    const int vgrid = (int)(want < (long long)num_sms() * 4 ? want : (long long)num_sms() * 4);   // 4 resident blocks of 8 warps per SM (58-70 registers)
    if (d == 384) DIG_CHECK_CUDA(launch_pdl(dig::layernorm_fwd_vec_kernel<3>, dim3(vgrid), dim3(256), 0, (cudaStream_t)stream, x, gamma, beta, (__nv_bfloat16*)y, mean, rstd, (long long)rows, eps));
    else DIG_CHECK_CUDA(launch_pdl(dig::layernorm_fwd_vec_kernel<4>, dim3(vgrid), dim3(256), 0, (cudaStream_t)stream, x, gamma, beta, (__nv_bfloat16*)y, mean, rstd, (long long)rows, eps));
    return 0;
  }
  const int grid = blocks_for(rows, 8);
  if (gelu) layernorm_fwd_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(x, gamma, beta, (__nv_bfloat16*)y, mean, rstd, rows, d, eps);
  else layernorm_fwd_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(x, gamma, beta, (__nv_bfloat16*)y, mean, rstd, rows, d, eps);
  DIG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

static int layernorm_bwd_launch(const void* dy, const float* x, const float* mean, const float* rstd, const float* gamma, const float* beta,
                                const void* dres, bool dres_bf16, float* dx_f32, void* dx_bf16, float* dgamma, float* dbeta, float* dxsum,
                                int64_t rows, int32_t d, int32_t gelu, void* stream) {
  DIG_REQUIRE(dy && x && mean && rstd && gamma && dgamma && dbeta && rows > 0, "dig_layernorm_bwd: bad arguments");
  DIG_REQUIRE(d % 64 == 0 && d <= 512, "dig_layernorm_bwd: d must be a multiple of 64 and <= 512 (got %d)", d);
  DIG_REQUIRE(!gelu || beta, "dig_layernorm_bwd: gelu variant needs beta");
  DIG_REQUIRE(!dres_bf16 || (dres && !gelu), "dig_layernorm_bwd_bf16res: needs dres and has no gelu variant");
  int grid = num_sms() * 4;
  if (grid > blocks_for(rows, 8)) grid = blocks_for(rows, 8);
  cudaStream_t s = (cudaStream_t)stream;
  const __nv_bfloat16* dyh = (const __nv_bfloat16*)dy;
  __nv_bfloat16* dxh = (__nv_bfloat16*)dx_bf16;
  const size_t sm = 3 * d * sizeof(float);
#define DIG_LN_ARGS dyh, x, mean, rstd, gamma, beta, dres, dx_f32, dxh, dgamma, dbeta, dxsum, (long long)rows
  if (dres_bf16) {
    if (d == 384) DIG_CHECK_CUDA(launch_pdl(layernorm_bwd_kernel<384, false, true>, dim3(grid), dim3(256), 0, s, DIG_LN_ARGS));
    else if (d == 512) DIG_CHECK_CUDA(launch_pdl(layernorm_bwd_kernel<512, false, true>, dim3(grid), dim3(256), 0, s, DIG_LN_ARGS));
    else layernorm_bwd_generic_kernel<false, true><<<grid, 256, sm, s>>>(DIG_LN_ARGS, d);
  } else if (!gelu && d == 384) {
    DIG_CHECK_CUDA(launch_pdl(layernorm_bwd_kernel<384, false, false>, dim3(grid), dim3(256), 0, s, DIG_LN_ARGS));
  } else if (!gelu && d == 512) {
    DIG_CHECK_CUDA(launch_pdl(layernorm_bwd_kernel<512, false, false>, dim3(grid), dim3(256), 0, s, DIG_LN_ARGS));
  } else {
    if (gelu) layernorm_bwd_generic_kernel<true, false><<<grid, 256, sm, s>>>(DIG_LN_ARGS, d);
    else layernorm_bwd_generic_kernel<false, false><<<grid, 256, sm, s>>>(DIG_LN_ARGS, d);
  }
#undef DIG_LN_ARGS
  DIG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int dig_layernorm_bwd(const void* dy, const float* x, const float* mean, const float* rstd, const float* gamma,
                                 const float* beta, const float* dres, float* dx_f32, void* dx_bf16, float* dgamma, float* dbeta,
                                 float* dxsum, int64_t rows, int32_t d, int32_t gelu, void* stream) {
  return layernorm_bwd_launch(dy, x, mean, rstd, gamma, beta, dres, false, dx_f32, dx_bf16, dgamma, dbeta, dxsum, rows, d, gelu, stream);
}

extern "C" int dig_layernorm_bwd_bf16res(const void* dy, const float* x, const float* mean, const float* rstd, const float* gamma,
                                         const void* dres_bf16, void* dx_bf16, float* dgamma, float* dbeta, float* dxsum, int64_t rows,
                                         int32_t d, void* stream) {
  DIG_REQUIRE(dres_bf16 && dx_bf16, "dig_layernorm_bwd_bf16res: dres_bf16 and dx_bf16 are required");
  return layernorm_bwd_launch(dy, x, mean, rstd, gamma, nullptr, dres_bf16, true, nullptr, dx_bf16, dgamma, dbeta, dxsum, rows, d, 0, stream);
}

extern "C" int dig_pool_fwd(const float* x0, const float* x1, int64_t split, void* out, int64_t num_seqs, int32_t d, int32_t num_windows,
                            void* stream) {
  DIG_REQUIRE(x0 && out && num_seqs > 0 && d % 4 == 0, "dig_pool_fwd: bad arguments");
  DIG_REQUIRE(num_windows > 0 && 32 % num_windows == 0, "dig_pool_fwd: num_windows must divide 32 (got %d)", num_windows);
  DIG_REQUIRE(split >= num_seqs || x1, "dig_pool_fwd: x1 missing");
  const long long total = (long long)num_seqs * num_windows * (d / 4);
  pool_fwd_kernel<<<blocks_for(total, 256), 256, 0, (cudaStream_t)stream>>>(x0, x1, split, (__nv_bfloat16*)out, num_seqs, d, num_windows);
  DIG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int dig_pool_bwd(const float* dpool, int64_t split, float* dx0, float* dx1, int64_t num_seqs, int32_t d, int32_t num_windows,
                            void* stream) {
  DIG_REQUIRE(dpool && dx0 && num_seqs > 0 && d % 4 == 0, "dig_pool_bwd: bad arguments");
  DIG_REQUIRE(num_windows > 0 && 32 % num_windows == 0, "dig_pool_bwd: num_windows must divide 32 (got %d)", num_windows);
  const long long total = (long long)num_seqs * 256 * (d / 4);
  pool_bwd_kernel<<<blocks_for(total, 256), 256, 0, (cudaStream_t)stream>>>(dpool, split, dx0, dx1, num_seqs, d, num_windows);
  DIG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int dig_gather_rows(const float* x, const int32_t* idx, void* out, int64_t n, int32_t d, void* stream) {
  DIG_REQUIRE(x && idx && out && d % 4 == 0, "dig_gather_rows: bad arguments");
  if (n == 0) return 0;
  gather_rows_kernel<<<blocks_for(n * (d / 4), 256), 256, 0, (cudaStream_t)stream>>>(x, idx, (__nv_bfloat16*)out, n, d);
  DIG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int dig_scatter_add_rows(const float* src, const int32_t* idx, float* dst, int64_t n, int32_t d, void* stream) {
  DIG_REQUIRE(src && idx && dst && d % 4 == 0, "dig_scatter_add_rows: bad arguments");
  if (n == 0) return 0;
  scatter_add_rows_kernel<<<blocks_for(n * (d / 4), 256), 256, 0, (cudaStream_t)stream>>>(src, idx, dst, n, d);
  DIG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

static int colsum_launch(const void* x, int32_t x_is_fp32, int64_t ld, float* out, float* out_sq, int64_t rows, int32_t cols, const PeerReduce& pr,
                         void* stream) {
  DIG_REQUIRE(x && out && rows > 0 && cols > 0, "dig_colsum: bad arguments");
  const int vec = x_is_fp32 ? 4 : 8;
  DIG_REQUIRE(cols % vec == 0 && ld % vec == 0 && ((uintptr_t)x & 15) == 0, "dig_colsum: cols/ld must be multiples of %d and x 16-byte aligned",
              vec);
  const int gx = (cols + 32 * vec - 1) / (32 * vec);
  int gy = (num_sms() * 4 + gx - 1) / gx;
  if (gy > (rows + 63) / 64) gy = (int)((rows + 63) / 64);
  if (gy < 1) gy = 1;
  const int rpb = (int)((rows + gy - 1) / gy);
  dim3 grid(gx, gy);
  cudaStream_t s = (cudaStream_t)stream;
  const bool peer = pr.world > 1;
  if (peer) {      // the fused exchange is built for the BatchNorm statistics call (fp32 input, sums and squares)
    DIG_REQUIRE(x_is_fp32 && out_sq, "dig_colsum: the peer variant needs fp32 input and out_sq");
    colsum_kernel<float, true, true><<<grid, 256, 0, s>>>((const float*)x, ld, out, out_sq, rows, cols, rpb, pr);
  } else if (x_is_fp32) {
    if (out_sq) colsum_kernel<float, true, false><<<grid, 256, 0, s>>>((const float*)x, ld, out, out_sq, rows, cols, rpb, pr);
    else colsum_kernel<float, false, false><<<grid, 256, 0, s>>>((const float*)x, ld, out, nullptr, rows, cols, rpb, pr);
  } else {
    if (out_sq) colsum_kernel<__nv_bfloat16, true, false><<<grid, 256, 0, s>>>((const __nv_bfloat16*)x, ld, out, out_sq, rows, cols, rpb, pr);
    else colsum_kernel<__nv_bfloat16, false, false><<<grid, 256, 0, s>>>((const __nv_bfloat16*)x, ld, out, nullptr, rows, cols, rpb, pr);
  }
  DIG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

static PeerReduce no_peer() {
  PeerReduce pr;
  memset(&pr, 0, sizeof(pr));
  return pr;
}

static int make_peer_reduce(PeerReduce* pr, const int64_t* bases, int32_t world, int32_t rank, int32_t channel, int64_t epoch, float* buf, int n,
                            float* loc0, float* loc1, int nloc) {
  DIG_REQUIRE(rank >= 0 && rank < world && channel >= 0 && channel < kPeerChannels && epoch >= 1, "peer: bad rank/channel/epoch");
  DIG_REQUIRE(n > 0 && n <= kPeerMaxFloats, "peer: message of %d floats exceeds %d", n, kPeerMaxFloats);
  if (int rc = fill_peer_table(&pr->pt, bases, world)) return rc;
  pr->world = world; pr->rank = rank; pr->channel = channel; pr->epoch = (uint32_t)epoch; pr->buf = buf; pr->n = n;
  pr->loc0 = loc0; pr->loc1 = loc1; pr->nloc = nloc;
  return 0;
}

extern "C" int dig_colsum(const void* x, int32_t x_is_fp32, int64_t ld, float* out, float* out_sq, int64_t rows, int32_t cols,
                          void* stream) {
  return colsum_launch(x, x_is_fp32, ld, out, out_sq, rows, cols, no_peer(), stream);
}

extern "C" int dig_bn_stats_allreduce(const float* x, float* stats, int64_t rows, int32_t C, const int64_t* bases, int32_t world, int32_t rank,
                                      int32_t channel, int64_t epoch, void* stream) {
  DIG_REQUIRE(stats != nullptr && world > 1, "dig_bn_stats_allreduce: needs a stats buffer and world > 1");
  PeerReduce pr;
  if (int rc = make_peer_reduce(&pr, bases, world, rank, channel, epoch, stats, 2 * C, nullptr, nullptr, 0)) return rc;
  return colsum_launch(x, 1, C, stats, stats + C, rows, C, pr, stream);
}

extern "C" int dig_bn_apply(const float* x, const float* stats, float count, const float* gamma, const float* beta, int32_t relu, float eps,
                            void* y_bf16, float* y_f32, int64_t rows, int32_t C, void* stream) {
  DIG_REQUIRE(x && stats && rows > 0 && C % 4 == 0 && count > 0, "dig_bn_apply: bad arguments");
  DIG_REQUIRE((gamma == nullptr) == (beta == nullptr), "dig_bn_apply: gamma and beta must both be given or both be NULL");
  bn_apply_kernel<<<blocks_for(((rows + kBnRows - 1) / kBnRows) * (C / 4), 256), 256, 0, (cudaStream_t)stream>>>(x, stats, count, gamma, beta, relu, eps,
                                                                                   (__nv_bfloat16*)y_bf16, y_f32, rows, C);
  DIG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int dig_bn_running(const float* stats, float count, float momentum, float* running_mean, float* running_var,
                              int64_t* num_batches_tracked, int32_t C, void* stream) {
  DIG_REQUIRE(stats && running_mean && running_var && C > 0 && count > 0, "dig_bn_running: bad arguments");
  bn_running_kernel<<<blocks_for(C, 256), 256, 0, (cudaStream_t)stream>>>(stats, count, momentum, running_mean, running_var,
                                                                        (long long*)num_batches_tracked, C);
  DIG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

static int bn_bwd_stats_launch(const float* dy, const float* x, const float* stats, float count, float eps, float* bstats, int64_t rows,
                               int32_t C, const PeerReduce& pr, void* stream) {
  DIG_REQUIRE(dy && x && stats && bstats && rows > 0 && C > 0 && count > 0, "dig_bn_bwd_stats: bad arguments");
  const int gx = (C + 31) / 32;
  int gy = (num_sms() * 8 + gx - 1) / gx;
  if (gy > (rows + 63) / 64) gy = (int)((rows + 63) / 64);
  if (gy < 1) gy = 1;
  const int rpb = (int)((rows + gy - 1) / gy);
  if (pr.world > 1) bn_bwd_stats_kernel<true><<<dim3(gx, gy), 256, 0, (cudaStream_t)stream>>>(dy, x, stats, count, eps, bstats, rows, C, rpb, pr);
  else bn_bwd_stats_kernel<false><<<dim3(gx, gy), 256, 0, (cudaStream_t)stream>>>(dy, x, stats, count, eps, bstats, rows, C, rpb, pr);
  DIG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int dig_bn_bwd_stats(const float* dy, const float* x, const float* stats, float count, float eps, float* bstats, int64_t rows,
                                int32_t C, void* stream) {
  return bn_bwd_stats_launch(dy, x, stats, count, eps, bstats, rows, C, no_peer(), stream);
}

extern "C" int dig_bn_bwd_stats_allreduce(const float* dy, const float* x, const float* stats, float count, float eps, float* bstats,
                                          float* dbeta_local, float* dgamma_local, int64_t rows, int32_t C, const int64_t* bases,
                                          int32_t world, int32_t rank, int32_t channel, int64_t epoch, void* stream) {
  DIG_REQUIRE(world > 1, "dig_bn_bwd_stats_allreduce: world must be > 1");
  PeerReduce pr;
  if (int rc = make_peer_reduce(&pr, bases, world, rank, channel, epoch, bstats, 2 * C, dbeta_local, dgamma_local, C)) return rc;
  return bn_bwd_stats_launch(dy, x, stats, count, eps, bstats, rows, C, pr, stream);
}

extern "C" int dig_bn_bwd_apply(const float* dy, const float* x, const float* stats, const float* bstats, float count, const float* gamma,
                                float eps, void* dx_bf16, float* dx_f32, int64_t rows, int32_t C, void* stream) {
  DIG_REQUIRE(dy && x && stats && bstats && rows > 0 && C > 0 && C % 4 == 0 && count > 0, "dig_bn_bwd_apply: bad arguments");
  bn_bwd_apply_kernel<<<blocks_for(((rows + kBnRows - 1) / kBnRows) * (C / 4), 256), 256, 0, (cudaStream_t)stream>>>(dy, x, stats, bstats, count, gamma, eps,
                                                                                 (__nv_bfloat16*)dx_bf16, dx_f32, rows, C);
  DIG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int dig_cast_f32_bf16(const float* x, void* y, int64_t n, void* stream) {
  DIG_REQUIRE(x && y && n > 0 && n % 4 == 0, "dig_cast_f32_bf16: n must be a positive multiple of 4");
  cast_f32_bf16_kernel<<<blocks_for(n / 4, 256), 256, 0, (cudaStream_t)stream>>>(x, (__nv_bfloat16*)y, n / 4);
  DIG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int dig_zero_masked_rows(const float* x, const uint8_t* row_mask, void* y, int64_t rows, int32_t d, void* stream) {
  DIG_REQUIRE(x && row_mask && y && rows > 0 && d % 4 == 0, "dig_zero_masked_rows: bad arguments");
  zero_masked_rows_kernel<<<blocks_for(rows * (d / 4), 256), 256, 0, (cudaStream_t)stream>>>(x, row_mask, (__nv_bfloat16*)y, rows, d);
  DIG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int dig_zero_masked_rows_bf16(const void* x, const uint8_t* row_mask, void* y, int64_t rows, int32_t d, void* stream) {
  DIG_REQUIRE(x && row_mask && y && rows > 0 && d % 8 == 0, "dig_zero_masked_rows_bf16: bad arguments");
  zero_masked_rows_bf16_kernel<<<blocks_for(rows * (d / 8), 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, row_mask,
                                                                                                  (__nv_bfloat16*)y, rows, d);
  DIG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int dig_mask_to_index(const uint8_t* mask, int32_t* idx, int32_t* err, int32_t B, int32_t n_per, void* stream) {
  DIG_REQUIRE(mask && idx && err && B > 0 && n_per >= 0 && n_per <= 256, "dig_mask_to_index: bad arguments");
  mask_to_index_kernel<<<blocks_for(B, 8), 256, 0, (cudaStream_t)stream>>>(mask, idx, err, B, n_per);
  DIG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// dig_b200 -- loss-side kernels, fp32 (the InfoNCE logits carry a 1/T = 5 gain: they run on tcgen05 as a bf16 x 3 split GEMM):
//   L2 row normalisation fwd/bwd (F.normalize, M:446-447), a small fp32 tiled GEMM for the q.k^T logits and their
//   gradient (torch.einsum, M:451), the fused row-wise cross-entropy / top-k accuracy (M:453-461, M:593-625) and the
//   masked-pixel MSE with on-the-fly target patchify (E:85-111, E:141).
#include "common.cuh"
#include "../../include/dig_b200.h"

namespace dig {

// ---- L2 normalise: warp per row ---------------------------------------------------------------------
__global__ void l2norm_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, float* __restrict__ inv_norm, long long rows, int C) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) { const float v = x[row * C + c]; s += v * v; }
  s = warp_sum(s);
  const float inv = 1.0f / fmaxf(sqrtf(s), 1e-12f);
  for (int c = lane; c < C; c += 32) y[row * C + c] = x[row * C + c] * inv;
  if (lane == 0 && inv_norm) inv_norm[row] = inv;
}

// dx = (dy - y * <y, dy>) * inv_norm * gscale[0]
__global__ void l2norm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, const float* __restrict__ inv_norm,
                                  const float* __restrict__ gscale, float* __restrict__ dx, long long rows, int C) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += y[row * C + c] * dy[row * C + c];
  s = warp_sum(s);
  const float k = inv_norm[row] * (gscale ? gscale[0] : 1.f);
  for (int c = lane; c < C; c += 32) dx[row * C + c] = (dy[row * C + c] - y[row * C + c] * s) * k;
}

// ---- bf16 x 3 operand split for the InfoNCE logits on tensor cores --------------------------------------
// x = hi + lo + O(2^-17 |x|) with hi = bf16(x), lo = bf16(x - hi).  q.k = qh.kh + qh.kl + ql.kh + O(2^-16): ONE bf16 tcgen05 GEMM over
// the 3x longer reduction axis [hi | hi | lo] . [hi | lo | hi] reproduces the fp32 logits (gain 1/T = 5, M:451) to ~1e-5 -- every bf16
// product is exact in the fp32 TMEM accumulator.
//   out_cat   [rows, 3*cols]            parts side by side (K-major operand); part `cat_lo_part` holds lo, the other two hi
//   out_stack [rows/stack_rows, 3, stack_rows, cols]   planes (hi, lo, hi) per batch (MN-major B operand of the gradient GEMM)
__global__ void __launch_bounds__(256)
split_bf16x3_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out_cat, int cat_lo_part, __nv_bfloat16* __restrict__ out_stack,
                    long long stack_rows, long long rows, int cols) {
  const long long i4 = (long long)blockIdx.x * blockDim.x + threadIdx.x;     // one float4 per thread
  const int c4n = cols >> 2;
  if (i4 >= rows * c4n) return;
  const long long r = i4 / c4n;
  const int c = (int)(i4 % c4n) * 4;
  const float4 v = *reinterpret_cast<const float4*>(x + r * cols + c);
  const uint2 hi = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
  const uint2 lo = make_uint2(pack_bf16(v.x - bf16_lo(hi.x), v.y - bf16_hi(hi.x)), pack_bf16(v.z - bf16_lo(hi.y), v.w - bf16_hi(hi.y)));
  if (out_cat != nullptr) {
    __nv_bfloat16* o = out_cat + r * 3 * cols + c;
#pragma unroll
    for (int j = 0; j < 3; ++j) *reinterpret_cast<uint2*>(o + (long long)j * cols) = (j == cat_lo_part) ? lo : hi;
  }
  if (out_stack != nullptr) {
    const long long b = r / stack_rows, rr = r % stack_rows;
    __nv_bfloat16* o = out_stack + ((b * 3) * stack_rows + rr) * cols + c;
    *reinterpret_cast<uint2*>(o) = hi;
    *reinterpret_cast<uint2*>(o + stack_rows * cols) = lo;
    *reinterpret_cast<uint2*>(o + 2 * stack_rows * cols) = hi;
  }
}

// ---- fp32 GEMM on CUDA cores: C[M,N] = alpha * A[M,K] . (B_NT ? B[N,K]^T : B[K,N]) ----------------------
// The InfoNCE logits (q.k^T / T, gain 5 on unit vectors) and their gradient stay in fp32; the problem is tiny (2 x 512 x 512W x 256 MAC),
// so the tiles are small (32 x 32 outputs per CTA, 256 threads, 2 x 2 per thread) to put a few hundred CTAs on the 148 SMs, and every
// global read is a 128-bit load along the contiguous dimension.  Requires K % 4 == 0 (and N % 4 == 0 for the [K,N] form).
template <bool B_NT>
__global__ void __launch_bounds__(256)
sgemm_f32_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C, int M, int N, int K, float alpha) {
  __shared__ float sA[32][32 + 1];   // [k][m]
  __shared__ float sB[32][32 + 1];   // [k][n]
  const int t = threadIdx.x;
  const int tx = t & 15, ty = t >> 4;
  const int m0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  float acc[2][2] = {};
  for (int k0 = 0; k0 < K; k0 += 32) {
    {
      const int r = t >> 3, c4 = (t & 7) * 4;   // 32 rows x 8 float4 along K
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m0 + r < M && k0 + c4 < K) v = *reinterpret_cast<const float4*>(A + (long long)(m0 + r) * K + k0 + c4);
      sA[c4 + 0][r] = v.x; sA[c4 + 1][r] = v.y; sA[c4 + 2][r] = v.z; sA[c4 + 3][r] = v.w;
      float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
      if (B_NT) {
        if (n0 + r < N && k0 + c4 < K) w = *reinterpret_cast<const float4*>(B + (long long)(n0 + r) * K + k0 + c4);
        sB[c4 + 0][r] = w.x; sB[c4 + 1][r] = w.y; sB[c4 + 2][r] = w.z; sB[c4 + 3][r] = w.w;
      } else {                                   // 32 k-rows x 8 float4 along N
        if (k0 + r < K && n0 + c4 < N) w = *reinterpret_cast<const float4*>(B + (long long)(k0 + r) * N + n0 + c4);
        sB[r][c4 + 0] = w.x; sB[r][c4 + 1] = w.y; sB[r][c4 + 2] = w.z; sB[r][c4 + 3] = w.w;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 32; ++kk) {
      const float a0 = sA[kk][ty * 2], a1 = sA[kk][ty * 2 + 1], b0 = sB[kk][tx * 2], b1 = sB[kk][tx * 2 + 1];
      acc[0][0] = fmaf(a0, b0, acc[0][0]); acc[0][1] = fmaf(a0, b1, acc[0][1]);
      acc[1][0] = fmaf(a1, b0, acc[1][0]); acc[1][1] = fmaf(a1, b1, acc[1][1]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int gm = m0 + ty * 2 + i, gn = n0 + tx * 2;
    if (gm < M && gn + 1 < N) *reinterpret_cast<float2*>(C + (long long)gm * N + gn) = make_float2(acc[i][0] * alpha, acc[i][1] * alpha);
    else if (gm < M && gn < N) C[(long long)gm * N + gn] = acc[i][0] * alpha;
  }
}

// ---- InfoNCE row pass: one block per query row over logits[Q, Nk] (already divided by T) ---------------
// result[0] += (lse - z_label) * 2T/Q ; result[1] += 100/Q * [rank<1] ; result[2] += 100/Q * [rank<5]
// logits row is overwritten with d(loss)/d(q.k) = (2/Q) * (softmax - onehot)
__global__ void __launch_bounds__(256)
infonce_rows_kernel(float* __restrict__ logits, int Q, int Nk, long long label_offset, float T, float* __restrict__ result) {
  __shared__ float red[8];
  __shared__ float bcast;
  const int row = blockIdx.x;
  float* z = logits + (long long)row * Nk;
  const int label = (int)(label_offset + row);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float mx = -INFINITY;
  for (int c = threadIdx.x; c < Nk; c += 256) mx = fmaxf(mx, z[c]);
  mx = warp_max(mx);
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  if (threadIdx.x == 0) { float m = red[0]; for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]); bcast = m; }
  __syncthreads();
  mx = bcast;
  const float zl = z[label];
  float s = 0.f, rank = 0.f;
  for (int c = threadIdx.x; c < Nk; c += 256) {
    const float v = z[c];
    s += __expf(v - mx);
    rank += (v > zl) ? 1.f : 0.f;
  }
  s = warp_sum(s);
  rank = warp_sum(rank);
  __syncthreads();
  if (lane == 0) red[warp] = s;
  __syncthreads();
  if (threadIdx.x == 0) { float t = 0.f; for (int i = 0; i < 8; ++i) t += red[i]; bcast = t; }
  __syncthreads();
  const float sum = bcast;
  __syncthreads();
  if (lane == 0) red[warp] = rank;
  __syncthreads();
  if (threadIdx.x == 0) {
    float r = 0.f;
    for (int i = 0; i < 8; ++i) r += red[i];
    atomicAdd(result + 0, (mx + logf(sum) - zl) * (2.f * T / Q));
    if (r < 1.f) atomicAdd(result + 1, 100.f / Q);
    if (r < 5.f) atomicAdd(result + 2, 100.f / Q);
  }
  const float inv = 1.f / sum, coef = 2.f / Q;
  for (int c = threadIdx.x; c < Nk; c += 256) {
    const float p = __expf(z[c] - mx) * inv;
    z[c] = coef * (p - (c == label ? 1.f : 0.f));
  }
}

// ---- masked-pixel MSE (E:85-111,141): target = un-normalised RGB patch '(p1 p2 c)' of view 0 at token row idx[r] -----
// loss[0] += sum (pred - target)^2 / numel ; dpred = 2 (pred - target) / numel
// One thread per (row, p1): the 12 consecutive prediction columns (p2, c) of one patch line = three 128-bit loads of pred, three 128-bit
// loads of the image (the four p2 pixels of each colour plane are contiguous and 16-byte aligned), three 128-bit stores of dpred.
// NORM (normlize_target=True, E:89-94): every colour plane of a patch is standardised over its 16 pixels first -- mean and UNBIASED
// variance, target = (x - mean) / (sqrt(var) + 1e-6).  The four lines of a patch sit in four adjacent lanes: two xor-shuffles per sum.
template <bool NORM>
__global__ void __launch_bounds__(256)
masked_mse_kernel(const float* __restrict__ pred, const float* __restrict__ images, const int* __restrict__ idx, float* __restrict__ loss,
                  float* __restrict__ dpred, long long n_rows) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const float inv_numel = 1.0f / (float)(n_rows * 48);
  float e = 0.f;
  const bool active = i < n_rows * 4;       // whole groups of four lanes are active or not (n_rows * 4 is a multiple of 4)
  float t[3][4] = {};
  long long r = 0;
  int p1 = 0;
  if (active) {
    r = i >> 2;
    p1 = (int)(i & 3);
    const int tok_row = idx[r];
    const long long b = tok_row >> 8;
    const int tok = tok_row & 255, ph = tok >> 5, pw = tok & 31;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(images + ((b * 3 + c) * 32 + ph * 4 + p1) * 128 + pw * 4));
      t[c][0] = v.x * 0.5f + 0.5f; t[c][1] = v.y * 0.5f + 0.5f; t[c][2] = v.z * 0.5f + 0.5f; t[c][3] = v.w * 0.5f + 0.5f;
    }
  }
  if (NORM) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float s = t[c][0] + t[c][1] + t[c][2] + t[c][3];
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      const float mean = s * (1.0f / 16.0f);
      float q = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) { t[c][k] -= mean; q += t[c][k] * t[c][k]; }
      q += __shfl_xor_sync(0xffffffffu, q, 1);
      q += __shfl_xor_sync(0xffffffffu, q, 2);
      const float inv = 1.0f / (sqrtf(q * (1.0f / 15.0f)) + 1e-6f);
#pragma unroll
      for (int k = 0; k < 4; ++k) t[c][k] *= inv;
    }
  }
  if (active) {
    const float* pr = pred + r * 48 + p1 * 12;
    float d[12];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const float4 v = *reinterpret_cast<const float4*>(pr + 4 * q);
      d[4 * q] = v.x; d[4 * q + 1] = v.y; d[4 * q + 2] = v.z; d[4 * q + 3] = v.w;
    }
#pragma unroll
    for (int j = 0; j < 12; ++j) {     // column j of the line = (p2 = j / 3, c = j % 3)
      d[j] -= t[j % 3][j / 3];
      e += d[j] * d[j];
    }
    if (dpred) {
      float* dp = dpred + r * 48 + p1 * 12;
      const float k = 2.f * inv_numel;
#pragma unroll
      for (int q = 0; q < 3; ++q)
        *reinterpret_cast<float4*>(dp + 4 * q) = make_float4(d[4 * q] * k, d[4 * q + 1] * k, d[4 * q + 2] * k, d[4 * q + 3] * k);
    }
  }
  e = warp_sum(e);
  __shared__ float red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = e;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    atomicAdd(loss, t * inv_numel);
  }
}

// y[i] = x[i] * s[0]  (apply an upstream scalar gradient that lives on the device)
__global__ void scale_by_device_scalar_kernel(const float* __restrict__ x, const float* __restrict__ s, float* __restrict__ y, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = x[i] * s[0];
}

}  // namespace dig

using namespace dig;

extern "C" int dig_l2norm_fwd(const float* x, float* y, float* inv_norm, int64_t rows, int32_t C, void* stream) {
  DIG_REQUIRE(x && y && rows > 0 && C > 0, "dig_l2norm_fwd: bad arguments");
  l2norm_fwd_kernel<<<(int)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(x, y, inv_norm, rows, C);
  DIG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int dig_l2norm_bwd(const float* dy, const float* y, const float* inv_norm, const float* gscale, float* dx, int64_t rows,
                              int32_t C, void* stream) {
  DIG_REQUIRE(dy && y && inv_norm && dx && rows > 0 && C > 0, "dig_l2norm_bwd: bad arguments");
  l2norm_bwd_kernel<<<(int)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(dy, y, inv_norm, gscale, dx, rows, C);
  DIG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int dig_split_bf16x3(const float* x, void* out_cat, int32_t cat_lo_part, void* out_stack, int64_t stack_rows, int64_t rows,
                                int32_t cols, void* stream) {
  DIG_REQUIRE(x && (out_cat || out_stack) && rows > 0 && cols > 0 && cols % 4 == 0, "dig_split_bf16x3: bad arguments (cols must be a multiple of 4)");
  DIG_REQUIRE(cat_lo_part >= 0 && cat_lo_part < 3, "dig_split_bf16x3: cat_lo_part must be 0, 1 or 2");
  DIG_REQUIRE(!out_stack || (stack_rows > 0 && rows % stack_rows == 0), "dig_split_bf16x3: rows must be a multiple of stack_rows");
  DIG_REQUIRE(((((uintptr_t)x) & 15) | (((uintptr_t)out_cat) & 7) | (((uintptr_t)out_stack) & 7)) == 0, "dig_split_bf16x3: misaligned operand");
  const long long n4 = rows * (cols >> 2);
  split_bf16x3_kernel<<<(int)((n4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, (__nv_bfloat16*)out_cat, cat_lo_part, (__nv_bfloat16*)out_stack,
                                                                               stack_rows > 0 ? stack_rows : 1, rows, cols);
  DIG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int dig_sgemm_f32(const float* A, const float* B, float* C, int32_t M, int32_t N, int32_t K, int32_t b_is_nk, float alpha,
                             void* stream) {
  DIG_REQUIRE(A && B && C && M > 0 && N > 0 && K > 0, "dig_sgemm_f32: bad arguments");
  DIG_REQUIRE(K % 4 == 0 && N % 2 == 0 && (b_is_nk || N % 4 == 0), "dig_sgemm_f32: K must be a multiple of 4 and N of %d (M=%d N=%d K=%d)",
              b_is_nk ? 2 : 4, M, N, K);
  DIG_REQUIRE((((uintptr_t)A | (uintptr_t)B) & 15) == 0 && ((uintptr_t)C & 7) == 0, "dig_sgemm_f32: operands must be 16-byte aligned");
  dim3 grid((N + 31) / 32, (M + 31) / 32);
  if (b_is_nk) sgemm_f32_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(A, B, C, M, N, K, alpha);
  else sgemm_f32_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(A, B, C, M, N, K, alpha);
  DIG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int dig_infonce_rows(float* logits, int32_t Q, int32_t Nk, int64_t label_offset, float T, float* result, void* stream) {
  DIG_REQUIRE(logits && result && Q > 0 && Nk > 0, "dig_infonce_rows: bad arguments");
  DIG_REQUIRE(label_offset >= 0 && label_offset + Q <= Nk, "dig_infonce_rows: labels [%lld, %lld) outside %d keys", (long long)label_offset,
              (long long)label_offset + Q, Nk);
  infonce_rows_kernel<<<Q, 256, 0, (cudaStream_t)stream>>>(logits, Q, Nk, label_offset, T, result);
  DIG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int dig_masked_mse(const float* pred, const float* images, const int32_t* idx, float* loss, float* dpred, int64_t n_rows,
                              int32_t normalize_target, void* stream) {
  DIG_REQUIRE(pred && images && idx && loss && n_rows > 0, "dig_masked_mse: bad arguments");
  DIG_REQUIRE(((((uintptr_t)pred) | ((uintptr_t)images) | ((uintptr_t)dpred)) & 15) == 0, "dig_masked_mse: pred, images and dpred must be 16-byte aligned");
  if (normalize_target) masked_mse_kernel<true><<<(int)((n_rows * 4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(pred, images, idx, loss, dpred, n_rows);
  else masked_mse_kernel<false><<<(int)((n_rows * 4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(pred, images, idx, loss, dpred, n_rows);
  DIG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int dig_scale_by_device_scalar(const float* x, const float* s, float* y, int64_t n, void* stream) {
  DIG_REQUIRE(x && s && y && n > 0, "dig_scale_by_device_scalar: bad arguments");
  scale_by_device_scalar_kernel<<<(int)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, s, y, n);
  DIG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// dig_b200 -- kernels of the fine-tuning step's transformer decoder (reference models/decoder.py:107-222,
// models/transformer_layer.py:47-118 and :204-281, loss/seqCrossEntropyLoss.py:19-63).  The decoder's Linears run on the tcgen05 GEMMs
// (dig_gemm); what is left is small and irregular and runs on CUDA cores in fp32:
//   * target embedding + sinusoid position table with the shift-right <BOS> (decoder.py:212-214, :173-178) and its backward
//   * multi-head attention with T <= 32 queries per (sample, head): decoder self-attention (T x T keys, causal & length mask,
//     transformer_layer.py:433-456) and encoder-decoder attention (T x 256 keys, no mask), forward and backward.  The query count is
//     25 (max_len), i.e. 0.5 % of the step's FLOPs: one CTA per (sample, head) with K and V of that head resident in shared memory.
//   * SeqCrossEntropyLoss with sample_normalize (loss summed over valid positions / batch) fused with its gradient and the arg-max.
#include "common.cuh"
#include "../../include/dig_b200.h"

namespace dig {

static constexpr int kDecHd = 64;        // d_k = d_v (decoder.py:141-142)
static constexpr int kDecMaxQ = 32;      // queries per (sample, head)
static constexpr int kDecMaxK = 256;     // keys per (sample, head)
static constexpr int kDecThreads = 128;
static constexpr int kKRow = kDecHd + 2; // padded bf16 row (132 bytes): lane-per-key reads hit distinct banks

// x[b*T + t, :] = emb[tok, :] + pos[t, :], tok = (t == 0 ? start_idx : targets[b, t-1])
__global__ void embed_pos_fwd_kernel(const long long* __restrict__ targets, const float* __restrict__ emb, const float* __restrict__ pos,
                                     float* __restrict__ x, int B, int T, int D, int start_idx) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int dv = D >> 2;
  if (i >= (long long)B * T * dv) return;
  const int c = (int)(i % dv) * 4;
  const long long row = i / dv;
  const int t = (int)(row % T);
  const long long b = row / T;
  const long long tok = t == 0 ? start_idx : targets[b * T + t - 1];
  const float4 e = *reinterpret_cast<const float4*>(emb + tok * D + c);
  const float4 p = *reinterpret_cast<const float4*>(pos + (long long)t * D + c);
  *reinterpret_cast<float4*>(x + row * D + c) = make_float4(e.x + p.x, e.y + p.y, e.z + p.z, e.w + p.w);
}

__global__ void embed_bwd_kernel(const float* __restrict__ dx, const long long* __restrict__ targets, float* __restrict__ demb, int B, int T,
                                 int D, int start_idx) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * T * D) return;
  const int c = (int)(i % D);
  const long long row = i / D;
  const int t = (int)(row % T);
  const long long b = row / T;
  const long long tok = t == 0 ? start_idx : targets[b * T + t - 1];
  atomicAdd(demb + tok * D + c, dx[i]);
}

struct DecAttnArgs {
  const __nv_bfloat16 *q, *k, *v;      // q [B*Lq, ldq], k/v [B*Lk, ldk/ldv]; head h = columns h*64 .. h*64+63
  long long ldq, ldk, ldv;
  __nv_bfloat16* out;                  // [B*Lq, ldo]
  long long ldo;
  float* lse;                          // [B, H, Lq]
  const long long* lens;               // self-attention: key j visible to query i iff j <= i && j < lens[b]; NULL: no mask
  float* maps;                         // optional [B, Lq, Lk]: += softmax / H (mean over heads, transformer_layer.py:270)
  int B, H, Lq, Lk;
  float scale;
};

__device__ __forceinline__ void load_kv_tile(const __nv_bfloat16* src, long long ld, int rows, __nv_bfloat16* dst) {
  // rows x 64 bf16 -> padded rows of kKRow; 32-bit copies, coalesced along the row
  for (int i = threadIdx.x; i < rows * 32; i += blockDim.x) {
    const int r = i >> 5, c2 = i & 31;
    reinterpret_cast<uint32_t*>(dst + r * kKRow)[c2] = *reinterpret_cast<const uint32_t*>(src + (long long)r * ld + c2 * 2);
  }
}

__global__ void __launch_bounds__(kDecThreads)
dec_attn_fwd_kernel(DecAttnArgs a) {
  extern __shared__ uint8_t smem[];
  __nv_bfloat16* sK = reinterpret_cast<__nv_bfloat16*>(smem);
  const int LkP = (a.Lk + 3) & ~3;                               // shared-memory sizes follow the actual key count (occupancy)
  __nv_bfloat16* sV = sK + LkP * kKRow;
  float* sQ = reinterpret_cast<float*>(sV + LkP * kKRow);        // [4 warps][64]
  float* sP = sQ + 4 * kDecHd;                                  // [4 warps][LkP]
  const int b = blockIdx.x / a.H, h = blockIdx.x % a.H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  load_kv_tile(a.k + (long long)b * a.Lk * a.ldk + h * kDecHd, a.ldk, a.Lk, sK);
  load_kv_tile(a.v + (long long)b * a.Lk * a.ldv + h * kDecHd, a.ldv, a.Lk, sV);
  __syncthreads();
  const int len = a.lens ? (int)a.lens[b] : a.Lk;
  float* q = sQ + warp * kDecHd;
  float* p = sP + warp * LkP;
  for (int i = warp; i < a.Lq; i += 4) {
    const long long qrow = (long long)b * a.Lq + i;
    {
      const uint32_t w = *reinterpret_cast<const uint32_t*>(a.q + qrow * a.ldq + h * kDecHd + lane * 2);
      q[lane * 2] = bf16_lo(w) * a.scale;
      q[lane * 2 + 1] = bf16_hi(w) * a.scale;
    }
    __syncwarp();
    const int nvis = a.lens ? min(i + 1, len) : a.Lk;       // visible keys are a prefix in both cases
    float mx = -INFINITY;
    for (int j = lane; j < a.Lk; j += 32) {
      float s = -INFINITY;
      if (j < nvis) {
        s = 0.f;
        const uint32_t* kr = reinterpret_cast<const uint32_t*>(sK + j * kKRow);
#pragma unroll 8
        for (int c = 0; c < 32; ++c) {
          const uint32_t w = kr[c];
          s = fmaf(q[2 * c], bf16_lo(w), s);
          s = fmaf(q[2 * c + 1], bf16_hi(w), s);
        }
      }
      p[j] = s;
      mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < a.Lk; j += 32) {
      const float e = j < nvis ? __expf(p[j] - mx) : 0.f;
      p[j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    __syncwarp();
    float o0 = 0.f, o1 = 0.f;
    for (int j = 0; j < nvis; ++j) {
      const float pj = p[j];
      const uint32_t w = reinterpret_cast<const uint32_t*>(sV + j * kKRow)[lane];
      o0 = fmaf(pj, bf16_lo(w), o0);
      o1 = fmaf(pj, bf16_hi(w), o1);
    }
    *reinterpret_cast<uint32_t*>(a.out + qrow * a.ldo + h * kDecHd + lane * 2) = pack_bf16(o0 * inv, o1 * inv);
    if (lane == 0) a.lse[((long long)b * a.H + h) * a.Lq + i] = mx + logf(sum);
    if (a.maps != nullptr) {
      const float k = inv / (float)a.H;
      for (int j = lane; j < a.Lk; j += 32) atomicAdd(a.maps + qrow * a.Lk + j, p[j] * k);
    }
    __syncwarp();
  }
}

struct DecAttnBwdArgs {
  const __nv_bfloat16 *q, *k, *v, *out, *dout;
  long long ldq, ldk, ldv, ldo, lddo;
  const float* lse;
  const long long* lens;
  __nv_bfloat16 *dq, *dk, *dv;
  long long lddq, lddk, lddv;
  int B, H, Lq, Lk;
  float scale;
};

// One CTA per (sample, head).  Phase A (warp per query row): p = exp(scale q.k - lse), D = dO.O, dP = dO.V^T, dS = scale p (dP - D),
// dq = dS.K; p and dS of all Lq rows stay in shared memory.  Phase B (thread per key x column pair): dV = P^T dO, dK = dS^T Q.
__global__ void __launch_bounds__(kDecThreads)
dec_attn_bwd_kernel(DecAttnBwdArgs a) {
  extern __shared__ uint8_t smem[];
  __nv_bfloat16* sK = reinterpret_cast<__nv_bfloat16*>(smem);
  const int LkP = (a.Lk + 3) & ~3;
  __nv_bfloat16* sV = sK + LkP * kKRow;
  float* sQ = reinterpret_cast<float*>(sV + LkP * kKRow);        // [Lq][64] unscaled q
  float* sdO = sQ + kDecMaxQ * kDecHd;                           // [Lq][64]
  float* sP = sdO + kDecMaxQ * kDecHd;                           // [Lq][LkP]
  float* sdS = sP + kDecMaxQ * LkP;                              // [Lq][LkP]
  const int b = blockIdx.x / a.H, h = blockIdx.x % a.H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  load_kv_tile(a.k + (long long)b * a.Lk * a.ldk + h * kDecHd, a.ldk, a.Lk, sK);
  load_kv_tile(a.v + (long long)b * a.Lk * a.ldv + h * kDecHd, a.ldv, a.Lk, sV);
  for (int i = threadIdx.x; i < a.Lq * 32; i += blockDim.x) {
    const int r = i >> 5, c2 = i & 31;
    const long long row = (long long)b * a.Lq + r;
    const uint32_t wq = *reinterpret_cast<const uint32_t*>(a.q + row * a.ldq + h * kDecHd + c2 * 2);
    const uint32_t wd = *reinterpret_cast<const uint32_t*>(a.dout + row * a.lddo + h * kDecHd + c2 * 2);
    sQ[r * kDecHd + c2 * 2] = bf16_lo(wq); sQ[r * kDecHd + c2 * 2 + 1] = bf16_hi(wq);
    sdO[r * kDecHd + c2 * 2] = bf16_lo(wd); sdO[r * kDecHd + c2 * 2 + 1] = bf16_hi(wd);
  }
  __syncthreads();
  const int len = a.lens ? (int)a.lens[b] : a.Lk;
  for (int i = warp; i < a.Lq; i += 4) {
    const long long row = (long long)b * a.Lq + i;
    const float* q = sQ + i * kDecHd;
    const float* dO = sdO + i * kDecHd;
    const int nvis = a.lens ? min(i + 1, len) : a.Lk;
    const float l = a.lse[((long long)b * a.H + h) * a.Lq + i];
    float dsum;
    {
      const uint32_t wo = *reinterpret_cast<const uint32_t*>(a.out + row * a.ldo + h * kDecHd + lane * 2);
      dsum = warp_sum(bf16_lo(wo) * dO[lane * 2] + bf16_hi(wo) * dO[lane * 2 + 1]);
    }
    float* p = sP + i * LkP;
    float* ds = sdS + i * LkP;
    for (int j = lane; j < a.Lk; j += 32) {
      float pj = 0.f, dsj = 0.f;
      if (j < nvis) {
        float s = 0.f, dp = 0.f;
        const uint32_t* kr = reinterpret_cast<const uint32_t*>(sK + j * kKRow);
        const uint32_t* vr = reinterpret_cast<const uint32_t*>(sV + j * kKRow);
#pragma unroll 8
        for (int c = 0; c < 32; ++c) {
          const uint32_t wk = kr[c], wv = vr[c];
          s = fmaf(q[2 * c], bf16_lo(wk), s);
          s = fmaf(q[2 * c + 1], bf16_hi(wk), s);
          dp = fmaf(dO[2 * c], bf16_lo(wv), dp);
          dp = fmaf(dO[2 * c + 1], bf16_hi(wv), dp);
        }
        pj = __expf(s * a.scale - l);
        dsj = a.scale * pj * (dp - dsum);
      }
      p[j] = pj;
      ds[j] = dsj;
    }
    __syncwarp();
    float g0 = 0.f, g1 = 0.f;
    for (int j = 0; j < nvis; ++j) {
      const float d = ds[j];
      const uint32_t w = reinterpret_cast<const uint32_t*>(sK + j * kKRow)[lane];
      g0 = fmaf(d, bf16_lo(w), g0);
      g1 = fmaf(d, bf16_hi(w), g1);
    }
    *reinterpret_cast<uint32_t*>(a.dq + row * a.lddq + h * kDecHd + lane * 2) = pack_bf16(g0, g1);
  }
  __syncthreads();
  for (int w = threadIdx.x; w < a.Lk * 32; w += blockDim.x) {
    const int j = w >> 5, c2 = w & 31;
    float v0 = 0.f, v1 = 0.f, k0 = 0.f, k1 = 0.f;
    for (int i = 0; i < a.Lq; ++i) {
      const float pj = sP[i * LkP + j], dsj = sdS[i * LkP + j];
      v0 = fmaf(pj, sdO[i * kDecHd + c2 * 2], v0);
      v1 = fmaf(pj, sdO[i * kDecHd + c2 * 2 + 1], v1);
      k0 = fmaf(dsj, sQ[i * kDecHd + c2 * 2], k0);
      k1 = fmaf(dsj, sQ[i * kDecHd + c2 * 2 + 1], k1);
    }
    const long long krow = (long long)b * a.Lk + j;
    *reinterpret_cast<uint32_t*>(a.dv + krow * a.lddv + h * kDecHd + c2 * 2) = pack_bf16(v0, v1);
    *reinterpret_cast<uint32_t*>(a.dk + krow * a.lddk + h * kDecHd + c2 * 2) = pack_bf16(k0, k1);
  }
}

// One warp per (sample, position): loss += mask * (lse - z[target]) / B; dlogits = mask * (softmax - onehot) / B; pred = argmax.
__global__ void __launch_bounds__(256)
seq_ce_kernel(const float* __restrict__ logits, long long ld, const long long* __restrict__ targets, const long long* __restrict__ lens,
              float* __restrict__ loss, float* __restrict__ dlogits, long long ldd, int* __restrict__ pred, int B, int T, int C) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= (long long)B * T) return;
  const int t = (int)(row % T);
  const long long b = row / T;
  const float* z = logits + row * ld;
  float mx = -INFINITY;
  int arg = 0;
  for (int c = lane; c < C; c += 32) {
    const float v = z[c];
    if (v > mx) { mx = v; arg = c; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, mx, o);
    const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
    if (om > mx || (om == mx && oa < arg)) { mx = om; arg = oa; }
  }
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += __expf(z[c] - mx);
  s = warp_sum(s);
  const bool valid = t < (int)lens[b];
  const int tgt = (int)targets[row];
  const float invB = 1.f / (float)B;
  if (lane == 0) {
    if (pred) pred[row] = arg;
    if (valid) atomicAdd(loss, (mx + logf(s) - z[tgt]) * invB);
  }
  if (dlogits != nullptr) {
    const float inv = 1.f / s;
    for (int c = lane; c < (int)ldd; c += 32) {
      float g = 0.f;
      if (valid && c < C) g = (__expf(z[c] - mx) * inv - (c == tgt ? 1.f : 0.f)) * invB;
      dlogits[row * ldd + c] = g;
    }
  }
}

static size_t dec_fwd_smem(int Lk) { const size_t k = (Lk + 3) & ~3; return 2 * k * kKRow * 2 + 4 * kDecHd * 4 + 4 * k * 4; }
static size_t dec_bwd_smem(int Lk) { const size_t k = (Lk + 3) & ~3; return 2 * k * kKRow * 2 + 2 * kDecMaxQ * kDecHd * 4 + 2 * kDecMaxQ * k * 4; }

}  // namespace dig

using namespace dig;

extern "C" int dig_embed_pos_fwd(const int64_t* targets, const float* emb, const float* pos, float* x, int32_t B, int32_t T, int32_t D,
                                 int32_t start_idx, void* stream) {
  DIG_REQUIRE(targets && emb && pos && x && B > 0 && T > 0 && D % 4 == 0, "dig_embed_pos_fwd: bad arguments");
  const long long n = (long long)B * T * (D / 4);
  embed_pos_fwd_kernel<<<(int)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const long long*)targets, emb, pos, x, B, T, D, start_idx);
  DIG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int dig_embed_bwd(const float* dx, const int64_t* targets, float* demb, int32_t B, int32_t T, int32_t D, int32_t start_idx,
                             void* stream) {
  DIG_REQUIRE(dx && targets && demb && B > 0 && T > 0 && D > 0, "dig_embed_bwd: bad arguments");
  const long long n = (long long)B * T * D;
  embed_bwd_kernel<<<(int)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(dx, (const long long*)targets, demb, B, T, D, start_idx);
  DIG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

static int check_dec_shapes(int32_t B, int32_t H, int32_t Lq, int32_t Lk, const char* what) {
  DIG_REQUIRE(B > 0 && H > 0 && Lq > 0 && Lq <= kDecMaxQ && Lk > 0 && Lk <= kDecMaxK, "%s: needs Lq <= %d and Lk <= %d (got %d, %d)", what,
              kDecMaxQ, kDecMaxK, Lq, Lk);
  return 0;
}

extern "C" int dig_dec_attention_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* out, int64_t ldo,
                                     float* lse, const int64_t* lens, float* maps, int32_t B, int32_t H, int32_t Lq, int32_t Lk, float scale,
                                     void* stream) {
  DIG_REQUIRE(q && k && v && out && lse, "dig_dec_attention_fwd: null pointer");
  if (int rc = check_dec_shapes(B, H, Lq, Lk, "dig_dec_attention_fwd")) return rc;
  DIG_REQUIRE(ldq % 2 == 0 && ldk % 2 == 0 && ldv % 2 == 0 && ldo % 2 == 0, "dig_dec_attention_fwd: leading dimensions must be even");
  DIG_REQUIRE(!lens || Lq == Lk, "dig_dec_attention_fwd: the causal/length mask is defined for self-attention (Lq == Lk)");
  DecAttnArgs a;
  a.q = (const __nv_bfloat16*)q; a.k = (const __nv_bfloat16*)k; a.v = (const __nv_bfloat16*)v; a.ldq = ldq; a.ldk = ldk; a.ldv = ldv;
  a.out = (__nv_bfloat16*)out; a.ldo = ldo; a.lse = lse; a.lens = (const long long*)lens; a.maps = maps;
  a.B = B; a.H = H; a.Lq = Lq; a.Lk = Lk; a.scale = scale;
  static bool attr = false;
  if (!attr) {
    DIG_CHECK_CUDA(cudaFuncSetAttribute(dec_attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dec_fwd_smem(kDecMaxK)));
    attr = true;
  }
  dec_attn_fwd_kernel<<<B * H, kDecThreads, dec_fwd_smem(Lk), (cudaStream_t)stream>>>(a);
  DIG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int dig_dec_attention_bwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, const void* out,
                                     int64_t ldo, const void* dout, int64_t lddo, const float* lse, const int64_t* lens, void* dq, int64_t lddq,
                                     void* dk, int64_t lddk, void* dv, int64_t lddv, int32_t B, int32_t H, int32_t Lq, int32_t Lk, float scale,
                                     void* stream) {
  DIG_REQUIRE(q && k && v && out && dout && lse && dq && dk && dv, "dig_dec_attention_bwd: null pointer");
  if (int rc = check_dec_shapes(B, H, Lq, Lk, "dig_dec_attention_bwd")) return rc;
  DIG_REQUIRE(!lens || Lq == Lk, "dig_dec_attention_bwd: the causal/length mask is defined for self-attention (Lq == Lk)");
  DecAttnBwdArgs a;
  a.q = (const __nv_bfloat16*)q; a.k = (const __nv_bfloat16*)k; a.v = (const __nv_bfloat16*)v; a.out = (const __nv_bfloat16*)out;
  a.dout = (const __nv_bfloat16*)dout; a.ldq = ldq; a.ldk = ldk; a.ldv = ldv; a.ldo = ldo; a.lddo = lddo; a.lse = lse;
  a.lens = (const long long*)lens; a.dq = (__nv_bfloat16*)dq; a.dk = (__nv_bfloat16*)dk; a.dv = (__nv_bfloat16*)dv;
  a.lddq = lddq; a.lddk = lddk; a.lddv = lddv; a.B = B; a.H = H; a.Lq = Lq; a.Lk = Lk; a.scale = scale;
  static bool attr = false;
  if (!attr) {
    DIG_CHECK_CUDA(cudaFuncSetAttribute(dec_attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dec_bwd_smem(kDecMaxK)));
    attr = true;
  }
  dec_attn_bwd_kernel<<<B * H, kDecThreads, dec_bwd_smem(Lk), (cudaStream_t)stream>>>(a);
  DIG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int dig_seq_cross_entropy(const float* logits, int64_t ld, const int64_t* targets, const int64_t* lens, float* loss, float* dlogits,
                                     int64_t ldd, int32_t* pred, int32_t B, int32_t T, int32_t C, void* stream) {
  DIG_REQUIRE(logits && targets && lens && loss && B > 0 && T > 0 && C > 0 && ld >= C && (!dlogits || ldd >= C),
              "dig_seq_cross_entropy: bad arguments");
  const long long rows = (long long)B * T;
  seq_ce_kernel<<<(int)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(logits, ld, (const long long*)targets, (const long long*)lens, loss,
                                                                       dlogits, ldd, pred, B, T, C);
  DIG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

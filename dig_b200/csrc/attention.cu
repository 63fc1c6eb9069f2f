// dig_b200 -- fused multi-head self-attention over the 256 patch tokens of a 32x128 crop (head_dim 64).
//
// Replaces modeling_finetune.py:97-118 of the reference (q*scale, q@k^T, softmax, @v -- two cuBLAS bmm and
// an ATen softmax that materialise a [S,h,256,256] score tensor in HBM) and its autograd backward.  Scores
// never leave the SM: S = Q.K^T is accumulated in TMEM by tcgen05.mma, each thread owns one query row for
// the softmax (a whole row of 256 keys is resident, so no online rescaling), P goes back to the tensor core
// as a bf16 operand (from TMEM in the forward, from swizzled shared memory in the backward).
//
// Layout: qkv is the QKV projection output [S*256, 3*d] bf16 (q | k | v, each head a 64-column slice, exactly
// the memory order F:93-95 reshapes); TMA pulls the per-head tiles straight out of it, so there is no
// permute/contiguous pass.  The context is written as [S*256, d] bf16, ready for the output projection.
#include <stdlib.h>

#include "common.cuh"
#include "../../include/dig_b200.h"

namespace dig {

static constexpr int kTok = 256;  // tokens per sequence (8 x 32 patch grid)
static constexpr int kHd = 64;    // head dim
static constexpr float kLog2e = 1.4426950408889634f;

// ------------------------------------------------------------------------------------------------
// forward: one CTA per (sequence, head, 128-query tile)
// ------------------------------------------------------------------------------------------------
template <bool P_TMEM>
__global__ void __launch_bounds__(160, P_TMEM ? 2 : 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tm_qkv, __nv_bfloat16* __restrict__ out, float* __restrict__ lse, int heads,
                float scale) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                 // 128 x 64 bf16, K-major SW128               16 KB
  uint8_t* sK = smem + 16384;         // 256 x 64 bf16, K-major SW128 (B of Q.K^T)   32 KB
  uint8_t* sV = smem + 49152;         // 256 x 64 bf16, rows = keys -> MN-major B of P.V  32 KB
  uint8_t* sP = smem + 81920;         // !P_TMEM: 4 x (128 x 64) bf16 K-major SW128   64 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 81920 + (P_TMEM ? 0 : 65536));
  uint64_t* bar_qk = bars + 0;
  uint64_t* bar_v = bars + 1;
  uint64_t* bar_s = bars + 2;
  uint64_t* bar_p = bars + 3;
  uint64_t* bar_o = bars + 4;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 5);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x & 1;
  const int head = (blockIdx.x >> 1) % heads;
  const int seq = (blockIdx.x >> 1) / heads;
  const int d = heads * kHd;
  const int row0 = seq * kTok;

  if (warp == 4) {
    if (lane == 0) {
      tma_prefetch_desc(&tm_qkv);
      mbar_init(bar_qk, 1);
      mbar_init(bar_v, 1);
      mbar_init(bar_s, 1);
      mbar_init(bar_p, 128);
      mbar_init(bar_o, 1);
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc(tmem_holder, 256);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_holder;
  constexpr uint32_t kColO = 192;

  if (warp == 4) {
    if (lane == 0) {
      mbar_expect_tx(bar_qk, 16384 + 32768);
      tma_load_2d(sQ, &tm_qkv, bar_qk, head * kHd, row0 + qt * 128);
      tma_load_2d(sK, &tm_qkv, bar_qk, d + head * kHd, row0);
      tma_load_2d(sK + 16384, &tm_qkv, bar_qk, d + head * kHd, row0 + 128);
      mbar_expect_tx(bar_v, 32768);
      tma_load_2d(sV, &tm_qkv, bar_v, 2 * d + head * kHd, row0);
      tma_load_2d(sV + 16384, &tm_qkv, bar_v, 2 * d + head * kHd, row0 + 128);

      mbar_wait(bar_qk, 0);
      tc_fence_after();
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 256, false, false);
#pragma unroll
      for (int k = 0; k < kHd / 16; ++k)
        tc_mma_ss(tmem, make_sdesc_sw128(smem_u32(sQ) + k * 32, 16, 1024), make_sdesc_sw128(smem_u32(sK) + k * 32, 16, 1024), idesc_s,
                  k > 0);
      tc_commit(bar_s);

      mbar_wait(bar_p, 0);
      mbar_wait(bar_v, 0);
      tc_fence_after();
      constexpr uint32_t idesc_o = make_idesc_bf16(128, 64, false, true);
#pragma unroll
      for (int k = 0; k < kTok / 16; ++k) {
        const uint64_t dv = make_sdesc_sw128(smem_u32(sV) + k * 2048, 8192, 1024);
        if (P_TMEM) tc_mma_ts(tmem + kColO, tmem + k * 8, dv, idesc_o, k > 0);
        else
          tc_mma_ss(tmem + kColO, make_sdesc_sw128(smem_u32(sP) + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024), dv, idesc_o, k > 0);
      }
      tc_commit(bar_o);
    }
  } else {
    // softmax warps: thread t owns query row t of this tile == TMEM lane t
    const int t = threadIdx.x;
    const uint32_t tl = tmem + ((uint32_t)(warp * 32) << 16);
    const float sl2 = scale * kLog2e;
    mbar_wait(bar_s, 0);
    tc_fence_after();
    float mx = -INFINITY;
#pragma unroll 1
    for (int c = 0; c < kTok; c += 32) {
      uint32_t v[32];
      tmem_ld32(tl + c, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
    }
    const float mb = mx * sl2;
    float sum = 0.f;
#pragma unroll 1
    for (int c = 0; c < kTok; c += 32) {
      uint32_t v[32];
      tmem_ld32(tl + c, v);
      tmem_ld_wait();
      uint32_t pk[16];
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        const float p0 = exp2f(fmaf(__uint_as_float(v[j]), sl2, -mb));
        const float p1 = exp2f(fmaf(__uint_as_float(v[j + 1]), sl2, -mb));
        sum += p0 + p1;
        pk[j >> 1] = pack_bf16(p0, p1);
      }
      if (P_TMEM) {
        tmem_st16(tl + (c >> 1), pk);
      } else {
        const uint32_t base = smem_u32(sP) + (c >> 6) * 16384;
        const uint32_t ch0 = (c & 63) >> 3;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj)
          sts_u4(base + sw128_offset(t, ch0 + jj), make_uint4(pk[4 * jj], pk[4 * jj + 1], pk[4 * jj + 2], pk[4 * jj + 3]));
      }
    }
    if (P_TMEM) tmem_st_wait();
    else fence_proxy_async_smem();
    tc_fence_before();
    mbar_arrive(bar_p);

    mbar_wait(bar_o, 0);
    tc_fence_after();
    const float inv = 1.0f / sum;
    const long long grow = (long long)row0 + qt * 128 + t;
    __nv_bfloat16* o = out + grow * d + head * kHd;
#pragma unroll
    for (int c = 0; c < kHd; c += 32) {
      uint32_t v[32];
      tmem_ld32(tl + kColO + c, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        uint4 p;
        p.x = pack_bf16(__uint_as_float(v[j + 0]) * inv, __uint_as_float(v[j + 1]) * inv);
        p.y = pack_bf16(__uint_as_float(v[j + 2]) * inv, __uint_as_float(v[j + 3]) * inv);
        p.z = pack_bf16(__uint_as_float(v[j + 4]) * inv, __uint_as_float(v[j + 5]) * inv);
        p.w = pack_bf16(__uint_as_float(v[j + 6]) * inv, __uint_as_float(v[j + 7]) * inv);
        *reinterpret_cast<uint4*>(o + c + j) = p;
      }
    }
    if (lse != nullptr) lse[((long long)seq * heads + head) * kTok + qt * 128 + t] = mx * scale + logf(sum);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
}

// ------------------------------------------------------------------------------------------------
// forward, persistent: one CTA per SM walks (sequence, head) items; two softmax warpgroups ("slots") own the two 128-query tiles
// of an item and share its K/V.  Warp 0 = TMA producer (double-buffered Q/K/V, prefetches the next item while this one computes),
// warp 1 = tcgen05.mma issuer and TMEM owner (2 x 256 columns: S -> P in place, O in the last 64 columns of the slot),
// warps 2-5 = slot 0, warps 6-9 = slot 1 (a warp may only touch TMEM lanes 32*(warp%4)..+31; each slot covers all four quarters).
// Per slot and item: S = Q K^T (MMA) -> row max / exp2 / sum, bf16 P back into TMEM (threads) -> O = P V (MMA, A from TMEM) ->
// O / sum -> HBM (threads).  The tile-to-tile prologue, TMA latency and MMA latency of the one-shot kernel above are hidden behind the
// other slot's softmax, which is the MUFU-bound critical resource.
// ------------------------------------------------------------------------------------------------
#ifndef DIG_ATTN_STAGGER
#define DIG_ATTN_STAGGER 2200
#endif
static constexpr long long kAttnStagger = DIG_ATTN_STAGGER;  // SM clocks between the first S products of slot 0 and slot 1
static constexpr int kFwdPThreads = 352;  // warp 0: TMA, warps 1-2: MMA issue of slot 0 / 1, warps 3-10: softmax (two slots x 4 TMEM lane quarters)

// Bring-up instrumentation: when a buffer is registered (dig_attention_debug_buffer), CTA 0 records SM clock stamps of its MMA warps
// and of one softmax thread per slot for the first 12 items (scripts/attn_timeline.py prints them).  Null in normal operation.
__device__ long long* g_attn_dbg = nullptr;
#define DIG_STAMP(role, n, k)                                                                         \
  do {                                                                                                \
    if (dbg != nullptr && (n) < 12) dbg[((role) * 12 + (n)) * 8 + (k)] = clock64();                     \
  } while (0)
static constexpr int kFwdPBuf = 98304;  // Q0 16K | Q1 16K | K 32K | V 32K

__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// 2^x for x <= 0 on the FMA / ALU pipes (no MUFU): round-to-nearest split x = n + f with the 1.5 * 2^23 trick, a cubic minimax fit of
// 2^f on [-0.5, 0.5] (max relative error 7.5e-5; the result is rounded to bf16, 2e-3), then n is added into the exponent field.
// The softmax is MUFU-bound (16 ex2 per clock and SM against 64K scores per item); every DIG_ATTN_POLY-th score takes this path.
#ifndef DIG_ATTN_POLY
#define DIG_ATTN_POLY 0
#endif
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -125.0f);
  const float fi = x + 12582912.0f;
  const float f = x - (fi - 12582912.0f);
  float p = fmaf(f, 0.055170830339193344f, 0.24260906875133514f);
  p = fmaf(p, f, 0.693260908126831f);
  p = fmaf(p, f, 0.9999281764030457f);
  return __uint_as_float(__float_as_uint(p) + (__float_as_uint(fi) << 23));
}
template <int E>
__device__ __forceinline__ float ex2_sel(float x) {   // E = position of the score within its 32-column chunk (compile-time)
  if constexpr (DIG_ATTN_POLY > 0 && (E % (DIG_ATTN_POLY > 0 ? DIG_ATTN_POLY : 1)) == (DIG_ATTN_POLY > 0 ? DIG_ATTN_POLY - 1 : 0)) return ex2_poly(x);
  else return ex2_approx(x);
}

// exp2 of one 32-column chunk of scores: packed bf16 P for the PV product, fp32 row sums in two chains
template <int J>
__device__ __forceinline__ void exp_pair(const uint32_t (&v)[32], uint32_t (&pk)[16], float& sum0, float& sum1, float sl2, float mb) {
  const float p0 = ex2_sel<J>(fmaf(__uint_as_float(v[J]), sl2, -mb));
  const float p1 = ex2_sel<J + 1>(fmaf(__uint_as_float(v[J + 1]), sl2, -mb));
  sum0 += p0;
  sum1 += p1;
  pk[J >> 1] = pack_bf16(p0, p1);
  if constexpr (J + 2 < 32) exp_pair<J + 2>(v, pk, sum0, sum1, sl2, mb);
}
__device__ __forceinline__ void exp_chunk(const uint32_t (&v)[32], uint32_t (&pk)[16], float& sum0, float& sum1, float sl2, float mb) {
  exp_pair<0>(v, pk, sum0, sum1, sl2, mb);
}

// Slot layout in TMEM (256 columns per slot): S fp32 [128 q x 256 keys] fills all of it; the softmax overwrites it in place with
//   cols   0- 63  P (bf16 pairs) of keys   0-127          cols  64-127  O accumulator (dead S columns once keys 0-127 are consumed)
//   cols 128-191  P of keys 128-255                        cols 192-255  dead
// so the PV product of the first key half is issued while the second half is still being exponentiated.
__global__ void __launch_bounds__(kFwdPThreads, 1)
attn_fwd_persist_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_out, float* __restrict__ lse,
                        int heads, float scale, int num_items) {
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * kFwdPBuf + 2 * 16384);   // after the two 16 KB output staging tiles
  uint64_t* qk_full = bars + 0;    // [2 buffers] TMA -> MMA warps
  uint64_t* v_full = bars + 2;     // [2 buffers] TMA -> MMA warps
  uint64_t* v_empty = bars + 4;    // [2 buffers] MMA -> TMA (2 arrivals): both slots' PV products have read V
  uint64_t* s_full = bars + 6;     // [2 slots]   MMA -> softmax
  uint64_t* p_half = bars + 8;     // [2 slots][2 key halves] softmax -> MMA (128 arrivals)
  uint64_t* o_full = bars + 12;    // [2 slots]   MMA -> softmax
  uint64_t* s_free = bars + 14;    // [2 slots]   softmax -> MMA (128 arrivals): O is in registers, the slot's TMEM may be overwritten
  uint64_t* k_empty = bars + 16;   // [2 buffers] MMA -> TMA (2 arrivals): both S products have read K (a softmax phase before V)
  uint64_t* q_empty = bars + 18;   // [2 slots][2 buffers] MMA -> TMA: the slot's S product has read its Q tile
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 22);
  volatile long long* t_first = reinterpret_cast<volatile long long*>(bars + 23);   // clock of slot 0's first S issue (0 = not yet)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int d = heads * kHd;
  long long* const dbg = (blockIdx.x == 0 && (warp == 1 || warp == 2 || (lane == 0 && ((warp - 3) & 3) == 0))) ? g_attn_dbg : nullptr;
  if (dbg != nullptr && warp == 1 && lane == 0) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    dbg[11 * 8 + 0] = clock64();
    dbg[11 * 8 + 2] = (long long)gt;
  }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_qkv);
    tma_prefetch_desc(&tm_out);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&qk_full[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 2);
      mbar_init(&k_empty[i], 2);
      mbar_init(&q_empty[i], 1);
      mbar_init(&q_empty[2 + i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_half[2 * i], 128);
      mbar_init(&p_half[2 * i + 1], 128);
      mbar_init(&o_full[i], 1);
      mbar_init(&s_free[i], 128);
    }
    *t_first = 0;
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_holder, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_holder, 0);   // warp-uniform for the compiler too
  constexpr uint32_t kColO = 64, kColP1 = 128;
  pdl_wait();   // prologue done: the qkv tensor written by the preceding GEMM is read from here on

  if (warp == 0) {
    if (lane == 0) {
      int n = 0;
      for (int w = blockIdx.x; w < num_items; w += gridDim.x, ++n) {
        const int b = n & 1;
        const uint32_t u = (uint32_t)(n >> 1) & 1u;
        const int head = w % heads, row0 = (w / heads) * kTok;
        uint8_t* base = smem + b * kFwdPBuf;
        // Q tiles and K are handed back right after the S products (a whole softmax phase before V), so most of the next-but-one
        // item's bytes are requested early: the 2-deep buffer ring then covers the ~1.8 us it takes one SM to pull 96 KB from HBM.
        mbar_wait(&q_empty[b], u ^ 1u);          // slot 0's S of the item two back is done: that item's qk_full phase has completed
        mbar_expect_tx(&qk_full[b], 65536);
        tma_load_2d(base, &tm_qkv, &qk_full[b], head * kHd, row0);
        mbar_wait(&q_empty[2 + b], u ^ 1u);
        tma_load_2d(base + 16384, &tm_qkv, &qk_full[b], head * kHd, row0 + 128);
        mbar_wait(&k_empty[b], u ^ 1u);
        tma_load_2d(base + 32768, &tm_qkv, &qk_full[b], d + head * kHd, row0);
        tma_load_2d(base + 49152, &tm_qkv, &qk_full[b], d + head * kHd, row0 + 128);
        mbar_wait(&v_empty[b], u ^ 1u);
        mbar_expect_tx(&v_full[b], 32768);
        tma_load_2d(base + 65536, &tm_qkv, &v_full[b], 2 * d + head * kHd, row0);
        tma_load_2d(base + 81920, &tm_qkv, &v_full[b], 2 * d + head * kHd, row0 + 128);
      }
    }
  } else if (warp <= 2) {
    // One issuing warp per slot (blocking waits): each slot is its own S -> (softmax) -> PV pipeline and neither waits behind the
    // other's 16-instruction PV issue (with one polling warp for both, clock stamps showed 600-900 clocks between a barrier completing
    // and the MMAs it releases being issued).  Slot 1 is started half a period late, so one slot's MMAs, O drain and barrier round
    // trips hide behind the other slot's exp2 work instead of both slots hitting the MUFU pipe -- and then both leaving it -- together.
    // The whole warp walks the loop and ONE ELECTED lane issues (operands stay in uniform registers).
    const int s = warp - 1;
    constexpr uint32_t idesc_s = make_idesc_bf16(128, 256, false, false);
    constexpr uint32_t idesc_o = make_idesc_bf16(128, 64, false, true);
    const int my_items = (num_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const uint32_t smem_s = smem_u32(smem);
    const uint32_t ts = tmem + (uint32_t)s * 256u;
    for (int n = 0; n < my_items; ++n) {
      const int b = n & 1;
      const uint32_t u = (uint32_t)(n >> 1) & 1u, np = (uint32_t)n & 1u;
      const uint32_t base = smem_s + b * kFwdPBuf;
      mbar_wait(&qk_full[b], u);
      mbar_wait(&s_free[s], np ^ 1u);
      if (s == 1 && n == 0 && kAttnStagger > 0) {
        long long t0;
        while ((t0 = *t_first) == 0) {
        }
        while (clock64() - t0 < kAttnStagger) {
        }
      }
      tc_fence_after();
      if (elect_one()) {
        DIG_STAMP(0, n, 1 + s);
#pragma unroll
        for (int k = 0; k < kHd / 16; ++k)
          tc_mma_ss(ts, make_sdesc_sw128(base + s * 16384 + k * 32, 16, 1024), make_sdesc_sw128(base + 32768 + k * 32, 16, 1024), idesc_s,
                    k > 0);
        tc_commit(&s_full[s]);
        tc_commit(&q_empty[2 * s + b]);
        tc_commit(&k_empty[b]);
        if (s == 0 && n == 0) *t_first = clock64() | 1;
      }
      __syncwarp();
      mbar_wait(&v_full[b], u);
      mbar_wait(&p_half[2 * s], np);
      tc_fence_after();
      if (elect_one()) {
        DIG_STAMP(0, n, 4 + s);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          tc_mma_ts(ts + kColO, ts + k * 8, make_sdesc_sw128(base + 65536 + k * 2048, 8192, 1024), idesc_o, k > 0);
      }
      __syncwarp();
      mbar_wait(&p_half[2 * s + 1], np);
      tc_fence_after();
      if (elect_one()) {
        DIG_STAMP(0, n, 6 + s);
#pragma unroll
        for (int k = 8; k < 16; ++k)
          tc_mma_ts(ts + kColO, ts + kColP1 + (k - 8) * 8, make_sdesc_sw128(base + 65536 + k * 2048, 8192, 1024), idesc_o, 1);
        tc_commit(&o_full[s]);
        tc_commit(&v_empty[b]);   // second arrival (either order): both slots are through with this item's V
      }
      __syncwarp();
    }
  } else {
    const int s = (warp - 3) >> 2;
    const int quarter = warp & 3;
    const int t = quarter * 32 + lane;  // query row within the slot's 128-row tile == TMEM lane
    const uint32_t tl = tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)s * 256u;
    const float sl2 = scale * kLog2e;
    const uint32_t stage_s = smem_u32(smem + 2 * kFwdPBuf + s * 16384);
    int n = 0;
    for (int w = blockIdx.x; w < num_items; w += gridDim.x, ++n) {
      const uint32_t np = (uint32_t)n & 1u;
      const int head = w % heads, seq = w / heads;
      DIG_STAMP(1 + s, n, 0);
      mbar_wait(&s_full[s], np);
      tc_fence_after();
      DIG_STAMP(1 + s, n, 1);
      // both passes keep one TMEM load in flight behind the chunk being processed (two register buffers); the row maximum runs in
      // four independent chains (one chain of 128 dependent FMNMX3 was ~900 clocks per row)
      float mx;
      {
        float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
        uint32_t va[32], vb[32];
        tmem_ld32(tl, va);
#pragma unroll 1
        for (int c = 0; c < kTok; c += 64) {
          tmem_ld_wait();
          tmem_ld32(tl + c + 32, vb);
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            m0 = fmaxf(m0, fmaxf(__uint_as_float(va[j + 0]), __uint_as_float(va[j + 1])));
            m1 = fmaxf(m1, fmaxf(__uint_as_float(va[j + 2]), __uint_as_float(va[j + 3])));
            m2 = fmaxf(m2, fmaxf(__uint_as_float(va[j + 4]), __uint_as_float(va[j + 5])));
            m3 = fmaxf(m3, fmaxf(__uint_as_float(va[j + 6]), __uint_as_float(va[j + 7])));
          }
          tmem_ld_wait();
          if (c + 64 < kTok) tmem_ld32(tl + c + 64, va);
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            m0 = fmaxf(m0, fmaxf(__uint_as_float(vb[j + 0]), __uint_as_float(vb[j + 1])));
            m1 = fmaxf(m1, fmaxf(__uint_as_float(vb[j + 2]), __uint_as_float(vb[j + 3])));
            m2 = fmaxf(m2, fmaxf(__uint_as_float(vb[j + 4]), __uint_as_float(vb[j + 5])));
            m3 = fmaxf(m3, fmaxf(__uint_as_float(vb[j + 6]), __uint_as_float(vb[j + 7])));
          }
        }
        mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
      }
      const float mb = mx * sl2;
      DIG_STAMP(1 + s, n, 2);
      float sum0 = 0.f, sum1 = 0.f;
      {
        uint32_t va[32], vb[32];
        tmem_ld32(tl, va);
#pragma unroll 1
        for (int c = 0; c < kTok; c += 64) {
          const uint32_t pc = tl + (uint32_t)(c >> 1) + (c >= 128 ? 64u : 0u);   // P columns of keys c..c+63
          uint32_t pk[16];
          tmem_ld_wait();
          tmem_ld32(tl + c + 32, vb);
          exp_chunk(va, pk, sum0, sum1, sl2, mb);
          tmem_ld_wait();   // vb has landed: the S columns the next two P stores overwrite (<= c+31) have been read
          if (c + 64 < kTok) tmem_ld32(tl + c + 64, va);
          tmem_st16(pc, pk);
          exp_chunk(vb, pk, sum0, sum1, sl2, mb);
          tmem_st16(pc + 16, pk);
          if (c == 64 || c == 192) {   // a key half is complete (for c == 64 the load in flight reads columns 128-159: not touched by PV)
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(&p_half[2 * s + (c >> 7)]);
            DIG_STAMP(1 + s, n, 3 + (c >> 7));
          }
        }
      }

      const float sum = sum0 + sum1;
      const float inv = 1.0f / sum;
      if (lse != nullptr) lse[((long long)seq * heads + head) * kTok + s * 128 + t] = mx * scale + logf(sum);
      mbar_wait(&o_full[s], np);
      tc_fence_after();
      DIG_STAMP(1 + s, n, 5);
      // O leaves TMEM in one go and the slot is handed back before the scaling / packing / staging work, so the next S product is
      // issued ~300 clocks earlier.  O / sum goes out through a 128 x 64 bf16 swizzled staging tile and ONE TMA store per slot and
      // item (a row-per-thread STG.128 touches 32 different lines per warp instruction).
      uint32_t o0[32], o1[32];
      tmem_ld32(tl + kColO, o0);
      tmem_ld32(tl + kColO + 32, o1);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&s_free[s]);
      DIG_STAMP(1 + s, n, 6);
      if (t == 0) tma_store_wait_read_all();                 // last item's store has finished reading the staging tile
      asm volatile("bar.sync %0, 128;" ::"r"(1 + s) : "memory");
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        uint4 p;
        p.x = pack_bf16(__uint_as_float(o0[j + 0]) * inv, __uint_as_float(o0[j + 1]) * inv);
        p.y = pack_bf16(__uint_as_float(o0[j + 2]) * inv, __uint_as_float(o0[j + 3]) * inv);
        p.z = pack_bf16(__uint_as_float(o0[j + 4]) * inv, __uint_as_float(o0[j + 5]) * inv);
        p.w = pack_bf16(__uint_as_float(o0[j + 6]) * inv, __uint_as_float(o0[j + 7]) * inv);
        sts_u4(stage_s + sw128_offset((uint32_t)t, (uint32_t)(j >> 3)), p);
      }
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        uint4 p;
        p.x = pack_bf16(__uint_as_float(o1[j + 0]) * inv, __uint_as_float(o1[j + 1]) * inv);
        p.y = pack_bf16(__uint_as_float(o1[j + 2]) * inv, __uint_as_float(o1[j + 3]) * inv);
        p.z = pack_bf16(__uint_as_float(o1[j + 4]) * inv, __uint_as_float(o1[j + 5]) * inv);
        p.w = pack_bf16(__uint_as_float(o1[j + 6]) * inv, __uint_as_float(o1[j + 7]) * inv);
        sts_u4(stage_s + sw128_offset((uint32_t)t, (uint32_t)((32 + j) >> 3)), p);
      }
      fence_proxy_async_smem();
      asm volatile("bar.sync %0, 128;" ::"r"(1 + s) : "memory");
      if (t == 0) {
        tma_store_2d(&tm_out, stage_s, head * kHd, seq * kTok + s * 128);
        tma_store_commit();
      }
      DIG_STAMP(1 + s, n, 7);
    }
    if (t == 0) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (dbg != nullptr && warp == 1 && lane == 0) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    dbg[11 * 8 + 1] = clock64();
    dbg[11 * 8 + 3] = (long long)gt;
  }
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// backward: one CTA per (sequence, head); loops key half j x query half i
//   S = Q_i K_j^T, dP = dO_i V_j^T  (TMEM)  ->  P = exp(scale*S - lse), dS = scale * P o (dP - D)  (threads)
//   dV_j += P^T dO_i, dK_j += dS^T Q_i, dQ_i += dS K_j   (TMEM accumulators)
// ------------------------------------------------------------------------------------------------
static constexpr int kBwdThreads = 288;  // warps 0-3: key columns 0-63 of a block, warps 4-7: columns 64-127 (same TMEM lane quarters), warp 8: TMA + MMA

__global__ void __launch_bounds__(kBwdThreads, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do, const __grid_constant__ CUtensorMap tm_o,
                const __grid_constant__ CUtensorMap tm_dqkv, const float* __restrict__ lse, int heads, float scale) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;            // 256 x 64, rows = tokens (two 128-row tiles of 16 KB)
  uint8_t* sK = smem + 32768;
  uint8_t* sV = smem + 65536;
  uint8_t* sdO = smem + 98304;
  uint8_t* sP = smem + 131072;   // [128 queries x 128 keys] as two 64-key column chunks of 16 KB; first the O tiles (for D), last dQ staging
  uint8_t* sdS = smem + 163840;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 196608);
  uint64_t* bar_ld = bars + 0;
  uint64_t* bar_sdp = bars + 1;
  uint64_t* bar_pds = bars + 2;
  uint64_t* bar_mma = bars + 3;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.x % heads;
  const int seq = blockIdx.x / heads;
  const int d = heads * kHd;
  const int row0 = seq * kTok;
  long long* const dbg = (blockIdx.x == 0 && lane == 0 && (warp == 8 || warp == 0)) ? g_attn_dbg : nullptr;
  DIG_STAMP(warp == 8 ? 0 : 1, 0, 0);

  if (warp == 8) {
    if (lane == 0) {
      tma_prefetch_desc(&tm_qkv);
      tma_prefetch_desc(&tm_do);
      tma_prefetch_desc(&tm_o);
      tma_prefetch_desc(&tm_dqkv);
      mbar_init(bar_ld, 1);
      mbar_init(bar_sdp, 1);
      mbar_init(bar_pds, 256);
      mbar_init(bar_mma, 1);
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc(tmem_holder, 512);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_holder;
  constexpr uint32_t cS = 0, cdP = 128, cdV = 256, cdK = 320, cdQ = 384;

  if (warp == 8) {
    if (lane == 0) {
      mbar_expect_tx(bar_ld, 5 * 32768);
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        tma_load_2d(sdO + r * 16384, &tm_do, bar_ld, head * kHd, row0 + r * 128);
        tma_load_2d(sP + r * 16384, &tm_o, bar_ld, head * kHd, row0 + r * 128);      // O rides in the P buffer until D is computed
        tma_load_2d(sQ + r * 16384, &tm_qkv, bar_ld, head * kHd, row0 + r * 128);
        tma_load_2d(sK + r * 16384, &tm_qkv, bar_ld, d + head * kHd, row0 + r * 128);
        tma_load_2d(sV + r * 16384, &tm_qkv, bar_ld, 2 * d + head * kHd, row0 + r * 128);
      }
      DIG_STAMP(0, 0, 1);
      mbar_wait(bar_ld, 0);
      tc_fence_after();
      DIG_STAMP(0, 0, 2);
      constexpr uint32_t id_s = make_idesc_bf16(128, 128, false, false);
      constexpr uint32_t id_tt = make_idesc_bf16(128, 64, true, true);   // dV, dK: A = P^T / dS^T (MN-major), B MN-major
      constexpr uint32_t id_q = make_idesc_bf16(128, 64, false, true);   // dQ: A = dS (K-major), B = K (MN-major)
      const uint32_t aQ = smem_u32(sQ), aK = smem_u32(sK), aV = smem_u32(sV), adO = smem_u32(sdO), aP = smem_u32(sP),
                     adS = smem_u32(sdS);
      // Issue order: S,dP(0) | for each block: wait P,dS(it) -> S,dP(it+1) -> dV,dK,dQ(it).  The score products of the next block go
      // ahead of this block's gradient products, so the compute warps work on block it+1 (TMEM -> registers) while the tensor pipe
      // is busy with dV,dK,dQ(it); they only wait for those (bar_mma) before overwriting the P / dS buffers.
      auto issue_sdp = [&](int it) {
        const int j = it >> 1, i = it & 1;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          tc_mma_ss(tmem + cS, make_sdesc_sw128(aQ + i * 16384 + k * 32, 16, 1024), make_sdesc_sw128(aK + j * 16384 + k * 32, 16, 1024),
                    id_s, k > 0);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          tc_mma_ss(tmem + cdP, make_sdesc_sw128(adO + i * 16384 + k * 32, 16, 1024), make_sdesc_sw128(aV + j * 16384 + k * 32, 16, 1024),
                    id_s, k > 0);
        tc_commit(bar_sdp);
      };
      issue_sdp(0);
      DIG_STAMP(0, 1, 0);
      for (int it = 0; it < 4; ++it) {
        const int j = it >> 1, i = it & 1;
        mbar_wait(bar_pds, it & 1);
        tc_fence_after();
        DIG_STAMP(0, 1 + it, 1);
        if (it < 3) issue_sdp(it + 1);
#pragma unroll
        for (int k = 0; k < 8; ++k)  // dV_j += P^T dO_i   (reduction over 128 queries)
          tc_mma_ss(tmem + cdV, make_sdesc_sw128(aP + k * 2048, 16384, 1024), make_sdesc_sw128(adO + i * 16384 + k * 2048, 8192, 1024),
                    id_tt, (i > 0 || k > 0));
#pragma unroll
        for (int k = 0; k < 8; ++k)  // dK_j += dS^T Q_i
          tc_mma_ss(tmem + cdK, make_sdesc_sw128(adS + k * 2048, 16384, 1024), make_sdesc_sw128(aQ + i * 16384 + k * 2048, 8192, 1024),
                    id_tt, (i > 0 || k > 0));
#pragma unroll
        for (int k = 0; k < 8; ++k)  // dQ_i += dS K_j     (reduction over 128 keys)
          tc_mma_ss(tmem + cdQ + i * 64, make_sdesc_sw128(adS + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024),
                    make_sdesc_sw128(aK + j * 16384 + k * 2048, 8192, 1024), id_q, (j > 0 || k > 0));
        tc_commit(bar_mma);
        DIG_STAMP(0, 1 + it, 2);
      }
    }
  } else {
    const int t = (warp & 3) * 32 + lane;  // row within a 128-token tile == TMEM lane
    const int hh = warp >> 2;              // which 64 key columns of a 128-key block (and which 32 head-dim columns of an output tile)
    const uint32_t tl = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const float sl2 = scale * kLog2e;
    const uint32_t sP_s = smem_u32(sP), sdS_s = smem_u32(sdS), sdO_s = smem_u32(sdO);
    // this thread's 32 of the 64 columns of a 128-row fp32 accumulator -> bf16, times `mul`, into a swizzled staging tile (TMA store)
    auto stage_tile = [&](uint32_t tcol, uint32_t dst_s, float mul) {
      uint32_t v[32];
      tmem_ld32(tl + tcol + hh * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 32; q += 8) {
        uint4 p;
        p.x = pack_bf16(__uint_as_float(v[q + 0]) * mul, __uint_as_float(v[q + 1]) * mul);
        p.y = pack_bf16(__uint_as_float(v[q + 2]) * mul, __uint_as_float(v[q + 3]) * mul);
        p.z = pack_bf16(__uint_as_float(v[q + 4]) * mul, __uint_as_float(v[q + 5]) * mul);
        p.w = pack_bf16(__uint_as_float(v[q + 6]) * mul, __uint_as_float(v[q + 7]) * mul);
        sts_u4(dst_s + sw128_offset((uint32_t)t, (uint32_t)((hh * 32 + q) >> 3)), p);
      }
    };
    // D_i = rowsum(dO o O) from the TMA-staged tiles (a row-per-thread global read costs 32 LSU wavefronts per instruction), lse_i
    float Dr[2], Lr[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) Lr[i] = lse[((long long)seq * heads + head) * kTok + i * 128 + t] * kLog2e;
    mbar_wait(bar_ld, 0);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      float acc = 0.f;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint32_t off = (uint32_t)i * 16384u + sw128_offset((uint32_t)t, (uint32_t)c);
        const uint4 a = lds_u4(sP_s + off), b = lds_u4(sdO_s + off);
        acc += bf16_lo(a.x) * bf16_lo(b.x) + bf16_hi(a.x) * bf16_hi(b.x) + bf16_lo(a.y) * bf16_lo(b.y) + bf16_hi(a.y) * bf16_hi(b.y) +
               bf16_lo(a.z) * bf16_lo(b.z) + bf16_hi(a.z) * bf16_hi(b.z) + bf16_lo(a.w) * bf16_lo(b.w) + bf16_hi(a.w) * bf16_hi(b.w);
      }
      Dr[i] = acc;
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");   // every thread has read its O rows: the P buffer may be overwritten
    DIG_STAMP(1, 0, 1);
    for (int it = 0; it < 4; ++it) {
      const int j = it >> 1, i = it & 1;
      DIG_STAMP(1, 1 + it, 0);
      mbar_wait(bar_sdp, it & 1);
      tc_fence_after();
      DIG_STAMP(1, 1 + it, 1);
      // phase A: S, dP (TMEM) -> P, dS as packed bf16 in registers; overlaps the tensor pipe's dV,dK,dQ of the previous block
      uint32_t pp[32], ds[32];
#pragma unroll
      for (int c = 0; c < 64; c += 32) {
        uint32_t s[32], g[32];
        tmem_ld32(tl + cS + hh * 64 + c, s);
        tmem_ld32(tl + cdP + hh * 64 + c, g);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 32; q += 2) {
          // dS is kept unscaled here; `scale` is applied once per output element when dK and dQ leave TMEM
          const float p0 = ex2_approx(fmaf(__uint_as_float(s[q]), sl2, -Lr[i]));
          const float p1 = ex2_approx(fmaf(__uint_as_float(s[q + 1]), sl2, -Lr[i]));
          const float d0 = p0 * (__uint_as_float(g[q]) - Dr[i]);
          const float d1 = p1 * (__uint_as_float(g[q + 1]) - Dr[i]);
          pp[(c + q) >> 1] = pack_bf16(p0, p1);
          ds[(c + q) >> 1] = pack_bf16(d0, d1);
        }
      }
      // phase B: the previous block's products have read the P / dS buffers (and, for it == 2, so has the TMA store staged in them)
      if (it > 0) {
        mbar_wait(bar_mma, (it - 1) & 1);
        tc_fence_after();
      }
      DIG_STAMP(1, 1 + it, 2);
      if (it == 2) {
        // dV_0, dK_0 are complete (bar_mma of block 1): TMEM -> the idle P / dS buffers -> two TMA stores
        stage_tile(cdV, sP_s, 1.0f);
        stage_tile(cdK, sdS_s, scale);
        tc_fence_before();
        fence_proxy_async_smem();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (warp == 0 && lane == 0) {
          tma_store_2d(&tm_dqkv, sP_s, 2 * d + head * kHd, row0);
          tma_store_2d(&tm_dqkv, sdS_s, d + head * kHd, row0);
          tma_store_commit();
          tma_store_wait_read_all();
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const uint32_t off = (uint32_t)hh * 16384u + sw128_offset((uint32_t)t, (uint32_t)jj);
        sts_u4(sP_s + off, make_uint4(pp[4 * jj], pp[4 * jj + 1], pp[4 * jj + 2], pp[4 * jj + 3]));
        sts_u4(sdS_s + off, make_uint4(ds[4 * jj], ds[4 * jj + 1], ds[4 * jj + 2], ds[4 * jj + 3]));
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(bar_pds);
      DIG_STAMP(1, 1 + it, 3);
    }
    mbar_wait(bar_mma, 1);   // block 3: every accumulator is final
    tc_fence_after();
    DIG_STAMP(1, 5, 0);
    // dV_1, dK_1 and dQ for both query tiles: the four 16 KB halves of the P / dS buffers stage one tile each
    stage_tile(cdV, sP_s, 1.0f);
    stage_tile(cdK, sdS_s, scale);
    stage_tile(cdQ, sP_s + 16384, scale);
    stage_tile(cdQ + 64, sdS_s + 16384, scale);
    fence_proxy_async_smem();
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (warp == 0 && lane == 0) {
      tma_store_2d(&tm_dqkv, sP_s, 2 * d + head * kHd, row0 + 128);
      tma_store_2d(&tm_dqkv, sdS_s, d + head * kHd, row0 + 128);
      tma_store_2d(&tm_dqkv, sP_s + 16384, head * kHd, row0);
      tma_store_2d(&tm_dqkv, sdS_s + 16384, head * kHd, row0 + 128);
      tma_store_commit();
      tma_store_wait_all();
    }
  }

  DIG_STAMP(warp == 8 ? 0 : 1, 5, 1);
  tc_fence_before();
  __syncthreads();
  DIG_STAMP(warp == 8 ? 0 : 1, 5, 2);
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// backward, persistent: one CTA per SM walks (sequence, head) items.  Same block schedule as attn_bwd_kernel (4 blocks of 128 queries x
// 128 keys per item, P/dS held in registers across the previous block's gradient products), plus:
//   * D = rowsum(dO o O) comes in precomputed (the output-projection dgrad GEMM emits it, DIG_EPI_ROWDOT), so O is never read here;
//   * operand tiles live in rings -- three (Q_i | dO_i) slots and two (K_j | V_j) slots -- that a dedicated TMA warp refills as soon as
//     the MMA thread's commits release them: the next item's first tiles land while this item is still computing, which hides the
//     ~5 us load latency that the one-shot kernel pays per item (1 CTA per SM: nobody else covers it);
//   * the last outputs of item n (dV_1, dK_1, dQ_0, dQ_1) are staged and stored while the tensor pipe already works on item n+1.
// Warps 0-7 compute (warp & 3 = TMEM lane quarter, warp >> 2 = which 64 key columns / 32 output columns), warp 8 = MMA issue + TMEM
// owner, warp 9 = TMA producer.
// ------------------------------------------------------------------------------------------------
#ifndef DIG_ATTN_BWD_STG
#define DIG_ATTN_BWD_STG 0   // experiment, measured SLOWER (146.7 vs 107.5 us): see stg_regs below
#endif
static constexpr int kBwdPThreads = 320;
static constexpr uint32_t kBwdQdO = 0, kBwdKV = 98304, kBwdP = 163840, kBwdDS = 196608, kBwdBars = 229376;

__global__ void __launch_bounds__(kBwdPThreads, 1)
attn_bwd_persist_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do,
                        const __grid_constant__ CUtensorMap tm_dqkv, const float* __restrict__ lse, const float* __restrict__ Dsum, int heads,
                        float scale, int num_items, __nv_bfloat16* __restrict__ dqkv_out) {
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kBwdBars);
  uint64_t* qdo_full = bars + 0;    // [3]
  uint64_t* qdo_empty = bars + 3;   // [3]
  uint64_t* kv_full = bars + 6;     // [2]
  uint64_t* kv_empty = bars + 8;    // [2]
  uint64_t* bar_sdp = bars + 10;
  uint64_t* bar_pds = bars + 11;
  uint64_t* bar_mma = bars + 12;
  uint64_t* bar_sc = bars + 13;     // compute -> MMA (256 arrivals): S, dP of the block are in registers, their TMEM columns may be overwritten
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 14);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int d = heads * kHd;
  const int my_items = (num_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int total = my_items * 4;   // blocks (it = 2 j + i) this CTA processes
  long long* const dbg = (blockIdx.x == 0 && (warp == 8 || threadIdx.x == 0)) ? g_attn_dbg : nullptr;   // clock stamps (bring-up)

  if (warp == 9 && lane == 0) {
    tma_prefetch_desc(&tm_qkv);
    tma_prefetch_desc(&tm_do);
    tma_prefetch_desc(&tm_dqkv);
    for (int i = 0; i < 3; ++i) { mbar_init(&qdo_full[i], 1); mbar_init(&qdo_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    mbar_init(bar_sdp, 1);
    mbar_init(bar_pds, 256);
    mbar_init(bar_mma, 1);
    mbar_init(bar_sc, 256);
    mbar_fence_init();
  }
  if (warp == 8) tmem_alloc(tmem_holder, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_holder, 0);   // warp-uniform for the compiler too
  pdl_wait();   // prologue done: qkv / dO / lse / D of earlier kernels are read from here on
  constexpr uint32_t cS = 0, cdP = 128, cdV = 256, cdK = 320, cdQ = 384;

  if (warp == 9) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int n = 0;
      for (int w = blockIdx.x; w < num_items; w += gridDim.x, ++n) {
        const int head = w % heads, row0 = (w / heads) * kTok;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int T = 2 * n + i, sl = T % 3;
          const uint32_t use = (uint32_t)(T / 3) & 1u;
          uint8_t* q = smem + kBwdQdO + sl * 32768;
          mbar_wait(&qdo_empty[sl], use ^ 1u);
          mbar_expect_tx(&qdo_full[sl], 32768);
          tma_load_2d(q, &tm_qkv, &qdo_full[sl], head * kHd, row0 + i * 128);
          tma_load_2d(q + 16384, &tm_do, &qdo_full[sl], head * kHd, row0 + i * 128);
          uint8_t* kv = smem + kBwdKV + i * 32768;   // K_j | V_j with j = i
          mbar_wait(&kv_empty[i], ((uint32_t)n & 1u) ^ 1u);
          mbar_expect_tx(&kv_full[i], 32768);
          tma_load_2d(kv, &tm_qkv, &kv_full[i], d + head * kHd, row0 + i * 128);
          tma_load_2d(kv + 16384, &tm_qkv, &kv_full[i], 2 * d + head * kHd, row0 + i * 128);
        }
      }
    }
  } else if (warp == 8) {
    // ===================== MMA issuer (the whole warp walks the loop, one elected lane issues: no per-instruction waterfall) ==========
    if (total > 0) {
      constexpr uint32_t id_s = make_idesc_bf16(128, 128, false, false);
      constexpr uint32_t id_tt = make_idesc_bf16(128, 64, true, true);   // dV, dK: A = P^T / dS^T (MN-major), B MN-major
      constexpr uint32_t id_q = make_idesc_bf16(128, 64, false, true);   // dQ: A = dS (K-major), B = K (MN-major)
      const uint32_t base = smem_u32(smem), aP = base + kBwdP, adS = base + kBwdDS;
      auto qdo_slot = [&](int g) { return (2 * (g >> 2) + (g & 1)) % 3; };                       // ring slot of (Q_i | dO_i) for block g
      auto issue_sdp = [&](int g) {   // S = Q_i K_j^T, dP = dO_i V_j^T for block g (waits for its tiles)
        const int n = g >> 2, j = (g >> 1) & 1, T = 2 * n + (g & 1), sl = T % 3;
        mbar_wait(&qdo_full[sl], (uint32_t)(T / 3) & 1u);
        mbar_wait(&kv_full[j], (uint32_t)n & 1u);
        tc_fence_after();
        const uint32_t aQ = base + kBwdQdO + sl * 32768, adO = aQ + 16384, aK = base + kBwdKV + j * 32768, aV = aK + 16384;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc_mma_ss(tmem + cS, make_sdesc_sw128(aQ + k * 32, 16, 1024), make_sdesc_sw128(aK + k * 32, 16, 1024), id_s, k > 0);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc_mma_ss(tmem + cdP, make_sdesc_sw128(adO + k * 32, 16, 1024), make_sdesc_sw128(aV + k * 32, 16, 1024), id_s, k > 0);
          tc_commit(bar_sdp);
        }
        __syncwarp();
      };
      issue_sdp(0);
      for (int g = 0; g < total; ++g) {
        const int it = g & 3, j = it >> 1, i = it & 1, sl = qdo_slot(g);
        const uint32_t aQ = base + kBwdQdO + sl * 32768, adO = aQ + 16384, aK = base + kBwdKV + j * 32768;
        // The next block's scores are issued as soon as this block's S, dP have left TMEM -- while the compute warps are still storing
        // P / dS to shared memory -- so the tensor pipe's S,dP latency (8 MMAs + commit) hides behind those stores instead of
        // following them; they also go ahead of this block's gradient products.
        if (g + 1 < total) {
          mbar_wait(bar_sc, (uint32_t)g & 1u);
          tc_fence_after();
          if (lane == 0) DIG_STAMP(0, g, 0);
          issue_sdp(g + 1);
          if (lane == 0) DIG_STAMP(0, g, 1);
        }
        mbar_wait(bar_pds, (uint32_t)g & 1u);
        tc_fence_after();
        if (lane == 0) DIG_STAMP(0, g, 2);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 8; ++k)  // dV_j += P^T dO_i   (reduction over 128 queries)
            tc_mma_ss(tmem + cdV, make_sdesc_sw128(aP + k * 2048, 16384, 1024), make_sdesc_sw128(adO + k * 2048, 8192, 1024), id_tt,
                      (i > 0 || k > 0));
#pragma unroll
          for (int k = 0; k < 8; ++k)  // dK_j += dS^T Q_i
            tc_mma_ss(tmem + cdK, make_sdesc_sw128(adS + k * 2048, 16384, 1024), make_sdesc_sw128(aQ + k * 2048, 8192, 1024), id_tt,
                      (i > 0 || k > 0));
#pragma unroll
          for (int k = 0; k < 8; ++k)  // dQ_i += dS K_j     (reduction over 128 keys)
            tc_mma_ss(tmem + cdQ + i * 64, make_sdesc_sw128(adS + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024),
                      make_sdesc_sw128(aK + k * 2048, 8192, 1024), id_q, (j > 0 || k > 0));
          tc_commit(bar_mma);
          // ring releases: a commit arrives once every MMA issued so far (including the S,dP of block g+1 above) has completed
          if (it == 1) tc_commit(&kv_empty[0]);
          else if (it == 2) tc_commit(&qdo_empty[sl]);                          // (Q_0 | dO_0): blocks 0 and 2
          else if (it == 3) { tc_commit(&qdo_empty[sl]); tc_commit(&kv_empty[1]); }
        }
        __syncwarp();
        if (lane == 0) DIG_STAMP(0, g, 3);
      }
    }
  } else {
    // ===================== compute warps =====================
    const int t = (warp & 3) * 32 + lane;  // row within a 128-token tile == TMEM lane
    const int hh = warp >> 2;
    const uint32_t tl = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const float sl2 = scale * kLog2e;
    const uint32_t sP_s = smem_u32(smem) + kBwdP, sdS_s = smem_u32(smem) + kBwdDS;
    const bool storer = (warp == 0 && lane == 0);
    auto stage_regs = [&](const uint32_t (&v)[32], uint32_t dst_s, float mul) {
#pragma unroll
      for (int q = 0; q < 32; q += 8) {
        uint4 p;
        p.x = pack_bf16(__uint_as_float(v[q + 0]) * mul, __uint_as_float(v[q + 1]) * mul);
        p.y = pack_bf16(__uint_as_float(v[q + 2]) * mul, __uint_as_float(v[q + 3]) * mul);
        p.z = pack_bf16(__uint_as_float(v[q + 4]) * mul, __uint_as_float(v[q + 5]) * mul);
        p.w = pack_bf16(__uint_as_float(v[q + 6]) * mul, __uint_as_float(v[q + 7]) * mul);
        sts_u4(dst_s + sw128_offset((uint32_t)t, (uint32_t)((hh * 32 + q) >> 3)), p);
      }
    };
    // DIG_ATTN_BWD_STG=1 (experiment): finished gradient tiles go from registers straight to global memory (each thread owns 64
    // contiguous bytes of a row: 4 x STG.128).  The TMA staging in the P / dS buffers costs 2-3.4 k clocks per window (clock stamps: two
    // block-wide barriers plus the wait until the TMA engine has read 32-64 KB back out of shared memory), but the row-per-thread stores
    // are worse: every STG.128 touches 32 different lines and the window grew to 5.2 k clocks (kernel 107.5 -> 146.7 us).
    auto stg_regs = [&](const uint32_t (&v)[32], int col, int row, float mul) {
      uint4* dst = reinterpret_cast<uint4*>(dqkv_out + (long long)(row + t) * (3 * d) + col + hh * 32);
#pragma unroll
      for (int q = 0; q < 32; q += 8) {
        uint4 p;
        p.x = pack_bf16(__uint_as_float(v[q + 0]) * mul, __uint_as_float(v[q + 1]) * mul);
        p.y = pack_bf16(__uint_as_float(v[q + 2]) * mul, __uint_as_float(v[q + 3]) * mul);
        p.z = pack_bf16(__uint_as_float(v[q + 4]) * mul, __uint_as_float(v[q + 5]) * mul);
        p.w = pack_bf16(__uint_as_float(v[q + 6]) * mul, __uint_as_float(v[q + 7]) * mul);
        dst[q >> 3] = p;
      }
    };
    auto store_item_tail = [&](int head, int row0) {
#if DIG_ATTN_BWD_STG
      uint32_t v0[32], v1[32], v2[32], v3[32];
      tmem_ld32(tl + cdV + hh * 32, v0);
      tmem_ld32(tl + cdK + hh * 32, v1);
      tmem_ld32(tl + cdQ + hh * 32, v2);
      tmem_ld32(tl + cdQ + 64 + hh * 32, v3);
      tmem_ld_wait();
      tc_fence_before();
      stg_regs(v0, 2 * d + head * kHd, row0 + 128, 1.0f);
      stg_regs(v1, d + head * kHd, row0 + 128, scale);
      stg_regs(v2, head * kHd, row0, scale);
      stg_regs(v3, head * kHd, row0 + 128, scale);
      return;
#endif
      {   // all four TMEM loads in flight before the first conversion
        uint32_t v0[32], v1[32], v2[32], v3[32];
        tmem_ld32(tl + cdV + hh * 32, v0);
        tmem_ld32(tl + cdK + hh * 32, v1);
        tmem_ld32(tl + cdQ + hh * 32, v2);
        tmem_ld32(tl + cdQ + 64 + hh * 32, v3);
        tmem_ld_wait();
        stage_regs(v0, sP_s, 1.0f);
        stage_regs(v1, sdS_s, scale);
        stage_regs(v2, sP_s + 16384, scale);
        stage_regs(v3, sdS_s + 16384, scale);
      }
      tc_fence_before();
      fence_proxy_async_smem();
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (storer) {
        tma_store_2d(&tm_dqkv, sP_s, 2 * d + head * kHd, row0 + 128);
        tma_store_2d(&tm_dqkv, sdS_s, d + head * kHd, row0 + 128);
        tma_store_2d(&tm_dqkv, sP_s + 16384, head * kHd, row0);
        tma_store_2d(&tm_dqkv, sdS_s + 16384, head * kHd, row0 + 128);
        tma_store_commit();
      }
    };
    // Row statistics (log-sum-exp, D) of an item are fetched one item ahead (during its predecessor's third block): read at the item's
    // start, the two dependent global loads stalled every compute warp for ~2 k clocks per item (clock stamps).
    float Dr[2] = {0.f, 0.f}, Lr[2] = {0.f, 0.f}, Dn[2] = {0.f, 0.f}, Ln[2] = {0.f, 0.f};
    auto fetch_stats = [&](int item) {
      const int w = (int)blockIdx.x + item * (int)gridDim.x;
      const int hd = w % heads, sq = w / heads;
#pragma unroll
      for (int ii = 0; ii < 2; ++ii) {
        Ln[ii] = lse[((long long)sq * heads + hd) * kTok + ii * 128 + t];
        Dn[ii] = Dsum[((long long)sq * kTok + ii * 128 + t) * heads + hd];
      }
    };
    if (total > 0) fetch_stats(0);
    int head = 0, row0 = 0, prev_head = 0, prev_row0 = 0;
    for (int g = 0; g < total; ++g) {
      const int it = g & 3, i = it & 1;
      if (it == 0) {
        prev_head = head; prev_row0 = row0;
        const int w = (int)blockIdx.x + (g >> 2) * (int)gridDim.x;
        head = w % heads; row0 = (w / heads) * kTok;
#pragma unroll
        for (int ii = 0; ii < 2; ++ii) { Lr[ii] = Ln[ii] * kLog2e; Dr[ii] = Dn[ii]; }
      } else if (it == 2 && g + 2 < total) {
        fetch_stats((g >> 2) + 1);
      }
      DIG_STAMP(1, g, 0);
      mbar_wait(bar_sdp, (uint32_t)g & 1u);
      tc_fence_after();
      DIG_STAMP(1, g, 1);
      // phase A: S, dP (TMEM) -> P, dS as packed bf16 in registers; overlaps the tensor pipe's dV,dK,dQ of the previous block
      // (the second 32-column chunk's TMEM loads are in flight while the first chunk is processed)
      uint32_t pp[32], ds[32];
      {
        uint32_t sv0[32], gv0[32], sv1[32], gv1[32];
        tmem_ld32(tl + cS + hh * 64, sv0);
        tmem_ld32(tl + cdP + hh * 64, gv0);
        tmem_ld_wait();
        tmem_ld32(tl + cS + hh * 64 + 32, sv1);
        tmem_ld32(tl + cdP + hh * 64 + 32, gv1);
#pragma unroll
        for (int q = 0; q < 32; q += 2) {
          const float p0 = ex2_approx(fmaf(__uint_as_float(sv0[q]), sl2, -Lr[i]));
          const float p1 = ex2_approx(fmaf(__uint_as_float(sv0[q + 1]), sl2, -Lr[i]));
          const float d0 = p0 * (__uint_as_float(gv0[q]) - Dr[i]);
          const float d1 = p1 * (__uint_as_float(gv0[q + 1]) - Dr[i]);
          pp[q >> 1] = pack_bf16(p0, p1);
          ds[q >> 1] = pack_bf16(d0, d1);
        }
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(bar_sc);     // S, dP of this block are in registers: the MMA warp may issue the next block's scores
#pragma unroll
        for (int q = 0; q < 32; q += 2) {
          const float p0 = ex2_approx(fmaf(__uint_as_float(sv1[q]), sl2, -Lr[i]));
          const float p1 = ex2_approx(fmaf(__uint_as_float(sv1[q + 1]), sl2, -Lr[i]));
          const float d0 = p0 * (__uint_as_float(gv1[q]) - Dr[i]);
          const float d1 = p1 * (__uint_as_float(gv1[q + 1]) - Dr[i]);
          pp[(32 + q) >> 1] = pack_bf16(p0, p1);
          ds[(32 + q) >> 1] = pack_bf16(d0, d1);
        }
      }
      DIG_STAMP(1, g, 2);
      // phase B: the previous block's gradient products have read the P / dS buffers
      if (g > 0) {
        mbar_wait(bar_mma, (uint32_t)(g - 1) & 1u);
        tc_fence_after();
      }
      DIG_STAMP(1, g, 3);
      if (it == 2 && DIG_ATTN_BWD_STG) {
        // dV_0, dK_0 are final (block 1)
        uint32_t v0[32], v1[32];
        tmem_ld32(tl + cdV + hh * 32, v0);
        tmem_ld32(tl + cdK + hh * 32, v1);
        tmem_ld_wait();
        tc_fence_before();
        stg_regs(v0, 2 * d + head * kHd, row0, 1.0f);
        stg_regs(v1, d + head * kHd, row0, scale);
      } else if (it == 2) {
        // dV_0, dK_0 are final (block 1): TMEM -> the idle P / dS buffers -> two TMA stores
        {
          uint32_t v0[32], v1[32];
          tmem_ld32(tl + cdV + hh * 32, v0);
          tmem_ld32(tl + cdK + hh * 32, v1);
          tmem_ld_wait();
          stage_regs(v0, sP_s, 1.0f);
          stage_regs(v1, sdS_s, scale);
        }
        tc_fence_before();
        fence_proxy_async_smem();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (storer) {
          tma_store_2d(&tm_dqkv, sP_s, 2 * d + head * kHd, row0);
          tma_store_2d(&tm_dqkv, sdS_s, d + head * kHd, row0);
          tma_store_commit();
          tma_store_wait_read_all();
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
      } else if (it == 0 && g > 0) {
        // the previous item is complete (its block 3): drain its last four tiles, then reuse the buffers for this item's block 0
        store_item_tail(prev_head, prev_row0);
        if (!DIG_ATTN_BWD_STG) {
          if (storer) tma_store_wait_read_all();
          asm volatile("bar.sync 1, 256;" ::: "memory");
        }
      }
      DIG_STAMP(1, g, 4);
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const uint32_t off = (uint32_t)hh * 16384u + sw128_offset((uint32_t)t, (uint32_t)jj);
        sts_u4(sP_s + off, make_uint4(pp[4 * jj], pp[4 * jj + 1], pp[4 * jj + 2], pp[4 * jj + 3]));
        sts_u4(sdS_s + off, make_uint4(ds[4 * jj], ds[4 * jj + 1], ds[4 * jj + 2], ds[4 * jj + 3]));
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(bar_pds);
      DIG_STAMP(1, g, 5);
    }
    if (total > 0) {
      mbar_wait(bar_mma, (uint32_t)(total - 1) & 1u);
      tc_fence_after();
      store_item_tail(head, row0);
      if (storer) tma_store_wait_all();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

}  // namespace dig

// bring-up only (not part of include/dig_b200.h): device buffer of 3 x 12 x 8 int64 clock stamps, or NULL to switch recording off
extern "C" int dig_attention_debug_buffer(void* buf) {
  long long* p = reinterpret_cast<long long*>(buf);
  return cudaMemcpyToSymbol(dig::g_attn_dbg, &p, sizeof(p)) == cudaSuccess ? 0 : -2;
}

extern "C" int dig_attention_fwd(const void* qkv, void* out, float* lse, int64_t num_seqs, int32_t heads, float scale,
                                 int32_t p_in_smem, void* stream) {
  using namespace dig;
  DIG_REQUIRE(qkv && out, "dig_attention_fwd: null pointer");
  DIG_REQUIRE(num_seqs > 0 && heads > 0, "dig_attention_fwd: empty problem");
  const int d = heads * kHd;
  CUtensorMap tm;
  int rc = make_tmap_bf16_2d(&tm, qkv, (uint64_t)num_seqs * kTok, (uint64_t)3 * d, (uint64_t)3 * d, 128, 64);
  if (rc) return rc;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int grid = (int)(num_seqs * heads * 2);
  static int one_shot = -1;  // DIG_ATTN_FWD_ONESHOT=1: the non-persistent kernel (A/B experiments)
  if (one_shot < 0) { const char* e = getenv("DIG_ATTN_FWD_ONESHOT"); one_shot = (e && e[0] == '1') ? 1 : 0; }
  if (!p_in_smem && !one_shot) {
    const int smem = 2 * kFwdPBuf + 2 * 16384 + 1024 + 256;
    CUtensorMap to;
    rc = make_tmap_bf16_2d(&to, out, (uint64_t)num_seqs * kTok, (uint64_t)d, (uint64_t)d, 128, 64);
    if (rc) return rc;
    static bool set = false;
    if (!set) { DIG_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_persist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); set = true; }
    const int items = (int)(num_seqs * heads);
    const int ctas = items < num_sms() ? items : num_sms();
    DIG_CHECK_CUDA(launch_pdl(attn_fwd_persist_kernel, dim3(ctas), dim3(kFwdPThreads), smem, s, tm, to, lse, heads, scale, items));
  } else if (!p_in_smem) {
    const int smem = 81920 + 1024 + 128;
    static bool set = false;
    if (!set) { DIG_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); set = true; }
    attn_fwd_kernel<true><<<grid, 160, smem, s>>>(tm, reinterpret_cast<__nv_bfloat16*>(out), lse, heads, scale);
  } else {
    const int smem = 81920 + 65536 + 1024 + 128;
    static bool set = false;
    if (!set) { DIG_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); set = true; }
    attn_fwd_kernel<false><<<grid, 160, smem, s>>>(tm, reinterpret_cast<__nv_bfloat16*>(out), lse, heads, scale);
  }
  DIG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int dig_attention_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, int64_t num_seqs,
                                 int32_t heads, float scale, void* stream) {
  using namespace dig;
  DIG_REQUIRE(qkv && out && dout && lse && dqkv, "dig_attention_bwd: null pointer");
  DIG_REQUIRE(num_seqs > 0 && heads > 0, "dig_attention_bwd: empty problem");
  const int d = heads * kHd;
  CUtensorMap tq, td, to, tg;
  int rc = make_tmap_bf16_2d(&tq, qkv, (uint64_t)num_seqs * kTok, (uint64_t)3 * d, (uint64_t)3 * d, 128, 64);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&td, dout, (uint64_t)num_seqs * kTok, (uint64_t)d, (uint64_t)d, 128, 64);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&to, out, (uint64_t)num_seqs * kTok, (uint64_t)d, (uint64_t)d, 128, 64);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tg, dqkv, (uint64_t)num_seqs * kTok, (uint64_t)3 * d, (uint64_t)3 * d, 128, 64);
  if (rc) return rc;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int smem = 196608 + 1024 + 128;
  static bool set = false;
  if (!set) { DIG_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); set = true; }
  attn_bwd_kernel<<<(int)(num_seqs * heads), kBwdThreads, smem, s>>>(tq, td, to, tg, lse, heads, scale);
  DIG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// Persistent backward: D = rowsum(dO o O) [num_seqs*256, heads] fp32 is supplied by the caller (dig_gemm with DIG_EPI_ROWDOT on the
// output-projection dgrad), so the forward output is not read.
extern "C" int dig_attention_bwd_d(const void* qkv, const void* dout, const float* lse, const float* dsum, void* dqkv, int64_t num_seqs,
                                   int32_t heads, float scale, void* stream) {
  using namespace dig;
  DIG_REQUIRE(qkv && dout && lse && dsum && dqkv, "dig_attention_bwd_d: null pointer");
  DIG_REQUIRE(num_seqs > 0 && heads > 0, "dig_attention_bwd_d: empty problem");
  const int d = heads * kHd;
  CUtensorMap tq, td, tg;
  int rc = make_tmap_bf16_2d(&tq, qkv, (uint64_t)num_seqs * kTok, (uint64_t)3 * d, (uint64_t)3 * d, 128, 64);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&td, dout, (uint64_t)num_seqs * kTok, (uint64_t)d, (uint64_t)d, 128, 64);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tg, dqkv, (uint64_t)num_seqs * kTok, (uint64_t)3 * d, (uint64_t)3 * d, 128, 64);
  if (rc) return rc;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int smem = (int)kBwdBars + 256 + 1024;
  static bool set = false;
  if (!set) { DIG_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_persist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); set = true; }
  const int items = (int)(num_seqs * heads);
  const int ctas = items < num_sms() ? items : num_sms();
  DIG_CHECK_CUDA(launch_pdl(attn_bwd_persist_kernel, dim3(ctas), dim3(kBwdPThreads), smem, s, tq, td, tg, lse, dsum, heads, scale, items,
                            reinterpret_cast<__nv_bfloat16*>(dqkv)));
  DIG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

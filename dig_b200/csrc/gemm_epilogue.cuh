// dig_b200 -- fused GEMM epilogue shared by the 1-CTA (gemm.cu) and 2-CTA (gemm2.cu) tcgen05 kernels.
//
// Eight epilogue warps drain one 128-row accumulator: warp `ew` owns TMEM lane quarter (warp index & 3) and one half of the
// tile's columns.  Per 32-column chunk: tcgen05.ld (one accumulator row per thread) -> 32x32 fp32 transpose through a per-warp
// XOR-swizzled shared-memory tile -> lane l owns 4 consecutive columns (l & 7) of rows (l >> 3) + 4 i, so bias is one float4 per
// chunk and every global access covers whole 128-byte row segments.  Operands the epilogue reads from HBM (residual rows or the
// bf16 aux rows) are requested one chunk ahead -- the first chunk before the accumulator wait, so that latency hides behind the
// tile's MMAs -- and are held as raw bits until used.
#pragma once
#include "common.cuh"
#include "../../include/dig_b200.h"

namespace dig {

static constexpr int kEpiWarps = 8;
static constexpr int kEpiAtomic = 4;  // MODE value: split-K fp32 atomic accumulate
// internal MODE values: the GELU epilogues with the pre-activation kept as 8-bit codes (dig_gemm_t.aux_q8) are separate instantiations,
// so neither variant carries the other's code
static constexpr int kEpiGeluQ8 = 6, kEpiGeluBwdQ8 = 7;
__host__ __device__ constexpr bool epi_is_gelu(int m) { return m == DIG_EPI_GELU || m == kEpiGeluQ8; }
__host__ __device__ constexpr bool epi_is_gelu_bwd(int m) { return m == DIG_EPI_GELU_BWD || m == kEpiGeluBwdQ8; }
__host__ __device__ constexpr bool epi_is_q8(int m) { return m == kEpiGeluQ8 || m == kEpiGeluBwdQ8; }

struct GemmEpilogue {
  void* out;
  long long ldo;
  const float* bias;
  const float* residual;
  long long ldr;
  long long res_row_mod;
  const uint8_t* row_mask;
  const float* row_mask_value;
  void* aux;
  long long ldaux;
  float alpha;
  float* colsum;
  float* rowdot;      // DIG_EPI_ROWDOT (TMA epilogue only)
  long long ldrowdot;
  int M;
  const float* lut;   // aux_q8: device table [256] of gelu_erf'(decoded pre-activation)
  int dbg;  // bring-up only (env DIG_GEMM_DBG): 1 = skip the global stores, 2 = skip the whole epilogue body
};

// A device-resident zero vector standing in for a missing bias in the GELU epilogue (so that the bias add is unconditional there).
const float* zero_bias();
static constexpr int kZeroBiasLen = 8192;

// 8-bit code of a GELU pre-activation (dig_gemm_t.aux_q8): 256 uniform levels over [-kQ8Range, kQ8Range].  The backward looks
// gelu_erf'(level) up in a 256-entry table (exact values, computed in double on the host) instead of evaluating the rational.
static constexpr float kQ8Range = 4.0f;
const float* gelu_grad_lut();
__device__ __forceinline__ uint32_t q8_bits(float x) {   // low byte = code; two FFMA (saturating scale, then the 2^23 rounding trick)
  float t;
  asm("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(t) : "f"(x), "f"(0.5f / kQ8Range), "f"(0.5f));
  return __float_as_uint(fmaf(t, 255.0f, 8388608.0f));
}
__device__ __forceinline__ uint32_t q8_encode4(float a, float b, float c, float d) {
  return __byte_perm(__byte_perm(q8_bits(a), q8_bits(b), 0x0040), __byte_perm(q8_bits(c), q8_bits(d), 0x0040), 0x5410);
}
// table offset (bytes) of code k (0..3) of a packed word
template <int K>
__device__ __forceinline__ uint32_t q8_lut_off(uint32_t w) {
  if constexpr (K == 0) return (w << 2) & 0x3FCu;
  else return (w >> (8 * K - 2)) & 0x3FCu;
}

__device__ __forceinline__ float4 ld_bf16x4(const __nv_bfloat16* p) {
  const uint2 v = *reinterpret_cast<const uint2*>(p);
  return make_float4(bf16_lo(v.x), bf16_hi(v.x), bf16_lo(v.y), bf16_hi(v.y));
}
__device__ __forceinline__ void st_bf16x4(__nv_bfloat16* p, float4 v) {
  *reinterpret_cast<uint2*>(p) = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
}

template <int MODE>
struct EpiPre {
  float4 v[8];
  uint32_t maskbits;
};

template <int MODE>
__device__ __forceinline__ void epi_prefetch(EpiPre<MODE>& p, const GemmEpilogue& ep, int gcol, long long row_base, int rows_left, int N,
                                             int rsub) {
  p.maskbits = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = i * 4 + rsub;
    p.v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (gcol < N && r < rows_left) {
      const long long grow = row_base + r;
      if (MODE == DIG_EPI_LINEAR) {
        if (ep.residual != nullptr)
          p.v[i] = *reinterpret_cast<const float4*>(ep.residual + (ep.res_row_mod > 0 ? (grow % ep.res_row_mod) : grow) * ep.ldr + gcol);
        if (ep.row_mask != nullptr && ep.row_mask[grow] != 0) p.maskbits |= 1u << i;
      } else if (epi_is_gelu_bwd(MODE) && epi_is_q8(MODE)) {
        p.v[i].x = __uint_as_float(*reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint8_t*>(ep.aux) + grow * ep.ldaux + gcol));
      } else if (epi_is_gelu_bwd(MODE) || MODE == DIG_EPI_RELU_MASK) {
        const uint2 t = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(ep.aux) + grow * ep.ldaux + gcol);
        p.v[i].x = __uint_as_float(t.x);
        p.v[i].y = __uint_as_float(t.y);
      }
    }
  }
}

// Process one 32-column chunk whose accumulator values are already in `v` (this thread's row).
template <int MODE, bool OUT_F32>
__device__ __forceinline__ void epi_chunk(const uint32_t (&v)[32], const EpiPre<MODE>& pre, const GemmEpilogue& ep, int gcol, long long row_base,
                                          int rows_left, int N, float* tile, float* cta_colsum, int lane) {
  const int col4 = lane & 7, rsub = lane >> 3;
  const uint32_t tile_s = smem_u32(tile);
#pragma unroll
  for (int j = 0; j < 8; ++j)
    sts_f4(tile_s + (uint32_t)(lane * 32 + ((j ^ (lane & 7)) << 2)) * 4u,
           make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])));
  __syncwarp();
  const bool col_ok = gcol < N;
  float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
  if (col_ok) {
    float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (MODE != kEpiAtomic && ep.bias != nullptr) b4 = __ldg(reinterpret_cast<const float4*>(ep.bias + gcol));
    float* const out_f = reinterpret_cast<float*>(ep.out);
    __nv_bfloat16* const out_h = reinterpret_cast<__nv_bfloat16*>(ep.out);
    const float alpha = ep.alpha;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = i * 4 + rsub;
      if (r >= rows_left) continue;
      const long long grow = row_base + r;
      float4 f = lds_f4(tile_s + (uint32_t)(r * 32 + ((col4 ^ (r & 7)) << 2)) * 4u);
      if (alpha != 1.0f) { f.x *= alpha; f.y *= alpha; f.z *= alpha; f.w *= alpha; }
      if (MODE == kEpiAtomic) {
        float* o = out_f + grow * ep.ldo + gcol;
        atomicAdd(o, f.x); atomicAdd(o + 1, f.y); atomicAdd(o + 2, f.z); atomicAdd(o + 3, f.w);
        continue;
      }
      f.x += b4.x; f.y += b4.y; f.z += b4.z; f.w += b4.w;
      if (epi_is_gelu(MODE)) {
        if (ep.aux != nullptr) {
          if (epi_is_q8(MODE)) *reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(ep.aux) + grow * ep.ldaux + gcol) = q8_encode4(f.x, f.y, f.z, f.w);
          else st_bf16x4(reinterpret_cast<__nv_bfloat16*>(ep.aux) + grow * ep.ldaux + gcol, f);
        }
        f.x = gelu_erf(f.x); f.y = gelu_erf(f.y); f.z = gelu_erf(f.z); f.w = gelu_erf(f.w);
      } else if (epi_is_gelu_bwd(MODE) || MODE == DIG_EPI_RELU_MASK) {
        const uint32_t lo = __float_as_uint(pre.v[i].x), hi = __float_as_uint(pre.v[i].y);
        const float4 x = make_float4(bf16_lo(lo), bf16_hi(lo), bf16_lo(hi), bf16_hi(hi));
        if (epi_is_gelu_bwd(MODE)) {
          if (epi_is_q8(MODE)) {
            const char* lut = reinterpret_cast<const char*>(ep.lut);
            f.x *= __ldg(reinterpret_cast<const float*>(lut + q8_lut_off<0>(lo))); f.y *= __ldg(reinterpret_cast<const float*>(lut + q8_lut_off<1>(lo)));
            f.z *= __ldg(reinterpret_cast<const float*>(lut + q8_lut_off<2>(lo))); f.w *= __ldg(reinterpret_cast<const float*>(lut + q8_lut_off<3>(lo)));
          } else {
            f.x *= gelu_erf_grad(x.x); f.y *= gelu_erf_grad(x.y); f.z *= gelu_erf_grad(x.z); f.w *= gelu_erf_grad(x.w);
          }
          cs.x += f.x; cs.y += f.y; cs.z += f.z; cs.w += f.w;
        } else {
          f.x = x.x > 0.f ? f.x : 0.f; f.y = x.y > 0.f ? f.y : 0.f; f.z = x.z > 0.f ? f.z : 0.f; f.w = x.w > 0.f ? f.w : 0.f;
        }
      } else if (MODE == DIG_EPI_LINEAR) {
        if (pre.maskbits & (1u << i)) f = __ldg(reinterpret_cast<const float4*>(ep.row_mask_value + gcol));
        f.x += pre.v[i].x; f.y += pre.v[i].y; f.z += pre.v[i].z; f.w += pre.v[i].w;
      }
      if (ep.dbg == 1) continue;
      if (OUT_F32) *reinterpret_cast<float4*>(out_f + grow * ep.ldo + gcol) = f;
      else st_bf16x4(out_h + grow * ep.ldo + gcol, f);
    }
  }
  if (epi_is_gelu_bwd(MODE) && ep.colsum != nullptr) {  // column sums of the written tile: bias gradient of fc1
#pragma unroll
    for (int o = 8; o <= 16; o <<= 1) {
      cs.x += __shfl_xor_sync(0xffffffffu, cs.x, o); cs.y += __shfl_xor_sync(0xffffffffu, cs.y, o);
      cs.z += __shfl_xor_sync(0xffffffffu, cs.z, o); cs.w += __shfl_xor_sync(0xffffffffu, cs.w, o);
    }
    if (col_ok && rsub == 0) {
      const uint32_t cs_s = smem_u32(cta_colsum + gcol);
      red_shared_add_f32(cs_s, cs.x); red_shared_add_f32(cs_s + 4, cs.y);
      red_shared_add_f32(cs_s + 8, cs.z); red_shared_add_f32(cs_s + 12, cs.w);
    }
  }
  __syncwarp();
}

// One epilogue warp, one output tile.  NCOLS = columns this warp owns (multiple of 32); tmem_warp = TMEM address of this warp's lane
// quarter at its first column; gcol0 = global column of that first column.  `release()` is called (by the whole warp, converged) right
// after the warp's last TMEM read so the MMA warp can reuse the accumulator while the stores are still in flight.
template <int NCOLS, int MODE, bool OUT_F32, typename Release>
__device__ __forceinline__ void epilogue_warp_tile(const GemmEpilogue& ep, uint32_t tmem_warp, int gcol0, long long row_base, int M, int N,
                                                   float* tile, float* cta_colsum, int lane, uint64_t* tmem_full, uint32_t full_phase,
                                                   Release release) {
  constexpr int NCH = NCOLS / 32;
  const int col4 = lane & 7, rsub = lane >> 3;
  const int rows_left = (int)min((long long)32, (long long)M - row_base);
  EpiPre<MODE> pre[2];
  epi_prefetch<MODE>(pre[0], ep, gcol0 + col4 * 4, row_base, rows_left, N, rsub);
  mbar_wait(tmem_full, full_phase);
  tc_fence_after();
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) {
    if (ch + 1 < NCH) epi_prefetch<MODE>(pre[(ch + 1) & 1], ep, gcol0 + (ch + 1) * 32 + col4 * 4, row_base, rows_left, N, rsub);
    uint32_t v[32];
    tmem_ld32(tmem_warp + ch * 32, v);
    tmem_ld_wait();
    if (ch == NCH - 1) {
      tc_fence_before();
      __syncwarp();
      release();
    }
    if (ep.dbg == 2) continue;
    epi_chunk<MODE, OUT_F32>(v, pre[ch & 1], ep, gcol0 + ch * 32 + col4 * 4, row_base, rows_left, N, tile, cta_colsum, lane);
  }
}

}  // namespace dig

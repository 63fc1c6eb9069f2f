// dig_b200 -- TMA-staged GEMM epilogue (the fast path; gemm_epilogue.cuh keeps the generic transposing one).
//
// Each epilogue warp keeps its accumulator rows in registers in the natural tcgen05.ld layout (one output row per thread), does the
// fused math there (bias straight from L1 with warp-uniform addresses), and writes 128-byte row segments into a per-warp, 128B-
// swizzled 32 x 128 B staging tile that a single elected lane hands to the TMA engine (cp.async.bulk.tensor store, or
// cp.reduce.async.bulk .add for split-K accumulation).  Operands the epilogue needs from HBM -- the fp32 residual rows, or the bf16
// GELU pre-activation -- arrive the same way: a TMA load into the staging tile, issued one group ahead (the first one before the
// accumulator wait), and the result is written back in place.  Epilogue warps therefore issue no global loads/stores at all, HBM
// traffic is whole 128-byte lines, and tails are clipped by the tensor maps.
#pragma once
#include <stdlib.h>

#include "gemm_epilogue.cuh"

#ifndef DIG_EPI_EARLY_LD
#define DIG_EPI_EARLY_LD 1  // 0: request an epilogue operand tile only one column group ahead (round-1 behaviour, A/B switch)
#endif
#ifndef DIG_GELU_PACKED
#define DIG_GELU_PACKED 1   // 0: scalar FFMA GELU in the epilogues (A/B switch)
#endif

namespace dig {

static constexpr int kStageTileBytes = 4096;                       // 32 rows x 128 B
static constexpr int kEpiTmaBytes = kEpiWarps * 2 * kStageTileBytes;  // two staging tiles per epilogue warp

__device__ __forceinline__ void tma_load_2d_addr(uint32_t smem_dst, const CUtensorMap* m, uint64_t* bar, int c_inner, int c_outer) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c_inner), "r"(c_outer)
      : "memory");
}
// Per-warp persistent state of the TMA epilogue (lives in registers across tiles).
struct EpiTmaState {
  uint32_t stage_s;   // shared address of this warp's two staging tiles
  uint64_t* ld_bar;   // two mbarriers (one per staging tile) for TMA loads into them
  uint32_t uses0, uses1;  // completed-load counters -> phase parity of ld_bar[0/1]
};

// One epilogue warp, one output tile.  BNT = tile width in columns; this warp handles the 128-byte column groups gi = half, half+2, ...
// of its 32 rows.  tmem_warp = TMEM address of the warp's lane quarter at the tile's first accumulator column.
template <int BNT, int MODE, bool OUT_F32, typename Release>
__device__ __forceinline__ void epilogue_warp_tile_tma(const GemmEpilogue& ep, const CUtensorMap* tm_out, const CUtensorMap* tm_aux,
                                                       EpiTmaState& st, uint32_t tmem_warp, int n0, int row_base, int N, int half,
                                                       float* cta_colsum, int lane, uint64_t* tmem_full, uint32_t full_phase,
                                                       Release release) {
  constexpr int G = OUT_F32 ? 32 : 64;            // columns per 128-byte staging row
  constexpr int NGT = (BNT + G - 1) / G;          // column groups in the tile
  static_assert(BNT % G == 0, "tile width must be a multiple of the staging group");
  const bool has_ld = (epi_is_gelu_bwd(MODE)) || (MODE == DIG_EPI_ROWDOT) || (MODE == DIG_EPI_LINEAR && OUT_F32 && ep.residual != nullptr);
  const bool has_bias = ep.bias != nullptr;
  const uint32_t row_s = (uint32_t)lane * 128u;
  const uint32_t sw = (uint32_t)(lane & 7);
  // 8-bit pre-activation codes (dig_gemm_t.aux_q8): the aux tile is 32 rows x 64 B in a 64-byte-swizzled staging tile (16-byte chunk c
  // of row r at r*64 + ((c ^ ((r >> 1) & 3)) << 4)), a quarter of the bf16 tile's bytes; the backward multiplies by a 256-entry table of
  // gelu' held in shared memory (behind the CTA's column-sum scratch) instead of evaluating the rational.
  constexpr bool q8 = epi_is_q8(MODE);
  const uint32_t q8_row = (uint32_t)lane * 64u, q8_sw = ((uint32_t)lane >> 1) & 3u;
  const uint32_t lut_s = smem_u32(cta_colsum) + 2048u * 4u;
  const uint32_t ld_bytes = (epi_is_gelu_bwd(MODE) && q8) ? kStageTileBytes / 2 : kStageTileBytes;
  // patch embed (V:95-99): the residual is the position table, indexed by row % res_row_mod (a multiple of 32, so a 32-row tile never
  // straddles the wrap), and rows flagged in row_mask are replaced by the mask token before the table is added
  const int aux_row = (MODE == DIG_EPI_LINEAR && ep.res_row_mod > 0) ? (int)(row_base % (int)ep.res_row_mod) : row_base;
  const bool row_masked = MODE == DIG_EPI_LINEAR && ep.row_mask != nullptr && row_base + lane < ep.M && ep.row_mask[row_base + lane] != 0;

  // first group's operand load, before the accumulator wait
  if (lane == 0) tma_store_wait_read_all();       // staging tiles of the previous output tile are free again
  __syncwarp();
  if (has_ld && half < NGT && lane == 0) {
    mbar_expect_tx(&st.ld_bar[0], ld_bytes);
    tma_load_2d_addr(st.stage_s, tm_aux, &st.ld_bar[0], n0 + half * G, aux_row);
#if DIG_EPI_EARLY_LD
    if (half + 2 < NGT) {   // the second group's operand as well: both staging tiles are free here, and the load then has the accumulator
      mbar_expect_tx(&st.ld_bar[1], ld_bytes);   // wait plus one whole group of math to arrive (it waited ~15 % of the epilogue's time)
      tma_load_2d_addr(st.stage_s + kStageTileBytes, tm_aux, &st.ld_bar[1], n0 + (half + 2) * G, aux_row);
    }
#endif
  }
  mbar_wait(tmem_full, full_phase);
  tc_fence_after();
  if (half >= NGT) {  // narrow tile: this warp owns no column group, it only hands the accumulator back
    __syncwarp();
    release();
    return;
  }

  int k = 0;
#pragma unroll
  for (int gi0 = 0; gi0 < NGT; gi0 += 2, ++k) {
    const int gi = gi0 + half;
    if (gi >= NGT) break;
    const bool last = (gi + 2 >= NGT);
    const int gcol = n0 + gi * G;
    const uint32_t buf = st.stage_s + (uint32_t)((epi_is_gelu(MODE)) ? 0 : (k & 1)) * kStageTileBytes;
    const uint32_t buf2 = st.stage_s + kStageTileBytes;   // GELU forward: second output (post-activation)
    constexpr int kFirstPrefetch = DIG_EPI_EARLY_LD ? 1 : 0;   // group k prefetches group k+1 unless the tile start already requested it
    if (k > 0 && (!has_ld || (!last && k >= kFirstPrefetch))) {  // the staging tile we are about to overwrite (or prefetch into) must have been read by its TMA store
      if (lane == 0) {
        if (has_ld || epi_is_gelu(MODE)) tma_store_wait_read_all();
        else tma_store_wait_read_1();      // double-buffered: only the store issued two groups ago has to be done
      }
      __syncwarp();
    }
    if (has_ld) {
      if (!last && k >= kFirstPrefetch && lane == 0) {  // prefetch the next group's operand into the other staging tile
        uint64_t* nb = &st.ld_bar[(k + 1) & 1];
        mbar_expect_tx(nb, ld_bytes);
        tma_load_2d_addr(st.stage_s + (uint32_t)((k + 1) & 1) * kStageTileBytes, tm_aux, nb, gcol + 2 * G, aux_row);
      }
      if (k & 1) { mbar_wait(&st.ld_bar[1], st.uses1 & 1); ++st.uses1; }
      else { mbar_wait(&st.ld_bar[0], st.uses0 & 1); ++st.uses0; }
    }
    // accumulator rows -> registers
    uint32_t v[G];
    {
      uint32_t (&v0)[32] = *reinterpret_cast<uint32_t (*)[32]>(&v[0]);
      tmem_ld32(tmem_warp + gi * G, v0);
      if (G == 64) {
        uint32_t (&v1)[32] = *reinterpret_cast<uint32_t (*)[32]>(&v[G - 32]);
        tmem_ld32(tmem_warp + gi * G + 32, v1);
      }
      tmem_ld_wait();
    }
    if (last) {
      tc_fence_before();
      __syncwarp();
      release();
    }
    uint32_t cw[16];   // GELU backward with 8-bit codes: this row's 64 codes, read before any lane overwrites the tile with its output row
    if (epi_is_gelu_bwd(MODE) && q8) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const uint4 t = lds_u4(buf + q8_row + ((((uint32_t)c) ^ q8_sw) << 4));
        cw[4 * c] = t.x; cw[4 * c + 1] = t.y; cw[4 * c + 2] = t.z; cw[4 * c + 3] = t.w;
      }
      __syncwarp();
    }
    uint2 qheld = make_uint2(0u, 0u);   // GELU forward with 8-bit codes: the even unit's 8 codes, stored together with the odd unit's
    const float alpha = ep.alpha;
    if (OUT_F32) {
      // 8 units of 4 fp32 columns
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float4 f = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                               __uint_as_float(v[4 * j + 3]));
        if (alpha != 1.0f) { f.x *= alpha; f.y *= alpha; f.z *= alpha; f.w *= alpha; }
        const uint32_t a = buf + row_s + (((uint32_t)j ^ sw) << 4);
        if (MODE == DIG_EPI_LINEAR) {
          if (has_bias) {   // uniform (kernel parameter); columns past N are clipped by the TMA store, so their address is only clamped
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(ep.bias + min(gcol + 4 * j, N - 4)));
            f.x += b4.x; f.y += b4.y; f.z += b4.z; f.w += b4.w;
          }
          if (row_masked) f = __ldg(reinterpret_cast<const float4*>(ep.row_mask_value + min(gcol + 4 * j, N - 4)));
          if (has_ld) {
            const float4 r4 = lds_f4(a);
            f.x += r4.x; f.y += r4.y; f.z += r4.z; f.w += r4.w;
          }
        }
        sts_f4(a, f);
      }
    } else {
      // 8 units of 8 bf16 columns; the arithmetic runs on fp32 pairs (FADD2 / FFMA2), which halves the issue slots of this issue-bound loop
      const bool scaled = alpha != 1.0f;
      float dot = 0.f;   // DIG_EPI_ROWDOT: this row's dot product with the aux tile over the 64-column group
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[8 * j + e]);
        if (scaled) {
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] *= alpha;
        }
        // GELU forward: the host always supplies a bias (a zero vector if the caller passed none), so the loads sit in the same basic
        // block as the polynomial and ptxas hoists them ahead of it; other modes: uniform branch on the kernel parameter.  Columns past
        // N are clipped by the TMA store, so their bias address is only clamped.
        if (epi_is_gelu(MODE) || has_bias) {
          const int bc = min(gcol + 8 * j, N - 8);
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(ep.bias + bc));
          const float4 b1 = __ldg(reinterpret_cast<const float4*>(ep.bias + bc + 4));
          f2_unpack(f2_add(f2_pack(f[0], f[1]), f2_pack(b0.x, b0.y)), f[0], f[1]);
          f2_unpack(f2_add(f2_pack(f[2], f[3]), f2_pack(b0.z, b0.w)), f[2], f[3]);
          f2_unpack(f2_add(f2_pack(f[4], f[5]), f2_pack(b1.x, b1.y)), f[4], f[5]);
          f2_unpack(f2_add(f2_pack(f[6], f[7]), f2_pack(b1.z, b1.w)), f[6], f[7]);
        }
        const uint32_t off = row_s + (((uint32_t)j ^ sw) << 4);
        if (epi_is_gelu(MODE)) {
          if (ep.aux != nullptr) {   // pre-activation copy for the backward; the no-grad momentum branch skips it (half the stores)
            if (q8) {
              const uint2 cur = make_uint2(q8_encode4(f[0], f[1], f[2], f[3]), q8_encode4(f[4], f[5], f[6], f[7]));
              if (j & 1) sts_u4(buf + q8_row + ((((uint32_t)(j >> 1)) ^ q8_sw) << 4), make_uint4(qheld.x, qheld.y, cur.x, cur.y));
              else qheld = cur;
            } else {
              sts_u4(buf + off, make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7])));
            }
          }
#if DIG_GELU_PACKED
#pragma unroll
          for (int e = 0; e < 8; e += 2) gelu_erf_x2(f[e], f[e + 1], f[e], f[e + 1]);
#else
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = gelu_erf(f[e]);
#endif
          sts_u4(buf2 + off, make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7])));
        } else {
          if (MODE == DIG_EPI_ROWDOT) {
            const uint4 x = lds_u4(buf + off);
            dot += f[0] * bf16_lo(x.x) + f[1] * bf16_hi(x.x) + f[2] * bf16_lo(x.y) + f[3] * bf16_hi(x.y) + f[4] * bf16_lo(x.z) +
                   f[5] * bf16_hi(x.z) + f[6] * bf16_lo(x.w) + f[7] * bf16_hi(x.w);
          }
          if (epi_is_gelu_bwd(MODE) && q8) {
            const uint32_t w0 = cw[2 * j], w1 = cw[2 * j + 1];
            f[0] *= lds_f32(lut_s + q8_lut_off<0>(w0)); f[1] *= lds_f32(lut_s + q8_lut_off<1>(w0));
            f[2] *= lds_f32(lut_s + q8_lut_off<2>(w0)); f[3] *= lds_f32(lut_s + q8_lut_off<3>(w0));
            f[4] *= lds_f32(lut_s + q8_lut_off<0>(w1)); f[5] *= lds_f32(lut_s + q8_lut_off<1>(w1));
            f[6] *= lds_f32(lut_s + q8_lut_off<2>(w1)); f[7] *= lds_f32(lut_s + q8_lut_off<3>(w1));
          } else if (epi_is_gelu_bwd(MODE)) {
            const uint4 x = lds_u4(buf + off);
#if DIG_GELU_PACKED
            gelu_erf_grad_mul_x2(bf16_lo(x.x), bf16_hi(x.x), f[0], f[1]);
            gelu_erf_grad_mul_x2(bf16_lo(x.y), bf16_hi(x.y), f[2], f[3]);
            gelu_erf_grad_mul_x2(bf16_lo(x.z), bf16_hi(x.z), f[4], f[5]);
            gelu_erf_grad_mul_x2(bf16_lo(x.w), bf16_hi(x.w), f[6], f[7]);
#else
            f[0] *= gelu_erf_grad(bf16_lo(x.x)); f[1] *= gelu_erf_grad(bf16_hi(x.x));
            f[2] *= gelu_erf_grad(bf16_lo(x.y)); f[3] *= gelu_erf_grad(bf16_hi(x.y));
            f[4] *= gelu_erf_grad(bf16_lo(x.z)); f[5] *= gelu_erf_grad(bf16_hi(x.z));
            f[6] *= gelu_erf_grad(bf16_lo(x.w)); f[7] *= gelu_erf_grad(bf16_hi(x.w));
#endif
          }
          sts_u4(buf + off, make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7])));
        }
      }
      if (MODE == DIG_EPI_ROWDOT) {
        if (row_base + lane < ep.M && gcol < N) ep.rowdot[(long long)(row_base + lane) * ep.ldrowdot + (gcol >> 6)] = dot;
      }
    }
    if (epi_is_gelu_bwd(MODE) && ep.colsum != nullptr) {
      // column sums of the staged 32 x 64 bf16 tile: lane l owns columns 2l, 2l+1 (bias gradient of fc1)
      __syncwarp();
      float s0 = 0.f, s1 = 0.f;   // rows past M are exact zeros (zero-filled A rows), so they need no masking
#pragma unroll 8
      for (int r = 0; r < 32; ++r) {
        const uint32_t w = lds_u32(buf + (uint32_t)r * 128u + ((((uint32_t)lane >> 2) ^ ((uint32_t)r & 7u)) << 4) + (((uint32_t)lane & 3u) << 2));
        s0 += bf16_lo(w);
        s1 += bf16_hi(w);
      }
      const int c = gcol + 2 * lane;
      if (c < N) {
        const uint32_t cs_s = smem_u32(cta_colsum + c);
        red_shared_add_f32(cs_s, s0);
        red_shared_add_f32(cs_s + 4, s1);
      }
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      if (MODE == kEpiAtomic) tma_reduce_add_2d(tm_out, buf, gcol, row_base);
      else if (epi_is_gelu(MODE)) {
        if (ep.aux != nullptr) tma_store_2d(tm_aux, buf, gcol, row_base);   // pre-activation
        tma_store_2d(tm_out, buf2, gcol, row_base);  // gelu(pre)
      } else tma_store_2d(tm_out, buf, gcol, row_base);
      tma_store_commit();
    }
  }
}

// Host: can this problem use the TMA-staged epilogue?  (tensor maps need 16-byte aligned bases / strides; the row-mask and position-
// table epilogue of the patch embed and the ReLU-mask epilogue stay on the generic path).  DIG_GEMM_TMA_EPI=0 disables it.
static inline bool tma_epilogue_ok(const dig_gemm_t* g) {
  static int enabled = -1;
  if (enabled < 0) { const char* e = getenv("DIG_GEMM_TMA_EPI"); enabled = (e && e[0] == '0') ? 0 : 1; }
  if (!enabled) return false;
  const int es = g->out_fp32 ? 4 : 2;
  if ((g->row_mask || g->res_row_mod > 0) && !(g->out_fp32 && g->epilogue == DIG_EPI_LINEAR && g->res_row_mod % 32 == 0 && g->split_k <= 1)) return false;
  if (g->res_row_mod > 0 && !g->residual) return false;
  if (g->epilogue == DIG_EPI_RELU_MASK) return false;
  if ((g->N * es) % 16 || (g->ldo * es) % 16 || ((uintptr_t)g->out & 15)) return false;
  if (g->epilogue == DIG_EPI_GELU || g->epilogue == DIG_EPI_GELU_BWD || g->epilogue == DIG_EPI_ROWDOT) {
    const int aes = (g->aux_q8 && g->epilogue != DIG_EPI_ROWDOT) ? 1 : 2;
    if (g->out_fp32 || (g->ldaux * aes) % 16 || ((uintptr_t)g->aux & 15)) return false;
  }
  if (g->epilogue == DIG_EPI_ROWDOT && (g->N % 64 != 0 || g->rowdot == nullptr)) return false;
  if (g->residual && (!g->out_fp32 || g->epilogue != DIG_EPI_LINEAR || (g->ldr * 4) % 16 || ((uintptr_t)g->residual & 15))) return false;
  return true;
}

}  // namespace dig

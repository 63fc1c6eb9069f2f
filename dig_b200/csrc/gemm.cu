// dig_b200 -- persistent warp-specialised tcgen05 GEMM for sm_100a.
//
//   out[M,N] = epilogue( alpha * A[M,K] . B[N,K]^T ),  bf16 operands, fp32 accumulation in TMEM.
//
// One CTA per SM walks output tiles of 128 x BN.  Warp 0 feeds a ring of shared-memory stages with TMA
// (128-byte swizzle), warp 1 issues tcgen05.mma from one thread and owns the TMEM allocation, warps 2..9
// drain the double-buffered TMEM accumulator (tcgen05.ld, one row per thread), transpose 32x32 chunks through
// per-warp shared memory and apply the fused epilogue with fully coalesced 128-bit global accesses, so tile i's
// epilogue overlaps tile i+1's MMAs.  Either operand may be K-major or MN-major (see
// include/dig_b200.h), which covers forward (x.W^T), dgrad (dy.W) and wgrad (dy^T.x) without any transposed
// copies.  Replaces the cuBLAS calls behind F.linear in the reference (modeling_finetune.py:93,119,54,58;
// modeling_pretrain_moco_mim_ori.py:463-482,422-426) and their autograd counterparts.
#include <stdlib.h>

#include <math.h>
#include <mutex>

#include "gemm_epilogue_tma.cuh"

namespace dig {

static constexpr int BM = 128;
static constexpr int BK = 64;
static constexpr int kGemmThreads = 64 + kEpiWarps * 32;

template <int BN, bool TMA_EPI>
struct GemmSmem {
  static constexpr int kStageA = BM * BK * 2;
  static constexpr int kStageB = BN * BK * 2;
  static constexpr int kStage = kStageA + kStageB;
  static constexpr int kStages = (BN == 128) ? (TMA_EPI ? 4 : 5) : 6;
  // epilogue scratch: one 32x32 fp32 transpose tile per warp (generic epilogue) or two 32 x 128 B TMA staging tiles per warp
  static constexpr int kEpi = TMA_EPI ? kEpiTmaBytes : kEpiWarps * 32 * 32 * 4;
  static constexpr int kColsum = 2048 * 4 + 1024;        // per-CTA column-sum scratch (N <= 2048 when colsum is requested) + the 256-entry gelu' table (8-bit codes)
  static constexpr int kBytes = kStages * kStage + kEpi + kColsum + 1024 /*align slack*/ + 512 /*barriers*/;
};

// MODE: DIG_EPI_* or kEpiAtomic (split-K fp32 accumulate); OUT_F32: output element type.

template <int BN, bool A_MN, bool B_MN, int MODE, bool OUT_F32, bool TMA_EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_tcgen05(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                  const __grid_constant__ CUtensorMap tma_out, const __grid_constant__ CUtensorMap tma_aux, GemmEpilogue ep, int M,
                  int N, int K, int split_k, int kb_per_split) {
  pdl_launch_dependents();
  using S = GemmSmem<BN, TMA_EPI>;
  constexpr int kStages = S::kStages;
  constexpr uint32_t kTmemCols = (2 * BN <= 128) ? 128 : (2 * BN <= 256 ? 256 : 512);
  constexpr uint32_t kIdesc = make_idesc_bf16(BM, BN, A_MN, B_MN);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* epi_smem = reinterpret_cast<float*>(smem + kStages * S::kStage);
  float* cta_colsum = reinterpret_cast<float*>(smem + kStages * S::kStage + S::kEpi);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * S::kStage + S::kEpi + S::kColsum);
  uint64_t* full_bar = bars;                      // [kStages]  TMA -> MMA
  uint64_t* empty_bar = bars + kStages;           // [kStages]  MMA -> TMA
  uint64_t* tmem_full = bars + 2 * kStages;       // [2]        MMA -> epilogue
  uint64_t* tmem_empty = bars + 2 * kStages + 2;  // [2]        epilogue -> MMA
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);
  uint64_t* epi_ld_bar = bars + 2 * kStages + 5;  // [kEpiWarps][2] TMA loads into the epilogue staging tiles

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int num_m = (M + BM - 1) / BM;
  const int num_n = (N + BN - 1) / BN;
  const int num_kb = (K + BK - 1) / BK;
  const int num_work = num_m * num_n * split_k;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    if (TMA_EPI) { tma_prefetch_desc(&tma_out); tma_prefetch_desc(&tma_aux); }
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], kEpiWarps);
    }
    if (TMA_EPI)
      for (int i = 0; i < 2 * kEpiWarps; ++i) mbar_init(&epi_ld_bar[i], 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_holder, kTmemCols);
  if (epi_is_gelu_bwd(MODE) && ep.colsum != nullptr)
    for (int i = threadIdx.x; i < N; i += kGemmThreads) cta_colsum[i] = 0.f;
  if (epi_is_gelu_bwd(MODE) && epi_is_q8(MODE))   // constant table (written once at library initialisation): safe to read before pdl_wait
    for (int i = threadIdx.x; i < 256; i += kGemmThreads) cta_colsum[2048 + i] = __ldg(ep.lut + i);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_holder, 0);   // warp-uniform for the compiler too
  pdl_wait();   // prologue done (shared memory / TMEM only): global memory written by earlier kernels is touched from here on

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = blockIdx.x; w < num_work; w += gridDim.x) {
        const int n_blk = w % num_n;
        const int m_blk = (w / num_n) % num_m;
        const int split = w / (num_n * num_m);
        const int kb0 = split * kb_per_split;
        const int kb1 = min(kb0 + kb_per_split, num_kb);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * S::kStage;
          uint8_t* sb = sa + S::kStageA;
          mbar_expect_tx(&full_bar[stage], S::kStage);
          if (!A_MN) {
            tma_load_2d(sa, &tma_a, &full_bar[stage], kb * BK, m_blk * BM);
          } else {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j) tma_load_2d(sa + j * 8192, &tma_a, &full_bar[stage], m_blk * BM + j * 64, kb * BK);
          }
          if (!B_MN) {
            tma_load_2d(sb, &tma_b, &full_bar[stage], kb * BK, n_blk * BN);
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j) tma_load_2d(sb + j * 8192, &tma_b, &full_bar[stage], n_blk * BN + j * 64, kb * BK);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (the whole warp walks the loop, one elected lane issues) =====================
    {
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      const uint32_t smem_s = smem_u32(smem);
      for (int w = blockIdx.x; w < num_work; w += gridDim.x, ++it) {
        const int split = w / (num_n * num_m);
        const int kb0 = split * kb_per_split;
        const int kb1 = min(kb0 + kb_per_split, num_kb);
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_s + stage * S::kStage;
          const uint32_t sb = sa + S::kStageA;
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              const uint64_t da = A_MN ? make_sdesc_sw128(sa + k * 2048, 8192, 1024) : make_sdesc_sw128(sa + k * 32, 16, 1024);
              const uint64_t db = B_MN ? make_sdesc_sw128(sb + k * 2048, 8192, 1024) : make_sdesc_sw128(sb + k * 32, 16, 1024);
              tc_mma_ss(d_tmem, da, db, kIdesc, (kb > kb0 || k > 0) ? 1u : 0u);
            }
            tc_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs have read it
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        if (elect_one()) tc_commit(&tmem_full[acc]);  // accumulator complete
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue: 8 warps; TMEM -> regs -> per-warp smem transpose -> coalesced fused epilogue ==========
    // warp ew: TMEM lane quarter (warp & 3), column half (ew >> 2).  After the transpose lane l owns 4 consecutive columns
    // (l & 7) of rows (l >> 3) + 4*i, so bias is one float4 per chunk and every global access covers whole 128-byte rows.
    const int ew = warp - 2;
    const int quarter = warp & 3;
    const int half = ew >> 2;
    float* tile = epi_smem + ew * 1024;
    EpiTmaState st;
    st.stage_s = smem_u32(epi_smem) + (uint32_t)ew * 2u * kStageTileBytes;
    st.ld_bar = epi_ld_bar + 2 * ew;
    st.uses0 = st.uses1 = 0;
    int it = 0;
    for (int w = blockIdx.x; w < num_work; w += gridDim.x, ++it) {
      const int n_blk = w % num_n;
      const int m_blk = (w / num_n) % num_m;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const long long row_base = (long long)m_blk * BM + quarter * 32;
      auto release = [&]() { if (lane == 0) mbar_arrive(&tmem_empty[acc]); };
      if constexpr (TMA_EPI) {
        const uint32_t tw = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN;
        epilogue_warp_tile_tma<BN, MODE, OUT_F32>(ep, &tma_out, &tma_aux, st, tw, n_blk * BN, (int)row_base, N, half, cta_colsum, lane,
                                                  &tmem_full[acc], acc_phase, release);
      } else {
        const uint32_t tw = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN + half * (BN / 2);
        epilogue_warp_tile<BN / 2, MODE, OUT_F32>(ep, tw, n_blk * BN + half * (BN / 2), row_base, M, N, tile, cta_colsum, lane,
                                                  &tmem_full[acc], acc_phase, release);
      }
    }
    if (TMA_EPI && lane == 0) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (epi_is_gelu_bwd(MODE) && ep.colsum != nullptr)  // one global atomic per column per CTA
    for (int i = threadIdx.x; i < N; i += kGemmThreads) atomicAdd(ep.colsum + i, cta_colsum[i]);
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

template <int BN, bool A_MN, bool B_MN, int MODE, bool OUT_F32, bool TMA_EPI>
static int launch_gemm(const dig_gemm_t* g, cudaStream_t stream) {
  using S = GemmSmem<BN, TMA_EPI>;
  CUtensorMap ta, tb;
  int rc;
  if (!g->a_mn_major) rc = make_tmap_bf16_2d(&ta, g->A, (uint64_t)g->M, (uint64_t)g->K, (uint64_t)g->lda, BM, BK);
  else rc = make_tmap_bf16_2d(&ta, g->A, (uint64_t)g->K, (uint64_t)g->M, (uint64_t)g->lda, BK, 64);
  if (rc) return rc;
  if (!g->b_mn_major) rc = make_tmap_bf16_2d(&tb, g->B, (uint64_t)g->N, (uint64_t)g->K, (uint64_t)g->ldb, BN, BK);
  else rc = make_tmap_bf16_2d(&tb, g->B, (uint64_t)g->K, (uint64_t)g->N, (uint64_t)g->ldb, BK, 64);
  if (rc) return rc;
  CUtensorMap to = ta, tx = ta;  // unused by the generic epilogue
  if (TMA_EPI) {
    rc = make_tmap_2d(&to, g->out, OUT_F32 ? 1 : 0, (uint64_t)g->M, (uint64_t)g->N, (uint64_t)g->ldo, 32, OUT_F32 ? 32 : 64);
    if (rc) return rc;
    if (epi_is_q8(MODE)) rc = make_tmap_u8_2d(&tx, g->aux, (uint64_t)g->M, (uint64_t)g->N, (uint64_t)g->ldaux, 32);
    else if ((epi_is_gelu(MODE) && g->aux != nullptr) || epi_is_gelu_bwd(MODE) || MODE == DIG_EPI_ROWDOT) rc = make_tmap_2d(&tx, g->aux, 0, (uint64_t)g->M, (uint64_t)g->N, (uint64_t)g->ldaux, 32, 64);
    else if (MODE == DIG_EPI_LINEAR && OUT_F32 && g->residual) rc = make_tmap_2d(&tx, g->residual, 1, (uint64_t)(g->res_row_mod > 0 ? g->res_row_mod : g->M), (uint64_t)g->N, (uint64_t)g->ldr, 32, 32);
    if (rc) return rc;
  }

  const int num_m = (int)((g->M + BM - 1) / BM), num_n = (int)((g->N + BN - 1) / BN), num_kb = (int)((g->K + BK - 1) / BK);
  int split = g->split_k > 1 ? g->split_k : 1;
  if (split > num_kb) split = num_kb;
  const int per = (num_kb + split - 1) / split;
  split = (num_kb + per - 1) / per;

  GemmEpilogue ep;
  ep.out = g->out; ep.ldo = g->ldo;
  ep.bias = g->bias ? g->bias : (epi_is_gelu(MODE) ? zero_bias() : nullptr); ep.residual = g->residual; ep.ldr = g->ldr; ep.res_row_mod = g->res_row_mod;
  ep.row_mask = g->row_mask; ep.row_mask_value = g->row_mask_value;
  ep.aux = g->aux; ep.ldaux = g->ldaux; ep.alpha = g->alpha;
  ep.colsum = g->colsum;
  ep.rowdot = g->rowdot; ep.ldrowdot = g->ldrowdot; ep.M = (int)g->M;
  ep.lut = epi_is_q8(MODE) ? gelu_grad_lut() : nullptr;
  if (epi_is_q8(MODE)) DIG_REQUIRE(ep.lut != nullptr, "dig_gemm: could not initialise the gelu' table");
  { static int dbg = -1; if (dbg < 0) { const char* e = getenv("DIG_GEMM_DBG"); dbg = e ? atoi(e) : 0; } ep.dbg = dbg; }
  if (g->colsum) DIG_REQUIRE(g->epilogue == DIG_EPI_GELU_BWD && g->N <= 2048, "dig_gemm: colsum is built for DIG_EPI_GELU_BWD with N <= 2048 only");

  auto kern = gemm_bf16_tcgen05<BN, A_MN, B_MN, MODE, OUT_F32, TMA_EPI>;
  static bool attr_set = false;
  if (!attr_set) {
    DIG_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kBytes));
    attr_set = true;
  }
  const long long work = (long long)num_m * num_n * split;
  const int grid = (int)(work < num_sms() ? work : num_sms());
  DIG_CHECK_CUDA(launch_pdl(kern, dim3(grid), dim3(kGemmThreads), S::kBytes, stream, ta, tb, to, tx, ep, (int)g->M, (int)g->N, (int)g->K, split, per));
  return 0;
}

// Instantiated (operand majors x epilogue x output type) combinations -- the ones the pre-training step uses, plus plain
// LINEAR for every major combination.  Anything else is rejected with an error instead of a slow generic path.
template <int BN>
static int dispatch(const dig_gemm_t* g, cudaStream_t s) {
  const bool amn = g->a_mn_major != 0, bmn = g->b_mn_major != 0, f32 = g->out_fp32 != 0;
  int mode = g->split_k > 1 ? kEpiAtomic : g->epilogue;
  if (g->aux_q8 && g->aux != nullptr) mode = (mode == DIG_EPI_GELU) ? kEpiGeluQ8 : kEpiGeluBwdQ8;
  const bool tma_ok = tma_epilogue_ok(g);
#define DIG_CASE(A, B, MODE, F32) \
  if (amn == A && bmn == B && mode == MODE && f32 == F32) return launch_gemm<BN, A, B, MODE, F32, false>(g, s);
#define DIG_CASE_T(A, B, MODE, F32)                                              \
  if (amn == A && bmn == B && mode == MODE && f32 == F32) {                      \
    if (tma_ok) return launch_gemm<BN, A, B, MODE, F32, true>(g, s);                      \
    return launch_gemm<BN, A, B, MODE, F32, false>(g, s);                                 \
  }
  DIG_CASE_T(false, false, DIG_EPI_LINEAR, false)   // forward Linear -> bf16 (qkv, pix_decoder)
  DIG_CASE_T(false, false, DIG_EPI_LINEAR, true)    // forward Linear -> fp32 (+bias +residual, patch embed, BN-MLP heads)
  DIG_CASE_T(false, false, DIG_EPI_GELU, false)     // fc1 + GELU
  DIG_CASE_T(false, false, kEpiGeluQ8, false)       // fc1 + GELU, pre-activation as 8-bit codes
  DIG_CASE_T(false, true, DIG_EPI_LINEAR, false)    // dgrad -> bf16
  DIG_CASE_T(false, true, DIG_EPI_LINEAR, true)     // dgrad -> fp32
  DIG_CASE_T(false, true, DIG_EPI_GELU_BWD, false)  // fc2 dgrad * gelu'
  DIG_CASE_T(false, true, kEpiGeluBwdQ8, false)     // fc2 dgrad * gelu'(8-bit level)
  if (tma_ok) { DIG_CASE_T(false, true, DIG_EPI_ROWDOT, false) }   // proj dgrad + D = rowsum(dO o O) per head (TMA epilogue only)
  DIG_CASE(false, true, DIG_EPI_RELU_MASK, true)  // BN-MLP dgrad through ReLU
  DIG_CASE_T(true, true, DIG_EPI_LINEAR, true)      // wgrad, single pass
  DIG_CASE_T(true, true, kEpiAtomic, true)          // wgrad, split-K
  DIG_CASE(true, true, DIG_EPI_LINEAR, false)
  DIG_CASE(true, false, DIG_EPI_LINEAR, true)
  DIG_CASE(true, false, DIG_EPI_LINEAR, false)
  DIG_CASE(false, false, kEpiAtomic, true)
  DIG_CASE(false, true, kEpiAtomic, true)
  DIG_CASE(true, false, kEpiAtomic, true)
#undef DIG_CASE
#undef DIG_CASE_T
  set_last_error("dig_gemm: combination not built (a_mn=%d b_mn=%d epilogue=%d split_k=%d out_fp32=%d)", (int)amn, (int)bmn, g->epilogue,
                 g->split_k, (int)f32);
  return -1;
}

__device__ float g_zero_bias[kZeroBiasLen];   // zero-initialised
const float* zero_bias() {
  static const float* p = nullptr;
  if (p == nullptr) {
    void* a = nullptr;
    if (cudaGetSymbolAddress(&a, g_zero_bias) == cudaSuccess) p = reinterpret_cast<const float*>(a);
  }
  return p;
}

// gelu_erf'(x) = Phi(x) + x phi(x) at the 256 code levels of dig_gemm_t.aux_q8 (computed in double, uploaded once)
__device__ float g_gelu_grad_lut[256];
const float* gelu_grad_lut() {
  static const float* p = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    float h[256];
    for (int i = 0; i < 256; ++i) {
      const double x = (double)i * (2.0 * kQ8Range / 255.0) - (double)kQ8Range;
      h[i] = (float)(0.5 * (1.0 + erf(x * 0.70710678118654752440)) + x * exp(-0.5 * x * x) * 0.39894228040143267794);
    }
    void* a = nullptr;
    if (cudaMemcpyToSymbol(g_gelu_grad_lut, h, sizeof(h)) == cudaSuccess && cudaGetSymbolAddress(&a, g_gelu_grad_lut) == cudaSuccess)
      p = reinterpret_cast<const float*>(a);
  });
  return p;
}

int gemm2_try(const dig_gemm_t* g, cudaStream_t s);  // gemm2.cu: 2-CTA kernel; returns 1 when it does not take the problem

static bool use_2cta() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DIG_GEMM_1CTA");
    v = (e && e[0] == '1') ? 0 : 1;
  }
  return v == 1;
}

}  // namespace dig

extern "C" int dig_gemm(const dig_gemm_t* g_in, void* stream) {
  using namespace dig;
  DIG_REQUIRE(g_in != nullptr, "dig_gemm: null descriptor");
  dig_gemm_t gg = *g_in;
  const dig_gemm_t* g = &gg;
  DIG_REQUIRE(g->M > 0 && g->N > 0 && g->K > 0, "dig_gemm: empty problem M=%lld N=%lld K=%lld", (long long)g->M, (long long)g->N,
              (long long)g->K);
  DIG_REQUIRE(g->A && g->B && g->out, "dig_gemm: null operand pointer");
  DIG_REQUIRE(g->lda % 8 == 0 && g->ldb % 8 == 0, "dig_gemm: leading dimensions must be multiples of 8 elements (lda=%lld ldb=%lld)",
              (long long)g->lda, (long long)g->ldb);
  DIG_REQUIRE(((uintptr_t)g->A & 15) == 0 && ((uintptr_t)g->B & 15) == 0 && ((uintptr_t)g->out & 15) == 0,
              "dig_gemm: operands must be 16-byte aligned");
  DIG_REQUIRE(g->ldo % 4 == 0 && g->N % 4 == 0, "dig_gemm: N and ldo must be multiples of 4 (N=%lld ldo=%lld)", (long long)g->N, (long long)g->ldo);
  DIG_REQUIRE(!g->residual || g->ldr % 4 == 0, "dig_gemm: ldr must be a multiple of 4");
  if (g->split_k > 1 || g->split_k < 0)
    DIG_REQUIRE(g->out_fp32 && g->epilogue == DIG_EPI_LINEAR && !g->bias && !g->residual && !g->row_mask && !g->colsum,
                "dig_gemm: split_k needs a plain fp32 accumulate epilogue");
  if (g->epilogue == DIG_EPI_GELU && g->bias == nullptr)
    DIG_REQUIRE(g->N <= kZeroBiasLen, "dig_gemm: DIG_EPI_GELU without a bias supports N <= %d", kZeroBiasLen);
  if (g->epilogue == DIG_EPI_ROWDOT)
    DIG_REQUIRE(g->rowdot != nullptr && g->ldrowdot * 64 >= g->N && g->N % 64 == 0 && !g->out_fp32 && g->split_k == 1,
                "dig_gemm: DIG_EPI_ROWDOT needs bf16 out, N %% 64 == 0 and rowdot[M, >= N/64]");
  if (g->epilogue != DIG_EPI_LINEAR && !(g->epilogue == DIG_EPI_GELU && g->aux == nullptr))   // GELU forward: the pre-activation copy is optional
    DIG_REQUIRE(g->aux != nullptr && g->ldaux % 4 == 0, "dig_gemm: epilogue %d needs aux", g->epilogue);
  if (g->aux_q8)
    DIG_REQUIRE((g->epilogue == DIG_EPI_GELU || g->epilogue == DIG_EPI_GELU_BWD) && g->aux != nullptr,
                "dig_gemm: aux_q8 applies to the GELU / GELU' epilogues with an aux tensor only");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (use_2cta()) {
    const int r = gemm2_try(g, s);
    if (r <= 0) return r;
  }
  // tile width: 128 wherever N fills it, 64 for the narrow heads (pix_decoder 192/48)
  const int bn = (g->N % 128 == 0 || g->N > 256) ? 128 : 64;
  if (g->split_k < 0) {  // auto split-K on the 1-CTA kernel: fill the SMs, at least 2 slices (the split epilogue accumulates into out)
    const long long tiles = ((g->M + BM - 1) / BM) * ((g->N + bn - 1) / bn), num_kb = (g->K + BK - 1) / BK;
    long long sp = num_sms() / (tiles > 0 ? tiles : 1);
    if (sp < 2) sp = 2;
    if (sp > num_kb) sp = num_kb;
    gg.split_k = (int)sp;     // a single k-block problem degenerates to a plain store: out must then be zero-initialised (it is: gradients)
  }
  if (bn == 128) return dispatch<128>(g, s);
  return dispatch<64>(g, s);
}

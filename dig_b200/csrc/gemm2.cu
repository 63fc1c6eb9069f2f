// dig_b200 -- 2-CTA (cta_group::2) variant of the persistent tcgen05 GEMM.
//
// A CTA pair (cluster of 2 on one TPC) owns a 256 x BN output tile.  Each CTA stages only ITS half of the operands -- its own 128
// rows of A and BN/2 of B's columns -- and the leader CTA issues one tcgen05.mma.cta_group::2 (M = 256) that reads both halves,
// so every byte fetched over TMA / read from shared memory feeds twice the FLOPs of the 1-CTA 128 x 128 kernel, which is operand-
// bandwidth bound.  Each CTA drains its own 128 accumulator rows from its own TMEM with the shared epilogue.
//
// Barrier protocol (barriers live at identical shared-memory offsets in both CTAs):
//   full[s]        leader's copy only; 2 arrivals (leader: arrive.expect_tx for BOTH CTAs' bytes, peer: remote arrive); both CTAs'
//                  TMA loads complete_tx on it (cta_group::2 loads, peer bit cleared in the barrier address)
//   empty[s]       one per CTA; the leader's tcgen05.commit multicasts the arrival to both
//   tmem_full[a]   one per CTA; multicast commit after the tile's last MMA
//   tmem_empty[a]  leader's copy only; 2 x 8 epilogue warps arrive (peer: remote arrive)
#include <stdlib.h>

#include "gemm_epilogue_tma.cuh"

namespace dig {

static constexpr int BK = 64;
static constexpr int kGemm2Threads = 64 + kEpiWarps * 32;

template <int BN, bool TMA_EPI, int MODE>
struct Gemm2Smem {
  static constexpr int kStageA = 128 * BK * 2;
  static constexpr int kStageB = (BN / 2) * BK * 2;
  static constexpr int kStage = kStageA + kStageB;
  static constexpr int kEpi = TMA_EPI ? kEpiTmaBytes : kEpiWarps * 32 * 32 * 4;
  static constexpr int kColsum = (epi_is_gelu_bwd(MODE)) ? 2048 * 4 + 1024 : 0;   // per-CTA column-sum scratch + the 256-entry gelu' table (8-bit codes): only the GELU' epilogue uses them
  // as many ring stages as fit in 227 KB: the ring holds only ~1 us of MMA work and every tile's operands are first touched from DRAM
  static constexpr int kStages = (232448 - 1024 - 512 - kEpi - kColsum) / kStage > 6 ? 6 : (232448 - 1024 - 512 - kEpi - kColsum) / kStage;
  static constexpr int kBytes = kStages * kStage + kEpi + kColsum + 1024 + 512;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar)),
      "r"(cta)
      : "memory");
}
// TMA load issued by either CTA of the pair; transaction bytes are credited to the LEADER's barrier.
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c_inner, int c_outer) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c_inner), "r"(c_outer)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* holder, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(holder)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_commit_2sm(uint64_t* bar) {  // arrive on `bar` in both CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void tc_mma_ss_2sm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

template <int BN, bool A_MN, bool B_MN, int MODE, bool OUT_F32, bool TMA_EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemm2Threads, 1)
gemm2_bf16_tcgen05(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                   const __grid_constant__ CUtensorMap tma_out, const __grid_constant__ CUtensorMap tma_aux, GemmEpilogue ep, int M, int N,
                   int K, int split_k, int kb_per_split) {
  using S = Gemm2Smem<BN, TMA_EPI, MODE>;
  constexpr int kStages = S::kStages;
  constexpr uint32_t kTmemCols = (2 * BN <= 256) ? 256 : 512;
  constexpr uint32_t kIdesc = make_idesc_bf16(256, BN, A_MN, B_MN);
  constexpr int HB = BN / 2;  // B columns staged by each CTA

  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* epi_smem = reinterpret_cast<float*>(smem + kStages * S::kStage);
  float* cta_colsum = reinterpret_cast<float*>(smem + kStages * S::kStage + S::kEpi);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * S::kStage + S::kEpi + S::kColsum);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kStages;
  uint64_t* tmem_full = bars + 2 * kStages;
  uint64_t* tmem_empty = bars + 2 * kStages + 2;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);
  uint64_t* epi_ld_bar = bars + 2 * kStages + 5;  // [kEpiWarps][2] TMA loads into the epilogue staging tiles

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;

  const int num_m = (M + 255) / 256;
  const int num_n = (N + BN - 1) / BN;
  const int num_kb = (K + BK - 1) / BK;
  const int num_work = num_m * num_n * split_k;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    if (TMA_EPI) { tma_prefetch_desc(&tma_out); tma_prefetch_desc(&tma_aux); }
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 2);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 2 * kEpiWarps);
    }
    if (TMA_EPI)
      for (int i = 0; i < 2 * kEpiWarps; ++i) mbar_init(&epi_ld_bar[i], 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc_2sm(tmem_holder, kTmemCols);
  if (epi_is_gelu_bwd(MODE) && ep.colsum != nullptr)
    for (int i = threadIdx.x; i < N; i += kGemm2Threads) cta_colsum[i] = 0.f;
  if (epi_is_gelu_bwd(MODE) && epi_is_q8(MODE))   // constant table (written once at library initialisation): safe to read before pdl_wait
    for (int i = threadIdx.x; i < 256; i += kGemm2Threads) cta_colsum[2048 + i] = __ldg(ep.lut + i);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // peer barriers are initialised before anyone arrives on them remotely
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_holder, 0);   // warp-uniform for the compiler too
  pdl_wait();   // everything above touched only this CTA's shared memory / TMEM; from here on global memory of earlier kernels is read

  if (warp == 0) {
    // ===================== TMA producer (both CTAs, own halves; the whole warp walks the loop, one elected lane issues) ==========
    // Activation operands stream from DRAM (each tile is first touched here), and the shared-memory ring holds only ~1 us of MMA work:
    // an L2 prefetch iterator runs kPfDist k-blocks (possibly one work item) ahead of the loads so those find their lines in L2.
    {
      constexpr bool kPrefetchB = A_MN && B_MN;   // weight gradients: both operands are activations; otherwise B = weights (L2-resident)
      const int pf_dist = ep.dbg >= 10 ? ep.dbg - 10 : 0;   // off by default: measured 10-45 % SLOWER with it (scripts/gemm_ab.py, DIG_GEMM_DBG=10+distance)
      int pw = cluster_id, pkb = 0, pkb1 = 0;
      auto pf_set = [&]() {
        if (pw < num_work) {
          const int split = pw / (num_n * num_m);
          pkb = split * kb_per_split;
          pkb1 = min(pkb + kb_per_split, num_kb);
        }
      };
      auto pf_issue = [&]() {
        if (pw >= num_work) return;
        const int m0 = ((pw / num_n) % num_m) * 256 + (int)rank * 128;
        const int n0 = (pw % num_n) * BN + (int)rank * HB;
        if (elect_one()) {
          if (!A_MN) tma_prefetch_l2_2d(&tma_a, pkb * BK, m0);
          else {
#pragma unroll
            for (int j = 0; j < 2; ++j) tma_prefetch_l2_2d(&tma_a, m0 + j * 64, pkb * BK);
          }
          if (kPrefetchB) {
#pragma unroll
            for (int j = 0; j < HB / 64; ++j) tma_prefetch_l2_2d(&tma_b, n0 + j * 64, pkb * BK);
          }
        }
        __syncwarp();
        if (++pkb == pkb1) { pw += num_clusters; pf_set(); }
      };
      pf_set();
      for (int i = 0; i < pf_dist; ++i) pf_issue();
      int stage = 0;
      uint32_t phase = 0;
      for (int w = cluster_id; w < num_work; w += num_clusters) {
        const int n_blk = w % num_n;
        const int m_blk = (w / num_n) % num_m;
        const int split = w / (num_n * num_m);
        const int kb0 = split * kb_per_split;
        const int kb1 = min(kb0 + kb_per_split, num_kb);
        const int m0 = m_blk * 256 + (int)rank * 128;
        const int n0 = n_blk * BN + (int)rank * HB;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * S::kStage;
          uint8_t* sb = sa + S::kStageA;
          if (elect_one()) {
            if (leader) mbar_expect_tx(&full_bar[stage], 2 * S::kStage);
            else mbar_arrive_remote(&full_bar[stage], 0);
            if (!A_MN) {
              tma_load_2d_2sm(sa, &tma_a, &full_bar[stage], kb * BK, m0);
            } else {
#pragma unroll
              for (int j = 0; j < 2; ++j) tma_load_2d_2sm(sa + j * 8192, &tma_a, &full_bar[stage], m0 + j * 64, kb * BK);
            }
            if (!B_MN) {
              tma_load_2d_2sm(sb, &tma_b, &full_bar[stage], kb * BK, n0);
            } else {
#pragma unroll
              for (int j = 0; j < HB / 64; ++j) tma_load_2d_2sm(sb + j * 8192, &tma_b, &full_bar[stage], n0 + j * 64, kb * BK);
            }
          }
          __syncwarp();
          if (pf_dist > 0) pf_issue();
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA; the whole warp walks the loop, one elected lane issues) =====================
    if (leader) {
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      const uint32_t smem_s = smem_u32(smem);
      for (int w = cluster_id; w < num_work; w += num_clusters, ++it) {
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
              tc_mma_ss_2sm(d_tmem, da, db, kIdesc, (kb > kb0 || k > 0) ? 1u : 0u);
            }
            tc_commit_2sm(&empty_bar[stage]);
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        if (elect_one()) tc_commit_2sm(&tmem_full[acc]);
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue (both CTAs, own 128 rows) =====================
    const int ew = warp - 2;
    const int quarter = warp & 3;
    const int half = ew >> 2;
    float* tile = epi_smem + ew * 1024;
    EpiTmaState st;
    st.stage_s = smem_u32(epi_smem) + (uint32_t)ew * 2u * kStageTileBytes;
    st.ld_bar = epi_ld_bar + 2 * ew;
    st.uses0 = st.uses1 = 0;
    int it = 0;
    for (int w = cluster_id; w < num_work; w += num_clusters, ++it) {
      const int n_blk = w % num_n;
      const int m_blk = (w / num_n) % num_m;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const long long row_base = (long long)m_blk * 256 + rank * 128 + quarter * 32;
      auto release = [&]() {
        if (lane == 0) {
          if (leader) mbar_arrive(&tmem_empty[acc]);
          else mbar_arrive_remote(&tmem_empty[acc], 0);
        }
      };
      if constexpr (TMA_EPI) {
        const uint32_t tw = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN;
        epilogue_warp_tile_tma<BN, MODE, OUT_F32>(ep, &tma_out, &tma_aux, st, tw, n_blk * BN, (int)row_base, N, half, cta_colsum, lane,
                                                  &tmem_full[acc], acc_phase, release);
      } else {
        const uint32_t tw = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN + half * HB;
        epilogue_warp_tile<HB, MODE, OUT_F32>(ep, tw, n_blk * BN + half * HB, row_base, M, N, tile, cta_colsum, lane, &tmem_full[acc],
                                              acc_phase, release);
      }
    }
    if (TMA_EPI && lane == 0) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the pair leaves together: the leader's MMAs read the peer's shared memory, arrivals are remote
  if (epi_is_gelu_bwd(MODE) && ep.colsum != nullptr)
    for (int i = threadIdx.x; i < N; i += kGemm2Threads) atomicAdd(ep.colsum + i, cta_colsum[i]);
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, kTmemCols);
  }
}

template <int BN, bool A_MN, bool B_MN, int MODE, bool OUT_F32, bool TMA_EPI>
static int launch_gemm2(const dig_gemm_t* g, cudaStream_t stream) {
  using S = Gemm2Smem<BN, TMA_EPI, MODE>;
  CUtensorMap ta, tb;
  int rc;
  if (!g->a_mn_major) rc = make_tmap_bf16_2d(&ta, g->A, (uint64_t)g->M, (uint64_t)g->K, (uint64_t)g->lda, 128, BK);
  else rc = make_tmap_bf16_2d(&ta, g->A, (uint64_t)g->K, (uint64_t)g->M, (uint64_t)g->lda, BK, 64);
  if (rc) return rc;
  if (!g->b_mn_major) rc = make_tmap_bf16_2d(&tb, g->B, (uint64_t)g->N, (uint64_t)g->K, (uint64_t)g->ldb, BN / 2, BK);
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

  const int num_m = (int)((g->M + 255) / 256), num_n = (int)((g->N + BN - 1) / BN), num_kb = (int)((g->K + BK - 1) / BK);
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

  auto kern = gemm2_bf16_tcgen05<BN, A_MN, B_MN, MODE, OUT_F32, TMA_EPI>;
  static bool attr_set = false;
  if (!attr_set) {
    DIG_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kBytes));
    attr_set = true;
  }
  const long long work = (long long)num_m * num_n * split;
  const int max_clusters = num_sms() / 2;
  const int clusters = (int)(work < max_clusters ? work : max_clusters);
  DIG_CHECK_CUDA(launch_pdl(kern, dim3(clusters * 2), dim3(kGemm2Threads), S::kBytes, stream, ta, tb, to, tx, ep, (int)g->M, (int)g->N, (int)g->K, split, per));
  return 0;
}

// Returns 1 when the combination is not built for the 2-CTA kernel (caller falls back to the 1-CTA kernel), 0 on success, <0 on error.
template <int BN>
static int dispatch2(const dig_gemm_t* g, cudaStream_t s) {
  const bool amn = g->a_mn_major != 0, bmn = g->b_mn_major != 0, f32 = g->out_fp32 != 0;
  int mode = g->split_k > 1 ? kEpiAtomic : g->epilogue;
  if (g->aux_q8 && g->aux != nullptr) mode = (mode == DIG_EPI_GELU) ? kEpiGeluQ8 : kEpiGeluBwdQ8;
  const bool tma_ok = tma_epilogue_ok(g);
#define DIG_CASE(A, B, MODE, F32) \
  if (amn == A && bmn == B && mode == MODE && f32 == F32) return launch_gemm2<BN, A, B, MODE, F32, false>(g, s);
#define DIG_CASE_T(A, B, MODE, F32)                                              \
  if (amn == A && bmn == B && mode == MODE && f32 == F32) {                      \
    if (tma_ok) return launch_gemm2<BN, A, B, MODE, F32, true>(g, s);                      \
    return launch_gemm2<BN, A, B, MODE, F32, false>(g, s);                                 \
  }
  DIG_CASE_T(false, false, DIG_EPI_LINEAR, false)
  DIG_CASE_T(false, false, DIG_EPI_LINEAR, true)
  DIG_CASE_T(false, false, DIG_EPI_GELU, false)
  DIG_CASE_T(false, false, kEpiGeluQ8, false)
  if constexpr (BN != 192) {  // MN-major B is staged in 64-column boxes: BN/2 must be a multiple of 64
    DIG_CASE_T(false, true, DIG_EPI_LINEAR, false)
    DIG_CASE_T(false, true, DIG_EPI_LINEAR, true)
    DIG_CASE_T(false, true, DIG_EPI_GELU_BWD, false)
    DIG_CASE_T(false, true, kEpiGeluBwdQ8, false)
    if (tma_ok) { DIG_CASE_T(false, true, DIG_EPI_ROWDOT, false) }
    DIG_CASE(false, true, DIG_EPI_RELU_MASK, true)
    DIG_CASE_T(true, true, DIG_EPI_LINEAR, true)
    DIG_CASE_T(true, true, kEpiAtomic, true)
  }
#undef DIG_CASE
#undef DIG_CASE_T
  return 1;
}

int gemm2_try(const dig_gemm_t* g_in, cudaStream_t s) {
  // 2-CTA tiles are 256 rows tall: keep small problems (few tiles) on the 1-CTA kernel so they still spread over the SMs
  dig_gemm_t gg = *g_in;
  const dig_gemm_t* g = &gg;
  const long long m_tiles = (g->M + 255) / 256;
  const bool bmn = g->b_mn_major != 0;
  const bool accumulate = g->split_k > 1 || g->split_k < 0;
  static int wgrad_bn256 = -1, splitk_2cta = -1;
  if (wgrad_bn256 < 0) { const char* e = getenv("DIG_GEMM_WGRAD_BN256"); wgrad_bn256 = e ? atoi(e) : 1; }
  if (splitk_2cta < 0) { const char* e = getenv("DIG_GEMM_2CTA_SPLITK"); splitk_2cta = e ? atoi(e) : 1; }
  if (accumulate && (!splitk_2cta || g->M * g->N <= 384 * 384)) return 1;   // small outputs: the 1-CTA kernel's 128-wide tiles waste less
  int bn;
  if (g->N % 256 == 0) bn = 256;
  else if (accumulate && wgrad_bn256 && g->N > 256) bn = 256;   // split-K accumulate: ragged last tile, clipped by the tensor maps
  else if (g->N % 192 == 0 && !bmn) bn = 192;
  else if (g->N % 128 == 0) bn = 128;
  else return 1;
  const long long tiles = m_tiles * ((g->N + bn - 1) / bn);
  if (g->split_k < 0) {  // auto: as many K slices as fill the 74 CTA pairs (at least 2, so the epilogue accumulates)
    const long long num_kb = (g->K + BK - 1) / BK;
    long long sp = (num_sms() / 2) / (tiles > 0 ? tiles : 1);
    if (sp < 2) sp = 2;
    if (sp > num_kb) sp = num_kb;
    if (sp < 2) return 1;
    gg.split_k = (int)sp;
  }
  const long long split = g->split_k > 1 ? g->split_k : 1;
  if (tiles * split < 32) return 1;
  // Measured on B200 (scripts/gemm_ab.py): with the TMA-staged epilogue the 2-CTA kernel wins or ties everywhere (erf-GELU forward:
  // 110 us vs 130 us at 65536 x 1536 x 384 once the epilogue arithmetic is packed FFMA2).  Split-K accumulation runs on 256 x 256
  // pair tiles as well: the 128 x 128 1-CTA tile is shared-memory-bandwidth bound (TMA fill + UMMA operand reads of 250 B/clk against
  // 128 B/clk/SM), the pair tile needs half of that per FLOP.  DIG_GEMM_2CTA_MINK raises the K threshold, DIG_GEMM_2CTA_GELU=0 /
  // DIG_GEMM_2CTA_SPLITK=0 / DIG_GEMM_WGRAD_BN256=0 switch the respective choices off (experiments).
  static int min_k = -1;
  if (min_k < 0) { const char* e = getenv("DIG_GEMM_2CTA_MINK"); min_k = e ? atoi(e) : 0; }
  static int gelu_2cta = -1;
  if (gelu_2cta < 0) { const char* e = getenv("DIG_GEMM_2CTA_GELU"); gelu_2cta = e ? atoi(e) : 1; }
  if (g->K < min_k || (g->epilogue == DIG_EPI_GELU && !gelu_2cta)) return 1;
  if (bn == 256) return dispatch2<256>(g, s);
  if (bn == 192) return dispatch2<192>(g, s);
  return dispatch2<128>(g, s);
}

}  // namespace dig

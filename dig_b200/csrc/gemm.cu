// dig_b200 -- persistent warp-specialised tcgen05 GEMM for sm_100a.
//
//   out[M,N] = epilogue( alpha * A[M,K] . B[N,K]^T ),  bf16 operands, fp32 accumulation in TMEM.
//
// One CTA per SM walks output tiles of 128 x BN.  Warp 0 feeds a ring of shared-memory stages with TMA
// (128-byte swizzle), warp 1 issues tcgen05.mma from one thread and owns the TMEM allocation, warps 2..5
// drain the double-buffered TMEM accumulator (one output row per thread) and apply the fused epilogue, so
// tile i's epilogue overlaps tile i+1's MMAs.  Either operand may be K-major or MN-major (see
// include/dig_b200.h), which covers forward (x.W^T), dgrad (dy.W) and wgrad (dy^T.x) without any transposed
// copies.  Replaces the cuBLAS calls behind F.linear in the reference (modeling_finetune.py:93,119,54,58;
// modeling_pretrain_moco_mim_ori.py:463-482,422-426) and their autograd counterparts.
#include "common.cuh"
#include "../../include/dig_b200.h"

namespace dig {

static constexpr int BM = 128;
static constexpr int BK = 64;
static constexpr int kGemmThreads = 192;

struct GemmEpilogue {
  void* out;
  long long ldo;
  int out_fp32;
  const float* bias;
  const float* residual;
  long long ldr;
  long long res_row_mod;
  const uint8_t* row_mask;
  const float* row_mask_value;
  int mode;
  void* aux;
  long long ldaux;
  float alpha;
  int atomic;
};

template <int BN>
struct GemmSmem {
  static constexpr int kStageA = BM * BK * 2;
  static constexpr int kStageB = BN * BK * 2;
  static constexpr int kStage = kStageA + kStageB;
  static constexpr int kStages = (BN == 256) ? 4 : ((BN == 128) ? 6 : 8);
  static constexpr int kBytes = kStages * kStage + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_tcgen05(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, GemmEpilogue ep, int M,
                  int N, int K, int split_k, int kb_per_split) {
  using S = GemmSmem<BN>;
  constexpr int kStages = S::kStages;
  constexpr uint32_t kTmemCols = (2 * BN <= 32) ? 32 : (2 * BN <= 64 ? 64 : (2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512)));
  constexpr uint32_t kIdesc = make_idesc_bf16(BM, BN, A_MN, B_MN);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * S::kStage);
  uint64_t* full_bar = bars;                   // [kStages]  TMA -> MMA
  uint64_t* empty_bar = bars + kStages;        // [kStages]  MMA -> TMA
  uint64_t* tmem_full = bars + 2 * kStages;    // [2]        MMA -> epilogue
  uint64_t* tmem_empty = bars + 2 * kStages + 2;  // [2]     epilogue -> MMA
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int num_m = (M + BM - 1) / BM;
  const int num_n = (N + BN - 1) / BN;
  const int num_kb = (K + BK - 1) / BK;
  const int num_work = num_m * num_n * split_k;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 128);
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_holder, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

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
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
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
          const uint32_t sa = smem_u32(smem + stage * S::kStage);
          const uint32_t sb = sa + S::kStageA;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t da = A_MN ? make_sdesc_sw128(sa + k * 2048, 8192, 1024) : make_sdesc_sw128(sa + k * 32, 16, 1024);
            const uint64_t db = B_MN ? make_sdesc_sw128(sb + k * 2048, 8192, 1024) : make_sdesc_sw128(sb + k * 32, 16, 1024);
            tc_mma_ss(d_tmem, da, db, kIdesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          tc_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs have read it
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        tc_commit(&tmem_full[acc]);  // accumulator complete
      }
    }
  } else {
    // ===================== epilogue warps (TMEM -> registers -> global) =====================
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    int it = 0;
    for (int w = blockIdx.x; w < num_work; w += gridDim.x, ++it) {
      const int n_blk = w % num_n;
      const int m_blk = (w / num_n) % num_m;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const long long row = (long long)m_blk * BM + quarter * 32 + lane;
      const bool row_ok = row < M;
      const bool masked = row_ok && ep.row_mask != nullptr && ep.row_mask[row] != 0;
      const float* res_row = nullptr;
      if (ep.residual != nullptr && row_ok)
        res_row = ep.residual + (ep.res_row_mod > 0 ? (row % ep.res_row_mod) : row) * ep.ldr;
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN + c, v);
        tmem_ld_wait();
        const int col0 = n_blk * BN + c;
        if (row_ok && col0 < N) {
          const bool full = (col0 + 32 <= N);
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]) * ep.alpha;
          if (ep.atomic) {
            float* o = reinterpret_cast<float*>(ep.out) + row * ep.ldo + col0;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (full || col0 + j < N) atomicAdd(o + j, f[j]);
            continue;
          }
          if (ep.bias != nullptr) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (full || col0 + j < N) f[j] += __ldg(ep.bias + col0 + j);
          }
          if (ep.mode == DIG_EPI_GELU) {
            __nv_bfloat16* a = reinterpret_cast<__nv_bfloat16*>(ep.aux) + row * ep.ldaux + col0;
            if (full) {
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                uint4 p;
                p.x = pack_bf16(f[j + 0], f[j + 1]);
                p.y = pack_bf16(f[j + 2], f[j + 3]);
                p.z = pack_bf16(f[j + 4], f[j + 5]);
                p.w = pack_bf16(f[j + 6], f[j + 7]);
                *reinterpret_cast<uint4*>(a + j) = p;
              }
            } else {
              for (int j = 0; j < 32 && col0 + j < N; ++j) a[j] = __float2bfloat16(f[j]);
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = gelu_erf(f[j]);
          } else if (ep.mode == DIG_EPI_GELU_BWD || ep.mode == DIG_EPI_RELU_MASK) {
            const __nv_bfloat16* a = reinterpret_cast<const __nv_bfloat16*>(ep.aux) + row * ep.ldaux + col0;
            float x[32];
            if (full) {
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                const uint4 p = *reinterpret_cast<const uint4*>(a + j);
                x[j + 0] = bf16_lo(p.x); x[j + 1] = bf16_hi(p.x);
                x[j + 2] = bf16_lo(p.y); x[j + 3] = bf16_hi(p.y);
                x[j + 4] = bf16_lo(p.z); x[j + 5] = bf16_hi(p.z);
                x[j + 6] = bf16_lo(p.w); x[j + 7] = bf16_hi(p.w);
              }
            } else {
              for (int j = 0; j < 32; ++j) x[j] = (col0 + j < N) ? __bfloat162float(a[j]) : 0.f;
            }
            if (ep.mode == DIG_EPI_GELU_BWD) {
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] *= gelu_erf_grad(x[j]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] = x[j] > 0.f ? f[j] : 0.f;
            }
          }
          if (masked) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (full || col0 + j < N) f[j] = __ldg(ep.row_mask_value + col0 + j);
          }
          if (res_row != nullptr) {
            if (full) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 r = *reinterpret_cast<const float4*>(res_row + col0 + j);
                f[j] += r.x; f[j + 1] += r.y; f[j + 2] += r.z; f[j + 3] += r.w;
              }
            } else {
              for (int j = 0; j < 32 && col0 + j < N; ++j) f[j] += res_row[col0 + j];
            }
          }
          if (ep.out_fp32) {
            float* o = reinterpret_cast<float*>(ep.out) + row * ep.ldo + col0;
            if (full) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
            } else {
              for (int j = 0; j < 32 && col0 + j < N; ++j) o[j] = f[j];
            }
          } else {
            __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(ep.out) + row * ep.ldo + col0;
            if (full) {
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                uint4 p;
                p.x = pack_bf16(f[j + 0], f[j + 1]);
                p.y = pack_bf16(f[j + 2], f[j + 3]);
                p.z = pack_bf16(f[j + 4], f[j + 5]);
                p.w = pack_bf16(f[j + 6], f[j + 7]);
                *reinterpret_cast<uint4*>(o + j) = p;
              }
            } else {
              for (int j = 0; j < 32 && col0 + j < N; ++j) o[j] = __float2bfloat16(f[j]);
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

template <int BN, bool A_MN, bool B_MN>
static int launch_gemm(const dig_gemm_t* g, cudaStream_t stream) {
  using S = GemmSmem<BN>;
  CUtensorMap ta, tb;
  int rc;
  if (!g->a_mn_major) rc = make_tmap_bf16_2d(&ta, g->A, (uint64_t)g->M, (uint64_t)g->K, (uint64_t)g->lda, BM, BK);
  else rc = make_tmap_bf16_2d(&ta, g->A, (uint64_t)g->K, (uint64_t)g->M, (uint64_t)g->lda, BK, 64);
  if (rc) return rc;
  if (!g->b_mn_major) rc = make_tmap_bf16_2d(&tb, g->B, (uint64_t)g->N, (uint64_t)g->K, (uint64_t)g->ldb, BN, BK);
  else rc = make_tmap_bf16_2d(&tb, g->B, (uint64_t)g->K, (uint64_t)g->N, (uint64_t)g->ldb, BK, 64);
  if (rc) return rc;

  const int num_m = (int)((g->M + BM - 1) / BM), num_n = (int)((g->N + BN - 1) / BN), num_kb = (int)((g->K + BK - 1) / BK);
  int split = g->split_k > 1 ? g->split_k : 1;
  if (split > num_kb) split = num_kb;
  const int per = (num_kb + split - 1) / split;
  split = (num_kb + per - 1) / per;

  GemmEpilogue ep;
  ep.out = g->out; ep.ldo = g->ldo; ep.out_fp32 = g->out_fp32;
  ep.bias = g->bias; ep.residual = g->residual; ep.ldr = g->ldr; ep.res_row_mod = g->res_row_mod;
  ep.row_mask = g->row_mask; ep.row_mask_value = g->row_mask_value;
  ep.mode = g->epilogue; ep.aux = g->aux; ep.ldaux = g->ldaux; ep.alpha = g->alpha;
  ep.atomic = g->split_k > 1 ? 1 : 0;

  auto kern = gemm_bf16_tcgen05<BN, A_MN, B_MN>;
  static bool attr_set = false;
  if (!attr_set) {
    DIG_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kBytes));
    attr_set = true;
  }
  const long long work = (long long)num_m * num_n * split;
  const int grid = (int)(work < num_sms() ? work : num_sms());
  kern<<<grid, kGemmThreads, S::kBytes, stream>>>(ta, tb, ep, (int)g->M, (int)g->N, (int)g->K, split, per);
  DIG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

template <int BN>
static int dispatch_major(const dig_gemm_t* g, cudaStream_t s) {
  if (!g->a_mn_major && !g->b_mn_major) return launch_gemm<BN, false, false>(g, s);
  if (!g->a_mn_major && g->b_mn_major) return launch_gemm<BN, false, true>(g, s);
  if (g->a_mn_major && !g->b_mn_major) return launch_gemm<BN, true, false>(g, s);
  return launch_gemm<BN, true, true>(g, s);
}

}  // namespace dig

extern "C" int dig_gemm(const dig_gemm_t* g, void* stream) {
  using namespace dig;
  DIG_REQUIRE(g != nullptr, "dig_gemm: null descriptor");
  DIG_REQUIRE(g->M > 0 && g->N > 0 && g->K > 0, "dig_gemm: empty problem M=%lld N=%lld K=%lld", (long long)g->M, (long long)g->N,
              (long long)g->K);
  DIG_REQUIRE(g->A && g->B && g->out, "dig_gemm: null operand pointer");
  DIG_REQUIRE(g->lda % 8 == 0 && g->ldb % 8 == 0, "dig_gemm: leading dimensions must be multiples of 8 elements (lda=%lld ldb=%lld)",
              (long long)g->lda, (long long)g->ldb);
  DIG_REQUIRE(((uintptr_t)g->A & 15) == 0 && ((uintptr_t)g->B & 15) == 0 && ((uintptr_t)g->out & 15) == 0,
              "dig_gemm: operands must be 16-byte aligned");
  DIG_REQUIRE(g->ldo % 8 == 0, "dig_gemm: ldo must be a multiple of 8");
  if (g->split_k > 1)
    DIG_REQUIRE(g->out_fp32 && g->epilogue == DIG_EPI_LINEAR && !g->bias && !g->residual && !g->row_mask,
                "dig_gemm: split_k>1 needs a plain fp32 accumulate epilogue");
  if (g->epilogue != DIG_EPI_LINEAR) DIG_REQUIRE(g->aux != nullptr && g->ldaux % 8 == 0, "dig_gemm: epilogue %d needs aux", g->epilogue);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  // tile width: 128 wherever N fills it, 64 for the narrow heads (pix_decoder 192/48)
  if (g->N % 128 == 0 || g->N > 256) return dispatch_major<128>(g, s);
  return dispatch_major<64>(g, s);
}

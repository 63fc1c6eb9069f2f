// dig_b200 -- shared device helpers for the sm_100a kernels (inline PTX: mbarrier, TMA, tcgen05, TMEM).
// Everything here is written against the PTX ISA for sm_100a; no CUTLASS/CuTe types are used.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace dig {

// ------------------------------------------------------------------------------------------------
// error plumbing (host)
// ------------------------------------------------------------------------------------------------
void set_last_error(const char* fmt, ...);
#define DIG_CHECK_CUDA(expr)                                                                     \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess) {                                                                     \
      dig::set_last_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return -2;                                                                                 \
    }                                                                                            \
  } while (0)
#define DIG_REQUIRE(cond, ...)          \
  do {                                  \
    if (!(cond)) {                      \
      dig::set_last_error(__VA_ARGS__); \
      return -1;                        \
    }                                   \
  } while (0)

// Tensor-map helper: 2-D bf16 row-major tensor [rows, cols] (cols contiguous), box {box_cols, box_rows},
// 128-byte swizzle.  Returns 0 on success.
int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t row_stride_elems,
                      uint32_t box_rows, uint32_t box_cols);
int make_tmap_2d(CUtensorMap* out, const void* base, int is_fp32, uint64_t rows, uint64_t cols, uint64_t row_stride_elems,
                 uint32_t box_rows, uint32_t box_cols);
// uint8 [rows, cols], box {64 bytes, box_rows}, 64-byte swizzle (8-bit GELU pre-activation codes)
int make_tmap_u8_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t row_stride_bytes, uint32_t box_rows);
int num_sms();

bool pdl_enabled();   // DIG_PDL (default on): launch with programmatic stream serialization

#ifdef __CUDACC__
// ------------------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL).  A kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may become resident as
// soon as every CTA of the kernel before it in the stream has executed pdl_launch_dependents() (or exited), so its launch latency and
// prologue -- barrier initialisation, TMEM allocation, descriptor prefetch: 1.5-3 us per launch, ~480 launches per step -- overlap the
// tail of its predecessor.  pdl_wait() blocks until ALL prerequisite grids have completed and flushed their memory; every kernel calls
// it before its first access to global memory, so the stream-order semantics are unchanged.  Both are no-ops in a plain launch.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// Launch helper: kern<<<grid, block, smem, stream>>>(args...) with the PDL attribute when enabled.
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// ------------------------------------------------------------------------------------------------
// small device utilities
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }

// erf(x) ~= x P(x^2) / Q(x^2) on |x| <= 3.925 (clamped beyond), max abs error 3.9e-7 in fp32: one MUFU (rcp) + 12 FMA.
// Coefficients fitted against scipy.special.erf (all positive, Q >= 1, so no cancellation or poles).
__device__ __forceinline__ float erf_fast(float x) {
  x = fminf(fmaxf(x, -3.925f), 3.925f);
  const float t = x * x;
  float p = 2.057485900e-06f;
  p = fmaf(p, t, 2.862000669e-04f);
  p = fmaf(p, t, 3.786277615e-03f);
  p = fmaf(p, t, 5.280982382e-02f);
  p = fmaf(p, t, 1.907734496e-01f);
  p = fmaf(p, t, 1.128379076e+00f);
  float q = 3.823944817e-05f;
  q = fmaf(q, t, 1.169261335e-03f);
  q = fmaf(q, t, 1.500781038e-02f);
  q = fmaf(q, t, 1.142729919e-01f);
  q = fmaf(q, t, 5.024007393e-01f);
  q = fmaf(q, t, 1.0f);
  return __fdividef(x * p, q);
}
__device__ __forceinline__ float rcp_approx(float x) {  // one MUFU.RCP, no range fix-up (denominators below are in [1, 1e3])
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// exact-form (erf) GELU and its derivative, fp32 (nn.GELU, F:55).
// gelu(x) = x (0.5 + xc P(xc^2) / Q(xc^2)), xc = clamp(x): a rational fit of (Phi(x) - 1/2)/x in x^2 with the 1/sqrt(2) argument scale and
// the factor 0.5 folded into the coefficients (Lawson-weighted least-squares fits against scipy.special.erf, all Q coefficients positive).
//   P5/Q5, |x| <= 5.55: max abs error 3.9e-7 -- the scalar functions below (row kernels, generic epilogue): 10 FMA + 1 MUFU per element.
//   P3/Q3, |x| <= 4.75: max abs error of gelu 5.4e-6 -- the packed functions of the TMA GEMM epilogues (DIG_GELU_ORDER 3, default): 6 FMA +
//   1 MUFU.  Those outputs are rounded to bf16 (relative 2^-9 = 2e-3), three orders of magnitude above the fit error, and the epilogues
//   are FMA-pipe / issue bound, so the polynomial degree is step time.  DIG_GELU_ORDER 5 selects P5/Q5 there too (A/B switch).
#ifndef DIG_GELU_ORDER
#define DIG_GELU_ORDER 3
#endif
#define DIG_GELU_CLAMP 5.5507882f
#define DIG_GELU_P0 3.989422482e-01f
#define DIG_GELU_P1 3.372429997e-02f
#define DIG_GELU_P2 4.667773067e-03f
#define DIG_GELU_P3 1.673314111e-04f
#define DIG_GELU_P4 6.324187753e-06f
#define DIG_GELU_P5 2.273222238e-08f
#define DIG_GELU_Q1 2.512003696e-01f
#define DIG_GELU_Q2 2.856824797e-02f
#define DIG_GELU_Q3 1.875976298e-03f
#define DIG_GELU_Q4 7.307883344e-05f
#define DIG_GELU_Q5 1.194982755e-06f
#define DIG_GELU3_CLAMP 4.75f
#define DIG_GELU3_P0 3.989111871e-01f
#define DIG_GELU3_P1 2.820760287e-02f
#define DIG_GELU3_P2 3.793992200e-03f
#define DIG_GELU3_P3 3.181064343e-05f
#define DIG_GELU3_Q1 2.372292233e-01f
#define DIG_GELU3_Q2 2.412535149e-02f
#define DIG_GELU3_Q3 1.133673430e-03f
__device__ __forceinline__ float gelu_erf(float x) {
  const float xc = fminf(fmaxf(x, -DIG_GELU_CLAMP), DIG_GELU_CLAMP);
  const float t = xc * xc;
  float p = DIG_GELU_P5;
  p = fmaf(p, t, DIG_GELU_P4);
  p = fmaf(p, t, DIG_GELU_P3);
  p = fmaf(p, t, DIG_GELU_P2);
  p = fmaf(p, t, DIG_GELU_P1);
  p = fmaf(p, t, DIG_GELU_P0);
  float q = DIG_GELU_Q5;
  q = fmaf(q, t, DIG_GELU_Q4);
  q = fmaf(q, t, DIG_GELU_Q3);
  q = fmaf(q, t, DIG_GELU_Q2);
  q = fmaf(q, t, DIG_GELU_Q1);
  q = fmaf(q, t, 1.0f);
  return x * fmaf(xc * p, rcp_approx(q), 0.5f);
}
// d/dx gelu_erf(x) = Phi(x) + x phi(x) = 0.5 + x P(x^2) / Q(x^2) on the clamped range (the odd part saturates at 0.5 beyond):
//   P5/Q5, |x| <= 6: max abs error 3.9e-7 (scalar);  P3/Q3, |x| <= 4.5: 5.7e-5 (packed; the product dy * gelu' is rounded to bf16).
#define DIG_GELUG_CLAMP 6.0f
#define DIG_GELUG_P0 7.978851765e-01f
#define DIG_GELUG_P1 -3.087451237e-02f
#define DIG_GELUG_P2 1.421916872e-02f
#define DIG_GELUG_P3 2.167418626e-05f
#define DIG_GELUG_P4 3.587825389e-05f
#define DIG_GELUG_P5 1.916786770e-07f
#define DIG_GELUG_Q1 2.946448083e-01f
#define DIG_GELUG_Q2 4.101880957e-02f
#define DIG_GELUG_Q3 3.525759290e-03f
#define DIG_GELUG_Q4 1.926611686e-04f
#define DIG_GELUG_Q5 8.911869432e-06f
#define DIG_GELUG3_CLAMP 4.5f
#define DIG_GELUG3_P0 7.982111506e-01f
#define DIG_GELUG3_P1 -2.699143690e-02f
#define DIG_GELUG3_P2 1.421469197e-02f
#define DIG_GELUG3_P3 1.283048511e-04f
#define DIG_GELUG3_Q1 3.016213812e-01f
#define DIG_GELUG3_Q2 4.027552325e-02f
#define DIG_GELUG3_Q3 4.900047771e-03f
__device__ __forceinline__ float gelu_erf_grad(float x) {
  x = fminf(fmaxf(x, -DIG_GELUG_CLAMP), DIG_GELUG_CLAMP);
  const float t = x * x;
  float p = DIG_GELUG_P5;
  p = fmaf(p, t, DIG_GELUG_P4);
  p = fmaf(p, t, DIG_GELUG_P3);
  p = fmaf(p, t, DIG_GELUG_P2);
  p = fmaf(p, t, DIG_GELUG_P1);
  p = fmaf(p, t, DIG_GELUG_P0);
  float q = DIG_GELUG_Q5;
  q = fmaf(q, t, DIG_GELUG_Q4);
  q = fmaf(q, t, DIG_GELUG_Q3);
  q = fmaf(q, t, DIG_GELUG_Q2);
  q = fmaf(q, t, DIG_GELUG_Q1);
  q = fmaf(q, t, 1.0f);
  return fmaf(x * p, rcp_approx(q), 0.5f);
}

// ---- the same two functions on PAIRS of values with sm_100's packed fp32 instructions (FFMA2 / FMUL2 / FADD2): half the issue slots
// of the scalar forms for the polynomial part.  Used by the GEMM epilogues, which are issue-bound.
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t f2_pack(float lo, float hi) {
  f32x2_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void f2_unpack(f32x2_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2_t f2_fma(f32x2_t a, f32x2_t b, f32x2_t c) {
  f32x2_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ f32x2_t f2_mul(f32x2_t a, f32x2_t b) {
  f32x2_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2_t f2_add(f32x2_t a, f32x2_t b) {
  f32x2_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
#define DIG_F2C(c) f2_pack(c, c)
// (g0, g1) = gelu(x0, x1)
__device__ __forceinline__ void gelu_erf_x2(float x0, float x1, float& g0, float& g1) {
#ifdef DIG_GELU_IDENTITY   // experiment: epilogue without the activation arithmetic (what does the data movement alone cost?)
  g0 = x0; g1 = x1; return;
#endif
#if DIG_GELU_ORDER == 5
  const float c0 = fminf(fmaxf(x0, -DIG_GELU_CLAMP), DIG_GELU_CLAMP);
  const float c1 = fminf(fmaxf(x1, -DIG_GELU_CLAMP), DIG_GELU_CLAMP);
#else
  const float c0 = fminf(fmaxf(x0, -DIG_GELU3_CLAMP), DIG_GELU3_CLAMP);
  const float c1 = fminf(fmaxf(x1, -DIG_GELU3_CLAMP), DIG_GELU3_CLAMP);
#endif
  const f32x2_t xc = f2_pack(c0, c1);
  const f32x2_t t = f2_mul(xc, xc);
#if DIG_GELU_ORDER == 5
  f32x2_t p = f2_fma(DIG_F2C(DIG_GELU_P5), t, DIG_F2C(DIG_GELU_P4));
  p = f2_fma(p, t, DIG_F2C(DIG_GELU_P3));
  p = f2_fma(p, t, DIG_F2C(DIG_GELU_P2));
  p = f2_fma(p, t, DIG_F2C(DIG_GELU_P1));
  p = f2_fma(p, t, DIG_F2C(DIG_GELU_P0));
#else
  f32x2_t p = f2_fma(DIG_F2C(DIG_GELU3_P3), t, DIG_F2C(DIG_GELU3_P2));
  p = f2_fma(p, t, DIG_F2C(DIG_GELU3_P1));
  p = f2_fma(p, t, DIG_F2C(DIG_GELU3_P0));
#endif
#if DIG_GELU_ORDER == 5
  f32x2_t q = f2_fma(DIG_F2C(DIG_GELU_Q5), t, DIG_F2C(DIG_GELU_Q4));
  q = f2_fma(q, t, DIG_F2C(DIG_GELU_Q3));
  q = f2_fma(q, t, DIG_F2C(DIG_GELU_Q2));
  q = f2_fma(q, t, DIG_F2C(DIG_GELU_Q1));
#else
  f32x2_t q = f2_fma(DIG_F2C(DIG_GELU3_Q3), t, DIG_F2C(DIG_GELU3_Q2));
  q = f2_fma(q, t, DIG_F2C(DIG_GELU3_Q1));
#endif
  q = f2_fma(q, t, DIG_F2C(1.0f));
  float q0, q1;
  f2_unpack(q, q0, q1);
  const f32x2_t w = f2_fma(f2_mul(xc, p), f2_pack(rcp_approx(q0), rcp_approx(q1)), DIG_F2C(0.5f));
  f2_unpack(f2_mul(f2_pack(x0, x1), w), g0, g1);
}
// (d0, d1) *= gelu'(x0, x1)
__device__ __forceinline__ void gelu_erf_grad_mul_x2(float x0, float x1, float& d0, float& d1) {
#ifdef DIG_GELU_IDENTITY
  d0 += 0.f * x0; d1 += 0.f * x1; return;
#endif
#if DIG_GELU_ORDER == 5
  x0 = fminf(fmaxf(x0, -DIG_GELUG_CLAMP), DIG_GELUG_CLAMP);
  x1 = fminf(fmaxf(x1, -DIG_GELUG_CLAMP), DIG_GELUG_CLAMP);
#else
  x0 = fminf(fmaxf(x0, -DIG_GELUG3_CLAMP), DIG_GELUG3_CLAMP);
  x1 = fminf(fmaxf(x1, -DIG_GELUG3_CLAMP), DIG_GELUG3_CLAMP);
#endif
  const f32x2_t x = f2_pack(x0, x1);
  const f32x2_t t = f2_mul(x, x);
#if DIG_GELU_ORDER == 5
  f32x2_t p = f2_fma(DIG_F2C(DIG_GELUG_P5), t, DIG_F2C(DIG_GELUG_P4));
  p = f2_fma(p, t, DIG_F2C(DIG_GELUG_P3));
  p = f2_fma(p, t, DIG_F2C(DIG_GELUG_P2));
  p = f2_fma(p, t, DIG_F2C(DIG_GELUG_P1));
  p = f2_fma(p, t, DIG_F2C(DIG_GELUG_P0));
#else
  f32x2_t p = f2_fma(DIG_F2C(DIG_GELUG3_P3), t, DIG_F2C(DIG_GELUG3_P2));
  p = f2_fma(p, t, DIG_F2C(DIG_GELUG3_P1));
  p = f2_fma(p, t, DIG_F2C(DIG_GELUG3_P0));
#endif
#if DIG_GELU_ORDER == 5
  f32x2_t q = f2_fma(DIG_F2C(DIG_GELUG_Q5), t, DIG_F2C(DIG_GELUG_Q4));
  q = f2_fma(q, t, DIG_F2C(DIG_GELUG_Q3));
  q = f2_fma(q, t, DIG_F2C(DIG_GELUG_Q2));
  q = f2_fma(q, t, DIG_F2C(DIG_GELUG_Q1));
#else
  f32x2_t q = f2_fma(DIG_F2C(DIG_GELUG3_Q3), t, DIG_F2C(DIG_GELUG3_Q2));
  q = f2_fma(q, t, DIG_F2C(DIG_GELUG3_Q1));
#endif
  q = f2_fma(q, t, DIG_F2C(1.0f));
  float q0, q1;
  f2_unpack(q, q0, q1);
  const f32x2_t w = f2_fma(f2_mul(x, p), f2_pack(rcp_approx(q0), rcp_approx(q1)), DIG_F2C(0.5f));
  f2_unpack(f2_mul(f2_pack(d0, d1), w), d0, d1);
}

// ------------------------------------------------------------------------------------------------
// mbarrier
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// non-blocking probe (try_wait may suspend the thread for a system-dependent time before it reports failure: not for polling loops)
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// ------------------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor) -- 2-D tile load, completion on an mbarrier
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c_inner, int c_outer) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c_inner), "r"(c_outer)
      : "memory");
}
// TMA stores (shared -> global tile, bulk async-group completion)
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, uint32_t smem, int c_inner, int c_outer) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(m)),
               "r"(smem), "r"(c_inner), "r"(c_outer)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* m, uint32_t smem, int c_inner, int c_outer) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem), "r"(c_inner), "r"(c_outer)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read_1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// TMA prefetch of a tensor-map box into L2 (no shared-memory destination, no completion)
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* m, int c_inner, int c_outer) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c_inner), "r"(c_outer)
               : "memory");
}
// generic-proxy writes to smem -> visible to the async proxy (UMMA / TMA reads)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_holder, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_holder)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// one lane of a converged warp (warp-uniform control flow around it keeps the MMA operands in uniform registers: issuing from inside
// `if (lane == 0)` makes ptxas wrap every tcgen05.mma in an ELECT / R2UR waterfall of ~20 instructions, which caps the issue rate)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// MMA completion -> mbarrier arrive (implies fence::before_thread_sync)
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], bf16 inputs, fp32 accumulate
__device__ __forceinline__ void tc_mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void tc_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Instruction descriptor, kind::f16, bf16 x bf16 -> fp32 (bit layout: cute/arch/mma_sm100_desc.hpp InstrDescriptor)
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, bool a_mn_major, bool b_mn_major) {
  return (1u << 4)                              // c_format  = F32
         | (1u << 7)                            // a_format  = BF16
         | (1u << 10)                           // b_format  = BF16
         | ((a_mn_major ? 1u : 0u) << 15)       // a_major
         | ((b_mn_major ? 1u : 0u) << 16)       // b_major
         | ((uint32_t)(N >> 3) << 17)           // n_dim
         | ((uint32_t)(M >> 4) << 24);          // m_dim
}

// Shared-memory matrix descriptor, 128-byte swizzle.  Byte offsets are encoded >> 4.
//   K-major  operand: rows of 64 bf16 (128 B); 8-row groups SBO bytes apart (1024 for a dense tile); LBO unused (1).
//   MN-major operand: 64-element MN chunks x 8 K-rows form one 1024-B atom; next 8 K-rows SBO bytes on,
//                     next 64-element MN chunk LBO bytes on.
__device__ __forceinline__ uint64_t make_sdesc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;  // SWIZZLE_128B
  return d;
}

// TMEM -> registers: this thread's lane, 32 consecutive 32-bit columns
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// registers -> TMEM: this thread's lane, 16 consecutive 32-bit columns
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
      "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// Explicit shared-memory accesses by 32-bit shared address.  The kernels carve their dynamic shared memory by hand (1024-byte
// alignment arithmetic), after which the compiler no longer knows the address space and would emit generic LD/ST.
__device__ __forceinline__ float4 lds_f4(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_f4(uint32_t a, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void sts_u4(uint32_t a, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void red_shared_add_f32(uint32_t a, float v) {
  asm volatile("red.shared.add.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory");
}

__device__ __forceinline__ uint4 lds_u4(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ float lds_f32(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}

// 128-byte swizzle: byte offset of (row, 16-byte chunk) inside a tile whose rows are 128 bytes wide
__device__ __forceinline__ uint32_t sw128_offset(uint32_t row, uint32_t chunk16) {
  return row * 128u + (((chunk16 ^ row) & 7u) << 4);
}
#endif  // __CUDACC__

}  // namespace dig

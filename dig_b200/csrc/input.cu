// dig_b200 -- GPU-side input stage of the pre-training step: what the reference's CPU workers do AFTER the
// image has been decoded, augmented and resized to 32 x 128 uint8 RGB, and per sample:
//   * transforms.ToTensor + Normalize(mean = std = 0.5) of both views (dataset/datasets.py:30-37, dataset/dataset_image.py:39-52):
//     uint8 HWC -> fp32 CHW, ((x / 255) - 0.5) / 0.5, evaluated with the same IEEE operations as torchvision (bit-exact);
//   * transforms.RandomGrayscale(p) of the augmented view (dataset_image.py:46): PIL's ITU-R 601-2 luma (R*19595 + G*38470 + B*7471 +
//     0x8000) >> 16 replicated to the three channels, applied to a sample when its uniform draw is < p;
//   * RandomMaskingGenerator (masking_generator.py:12-46): for every (sample, view) a uniformly random subset of n_mask of the 256
//     patch positions.  The reference shuffles [0]*(256-n) + [1]*n with numpy's global Mersenne twister inside each worker; here each
//     position draws a 64-bit key from a counter-based hash of (seed, step, sample, view, position) and the n_mask smallest keys are
//     masked -- the same distribution (every subset equally likely), reproducible from (seed, step) alone, independent of the batch split.
// Host batches then cross PCIe as uint8 (4x fewer bytes than the fp32 tensors the reference's loader ships) and no CPU touches a pixel.
#include "common.cuh"
#include "../../include/dig_b200.h"

namespace dig {

__host__ __device__ inline unsigned long long mix64(unsigned long long z) {   // splitmix64 finaliser
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__host__ __device__ inline unsigned long long draw64(unsigned long long seed, unsigned long long step, unsigned long long a, unsigned long long b,
                                                      unsigned long long c) {
  return mix64(mix64(mix64(mix64(seed) ^ step) ^ (a * 0x100000001B3ull + b)) ^ c);
}

// one thread per 4 horizontally adjacent pixels of one image row: 12 input bytes (3 x 32-bit loads), one float4 store per channel
__global__ void __launch_bounds__(256)
normalize_views_kernel(const uint8_t* __restrict__ img, const uint8_t* __restrict__ aug, float* __restrict__ img_out, float* __restrict__ aug_out,
                       long long B, float gray_p, unsigned long long seed, unsigned long long step, long long sample0) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // over 2 views x B x 32 rows x 32 pixel quads
  if (i >= 2 * B * 32 * 32) return;
  const int quad = (int)(i & 31), row = (int)((i >> 5) & 31);
  const long long b = (i >> 10) % B;
  const int view = (int)((i >> 10) / B);
  const uint8_t* src = (view ? aug : img) + ((b * 32 + row) * 128 + quad * 4) * 3;
  float* dst = (view ? aug_out : img_out) + (b * 3 * 32 + row) * 128 + quad * 4;
  const uint32_t w0 = *reinterpret_cast<const uint32_t*>(src), w1 = *reinterpret_cast<const uint32_t*>(src + 4),
                 w2 = *reinterpret_cast<const uint32_t*>(src + 8);
  uint8_t px[12];
#pragma unroll
  for (int k = 0; k < 4; ++k) { px[k] = (uint8_t)(w0 >> (8 * k)); px[4 + k] = (uint8_t)(w1 >> (8 * k)); px[8 + k] = (uint8_t)(w2 >> (8 * k)); }
  bool gray = false;
  if (view == 1 && gray_p > 0.f) {
    const unsigned long long r = draw64(seed, step, (unsigned long long)(sample0 + b), 0xA5ull, 0x6772617900ull);
    gray = (float)(r >> 40) * (1.0f / 16777216.0f) < gray_p;      // 24-bit uniform in [0, 1)
  }
  float o[3][4];
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    uint32_t r = px[3 * p], g = px[3 * p + 1], bl = px[3 * p + 2];
    if (gray) { const uint32_t l = (r * 19595u + g * 38470u + bl * 7471u + 0x8000u) >> 16; r = g = bl = l; }
    o[0][p] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)r, 255.0f), 0.5f), 0.5f);
    o[1][p] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)g, 255.0f), 0.5f), 0.5f);
    o[2][p] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)bl, 255.0f), 0.5f), 0.5f);
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) *reinterpret_cast<float4*>(dst + (long long)c * 32 * 128) = make_float4(o[c][0], o[c][1], o[c][2], o[c][3]);
}

// one block of 256 threads per (sample, view): thread = patch position; rank of its key among the 256 keys (ties by index) < n_mask -> 1
__global__ void __launch_bounds__(256)
random_masks_kernel(uint8_t* __restrict__ mask_u8, double* __restrict__ mask_f64, int V, int n_mask, unsigned long long seed,
                    unsigned long long step, long long sample0) {
  __shared__ unsigned long long keys[256];
  const long long b = blockIdx.x / V;
  const int v = blockIdx.x % V, t = threadIdx.x;
  const unsigned long long k = draw64(seed, step, (unsigned long long)(sample0 + b), (unsigned long long)v, (unsigned long long)t);
  keys[t] = k;
  __syncthreads();
  int rank = 0;
#pragma unroll 8
  for (int j = 0; j < 256; ++j) {
    const unsigned long long o = keys[j];
    rank += (o < k || (o == k && j < t)) ? 1 : 0;
  }
  const int m = rank < n_mask ? 1 : 0;
  const long long idx = (long long)blockIdx.x * 256 + t;
  if (mask_u8) mask_u8[idx] = (uint8_t)m;
  if (mask_f64) mask_f64[idx] = (double)m;
}

}  // namespace dig

using namespace dig;

extern "C" int dig_normalize_views(const uint8_t* img_u8, const uint8_t* aug_u8, float* img_out, float* aug_out, int64_t B, float gray_p,
                                   int64_t seed, int64_t step, int64_t sample0, void* stream) {
  DIG_REQUIRE(img_u8 && aug_u8 && img_out && aug_out && B > 0, "dig_normalize_views: bad arguments");
  DIG_REQUIRE(((((uintptr_t)img_u8) | ((uintptr_t)aug_u8)) & 3) == 0 && ((((uintptr_t)img_out) | ((uintptr_t)aug_out)) & 15) == 0,
              "dig_normalize_views: inputs must be 4-byte and outputs 16-byte aligned");
  const long long n = 2 * B * 32 * 32;
  normalize_views_kernel<<<(int)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(img_u8, aug_u8, img_out, aug_out, B, gray_p,
                                                                                 (unsigned long long)seed, (unsigned long long)step, sample0);
  DIG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int dig_random_masks(uint8_t* mask_u8, double* mask_f64, int64_t B, int32_t num_view, int32_t n_mask, int64_t seed, int64_t step,
                                int64_t sample0, void* stream) {
  DIG_REQUIRE((mask_u8 || mask_f64) && B > 0 && num_view > 0 && n_mask >= 0 && n_mask <= 256, "dig_random_masks: bad arguments");
  random_masks_kernel<<<(int)(B * num_view), 256, 0, (cudaStream_t)stream>>>(mask_u8, mask_f64, num_view, n_mask, (unsigned long long)seed,
                                                                           (unsigned long long)step, sample0);
  DIG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

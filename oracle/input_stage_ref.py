"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the GPU-side input stage (dig_b200/csrc/input.cu).

* `normalize_views`: transforms.ToTensor + transforms.Normalize(0.5, 0.5) (reference dataset/datasets.py:30-37 and
  dataset/dataset_image.py:39-52) and transforms.RandomGrayscale's conversion (dataset_image.py:46 -> PIL "L": ITU-R 601-2 luma,
  (R*19595 + G*38470 + B*7471 + 0x8000) >> 16), numpy fp32, pinned against torchvision itself in tests/test_input_stage.py.
* `random_masks`: RandomMaskingGenerator (masking_generator.py:12-46) draws, per view, a uniformly random arrangement of n_mask ones among
  256 positions with numpy's global Mersenne twister.  The CUDA kernel draws the same DISTRIBUTION from a counter-based hash (splitmix64
  finaliser) so that masks are a pure function of (seed, step, global sample index, view): this file restates that hash bit-for-bit, and
  the tests check (a) the kernel against it exactly, (b) both against the reference generator statistically (exact count per view,
  per-position frequency, independence of views).
"""
import numpy as np

MASK64 = (1 << 64) - 1


def mix64(z):
    z = (z + 0x9E3779B97F4A7C15) & MASK64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK64
    return z ^ (z >> 31)


def draw64(seed, step, a, b, c):
    return mix64(mix64(mix64(mix64(seed & MASK64) ^ (step & MASK64)) ^ ((a * 0x100000001B3 + b) & MASK64)) ^ (c & MASK64))


def random_masks(B, num_view, n_mask, seed, step, sample0=0):
    out = np.zeros((B, num_view, 256), dtype=np.uint8)
    for b in range(B):
        for v in range(num_view):
            keys = [draw64(seed, step, sample0 + b, v, t) for t in range(256)]
            order = sorted(range(256), key=lambda t: (keys[t], t))
            out[b, v, order[:n_mask]] = 1
    return out


def gray_decision(seed, step, sample, gray_p):
    r = draw64(seed, step, sample, 0xA5, 0x6772617900)
    return np.float32(r >> 40) * np.float32(1.0 / 16777216.0) < np.float32(gray_p)


def to_gray_u8(x):
    """x uint8 [...,3] -> PIL 'L' luma replicated to 3 channels."""
    x = x.astype(np.uint32)
    l = (x[..., 0] * 19595 + x[..., 1] * 38470 + x[..., 2] * 7471 + 0x8000) >> 16
    return np.stack([l, l, l], axis=-1).astype(np.uint8)


def normalize_view(x_u8):
    """uint8 [B,32,128,3] -> fp32 [B,3,32,128], ((x/255) - 0.5)/0.5 with fp32 IEEE operations (ToTensor then Normalize)."""
    x = x_u8.astype(np.float32) / np.float32(255.0)
    x = (x - np.float32(0.5)) / np.float32(0.5)
    return np.ascontiguousarray(x.transpose(0, 3, 1, 2))


def normalize_views(img_u8, aug_u8, gray_p, seed, step, sample0=0):
    aug = aug_u8.copy()
    for b in range(aug.shape[0]):
        if gray_p > 0 and gray_decision(seed, step, sample0 + b, gray_p):
            aug[b] = to_gray_u8(aug[b])
    return normalize_view(img_u8), normalize_view(aug)

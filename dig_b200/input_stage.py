"""GPU-side input stage: the per-sample work the reference leaves to its CPU DataLoader workers after
decode / augment / resize -- ToTensor + Normalize of both views (dataset/datasets.py:27-49, dataset/dataset_image.py:39-52),
RandomGrayscale of the augmented view (dataset_image.py:46) and the per-view random masks (masking_generator.py:12-46) -- as two
kernels behind the C-ABI (dig_normalize_views, dig_random_masks).  The loader then ships uint8 [B,32,128,3] views (4x fewer PCIe bytes than the
reference's fp32 tensors, no mask tensor at all) and `GpuInputStage` returns exactly the batch `train_one_epoch` expects.
"""
import torch

from . import ops
from .ops import call


class GpuInputStage:
    def __init__(self, mask_ratio=0.7, num_view=2, num_patches=256, gray_p=0.2, seed=0):
        if num_patches != 256:
            raise ops.DigError("the input stage is built for the 8 x 32 patch grid of 32x128 crops (256 patches)")
        self.num_view = int(num_view)
        self.num_mask = int(mask_ratio * num_patches)           # masking_generator.py:20
        self.gray_p = float(gray_p)
        self.seed = int(seed)
        self.step = 0

    def __call__(self, img_u8, aug_u8, sample0=0, step=None):
        """img_u8, aug_u8: uint8 [B,32,128,3] (CUDA, or pinned host tensors that are copied first).  Returns (images fp32 [B,3,32,128],
        aug_images fp32 [B,3,32,128], mask bool [B,num_view,256]) on the device.  `sample0` is the global index of the batch's first sample
        (so that masks do not depend on how a global batch is split over ranks); `step` defaults to an internal counter."""
        if img_u8.dtype != torch.uint8 or aug_u8.dtype != torch.uint8 or tuple(img_u8.shape[1:]) != (32, 128, 3) or img_u8.shape != aug_u8.shape:
            raise ops.DigError("GpuInputStage expects two uint8 [B,32,128,3] views, got %s / %s" % (tuple(img_u8.shape), tuple(aug_u8.shape)))
        dev = img_u8.device if img_u8.is_cuda else torch.device("cuda", torch.cuda.current_device())
        img_u8 = img_u8.to(dev, non_blocking=True).contiguous()
        aug_u8 = aug_u8.to(dev, non_blocking=True).contiguous()
        B = img_u8.shape[0]
        step = self.step if step is None else int(step)
        self.step = step + 1
        images = torch.empty(B, 3, 32, 128, dtype=torch.float32, device=dev)
        aug_images = torch.empty_like(images)
        call("dig_normalize_views", img_u8, aug_u8, images, aug_images, B, self.gray_p, self.seed, step, int(sample0))
        mask = torch.empty(B, self.num_view, 256, dtype=torch.uint8, device=dev)
        call("dig_random_masks", mask, None, B, self.num_view, self.num_mask, self.seed, step, int(sample0))
        return images, aug_images, mask.view(torch.bool)

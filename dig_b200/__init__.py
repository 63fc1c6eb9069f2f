"""dig_b200 -- B200-native (sm_100a) implementation of DiG's self-supervised pre-training step.

Public surface (mirrors the reference's two call sites):
  * dig_b200.modeling   -- `pretrain_simmim_moco_ori_vit_{tiny,small,base}_patch4_32x128` factories
  * dig_b200.engine     -- `train_one_epoch(...)` with the reference's signature
  * dig_b200.ops        -- ctypes binding of the C-ABI in include/dig_b200.h (libdig_b200.so)
"""
from . import registry  # noqa: F401
from .registry import create_model, register_model  # noqa: F401

__version__ = "0.1.0"

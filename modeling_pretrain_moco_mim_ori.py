"""Drop-in for the reference module of the same name: importing it registers the
`pretrain_simmim_moco_ori_vit_{tiny,small,base}_patch4_32x128` factories (run_mae_pretraining_moco.py:36,278-294).
The implementation lives in dig_b200/ (hand-written sm_100a kernels behind include/dig_b200.h)."""
from dig_b200.modeling import (DigMoCoViT as MoCo_ViT,  # noqa: F401
                               pretrain_simmim_moco_ori_vit_base_patch4_32x128,
                               pretrain_simmim_moco_ori_vit_small_patch4_32x128,
                               pretrain_simmim_moco_ori_vit_tiny_patch4_32x128)

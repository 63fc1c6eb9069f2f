"""Drop-in model boundary B1: the `pretrain_simmim_moco_ori_vit_*_patch4_32x128`
factories and an `nn.Module` whose parameters / buffers carry the reference's exact state-dict names
and shapes (reference: modeling_pretrain_moco_mim_ori.py:261-426 `MoCo_ViT.__init__`,
modeling_pretrain_vit.py:27-73, modeling_finetune.py:43-196), so that checkpoints interchange with
utils/utils.py:546-669 and run_class_finetuning.py:362-442.

The modules below are *parameter holders only*: none of them computes anything.  All arithmetic of
`forward` runs in the hand-written sm_100a kernels behind the C-ABI (`dig_b200/csrc`, bound in
`dig_b200/ops.py`, sequenced by `dig_b200/pretrain_step.py`).  There is no CPU / eager fallback: calling
`forward` without the CUDA extension or on a CPU tensor raises.

Construction order and the init calls mirror the reference so that, for a given `torch.manual_seed`,
every parameter is bit-identical to the reference's (checked in tests/test_boundary.py when the
reference is importable).
"""
import math
from functools import partial

import torch
import torch.nn as nn

from . import registry

_FROZEN_PREFIXES = ("momentum_encoder.", "momentum_projection_layer.", "pix_projector_m.")


class _Holder(nn.Module):
    """Parameter container; the fused pipeline reads its tensors, `forward` is never used."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("dig_b200 holder modules do not compute; call the top-level model")


class _Attn(_Holder):
    # reference layout: F:63-85 -- fused qkv weight [3d,d] without bias, learnable q/v bias, proj with bias
    def __init__(self, dim):
        super().__init__()
        self.qkv = nn.Linear(dim, dim * 3, bias=False)
        self.q_bias = nn.Parameter(torch.zeros(dim))
        self.v_bias = nn.Parameter(torch.zeros(dim))
        self.proj = nn.Linear(dim, dim)


class _Mlp(_Holder):
    # F:43-52
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)


class _Block(_Holder):
    # F:128-148 with init_values=0 -> no gamma_1/gamma_2
    def __init__(self, dim, mlp_ratio, eps):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=eps)
        self.attn = _Attn(dim)
        self.norm2 = nn.LayerNorm(dim, eps=eps)
        self.mlp = _Mlp(dim, int(dim * mlp_ratio))


class _PatchEmbed(_Holder):
    # F:173-188
    def __init__(self, img_size, patch_size, in_chans, embed_dim):
        super().__init__()
        self.img_size = tuple(img_size)
        self.patch_size = (patch_size, patch_size)
        self.patch_shape = (img_size[0] // patch_size, img_size[1] // patch_size)
        self.num_patches = self.patch_shape[0] * self.patch_shape[1]
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)


def sinusoid_table(n_position, d_hid):
    """Fixed sin/cos position table, F:200-210 (float64 math, stored fp32)."""
    pos = torch.arange(n_position, dtype=torch.float64)[:, None]
    j = torch.arange(d_hid, dtype=torch.float64)[None, :]
    ang = pos / torch.pow(torch.tensor(10000.0, dtype=torch.float64), 2 * torch.floor(j / 2) / d_hid)
    tab = torch.where((torch.arange(d_hid) % 2 == 0)[None, :], torch.sin(ang), torch.cos(ang))
    return tab.to(torch.float32).unsqueeze(0)


class _Encoder(_Holder):
    """Holder for PretrainVisionTransformerEncoder's tensors (V:27-73)."""

    def __init__(self, img_size, patch_size, in_chans, embed_dim, depth, num_heads, mlp_ratio, eps, final_norm=False):
        super().__init__()
        self.embed_dim = self.num_features = embed_dim
        self.num_heads = num_heads
        self.patch_embed = _PatchEmbed(img_size, patch_size, in_chans, embed_dim)
        self.mask_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = sinusoid_table(self.patch_embed.num_patches, embed_dim)  # plain attr, V:48
        self.blocks = nn.ModuleList([_Block(embed_dim, mlp_ratio, eps) for _ in range(depth)])
        # the pre-training model strips the final norm (M:362 sets encoder.norm = Identity); the fine-tuning encoder keeps it (V:57,104)
        self.norm = nn.LayerNorm(embed_dim, eps=eps) if final_norm else nn.Identity()
        self.head = nn.Identity()
        # V:63-73: xavier on every Linear (bias 0), LayerNorm 1/0, in module-traversal order
        for mod in self.modules():
            if isinstance(mod, nn.Linear):
                nn.init.xavier_uniform_(mod.weight)
                if mod.bias is not None:
                    nn.init.constant_(mod.bias, 0)
            elif isinstance(mod, nn.LayerNorm):
                nn.init.constant_(mod.bias, 0)
                nn.init.constant_(mod.weight, 1.0)

    def get_num_layers(self):
        return len(self.blocks)

    def no_weight_decay(self):
        return {"pos_embed", "cls_token"}


def _bn_mlp(num_layers, input_dim, mlp_dim, output_dim):
    """Linear(no bias) -> BN -> ReLU ... , last BN affine=False (M:463-482)."""
    mods = []
    for l in range(num_layers):
        d1 = input_dim if l == 0 else mlp_dim
        d2 = output_dim if l == num_layers - 1 else mlp_dim
        mods.append(nn.Linear(d1, d2, bias=False))
        if l < num_layers - 1:
            mods.append(nn.BatchNorm1d(d2))
            mods.append(nn.ReLU(inplace=True))
        else:
            mods.append(nn.BatchNorm1d(d2, affine=False))
    return nn.Sequential(*mods)


class _PatchPool(_Holder):
    """patchnet_name='no_patchtrans' PatchNet: parameter-free window pooling (M:136-157, M:189-193)."""

    def __init__(self, num_windows, patch_shape):
        super().__init__()
        self.num_windows = num_windows
        self.patch_shape = patch_shape
        self.use_patch_transformer = False


class DigMoCoViT(nn.Module):
    """State-dict compatible stand-in for the reference `MoCo_ViT` (M:261-577) restricted to the
    README configuration: use_pixel_target=True, use_moco_target=True, use_pix_projector=True,
    patchnet_name='no_patchtrans', sinusoid position table, drop_path 0."""

    def __init__(self, img_size=(32, 128), patch_size=4, in_chans=3, encoder_embed_dim=384, encoder_depth=12,
                 encoder_num_heads=6, decoder_num_classes=48, decoder_embed_dim=192, mlp_ratio=4.0,
                 qkv_bias=True, norm_eps=1e-6, drop_path_rate=0.0, num_classes=0, mlp_dim=4096, dim=256, T=1.0,
                 num_windows=5, encoder_type="vit", queue_size=65536, patchnet_name="regular",
                 label_smoothing=0.0, use_pix_projector=True, **unused):
        super().__init__()
        if patchnet_name != "no_patchtrans":
            raise NotImplementedError(
                "dig_b200 implements the README configuration patchnet_name='no_patchtrans' (README.md:77); "
                "got %r" % (patchnet_name,))
        if drop_path_rate not in (0, 0.0, None):
            raise NotImplementedError("drop_path must be 0.0 on the pre-training path (R:87)")
        if not qkv_bias or label_smoothing != 0.0 or not use_pix_projector:
            raise NotImplementedError("only qkv_bias=True, label_smoothing=0, use_pix_projector=True are built")
        if encoder_embed_dim != 64 * encoder_num_heads:
            raise NotImplementedError("the attention kernels are built for head_dim 64 (got d=%d, heads=%d)" % (encoder_embed_dim, encoder_num_heads))
        grid_w = img_size[1] // patch_size
        if num_windows <= 0 or grid_w % num_windows != 0:
            raise NotImplementedError("num_windows must divide the %d patch columns (README.md:76 runs --num_windows 4); got %r -- "
                                      "adaptive_avg_pool2d's uneven windows (M:192) are not built" % (grid_w, num_windows))
        self.T = T
        self.num_windows = num_windows
        self.use_pixel_target = True
        self.use_moco_target = True
        self.label_smoothing = label_smoothing
        enc_args = (img_size, patch_size, in_chans, encoder_embed_dim, encoder_depth, encoder_num_heads,
                    mlp_ratio, norm_eps)
        self.encoder = _Encoder(*enc_args)
        print("using moco branch.")
        self.momentum_encoder = _Encoder(*enc_args)
        # M:353-355: patch-embed weight re-drawn U(+-sqrt(6/(3*p*p+d))), bias 0
        val = math.sqrt(6.0 / float(3 * patch_size * patch_size + encoder_embed_dim))
        nn.init.uniform_(self.encoder.patch_embed.proj.weight, -val, val)
        nn.init.zeros_(self.encoder.patch_embed.proj.bias)
        self.encoder_projection_layer = _bn_mlp(3, encoder_embed_dim, mlp_dim, dim)
        self.momentum_projection_layer = _bn_mlp(3, encoder_embed_dim, mlp_dim, dim)
        self.predictor = _bn_mlp(2, dim, mlp_dim, dim)
        self.patch_extractor = _PatchPool(num_windows, self.encoder.patch_embed.patch_shape)
        self.momentum_patch_extractor = _PatchPool(num_windows, self.encoder.patch_embed.patch_shape)
        self._init_momentum(self.encoder, self.momentum_encoder)
        self._init_momentum(self.encoder_projection_layer, self.momentum_projection_layer)
        print("using mim branch.")
        self.pix_projector = _bn_mlp(3, encoder_embed_dim, 512, encoder_embed_dim)
        self.pix_projector_m = _bn_mlp(3, encoder_embed_dim, 512, encoder_embed_dim)
        self._init_momentum(self.pix_projector, self.pix_projector_m)
        self.pix_decoder = nn.Sequential(
            nn.Linear(encoder_embed_dim, decoder_embed_dim, bias=False),
            nn.Linear(decoder_embed_dim, decoder_embed_dim, bias=False),
            nn.LayerNorm(decoder_embed_dim, eps=1e-6),
            nn.GELU(),
            nn.Linear(decoder_embed_dim, decoder_num_classes))
        self._step = None  # lazily built dig_b200.pretrain_step.PretrainStep

    @staticmethod
    def _init_momentum(online, momentum):
        for pb, pm in zip(online.parameters(), momentum.parameters()):
            pm.data.copy_(pb.data)
            pm.requires_grad = False

    @torch.jit.ignore
    def no_weight_decay(self):
        return {"pos_embed", "cls_token"}

    # ---- fused forward -------------------------------------------------------------------------
    def _pipeline(self):
        from .pretrain_step import PretrainStep  # imports the C-ABI; raises loudly if it is missing
        if self._step is None or not self._step.matches(self):
            self._step = PretrainStep(self)
        return self._step

    def forward(self, image, aug_image, vis_mask_pos, m, only_mim_on_ori_img=True):
        """Same contract as M:488-577: returns {'contra_loss', 'q{1,2}_acc{1,5}', 'vis_out': [ [B,n,48] ]}."""
        if not image.is_cuda:
            raise RuntimeError("dig_b200 runs on sm_100a only: inputs must be CUDA tensors (no CPU fallback)")
        if not self.training:
            raise NotImplementedError("eval mode is not built: the BatchNorm heads always normalise with batch statistics and update their "
                                      "running estimates (the pre-training runner never leaves train mode, E:35)")
        from .pretrain_step import run_model
        return run_model(self._pipeline(), image, aug_image, vis_mask_pos, float(m), bool(only_mim_on_ori_img))


def _factory(embed_dim, heads, **kwargs):
    kwargs.pop("pretrained", None)
    kwargs.pop("in_chans", None)       # timm 0.3.2 create_model may inject these
    kwargs.pop("drop_block_rate", None)
    init_ckpt = kwargs.pop("init_ckpt", None)
    model = DigMoCoViT(img_size=(32, 128), patch_size=4, encoder_embed_dim=embed_dim, encoder_depth=12,
                       encoder_num_heads=heads, decoder_num_classes=48, decoder_embed_dim=192, mlp_ratio=4,
                       qkv_bias=True, norm_eps=1e-6, **kwargs)
    model.default_cfg = {"num_classes": 1000, "input_size": (3, 224, 224), "crop_pct": 0.9,
                         "interpolation": "bicubic", "mean": (0.5, 0.5, 0.5), "std": (0.5, 0.5, 0.5)}
    if init_ckpt:
        model.load_state_dict(torch.load(init_ckpt, map_location="cpu")["model"])
    return model


@registry.register_model
def pretrain_simmim_moco_ori_vit_tiny_patch4_32x128(pretrained=False, **kwargs):
    """M:765-789"""
    return _factory(192, 3, **kwargs)


@registry.register_model
def pretrain_simmim_moco_ori_vit_small_patch4_32x128(pretrained=False, **kwargs):
    """M:682-707"""
    return _factory(384, 6, **kwargs)


@registry.register_model
def pretrain_simmim_moco_ori_vit_base_patch4_32x128(pretrained=False, **kwargs):
    """M:792-817 -- this repo's 'base' is d=512, 8 heads"""
    return _factory(512, 8, **kwargs)

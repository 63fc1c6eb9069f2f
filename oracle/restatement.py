"""TEST INFRASTRUCTURE ONLY -- CPU/fp32 restatement of DiG's pre-training step.

This file is the *oracle* for the hot path named in BASELINE.json: a plain PyTorch fp32,
functional restatement (no nn.Module, no custom kernels) of what the reference computes in
`MoCo_ViT.forward` + `train_one_epoch` for the `pretrain_simmim_moco_ori_vit_*_patch4_32x128`
configuration.  It exists to CHECK the CUDA path; only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s cpu_baseline / `--impl reference` legs may import it.  Nothing under `dig_b200/`
imports it, and the product path raises if its CUDA extension is missing.

Parity pinning: the reference ships no tests or golden vectors, so this
restatement is pinned against the reference ITSELF, imported live in the authoring container by
`oracle/ref_shims.py`; `oracle/make_golden.py` writes the reference's outputs on seeded inputs to
`tests/golden/` and `tests/test_oracle_vs_reference.py` checks this file against them (and,
where /root/reference is present, against the live reference bit-for-bit in fp32).

Every function cites the reference lines it restates.  Tags: M = modeling_pretrain_moco_mim_ori.py,
V = modeling_pretrain_vit.py, F = modeling_finetune.py, E = engine_for_pretraining_moco.py,
U = utils/utils.py.
All tensors are keyed by the reference's own state-dict names (e.g. `encoder.blocks.0.attn.qkv.weight`).
"""
import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as Fn

CONFIGS = {
    # name -> (embed_dim, heads)                                   M:765-789, M:682-707, M:792-817
    "pretrain_simmim_moco_ori_vit_tiny_patch4_32x128": (192, 3),
    "pretrain_simmim_moco_ori_vit_small_patch4_32x128": (384, 6),
    "pretrain_simmim_moco_ori_vit_base_patch4_32x128": (512, 8),
}
DEPTH = 12
PATCH = 4
IMG_H, IMG_W = 32, 128
GRID_H, GRID_W = IMG_H // PATCH, IMG_W // PATCH   # (8, 32) token grid, F:181-182
NUM_PATCHES = GRID_H * GRID_W                      # 256
LN_EPS = 1e-6                                      # M:697 partial(nn.LayerNorm, eps=1e-6)
BN_EPS = 1e-5                                      # nn.BatchNorm1d default
BN_MOMENTUM = 0.1


def sinusoid_table(n_position=NUM_PATCHES, d_hid=384):
    """F:200-210 get_sinusoid_encoding_table -> [1, n, d] float32 (computed in float64 first)."""
    pos = np.arange(n_position, dtype=np.float64)[:, None]
    j = np.arange(d_hid)[None, :]
    table = pos / np.power(10000, 2 * (j // 2) / d_hid)
    table[:, 0::2] = np.sin(table[:, 0::2])
    table[:, 1::2] = np.cos(table[:, 1::2])
    return torch.tensor(table, dtype=torch.float32).unsqueeze(0)


def patch_embed(sd, prefix, x):
    """F:190-196: 4x4/s4 conv, flatten(2).transpose(1,2) -> [S, 256, d], row-major over (8,32)."""
    y = Fn.conv2d(x, sd[prefix + "patch_embed.proj.weight"], sd[prefix + "patch_embed.proj.bias"],
                  stride=PATCH)
    return y.flatten(2).transpose(1, 2)


def attention(sd, p, x, heads):
    """F:87-125 fused-qkv attention; k-bias is a constant zero (F:91)."""
    B, N, C = x.shape
    qkv_bias = torch.cat((sd[p + "q_bias"], torch.zeros_like(sd[p + "v_bias"]), sd[p + "v_bias"]))
    qkv = Fn.linear(x, sd[p + "qkv.weight"], qkv_bias)
    qkv = qkv.reshape(B, N, 3, heads, -1).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    q = q * (q.shape[-1] ** -0.5)                                   # F:97
    attn = (q @ k.transpose(-2, -1)).softmax(dim=-1)                # F:98, F:114
    out = (attn @ v).transpose(1, 2).reshape(B, N, -1)              # F:118
    return Fn.linear(out, sd[p + "proj.weight"], sd[p + "proj.bias"])  # F:119


def block(sd, p, x, heads):
    """F:150-158 pre-LN block, gamma_* None (init_values=0), DropPath off."""
    d = x.shape[-1]
    h = Fn.layer_norm(x, (d,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], LN_EPS)
    x = x + attention(sd, p + "attn.", h, heads)
    h = Fn.layer_norm(x, (d,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], LN_EPS)
    h = Fn.linear(h, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"])   # F:54
    h = Fn.gelu(h)                                                         # F:55 exact erf GELU
    h = Fn.linear(h, sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])   # F:58
    return x + h


def encoder_forward(sd, prefix, images, mask, heads, depth=DEPTH, taps=None):
    """V:89-106 forward_features with norm/head = Identity (M:362).  mask: bool [S,256] or None."""
    x = patch_embed(sd, prefix, images)
    S, N, C = x.shape
    if mask is not None:                                             # V:94-97
        m = mask.unsqueeze(-1)
        x = x * (~m) + sd[prefix + "mask_token"].expand(S, N, -1) * m
    x = x + sinusoid_table(N, C).to(x.device, x.dtype)               # V:99 (pos_embed is not a parameter)
    if taps is not None:
        taps.append(x)
    for i in range(depth):
        x = block(sd, "%sblocks.%d." % (prefix, i), x, heads)
        if taps is not None:
            taps.append(x)
    return x


def bn_train(x, weight, bias, stats=None):
    """nn.BatchNorm1d in train mode: biased batch variance, eps 1e-5.  `stats`, if given, is a
    (sum, sumsq, count) triple already reduced over ranks (SyncBatchNorm, R:390)."""
    if stats is None:
        mean = x.mean(0)
        var = x.var(0, unbiased=False)
        n = x.shape[0]
    else:
        s, ss, n = stats
        mean = s / n
        var = ss / n - mean * mean
    y = (x - mean) * torch.rsqrt(var + BN_EPS)
    if weight is not None:
        y = y * weight + bias
    return y, mean, var, n


def bn_mlp(sd, prefix, x, num_layers, update_running=True):
    """M:463-482 _build_mlp products: Linear(no bias) -> BN -> ReLU ..., last BN affine=False.
    Sequential indices: layer l has Linear at 3l and BN at 3l+1.  Running stats are updated
    in place in `sd` as nn.BatchNorm1d does (momentum 0.1, unbiased variance)."""
    for l in range(num_layers):
        x = Fn.linear(x, sd["%s%d.weight" % (prefix, 3 * l)])
        bn = "%s%d." % (prefix, 3 * l + 1)
        last = l == num_layers - 1
        w = None if last else sd[bn + "weight"]
        b = None if last else sd[bn + "bias"]
        x, mean, var, n = bn_train(x, w, b)
        if update_running:
            with torch.no_grad():
                unbiased = var * (n / max(n - 1, 1))
                sd[bn + "running_mean"].mul_(1 - BN_MOMENTUM).add_(mean.detach(), alpha=BN_MOMENTUM)
                sd[bn + "running_var"].mul_(1 - BN_MOMENTUM).add_(unbiased.detach(), alpha=BN_MOMENTUM)
                sd[bn + "num_batches_tracked"] += 1
        if not last:
            x = torch.relu(x)
    return x


def patch_pool(x, num_windows=4):
    """M:189-193 PatchNet, use_patch_transformer=False: [S,256,d] -> [S,num_windows,d], the mean
    over all 8 grid rows and each group of 32/num_windows grid columns."""
    S, _, C = x.shape
    x = x.reshape(S, GRID_H, GRID_W, C).permute(0, 3, 1, 2)
    x = Fn.adaptive_avg_pool2d(x, (1, num_windows)).permute(0, 2, 3, 1).squeeze(1)
    return x


EMA_PAIRS = (("encoder.", "momentum_encoder."),
             ("encoder_projection_layer.", "momentum_projection_layer."),
             ("pix_projector.", "pix_projector_m."))


def is_buffer(name):
    return name.endswith("running_mean") or name.endswith("running_var") or name.endswith("num_batches_tracked")


def ema_update(sd, m):
    """M:428-442 p_m = m*p_m + (1-m)*p over parameters (not BN buffers) of the three pairs;
    patch_extractor has no parameters with patchnet_name='no_patchtrans'."""
    with torch.no_grad():
        for name in list(sd.keys()):
            for src, dst in EMA_PAIRS:
                if name.startswith(src) and not is_buffer(name):
                    k = dst + name[len(src):]
                    sd[k] = sd[k] * m + sd[name].detach() * (1.0 - m)


def contrastive_loss(q, k_all, T, rank=0):
    """M:444-461 + M:593-625: normalise, logits = q k^T / T, labels arange(N)+N*rank, CE * 2T,
    top-1 / top-5 accuracy in percent.  `k_all` is the all_gather of the *normalised* keys."""
    q = Fn.normalize(q, dim=1)
    logits = q @ k_all.t() / T
    N = logits.shape[0]
    labels = torch.arange(N, dtype=torch.long, device=q.device) + N * rank
    logp = logits.log_softmax(dim=1)
    loss = -logp.gather(1, labels[:, None]).squeeze(1).mean() * (2 * T)
    with torch.no_grad():
        _, pred = logits.topk(min(5, logits.shape[1]), 1, True, True)
        correct = pred.eq(labels[:, None])
        acc1 = correct[:, :1].float().sum().mul(100.0 / N).reshape(1)
        acc5 = correct[:, :5].float().sum().mul(100.0 / N).reshape(1)
    return loss, acc1, acc5


def pix_decoder(sd, x):
    """M:422-426: Linear(d,192,no bias) -> Linear(192,192,no bias) -> LN(192,1e-6) -> GELU -> Linear(192,48)."""
    x = Fn.linear(x, sd["pix_decoder.0.weight"])
    x = Fn.linear(x, sd["pix_decoder.1.weight"])
    x = Fn.layer_norm(x, (x.shape[-1],), sd["pix_decoder.2.weight"], sd["pix_decoder.2.bias"], LN_EPS)
    x = Fn.gelu(x)
    return Fn.linear(x, sd["pix_decoder.4.weight"], sd["pix_decoder.4.bias"])


def moco_vit_forward(sd, image, aug_image, vis_mask_pos, m, heads, T=0.2, num_windows=4,
                     only_mim_on_ori_img=True, rank=0, gather=None, taps=None):
    """M:488-577.  `sd` is mutated: momentum parameters (EMA, M:526) and BN running stats.
    vis_mask_pos: bool [B, num_view, 256].  gather: callable all-gathering normalised keys over
    ranks (identity at W=1).  Returns the reference's out_dict plus a few taps for tests."""
    out = OrderedDict()
    all_images = torch.cat([image, aug_image], dim=0)                                   # M:491
    num_view = vis_mask_pos.size(1)
    mask = vis_mask_pos.permute(1, 0, 2).reshape(-1, vis_mask_pos.size(-1))            # M:496-497
    B = image.shape[0]

    enc = encoder_forward(sd, "encoder.", all_images, mask, heads)                      # M:502
    masked_o, aug_o = enc.chunk(2, dim=0)
    b, l, c = masked_o.shape
    masked_p = bn_mlp(sd, "pix_projector.", masked_o.reshape(b * l, c), 3).reshape(b, l, c)   # M:505
    enc_cat = torch.cat([masked_p, aug_o], dim=0)                                       # M:507
    patches = patch_pool(enc_cat, num_windows)                                          # M:513
    S, L, C = patches.shape
    qs = bn_mlp(sd, "encoder_projection_layer.", patches.reshape(S * L, C), 3)          # M:517
    qs = bn_mlp(sd, "predictor.", qs, 2)                                                # M:518
    q1, q2 = qs.reshape(S, L, -1).chunk(2, dim=0)
    q1 = q1.reshape(-1, q1.size(-1))
    q2 = q2.reshape(-1, q2.size(-1))

    with torch.no_grad():                                                               # M:525-549
        ema_update(sd, m)
        enc_m = encoder_forward(sd, "momentum_encoder.", all_images, mask, heads)
        masked_m, aug_m = enc_m.chunk(2, dim=0)
        masked_mp = bn_mlp(sd, "pix_projector_m.", masked_m.reshape(b * l, c), 3).reshape(b, l, c)
        mom_cat = torch.cat([masked_mp, aug_m], dim=0)
        mpatches = patch_pool(mom_cat, num_windows)
        ks = bn_mlp(sd, "momentum_projection_layer.", mpatches.reshape(S * L, C), 3)
        k1, k2 = ks.reshape(S, L, -1).chunk(2, dim=0)
        k1 = Fn.normalize(k1.reshape(-1, k1.size(-1)), dim=1)
        k2 = Fn.normalize(k2.reshape(-1, k2.size(-1)), dim=1)
        if gather is not None:
            k1, k2 = gather(k1), gather(k2)

    l1, a11, a15 = contrastive_loss(q1, k2, T, rank)                                    # M:551
    l2, a21, a25 = contrastive_loss(q2, k1, T, rank)                                    # M:552
    out["contra_loss"] = l1 + l2
    out["q1_acc1"], out["q1_acc5"], out["q2_acc1"], out["q2_acc5"] = a11, a15, a21, a25

    # M:561-575.  The reference decodes all 2B*256 rows and then gathers; the decoder is
    # row-wise, so gathering first is bit-identical.  We keep the reference order.
    dec = pix_decoder(sd, enc)
    Cd = dec.shape[-1]
    dec_list = list(dec.chunk(num_view, dim=0))
    mask_list = list(mask.chunk(num_view, dim=0))
    if only_mim_on_ori_img:
        out["vis_out"] = [dec_list[0][mask_list[0]].reshape(B, -1, Cd)]
    else:
        out["vis_out"] = [d_[m_].reshape(B, -1, Cd) for d_, m_ in zip(dec_list, mask_list)]
    if taps is not None:
        taps.update(enc=enc, enc_m=enc_m, q1=q1, q2=q2, k1=k1, k2=k2)
    return out


def build_targets(images, mask_bvn, only_mim_on_ori_img=True, patch_size=PATCH, normalize_target=False):
    """E:83-111: un-normalise, patchify '(p1 p2 c)', masked gather; normalize_target = the `normlize_target` branch (E:89-94): every
    colour plane of a patch standardised over its p1*p2 pixels (mean, unbiased variance, + 1e-6 on the standard deviation).
    mask_bvn: bool [B, num_view, 256] (already zeroed for view 1 if only_mim_on_ori_img).
    Returns list of [B, n_masked, 48]."""
    unnorm = images * 0.5 + 0.5
    B, C, H, W = unnorm.shape
    h, w = H // patch_size, W // patch_size
    p = unnorm.reshape(B, C, h, patch_size, w, patch_size).permute(0, 2, 4, 3, 5, 1)            # b h w p1 p2 c
    if normalize_target:
        sq = p.reshape(B, h * w, patch_size * patch_size, C)                                   # 'b (h w) (p1 p2) c'
        p = (sq - sq.mean(dim=-2, keepdim=True)) / (sq.var(dim=-2, unbiased=True, keepdim=True).sqrt() + 1e-6)
    patches = p.reshape(B, h * w, patch_size * patch_size * C)
    views = 1 if only_mim_on_ori_img else mask_bvn.shape[1]
    return [patches[mask_bvn[:, i, :]].reshape(B, -1, patches.shape[-1]) for i in range(views)]


def step_losses(sd, images, aug_images, mask_bvn, m, heads, w_contrast=0.1, w_pixel=1.0, T=0.2,
                num_windows=4, rank=0, gather=None, taps=None, only_mim_on_ori_img=True):
    """E:76-146 for one batch: zero view-1 mask (only_mim_on_ori_img), targets, forward, weighted loss; with only_mim_on_ori_img False
    every view's decoder output is compared with the masked patches of the ORIGINAL image, weighted 1/num_view (E:137-141)."""
    mask_bvn = mask_bvn.clone()
    if only_mim_on_ori_img:
        mask_bvn[:, 1, :] = False                                                       # E:103-104
    labels = build_targets(images, mask_bvn, only_mim_on_ori_img)
    out = moco_vit_forward(sd, images, aug_images, mask_bvn, m, heads, T, num_windows, only_mim_on_ori_img, rank,
                           gather, taps)
    loss_pixel = sum(Fn.mse_loss(o, l, reduction="mean") for o, l in zip(out["vis_out"], labels)) / len(labels)   # E:137-141
    loss = out["contra_loss"] * w_contrast + loss_pixel * w_pixel                       # E:122, E:143
    return loss, out, loss_pixel


def trainable_names(sd):
    """requires_grad=True parameters: everything except momentum copies (M:399-420) and buffers."""
    frozen = ("momentum_encoder.", "momentum_projection_layer.", "pix_projector_m.")
    return [k for k in sd if not is_buffer(k) and not k.startswith(frozen)]


def weight_decay_of(name, shape, weight_decay):
    """optim_factory.py:57-100 get_parameter_groups: 1-D / .bias / skip-list -> no decay."""
    if len(shape) == 1 or name.endswith(".bias") or name in ("pos_embed", "cls_token"):
        return 0.0
    return weight_decay


def adamw_step(param, grad, exp_avg, exp_avg_sq, step, lr, weight_decay, beta1=0.9, beta2=0.999, eps=1e-8):
    """custom_optim/_functional.py:115-140 (amsgrad False). In place on param and moments."""
    param.mul_(1 - lr * weight_decay)
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    exp_avg.mul_(beta1).add_(grad, alpha=1 - beta1)
    exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
    denom = (exp_avg_sq.sqrt() / math.sqrt(bc2)).add_(eps)
    param.addcdiv_(exp_avg, denom, value=-(lr / bc1))


def grad_norm(grads):
    """U:507-519 get_grad_norm_: L2 norm of per-tensor L2 norms."""
    return torch.norm(torch.stack([torch.norm(g.detach(), 2.0) for g in grads]), 2.0)


class OracleTrainer:
    """Drives whole training steps on a state dict: forward, autograd backward, grad-norm, AdamW.
    Restates E:58-199 + U:483-498 without AMP (fp32)."""

    def __init__(self, sd, heads, lr=1.5e-4, weight_decay=0.05, T=0.2, num_windows=4,
                 w_contrast=0.1, w_pixel=1.0, betas=(0.9, 0.999), eps=1e-8):
        self.sd = sd
        self.heads = heads
        self.lr, self.wd, self.T, self.nw = lr, weight_decay, T, num_windows
        self.wc, self.wp = w_contrast, w_pixel
        self.betas, self.eps = betas, eps
        self.names = trainable_names(sd)
        self.state = {n: (torch.zeros_like(sd[n]), torch.zeros_like(sd[n])) for n in self.names}
        self.t = 0

    def step(self, images, aug_images, mask_bvn, m, lr=None, weight_decay=None):
        sd = self.sd
        for n in self.names:
            sd[n] = sd[n].detach().requires_grad_(True)
        loss, out, loss_pixel = step_losses(sd, images, aug_images, mask_bvn, m, self.heads, self.wc,
                                            self.wp, self.T, self.nw)
        grads = torch.autograd.grad(loss, [sd[n] for n in self.names], allow_unused=True)
        grads = [g if g is not None else torch.zeros_like(sd[n]) for g, n in zip(grads, self.names)]
        gn = grad_norm(grads)
        self.t += 1
        lr = self.lr if lr is None else lr
        wd = self.wd if weight_decay is None else weight_decay
        with torch.no_grad():
            for n, g in zip(self.names, grads):
                p = sd[n].detach()
                ea, eas = self.state[n]
                adamw_step(p, g, ea, eas, self.t, lr, weight_decay_of(n, p.shape, wd), self.betas[0],
                           self.betas[1], self.eps)
                sd[n] = p
        return {"loss": float(loss), "loss_contrast": float(out["contra_loss"]),
                "loss_pixel": float(loss_pixel), "grad_norm": float(gn),
                "q1_acc1": float(out["q1_acc1"]), "q1_acc5": float(out["q1_acc5"]),
                "q2_acc1": float(out["q2_acc1"]), "q2_acc5": float(out["q2_acc5"])}, grads, out


def synthetic_batch(B, seed=1, device="cpu", mask_ratio=0.7):
    """SURVEY 8(d) synthetic inputs: images/aug ~ U(-1,1) fp32 [B,3,32,128]; mask = per (sample,
    view) randperm(256)[:179] set, bool [B,2,256] (masking_generator.py:20,29-46)."""
    g = torch.Generator().manual_seed(seed)
    img = torch.rand(B, 3, IMG_H, IMG_W, generator=g) * 2 - 1
    aug = torch.rand(B, 3, IMG_H, IMG_W, generator=g) * 2 - 1
    n_mask = int(mask_ratio * NUM_PATCHES)
    mask = torch.zeros(B, 2, NUM_PATCHES, dtype=torch.bool)
    for b in range(B):
        for v in range(2):
            mask[b, v, torch.randperm(NUM_PATCHES, generator=g)[:n_mask]] = True
    return img.to(device), aug.to(device), mask.to(device)

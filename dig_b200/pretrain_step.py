"""Sequencing of DiG's pre-training forward/backward on the sm_100a kernels (boundary B1 -> B3).

`PretrainStep.forward` reproduces `MoCo_ViT.forward` (reference modeling_pretrain_moco_mim_ori.py:488-577, tag M)
and `PretrainStep.backward` its autograd, but every arithmetic step is a C-ABI call into libdig_b200.so
(dig_b200/ops.py).  torch is used for what the brief allows: owning device buffers, the current stream,
autograd plumbing (one custom Function for the whole model) and torch.distributed collectives (MoCo key
all_gather M:580-591, SyncBatchNorm statistics R:390).

Activation layout in HBM (tokens = 2B sequences x 256 patch tokens, view-major: rows [0, B*256) are the masked
view, rows [B*256, 2B*256) the augmented view, exactly torch.cat([image, aug_image]) of M:491):
  * residual stream           fp32 [tokens, d]     one buffer per half-block (kept for the backward)
  * GEMM operands             bf16 [tokens, d|3d|4d] (LayerNorm outputs, qkv, attention context, GELU in/out)
  * weights                   fp32 master nn.Parameters + bf16 shadows refreshed once per step
  * head activations          fp32 pre-BatchNorm, bf16 post-BatchNorm (GEMM operands)
"""
import math
import os

import torch
import torch.distributed as dist

from . import dist_layout, ops, peer
from .ops import call

F32, BF16 = torch.float32, torch.bfloat16
TOK = 256


def _world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(), dist.get_rank()
    return 1, 0


class MtTable:
    """Device-side pointer table for the multi-tensor kernels (dig_b200/csrc/optim.cu)."""

    def __init__(self, device, *tensor_lists):
        n = len(tensor_lists[0])
        chunk = ops.load().dig_mt_chunk()
        self.keep = tensor_lists
        self.ptrs = [torch.tensor([0 if t is None else t.data_ptr() for t in lst], dtype=torch.int64, device=device)
                     for lst in tensor_lists]
        numel = [tensor_lists[0][i].numel() for i in range(n)]
        self.numel = torch.tensor(numel, dtype=torch.int64, device=device)
        bt, bc = [], []
        for i, ne in enumerate(numel):
            for c in range((ne + chunk - 1) // chunk):
                bt.append(i)
                bc.append(c)
        self.blk_tensor = torch.tensor(bt, dtype=torch.int32, device=device)
        self.blk_chunk = torch.tensor(bc, dtype=torch.int32, device=device)
        self.num_blocks = len(bt)
        self.signature = tuple(0 if t is None else t.data_ptr() for lst in tensor_lists for t in lst)


class _Bufs:
    """Named, shape-checked device buffers (allocated once per batch size)."""

    def __init__(self, device):
        self.device = device
        self.d = {}
        self.zero_lists = {}

    def get(self, name, shape, dtype):
        t = self.d.get(name)
        shape = tuple(int(s) for s in shape)
        if t is None or tuple(t.shape) != shape or t.dtype != dtype:
            t = torch.empty(shape, dtype=dtype, device=self.device)
            self.d[name] = t
            self.zero_lists = {k: [z for z in v if z[0] != name] for k, v in self.zero_lists.items()}
        return t

    def zeroed(self, name, shape, dtype, phase):
        """A small accumulator buffer that must be zero when its consumer runs.  The first time it is zeroed on the spot; from then on
        `zero_phase(phase)` clears every accumulator of the phase ("fwd" / "bwd") with ONE multi-tensor launch at the start of the
        phase instead of one memset per layer (41 launches per step before)."""
        t = self.get(name, shape, dtype)
        lst = self.zero_lists.setdefault(phase, [])
        if not any(z[0] == name for z in lst):
            t.zero_()
            lst.append((name, t))
        return t

    def zero_phase(self, phase):
        lst = self.zero_lists.get(phase)
        if lst:
            torch._foreach_zero_([t for _, t in lst])


class PretrainStep:
    GEMM_WEIGHTS_BLOCK = ("attn.qkv.weight", "attn.proj.weight", "mlp.fc1.weight", "mlp.fc2.weight")

    def __init__(self, model):
        ops.load()
        p0 = next(model.parameters())
        if not p0.is_cuda:
            raise ops.DigError("dig_b200 runs on sm_100a only: move the model to a CUDA device first")
        self.model = model
        self.device = p0.device
        enc = model.encoder
        self.d = enc.embed_dim
        self.heads = enc.num_heads
        if self.d != self.heads * 64:
            raise ops.DigError("head_dim must be 64 (d=%d heads=%d)" % (self.d, self.heads))
        self.depth = len(enc.blocks)
        for mod in model.modules():
            if isinstance(mod, torch.nn.modules.batchnorm._BatchNorm) and mod.track_running_stats and mod.momentum is None:
                raise ops.DigError("BatchNorm momentum=None (cumulative moving average) is not built; the reference uses the default 0.1 (M:463-482)")
        self.T = float(model.T)
        self.num_windows = int(model.num_windows)
        self.scale = 64 ** -0.5
        self.bufs = _Bufs(self.device)
        self.pos = enc.pos_embed.reshape(TOK, self.d).to(self.device, F32).contiguous()
        self._named = dict(model.named_parameters())
        self._ptr_sig = self._signature()
        self._build_shadows()
        self.saved = None
        self.forward_serial = 0
        self._n_masked = {}
        self._mask_err = None      # (event, pinned int32[1]) of the previous forward's dig_mask_to_index error flag
        self._peer = None
        self._grad_flat = None
        self._grad_peer = None      # table of every rank's flat gradient buffer when it lives in IPC-mapped memory
        self._comm_stream = None    # stream of the gradient exchanges that overlap the backward
        # The momentum branch (EMA update + no-grad forward) and the online branch are independent until the InfoNCE logits: they run
        # on two streams so that one branch's HBM-bound kernels (LayerNorm, BatchNorm, casts) fill in under the other's GEMMs.
        import os
        self._two_streams = os.environ.get("DIG_TWO_STREAMS", "1") != "0"
        self._always_cast = os.environ.get("DIG_ALWAYS_CAST", "0") == "1"
        # The running gradient of the residual stream through the encoder backward is a bf16 stream (each LayerNorm backward reads it and
        # writes the next one; it is the dgrad GEMM operand anyway) -- DIG_BF16_GRAD_STREAM=0 keeps the fp32 copy alongside (round 1).
        self._bf16_grad_stream = os.environ.get("DIG_BF16_GRAD_STREAM", "1") != "0"
        # The GELU pre-activation is kept for the backward as 8-bit codes (include/dig_b200.h, dig_gemm_t.aux_q8): 100 instead of 200 MB
        # written by fc1 and read by the GELU' dgrad per block -- DIG_GELU_Q8=0 keeps the bf16 copy (round 1).
        self._gelu_q8 = os.environ.get("DIG_GELU_Q8", "1") != "0"
        self._side = torch.cuda.Stream(device=self.device) if self._two_streams else None

    # ------------------------------------------------------------------ parameter bookkeeping
    def _signature(self):
        return tuple(p.data_ptr() for p in self._named.values())

    def matches(self, model):
        return model is self.model and self._signature() == self._ptr_sig

    def _gemm_weight_names(self, prefix_enc, prefix_proj, prefix_pix, predictor):
        names = [prefix_enc + "patch_embed.proj.weight"]
        for l in range(self.depth):
            names += ["%sblocks.%d.%s" % (prefix_enc, l, w) for w in self.GEMM_WEIGHTS_BLOCK]
        names += ["%s%d.weight" % (prefix_proj, i) for i in (0, 3, 6)]
        names += ["%s%d.weight" % (prefix_pix, i) for i in (0, 3, 6)]
        if predictor:
            names += ["predictor.0.weight", "predictor.3.weight", "pix_decoder.0.weight", "pix_decoder.1.weight",
                      "pix_decoder.4.weight"]
        return names

    def _build_shadows(self):
        """bf16 shadows of every GEMM weight (online + momentum), views into one flat buffer (16-byte aligned)."""
        on = self._gemm_weight_names("encoder.", "encoder_projection_layer.", "pix_projector.", True)
        mo = self._gemm_weight_names("momentum_encoder.", "momentum_projection_layer.", "pix_projector_m.", False)
        total = 0
        offs = {}
        for n in on + mo:
            offs[n] = total
            total += (self._named[n].numel() + 7) // 8 * 8
        flat = torch.zeros(total, dtype=BF16, device=self.device)
        self.shadow = {}
        for n in on + mo:
            p = self._named[n]
            shape = (p.shape[0], p.numel() // p.shape[0])
            self.shadow[n] = flat[offs[n]:offs[n] + p.numel()].view(shape)
        self.shadow_flat = flat
        self.tab_cast_online = MtTable(self.device, [self._named[n].data for n in on], [self.shadow[n] for n in on])
        for n in on:
            ops.register_shadow(self._named[n], self.shadow[n])     # FusedAdamW refreshes these in its own launch
        self._online_gemm_params = [self._named[n] for n in on]
        self._cast_state = None
        self.tab_cast_momentum = MtTable(self.device, [self._named[n].data for n in mo], [self.shadow[n] for n in mo])
        # EMA pairs: every parameter of encoder / projector / pix_projector (M:428-442)
        pairs = []
        for src, dst in (("encoder.", "momentum_encoder."), ("encoder_projection_layer.", "momentum_projection_layer."),
                         ("pix_projector.", "pix_projector_m.")):
            for n, p in self._named.items():
                if n.startswith(src):
                    pairs.append((p.data, self._named[dst + n[len(src):]].data, self.shadow.get(dst + n[len(src):])))
        self.tab_ema = MtTable(self.device, [a for a, _, _ in pairs], [b for _, b, _ in pairs], [c for _, _, c in pairs])
        # fused qkv bias [q_bias | 0 | v_bias] per block (F:91) for both encoders
        d, L = self.d, self.depth
        self.qkv_bias = {"encoder.": torch.zeros(L, 3 * d, device=self.device), "momentum_encoder.": torch.zeros(L, 3 * d, device=self.device)}
        self.tab_qkv_bias = {}
        for pre in ("encoder.", "momentum_encoder."):
            src, dst = [], []
            for l in range(L):
                src += [self._named["%sblocks.%d.attn.q_bias" % (pre, l)].data, self._named["%sblocks.%d.attn.v_bias" % (pre, l)].data]
                dst += [self.qkv_bias[pre][l, :d], self.qkv_bias[pre][l, 2 * d:]]
            self.tab_qkv_bias[pre] = MtTable(self.device, src, dst)
        # trainable parameters, in named_parameters order, and their flat gradient buffer
        self.train_names = [n for n, p in self._named.items() if p.requires_grad]
        goff, total = {}, 0
        for n in self.train_names:
            goff[n] = total
            total += (self._named[n].numel() + 3) // 4 * 4
        self.grad_off, self.grad_total = goff, total
        from .parallel import grad_segments
        self.grad_seg = grad_segments(self.train_names, goff, total)     # heads / block l / embed -> [start, end) of the flat buffer
        self.momentum_warm = False

    def trainable_params(self):
        return [self._named[n] for n in self.train_names]

    def _grad_views(self, flat):
        out = {}
        for n in self.train_names:
            p = self._named[n]
            out[n] = flat[self.grad_off[n]:self.grad_off[n] + p.numel()].view(p.shape)
        return out

    # ------------------------------------------------------------------ small helpers
    def _mt(self, fn, tab, *extra):
        lib = ops.load()
        rc = getattr(lib, fn)(*[t.data_ptr() for t in tab.ptrs], tab.numel.data_ptr(), tab.blk_tensor.data_ptr(),
                              tab.blk_chunk.data_ptr(), tab.num_blocks, *extra, torch.cuda.current_stream().cuda_stream)
        ops.count_launch()
        if rc != 0:
            raise ops.DigError("%s failed: %s" % (fn, lib.dig_last_error().decode()))

    def _bn_sync(self, bn):
        w, _ = _world()
        return w > 1 and isinstance(bn, torch.nn.SyncBatchNorm)

    def _ln(self, x, w, b, y, mean, rstd, gelu=0, eps=1e-6):
        call("dig_layernorm_fwd", x, w, b, y, mean, rstd, x.shape[0], x.shape[1], eps, gelu)

    # ------------------------------------------------------------------ encoder
    def _enc_weights(self, pre):
        N, S = self._named, self.shadow
        blocks = []
        for l in range(self.depth):
            b = "%sblocks.%d." % (pre, l)
            blocks.append(dict(n1w=N[b + "norm1.weight"], n1b=N[b + "norm1.bias"], qkvw=S[b + "attn.qkv.weight"],
                               qkvb=self.qkv_bias[pre][l], pw=S[b + "attn.proj.weight"], pb=N[b + "attn.proj.bias"],
                               n2w=N[b + "norm2.weight"], n2b=N[b + "norm2.bias"], f1w=S[b + "mlp.fc1.weight"],
                               f1b=N[b + "mlp.fc1.bias"], f2w=S[b + "mlp.fc2.weight"], f2b=N[b + "mlp.fc2.bias"], name=b))
        return dict(pew=S[pre + "patch_embed.proj.weight"], peb=N[pre + "patch_embed.proj.bias"],
                    mtok=N[pre + "mask_token"].reshape(-1), blocks=blocks, pre=pre)

    def _encoder_fwd(self, W, images, mask_u8, tag, save):
        """V:89-106. images fp32 [S,3,32,128]; mask_u8 [S*256]. Returns fp32 [S*256, d]; saves activations when `save`."""
        B = self.bufs
        S = images.shape[0]
        M, d, h = S * TOK, self.d, self.heads
        a0 = B.get(tag + "a0", (M, 48), BF16)
        call("dig_im2col_patch4", images, a0, S)
        x = B.get(tag + "x0", (M, d), F32)
        ops.gemm(a0, W["pew"], x, bias=W["peb"], residual=self.pos, res_row_mod=TOK, row_mask=mask_u8, row_mask_value=W["mtok"])
        acts = []
        for l, bw in enumerate(W["blocks"]):
            t = (tag + "b%d." % l) if save else (tag + "b.")
            ln1 = B.get(t + "ln1", (M, d), BF16)
            mean1, rstd1 = B.get(t + "m1", (M,), F32), B.get(t + "r1", (M,), F32)
            self._ln(x, bw["n1w"], bw["n1b"], ln1, mean1, rstd1)
            qkv = B.get(t + "qkv", (M, 3 * d), BF16)
            ops.gemm(ln1, bw["qkvw"], qkv, bias=bw["qkvb"])
            att = B.get(t + "att", (M, d), BF16)
            lse = B.get(t + "lse", (S, h, TOK), F32) if save else None      # row log-sum-exp: only the backward reads it
            ops.attention_fwd(qkv, att, lse, h, self.scale)
            xm = B.get(t + "xm" if save else tag + "xm", (M, d), F32)
            ops.gemm(att, bw["pw"], xm, bias=bw["pb"], residual=x)
            ln2 = B.get(t + "ln2", (M, d), BF16)
            mean2, rstd2 = B.get(t + "m2", (M,), F32), B.get(t + "r2", (M,), F32)
            self._ln(xm, bw["n2w"], bw["n2b"], ln2, mean2, rstd2)
            hpre = B.get(t + "hpre", (M, 4 * d), torch.uint8 if self._gelu_q8 else BF16) if save else None     # GELU pre-activation: only the backward reads it
            hpost = B.get(t + "hpost", (M, 4 * d), BF16)
            ops.gemm(ln2, bw["f1w"], hpost, bias=bw["f1b"], epilogue=ops.EPI_GELU, aux=hpre)
            xn = B.get((tag + "x%d" % (l + 1)) if save else (tag + "x%d" % ((l + 1) % 2 + 1)), (M, d), F32)
            ops.gemm(hpost, bw["f2w"], xn, bias=bw["f2b"], residual=xm)
            if save:
                acts.append(dict(x=x, ln1=ln1, mean1=mean1, rstd1=rstd1, qkv=qkv, att=att, lse=lse, xm=xm, ln2=ln2, mean2=mean2,
                                 rstd2=rstd2, hpre=hpre, hpost=hpost))
            x = xn
        return x, dict(a0=a0, acts=acts, mask=mask_u8)

    def _encoder_bwd(self, W, sv, g, gb, grads):
        """g fp32 / gb bf16 [M,d]: gradient w.r.t. the encoder output.  Accumulates parameter gradients into `grads`.

        The activation-gradient chain (GELU' dgrad -> fc1 dgrad -> LN2 bwd -> proj dgrad -> attention bwd -> qkv dgrad -> LN1 bwd) runs on
        the current stream; the weight-gradient GEMMs and the qkv bias column sums are off that chain and go to the side stream, so they
        fill in next to the chain's HBM-bound kernels.  The scratch tensors both streams touch (gb, dh, dqkv) rotate through small rings;
        before the chain overwrites a ring slot it waits for the event its last side-stream reader recorded."""
        B = self.bufs
        M, d, h = g.shape[0], self.d, self.heads
        cur = torch.cuda.current_stream()
        side = self._side if self._two_streams else cur
        two = side is not cur
        ring = {"gb": 4, "dh": 2, "dqkv": 2} if two else {"gb": 1, "dh": 1, "dqkv": 1}
        shapes = {"gb": (M, d), "dh": (M, 4 * d), "dqkv": (M, 3 * d)}
        bufs = {k: [B.get("bw.%s%d" % (k, i), shapes[k], BF16) for i in range(n)] for k, n in ring.items()}
        busy = {k: [None] * n for k, n in ring.items()}    # event recorded on the side stream after the slot's last reader there
        nxt = {k: 0 for k in ring}

        def take(kind):
            """Next ring slot of `kind` for the chain to write: waits for its pending side-stream readers."""
            i = nxt[kind]
            nxt[kind] = (i + 1) % ring[kind]
            if busy[kind][i] is not None:
                cur.wait_event(busy[kind][i])
                busy[kind][i] = None
            return i, bufs[kind][i]

        def on_side(fn, reads):
            """Run fn() on the side stream once everything enqueued on the chain so far is done; mark the ring slots it reads busy."""
            if not two:
                fn()
                return
            ev = torch.cuda.Event()
            ev.record(cur)
            side.wait_event(ev)
            with torch.cuda.stream(side):
                fn()
                done = torch.cuda.Event()
                done.record(side)
            for kind, i in reads:
                busy[kind][i] = done

        if two:
            side.wait_stream(cur)          # zeroed gradient buffer, head gradients
        gi, gb0 = take("gb")
        gb0.copy_(gb)                       # incoming gradient into ring slot 0 (the caller's buffer stays untouched)
        gb = gb0
        dln = B.get("bw.dln", (M, d), BF16)
        dat = B.get("bw.dat", (M, d), BF16)
        dsum = B.get("bw.dsum", (M, h), F32)
        # fc2 bias gradient of the last block = column sums of the incoming gradient; for the other blocks it falls out of
        # the LayerNorm-1 backward of the block above (dxsum), like the proj bias gradient out of the LayerNorm-2 backward.
        call("dig_colsum", gb, 0, d, grads[W["blocks"][-1]["name"] + "mlp.fc2.bias"], None, M, d)
        for l in reversed(range(self.depth)):
            bw, a = W["blocks"][l], sv["acts"][l]
            nm = bw["name"]
            # ---- MLP (F:53-60) ----
            hi, dh = take("dh")
            ops.gemm(gb, bw["f2w"], dh, b_mn_major=True, epilogue=ops.EPI_GELU_BWD, aux=a["hpre"], colsum=grads[nm + "mlp.fc1.bias"])

            def mlp_wgrads(gb=gb, dh=dh, a=a, nm=nm):
                ops.gemm(gb, a["hpost"], grads[nm + "mlp.fc2.weight"], a_mn_major=True, b_mn_major=True, split_k=-1)
                ops.gemm(dh, a["ln2"], grads[nm + "mlp.fc1.weight"], a_mn_major=True, b_mn_major=True, split_k=-1)
            on_side(mlp_wgrads, [("gb", gi), ("dh", hi)])
            ops.gemm(dh, bw["f1w"], dln, b_mn_major=True)
            gprev = gb                      # the running residual gradient (bf16 stream): read by the LayerNorm backward, then dead
            gi, gb = take("gb")
            if self._bf16_grad_stream:
                call("dig_layernorm_bwd_bf16res", dln, a["xm"], a["mean2"], a["rstd2"], bw["n2w"], gprev, gb,
                     grads[nm + "norm2.weight"], grads[nm + "norm2.bias"], grads[nm + "attn.proj.bias"], M, d)
            else:
                call("dig_layernorm_bwd", dln, a["xm"], a["mean2"], a["rstd2"], bw["n2w"], None, g, g, gb,
                     grads[nm + "norm2.weight"], grads[nm + "norm2.bias"], grads[nm + "attn.proj.bias"], M, d, 0)
            # ---- attention (F:87-125) ----
            # output-projection dgrad; its epilogue also emits D = rowsum(dO o O) per (token, head) for the attention backward
            ops.gemm(gb, bw["pw"], dat, b_mn_major=True, epilogue=ops.EPI_ROWDOT, aux=a["att"], rowdot=dsum)

            def proj_wgrad(gb=gb, a=a, nm=nm):
                ops.gemm(gb, a["att"], grads[nm + "attn.proj.weight"], a_mn_major=True, b_mn_major=True, split_k=-1)
            on_side(proj_wgrad, [("gb", gi)])
            qi, dqkv = take("dqkv")
            ops.attention_bwd_d(a["qkv"], dat, a["lse"], dsum, dqkv, h, self.scale)

            def qkv_wgrads(dqkv=dqkv, a=a, nm=nm, l=l):
                dqkvb = B.zeroed("bw.dqkvb%d" % l, (3 * d,), F32, "bwd")
                call("dig_colsum", dqkv, 0, 3 * d, dqkvb, None, M, 3 * d)
                grads[nm + "attn.q_bias"].copy_(dqkvb[:d])
                grads[nm + "attn.v_bias"].copy_(dqkvb[2 * d:])
                ops.gemm(dqkv, a["ln1"], grads[nm + "attn.qkv.weight"], a_mn_major=True, b_mn_major=True, split_k=-1)
            on_side(qkv_wgrads, [("dqkv", qi)])
            ops.gemm(dqkv, bw["qkvw"], dln, b_mn_major=True)
            prev_b2 = grads[W["blocks"][l - 1]["name"] + "mlp.fc2.bias"] if l > 0 else None
            gprev = gb
            gi, gb = take("gb")
            if self._bf16_grad_stream:
                call("dig_layernorm_bwd_bf16res", dln, a["x"], a["mean1"], a["rstd1"], bw["n1w"], gprev, gb,
                     grads[nm + "norm1.weight"], grads[nm + "norm1.bias"], prev_b2, M, d)
            else:
                call("dig_layernorm_bwd", dln, a["x"], a["mean1"], a["rstd1"], bw["n1w"], None, g, g, gb,
                     grads[nm + "norm1.weight"], grads[nm + "norm1.bias"], prev_b2, M, d, 0)
            # block l is final once this LayerNorm backward (chain) and its weight gradients (side stream) are done
            self._allreduce_segment("block%d" % l, stream=side)
        if two:
            cur.wait_stream(side)
        # ---- patch embed + mask token (F:190-196, V:95-99); pos_embed carries no gradient (V:99 detach) ----
        pre = W["pre"]
        gz = B.get("bw.gz", (M, d), BF16)
        tot = B.zeroed("bw.tot", (d,), F32, "bwd")
        if self._bf16_grad_stream:
            call("dig_zero_masked_rows_bf16", gb, sv["mask"], gz, M, d)
            call("dig_colsum", gb, 0, d, tot, None, M, d)
        else:
            call("dig_zero_masked_rows", g, sv["mask"], gz, M, d)
            call("dig_colsum", g, 1, d, tot, None, M, d)
        call("dig_colsum", gz, 0, d, grads[pre + "patch_embed.proj.bias"], None, M, d)
        torch.sub(tot, grads[pre + "patch_embed.proj.bias"], out=grads[pre + "mask_token"].view(-1))
        ops.gemm(gz, sv["a0"], grads[pre + "patch_embed.proj.weight"].view(d, 48), a_mn_major=True, b_mn_major=True,
                 split_k=-1)

    # ------------------------------------------------------------------ BatchNorm MLP heads (M:463-482)
    def _mlp_layers(self, prefix, num_layers, seq):
        return [(self.shadow["%s%d.weight" % (prefix, 3 * i)], seq[3 * i + 1], "%s%d.weight" % (prefix, 3 * i),
                 "%s%d." % (prefix, 3 * i + 1)) for i in range(num_layers)]

    def _mlp_fwd(self, x_bf16, layers, tag, want_bf16_out=False):
        B = self.bufs
        rows = x_bf16.shape[0]
        world, _ = _world()
        saved, a = [], x_bf16
        out_f32 = None
        for li, (w, bn, _, _) in enumerate(layers):
            last = li == len(layers) - 1
            C = w.shape[0]
            z = B.get("%s.z%d" % (tag, li), (rows, C), F32)
            ops.gemm(a, w, z)
            stats = B.zeroed("%s.st%d" % (tag, li), (2 * C,), F32, "fwd")
            pc = self._peer if (self._bn_sync(bn) and 2 * C <= peer.MAX_FLOATS) else None
            if pc is not None:
                # SyncBatchNorm (R:390): column sums + their cross-rank sum in ONE kernel over NVLink peer memory (csrc/peer.cu)
                ch = peer.CH_MOMENTUM if tag.startswith("m.") else peer.CH_ONLINE
                call("dig_bn_stats_allreduce", z, stats, rows, C, pc.bases, pc.world, pc.rank, ch, pc.next_epoch(ch))
                count = float(rows * pc.world)
            else:
                call("dig_colsum", z, 1, C, stats, stats[C:], rows, C)
                count = dist_layout.sync_batch_stats(stats, rows) if self._bn_sync(bn) else float(rows)
            a_out = B.get("%s.a%d" % (tag, li), (rows, C), BF16) if (not last or want_bf16_out) else None
            out_f32 = B.get("%s.out" % tag, (rows, C), F32) if last else None
            gamma, beta = (bn.weight, bn.bias) if bn.affine else (None, None)
            call("dig_bn_apply", z, stats, count, gamma, beta, 0 if last else 1, bn.eps, a_out, out_f32, rows, C)
            if bn.track_running_stats and bn.running_mean is not None:
                call("dig_bn_running", stats, count, bn.momentum, bn.running_mean, bn.running_var, bn.num_batches_tracked, C)
            saved.append(dict(a_in=a, z=z, stats=stats, count=count, a_out=a_out))
            a = a_out
        return out_f32, a, saved

    def _mlp_bwd(self, dy, layers, saved, tag, grads, dx_out):
        """dy fp32 [rows, C_last] w.r.t. the last BatchNorm output.  Writes d input (fp32) into dx_out when given."""
        B = self.bufs
        rows = dy.shape[0]
        for li in reversed(range(len(layers))):
            w, bn, wname, bnname = layers[li]
            sv = saved[li]
            C = w.shape[0]
            bst = B.zeroed("%s.bst%d" % (tag, li), (2 * C,), F32, "bwd")
            pc = self._peer if (self._bn_sync(bn) and 2 * C <= peer.MAX_FLOATS) else None
            if pc is not None:
                # [sum dy | sum dy xhat] + cross-rank sum in one kernel; the rank-local sums are the BatchNorm bias / weight gradients
                call("dig_bn_bwd_stats_allreduce", dy, sv["z"], sv["stats"], sv["count"], bn.eps, bst,
                     grads[bnname + "bias"] if bn.affine else None, grads[bnname + "weight"] if bn.affine else None, rows, C,
                     pc.bases, pc.world, pc.rank, peer.CH_BACKWARD, pc.next_epoch(peer.CH_BACKWARD))
            else:
                call("dig_bn_bwd_stats", dy, sv["z"], sv["stats"], sv["count"], bn.eps, bst, rows, C)
                if bn.affine:
                    grads[bnname + "bias"].copy_(bst[:C])
                    grads[bnname + "weight"].copy_(bst[C:])
                if self._bn_sync(bn):
                    dist.all_reduce(bst)
            dz = B.get("%s.dz%d" % (tag, li), (rows, C), BF16)
            call("dig_bn_bwd_apply", dy, sv["z"], sv["stats"], bst, sv["count"], bn.weight if bn.affine else None, bn.eps, dz, None, rows, C)
            a_in = sv["a_in"]
            Cin = a_in.shape[1]
            ops.gemm(dz, a_in, grads[wname], a_mn_major=True, b_mn_major=True, split_k=-1)
            if li > 0:
                dy = B.get("%s.dy%d" % (tag, li), (rows, Cin), F32)
                ops.gemm(dz, w, dy, b_mn_major=True, epilogue=ops.EPI_RELU_MASK, aux=a_in)
            elif dx_out is not None:
                ops.gemm(dz, w, dx_out, b_mn_major=True)

    # ------------------------------------------------------------------ forward (M:488-577)
    def forward(self, image, aug_image, vis_mask_pos, m, only_mim_on_ori_img, need_grad=True):
        model, Bf = self.model, self.bufs
        dev, d = self.device, self.d
        Bsz = image.shape[0]
        S, M, half = 2 * Bsz, 2 * Bsz * TOK, Bsz * TOK
        world, rank = _world()
        if vis_mask_pos.dim() != 3 or vis_mask_pos.shape[0] != Bsz or vis_mask_pos.shape[-1] != TOK:
            raise ops.DigError("vis_mask_pos must be [B, num_view, 256], got %s" % (tuple(vis_mask_pos.shape),))
        if vis_mask_pos.shape[1] != 2:
            raise ops.DigError("num_view must be 2 (README.md:66), got %d" % vis_mask_pos.shape[1])
        self._check_mask_err()
        if world > 1 and isinstance(model.predictor[1], torch.nn.SyncBatchNorm):
            # collective on first use (every rank runs its first forward): NVLink peer-memory workspaces for the SyncBN / key exchanges;
            # key table = [k1 of all ranks ; k2 of all ranks], Q = B * num_windows rows per rank and half
            self._peer = peer.get(dev, 2 * world * (Bsz * self.num_windows) * int(model.encoder_projection_layer[-1].num_features) * 4)
        else:
            self._peer = None
        images = Bf.get("images", (S, 3, 32, 128), F32)
        images[:Bsz].copy_(image)
        images[Bsz:].copy_(aug_image)
        mask_u8 = Bf.get("mask", (M,), torch.uint8)
        mask_u8.view(2, Bsz, TOK).copy_(vis_mask_pos.permute(1, 0, 2))      # M:496-497 view-major

        Bf.zero_phase("fwd")
        # masked rows of view 0 as an index list (M:569); done first so that its error flag is back on the host early (see _post_mask_err).
        # only_mim_on_ori_img False (M:571-575): the masked rows of BOTH views -- the mask buffer is view-major like the token rows, so the
        # same kernel over 2B "samples" yields the indices of view 0 followed by those of view 1.
        views = 1 if only_mim_on_ori_img else 2
        n_per = self._masked_per_sample(vis_mask_pos, views)
        n_m = views * Bsz * n_per
        idx = Bf.get("dec.idx", (max(n_m, 1),), torch.int32)
        err = Bf.zeroed("dec.err", (1,), torch.int32, "fwd")
        call("dig_mask_to_index", mask_u8, idx, err, views * Bsz, n_per)
        self._post_mask_err(err)
        # bf16 shadows of the online weights + fused qkv bias
        cur = torch.cuda.current_stream()
        # bf16 shadows of the online GEMM weights: FusedAdamW rewrites them together with the fp32 masters, so the cast pass only runs
        # when something else touched the parameters (first step, load_state_dict, another optimizer: in-place torch ops bump _version)
        state = (sum(p._version for p in self._online_gemm_params), ops.raw_parameter_writes())
        if state != self._cast_state or self._always_cast:
            self._mt("dig_mt_cast_bf16", self.tab_cast_online)
            self._cast_state = state
        self._mt("dig_mt_copy_f32", self.tab_qkv_bias["encoder."])
        side = self._side if self._two_streams else cur
        if side is not cur:
            side.wait_stream(cur)       # staged inputs, and the optimizer step that produced the weights the EMA reads
        with torch.cuda.stream(side):
            if not self.momentum_warm:
                self._mt("dig_mt_cast_bf16", self.tab_cast_momentum)
                self.momentum_warm = True
            self._mt("dig_mt_ema", self.tab_ema, float(m))                       # M:526 (before the momentum forward)
            self._mt("dig_mt_copy_f32", self.tab_qkv_bias["momentum_encoder."])
            Wm = self._enc_weights("momentum_encoder.")
            enc_m, _ = self._encoder_fwd(Wm, images, mask_u8, "m.", save=False)
            a0 = Bf.get("m.pp.in", (half, d), BF16)
            call("dig_cast_f32_bf16", enc_m, a0, half * d)
            ppm_layers = self._mlp_layers("pix_projector_m.", 3, model.pix_projector_m)
            ppm_out, _, _ = self._mlp_fwd(a0, ppm_layers, "m.pp")
            pooled_m = Bf.get("m.pooled", (S * self.num_windows, d), BF16)
            call("dig_pool_fwd", ppm_out, enc_m[half:], Bsz, pooled_m, S, d, self.num_windows)
            k, _, _ = self._mlp_fwd(pooled_m, self._mlp_layers("momentum_projection_layer.", 3, model.momentum_projection_layer), "m.proj")
            R = k.shape[0]            # 2 * B * num_windows rows: [k1 ; k2]
            Q, C = R // 2, k.shape[1]
            pc = self._peer
            if pc is not None and 2 * world * Q * C * 4 <= pc.key_table_bytes:
                # F.normalize + concat_all_gather (M:446-447, M:580-591) in one kernel: every row goes straight into every rank's table
                ep = pc.next_epoch(peer.CH_KEYS)
                call("dig_peer_l2norm_allgather", pc.bases, pc.world, pc.rank, peer.CH_KEYS, ep, k, None, Q, C, pc.key_table_bytes)
                k_all2 = pc.key_table_ptr(ep)
            else:
                kn = Bf.get("kn", (R, C), F32)
                call("dig_l2norm_fwd", k, kn, None, R, C)
                k_all2 = dist_layout.gather_keys(kn, Bf.get("kall", (world, R, C), F32) if world > 1 else None,
                                                 Bf.get("kall2", (2, world * Q, C), F32) if world > 1 else None)       # M:580-591
            # operands of the logits GEMM (K-major [hi|lo|hi]) and of its gradient GEMM (MN-major planes), see dig_split_bf16x3
            Nk = world * Q
            k3 = Bf.get("nce.k3", (2, Nk, 3 * C), BF16)
            k3s = Bf.get("nce.k3s", (2, 3 * Nk, C), BF16) if need_grad else None
            call("dig_split_bf16x3", k_all2, k3, 1, k3s, Nk, 2 * Nk, C)

        # ---- online branch ----
        W = self._enc_weights("encoder.")
        enc, sv_enc = self._encoder_fwd(W, images, mask_u8, "o.", save=True)
        pp_in = Bf.get("o.pp.in", (half, d), BF16)
        call("dig_cast_f32_bf16", enc, pp_in, half * d)
        pp_layers = self._mlp_layers("pix_projector.", 3, model.pix_projector)
        pp_out, _, sv_pp = self._mlp_fwd(pp_in, pp_layers, "o.pp")
        pooled = Bf.get("o.pooled", (S * self.num_windows, d), BF16)
        call("dig_pool_fwd", pp_out, enc[half:], Bsz, pooled, S, d, self.num_windows)
        proj_layers = self._mlp_layers("encoder_projection_layer.", 3, model.encoder_projection_layer)
        _, proj_bf16, sv_proj = self._mlp_fwd(pooled, proj_layers, "o.proj", want_bf16_out=True)
        pred_layers = self._mlp_layers("predictor.", 2, model.predictor)
        q, _, sv_pred = self._mlp_fwd(proj_bf16, pred_layers, "o.pred")
        qn = Bf.get("qn", (R, C), F32)
        qinv = Bf.get("qinv", (R,), F32)
        call("dig_l2norm_fwd", q, qn, qinv, R, C)

        # ---- contrastive loss (M:444-461): q1.k2 + q2.k1 ----
        if side is not cur:
            cur.wait_stream(side)       # the keys (and the momentum weights the next EMA overwrites) are ready
        # logits on tcgen05: [qh|qh|ql] . [kh|kl|kh]^T over K = 3C reproduces the fp32 einsum of M:451 to ~1e-5 (q1.k2 and q2.k1)
        res = Bf.zeroed("nce.res", (2, 4), F32, "fwd")
        q3 = Bf.get("nce.q3", (R, 3 * C), BF16)
        call("dig_split_bf16x3", qn, q3, 2, None, 0, R, C)
        lg = Bf.get("nce.lg", (2, Q, Nk), F32)
        ops.gemm(q3[:Q], k3[1], lg[0], alpha=1.0 / self.T)
        ops.gemm(q3[Q:], k3[0], lg[1], alpha=1.0 / self.T)
        call("dig_infonce_rows", lg[0], Q, Nk, dist_layout.label_offset(Q, rank), self.T, res[0])
        call("dig_infonce_rows", lg[1], Q, Nk, dist_layout.label_offset(Q, rank), self.T, res[1])

        # ---- masked-pixel decoder on the masked rows of view 0 (M:561-570; row-wise, so gather first) ----
        g0 = Bf.get("dec.g0", (n_m, d), BF16)
        call("dig_gather_rows", enc, idx, g0, n_m, d)
        S_ = self.shadow
        t1 = Bf.get("dec.t1", (n_m, 192), BF16)
        ops.gemm(g0, S_["pix_decoder.0.weight"], t1)
        t2 = Bf.get("dec.t2", (n_m, 192), F32)
        ops.gemm(t1, S_["pix_decoder.1.weight"], t2)
        t3 = Bf.get("dec.t3", (n_m, 192), BF16)
        dmean, drstd = Bf.get("dec.mean", (n_m,), F32), Bf.get("dec.rstd", (n_m,), F32)
        ln = model.pix_decoder[2]
        self._ln(t2, ln.weight, ln.bias, t3, dmean, drstd, gelu=1, eps=ln.eps)
        vis = torch.empty(n_m, 48, dtype=F32, device=dev)
        ops.gemm(t3, S_["pix_decoder.4.weight"], vis, bias=self._named["pix_decoder.4.bias"])

        self.forward_serial += 1
        self.saved = dict(serial=self.forward_serial, enc=sv_enc, W=W, pp=(pp_layers, sv_pp), proj=(proj_layers, sv_proj), pred=(pred_layers, sv_pred),
                          qn=qn, qinv=qinv, k3s=k3s, lg=lg, Q=Q, Nk=Nk, C=C, Bsz=Bsz,
                          idx=idx, n_m=n_m, g0=g0, t1=t1, t2=t2, t3=t3, dmean=dmean, drstd=drstd, pooled=pooled)
        contra = res[:, 0].sum()
        accs = res[:, 1:3].clone()      # [[q1_acc1, q1_acc5], [q2_acc1, q2_acc5]]
        return contra, vis.view(views * Bsz, n_per, 48), accs      # both views: view 0's samples, then view 1's

    def _post_mask_err(self, err):
        """The number of masked patches per sample is validated on the host for the first batch of a shape only (a blocking read); every
        later batch is validated by dig_mask_to_index on the device.  Its flag comes back through a pinned slot and is checked when the
        NEXT forward starts (no host block inside the step): a ragged mask raises one step late, and never indexes out of bounds."""
        if self._mask_err is None:
            self._mask_err = [None, torch.zeros(1, dtype=torch.int32).pin_memory()]
        self._mask_err[1].copy_(err, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._mask_err[0] = ev

    def _check_mask_err(self):
        if self._mask_err is not None and self._mask_err[0] is not None:
            self._mask_err[0].synchronize()
            self._mask_err[0] = None
            if int(self._mask_err[1][0]) != 0:
                raise ops.DigError("the previous batch did not mask the same number of patches in every sample (masking_generator.py:20)")

    def _masked_per_sample(self, vis_mask_pos, views=1):
        key = tuple(vis_mask_pos.shape) + (views,)
        if key not in self._n_masked:
            cnt = vis_mask_pos[:, :views].sum(dim=2)
            n = int(cnt[0, 0].item())
            if not bool((cnt == n).all().item()):
                raise ops.DigError("every sample (and view) must mask the same number of patches (masking_generator.py:20)")
            self._n_masked[key] = n
        return self._n_masked[key]

    # ------------------------------------------------------------------ backward
    def backward(self, d_contra, d_vis):
        """d_contra: 0-d fp32 tensor or None; d_vis: fp32 [B, n, 48] or None.  Returns gradients in trainable_params() order."""
        sv, Bf, model = self.saved, self.bufs, self.model
        if sv is None:
            raise ops.DigError("backward called without a saved forward")
        d, Bsz = self.d, sv["Bsz"]
        S, M, half = 2 * Bsz, 2 * Bsz * TOK, Bsz * TOK
        # One persistent flat gradient buffer: the views handed to autograd keep their addresses from step to step, so the pointer
        # tables of the multi-tensor grad-norm / AdamW launches are built once (rebuilding them costs synchronous pageable H2D copies).
        # If gradients of an earlier backward are still attached to the parameters (accumulation), they must not be overwritten.
        p0 = self._named[self.train_names[0]]
        sync = getattr(model, "_dig_grad_sync", None)
        world, _ = _world()
        if self._grad_flat is None:
            # DigDataParallel on one NVSwitch node: the flat buffer lives in IPC-mapped memory, so that ONE kernel at the end of the backward
            # averages it over the ranks through NVLink peer loads / stores (csrc/peer.cu) instead of 14 NCCL all-reduces that hold SMs
            # next to the backward's statically scheduled persistent GEMMs (collective allocation: every rank's first backward).
            self._grad_peer = None
            if (sync is not None and world > 1 and self._peer is not None and os.environ.get("DIG_PEER_GRADS", "1") != "0"
                    and world in (2, 4, 8) and self.grad_total % 4 == 0 and sync.process_group is None):
                r = peer.shared_float_buffer(self._peer, self.grad_total)
                if r is not None:
                    self._grad_flat, self._grad_peer = r
            if self._grad_flat is None:
                self._grad_flat = torch.zeros(self.grad_total, dtype=F32, device=self.device)
            flat = self._grad_flat
        elif p0.grad is not None and p0.grad.data_ptr() == self._grad_flat.data_ptr():
            flat = torch.zeros(self.grad_total, dtype=F32, device=self.device)
        else:
            flat = self._grad_flat
            flat.zero_()
        grads = self._grad_views(flat)
        self._sync_works = []
        self._sync_flat = flat if (sync is not None and world > 1) else None
        self._sync_peer = self._grad_peer is not None and flat is self._grad_flat and self._peer is not None   # else: NCCL segments
        self._peer_pending = []
        nb = len(self.model.encoder.blocks)
        # DIG_PEER_GRAD_OVERLAP = k exchanges overlapped with the backward, issued when the upper 1/(k+1), 2/(k+1), ... of the encoder blocks
        # are final (0: one exchange at the end)
        k = max(0, min(int(os.environ.get("DIG_PEER_GRAD_OVERLAP", "3")), nb - 1))
        self._peer_flush_keys = {"block%d" % (nb - (j * nb) // (k + 1)) for j in range(1, k + 1)}
        self._sync_group = sync.process_group if sync is not None else None
        Bf.zero_phase("bwd")
        g = Bf.get("bw.g", (M, d), F32)
        S_ = self.shadow

        # ---- contrastive head ----
        if d_contra is not None:
            Q, Nk, C = sv["Q"], sv["Nk"], sv["C"]
            R = 2 * Q
            # d qn = dlogits . k_all on tcgen05 with the same three-term split (dlogits [hi|hi|lo] over K = 3 Nk, keys as MN-major planes)
            dl3 = Bf.get("bw.dl3", (2 * Q, 3 * Nk), BF16)
            call("dig_split_bf16x3", sv["lg"], dl3, 2, None, 0, 2 * Q, Nk)
            dqn = Bf.zeroed("bw.dqn", (R, C), F32, "bwd")
            ops.gemm(dl3[:Q], sv["k3s"][1], dqn[:Q], b_mn_major=True, split_k=-1)
            ops.gemm(dl3[Q:], sv["k3s"][0], dqn[Q:], b_mn_major=True, split_k=-1)
            dq = Bf.get("bw.dq", (R, C), F32)
            gs = d_contra.reshape(1).to(F32).contiguous()
            call("dig_l2norm_bwd", dqn, sv["qn"], sv["qinv"], gs, dq, R, C)
            dproj = Bf.get("bw.dproj", (R, C), F32)
            self._mlp_bwd(dq, sv["pred"][0], sv["pred"][1], "bw.pred", grads, dproj)
            dpooled = Bf.get("bw.dpooled", (R, d), F32)
            self._mlp_bwd(dproj, sv["proj"][0], sv["proj"][1], "bw.proj", grads, dpooled)
            dpp = Bf.get("bw.dpp", (half, d), F32)
            call("dig_pool_bwd", dpooled, Bsz, dpp, g[half:], S, d, self.num_windows)
            self._mlp_bwd(dpp, sv["pp"][0], sv["pp"][1], "bw.pp", grads, g[:half])
        else:
            g.zero_()

        # ---- masked-pixel decoder ----
        if d_vis is not None:
            n_m = sv["n_m"]
            dv = Bf.get("bw.dvis", (n_m, 48), BF16)
            dvf = d_vis.reshape(n_m, 48).to(F32).contiguous()
            call("dig_cast_f32_bf16", dvf, dv, n_m * 48)
            call("dig_colsum", dv, 0, 48, grads["pix_decoder.4.bias"], None, n_m, 48)
            ops.gemm(dv, sv["t3"], grads["pix_decoder.4.weight"], a_mn_major=True, b_mn_major=True, split_k=-1)
            dt3 = Bf.get("bw.dt3", (n_m, 192), BF16)
            ops.gemm(dv, S_["pix_decoder.4.weight"], dt3, b_mn_major=True)
            dt2 = Bf.get("bw.dt2", (n_m, 192), BF16)
            ln = model.pix_decoder[2]
            call("dig_layernorm_bwd", dt3, sv["t2"], sv["dmean"], sv["drstd"], ln.weight, ln.bias, None, None, dt2,
                 grads["pix_decoder.2.weight"], grads["pix_decoder.2.bias"], None, n_m, 192, 1)
            ops.gemm(dt2, sv["t1"], grads["pix_decoder.1.weight"], a_mn_major=True, b_mn_major=True, split_k=-1)
            dt1 = Bf.get("bw.dt1", (n_m, 192), BF16)
            ops.gemm(dt2, S_["pix_decoder.1.weight"], dt1, b_mn_major=True)
            ops.gemm(dt1, sv["g0"], grads["pix_decoder.0.weight"], a_mn_major=True, b_mn_major=True, split_k=-1)
            dg0 = Bf.get("bw.dg0", (n_m, d), F32)
            ops.gemm(dt1, S_["pix_decoder.0.weight"], dg0, b_mn_major=True)
            call("dig_scatter_add_rows", dg0, sv["idx"], g, n_m, d)

        # ---- encoder ----
        self._allreduce_segment("heads")          # every head gradient is final: average it while the encoder backward runs
        gb = Bf.get("bw.gb", (M, d), BF16)
        call("dig_cast_f32_bf16", g, gb, M * d)
        self._encoder_bwd(sv["W"], sv["enc"], g, gb, grads)
        self._allreduce_segment("embed")
        for w in self._sync_works:                 # the current stream waits for the NCCL work (no host block)
            w.wait()
        self._sync_works = []
        if self._sync_flat is not None and self._sync_peer:
            # every gradient of this rank is final (main stream; the side stream -- weight gradients, earlier exchanges -- was joined above)
            if self._comm_stream is not None:
                torch.cuda.current_stream().wait_stream(self._comm_stream)     # exchanges of the channel stay stream-ordered
            self._peer_flush(final=True)
        self.saved = None
        return [grads[n] for n in self.train_names]

    def _allreduce_segment(self, key, stream=None, after=None):
        """DigDataParallel: average one finished segment of the flat gradient buffer over the ranks (async NCCL all-reduce, issued on
        `stream` -- the side stream during the encoder backward -- after the event `after` recorded on the chain stream)."""
        if self._sync_flat is None:
            return
        if self._sync_peer:
            # peer-memory path: finished segments are collected and averaged in a few exchanges -- two from the side stream in the middle of
            # the backward (blocks small enough to sit next to the resident GEMM CTAs), the rest in one full-width kernel at the end
            self._peer_pending.append(key)
            if key in self._peer_flush_keys:
                # on a stream of its own (the exchange waits for the other ranks: it must not hold up the weight gradients queued behind it
                # on the side stream), after everything queued so far on the chain stream and on the side stream
                cur = torch.cuda.current_stream()
                if self._comm_stream is None:
                    self._comm_stream = torch.cuda.Stream(device=self.device)
                comm = self._comm_stream
                for st in (cur, stream):
                    if st is not None:
                        ev = torch.cuda.Event()
                        ev.record(st)
                        comm.wait_event(ev)
                with torch.cuda.stream(comm):
                    self._peer_flush(final=False)
            return
        a, b = self.grad_seg[key]
        cur = torch.cuda.current_stream()
        st = cur if stream is None else stream
        if st is not cur:
            if after is None:
                after = torch.cuda.Event()
                after.record(cur)
            st.wait_event(after)
        with torch.cuda.stream(st):
            self._sync_works.append(dist.all_reduce(self._sync_flat[a:b], op=dist.ReduceOp.AVG, group=self._sync_group, async_op=True))


def _peer_flush_impl(self, final):
    """Average the pending (finished) segments of the flat gradient buffer over the ranks with dig_peer_grad_allreduce: one call per
    contiguous range.  Every rank issues the same sequence (the segment order is a property of the model)."""
    if not self._peer_pending:
        return
    rng = sorted(self.grad_seg[k] for k in self._peer_pending)
    self._peer_pending = []
    merged = [list(rng[0])]
    for a, b in rng[1:]:
        if a == merged[-1][1]:
            merged[-1][1] = b
        else:
            merged.append([a, b])
    pc = self._peer
    for a, b in merged:
        if a % 4 or (b - a) % 4:
            raise ops.DigError("gradient segment [%d, %d) is not 16-byte aligned" % (a, b))
        call("dig_peer_grad_allreduce", pc.bases, self._grad_peer, pc.world, pc.rank, peer.CH_GRADS, pc.next_epoch(peer.CH_GRADS), a, b - a,
             int(os.environ.get("DIG_PEER_GRAD_BLOCKS", "0")) if final else 0, 0 if final else 1)


PretrainStep._peer_flush = _peer_flush_impl


class _PretrainFn(torch.autograd.Function):
    """One autograd node for the whole model: gradients reach ordinary leaf nn.Parameters (DDP-compatible)."""

    @staticmethod
    def forward(ctx, step, image, aug_image, vis_mask_pos, m, only_mim, *params):
        contra, vis, accs = step.forward(image, aug_image, vis_mask_pos, m, only_mim)
        ctx.step = step
        ctx.serial = step.forward_serial
        ctx.mark_non_differentiable(accs)
        return contra, vis, accs

    @staticmethod
    def backward(ctx, d_contra, d_vis, _d_accs):
        sv = ctx.step.saved
        if sv is None or sv.get("serial") != ctx.serial:
            # activations live in per-model buffers that every forward overwrites: only the most recent forward can be back-propagated
            raise ops.DigError("backward of a stale forward: dig_b200 keeps the activations of the most recent forward only "
                               "(run backward before the next forward of the same model)")
        grads = ctx.step.backward(d_contra, d_vis)
        return (None, None, None, None, None, None) + tuple(grads)


def run_model(step, image, aug_image, vis_mask_pos, m, only_mim_on_ori_img):
    params = step.trainable_params()
    if torch.is_grad_enabled() and any(p.requires_grad for p in params):
        contra, vis, accs = _PretrainFn.apply(step, image, aug_image, vis_mask_pos, m, only_mim_on_ori_img, *params)
    else:
        contra, vis, accs = step.forward(image, aug_image, vis_mask_pos, m, only_mim_on_ori_img)
    B = image.shape[0]
    vis_out = [vis] if only_mim_on_ori_img else [vis[:B], vis[B:]]          # M:566-575
    return {"contra_loss": contra, "q1_acc1": accs[0, 0:1], "q1_acc5": accs[0, 1:2], "q2_acc1": accs[1, 0:1],
            "q2_acc5": accs[1, 1:2], "vis_out": vis_out}

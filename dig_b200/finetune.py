"""Fine-tuning step on the sm_100a kernels (BASELINE configs[4]).

`DigRecModel` is a state-dict compatible stand-in for the reference's `RecModel` with `decoder_name='tf_decoder'`
(models/model_builder.py:74-169): the `simmim_vit_*_patch4_32x128` encoder WITH its final LayerNorm (modeling_pretrain_vit.py:27-111),
`linear_norm` (model_builder.py:86-89) and `TFDecoder` (models/decoder.py:107-222; six pre-LN TransformerDecoderLayers,
models/transformer_layer.py:47-118).  `FinetuneStep` sequences its training forward / backward on the C-ABI: the encoder blocks reuse
the pre-training path (dig_gemm, fused attention, LayerNorm kernels); every decoder Linear is a dig_gemm (q/k/v and the cross-attention
k/v projections fused into one GEMM each); the T <= 32-query attention, the <BOS>-shifted embedding and SeqCrossEntropyLoss run in
csrc/decoder.cu.  Built: training forward with every dropout / drop-path probability 0 (the README's 0.1 rates need RNG-matched dropout
inside the fused kernels), teacher forcing (`forward_train`) in train mode and greedy decoding (`forward_test`) in eval mode; beam search raises.
"""
import os

import torch
import torch.nn as nn

from . import ops, registry
from .modeling import _Encoder, _Holder
from .ops import call
from .pretrain_step import BF16, F32, TOK, MtTable, PretrainStep, _Bufs

DEC_LN_EPS = 1e-5


class _MHA(_Holder):
    def __init__(self, d_model=512, n_head=8, d_k=64):
        super().__init__()
        self.n_head, self.d_k = n_head, d_k
        self.linear_q = nn.Linear(n_head * d_k, n_head * d_k, bias=False)
        self.linear_k = nn.Linear(n_head * d_k, n_head * d_k, bias=False)
        self.linear_v = nn.Linear(n_head * d_k, n_head * d_k, bias=False)
        self.fc = nn.Linear(n_head * d_k, d_model, bias=False)


class _FFN(_Holder):
    def __init__(self, d_in, d_hid):
        super().__init__()
        self.w_1 = nn.Linear(d_in, d_hid)
        self.w_2 = nn.Linear(d_hid, d_in)


class _DecoderLayer(_Holder):
    def __init__(self, d_model, d_inner, n_head, d_k):
        super().__init__()
        self.self_attn = _MHA()                  # transformer_layer.py:61 first builds a default MultiHeadAttention under this name (replaced
        self.norm1 = nn.LayerNorm(d_model)       # below): same RNG draws, and self_attn keeps its place ahead of the norms in the state dict
        self.norm2 = nn.LayerNorm(d_model)
        self.norm3 = nn.LayerNorm(d_model)
        self.self_attn = _MHA(d_model, n_head, d_k)
        self.enc_attn = _MHA(d_model, n_head, d_k)
        self.mlp = _FFN(d_model, d_inner)


class _PositionalEncoding(_Holder):
    def __init__(self, d_hid=512, n_position=200):
        super().__init__()
        import numpy as np
        den = torch.Tensor([1.0 / np.power(10000, 2 * (j // 2) / d_hid) for j in range(d_hid)]).view(1, -1)      # transformer_layer.py:417-428
        tab = torch.arange(n_position).unsqueeze(-1).float() * den
        tab[:, 0::2] = torch.sin(tab[:, 0::2])
        tab[:, 1::2] = torch.cos(tab[:, 1::2])
        self.register_buffer("position_table", tab.unsqueeze(0))


class _TFDecoder(_Holder):
    def __init__(self, n_layers=6, d_embedding=512, n_head=8, d_k=64, d_model=512, d_inner=256, n_position=200, num_classes=97, max_seq_len=25):
        super().__init__()
        self.max_seq_len, self.start_idx, self.d_embedding, self.num_classes = max_seq_len, num_classes, d_embedding, num_classes
        self.n_head = n_head
        self.trg_word_emb = nn.Embedding(num_classes + 1, d_embedding)
        self.position_enc = _PositionalEncoding(d_embedding, n_position)
        self.layer_stack = nn.ModuleList([_DecoderLayer(d_model, d_inner, n_head, d_k) for _ in range(n_layers)])
        self.layer_norm = nn.LayerNorm(d_model, eps=1e-6)
        self.classifier = nn.Linear(d_model, num_classes)


class DigRecModel(nn.Module):
    def __init__(self, encoder_name="simmim_vit_small_patch4_32x128", nb_classes=97, max_len=25, decoder_name="tf_decoder", drop=0.0,
                 drop_path=0.0, attn_drop_rate=0.0, decoder_dropout=0.0, **unused):
        super().__init__()
        if decoder_name != "tf_decoder":
            raise NotImplementedError("only decoder_name='tf_decoder' (README.md:91-112) is built, got %r" % (decoder_name,))
        if any(float(v or 0.0) != 0.0 for v in (drop, drop_path, attn_drop_rate, decoder_dropout)):
            raise NotImplementedError("dropout / drop-path are not built on the fused path: run with --drop 0 --attn_drop_rate 0 --drop_path 0")
        dims = {"simmim_vit_tiny_patch4_32x128": (192, 3), "simmim_vit_small_patch4_32x128": (384, 6), "simmim_vit_base_patch4_32x128": (512, 8)}
        if encoder_name not in dims:
            raise NotImplementedError("unknown encoder %r" % (encoder_name,))
        d, h = dims[encoder_name]
        self.encoder = _Encoder((32, 128), 4, 3, d, 12, h, 4, 1e-6, final_norm=True)            # V:114-136
        self.decoder = _TFDecoder(num_classes=nb_classes, max_seq_len=max_len)
        self.linear_norm = nn.Sequential(nn.Linear(d, self.decoder.d_embedding), nn.LayerNorm(self.decoder.d_embedding))
        self.patch_embed = self.encoder.patch_embed          # model_builder.py:92-93 aliases (they appear in the state dict)
        self.pos_embed = self.encoder.pos_embed
        self.trg_word_emb = None
        self.insert_sem = False
        self._step = None

    def no_weight_decay(self):
        return {"encoder." + k for k in self.encoder.no_weight_decay()}

    def get_num_layers(self):
        return self.encoder.get_num_layers()

    def _pipeline(self):
        if self._step is None or not self._step.matches(self):
            self._step = FinetuneStep(self)
        return self._step

    def forward(self, x):
        """x = (images fp32 [B,3,32,128], targets int64 [B,T], tgt_lens int64 [B]) -> (logits [B,T,nb_classes], None, None, None),
        the tuple RecModel.forward returns in train mode (model_builder.py:124-169; the attention maps it also returns are only
        consumed by visualisation code and are available through FinetuneStep.forward(need_maps=True))."""
        images, targets, tgt_lens = x
        if not images.is_cuda:
            raise RuntimeError("dig_b200 runs on sm_100a only: inputs must be CUDA tensors (no CPU fallback)")
        step = self._pipeline()
        if not self.training:
            # eval mode (model_builder.py:137-139 drops the targets): greedy decoding, TFDecoder.forward_test (decoder.py:224-250)
            if getattr(self, "beam_width", 0):
                raise NotImplementedError("beam search (models/decoder.py:252-330) is not built; evaluate with --beam_width 0")
            probs, maps, _ = step.greedy_decode(images, need_maps=bool(getattr(self, "return_attn_maps", False)))
            return probs, None, None, maps
        params = step.trainable_params()
        if torch.is_grad_enabled():
            logits = _FinetuneFn.apply(step, images, targets, tgt_lens, *params)
        else:
            logits = step.forward(images, targets, tgt_lens)
        return logits, None, None, None


class FinetuneStep(PretrainStep):
    def __init__(self, model):      # noqa: deliberately does not call PretrainStep.__init__ (different parameter set)
        ops.load()
        p0 = next(model.parameters())
        if not p0.is_cuda:
            raise ops.DigError("dig_b200 runs on sm_100a only: move the model to a CUDA device first")
        self.model, self.device = model, p0.device
        enc, dec = model.encoder, model.decoder
        self.d, self.heads, self.depth = enc.embed_dim, enc.num_heads, len(enc.blocks)
        self.dm, self.dh, self.nl = dec.d_embedding, dec.n_head, len(dec.layer_stack)
        self.C = dec.num_classes
        self.Cp = (self.C + 7) // 8 * 8                       # classifier rows padded for the GEMM (N % 4, MN-major lda % 8)
        self.scale = 64 ** -0.5
        self.bufs = _Bufs(self.device)
        self.pos = enc.pos_embed.reshape(TOK, self.d).to(self.device, F32).contiguous()
        self._named = dict(model.named_parameters())
        self._ptr_sig = self._signature()
        self.saved, self.forward_serial = None, 0
        self._grad_flat = None
        self._two_streams = os.environ.get("DIG_TWO_STREAMS", "1") != "0"
        self._side = torch.cuda.Stream(device=self.device) if self._two_streams else None
        self._always_cast = os.environ.get("DIG_ALWAYS_CAST", "0") == "1"
        self._bf16_grad_stream = os.environ.get("DIG_BF16_GRAD_STREAM", "1") != "0"
        self._gelu_q8 = os.environ.get("DIG_GELU_Q8", "1") != "0"     # encoder + decoder FFN pre-activations as 8-bit codes (dig_gemm_t.aux_q8)
        self._sync_flat, self._sync_works, self._sync_group = None, [], None
        self._build()

    # ------------------------------------------------------------------ shadows / tables
    def _build(self):
        N, dev, dm = self._named, self.device, self.dm
        src, dst, self.shadow = [], [], {}

        def plain(name, rows=None):
            p = N[name]
            shape = (p.shape[0], p.numel() // p.shape[0])
            buf = torch.zeros((rows or shape[0], shape[1]), dtype=BF16, device=dev)
            self.shadow[name] = buf
            src.append(p.data); dst.append(buf[:shape[0]])

        def fused(key, names):
            rows = sum(N[n].shape[0] for n in names)
            buf = torch.zeros(rows, N[names[0]].shape[1], dtype=BF16, device=dev)
            self.shadow[key] = buf
            r = 0
            for n in names:
                src.append(N[n].data); dst.append(buf[r:r + N[n].shape[0]])
                r += N[n].shape[0]

        plain("encoder.patch_embed.proj.weight")
        for l in range(self.depth):
            for w in self.GEMM_WEIGHTS_BLOCK:
                plain("encoder.blocks.%d.%s" % (l, w))
        plain("linear_norm.0.weight")
        for l in range(self.nl):
            p = "decoder.layer_stack.%d." % l
            fused(p + "self_qkv", [p + "self_attn.linear_q.weight", p + "self_attn.linear_k.weight", p + "self_attn.linear_v.weight"])
            plain(p + "self_attn.fc.weight")
            plain(p + "enc_attn.linear_q.weight")
            fused(p + "enc_kv", [p + "enc_attn.linear_k.weight", p + "enc_attn.linear_v.weight"])
            plain(p + "enc_attn.fc.weight")
            plain(p + "mlp.w_1.weight")
            plain(p + "mlp.w_2.weight")
        plain("decoder.classifier.weight", rows=self.Cp)
        self.tab_cast_online = MtTable(dev, src, dst)
        for s_, d_ in zip(src, dst):
            ops._shadows[s_.data_ptr()] = d_
        self._shadow_keep = dst
        self._online_gemm_params = [p for p in N.values() if p.data_ptr() in {s_.data_ptr() for s_ in src}]
        self._cast_state = None
        self.cls_bias_pad = torch.zeros(self.Cp, device=dev)
        d, L = self.d, self.depth
        self.qkv_bias = {"encoder.": torch.zeros(L, 3 * d, device=dev)}
        bs, bd = [], []
        for l in range(L):
            bs += [N["encoder.blocks.%d.attn.q_bias" % l].data, N["encoder.blocks.%d.attn.v_bias" % l].data]
            bd += [self.qkv_bias["encoder."][l, :d], self.qkv_bias["encoder."][l, 2 * d:]]
        bs.append(N["decoder.classifier.bias"].data); bd.append(self.cls_bias_pad[:self.C])
        self.tab_qkv_bias = {"encoder.": MtTable(dev, bs, bd)}
        # trainable parameters and the flat gradient buffer (named_parameters order; encoder.mask_token is unused in fine-tuning and
        # returns no gradient -- run_class_finetuning.py:497 wraps with find_unused_parameters=True for that reason)
        self.train_names = [n for n, p in N.items() if p.requires_grad]
        goff, total = {}, 0
        for n in self.train_names:
            goff[n] = total
            total += (N[n].numel() + 3) // 4 * 4
        self.grad_off, self.grad_total = goff, total
        self.grad_seg = {}
        self.mask_zero = {}

    # ------------------------------------------------------------------ forward (model_builder.py:124-169, decoder.py:180-222)
    def greedy_decode(self, images, need_maps=False, force_tokens=None):
        """RecModel.forward in eval mode with beam_width 0 = TFDecoder.forward_test (models/decoder.py:224-250): max_seq_len decoder passes
        over the <BOS>-shifted sequence decoded so far (tgt_lens = step + 1), softmax of the classifier output at position `step`, its
        arg-max fed back.  The encoder and linear_norm run once; every step reuses their output (the decoder's memory).
        -> (probabilities fp32 [B, T, C], head-averaged cross-attention maps of the last layer [B, T, 256] or None, tokens int64 [B, T]).
        force_tokens [B, T] (tests only) replaces the fed-back arg-max."""
        B, T, dev = images.shape[0], int(self.model.decoder.max_seq_len), self.device
        toks = torch.zeros(B, T, dtype=torch.int64, device=dev)           # position t feeds the query at t + 1 (decoder.py:247)
        out = torch.zeros(B, T, dtype=torch.int64, device=dev)
        probs = torch.empty(B, T, self.C, dtype=F32, device=dev)
        maps = torch.empty(B, T, TOK, dtype=F32, device=dev) if need_maps else None
        with torch.no_grad():
            for t in range(T):
                lens = torch.full((B,), t + 1, dtype=torch.int64, device=dev)
                logits = self.forward(images, toks, lens, need_maps=need_maps, reuse_memory=t > 0)
                p = torch.softmax(logits[:, t].float(), dim=-1)
                probs[:, t] = p
                out[:, t] = p.argmax(-1)
                toks[:, t] = out[:, t] if force_tokens is None else force_tokens[:, t].to(dev)
                if need_maps:
                    maps[:, t] = self.last_maps[:, t]
        self.saved = None          # nothing of these passes may be back-propagated
        return probs, maps, out

    def forward(self, images, targets, tgt_lens, need_maps=False, reuse_memory=False):
        """reuse_memory (greedy decoding): skip the encoder / linear_norm and decode against the memory of the previous call."""
        model, Bf, dev = self.model, self.bufs, self.device
        N_, S_ = self._named, self.shadow
        B, T = targets.shape
        if tuple(images.shape[1:]) != (3, 32, 128) or images.shape[0] != B or T > 32:
            raise ops.DigError("fine-tune step expects images [B,3,32,128] and targets [B,T<=32]; got %s / %s" % (tuple(images.shape), tuple(targets.shape)))
        M, Md, d, dm, H = B * TOK, B * T, self.d, self.dm, self.dh
        Bf.zero_phase("fwd")
        state = (sum(p._version for p in self._online_gemm_params), ops.raw_parameter_writes())
        if state != self._cast_state or self._always_cast:
            self._mt("dig_mt_cast_bf16", self.tab_cast_online)
            self._cast_state = state
        self._mt("dig_mt_copy_f32", self.tab_qkv_bias["encoder."])
        targets = targets.to(dev, torch.int64).contiguous()
        tgt_lens = tgt_lens.to(dev, torch.int64).contiguous()
        W = self._enc_weights("encoder.")
        mem = Bf.get("f.mem", (M, dm), BF16)
        if reuse_memory:
            sv_enc = x = encn = me = re = lpre = ml = rl = None
        else:
            images = images.to(F32).contiguous()
            mask_u8 = self.mask_zero.get(M)
            if mask_u8 is None:
                mask_u8 = self.mask_zero[M] = torch.zeros(M, dtype=torch.uint8, device=dev)
            x, sv_enc = self._encoder_fwd(W, images, mask_u8, "f.", save=True)
            # final encoder norm (V:104) and linear_norm (model_builder.py:86-89, :146)
            encn = Bf.get("f.encn", (M, d), BF16)
            me, re = Bf.get("f.me", (M,), F32), Bf.get("f.re", (M,), F32)
            self._ln(x, N_["encoder.norm.weight"], N_["encoder.norm.bias"], encn, me, re, eps=model.encoder.norm.eps)
            lpre = Bf.get("f.lpre", (M, dm), F32)
            ops.gemm(encn, S_["linear_norm.0.weight"], lpre, bias=N_["linear_norm.0.bias"])
            ml, rl = Bf.get("f.ml", (M,), F32), Bf.get("f.rl", (M,), F32)
            self._ln(lpre, N_["linear_norm.1.weight"], N_["linear_norm.1.bias"], mem, ml, rl, eps=model.linear_norm[1].eps)
        # decoder input: <BOS>-shifted target embedding + position table (decoder.py:173-178, :212-214)
        xd = Bf.get("d.x0", (Md, dm), F32)
        call("dig_embed_pos_fwd", targets, N_["decoder.trg_word_emb.weight"], model.decoder.position_enc.position_table, xd, B, T, dm,
             model.decoder.start_idx)
        maps = Bf.zeroed("d.maps", (B, T, TOK), F32, "fwd") if need_maps else None
        layers = []
        for l in range(self.nl):
            p, t = "decoder.layer_stack.%d." % l, "d%d." % l
            eps = model.decoder.layer_stack[l].norm1.eps
            h1 = Bf.get(t + "h1", (Md, dm), BF16)
            m1, r1 = Bf.get(t + "m1", (Md,), F32), Bf.get(t + "r1", (Md,), F32)
            self._ln(xd, N_[p + "norm1.weight"], N_[p + "norm1.bias"], h1, m1, r1, eps=eps)
            qkv = Bf.get(t + "qkv", (Md, 3 * dm), BF16)
            ops.gemm(h1, S_[p + "self_qkv"], qkv)
            sa = Bf.get(t + "sa", (Md, dm), BF16)
            lse_s = Bf.get(t + "lse_s", (B, H, T), F32)
            call("dig_dec_attention_fwd", qkv, 3 * dm, qkv[:, dm:], 3 * dm, qkv[:, 2 * dm:], 3 * dm, sa, dm, lse_s, tgt_lens, None, B, H, T, T,
                 self.scale)
            x1 = Bf.get(t + "x1", (Md, dm), F32)
            ops.gemm(sa, S_[p + "self_attn.fc.weight"], x1, residual=xd)
            h2 = Bf.get(t + "h2", (Md, dm), BF16)
            m2, r2 = Bf.get(t + "m2", (Md,), F32), Bf.get(t + "r2", (Md,), F32)
            self._ln(x1, N_[p + "norm2.weight"], N_[p + "norm2.bias"], h2, m2, r2, eps=eps)
            qc = Bf.get(t + "qc", (Md, dm), BF16)
            ops.gemm(h2, S_[p + "enc_attn.linear_q.weight"], qc)
            # Encoder-decoder attention on the fused tcgen05 attention kernel of the encoder: the T <= 32 decoder queries of a sample are
            # placed in the first T rows of a 256-row query tile (the other rows stay zero and their outputs are never read), K and V of the
            # 256 memory tokens are written by the fused k|v projection GEMM straight into the q|k|v layout the kernel reads.
            qkvx = self._zero_once(t + "qkvx", (M, 3 * dm), BF16)
            kv = qkvx[:, dm:]
            ops.gemm(mem, S_[p + "enc_kv"], kv)
            qkvx.view(B, TOK, 3 * dm)[:, :T, :dm].copy_(qc.view(B, T, dm))
            attx = Bf.get(t + "attx", (M, dm), BF16)
            lse_c = Bf.get(t + "lse_c", (B, H, TOK), F32)
            ops.attention_fwd(qkvx, attx, lse_c, H, self.scale)
            ca = Bf.get(t + "ca", (Md, dm), BF16)
            ca.view(B, T, dm).copy_(attx.view(B, TOK, dm)[:, :T])
            if need_maps and l == self.nl - 1:      # head-averaged attention weights of the last layer (visualisation only): CUDA-core kernel
                scratch = Bf.get("d.maps_o", (Md, dm), BF16)
                call("dig_dec_attention_fwd", qc, dm, kv, 3 * dm, qkvx[:, 2 * dm:], 3 * dm, scratch, dm, Bf.get("d.maps_lse", (B, H, T), F32), None,
                     maps, B, H, T, TOK, self.scale)
            x2 = Bf.get(t + "x2", (Md, dm), F32)
            ops.gemm(ca, S_[p + "enc_attn.fc.weight"], x2, residual=x1)
            h3 = Bf.get(t + "h3", (Md, dm), BF16)
            m3, r3 = Bf.get(t + "m3", (Md,), F32), Bf.get(t + "r3", (Md,), F32)
            self._ln(x2, N_[p + "norm3.weight"], N_[p + "norm3.bias"], h3, m3, r3, eps=eps)
            di = S_[p + "mlp.w_1.weight"].shape[0]
            fpre, f1 = Bf.get(t + "fpre", (Md, di), torch.uint8 if self._gelu_q8 else BF16), Bf.get(t + "f1", (Md, di), BF16)
            ops.gemm(h3, S_[p + "mlp.w_1.weight"], f1, bias=N_[p + "mlp.w_1.bias"], epilogue=ops.EPI_GELU, aux=fpre)
            x3 = Bf.get(t + "x3", (Md, dm), F32)
            ops.gemm(f1, S_[p + "mlp.w_2.weight"], x3, bias=N_[p + "mlp.w_2.bias"], residual=x2)
            layers.append(dict(x0=xd, h1=h1, m1=m1, r1=r1, qkv=qkv, sa=sa, lse_s=lse_s, x1=x1, h2=h2, m2=m2, r2=r2, qc=qc, qkvx=qkvx, attx=attx, ca=ca,
                               lse_c=lse_c, x2=x2, h3=h3, m3=m3, r3=r3, fpre=fpre, f1=f1, p=p))
            xd = x3
        hf = Bf.get("d.hf", (Md, dm), BF16)
        mf, rf = Bf.get("d.mf", (Md,), F32), Bf.get("d.rf", (Md,), F32)
        self._ln(xd, N_["decoder.layer_norm.weight"], N_["decoder.layer_norm.bias"], hf, mf, rf, eps=model.decoder.layer_norm.eps)
        logits = torch.empty(Md, self.Cp, dtype=F32, device=dev)
        ops.gemm(hf, S_["decoder.classifier.weight"], logits, bias=self.cls_bias_pad)
        self.forward_serial += 1
        self.saved = dict(serial=self.forward_serial, enc=sv_enc, W=W, x_enc=x, encn=encn, me=me, re=re, lpre=lpre, mem=mem, ml=ml, rl=rl,
                          layers=layers, x_last=xd, hf=hf, mf=mf, rf=rf, B=B, T=T, targets=targets, lens=tgt_lens)
        self.last_maps = maps
        return logits.view(B, T, self.Cp)[:, :, :self.C]

    # ------------------------------------------------------------------ backward
    def backward(self, d_logits):
        sv, Bf, model = self.saved, self.bufs, self.model
        if sv is None:
            raise ops.DigError("backward called without a saved forward")
        N_, S_ = self._named, self.shadow
        B, T = sv["B"], sv["T"]
        M, Md, d, dm, H = B * TOK, B * T, self.d, self.dm, self.dh
        if self._grad_flat is None:
            self._grad_flat = torch.zeros(self.grad_total, dtype=F32, device=self.device)
        else:
            self._grad_flat.zero_()
        grads = self._grad_views(self._grad_flat)
        Bf.zero_phase("bwd")

        def wgrad(dy, act, out):
            ops.gemm(dy, act, out, a_mn_major=True, b_mn_major=True, split_k=-1)

        def ln_bwd(dy, x, mean, rstd, name, dres, dx_f32, dx_bf16=None, dxsum=None):
            call("dig_layernorm_bwd", dy, x, mean, rstd, N_[name + "weight"], None, dres, dx_f32, dx_bf16, grads[name + "weight"],
                 grads[name + "bias"], dxsum, x.shape[0], x.shape[1], 0)

        # classifier (decoder.py:171, :221)
        dl = Bf.get("b.dl", (Md, self.Cp), F32)
        dl.zero_()
        dl.view(B, T, self.Cp)[:, :, :self.C].copy_(d_logits)
        dlb = Bf.get("b.dlb", (Md, self.Cp), BF16)
        call("dig_cast_f32_bf16", dl, dlb, Md * self.Cp)
        dcls = Bf.zeroed("b.dcls", (self.Cp, dm), F32, "bwd")
        wgrad(dlb, sv["hf"], dcls)
        dbc = Bf.zeroed("b.dbc", (self.Cp,), F32, "bwd")
        call("dig_colsum", dl, 1, self.Cp, dbc, None, Md, self.Cp)
        dhf = Bf.get("b.dh", (Md, dm), BF16)
        ops.gemm(dlb, S_["decoder.classifier.weight"], dhf, b_mn_major=True)
        gx = Bf.get("b.gx", (Md, dm), F32)
        ln_bwd(dhf, sv["x_last"], sv["mf"], sv["rf"], "decoder.layer_norm.", None, gx)
        gxb = Bf.get("b.gxb", (Md, dm), BF16)
        dmem = Bf.get("b.dmem", (M, dm), F32)
        first_mem = True
        for l in reversed(range(self.nl)):
            a = sv["layers"][l]
            p = a["p"]
            di = a["f1"].shape[1]
            # ---- feed-forward (transformer_layer.py:386-404) ----
            call("dig_cast_f32_bf16", gx, gxb, Md * dm)
            call("dig_colsum", gxb, 0, dm, grads[p + "mlp.w_2.bias"], None, Md, dm)
            df1 = Bf.get("b.df1", (Md, di), BF16)
            ops.gemm(gxb, S_[p + "mlp.w_2.weight"], df1, b_mn_major=True, epilogue=ops.EPI_GELU_BWD, aux=a["fpre"], colsum=grads[p + "mlp.w_1.bias"])
            wgrad(gxb, a["f1"], grads[p + "mlp.w_2.weight"])
            wgrad(df1, a["h3"], grads[p + "mlp.w_1.weight"])
            dh = Bf.get("b.dh", (Md, dm), BF16)
            ops.gemm(df1, S_[p + "mlp.w_1.weight"], dh, b_mn_major=True)
            ln_bwd(dh, a["x2"], a["m3"], a["r3"], p + "norm3.", gx, gx)
            # ---- encoder-decoder attention (transformer_layer.py:106-112) ----
            call("dig_cast_f32_bf16", gx, gxb, Md * dm)
            dca = Bf.get("b.dca", (Md, dm), BF16)
            # the dgrad of the output projection also emits D = rowsum(dO o O) per head (DIG_EPI_ROWDOT), which lets the cross-attention
            # backward run on the persistent kernel (as in the encoder) instead of the one-shot one that recomputes D from O
            dsum_s = Bf.get("b.dsum_s", (Md, H), F32)
            ops.gemm(gxb, S_[p + "enc_attn.fc.weight"], dca, b_mn_major=True, epilogue=ops.EPI_ROWDOT, aux=a["ca"], rowdot=dsum_s)
            wgrad(gxb, a["ca"], grads[p + "enc_attn.fc.weight"])
            dqc = Bf.get("b.dqc", (Md, dm), BF16)
            dattx = self._zero_once("b.dattx", (M, dm), BF16)          # rows >= T of every 256-row tile stay zero
            dattx.view(B, TOK, dm)[:, :T].copy_(dca.view(B, T, dm))
            dqkvx = Bf.get("b.dqkvx", (M, 3 * dm), BF16)
            dsumx = self._zero_once("b.dsumx", (M, H), F32)            # D of the padding rows is zero (their dO is)
            dsumx.view(B, TOK, H)[:, :T].copy_(dsum_s.view(B, T, H))
            ops.attention_bwd_d(a["qkvx"], dattx, a["lse_c"], dsumx, dqkvx, H, self.scale)
            dqc.view(B, T, dm).copy_(dqkvx.view(B, TOK, 3 * dm)[:, :T, :dm])
            dkv = dqkvx[:, dm:]
            wgrad(dqc, a["h2"], grads[p + "enc_attn.linear_q.weight"])
            wgrad(dkv, sv["mem"], self._pair_view(p + "enc_attn.linear_k.weight", p + "enc_attn.linear_v.weight"))
            ops.gemm(dkv, S_[p + "enc_kv"], dmem, b_mn_major=True, residual=None if first_mem else dmem)
            first_mem = False
            ops.gemm(dqc, S_[p + "enc_attn.linear_q.weight"], dh, b_mn_major=True)
            ln_bwd(dh, a["x1"], a["m2"], a["r2"], p + "norm2.", gx, gx)
            # ---- masked self-attention (transformer_layer.py:98-104) ----
            call("dig_cast_f32_bf16", gx, gxb, Md * dm)
            dsa = Bf.get("b.dca", (Md, dm), BF16)
            ops.gemm(gxb, S_[p + "self_attn.fc.weight"], dsa, b_mn_major=True)
            wgrad(gxb, a["sa"], grads[p + "self_attn.fc.weight"])
            dqkv = Bf.get("b.dqkv", (Md, 3 * dm), BF16)
            qkv = a["qkv"]
            call("dig_dec_attention_bwd", qkv, 3 * dm, qkv[:, dm:], 3 * dm, qkv[:, 2 * dm:], 3 * dm, a["sa"], dm, dsa, dm, a["lse_s"], sv["lens"],
                 dqkv, 3 * dm, dqkv[:, dm:], 3 * dm, dqkv[:, 2 * dm:], 3 * dm, B, H, T, T, self.scale)
            wgrad(dqkv, a["h1"], self._pair_view(p + "self_attn.linear_q.weight", p + "self_attn.linear_v.weight"))
            ops.gemm(dqkv, S_[p + "self_qkv"], dh, b_mn_major=True)
            ln_bwd(dh, a["x0"], a["m1"], a["r1"], p + "norm1.", gx, gx)
        call("dig_embed_bwd", gx, sv["targets"], grads["decoder.trg_word_emb.weight"], B, T, dm, model.decoder.start_idx)
        grads["decoder.classifier.weight"].copy_(dcls[:self.C])
        grads["decoder.classifier.bias"].copy_(dbc[:self.C])
        # ---- linear_norm and the encoder's final norm ----
        dmemb = Bf.get("b.dmemb", (M, dm), BF16)
        call("dig_cast_f32_bf16", dmem, dmemb, M * dm)
        dlpre = Bf.get("b.dlpre", (M, dm), BF16)
        ln_bwd(dmemb, sv["lpre"], sv["ml"], sv["rl"], "linear_norm.1.", None, None, dlpre, grads["linear_norm.0.bias"])
        wgrad(dlpre, sv["encn"], grads["linear_norm.0.weight"])
        dencn = Bf.get("b.dencn", (M, d), BF16)
        ops.gemm(dlpre, S_["linear_norm.0.weight"], dencn, b_mn_major=True)
        g = Bf.get("bw.g", (M, d), F32)
        gb = Bf.get("bw.gb", (M, d), BF16)
        ln_bwd(dencn, sv["x_enc"], sv["me"], sv["re"], "encoder.norm.", None, g, gb)
        self._sync_flat, self._sync_works = None, []
        self._encoder_bwd(sv["W"], sv["enc"], g, gb, grads)
        self.saved = None
        return [None if n == "encoder.mask_token" else grads[n] for n in self.train_names]

    def _zero_once(self, name, shape, dtype):
        """A buffer whose never-written part must read as zero: cleared when (re)allocated, not every step."""
        t = self.bufs.d.get(name)
        fresh = t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype
        t = self.bufs.get(name, shape, dtype)
        if fresh:
            t.zero_()
        return t

    def _pair_view(self, first, last):
        """One [rows, cols] fp32 view over the gradients of adjacent weights of equal width (they are consecutive in named_parameters
        order, hence in the flat gradient buffer): the fused q|k|v and k|v projection weight gradients are written by ONE wgrad GEMM."""
        a = self.grad_off[first]
        b = self.grad_off[last] + self._named[last].numel()
        cols = self._named[first].shape[1]
        if (b - a) % cols or any(self._named[n].numel() % 4 for n in (first, last)):
            raise ops.DigError("fused weight gradients need contiguous, 4-aligned parameter blocks")
        return self._grad_flat[a:b].view((b - a) // cols, cols)


class _FinetuneFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, step, images, targets, tgt_lens, *params):
        logits = step.forward(images, targets, tgt_lens)
        ctx.step, ctx.serial = step, step.forward_serial
        return logits

    @staticmethod
    def backward(ctx, d_logits):
        sv = ctx.step.saved
        if sv is None or sv.get("serial") != ctx.serial:
            raise ops.DigError("backward of a stale forward: dig_b200 keeps the activations of the most recent forward only")
        return (None, None, None, None) + tuple(ctx.step.backward(d_logits))


class _SeqCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, targets, lens):
        B, T, C = logits.shape
        lg = logits if logits.stride(-1) == 1 and logits.stride(1) == logits.stride(0) // T else logits.contiguous()
        ld = lg.stride(1)
        loss = torch.zeros(1, dtype=torch.float32, device=logits.device)
        dl = torch.empty(B * T, C, dtype=torch.float32, device=logits.device)
        pred = torch.empty(B * T, dtype=torch.int32, device=logits.device)
        call("dig_seq_cross_entropy", lg, ld, targets.contiguous(), lens.contiguous(), loss, dl, C, pred, B, T, C)
        ctx.save_for_backward(dl)
        ctx.shape = (B, T, C)
        ctx.mark_non_differentiable(pred)
        return loss.reshape(()), pred.view(B, T)

    @staticmethod
    def backward(ctx, g, _gp):
        (dl,) = ctx.saved_tensors
        out = torch.empty_like(dl)
        call("dig_scale_by_device_scalar", dl, g.reshape(1).to(torch.float32).contiguous(), out, dl.numel())
        return out.view(ctx.shape), None, None


def seq_cross_entropy(logits, targets, lens):
    """SeqCrossEntropyLoss(sample_normalize=True) (loss/seqCrossEntropyLoss.py:19-63) -> (loss, arg-max predictions [B,T])."""
    if not logits.is_cuda:
        raise ops.DigError("seq_cross_entropy runs on CUDA tensors only")
    return _SeqCE.apply(logits.float(), targets.to(logits.device, torch.int64), lens.to(logits.device, torch.int64))


class SeqCrossEntropyLoss(nn.Module):
    """Drop-in for loss.SeqCrossEntropyLoss (the criterion run_class_finetuning.py hands to train_one_epoch)."""

    def __init__(self, weight=None, size_average=True, ignore_index=-100, sequence_normalize=False, sample_normalize=True):
        super().__init__()
        if sequence_normalize or not sample_normalize or weight is not None:
            raise NotImplementedError("only the default sample_normalize=True form is built")

    def forward(self, input, target, length):
        return seq_cross_entropy(input, target, length)[0]


def create_rec_model(args):
    """models.model_builder.RecModel(args) (run_class_finetuning.py builds it from its argparse namespace)."""
    return DigRecModel(encoder_name=args.model, nb_classes=args.nb_classes, max_len=args.max_len, decoder_name=args.decoder_name,
                       drop=getattr(args, "drop", 0.0), drop_path=getattr(args, "drop_path", 0.0),
                       attn_drop_rate=getattr(args, "attn_drop_rate", 0.0), decoder_dropout=getattr(args, "decoder_dropout", 0.0))

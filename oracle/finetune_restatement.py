"""TEST INFRASTRUCTURE ONLY -- fp32 functional restatement of DiG's fine-tuning forward (BASELINE configs[4]):
`RecModel.forward` in train mode (models/model_builder.py:124-169) = ViT encoder with its final LayerNorm (modeling_pretrain_vit.py:89-106)
-> `linear_norm` (model_builder.py:86-89) -> `TFDecoder.forward_train` (models/decoder.py:196-222: six pre-LN TransformerDecoderLayers,
models/transformer_layer.py:47-118, MultiHeadAttention :204-281, PositionwiseFeedForward :386-404) -> `SeqCrossEntropyLoss`
(loss/seqCrossEntropyLoss.py:19-63), all dropout probabilities 0 (the configuration dig_b200 builds; the README's 0.1 rates need RNG-matched
dropout and are not built).  Tensors are keyed by the reference's state-dict names.  Pinned against the unmodified reference by
oracle/make_golden_finetune.py -> tests/golden/ref_finetune_*.pt.  Only tests / smoke / bench's CPU legs may import this file.
"""
import numpy as np
import torch
import torch.nn.functional as Fn

from . import restatement as R

DEC_LN_EPS = 1e-5          # nn.LayerNorm default inside TransformerDecoderLayer (transformer_layer.py:64-66) and linear_norm
DEC_FINAL_LN_EPS = 1e-6    # decoder.py:168


def encoder_forward(sd, images, heads, depth=R.DEPTH):
    """V:89-106 without mask, WITH the final LayerNorm (self.norm, eps 1e-6 from the factory's norm_layer)."""
    x = R.encoder_forward(sd, "encoder.", images, None, heads, depth)
    return Fn.layer_norm(x, (x.shape[-1],), sd["encoder.norm.weight"], sd["encoder.norm.bias"], R.LN_EPS)


def mha(sd, p, q_in, kv_in, n_head, mask=None):
    """transformer_layer.py:241-281: bias-free q/k/v/fc Linears, d_k = d_v = d_model / n_head, scale d_k^-0.5, masked_fill(-inf), softmax."""
    B, Lq, D = q_in.shape
    Lk = kv_in.shape[1]
    dk = D // n_head
    q = Fn.linear(q_in, sd[p + "linear_q.weight"]).view(B, Lq, n_head, dk).permute(0, 2, 1, 3)
    k = Fn.linear(kv_in, sd[p + "linear_k.weight"]).view(B, Lk, n_head, dk).permute(0, 2, 3, 1)
    v = Fn.linear(kv_in, sd[p + "linear_v.weight"]).view(B, Lk, n_head, dk).permute(0, 2, 1, 3)
    logits = torch.matmul(q, k) * (dk ** -0.5)
    if mask is not None:
        logits = logits.masked_fill(mask.unsqueeze(1) == 0, float("-inf"))
    w = logits.softmax(dim=-1)
    out = torch.matmul(w, v).transpose(1, 2).reshape(B, Lq, D)
    return Fn.linear(out, sd[p + "fc.weight"]), w.mean(1)


def decoder_layer(sd, p, x, mem, self_mask, n_head):
    """transformer_layer.py:98-118."""
    D = x.shape[-1]
    h = Fn.layer_norm(x, (D,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], DEC_LN_EPS)
    a, _ = mha(sd, p + "self_attn.", h, h, n_head, self_mask)
    x = x + a
    h = Fn.layer_norm(x, (D,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], DEC_LN_EPS)
    a, maps = mha(sd, p + "enc_attn.", h, mem, n_head, None)
    x = x + a
    h = Fn.layer_norm(x, (D,), sd[p + "norm3.weight"], sd[p + "norm3.bias"], DEC_LN_EPS)
    h = Fn.linear(Fn.gelu(Fn.linear(h, sd[p + "mlp.w_1.weight"], sd[p + "mlp.w_1.bias"])), sd[p + "mlp.w_2.weight"], sd[p + "mlp.w_2.bias"])
    return x + h, maps


def self_attention_mask(tgt_lens, T, device):
    """get_pad_mask(seq, tgt_lens) & get_subsequent_mask(seq) (transformer_layer.py:433-456): [B, T, T], True = attend."""
    pad = torch.arange(T, device=device)[None, :] < tgt_lens.to(device)[:, None]          # [B, T] valid key positions
    causal = torch.tril(torch.ones(T, T, dtype=torch.bool, device=device))
    return pad[:, None, :] & causal[None]


def rec_forward(sd, images, targets, tgt_lens, heads, n_head=8, n_layers=6, num_classes=97):
    """RecModel.forward (train mode) -> logits [B, T, num_classes], cross-attention maps of the last layer [B, T, 256]."""
    enc = encoder_forward(sd, images, heads)
    mem = Fn.linear(enc, sd["linear_norm.0.weight"], sd["linear_norm.0.bias"])
    mem = Fn.layer_norm(mem, (mem.shape[-1],), sd["linear_norm.1.weight"], sd["linear_norm.1.bias"], DEC_LN_EPS)
    B, T = targets.shape
    start = torch.full((B, 1), num_classes, dtype=targets.dtype, device=targets.device)                # decoder.py:213 start_idx = num_classes
    query = torch.cat([start, targets], dim=-1)[:, :-1]                                              # decoder.py:214
    x = Fn.embedding(query, sd["decoder.trg_word_emb.weight"]) + sd["decoder.position_enc.position_table"][:, :T]
    mask = self_attention_mask(tgt_lens, T, x.device)
    maps = None
    for l in range(n_layers):
        x, maps = decoder_layer(sd, "decoder.layer_stack.%d." % l, x, mem, mask, n_head)
    x = Fn.layer_norm(x, (x.shape[-1],), sd["decoder.layer_norm.weight"], sd["decoder.layer_norm.bias"], DEC_FINAL_LN_EPS)
    return Fn.linear(x, sd["decoder.classifier.weight"], sd["decoder.classifier.bias"]), maps


def encode_memory(sd, images, heads):
    """model_builder.py:126-146: encoder (+ final norm) -> linear_norm: the decoder's memory [B, 256, d_model]."""
    enc = encoder_forward(sd, images, heads)
    mem = Fn.linear(enc, sd["linear_norm.0.weight"], sd["linear_norm.0.bias"])
    return Fn.layer_norm(mem, (mem.shape[-1],), sd["linear_norm.1.weight"], sd["linear_norm.1.bias"], DEC_LN_EPS)


def greedy_decode(sd, images, heads, n_head=8, n_layers=6, num_classes=97, max_len=25, force_tokens=None):
    """RecModel.forward in eval mode with beam_width 0 = TFDecoder.forward_test (models/decoder.py:224-250): the target sequence starts as
    [<BOS>, 0, 0, ...] (max_len + 1 positions); step t runs the whole decoder with tgt_lens = t + 1, takes softmax(classifier(output[:, t])),
    and writes its arg-max into position t + 1.  Returns (probabilities [B, max_len, C], cross-attention maps [B, max_len, 256], tokens
    [B, max_len]).  force_tokens [B, max_len] (tests only) replaces the fed-back arg-max."""
    mem = encode_memory(sd, images, heads)
    B, dev = images.shape[0], images.device
    T1 = max_len + 1
    seq = torch.zeros(B, T1, dtype=torch.long, device=dev)
    seq[:, 0] = num_classes                                                   # start_idx (decoder.py:229)
    probs, maps_out, toks = [], [], []
    pos = sd["decoder.position_enc.position_table"][:, :T1]
    for step in range(max_len):
        lens = torch.full((B,), step + 1, dtype=torch.long, device=dev)
        x = Fn.embedding(seq, sd["decoder.trg_word_emb.weight"]) + pos
        mask = self_attention_mask(lens, T1, dev)
        maps = None
        for l in range(n_layers):
            x, maps = decoder_layer(sd, "decoder.layer_stack.%d." % l, x, mem, mask, n_head)
        x = Fn.layer_norm(x, (x.shape[-1],), sd["decoder.layer_norm.weight"], sd["decoder.layer_norm.bias"], DEC_FINAL_LN_EPS)
        p = Fn.softmax(Fn.linear(x[:, step], sd["decoder.classifier.weight"], sd["decoder.classifier.bias"]), dim=-1)
        probs.append(p)
        maps_out.append(maps[:, step])
        nxt = p.argmax(-1) if force_tokens is None else force_tokens[:, step].to(dev)
        toks.append(p.argmax(-1))
        seq[:, step + 1] = nxt
    return torch.stack(probs, 1), torch.stack(maps_out, 1), torch.stack(toks, 1)


def seq_cross_entropy(logits, targets, tgt_lens):
    """loss/seqCrossEntropyLoss.py:47-63 with sample_normalize=True: sum over valid positions of -log p[target] / batch size."""
    B, T, C = logits.shape
    mask = (torch.arange(T, device=logits.device)[None, :] < tgt_lens.to(logits.device)[:, None]).reshape(-1, 1)
    logp = Fn.log_softmax(logits.reshape(-1, C), dim=1)
    out = -logp.gather(1, targets.reshape(-1, 1).long()) * mask
    return out.sum() / B


def trainable_names(sd):
    return [k for k in sd if k != "decoder.position_enc.position_table" and not k.startswith("patch_embed.")]


def synthetic_batch(B, seed=1, T=25, num_chars=94):
    g = torch.Generator().manual_seed(seed)
    img = torch.rand(B, 3, R.IMG_H, R.IMG_W, generator=g) * 2 - 1
    lens = torch.randint(1, T + 1, (B,), generator=g)
    tgt = torch.randint(0, num_chars, (B, T), generator=g)
    for b in range(B):          # EOS (94) at the last valid position, PADDING (95) beyond (dataset_lmdb.py label layout)
        tgt[b, lens[b] - 1] = num_chars
        tgt[b, lens[b]:] = num_chars + 1
    return img, tgt, lens

"""TEST INFRASTRUCTURE ONLY -- golden fixtures of the fine-tuning step (row f2) from the UNMODIFIED reference RecModel
(models/model_builder.py:74-169) + SeqCrossEntropyLoss, CPU fp32, every nn.Dropout set to p = 0, train mode.
    python oracle/make_golden_finetune.py      (needs /root/reference)
"""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shims  # noqa: E402
from oracle import finetune_restatement as FR  # noqa: E402


def reference_rec_model(model_name, seed=0):
    mb, crit = ref_shims.import_reference_finetune()
    args = types.SimpleNamespace(model=model_name, nb_classes=97, max_len=25, decoder_name="tf_decoder", text_cond_vis=False, drop=0.0,
                                 drop_path=0.0, attn_drop_rate=0.0, use_mean_pooling=False, init_scale=0.001, use_seq_cls_token=False,
                                 use_1d_attdec=False, beam_width=0)
    torch.manual_seed(seed)
    m = mb.RecModel(args)
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    return m.train(), crit()


def reference_step(model_name, B, seed_model=0, seed_data=1):
    model, crit = reference_rec_model(model_name, seed_model)
    sd0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
    img, tgt, lens = FR.synthetic_batch(B, seed=seed_data)
    out = model((img, tgt, lens))
    loss = crit(out[0], tgt, lens)
    loss.backward()
    grads = {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}
    # eval mode, beam_width 0: TFDecoder.forward_test (greedy decoding, decoder.py:224-250) of the same images
    with torch.no_grad():
        ev = model.eval()((img, tgt, lens))
    model.train()
    return dict(eval_probs=ev[0].detach().clone(), eval_maps=ev[3].detach().clone(),
                model=model_name, B=B, seed_model=seed_model, seed_data=seed_data, loss=float(loss), logits=out[0].detach().clone(),
                attn_maps=out[3].detach().clone(), state_keys=[(k, tuple(v.shape), str(v.dtype)) for k, v in sd0.items()],
                param_checksum={k: float(v.double().sum()) for k, v in sd0.items() if v.dtype.is_floating_point},
                grad_norms={n: float(g.norm()) for n, g in grads.items()},
                no_grad=[n for n, p in model.named_parameters() if p.grad is None],
                grad_samples={n: grads[n].flatten()[:64].clone() for n in
                              ("encoder.blocks.0.attn.qkv.weight", "encoder.norm.weight", "linear_norm.0.weight", "decoder.trg_word_emb.weight",
                               "decoder.layer_stack.0.self_attn.linear_q.weight", "decoder.layer_stack.5.enc_attn.linear_k.weight",
                               "decoder.layer_stack.3.mlp.w_1.weight", "decoder.classifier.weight")})


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    for name, B, tag in (("simmim_vit_tiny_patch4_32x128", 3, "tiny_b3"), ("simmim_vit_small_patch4_32x128", 4, "small_b4")):
        g = reference_step(name, B)
        torch.save(g, os.path.join(out_dir, "ref_finetune_%s.pt" % tag))
        print(tag, "loss %.6f" % g["loss"], "params without grad:", g["no_grad"])


if __name__ == "__main__":
    main()

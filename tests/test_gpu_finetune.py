"""-m gpu: the fine-tuning step through DigRecModel / seq_cross_entropy / train_one_epoch against the golden
fixtures written from the unmodified reference RecModel + SeqCrossEntropyLoss (dropout 0) and the fp32 oracle.
Tolerances: loss 1e-3 relative (north_star), logits 5e-2 absolute (bf16 operands, values of magnitude ~1), gradients per tensor cosine."""
import os
import types

import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _model(name):
    import __graft_entry__ as ge
    ge.build()
    from dig_b200.finetune import DigRecModel
    torch.manual_seed(0)
    return DigRecModel(name).train()


def test_decoder_kernels_against_torch():
    import __graft_entry__ as ge
    ge.build()
    from dig_b200 import ops
    torch.manual_seed(1)
    B, H, T, Lk, D = 5, 8, 25, 256, 512
    for (Lq, L2, lens) in ((T, T, torch.tensor([1, 25, 7, 13, 20], device="cuda")), (T, Lk, None)):
        q = (torch.randn(B * Lq, D, device="cuda") * 0.5).bfloat16()
        k = (torch.randn(B * L2, D, device="cuda") * 0.5).bfloat16()
        v = (torch.randn(B * L2, D, device="cuda") * 0.5).bfloat16()
        qf, kf, vf = [t.float().requires_grad_(True) for t in (q, k, v)]
        qh = qf.view(B, Lq, H, 64).permute(0, 2, 1, 3)
        kh = kf.view(B, L2, H, 64).permute(0, 2, 1, 3)
        vh = vf.view(B, L2, H, 64).permute(0, 2, 1, 3)
        logit = qh @ kh.transpose(-1, -2) * 0.125
        if lens is not None:
            m = (torch.arange(L2, device="cuda")[None, :] < lens[:, None])[:, None, :] & torch.tril(torch.ones(Lq, L2, dtype=torch.bool, device="cuda"))[None]
            logit = logit.masked_fill(~m[:, None], float("-inf"))
        w = logit.softmax(-1)
        ref = (w @ vh).permute(0, 2, 1, 3).reshape(B * Lq, D)
        out = torch.empty(B * Lq, D, device="cuda", dtype=torch.bfloat16)
        lse = torch.empty(B, H, Lq, device="cuda")
        maps = torch.zeros(B, Lq, L2, device="cuda")
        ops.call("dig_dec_attention_fwd", q, D, k, D, v, D, out, D, lse, lens, maps, B, H, Lq, L2, 0.125)
        assert torch.allclose(out.float(), ref, atol=2e-2)
        assert torch.allclose(maps, w.mean(1), atol=1e-4)
        do = (torch.randn(B * Lq, D, device="cuda") * 0.5).bfloat16()
        ref.backward(do.float())
        dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        ops.call("dig_dec_attention_bwd", q, D, k, D, v, D, out, D, do, D, lse, lens, dq, D, dk, D, dv, D, B, H, Lq, L2, 0.125)
        for got, want in ((dq, qf.grad), (dk, kf.grad), (dv, vf.grad)):
            assert torch.allclose(got.float(), want, atol=3e-2, rtol=3e-2)
    # embedding + position, sequence cross entropy
    from oracle import finetune_restatement as FR
    from dig_b200.finetune import seq_cross_entropy
    _, tgt, lens = FR.synthetic_batch(6, seed=2)
    tgt, lens = tgt.cuda(), lens.cuda()
    logits = torch.randn(6, 25, 97, device="cuda", requires_grad=True)
    loss, pred = seq_cross_entropy(logits, tgt, lens)
    ref = FR.seq_cross_entropy(logits.detach().clone().requires_grad_(True), tgt, lens)
    assert float(loss) == pytest.approx(float(ref), rel=1e-5)
    loss.backward()
    lg2 = logits.detach().clone().requires_grad_(True)
    FR.seq_cross_entropy(lg2, tgt, lens).backward()
    assert torch.allclose(logits.grad, lg2.grad, atol=1e-6)
    assert torch.equal(pred.long(), logits.argmax(-1))


@pytest.mark.parametrize("tag", ["tiny_b3", "small_b4"])
def test_finetune_step_matches_reference_golden(tag):
    from oracle import finetune_restatement as FR
    from dig_b200.finetune import seq_cross_entropy
    g = torch.load(os.path.join(GOLD, "ref_finetune_%s.pt" % tag), weights_only=False)
    model = _model(g["model"])
    sd = model.state_dict()
    assert [(k, tuple(v.shape), str(v.dtype)) for k, v in sd.items()] == g["state_keys"]
    for k, v in sd.items():
        if v.dtype.is_floating_point:
            assert float(v.double().sum()) == pytest.approx(g["param_checksum"][k], rel=1e-9, abs=1e-9), k     # bit-identical init
    model.cuda()
    img, tgt, lens = FR.synthetic_batch(g["B"], seed=g["seed_data"])
    out = model((img.cuda(), tgt.cuda(), lens.cuda()))
    assert out[1] is None and out[2] is None and tuple(out[0].shape) == tuple(g["logits"].shape)
    loss, pred = seq_cross_entropy(out[0], tgt.cuda(), lens.cuda())
    assert float(loss) == pytest.approx(g["loss"], rel=1e-3)
    assert torch.allclose(out[0].float().cpu(), g["logits"], atol=5e-2)
    loss.backward()
    named = dict(model.named_parameters())
    assert named["encoder.mask_token"].grad is None                                       # unused in fine-tuning, as in the reference
    worst = 1.0
    for n, ref in g["grad_norms"].items():
        gn = float(named[n].grad.float().norm())
        assert gn == pytest.approx(ref, rel=6e-2, abs=1e-6), (n, gn, ref)
    for n, ref in g["grad_samples"].items():
        got = named[n].grad.flatten()[:64].float().cpu()
        if float(ref.norm()) == 0.0:          # e.g. the embedding row of a token that does not occur in the batch
            assert float(got.norm()) == 0.0, n
            continue
        cos = float(torch.nn.functional.cosine_similarity(got, ref, dim=0))
        worst = min(worst, cos)
        assert cos > 0.99, (n, cos)
    # attention maps of the last decoder layer (what RecModel.forward returns as dec_attn_maps)
    step = model._step
    with torch.no_grad():
        step.forward(img.cuda(), tgt.cuda(), lens.cuda(), need_maps=True)
    assert torch.allclose(step.last_maps.cpu(), g["attn_maps"], atol=2e-3)


def test_finetune_train_one_epoch_learns():
    from oracle import finetune_restatement as FR
    from dig_b200.engine_finetune import train_one_epoch
    from dig_b200.finetune import SeqCrossEntropyLoss
    from dig_b200.optim import FusedAdamW
    from dig_b200.utils import NativeScalerWithGradNormCount
    model = _model("simmim_vit_tiny_patch4_32x128").cuda()
    opt = FusedAdamW([{"params": [p for p in model.parameters() if p.requires_grad], "weight_decay": 0.05, "lr_scale": 1.0}], lr=1e-3)
    img, tgt, lens = FR.synthetic_batch(8, seed=3)
    loader = [(img, tgt, lens)] * 12
    args = types.SimpleNamespace(w2v_path=None, use_seq_cls_token=False, eval_freq=10 ** 9)
    st0 = train_one_epoch(model, SeqCrossEntropyLoss(), loader, opt, torch.device("cuda"), 0, NativeScalerWithGradNormCount(), max_norm=None,
                          start_steps=0, num_training_steps_per_epoch=12, update_freq=1, args=args)
    st1 = train_one_epoch(model, SeqCrossEntropyLoss(), loader, opt, torch.device("cuda"), 1, NativeScalerWithGradNormCount(), max_norm=None,
                          start_steps=12, num_training_steps_per_epoch=12, update_freq=1, args=args)
    assert {"loss", "class_acc", "loss_scale", "lr", "min_lr", "weight_decay", "grad_norm", "max_accuracy"} <= set(st1)
    assert st1["loss"] < st0["loss"] and all(v == v for v in st1.values() if v is not None)


class _ToyDataset:
    """class_to_idx / idx_to_class of dataset_lmdb.py's 94 characters + EOS / PADDING / UNKNOWN."""

    def __init__(self):
        import string
        voc = list(string.digits + string.ascii_lowercase + string.ascii_uppercase + string.punctuation) + ["EOS", "PADDING", "UNKNOWN"]
        self.class_to_idx = {c: i for i, c in enumerate(voc)}
        self.idx_to_class = {i: c for i, c in enumerate(voc)}


@pytest.mark.parametrize("tag", ["tiny_b3", "small_b4"])
def test_greedy_decoding_matches_reference_eval_mode(tag):
    """Eval mode = TFDecoder.forward_test (decoder.py:224-250).  (a) With the reference's own fed-back tokens forced, the step probabilities
    equal the fixture of the unmodified reference; (b) the free-running decode is self-consistent: teacher-forcing its tokens through the
    training forward reproduces the same distributions at every position; (c) RecModel's eval-mode return tuple."""
    from oracle import finetune_restatement as FR
    g = torch.load(os.path.join(GOLD, "ref_finetune_%s.pt" % tag), weights_only=False)
    model = _model(g["model"]).cuda()
    img, tgt, lens = FR.synthetic_batch(g["B"], seed=g["seed_data"])
    step = model._pipeline()
    ref_tokens = g["eval_probs"].argmax(-1)
    probs, maps, _ = step.greedy_decode(img.cuda(), need_maps=True, force_tokens=ref_tokens.cuda())
    assert torch.allclose(probs.cpu(), g["eval_probs"], atol=2e-3, rtol=5e-2)
    assert torch.allclose(maps.cpu(), g["eval_maps"], atol=2e-3)
    model.eval()
    with torch.no_grad():
        out = model((img.cuda(), tgt.cuda(), lens.cuda()))
    assert out[1] is None and out[2] is None and tuple(out[0].shape) == (g["B"], 25, 97)
    assert torch.allclose(out[0].sum(-1), torch.ones(g["B"], 25, device="cuda"), atol=1e-4)
    toks = out[0].argmax(-1)
    model.train()
    with torch.no_grad():
        logits = model((img.cuda(), toks, torch.full((g["B"],), 25, device="cuda")))[0]
    assert torch.allclose(torch.softmax(logits.float(), -1), out[0], atol=1e-4)


def test_finetune_evaluate_contract():
    from oracle import finetune_restatement as FR
    from dig_b200.engine_finetune import evaluate, recognition_fmeasure, word_accuracy
    ds = _ToyDataset()
    t = torch.tensor([[10, 11, 94, 95, 95], [1, 96, 2, 94, 95]])          # "ab<EOS>", "1<UNK>2<EOS>"
    assert word_accuracy(t.clone(), t, ds) == 1.0 and recognition_fmeasure(t.clone(), t, ds) == pytest.approx(1.0, abs=1e-4)
    p = torch.tensor([[10, 12, 94, 0, 0], [1, 2, 94, 95, 95]])            # "ac" vs "ab": wrong; "12" vs "12" (UNKNOWN dropped): right
    assert word_accuracy(p, t, ds) == 0.5
    model = _model("simmim_vit_tiny_patch4_32x128").cuda()
    img, tgt, lens = FR.synthetic_batch(6, seed=5)

    class Loader(list):
        dataset = ds
    stats = evaluate(Loader([(img, tgt, lens)] * 2), model, torch.device("cuda"), args=types.SimpleNamespace(beam_width=0))
    assert {"loss", "acc", "recognition_fmeasure"} <= set(stats) and all(v == v for v in stats.values())
    assert 0.0 <= stats["acc"] <= 1.0 and not model.training

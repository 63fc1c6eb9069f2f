"""Fine-tuning step: the fp32 restatement against the golden fixtures written from the unmodified reference
RecModel + SeqCrossEntropyLoss (oracle/make_golden_finetune.py), and against the live reference where present."""
import os

import pytest
import torch

from oracle import finetune_restatement as FR

GOLD = os.path.join(os.path.dirname(__file__), "golden")
HEADS = {"simmim_vit_tiny_patch4_32x128": 3, "simmim_vit_small_patch4_32x128": 6}


def _reference_state(g):
    from oracle import ref_shims
    if not ref_shims.reference_available():
        pytest.skip("reference not available")
    from oracle.make_golden_finetune import reference_rec_model
    model, _ = reference_rec_model(g["model"], g["seed_model"])
    return {k: v.detach().clone() for k, v in model.state_dict().items()}


@pytest.mark.parametrize("tag", ["tiny_b3", "small_b4"])
def test_finetune_oracle_reproduces_reference_golden(tag):
    g = torch.load(os.path.join(GOLD, "ref_finetune_%s.pt" % tag), weights_only=False)
    sd = _reference_state(g)
    for k, v in sd.items():
        if v.dtype.is_floating_point:
            assert float(v.double().sum()) == pytest.approx(g["param_checksum"][k], rel=1e-9, abs=1e-9)
    names = FR.trainable_names(sd)
    for n in names:
        sd[n].requires_grad_(True)
    img, tgt, lens = FR.synthetic_batch(g["B"], seed=g["seed_data"])
    logits, maps = FR.rec_forward(sd, img, tgt, lens, HEADS[g["model"]])
    loss = FR.seq_cross_entropy(logits, tgt, lens)
    assert float(loss) == pytest.approx(g["loss"], rel=2e-5)
    assert torch.allclose(logits, g["logits"], atol=2e-4) and torch.allclose(maps, g["attn_maps"], atol=1e-5)
    grads = dict(zip(names, torch.autograd.grad(loss, [sd[n] for n in names], allow_unused=True)))
    assert grads["encoder.mask_token"] is None and g["no_grad"] == ["encoder.mask_token"]      # unused in fine-tuning (no mask), SURVEY f2
    for n, ref in g["grad_norms"].items():
        assert float(grads[n].norm()) == pytest.approx(ref, rel=2e-3, abs=1e-7), n
    for n, ref in g["grad_samples"].items():
        assert torch.allclose(grads[n].flatten()[:64], ref, atol=1e-4, rtol=2e-3), n


@pytest.mark.parametrize("tag", ["tiny_b3", "small_b4"])
def test_greedy_decoding_restatement_reproduces_reference_eval_mode(tag):
    """RecModel.forward in eval mode (TFDecoder.forward_test, models/decoder.py:224-250): step probabilities and cross-attention maps of the
    restated greedy decoding against the unmodified reference's (fixture written by oracle/make_golden_finetune.py)."""
    g = torch.load(os.path.join(GOLD, "ref_finetune_%s.pt" % tag), weights_only=False)
    sd = _reference_state(g)
    img, _, _ = FR.synthetic_batch(g["B"], seed=g["seed_data"])
    with torch.no_grad():
        probs, maps, toks = FR.greedy_decode(sd, img, HEADS[g["model"]])
    assert probs.shape == g["eval_probs"].shape == (g["B"], 25, 97)
    assert torch.allclose(probs, g["eval_probs"], atol=2e-6) and torch.allclose(maps, g["eval_maps"], atol=1e-5)
    assert torch.equal(toks, g["eval_probs"].argmax(-1))
    assert torch.allclose(probs.sum(-1), torch.ones(g["B"], 25), atol=1e-5)

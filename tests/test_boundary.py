"""Drop-in boundary B1/B2: factories, state-dict contract, init parity with the reference (via golden fixtures)."""
import os

import pytest
import torch

GOLD = os.path.join(os.path.dirname(__file__), "golden")
KW = dict(pretrained=False, drop_path_rate=0.0, drop_block_rate=None, mlp_dim=4096, dim=256, T=0.2, num_windows=4, encoder_type="vit",
          queue_size=65536, patchnet_name="no_patchtrans")


def make(name, **over):
    import dig_b200
    from dig_b200 import modeling  # noqa: F401
    torch.manual_seed(0)
    kw = dict(KW)
    kw.update(over)
    return dig_b200.create_model(name, **kw)


@pytest.mark.parametrize("tag,name", [("small_b2", "pretrain_simmim_moco_ori_vit_small_patch4_32x128"),
                                      ("base_b2", "pretrain_simmim_moco_ori_vit_base_patch4_32x128")])
def test_state_dict_contract_and_init_match_reference(tag, name):
    g = torch.load(os.path.join(GOLD, "ref_step_%s.pt" % tag), weights_only=False)
    m = make(name)
    sd = m.state_dict()
    assert [(k, tuple(v.shape), str(v.dtype)) for k, v in sd.items()] == g["state_keys"]
    assert "encoder.pos_embed" not in sd and "pos_embed" not in sd          # plain attribute in the reference (V:48)
    assert sum(p.numel() for p in m.parameters() if p.requires_grad) == g["trainable"]
    assert sum(p.numel() for p in m.parameters() if not p.requires_grad) == g["frozen"]
    # same RNG consumption order as the reference => identical initial parameters for a given seed
    for k, ref in g["param_checksum"].items():
        assert float(sd[k].double().sum()) == pytest.approx(ref, rel=1e-12, abs=1e-12), k


def test_small_counts_match_survey():
    m = make("pretrain_simmim_moco_ori_vit_small_patch4_32x128")
    assert sum(p.numel() for p in m.parameters() if p.requires_grad) == 43606192
    assert len(m.state_dict()) == 398
    assert len([p for p in m.parameters() if p.requires_grad]) == 183
    assert len([p for p in m.parameters() if not p.requires_grad]) == 173


def test_runner_facing_attributes():
    m = make("pretrain_simmim_moco_ori_vit_tiny_patch4_32x128")
    assert m.encoder.patch_embed.patch_size == (4, 4)                      # run_mae_pretraining_moco.py:320-323
    assert m.no_weight_decay() == {"pos_embed", "cls_token"}
    assert m.encoder.embed_dim == 192 and m.encoder.num_heads == 3
    for n, p in m.named_parameters():
        assert p.requires_grad == (not n.startswith(("momentum_", "pix_projector_m")))


def test_factory_tolerates_timm_kwargs_and_drops_none():
    m = make("pretrain_simmim_moco_ori_vit_tiny_patch4_32x128", num_classes=1000, in_chans=3)
    assert isinstance(m, torch.nn.Module)
    import dig_b200
    with pytest.raises(RuntimeError):
        dig_b200.create_model("no_such_model")


def test_unsupported_configurations_raise():
    with pytest.raises(NotImplementedError):
        make("pretrain_simmim_moco_ori_vit_tiny_patch4_32x128", patchnet_name="regular")
    with pytest.raises(NotImplementedError):
        make("pretrain_simmim_moco_ori_vit_tiny_patch4_32x128", drop_path_rate=0.1)


def test_cpu_forward_is_refused():
    m = make("pretrain_simmim_moco_ori_vit_tiny_patch4_32x128")
    x = torch.zeros(2, 3, 32, 128)
    with pytest.raises(RuntimeError):
        m(x, x, torch.zeros(2, 2, 256, dtype=torch.bool), 0.99)


def test_root_drop_in_modules():
    import engine_for_pretraining_moco as E
    import modeling_pretrain_moco_mim_ori as M
    import inspect
    assert hasattr(M, "pretrain_simmim_moco_ori_vit_small_patch4_32x128") and hasattr(M, "MoCo_ViT")
    params = list(inspect.signature(E.train_one_epoch).parameters)
    assert params == ["model", "teacher_model", "teacher_model_without_ddp", "data_loader", "word_data_loader", "optimizer", "device",
                      "epoch", "loss_scaler", "max_norm", "patch_size", "normlize_target", "log_writer", "lr_scheduler", "start_steps",
                      "lr_schedule_values", "wd_schedule_values", "momentum_schedule", "args"]


def test_checkpoint_round_trip(tmp_path):
    import types
    from dig_b200 import checkpoint
    from dig_b200.utils import NativeScalerWithGradNormCount
    m = make("pretrain_simmim_moco_ori_vit_tiny_patch4_32x128")
    opt = torch.optim.SGD([p for p in m.parameters() if p.requires_grad], lr=0.1)
    args = types.SimpleNamespace(output_dir=str(tmp_path), resume="", auto_resume=True, start_epoch=0)
    checkpoint.save_model(args, 3, m, m, opt, NativeScalerWithGradNormCount())
    m2 = make("pretrain_simmim_moco_ori_vit_tiny_patch4_32x128")
    with torch.no_grad():
        for p in m2.parameters():
            p.add_(1.0)
    checkpoint.auto_load_model(args, m2, m2, opt, NativeScalerWithGradNormCount())
    assert args.start_epoch == 4
    for (k, a), (_, b) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert torch.equal(a, b), k

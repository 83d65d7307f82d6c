"""Checkpoint / latent-stats ingestion (host logic, CPU): the reference's selection and shape-check rules."""
import torch

import va_vae_b200 as V
from va_vae_b200 import checkpoint as C

KW = dict(input_size=4, patch_size=1, in_channels=32, hidden_size=64, depth=1, num_heads=2, num_classes=10,
          use_swiglu=True, use_rope=True, use_rmsnorm=True)


def _rand_sd(model, seed=0):
    g = torch.Generator().manual_seed(seed)
    return {k: torch.randn(v.shape, generator=g) for k, v in model.state_dict().items()}


def test_ema_preferred_and_module_prefix_stripped(tmp_path):
    m = V.LightningDiT(**KW)
    ema, raw = _rand_sd(m, 1), _rand_sd(m, 2)
    ckpt = {"model": {"module." + k: v for k, v in raw.items()}, "ema": {"module." + k: v for k, v in ema.items()},
            "opt": {}, "config": {}}
    path = tmp_path / "0001000.pt"
    torch.save(ckpt, path)
    C.load_checkpoint(m, str(path), verbose=False)
    for k, v in m.state_dict().items():
        assert torch.equal(v, ema[k]), k
    # no 'ema' -> 'model'; neither -> bare state_dict
    assert C.select_weights({"model": {"a": 1}}) == {"a": 1}
    assert C.select_weights({"module.a": 1}) == {"a": 1}
    # only a LEADING 'module.' is DDP's prefix
    assert C.select_weights({"module.blocks.0.submodule.w": 1}) == {"blocks.0.submodule.w": 1}


def test_sampling_loader_is_strict_like_the_reference(tmp_path):
    """inference.py:91 calls load_state_dict strictly: a checkpoint of another configuration must raise, not sample from
    partly random weights; strict=False is the opt-in fine-tuning behaviour (train.py:275-301)."""
    import argparse
    import pytest
    m = V.LightningDiT(**KW)
    sd = _rand_sd(m, 4)
    sd.pop("blocks.0.attn.proj.bias")
    path = tmp_path / "partial.pt"
    torch.save({"ema": sd}, path)
    with pytest.raises(RuntimeError, match="Missing key"):
        C.load_checkpoint(m, str(path))
    C.load_checkpoint(m, str(path), strict=False, verbose=False)
    assert torch.equal(m.state_dict()["blocks.0.attn.qkv.weight"], sd["blocks.0.attn.qkv.weight"])
    # a pickle carrying arbitrary objects is refused unless the caller says it is trusted
    full = _rand_sd(m, 5)
    bad = tmp_path / "with_args.pt"
    torch.save({"ema": full, "args": argparse.Namespace(lr=1e-4)}, bad)
    with pytest.raises(RuntimeError, match="trust_pickle"):
        C.load_checkpoint(m, str(bad))
    C.load_checkpoint(m, str(bad), trust_pickle=True)
    assert torch.equal(m.state_dict()["pos_embed"], full["pos_embed"])


def test_shape_checked_partial_load_and_patch_embed_special_case():
    m = V.LightningDiT(**KW)
    before = {k: v.clone() for k, v in m.state_dict().items()}
    sd = _rand_sd(m, 3)
    sd["blocks.0.attn.qkv.weight"] = torch.randn(7, 7)                     # mismatch -> skipped
    sd["not.a.parameter"] = torch.randn(3)                                  # unknown -> skipped
    sd["x_embedder.proj.weight"] = torch.randn(64, 16, 1, 1)                # fewer input channels -> special case
    C.load_weights_with_shape_check(m, {"model": sd}, verbose=False)
    after = m.state_dict()
    assert torch.equal(after["blocks.0.attn.qkv.weight"], before["blocks.0.attn.qkv.weight"])
    assert torch.equal(after["blocks.0.attn.proj.weight"], sd["blocks.0.attn.proj.weight"])
    w = after["x_embedder.proj.weight"]
    assert torch.equal(w[:, :16], sd["x_embedder.proj.weight"]) and w[:, 16:].abs().sum() == 0


def test_latent_stats_and_denormalisation(tmp_path):
    mean, std = torch.randn(1, 32, 1, 1), torch.rand(1, 32, 1, 1) + 0.5
    path = tmp_path / "latents_stats.pt"
    torch.save({"mean": mean, "std": std}, path)
    m2, s2 = C.load_latent_stats(str(path))
    x = torch.randn(3, 32, 16, 16)
    out = C.denormalize_latents(x, m2, s2, latent_multiplier=0.18215)
    assert torch.allclose(out, (x * std) / 0.18215 + mean)

"""Host-side mirror of the VA-VAE decode interface (va_vae_b200/vavae.py): parameter names / shapes against the
reference module, checkpoint filtering (autoencoder.py:476-486), argument validation, no-CPU-fallback behaviour."""
import os

import pytest
import torch

import va_vae_b200 as V
from va_vae_b200.vavae import decoder_param_shapes
from oracle import vavae_decoder_oracle as D


def test_param_spec_matches_oracle_spec():
    for kw in (dict(), dict(z_channels=4, ch=32, ch_mult=(1, 2), num_res_blocks=1, attn_resolutions=(8,), resolution=16)):
        cfg = D.DecoderConfig(**kw)
        mine = decoder_param_shapes(cfg.z_channels, cfg.ch, cfg.ch_mult, cfg.num_res_blocks, cfg.attn_resolutions,
                                    cfg.resolution, cfg.out_ch)
        assert mine == D.state_dict_spec(cfg)


@pytest.mark.skipif(not os.path.isdir("/root/reference/LightningDiT"), reason="/root/reference not present")
def test_state_dict_keys_match_reference_autoencoder():
    import sys
    sys.path.insert(0, "/root/reference/LightningDiT")
    from tokenizer.autoencoder import AutoencoderKL as RefKL
    ref = RefKL(embed_dim=32, ch_mult=(1, 1, 2, 2, 4))
    ref_sd = {k: v.shape for k, v in ref.state_dict().items() if k.startswith(("decoder.", "post_quant_conv."))}
    mine = V.AutoencoderKL(embed_dim=32, ch_mult=(1, 1, 2, 2, 4))
    assert {k: v.shape for k, v in mine.state_dict().items()} == ref_sd


def test_init_from_ckpt_filters_like_the_reference(tmp_path):
    cfg = D.DecoderConfig(z_channels=4, ch=32, ch_mult=(1, 2), num_res_blocks=1, attn_resolutions=(8,), resolution=16)
    sd = D.init_state_dict(cfg, seed=1)
    extra = {"encoder.conv_in.weight": torch.zeros(3), "loss.logvar": torch.zeros(1),
             "foundation_model.model.x": torch.zeros(2), "quant_conv.weight": torch.zeros(1)}
    path = tmp_path / "vae.pt"
    torch.save({"state_dict": {**sd, **extra}}, path)
    m = V.AutoencoderKL(embed_dim=4, ch=32, ch_mult=(1, 2), num_res_blocks=1, resolution=16, attn_resolutions=(8,),
                        ckpt_path=str(path))
    got = m.state_dict()
    assert set(got) == set(sd) and all(torch.equal(got[k], sd[k]) for k in sd)
    # marvae checkpoints keep their weights under ['model'] and have no attention blocks (autoencoder.py:457-458,481-483)
    cfg2 = D.DecoderConfig(z_channels=4, ch=32, ch_mult=(1, 2), num_res_blocks=1, attn_resolutions=(), resolution=16)
    sd2 = D.init_state_dict(cfg2, seed=2)
    torch.save({"model": sd2}, path)
    m2 = V.AutoencoderKL(embed_dim=4, ch=32, ch_mult=(1, 2), num_res_blocks=1, resolution=16, model_type="marvae",
                         ckpt_path=str(path))
    keys2 = m2.state_dict().keys()       # the mid attention block exists in every Decoder (autoencoder.py:331)
    assert not any(k.startswith("decoder.up") and ".attn." in k for k in keys2) and "decoder.mid.attn_1.q.weight" in keys2
    assert torch.equal(m2.state_dict()["decoder.conv_out.bias"], sd2["decoder.conv_out.bias"])
    with pytest.raises(ValueError, match="Invalid model type"):
        V.AutoencoderKL(model_type="vqgan")


def test_decoder_refuses_cpu_and_encoding():
    m = V.AutoencoderKL(embed_dim=4, ch=32, ch_mult=(1, 2), num_res_blocks=1, resolution=16, attn_resolutions=(8,))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.decode(torch.zeros(1, 4, 8, 8))
    with pytest.raises(NotImplementedError):
        m.encode(torch.zeros(1, 3, 16, 16))

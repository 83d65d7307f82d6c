"""VA-VAE decode path: the CPU oracle against the golden fixtures written from the
unmodified reference module (oracle/gen_golden_decoder.py) and - in the build container only - against the live
reference tokenizer/autoencoder.py."""
import os

import pytest
import torch

from conftest import rel_l2
from oracle import vavae_decoder_oracle as D

DECODER_CASES = ["decoder_tiny", "decoder_small"]


def load_case(golden, name):
    g = golden(name)
    cfg = D.DecoderConfig(**{k: tuple(v) if isinstance(v, (list, tuple)) else v for k, v in g["config"].items()})
    sd = g.get("state_dict") or D.init_state_dict(cfg, seed=g["weight_seed"])
    digest = float(sum(v.double().abs().sum() for v in sd.values()))
    assert abs(digest - g["weight_digest"]) <= 1e-9 * g["weight_digest"], "seeded weights drifted from the fixture's"
    return cfg, sd, g


@pytest.mark.parametrize("name", DECODER_CASES)
def test_oracle_matches_reference_fixture(golden, name):
    cfg, sd, g = load_case(golden, name)
    out = D.decode(cfg, sd, g["z"])
    assert out.shape == g["out"].shape
    assert rel_l2(out, g["out"]) < 1e-6
    img = D.to_images(out)
    assert img.dtype == torch.uint8 and img.shape == g["img"].shape
    # identical float math -> identical truncation, allow the odd value sitting on an integer boundary
    assert (img.int() - g["img"].int()).abs().max() <= 1
    assert (img != g["img"]).float().mean() < 1e-3


def test_state_dict_spec_covers_production_config():
    cfg = D.DecoderConfig()                      # vavae_f16d32.yaml:25-41 + AutoencoderKL(ch_mult=(1,1,2,2,4)), vavae.py:38-41
    spec = dict(D.state_dict_spec(cfg))
    assert spec["decoder.conv_in.weight"] == (512, 32, 3, 3)
    assert spec["decoder.up.4.attn.2.q.weight"] == (512, 512, 1, 1)          # attention only at 16x16
    assert "decoder.up.3.attn.0.q.weight" not in spec
    assert spec["decoder.up.3.block.0.nin_shortcut.weight"] == (256, 512, 1, 1)
    assert spec["decoder.up.1.block.0.nin_shortcut.weight"] == (128, 256, 1, 1)
    assert spec["decoder.up.1.upsample.conv.weight"] == (128, 128, 3, 3) and "decoder.up.0.upsample.conv.weight" not in spec
    assert spec["decoder.conv_out.weight"] == (3, 128, 3, 3)
    assert 2.4e11 < D.flops_per_image(cfg) < 2.6e11


@pytest.mark.skipif(not os.path.isdir("/root/reference/LightningDiT"), reason="/root/reference not present")
def test_oracle_matches_live_reference_module():
    from oracle.gen_golden_decoder import reference_decode
    cfg = D.DecoderConfig(z_channels=8, ch=32, ch_mult=(1, 1, 2), num_res_blocks=2, attn_resolutions=(8, 16),
                          resolution=32)
    sd = D.init_state_dict(cfg, seed=5)
    z = torch.randn(2, 8, 8, 8, generator=torch.Generator().manual_seed(6))
    ref, img = reference_decode(cfg, sd, z)
    out = D.decode(cfg, sd, z)
    assert rel_l2(out, ref) < 1e-6
    assert torch.equal(D.to_images(out), img)

"""TEST INFRASTRUCTURE ONLY -- writes tests/golden/decoder_*.pt FROM THE UNMODIFIED REFERENCE MODULE.

Container only (needs /root/reference).  Imports LightningDiT/tokenizer/autoencoder.py as it is (its dependencies
torchvision / PIL / requests / tqdm are in the image), builds Decoder + post_quant_conv with the fixture's config,
loads the seeded weights, decodes a seeded latent batch on CPU in fp32 and stores config, weights, input, the float
output and the uint8 image the caller-side post-processing (tokenizer/vavae.py:80-81) produces.
Usage: python -m oracle.gen_golden_decoder
"""
import os
import sys

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import vavae_decoder_oracle as D

REF = "/root/reference/LightningDiT"

CASES = {
    # name: (config, weight seed, batch)
    "decoder_tiny": (D.DecoderConfig(z_channels=4, ch=32, ch_mult=(1, 2), num_res_blocks=1, attn_resolutions=(8,),
                                     resolution=16), 11, 2),
    "decoder_small": (D.DecoderConfig(z_channels=32, ch=32, ch_mult=(1, 2, 4), num_res_blocks=2, attn_resolutions=(16,),
                                      resolution=64), 12, 3),
}


def reference_decode(cfg, sd, z):
    sys.path.insert(0, REF)
    from tokenizer.autoencoder import Decoder           # the unmodified reference
    dec = Decoder(ch=cfg.ch, out_ch=cfg.out_ch, ch_mult=cfg.ch_mult, num_res_blocks=cfg.num_res_blocks,
                  attn_resolutions=cfg.attn_resolutions, z_channels=cfg.z_channels, resolution=cfg.resolution).eval()
    pq = nn.Conv2d(cfg.z_channels, cfg.z_channels, 1)    # AutoencoderKL.post_quant_conv, autoencoder.py:469
    dec.load_state_dict({k[len("decoder."):]: v for k, v in sd.items() if k.startswith("decoder.")})
    pq.load_state_dict({"weight": sd["post_quant_conv.weight"], "bias": sd["post_quant_conv.bias"]})
    with torch.no_grad():
        out = dec(pq(z))                                 # AutoencoderKL.decode, autoencoder.py:490-493
        img = torch.clamp(127.5 * out + 128.0, 0, 255).permute(0, 2, 3, 1).to("cpu", dtype=torch.uint8)  # vavae.py:80-81
    return out, img


def main():
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    for name, (cfg, seed, n) in CASES.items():
        sd = D.init_state_dict(cfg, seed=seed)
        z = torch.randn(n, cfg.z_channels, cfg.z_res, cfg.z_res, generator=torch.Generator().manual_seed(seed + 100))
        out, img = reference_decode(cfg, sd, z)
        path = os.path.join(ROOT, "tests", "golden", name + ".pt")
        digest = float(sum(v.double().abs().sum() for v in sd.values()))
        blob = {"config": cfg.__dict__, "weight_seed": seed, "weight_digest": digest, "z": z, "out": out, "img": img}
        if name == "decoder_tiny":          # the larger case regenerates its weights from the seed (digest-checked)
            blob["state_dict"] = sd
        torch.save(blob, path)
        print(name, tuple(out.shape), "std", float(out.std()), os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()

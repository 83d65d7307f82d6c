"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the VA-VAE (LDM AutoencoderKL) decode path.

Follows /root/reference/LightningDiT/tokenizer/autoencoder.py:
  nonlinearity (swish) :10-12 ; Normalize = GroupNorm(32, eps 1e-6, affine) :15-18 ; Upsample (nearest x2 + 3x3
  conv) :21-34 ; ResnetBlock :57-114 (temb is None in the decoder) ; AttnBlock :117-160 (single head, scale C^-1/2) ;
  Decoder :277-399 ; AutoencoderKL.decode = post_quant_conv (1x1) -> decoder :490-493 ;
and the caller-side post-processing of tokenizer/vavae.py:72-82 (clamp(127.5 x + 128, 0, 255) -> uint8 NHWC) and
inference.py:220 (latent de-normalisation, restated in va_vae_b200/checkpoint.py).

Pinned against the unmodified reference module by tests/test_decoder_oracle.py (container only: imports
tokenizer/autoencoder.py from /root/reference) and through the golden fixtures tests/golden/decoder_*.pt written by
oracle/gen_golden_decoder.py FROM THE REFERENCE MODULE.  Only tests/, __graft_entry__.smoke() and bench.py's CPU
baseline leg may import this file; the product path never does.
"""
from dataclasses import dataclass, field
from typing import Tuple

import torch
import torch.nn.functional as F


@dataclass
class DecoderConfig:
    z_channels: int = 32            # embed_dim (vavae_f16d32.yaml:8,28)
    ch: int = 128
    ch_mult: Tuple[int, ...] = (1, 1, 2, 2, 4)
    num_res_blocks: int = 2
    attn_resolutions: Tuple[int, ...] = (16,)
    resolution: int = 256
    out_ch: int = 3

    @property
    def num_resolutions(self):
        return len(self.ch_mult)

    @property
    def z_res(self):
        return self.resolution // 2 ** (self.num_resolutions - 1)


def state_dict_spec(cfg: DecoderConfig):
    """(name, shape) of every tensor AutoencoderKL.decode reads, in the reference's naming (autoencoder.py:277-399,469)."""
    spec = [("post_quant_conv.weight", (cfg.z_channels, cfg.z_channels, 1, 1)), ("post_quant_conv.bias", (cfg.z_channels,))]
    block_in = cfg.ch * cfg.ch_mult[-1]
    curr = cfg.z_res

    def conv(name, cin, cout, k):
        spec.append((name + ".weight", (cout, cin, k, k)))
        spec.append((name + ".bias", (cout,)))

    def norm(name, c):
        spec.append((name + ".weight", (c,)))
        spec.append((name + ".bias", (c,)))

    def resblock(name, cin, cout):
        norm(name + ".norm1", cin); conv(name + ".conv1", cin, cout, 3)
        norm(name + ".norm2", cout); conv(name + ".conv2", cout, cout, 3)
        if cin != cout:
            conv(name + ".nin_shortcut", cin, cout, 1)

    def attn(name, c):
        norm(name + ".norm", c)
        for p in ("q", "k", "v", "proj_out"):
            conv(name + "." + p, c, c, 1)

    conv("decoder.conv_in", cfg.z_channels, block_in, 3)
    resblock("decoder.mid.block_1", block_in, block_in)
    attn("decoder.mid.attn_1", block_in)
    resblock("decoder.mid.block_2", block_in, block_in)
    for lvl in reversed(range(cfg.num_resolutions)):
        block_out = cfg.ch * cfg.ch_mult[lvl]
        for ib in range(cfg.num_res_blocks + 1):
            resblock(f"decoder.up.{lvl}.block.{ib}", block_in, block_out)
            block_in = block_out
            if curr in cfg.attn_resolutions:
                attn(f"decoder.up.{lvl}.attn.{ib}", block_in)
        if lvl != 0:
            conv(f"decoder.up.{lvl}.upsample.conv", block_in, block_in, 3)
            curr *= 2
    norm("decoder.norm_out", block_in)
    conv("decoder.conv_out", block_in, cfg.out_ch, 3)
    return spec


def init_state_dict(cfg: DecoderConfig, seed: int = 0) -> dict:
    """Random weights with activations of order one at every depth (fan-in scaled convs, norm weights near one)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in state_dict_spec(cfg):
        if name.endswith(".weight") and len(shape) == 4:
            fan_in = shape[1] * shape[2] * shape[3]
            sd[name] = torch.randn(shape, generator=g) * (1.0 / fan_in) ** 0.5
        elif name.endswith(".weight"):
            sd[name] = 1.0 + 0.1 * torch.randn(shape, generator=g)
        else:
            sd[name] = 0.1 * torch.randn(shape, generator=g)
    return sd


def swish(x):
    return x * torch.sigmoid(x)                                   # autoencoder.py:10-12


def group_norm(sd, name, x):
    return F.group_norm(x, 32, sd[name + ".weight"], sd[name + ".bias"], eps=1e-6)   # :15-18


def conv(sd, name, x, pad):
    return F.conv2d(x, sd[name + ".weight"], sd[name + ".bias"], stride=1, padding=pad)


def resnet_block(sd, name, x):
    """autoencoder.py:93-114 with temb None, dropout 0."""
    h = conv(sd, name + ".conv1", swish(group_norm(sd, name + ".norm1", x)), 1)
    h = conv(sd, name + ".conv2", swish(group_norm(sd, name + ".norm2", h)), 1)
    if name + ".nin_shortcut.weight" in sd:
        x = conv(sd, name + ".nin_shortcut", x, 0)
    return x + h


def attn_block(sd, name, x):
    """autoencoder.py:135-160: single-head attention over the h*w positions, scale c^-1/2."""
    h_ = group_norm(sd, name + ".norm", x)
    q, k, v = conv(sd, name + ".q", h_, 0), conv(sd, name + ".k", h_, 0), conv(sd, name + ".v", h_, 0)
    b, c, h, w = q.shape
    q = q.reshape(b, c, h * w).permute(0, 2, 1)
    k = k.reshape(b, c, h * w)
    w_ = torch.softmax(torch.bmm(q, k) * (int(c) ** (-0.5)), dim=2)
    h_ = torch.bmm(v.reshape(b, c, h * w), w_.permute(0, 2, 1)).reshape(b, c, h, w)
    return x + conv(sd, name + ".proj_out", h_, 0)


@torch.no_grad()
def decode(cfg: DecoderConfig, sd: dict, z: torch.Tensor) -> torch.Tensor:
    """AutoencoderKL.decode (autoencoder.py:490-493): (n, z_channels, r, r) -> (n, out_ch, R, R) float."""
    h = conv(sd, "post_quant_conv", z, 0)
    h = conv(sd, "decoder.conv_in", h, 1)                          # :380
    h = resnet_block(sd, "decoder.mid.block_1", h)
    h = attn_block(sd, "decoder.mid.attn_1", h)
    h = resnet_block(sd, "decoder.mid.block_2", h)
    curr = cfg.z_res
    for lvl in reversed(range(cfg.num_resolutions)):               # :388-394
        for ib in range(cfg.num_res_blocks + 1):
            h = resnet_block(sd, f"decoder.up.{lvl}.block.{ib}", h)
            if curr in cfg.attn_resolutions:
                h = attn_block(sd, f"decoder.up.{lvl}.attn.{ib}", h)
        if lvl != 0:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")  # :31
            h = conv(sd, f"decoder.up.{lvl}.upsample.conv", h, 1)
            curr *= 2
    h = swish(group_norm(sd, "decoder.norm_out", h))                # :396-398
    return conv(sd, "decoder.conv_out", h, 1)


def to_images(dec: torch.Tensor) -> torch.Tensor:
    """tokenizer/vavae.py:80-81: clamp(127.5 x + 128, 0, 255) -> uint8 (truncation), NHWC."""
    return torch.clamp(127.5 * dec + 128.0, 0, 255).permute(0, 2, 3, 1).to(torch.uint8)


def flops_per_image(cfg: DecoderConfig) -> float:
    """2*M*N*K of every convolution and attention matmul of one decode."""
    block_in = cfg.ch * cfg.ch_mult[-1]
    curr = cfg.z_res
    f = 2.0 * curr * curr * cfg.z_channels * cfg.z_channels + 2.0 * curr * curr * 9 * cfg.z_channels * block_in

    def res(cin, cout, r):
        v = 2.0 * r * r * 9 * (cin * cout + cout * cout)
        return v + (2.0 * r * r * cin * cout if cin != cout else 0.0)

    def att(c, r):
        t = r * r
        return 4 * 2.0 * t * c * c + 2 * 2.0 * t * t * c

    f += res(block_in, block_in, curr) * 2 + att(block_in, curr)
    for lvl in reversed(range(cfg.num_resolutions)):
        block_out = cfg.ch * cfg.ch_mult[lvl]
        for ib in range(cfg.num_res_blocks + 1):
            f += res(block_in, block_out, curr)
            block_in = block_out
            if curr in cfg.attn_resolutions:
                f += att(block_in, curr)
        if lvl != 0:
            curr *= 2
            f += 2.0 * curr * curr * 9 * block_in * block_in
    f += 2.0 * curr * curr * 9 * block_in * cfg.out_ch
    return f

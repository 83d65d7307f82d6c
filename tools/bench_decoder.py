#!/usr/bin/env python
"""Decode throughput of the native VA-VAE decoder (production geometry: 16x16x32 latent -> 256x256x3) and its error
against the CPU oracle on one image.  Usage: python tools/bench_decoder.py [batch]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import va_vae_b200 as V
from va_vae_b200.vavae import DecoderEngine
from oracle import vavae_decoder_oracle as D

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 32
cfg = D.DecoderConfig()
sd = D.init_state_dict(cfg, seed=3)
conf = dict(z_channels=32, ch=128, ch_mult=(1, 1, 2, 2, 4), num_res_blocks=2, attn_resolutions=(16,), resolution=256)
dev = torch.device("cuda:0")
eng = DecoderEngine(conf, sd, dev, "bf16", max_batch=min(batch, 16))
z = torch.randn(batch, 32, 16, 16, generator=torch.Generator().manual_seed(4)).to(dev)
ref = D.decode(cfg, sd, z[:1].cpu())
out, img = eng.decode(z[:1], want_uint8=True)
rel = ((out.cpu().double() - ref.double()).norm() / ref.double().norm()).item()
d8 = (img.cpu().int() - D.to_images(ref).int()).abs().float()
print(f"one image vs fp32 CPU oracle: rel-L2 {rel:.3e}; uint8 mean |diff| {d8.mean():.3f}, max {int(d8.max())}, "
      f"share of pixels differing by more than 2: {(d8 > 2).float().mean():.2e}")
for _ in range(2):
    eng.decode(z, want_float=False, want_uint8=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
l0 = eng.launch_count()
e0.record()
reps = 3
for _ in range(reps):
    eng.decode(z, want_float=False, want_uint8=True)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
fl = D.flops_per_image(cfg) * batch
print(f"batch {batch}: {ms:.1f} ms per batch = {batch / ms * 1e3:.1f} images/s, {fl / ms / 1e9:.1f} TFLOP/s "
      f"({D.flops_per_image(cfg) / 1e9:.0f} GFLOP per image), {(eng.launch_count() - l0) // reps} kernel launches per batch, "
      f"workspace {eng.device_bytes() / 2**30:.2f} GiB")
# stock PyTorch eager on the same GPU: the oracle's op sequence (= the reference's) under autocast bf16
sd_dev = {k: v.to(dev) for k, v in sd.items()}
with torch.autocast("cuda", dtype=torch.bfloat16):
    for _ in range(2):
        D.decode(cfg, sd_dev, z)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        D.decode(cfg, sd_dev, z)
    e1.record()
torch.cuda.synchronize()
ms_t = e0.elapsed_time(e1) / reps
print(f"stock PyTorch eager (autocast bf16, cuDNN): {ms_t:.1f} ms per batch = {batch / ms_t * 1e3:.1f} images/s "
      f"-> native decoder {ms_t / ms:.2f}x")

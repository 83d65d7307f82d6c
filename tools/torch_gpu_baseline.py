#!/usr/bin/env python
"""Stock-PyTorch GPU baseline on the same B200 (informational; not the contract's reference arm, which is CPU).

Runs the restated LightningDiT forward_with_cfg (oracle/lightningdit_oracle.py, i.e. the reference's op sequence) on
cuda:0 with library kernels only -- cuBLAS GEMMs through F.linear under torch.autocast(bf16) and the flash / cuDNN
F.scaled_dot_product_attention (the reference's fused_attn branch, lightningdit.py:77-80) -- times a few evaluations
with CUDA events and extrapolates to the 250-point Heun sampler exactly like bench.py's cpu_baseline leg.  Then times
the engine on the same inputs.  Usage: python tools/torch_gpu_baseline.py [b1|xl1]
"""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F

import va_vae_b200 as V
from oracle import lightningdit_oracle as O


def attention_sdpa(cfg, sd, pre, x):
    B, T, D = x.shape
    H, hd = cfg.num_heads, cfg.head_dim
    qkv = F.linear(x, sd[pre + "qkv.weight"], sd[pre + "qkv.bias"]).reshape(B, T, 3, H, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    cos, sin = sd["feat_rope.freqs_cos"], sd["feat_rope.freqs_sin"]
    q = q * cos + O.rotate_half(q) * sin
    k = k * cos + O.rotate_half(k) * sin
    o = F.scaled_dot_product_attention(q.to(v.dtype), k.to(v.dtype), v)
    return F.linear(o.transpose(1, 2).reshape(B, T, D), sd[pre + "proj.weight"], sd[pre + "proj.bias"])


def timestep_embedding_dev(t, dim=256, max_period=10000):
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "b1"
    name, n, shift = ("LightningDiT-B/1", 256, 0.1) if which == "b1" else ("LightningDiT-XL/1", 128, 0.3)
    dev = torch.device("cuda:0")
    cfg = O.config_from_name(name, input_size=16, num_classes=1000)
    sd = O.init_state_dict(cfg, seed=0)
    sd_dev = {k: v.to(dev) for k, v in sd.items()}
    O.attention = attention_sdpa
    O.timestep_embedding = timestep_embedding_dev
    z = torch.randn(n, 32, 16, 16, generator=torch.Generator().manual_seed(1234)).to(dev)
    y = torch.randint(0, 1000, (n,), generator=torch.Generator().manual_seed(4321)).to(dev)
    x2, y2 = torch.cat([z, z]), torch.cat([y, torch.full_like(y, 1000)])
    evals = 498
    t = torch.full((2 * n,), 0.37, device=dev)

    def torch_eval():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            return O.forward_with_cfg(cfg, sd_dev, x2, t, y2, 10.0)

    for _ in range(3):
        ref = torch_eval()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        torch_eval()
    e1.record()
    torch.cuda.synchronize()
    ms_torch = e0.elapsed_time(e1) / reps

    m = V.LightningDiT_models[name](input_size=16, num_classes=1000, use_qknorm=False, use_swiglu=True, use_rope=True,
                                   use_rmsnorm=True, wo_shift=False)
    m.load_state_dict(sd)
    m = m.eval().to(dev)
    for _ in range(3):
        out = m.forward_with_cfg(x2, t, y2, 10.0)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        m.forward_with_cfg(x2, t, y2, 10.0)
    e1.record()
    torch.cuda.synchronize()
    ms_eng = e0.elapsed_time(e1) / reps
    rel = ((out.float() - ref.float()).norm() / ref.float().norm()).item()
    fl = 2 * n * O.flops_per_forward(cfg)
    print(f"{name} n={n} (2n={2 * n} model rows), one forward_with_cfg evaluation:")
    print(f"  stock PyTorch eager, autocast bf16 + SDPA : {ms_torch:8.2f} ms  {fl / ms_torch / 1e9:7.1f} TFLOP/s  "
          f"-> {n / (ms_torch * 1e-3 * evals):6.2f} samples/s extrapolated to {evals} evaluations")
    print(f"  this engine (bf16 mode, C ABI)            : {ms_eng:8.2f} ms  {fl / ms_eng / 1e9:7.1f} TFLOP/s  "
          f"-> {n / (ms_eng * 1e-3 * evals):6.2f} samples/s")
    print(f"  speed-up {ms_torch / ms_eng:.2f}x ; engine vs autocast-torch output rel-L2 {rel:.2e} "
          f"(both bf16; the parity tests compare against the fp32 oracle instead)")


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Isolated throughput of the four per-layer GEMMs and attention (B/1 and XL/1 shapes, 512 / 256 model rows)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import va_vae_b200 as V
from oracle import lightningdit_oracle as O
CASES = (("LightningDiT-B/1", 512), ("LightningDiT-XL/1", 256))
if os.environ.get("AB_MODELS") == "xl":
    CASES = CASES[1:]
for name, rows in CASES:
    cfg = O.config_from_name(name, input_size=16, num_classes=1000)
    cfg.depth = 1
    sd = O.init_state_dict(cfg, seed=0)
    m = V.LightningDiT(input_size=16, patch_size=1, in_channels=32, hidden_size=cfg.hidden_size, depth=1,
                       num_heads=cfg.num_heads, num_classes=1000, use_swiglu=True, use_rope=True, use_rmsnorm=True)
    m.load_state_dict(sd); m = m.eval().to("cuda:0"); eng = m.engine()
    for k in ("gemm_qkv", "gemm_proj", "gemm_w12", "gemm_w3", "attention"):
        r = []
        for _ in range(3):
            ms, work = eng.time_kernel(k, rows, 20); r.append((round(work / ms / 1e9), round(ms * 1e3 / 20)))
        print(name, k, "TFLOP/s, us:", r, flush=True)
    del m, eng

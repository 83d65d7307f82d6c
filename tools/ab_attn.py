#!/usr/bin/env python
"""Isolated attention throughput (B/1, XL/1 at 256 tokens; XL/1 at 1024 tokens) for library A/B runs."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import va_vae_b200 as V
from oracle import lightningdit_oracle as O
for name, size, rows in (("LightningDiT-B/1", 16, 512), ("LightningDiT-XL/1", 16, 256), ("LightningDiT-XL/1", 32, 64)):
    cfg = O.config_from_name(name, input_size=size, num_classes=1000)
    cfg.depth = 1
    sd = O.init_state_dict(cfg, seed=0)
    m = V.LightningDiT(input_size=size, patch_size=1, in_channels=32, hidden_size=cfg.hidden_size, depth=1,
                       num_heads=cfg.num_heads, num_classes=1000, use_swiglu=True, use_rope=True, use_rmsnorm=True)
    m.load_state_dict(sd); m = m.eval().to("cuda:0"); eng = m.engine()
    r = []
    for _ in range(4):
        ms, work = eng.time_kernel("attention", rows, 20); r.append(round(work / ms / 1e9))
    print(name, "T=%d" % (size * size), "attention TFLOP/s:", r, flush=True)
    del m, eng

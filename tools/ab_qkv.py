#!/usr/bin/env python
"""What does RoPE cost in the QKV epilogue?  Isolated QKV GEMM with and without RoPE tables (B/1 and XL/1 shapes)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import va_vae_b200 as V
from oracle import lightningdit_oracle as O
for name in ("LightningDiT-B/1", "LightningDiT-XL/1"):
    for rope in (True, False, True, False):
        cfg = O.config_from_name(name, input_size=16, num_classes=1000, use_rope=rope)
        cfg.depth = 1
        sd = O.init_state_dict(cfg, seed=0)
        m = V.LightningDiT(input_size=16, patch_size=1, in_channels=32, hidden_size=cfg.hidden_size, depth=1,
                           num_heads=cfg.num_heads, num_classes=1000, use_swiglu=True, use_rope=rope, use_rmsnorm=True)
        m.load_state_dict(sd); m = m.eval().to("cuda:0"); eng = m.engine()
        r = []
        for _ in range(3):
            ms, work = eng.time_kernel("gemm_qkv", 512, 20); r.append(round(work / ms / 1e9))
        ms2, w2 = eng.time_kernel("gemm_w12", 512, 20)
        print(name, "rope" if rope else "no-rope", "qkv TFLOP/s", r, "w12", round(w2 / ms2 / 1e9), flush=True)
        del m, eng

#!/usr/bin/env python
"""A/B of the single-CTA vs CTA-pair GEMM kernels on the engine's own layer-0 shapes (isolated, CUDA events)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import va_vae_b200 as V
from va_vae_b200 import engine as E
from oracle import lightningdit_oracle as O

for name in ("LightningDiT-B/1", "LightningDiT-XL/1"):
    cfg = O.config_from_name(name, input_size=16, num_classes=1000)
    cfg.depth = 2
    sd = O.init_state_dict(cfg, seed=0)
    m = V.LightningDiT(input_size=16, patch_size=1, in_channels=32, hidden_size=cfg.hidden_size, depth=2,
                       num_heads=cfg.num_heads, num_classes=1000, use_swiglu=True, use_rope=True, use_rmsnorm=True)
    m.load_state_dict(sd); m = m.eval().to("cuda:0"); eng = m.engine()
    res = {}
    for rep in range(3):
        for two in (0, 1):
            E.set_gemm_two_cta(bool(two))
            for k in ("gemm_qkv", "gemm_proj", "gemm_w12", "gemm_w3"):
                ms, work = eng.time_kernel(k, 512, 20)
                res.setdefault((k, two), []).append(work / ms / 1e9)
    for k in ("gemm_qkv", "gemm_proj", "gemm_w12", "gemm_w3"):
        print(name, k, "1cta", [round(x) for x in res[(k, 0)]], "2cta", [round(x) for x in res[(k, 1)]], flush=True)
    del m, eng

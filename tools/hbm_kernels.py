#!/usr/bin/env python
"""Achieved GB/s of the bandwidth-bound kernels at a DRAM-sized batch (512 model rows of XL/1: every tensor they touch is
larger than the 126 MB L2 except the latents), against the measured copy bandwidth in MEASURED_PEAKS.json.
Algorithmic bytes only (compulsory reads + writes).  python tools/hbm_kernels.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import va_vae_b200 as V
from oracle import lightningdit_oracle as O

peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
for name, rows in (("LightningDiT-XL/1", 512), ("LightningDiT-B/1", 512)):
    cfg = O.config_from_name(name, input_size=16, num_classes=1000)
    cfg.depth = 1
    sd = O.init_state_dict(cfg, seed=0)
    m = V.LightningDiT(input_size=16, patch_size=1, in_channels=32, hidden_size=cfg.hidden_size, depth=1,
                       num_heads=cfg.num_heads, num_classes=1000, use_swiglu=True, use_rope=True, use_rmsnorm=True)
    m.load_state_dict(sd)
    eng = m.eval().to("cuda:0").engine()
    for k in ("norm_modulate", "patchify", "ode_update", "cfg_combine"):
        best = None
        for _ in range(3):
            ms, work = eng.time_kernel(k, rows, 50)
            gbs = work / ms / 1e6
            best = max(best or 0.0, gbs)
        print(json.dumps(dict(model=name, rows=rows, kernel=k, bytes=work, us=round(work / best / 1e3, 2), gbs=round(best, 1),
                              frac_of_measured_copy_bw=round(best / peak, 3))), flush=True)
    del eng, m

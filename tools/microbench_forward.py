#!/usr/bin/env python
"""BASELINE.json configs[4]: LightningDiT-XL/1 single `forward` (no CFG, no sampler) latency sweep over the batch size,
CUDA events on the launching stream, plus isolated per-kernel timings (tensor TFLOP/s for GEMM/attention, GB/s for the
bandwidth kernels).  Prints one JSON line per batch size.   python tools/microbench_forward.py [--model LightningDiT-XL/1]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("TORCH_COMPILE_DISABLE", "1")

import torch

import va_vae_b200 as V
from oracle import lightningdit_oracle as O   # weight recipe + FLOP count only


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="LightningDiT-XL/1")
    ap.add_argument("--batches", default="1,2,4,8,16,32,64,128,256,512,1024,2048")
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    cfg = O.config_from_name(args.model, input_size=16, num_classes=1000)
    sd = O.init_state_dict(cfg, seed=0)
    m = V.LightningDiT_models[args.model](input_size=16, num_classes=1000, in_channels=32, use_swiglu=True, use_rope=True,
                                          use_rmsnorm=True)
    m.load_state_dict(sd)
    m = m.eval().to(dev)
    eng = m.engine()
    flops = O.flops_per_forward(cfg)
    for b in [int(x) for x in args.batches.split(",")]:
        x = torch.randn(b, 32, 16, 16, device=dev)
        t = torch.rand(b, device=dev)
        y = torch.randint(0, 1000, (b,), device=dev)
        for _ in range(3):
            m(x, t, y)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(args.iters):
            m(x, t, y)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.iters
        line = dict(model=args.model, batch=b, ms_per_forward=ms, samples_per_s=b / ms * 1e3,
                    tflops=flops * b / ms / 1e9, bound="latency / weight streaming" if b <= 4 else "tensor")
        if b in (1, 64, 512):
            rows = min(b, 512)
            bd = {}
            for k in ("gemm_qkv", "attention", "gemm_proj", "gemm_w12", "gemm_w3", "norm_modulate"):
                kms, work = eng.time_kernel(k, rows, 20)
                byt = k == "norm_modulate"
                bd[k] = dict(ms=kms, achieved=work / kms / (1e6 if byt else 1e9), unit="GB/s" if byt else "TFLOP/s")
            line["kernels"] = bd
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()

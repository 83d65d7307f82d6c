#!/usr/bin/env python
"""Benchmark of the hot path: 250-point Heun + CFG latent sampling with LightningDiT on B200.

    python bench.py --gpus N --steps K --warmup W            # this engine (default arm)
    python bench.py --impl reference --gpus N ...            # the reference's CPU path (oracle port) on host cores

One "step" = one full pass of the sampler over one batch of synthetic latents (n samples per GPU:
P-1 Heun intervals = 2(P-1) evaluations of forward_with_cfg on 2n model rows).  Default workload is
BASELINE.json configs[1]: LightningDiT-B/1, 16x16x32 latent, n=256 per GPU, 250-point Heun + CFG.
Rank 0 prints ONE JSON line.  Timing: CUDA events on the launching stream, barrier + synchronize on both
sides, max over ranks; L2 is flushed (256 MiB write) between timed steps.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

os.environ.setdefault("TORCH_COMPILE_DISABLE", "1")
ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch

WORKLOADS = {
    # name: (registry name, num_classes, cfg_scale, timestep_shift, default n per GPU)
    "b1": ("LightningDiT-B/1", 31, 10.0, 0.1, 256),     # BASELINE.json configs[1] (config_dit_base.yaml)
    "xl1": ("LightningDiT-XL/1", 1000, 10.0, 0.3, 128),   # BASELINE.json configs[2]: 1024 global = 128 per GPU at 8 GPUs
    "s1": ("LightningDiT-S/1", 31, 10.0, 0.1, 4),       # BASELINE.json configs[0]
    # BASELINE.json configs[3]: XL/1 at 512x512 = 32x32x32 latent (1024 tokens), 256 global = 32 per GPU at 8 GPUs;
    # run with --method euler --cfg-interval-start 0.11 for the config as named
    "xl1_512": ("LightningDiT-XL/1", 1000, 10.0, 0.3, 32),
}
INPUT_SIZE = {"xl1_512": 32}
METRIC = "latent samples/sec (250-step CFG)"
UNIT = "samples/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--workload", default="b1", choices=sorted(WORKLOADS))
    ap.add_argument("--n", type=int, default=0, help="samples per GPU (default: workload's)")
    ap.add_argument("--points", type=int, default=250)
    ap.add_argument("--method", default="heun", choices=["euler", "heun"])
    ap.add_argument("--chunk", type=int, default=0, help="engine max_chunk_rows (0 = default)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--breakdown", action="store_true", help="add isolated per-kernel timings to the JSON line")
    ap.add_argument("--profile-kernel", default="gemm_w12",
                    choices=["gemm_qkv", "attention", "gemm_proj", "gemm_w12", "gemm_w3", "norm_modulate", "ode_update"],
                    help="kernel class bracketed by CUDA events inside the timed region (roofline object)")
    ap.add_argument("--cfg-interval-start", type=float, default=None,
                    help="enable the CFG interval gate (guidance only for t >= start), e.g. 0.11 for configs[3]")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(tflops_sustained=p["bf16_tflops_sustained"], tflops_burst=p["bf16_tflops"], hbm=p["hbm_gbs"],
                    source="measured (MEASURED_PEAKS.json)")
    return dict(tflops_sustained=1400.0, tflops_burst=1590.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


def input_size(workload):
    return INPUT_SIZE.get(workload, 16)


def synth_inputs(cfg, n_total, K):
    """noise seed 1234, labels seed 4321, generated once on CPU for the GLOBAL batch."""
    z = torch.randn(n_total, cfg.in_channels, cfg.input_size, cfg.input_size, generator=torch.Generator().manual_seed(1234))
    y = torch.randint(0, K, (n_total,), generator=torch.Generator().manual_seed(4321))
    return z, y


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.QUERY}",
                                       "--format=csv,noheader,nounits", "-lms", "200"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
                power.append(float(parts[3]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        load = sorted(s for s, p in zip(sm, power) if p >= 0.5 * max(power)) or sorted(sm)
        return {"sm_mhz": load[len(load) // 2], "sm_max_mhz": max(mx), "power_w_max": max(power), "samples": len(sm),
                "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ CPU reference arm
def cpu_reference_rate(workload, points, method, n=4, evals=3, warm=1):
    """The reference's CPU path (oracle port; /root/reference does not exist on the GPU box), all host threads:
    time `evals` forward_with_cfg evaluations on n samples (2n rows) and extrapolate to the full sampler."""
    from oracle import lightningdit_oracle as O
    name, K, scale, shift, _ = WORKLOADS[workload]
    cfg = O.config_from_name(name, input_size=input_size(workload), num_classes=K)
    sd = O.init_state_dict(cfg, seed=0)
    z, y = synth_inputs(cfg, n, K)
    zz, yy = torch.cat([z, z], 0), torch.cat([y, torch.full((n,), K)])
    t = torch.full((2 * n,), 0.5)
    for _ in range(warm):
        O.forward_with_cfg(cfg, sd, zz, t, yy, scale)
    t0 = time.perf_counter()
    for _ in range(evals):
        O.forward_with_cfg(cfg, sd, zz, t, yy, scale)
    per_eval = (time.perf_counter() - t0) / evals
    n_evals = (points - 1) * (2 if method == "heun" else 1)
    return dict(value=n / (per_eval * n_evals), unit=UNIT, cores=torch.get_num_threads(), kind="port",
                sample=f"{evals} forward_with_cfg evals of {name} on n={n} (2n={2 * n} rows), fp32 eager torch CPU, "
                       f"extrapolated to {n_evals} evals; os.cpu_count()={os.cpu_count()}"), per_eval


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name, K, scale, shift, n_def = WORKLOADS[args.workload]
    for _ in range(max(args.warmup - 1, 0)):
        cpu_reference_rate(args.workload, args.points, args.method, n=2, evals=1, warm=0)
    vals, t0 = [], time.perf_counter()
    for _ in range(args.steps):
        cb, _ = cpu_reference_rate(args.workload, args.points, args.method, n=4, evals=2, warm=0)
        vals.append(cb["value"])
    elapsed = time.perf_counter() - t0
    value = sum(vals) / len(vals)
    cb["value"] = value
    n = args.n or n_def
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1e3 * elapsed / max(args.steps, 1), higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f32", data="synthetic", impl="reference",
                config=dict(workload=f"{name} 16x16x32 latent, n={n} per GPU, {args.points}-point {args.method}+CFG "
                                     f"(cfg {scale}, shift {shift})", bounded_sample=cb["sample"]),
                cpu_baseline=cb, e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0)
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ engine arm
def run_engine(args):
    import torch.distributed as dist
    import va_vae_b200 as V
    from va_vae_b200 import dist as vdist
    from oracle import lightningdit_oracle as O   # weight recipe + FLOP count only (not on the timed path)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl engine needs a B200: the engine has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    name, K, scale, shift, n_def = WORKLOADS[args.workload]
    n = args.n or n_def
    n_total = n * world
    isz = input_size(args.workload)
    cfg = O.config_from_name(name, input_size=isz, num_classes=K)
    sd = O.init_state_dict(cfg, seed=0)
    model = V.LightningDiT_models[name](input_size=isz, num_classes=K, in_channels=32, use_swiglu=True, use_rope=True,
                                        use_rmsnorm=True, precision=args.precision, max_chunk_rows=args.chunk)
    model.load_state_dict(sd)
    model = model.eval().to(dev)
    eng = model.engine()
    del sd
    grid = O.time_grid(args.points, shift)
    z_all, y_all = synth_inputs(cfg, n_total, K)
    lo, hi = vdist.shard_range(n_total, rank, world)
    zz, yy = vdist.build_cfg_batch(z_all[lo:hi], y_all[lo:hi], K)      # inference.py:205-209
    z_dev, y_dev = zz.to(dev), yy.to(dev)
    z_pin, y_pin = zz.pin_memory(), yy.pin_memory()
    out_pin = torch.empty_like(zz).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    n_evals = (args.points - 1) * (2 if args.method == "heun" else 1)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    gate = dict(cfg_interval=True, cfg_interval_start=args.cfg_interval_start) if args.cfg_interval_start is not None else {}

    def step_resident():
        flush.zero_()
        out = eng.sample_ode(z_dev, y_dev, grid, args.method, True, scale, **gate)
        samples = out[: hi - lo]
        if world > 1:
            gathered = torch.empty((n_total,) + tuple(samples.shape[1:]), device=dev)
            dist.all_gather_into_tensor(gathered, samples.contiguous())
            return gathered
        return samples

    def step_e2e():
        flush.zero_()
        eng.sample_ode_host(z_pin, y_pin, grid, out_pin, args.method, True, scale, **gate)
        return out_pin

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(args.warmup):
        step_resident()
    clocks = ClockSampler(local) if rank == 0 else None
    l0 = eng.launch_count
    eng.profile_begin(args.profile_kernel)
    ms = timed(step_resident, args.steps)
    prof_ms, prof_n, prof_work = eng.profile_end()
    launches = torch.tensor([eng.launch_count - l0], device=dev, dtype=torch.int64)
    if world > 1:
        dist.all_reduce(launches)
    clk = clocks.stop() if clocks else None
    value = n_total * args.steps / (ms / 1e3)

    k_e2e = min(args.steps, 2)
    step_e2e()
    ms_e2e = timed(step_e2e, k_e2e)
    e2e_value = n_total * k_e2e / (ms_e2e / 1e3)

    pk = peaks()
    flops_fwd = O.flops_per_forward(cfg)
    achieved_total = value * flops_fwd * 2 * n_evals / 1e12 / world      # per-GPU TFLOP/s over the whole step
    hbm_kernel = args.profile_kernel in ("norm_modulate", "ode_update")
    roof_achieved = prof_work / (prof_ms / 1e3) / (1e9 if hbm_kernel else 1e12) if prof_ms > 0 else None
    roof_peak = pk["hbm"] if hbm_kernel else pk["tflops_sustained"]
    kernel_names = {"gemm_w12": "k_gemm_tc<SWIGLU> (mlp.w12, 42.8% of FLOPs)", "gemm_qkv": "k_gemm_tc<QKV> (attn.qkv + RoPE + scatter)",
                    "gemm_proj": "k_gemm_tc<GATED_RESIDUAL> (attn.proj)", "gemm_w3": "k_gemm_tc<GATED_RESIDUAL> (mlp.w3)",
                    "attention": "k_attention_tc", "norm_modulate": "k_norm_modulate", "ode_update": "k_ode_update"}
    line = dict(
        metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
        ms_per_step=ms / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
        dtype="bf16" if args.precision == "bf16" else "f32", data="synthetic",
        config=dict(workload=f"{name} {isz}x{isz}x32 latent, n={n} per GPU ({2 * n} model rows), {args.points}-point "
                             f"{args.method}+CFG = {n_evals} evals (cfg {scale}, shift {shift}, 3 guided channels"
                             + (f", interval start {args.cfg_interval_start}: every forward executed, none skipped)"
                                if args.cfg_interval_start is not None else ")"),
                    parallelism=f"dp{world} (samples sharded, CFG pair per GPU, one all-gather of final latents)",
                    weights="random init (oracle.init_state_dict seed 0)", l2="flushed between steps (256 MiB write)",
                    chunk_rows=args.chunk or 512, tflops_per_gpu_whole_step=achieved_total,
                    frac_of_sustained_peak_whole_step=achieved_total / pk["tflops_sustained"]),
        clocks=clk,
        e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=int(zz.numel() * 4 + yy.numel() * 8) * world,
                 d2h_bytes_per_step=int(zz.numel() * 4) * world, steps=k_e2e,
                 api="Engine.sample_ode_host -> vavae_sample_ode_host (pinned host buffers)"),
        gpu_launches=int(launches.item()),
        roofline=dict(bound="hbm" if hbm_kernel else "tensor", kernel=kernel_names[args.profile_kernel], achieved=roof_achieved,
                      peak=roof_peak, unit="GB/s" if hbm_kernel else "TFLOP/s",
                      frac=(roof_achieved / roof_peak) if roof_achieved else None,
                      # dram__bytes_read+write per launch from profiles/r1_w12_ncu.md (B/1, 512-row chunks) vs
                      # 744.5e6 algorithmic bytes; only valid for that exact launch shape
                      traffic=(728.1e6 if (args.workload == "b1" and 2 * n >= 512 and not args.chunk
                                           and args.profile_kernel == "gemm_w12") else None),
                      launches_timed=prof_n, peak_source=pk["source"] + ", sustained figure (kernel timed inside a long step)"),
    )
    if args.breakdown and rank == 0:
        bd = {}
        rows = min(2 * n, eng_chunk(eng, args))
        for kname in ("gemm_qkv", "attention", "gemm_proj", "gemm_w12", "gemm_w3", "norm_modulate", "ode_update"):
            kms, work = eng.time_kernel(kname, rows, 20)
            is_bytes = kname in ("norm_modulate", "ode_update")
            bd[kname] = dict(ms=kms, achieved=work / (kms / 1e3) / (1e9 if is_bytes else 1e12),
                             unit="GB/s" if is_bytes else "TFLOP/s")
        line["kernel_breakdown"] = dict(rows=rows, kernels=bd)
    if rank == 0:
        if not args.no_cpu_baseline:
            cb, _ = cpu_reference_rate(args.workload, args.points, args.method, n=4, evals=3, warm=1)
            line["cpu_baseline"] = cb
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def eng_chunk(eng, args):
    return args.chunk or 512


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_engine(a)

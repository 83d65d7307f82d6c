#!/usr/bin/env python
"""Benchmark of the hot path: 250-point Heun + CFG latent sampling with LightningDiT on B200.

    python bench.py --gpus N --steps K --warmup W            # this engine (default arm)
    python bench.py --impl reference --gpus N ...            # the reference's CPU path (oracle port) on host cores

One "step" = one full pass of the sampler over one batch of synthetic latents (n samples per GPU:
P-1 Heun intervals = 2(P-1) evaluations of forward_with_cfg on 2n model rows).  The default workload is
the configuration BASELINE.json's metric is quoted on, configs[2]: LightningDiT-XL/1 (256 tokens, d=1152,
28 layers), batch 1024 over 8 GPUs = n=128 per GPU (weak scaling: per-GPU work fixed), 250-point Heun + CFG.
`--global-n 1024` instead shards a FIXED global batch over the ranks (strong scaling, 1024/512/256/128 per GPU).
Rank 0 prints ONE JSON line.  Timing: CUDA events on the launching stream, barrier + synchronize on both
sides, max over ranks; L2 is flushed (256 MiB write) between timed steps.

Warm-up: W untimed passes of the SAME kernels on the SAME batch, integrated over a short `--warmup-points` grid
(default 8 points = 14 evaluations instead of 498) so that `--steps 20 --warmup 5` on XL/1 fits the driver's
per-run budget (a full step is ~30 s); the first timed step is therefore preceded by >= 70 evaluations of every
kernel.  `--warmup-points 0` uses the full grid.
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
    "xl1": ("LightningDiT-XL/1", 1000, 10.0, 0.3, 128),   # BASELINE.json configs[2]: 1024 global = 128 per GPU at 8 GPUs
    "b1": ("LightningDiT-B/1", 31, 10.0, 0.1, 256),     # BASELINE.json configs[1] (config_dit_base.yaml)
    "s1": ("LightningDiT-S/1", 31, 10.0, 0.1, 4),       # BASELINE.json configs[0]
    # BASELINE.json configs[3]: XL/1 at 512x512 = 32x32x32 latent (1024 tokens), 256 global = 32 per GPU at 8 GPUs;
    # run with --method euler --cfg-interval-start 0.11 for the config as named
    "xl1_512": ("LightningDiT-XL/1", 1000, 10.0, 0.3, 32),
}
INPUT_SIZE = {"xl1_512": 32}
METRIC = "latent samples/sec (250-step CFG)"
UNIT = "samples/s"
KERNELS = ["gemm_qkv", "attention", "gemm_proj", "gemm_w12", "gemm_w3", "norm_modulate", "ode_update"]


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--workload", default="xl1", choices=sorted(WORKLOADS))
    ap.add_argument("--n", type=int, default=0, help="samples per GPU (default: workload's)")
    ap.add_argument("--global-n", type=int, default=0,
                    help="strong scaling: total samples, sharded over the ranks (overrides --n)")
    ap.add_argument("--points", type=int, default=250)
    ap.add_argument("--warmup-points", type=int, default=8, help="grid points of a warm-up pass (0 = --points)")
    ap.add_argument("--e2e-steps", type=int, default=1, help="timed end-to-end (host buffer) steps")
    ap.add_argument("--method", default="heun", choices=["euler", "heun"])
    ap.add_argument("--chunk", type=int, default=0, help="engine max_chunk_rows (0 = default)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--breakdown", action="store_true", help="add isolated per-kernel timings to the JSON line")
    ap.add_argument("--profile-kernel", default="gemm_w12", choices=KERNELS,
                    help="kernel class bracketed by CUDA events inside the timed region (roofline object)")
    ap.add_argument("--cfg-interval-start", type=float, default=None,
                    help="enable the CFG interval gate (guidance only for t >= start), e.g. 0.11 for configs[3]")
    ap.add_argument("--literal-2n", action="store_true",
                    help="with --cfg-interval-start: also run the unconditional forward whose output the gate discards "
                         "(the reference's behaviour, lightningdit.py:436-439); default skips it")
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


def per_gpu_n(args, world):
    """(samples per GPU, global samples, scaling) of this run."""
    if args.global_n:
        return -(-args.global_n // world), args.global_n, "strong"
    n = args.n or WORKLOADS[args.workload][4]
    return n, n * world, "weak"


def n_evals_of(args):
    return (args.points - 1) * (2 if args.method == "heun" else 1)


def workload_desc(args, world):
    """ONE description of the workload, printed verbatim by both arms (the driver compares them)."""
    name, K, scale, shift, _ = WORKLOADS[args.workload]
    isz = input_size(args.workload)
    n, n_total, scaling = per_gpu_n(args, world)
    s = (f"{name} {isz}x{isz}x32 latent, n={n} per GPU ({2 * n} model rows), {args.points}-point {args.method}+CFG = "
         f"{n_evals_of(args)} evals (cfg {scale}, shift {shift}, 3 guided channels")
    if args.cfg_interval_start is not None:
        s += f", interval start {args.cfg_interval_start}"
    s += ")"
    if scaling == "strong":
        s += f", global batch {n_total} sharded over the GPUs"
    return s


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
def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU legs are reported with all host cores."""
    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        pass
    torch.set_num_threads(max(cores, 1))
    return torch.get_num_threads()


def cpu_reference_rate(args, n=4, evals=3, warm=1, budget_s=12.0):
    """The reference's CPU path (oracle port; /root/reference does not exist on the GPU box), all host threads:
    time up to `evals` forward_with_cfg evaluations on n samples (2n rows; stops early once `budget_s` of CPU work is
    spent) and extrapolate to the full sampler."""
    from oracle import lightningdit_oracle as O
    threads = use_all_host_threads()
    name, K, scale, shift, _ = WORKLOADS[args.workload]
    cfg = O.config_from_name(name, input_size=input_size(args.workload), num_classes=K)
    sd = O.init_state_dict(cfg, seed=0)
    z, y = synth_inputs(cfg, n, K)
    zz, yy = torch.cat([z, z], 0), torch.cat([y, torch.full((n,), K)])
    t = torch.full((2 * n,), 0.5)
    for _ in range(warm):
        O.forward_with_cfg(cfg, sd, zz, t, yy, scale)
    done, t0 = 0, time.perf_counter()
    while done < evals and (done == 0 or time.perf_counter() - t0 < budget_s):
        O.forward_with_cfg(cfg, sd, zz, t, yy, scale)
        done += 1
    per_eval = (time.perf_counter() - t0) / done
    n_evals = n_evals_of(args)
    return dict(value=n / (per_eval * n_evals), unit=UNIT, cores=threads, kind="port",
                sample=f"{done} forward_with_cfg evals of {name} on n={n} (2n={2 * n} rows), fp32 eager torch CPU "
                       f"({threads} threads), extrapolated to {n_evals} evals; os.cpu_count()={os.cpu_count()}"), per_eval


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    from oracle import lightningdit_oracle as O
    threads = use_all_host_threads()
    name, K, scale, shift, _ = WORKLOADS[args.workload]
    cfg = O.config_from_name(name, input_size=input_size(args.workload), num_classes=K)
    sd = O.init_state_dict(cfg, seed=0)
    n = 4 if cfg.tokens <= 256 else 1
    z, y = synth_inputs(cfg, n, K)
    zz, yy = torch.cat([z, z], 0), torch.cat([y, torch.full((n,), K)])
    t = torch.full((2 * n,), 0.5)
    n_evals = n_evals_of(args)

    def step():   # one bounded sample of the workload: one forward_with_cfg evaluation of 2n rows
        t0 = time.perf_counter()
        O.forward_with_cfg(cfg, sd, zz, t, yy, scale)
        return time.perf_counter() - t0

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    per = [step() for _ in range(args.steps)]
    elapsed = time.perf_counter() - t0
    per_eval = sum(per) / len(per)
    value = n / (per_eval * n_evals)
    sample = (f"each step = 1 forward_with_cfg eval of {name} on n={n} (2n={2 * n} rows), fp32 eager torch CPU "
              f"({threads} threads), extrapolated to {n_evals} evals; os.cpu_count()={os.cpu_count()}")
    cb = dict(value=value, unit=UNIT, cores=threads, kind="port", sample=sample)
    _, _, scaling = per_gpu_n(args, world)
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1e3 * elapsed / max(args.steps, 1), higher_is_better=True, scaling=scaling, vs_baseline=None,
                dtype="f32", data="synthetic", impl="reference",
                config=dict(workload=workload_desc(args, world), bounded_sample=sample),
                cpu_baseline=cb, e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0)
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ engine arm
def measured_traffic(workload, rows, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed ncu --set full capture
    of this workload (profiles/ncu_traffic.json, written by tools/ncu_traffic.py); None when that shape was not captured."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(path):
        return None
    return json.load(open(path)).get(f"{workload}:{rows}:{kernel}")


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
    name, K, scale, shift, _ = WORKLOADS[args.workload]
    n, n_total, scaling = per_gpu_n(args, world)
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
    wpoints = args.warmup_points if 0 < args.warmup_points < args.points else args.points
    wgrid = O.time_grid(wpoints, shift)
    z_all, y_all = synth_inputs(cfg, n_total, K)
    lo, hi = vdist.shard_range(n_total, rank, world)
    z_dev, y_dev = z_all.to(dev), y_all.to(dev)                        # global batch resident on every rank (33 MB at 1024)
    zz, yy = vdist.build_cfg_batch(z_all[lo:hi], y_all[lo:hi], K)      # this rank's shard, inference.py:205-209
    z_pin, y_pin = zz.pin_memory(), yy.pin_memory()
    out_pin = torch.empty_like(zz).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    n_evals = n_evals_of(args)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    gate = {}
    if args.cfg_interval_start is not None:
        gate = dict(cfg_interval=True, cfg_interval_start=args.cfg_interval_start, skip_discarded=not args.literal_2n)

    def sampler(g):
        def sample_local(z, y):   # one rank's shard: CFG doubling (inference.py:205-209) + the native sampler
            zc, yc = vdist.build_cfg_batch(z, y, K)
            return eng.sample_ode(zc, yc, g, args.method, True, scale, **gate)[: z.shape[0]]
        return sample_local

    def step_resident(g=grid):
        flush.zero_()
        # the product's multi-GPU helper: shard, sample, ONE all-gather of the final latents (no-op at world 1)
        return vdist.sample_sharded(sampler(g), z_dev, y_dev, device=dev)

    def step_e2e(g=grid):
        flush.zero_()
        eng.sample_ode_host(z_pin, y_pin, g, out_pin, args.method, True, scale, **gate)
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
        step_resident(wgrid)
    clocks = ClockSampler(local) if rank == 0 else None
    l0, f0 = eng.launch_count, eng.forward_rows
    eng.profile_begin(args.profile_kernel)
    ms = timed(step_resident, args.steps)
    prof_ms, prof_n, prof_work = eng.profile_end()
    counts = torch.tensor([eng.launch_count - l0, eng.forward_rows - f0], device=dev, dtype=torch.int64)
    if world > 1:
        dist.all_reduce(counts)
    clk = clocks.stop() if clocks else None
    value = n_total * args.steps / (ms / 1e3)

    k_e2e = min(args.steps, args.e2e_steps)
    e2e_value = None
    if k_e2e > 0:                       # --e2e-steps 0: developer runs that only need the resident number
        step_e2e(wgrid)
        ms_e2e = timed(step_e2e, k_e2e)
        e2e_value = n_total * k_e2e / (ms_e2e / 1e3)

    pk = peaks()
    flops_fwd = O.flops_per_forward(cfg)
    # utilisation is reported against the forwards actually EXECUTED (a skipped unconditional forward is not claimed)
    rows_executed = int(counts[1].item())
    achieved_total = rows_executed * flops_fwd / (ms / 1e3) / 1e12 / world      # per-GPU TFLOP/s over the timed region
    hbm_kernel = args.profile_kernel in ("norm_modulate", "ode_update")
    roof_achieved = prof_work / (prof_ms / 1e3) / (1e9 if hbm_kernel else 1e12) if prof_ms > 0 else None
    roof_peak = pk["hbm"] if hbm_kernel else pk["tflops_sustained"]
    kernel_names = {"gemm_w12": "k_gemm_tc<SWIGLU> (mlp.w12, 42.8% of FLOPs)", "gemm_qkv": "k_gemm_tc<QKV> (attn.qkv + RoPE + scatter)",
                    "gemm_proj": "k_gemm_tc<GATED_RESIDUAL> (attn.proj)", "gemm_w3": "k_gemm_tc<GATED_RESIDUAL> (mlp.w3)",
                    "attention": "k_attention_tc", "norm_modulate": "k_norm_modulate", "ode_update": "k_ode_update"}
    chunk_rows = min(2 * n, args.chunk or 512)
    line = dict(
        metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
        ms_per_step=ms / args.steps, higher_is_better=True, scaling=scaling, vs_baseline=None,
        dtype="bf16" if args.precision == "bf16" else "f32", data="synthetic",
        config=dict(workload=workload_desc(args, world),
                    parallelism=f"dp{world} (samples sharded, CFG pair per GPU, one all-gather of final latents via "
                                f"va_vae_b200.dist.sample_sharded)",
                    weights="random init (oracle.init_state_dict seed 0)", l2="flushed between steps (256 MiB write)",
                    warmup=f"{args.warmup} passes over a {wpoints}-point grid (same batch, kernels and shapes)",
                    chunk_rows=chunk_rows, forwards_executed_per_step=rows_executed // max(args.steps, 1),
                    forwards_reference_per_step=2 * n_total * n_evals,
                    tflops_per_gpu_whole_step=achieved_total,
                    frac_of_burst_peak_whole_step=achieved_total / pk["tflops_burst"],
                    frac_of_sustained_peak_whole_step=achieved_total / pk["tflops_sustained"]),
        clocks=clk,
        e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=int(zz.numel() * 4 + yy.numel() * 8) * world,
                 d2h_bytes_per_step=int(zz.numel() * 4) * world, steps=k_e2e,
                 api="Engine.sample_ode_host -> vavae_sample_ode_host (pinned host buffers)"),
        gpu_launches=int(counts[0].item()),
        roofline=dict(bound="hbm" if hbm_kernel else "tensor", kernel=kernel_names[args.profile_kernel], achieved=roof_achieved,
                      peak=roof_peak, unit="GB/s" if hbm_kernel else "TFLOP/s",
                      frac=(roof_achieved / roof_peak) if roof_achieved else None,
                      traffic=measured_traffic(args.workload, chunk_rows, args.profile_kernel),
                      launches_timed=prof_n, peak_source=pk["source"] + ", sustained figure (kernel timed inside a long step)"),
    )
    if args.breakdown and rank == 0:
        bd = {}
        for kname in KERNELS:
            kms, work = eng.time_kernel(kname, chunk_rows, 20)
            is_bytes = kname in ("norm_modulate", "ode_update")
            bd[kname] = dict(ms=kms, achieved=work / (kms / 1e3) / (1e9 if is_bytes else 1e12),
                             unit="GB/s" if is_bytes else "TFLOP/s")
        line["kernel_breakdown"] = dict(rows=chunk_rows, kernels=bd)
    if world > 1:
        # the other ranks leave before the CPU leg so that it has the host cores to itself
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        if not args.no_cpu_baseline:
            cb, _ = cpu_reference_rate(args, n=4 if cfg.tokens <= 256 else 1, evals=3, warm=1)
            line["cpu_baseline"] = cb
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_engine(a)

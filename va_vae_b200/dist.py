"""Data-parallel sampling across the GPUs of one box: one process per GPU, no collective inside the
denoising loop, one all-gather of the final latents.

The reference runs independent per-rank replicas with rank-offset seeds and exchanges only barriers
(LightningDiT/inference.py:75,134,228; generate_conditional_samples_distributed.py:241-250).  Here the noise
and labels are inputs, generated once and sharded contiguously, so the gathered result is identical for every
world size.  Each conditional / unconditional CFG pair stays on one rank.
"""
from __future__ import annotations

from typing import Callable, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of n samples; the first n % world ranks take one extra sample."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def build_cfg_batch(z: torch.Tensor, y: torch.Tensor, null_label: int):
    """inference.py:205-209: z -> cat([z, z]); y -> cat([y, full(n, null)])."""
    return torch.cat([z, z], 0), torch.cat([y, torch.full_like(y, null_label)], 0)


def _collective_device(z: torch.Tensor, group, device):
    """Device the gathered tensors must live on: the caller's choice, else the current CUDA device under NCCL (the
    noise z is typically generated once on the CPU, while NCCL only moves CUDA tensors), else z's own device."""
    if device is not None:
        return torch.device(device)
    if dist.get_backend(group) == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return z.device


def sample_sharded(sample_local: Callable[[torch.Tensor, torch.Tensor], torch.Tensor], z: torch.Tensor, y: torch.Tensor,
                   group=None, gather: bool = True, device=None) -> torch.Tensor:
    """Shard (z, y) over the ranks of `group`, run `sample_local(z_shard, y_shard) -> latents_shard` on each,
    all-gather the final latents.  `sample_local` wraps Engine.sample_ode (CFG doubling included).
    Works with any backend (NCCL on the GPU box, gloo in the CPU tests).  `device`: where the shard results and the
    gathered tensor live (default: the current CUDA device under NCCL, z's device otherwise); a rank whose shard is
    empty (n < world size) contributes an empty tensor on that device."""
    if not (dist.is_available() and dist.is_initialized()):
        return sample_local(z, y)
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    n = z.shape[0]
    lo, hi = shard_range(n, rank, world)
    dev = _collective_device(z, group, device)
    if hi > lo:
        local = sample_local(z[lo:hi], y[lo:hi]).to(dev)
    else:
        local = torch.empty((0,) + tuple(z.shape[1:]), dtype=torch.float32, device=dev)
    if not gather:
        return local
    sizes = [shard_range(n, r, world) for r in range(world)]
    if n % world == 0:
        out = torch.empty((n,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    # uneven shards: pad every shard to the largest one so a single equal-size all-gather still suffices
    m = max(b - a for a, b in sizes)
    padded = torch.zeros((m,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    out = torch.empty((world * m,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    return torch.cat([out[r * m: r * m + (b - a)] for r, (a, b) in enumerate(sizes)], 0)

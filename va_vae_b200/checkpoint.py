"""Checkpoint / latent-statistics ingestion at the boundary of the sampling path.

Mirrors what the reference's callers do before sampling:
  * weight selection: `checkpoint["ema"]` preferred, else `["model"]`, else the dict itself; `module.` prefixes stripped
    (LightningDiT/inference.py:88-91; generate_conditional_samples_distributed.py:94-111);
  * shape-checked copy with the `x_embedder.proj.weight` channel special case
    (LightningDiT/train.py:275-301; generate_conditional_samples_distributed.py:36-62);
  * `latents_stats.pt` = {'mean', 'std'} of shape (1, C, 1, 1) (LightningDiT/datasets/img_latent_dataset.py:43-50);
  * de-normalisation of sampled latents `x * std / latent_multiplier + mean` (LightningDiT/inference.py:220).
Host logic only; the weights land in the CUDA engine through LightningDiT.load_state_dict.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch


def select_weights(checkpoint: Dict) -> Dict[str, torch.Tensor]:
    """EMA first (sampling uses the EMA copy), then 'model', then a bare state_dict; strips DDP's 'module.' prefix."""
    if "ema" in checkpoint:
        weights = checkpoint["ema"]
    elif "model" in checkpoint:
        weights = checkpoint["model"]
    else:
        weights = checkpoint
    return {(k[len("module."):] if k.startswith("module.") else k): v for k, v in weights.items()}


def load_weights_with_shape_check(model, checkpoint: Dict, rank: int = 0, verbose: bool = True):
    """Same contract as the reference function of this name: `checkpoint['model']` is copied tensor by tensor where
    shapes agree; `x_embedder.proj.weight` with a different channel count keeps the first 16 input channels and zeros
    the rest; everything else with a mismatching shape or an unknown name is skipped (reported on rank 0)."""
    state = model.state_dict()
    skipped = []
    for name, param in checkpoint["model"].items():
        if name not in state:
            skipped.append((name, "not in model"))
            continue
        if param.shape == state[name].shape:
            state[name].copy_(param)
        elif name == "x_embedder.proj.weight":
            weight = torch.zeros_like(state[name])
            weight[:, :16] = param[:, :16]
            state[name] = weight
        else:
            skipped.append((name, f"shape {tuple(param.shape)} vs model {tuple(state[name].shape)}"))
    if verbose and rank == 0:
        for name, why in skipped:
            print(f"Skipping parameter '{name}': {why}")
    model.load_state_dict(state, strict=False)
    return model


def _torch_load(path: str, map_location, trust_pickle: bool):
    """weights_only=True by default: a downloaded .pt is untrusted input and the full unpickler executes code.
    Training checkpoints that carry non-tensor objects (an argparse Namespace under 'args', optimizer state with
    custom classes) need trust_pickle=True."""
    try:
        return torch.load(path, map_location=map_location, weights_only=True)
    except Exception as exc:
        if not trust_pickle:
            raise RuntimeError(f"{path} cannot be read with weights_only=True ({type(exc).__name__}: {exc}); pass "
                               "trust_pickle=True if the file comes from a trusted source") from exc
        return torch.load(path, map_location=map_location, weights_only=False)


def load_checkpoint(model, path: str, map_location="cpu", rank: int = 0, verbose: bool = True, strict: bool = True,
                    trust_pickle: bool = False):
    """The sampling-path loader (LightningDiT/inference.py:88-93): torch.load -> EMA-first selection -> leading
    'module.' stripped -> `model.load_state_dict(weights)`, STRICT like the reference, so a checkpoint that does not
    match the model configuration raises instead of sampling from partly random weights.
    strict=False gives the fine-tuning behaviour of load_weights_with_shape_check (train.py:275-301): tensors whose
    shape differs or whose name is unknown are skipped and reported."""
    ckpt = _torch_load(path, map_location, trust_pickle)
    weights = select_weights(ckpt)
    if strict:
        model.load_state_dict(weights, strict=True)
        return model
    return load_weights_with_shape_check(model, {"model": weights}, rank=rank, verbose=verbose)


def load_latent_stats(path: str, map_location="cpu") -> Tuple[torch.Tensor, torch.Tensor]:
    stats = torch.load(path, map_location=map_location, weights_only=True)
    return stats["mean"], stats["std"]


def denormalize_latents(samples: torch.Tensor, latent_mean: torch.Tensor, latent_std: torch.Tensor,
                        latent_multiplier: float = 1.0) -> torch.Tensor:
    """inference.py:220 -- the step between the sampler and the VA-VAE decoder."""
    return (samples * latent_std.to(samples)) / latent_multiplier + latent_mean.to(samples)

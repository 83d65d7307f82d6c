"""TEST INFRASTRUCTURE -- golden vectors for the two "next" rows either side of the sampler,
written FROM the imported, unmodified reference (build container only):

  tests/golden/training_losses.pt   Transport.training_losses (LightningDiT/transport/transport.py:169-215) on seeded
                                    inputs for four transport configurations, with a fixed stand-in model
  tests/golden/latent_shard/*.safetensors + latent_dataset.pt
                                    two shards written with safetensors.torch.save_file exactly as
                                    LightningDiT/extract_features.py:113-128 does, and what the reference's
                                    ImgLatentDataset (LightningDiT/datasets/img_latent_dataset.py) returns for them

    python -m oracle.gen_golden_train_data
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

from oracle import ref_shim

GOLD = os.path.join(ROOT, "tests", "golden")
LOSS_CASES = {
    "uniform": dict(use_cosine_loss=False, use_lognorm=False),
    "lognorm_cos": dict(use_cosine_loss=True, use_lognorm=True),
    "lognorm_shift": dict(use_cosine_loss=True, use_lognorm=True, shift_lg=True),
    "partial": dict(use_cosine_loss=True, use_lognorm=True, partitial_train=[0.0, 0.5], partial_ratio=1.0),
}


def stand_in_model(x, t, y=None):
    """Deterministic velocity 'model' (any callable works for the loss arithmetic)."""
    return torch.tanh(x * (0.5 + t.view(-1, 1, 1, 1))) + 0.01 * y.view(-1, 1, 1, 1).float()


def gen_losses():
    _, tr = ref_shim.import_reference()
    out = {}
    for name, kw in LOSS_CASES.items():
        transport = tr.create_transport("Linear", "velocity", None, None, None, **kw)
        torch.manual_seed(11)
        np.random.seed(12)
        x1 = torch.randn(6, 4, 8, 8, generator=torch.Generator().manual_seed(5))
        y = torch.arange(6)
        terms = transport.training_losses(stand_in_model, x1, dict(y=y), shifted_mu=0.3 if kw.get("shift_lg") else 0)
        out[name] = dict(kw=kw, x1=x1, y=y, terms={k: v.clone() for k, v in terms.items()})
    torch.save(out, os.path.join(GOLD, "training_losses.pt"))
    print("wrote training_losses.pt:", {k: float(v["terms"]["loss"].mean()) for k, v in out.items()})


def gen_shards():
    from safetensors.torch import save_file
    d = os.path.join(GOLD, "latent_shard")
    shutil.rmtree(d, ignore_errors=True)
    os.makedirs(d)
    g = torch.Generator().manual_seed(21)
    for shard, n in enumerate((5, 3)):
        latents = torch.randn(n, 4, 2, 2, generator=g) * 2 + 0.5
        save_dict = {"latents": latents, "latents_flip": latents.flip(-1).contiguous(),
                     "labels": torch.randint(0, 1000, (n,), generator=g)}
        save_file(save_dict, os.path.join(d, f"latents_rank00_shard{shard:03d}.safetensors"),
                  metadata={"total_size": f"{n}", "dtype": f"{latents.dtype}", "device": f"{latents.device}"})
    sys.path.insert(0, ref_shim.REFERENCE_ROOT)
    from datasets.img_latent_dataset import ImgLatentDataset
    np.random.seed(31)
    ds = ImgLatentDataset(d, latent_norm=True, latent_multiplier=0.5)      # writes latents_stats.pt from the shards
    stats = torch.load(os.path.join(d, "latents_stats.pt"))
    np.random.seed(32)
    items = [ds[i] for i in range(len(ds))]
    ds_raw = ImgLatentDataset(d, latent_norm=False, latent_multiplier=1.0)
    np.random.seed(33)
    raw = [ds_raw[i] for i in range(len(ds_raw))]
    torch.save(dict(length=len(ds), mean=stats["mean"], std=stats["std"], features=torch.stack([f for f, _ in items]),
                    labels=torch.stack([l for _, l in items]), raw_features=torch.stack([f for f, _ in raw])),
               os.path.join(d, "latent_dataset.pt"))
    print("wrote", sorted(os.listdir(d)))


if __name__ == "__main__":
    assert ref_shim.reference_available(), "needs /root/reference"
    gen_losses()
    gen_shards()

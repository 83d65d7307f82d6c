"""Latent-shard reader against fixtures written by the reference's own writer library
(safetensors.torch.save_file, as in LightningDiT/extract_features.py:113-128) and read back by the reference's own
ImgLatentDataset (LightningDiT/datasets/img_latent_dataset.py) -- oracle/gen_golden_train_data.py made them."""
import os
import shutil

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR
from va_vae_b200.latent_shards import ImgLatentDataset, SafetensorsFile, write_latent_shard, write_safetensors

SHARDS = os.path.join(GOLDEN_DIR, "latent_shard")


@pytest.fixture()
def shard_dir(tmp_path):
    d = tmp_path / "shards"
    shutil.copytree(SHARDS, d)
    os.remove(d / "latent_dataset.pt")
    return str(d)


def test_dataset_matches_reference_dataset_outputs(shard_dir):
    want = torch.load(os.path.join(SHARDS, "latent_dataset.pt"), weights_only=True)
    np.random.seed(31)
    ds = ImgLatentDataset(shard_dir, latent_norm=True, latent_multiplier=0.5)     # uses the committed latents_stats.pt
    assert len(ds) == want["length"] == 8
    assert torch.equal(ds._latent_mean, want["mean"]) and torch.equal(ds._latent_std, want["std"])
    np.random.seed(32)                                                           # same flip choices as the reference run
    items = [ds[i] for i in range(len(ds))]
    assert torch.equal(torch.stack([f for f, _ in items]), want["features"])
    assert torch.equal(torch.stack([l for _, l in items]), want["labels"])
    np.random.seed(33)
    raw = ImgLatentDataset(shard_dir, latent_norm=False)
    assert torch.equal(torch.stack([raw[i][0] for i in range(len(raw))]), want["raw_features"])


def test_latent_stats_are_recomputed_like_the_reference_when_the_cache_is_missing(shard_dir):
    want = torch.load(os.path.join(SHARDS, "latent_dataset.pt"), weights_only=True)
    os.remove(os.path.join(shard_dir, "latents_stats.pt"))
    np.random.seed(31)                         # the reference drew its subset permutation from this seed
    ds = ImgLatentDataset(shard_dir, latent_norm=True)
    assert os.path.exists(os.path.join(shard_dir, "latents_stats.pt"))          # cache written (img_latent_dataset.py:46-48)
    assert torch.allclose(ds._latent_mean, want["mean"], rtol=0, atol=1e-6)
    assert torch.allclose(ds._latent_std, want["std"], rtol=0, atol=1e-6)
    assert ds._latent_mean.shape == (1, 4, 1, 1)


def test_reader_edge_cases(tmp_path):
    p = str(tmp_path / "x.safetensors")
    tensors = {"a": torch.arange(24, dtype=torch.float32).reshape(6, 2, 2), "e": torch.empty(0, 3),
               "h": torch.randn(5, 3).bfloat16(), "l": torch.tensor([7, -1, 2 ** 40]), "s": torch.tensor(3.5)}
    write_safetensors(p, tensors, metadata={"k": "v"})
    with SafetensorsFile(p) as f:
        assert sorted(f.keys()) == sorted(tensors) and f.metadata == {"k": "v"}
        for k, v in tensors.items():
            assert torch.equal(f.get_tensor(k), v) and f.get_shape(k) == tuple(v.shape)
        assert torch.equal(f.get_slice("a", 2, 4), tensors["a"][2:4])
        assert f.get_slice("a", 5, 99).shape == (1, 2, 2) and f.get_slice("a", 4, 4).shape == (0, 2, 2)   # ragged / empty
    with open(p, "r+b") as fh:                 # corrupt the length prefix
        fh.write((1 << 40).to_bytes(8, "little"))
    with pytest.raises(ValueError):
        SafetensorsFile(p)


def test_round_trip_with_the_safetensors_package(tmp_path):
    st = pytest.importorskip("safetensors.torch")
    lat = torch.randn(4, 32, 16, 16)
    lab = torch.randint(0, 1000, (4,))
    p = str(tmp_path / "mine.safetensors")
    write_latent_shard(p, lat, lat.flip(-1).contiguous(), lab)
    back = st.load_file(p)                                                       # the reference's loader library
    assert torch.equal(back["latents"], lat) and torch.equal(back["latents_flip"], lat.flip(-1)) and torch.equal(back["labels"], lab)
    q = str(tmp_path / "theirs.safetensors")
    st.save_file({"latents": lat, "labels": lab}, q)
    with SafetensorsFile(q) as f:
        assert torch.equal(f.get_tensor("latents"), lat) and torch.equal(f.get_slice("labels", 1, 3), lab[1:3])


def test_against_live_reference_dataset(shard_dir):
    from oracle import ref_shim
    if not ref_shim.reference_available():
        pytest.skip("/root/reference only exists in the build container")
    pytest.importorskip("safetensors")
    import sys
    sys.path.insert(0, ref_shim.REFERENCE_ROOT)
    from datasets.img_latent_dataset import ImgLatentDataset as RefDataset
    np.random.seed(5)
    ref = RefDataset(shard_dir, latent_norm=True, latent_multiplier=1.0)
    a = [ref[i] for i in range(len(ref))]
    np.random.seed(5)
    mine = ImgLatentDataset(shard_dir, latent_norm=True, latent_multiplier=1.0)
    b = [mine[i] for i in range(len(mine))]
    assert len(a) == len(b)
    for (fa, la), (fb, lb) in zip(a, b):
        assert torch.equal(fa, fb) and torch.equal(la, lb)

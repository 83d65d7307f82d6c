"""Latent shards on disk: the data format on the input side of the path.

The reference stores pre-extracted VA-VAE latents as safetensors shards written by
LightningDiT/extract_features.py:105-149 -- three tensors per file, `latents` (N, C, H, W), `latents_flip`
(N, C, H, W: the horizontally flipped image's latent) and `labels` (N,), plus string metadata -- and reads them back
through `ImgLatentDataset` (LightningDiT/datasets/img_latent_dataset.py:16-90), which also owns the
`latents_stats.pt` cache ({'mean', 'std'} of shape (1, C, 1, 1)) that the sampler's de-normalisation consumes
(LightningDiT/inference.py:150-157,220).

`ImgLatentDataset` below mirrors that class (same constructor, `__len__`, `__getitem__`, `get_latent_stats`,
`compute_latent_stats`, same use of numpy's global RNG for the flip choice and the statistics subset).  The file
format is read by a small zero-copy reader of the published safetensors layout -- 8-byte little-endian header length,
JSON header {name: {dtype, shape, data_offsets}, "__metadata__": {...}}, then one flat byte buffer -- through a
memory map, so slicing one image out of a 10 000-image shard touches only that image's pages and the `safetensors`
package is not required on the sampling box.  `write_latent_shard` produces files the reference's loader accepts.
"""
from __future__ import annotations

import json
import mmap
import os
import struct
from glob import glob
from typing import Dict, Optional, Tuple

import numpy as np
import torch

_DTYPES = {
    "F64": (torch.float64, 8), "F32": (torch.float32, 4), "F16": (torch.float16, 2), "BF16": (torch.bfloat16, 2),
    "I64": (torch.int64, 8), "I32": (torch.int32, 4), "I16": (torch.int16, 2), "I8": (torch.int8, 1),
    "U8": (torch.uint8, 1), "BOOL": (torch.bool, 1),
}
_DTYPE_NAMES = {v[0]: k for k, v in _DTYPES.items()}


class SafetensorsFile:
    """Read-only, memory-mapped view of one .safetensors file."""

    def __init__(self, path: str):
        self.path = path
        self._f = open(path, "rb")
        size = os.fstat(self._f.fileno()).st_size
        if size < 8:
            raise ValueError(f"{path}: not a safetensors file (shorter than its length prefix)")
        (hlen,) = struct.unpack("<Q", self._f.read(8))
        if hlen > size - 8 or hlen > (100 << 20):
            raise ValueError(f"{path}: header length {hlen} exceeds the file")
        header = json.loads(self._f.read(hlen).decode("utf-8"))
        self.metadata: Dict[str, str] = header.pop("__metadata__", {}) or {}
        self._base = 8 + hlen
        self._entries = {}
        for name, ent in header.items():
            if ent["dtype"] not in _DTYPES:
                raise ValueError(f"{path}: tensor {name!r} has unsupported dtype {ent['dtype']}")
            dtype, esz = _DTYPES[ent["dtype"]]
            shape = tuple(int(d) for d in ent["shape"])
            lo, hi = (int(v) for v in ent["data_offsets"])
            if hi - lo != esz * int(np.prod(shape, dtype=np.int64)) or self._base + hi > size or lo < 0:
                raise ValueError(f"{path}: tensor {name!r} has inconsistent offsets {lo}:{hi} for shape {shape}")
            self._entries[name] = (dtype, esz, shape, lo)
        self._mm = mmap.mmap(self._f.fileno(), 0, access=mmap.ACCESS_READ) if size > 0 else None

    def keys(self):
        return list(self._entries)

    def get_shape(self, name: str) -> Tuple[int, ...]:
        return self._entries[name][2]

    def get_slice(self, name: str, start: int, stop: int) -> torch.Tensor:
        """Rows [start, stop) of tensor `name` along its first dimension, as an owning tensor (one copy of those rows)."""
        dtype, esz, shape, lo = self._entries[name]
        if not shape:
            raise IndexError(f"{name!r} is a scalar")
        start, stop = max(0, start), min(shape[0], stop)
        row = esz * int(np.prod(shape[1:], dtype=np.int64))
        off = self._base + lo + start * row
        nbytes = max(stop - start, 0) * row
        out = torch.empty((max(stop - start, 0),) + shape[1:], dtype=dtype)
        if nbytes:
            # np.frombuffer accepts the read-only map; the copy into `out` is the only copy of these rows
            src = np.frombuffer(self._mm, dtype=np.uint8, count=nbytes, offset=off)
            out.reshape(-1).view(torch.uint8).numpy()[:] = src
        return out

    def get_tensor(self, name: str) -> torch.Tensor:
        shape = self._entries[name][2]
        if not shape:
            dtype, esz, _, lo = self._entries[name]
            off = self._base + lo
            return torch.frombuffer(bytearray(self._mm[off: off + esz]), dtype=dtype).reshape(())
        return self.get_slice(name, 0, shape[0])

    def close(self):
        if self._mm is not None:
            self._mm.close()
            self._mm = None
        self._f.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def write_safetensors(path: str, tensors: Dict[str, torch.Tensor], metadata: Optional[Dict[str, str]] = None):
    """Writes the published layout (header padded with spaces to 8 bytes, tensors in name order, contiguous)."""
    header, chunks, off = {}, [], 0
    if metadata:
        header["__metadata__"] = {str(k): str(v) for k, v in metadata.items()}
    for name in sorted(tensors):
        t = tensors[name].detach().cpu().contiguous()
        if t.dtype not in _DTYPE_NAMES:
            raise ValueError(f"unsupported dtype {t.dtype} for {name!r}")
        raw = t.reshape(-1).view(torch.uint8).numpy().tobytes() if t.numel() else b""
        header[name] = {"dtype": _DTYPE_NAMES[t.dtype], "shape": list(t.shape), "data_offsets": [off, off + len(raw)]}
        chunks.append(raw)
        off += len(raw)
    hj = json.dumps(header, separators=(",", ":")).encode("utf-8")
    hj += b" " * ((8 - len(hj) % 8) % 8)
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(hj)))
        f.write(hj)
        for raw in chunks:
            f.write(raw)


def write_latent_shard(path: str, latents: torch.Tensor, latents_flip: torch.Tensor, labels: torch.Tensor):
    """One shard in the layout of extract_features.py:113-128 (tensor names and metadata keys included)."""
    if latents.shape != latents_flip.shape or labels.shape[0] != latents.shape[0]:
        raise ValueError("latents / latents_flip / labels disagree in shape")
    write_safetensors(path, {"latents": latents, "latents_flip": latents_flip, "labels": labels},
                      metadata={"total_size": f"{latents.shape[0]}", "dtype": f"{latents.dtype}",
                                "device": f"{latents.device}"})


class ImgLatentDataset(torch.utils.data.Dataset):
    """Mirror of datasets/img_latent_dataset.py:16-90 (ImageNet latent dataset over safetensors shards)."""

    def __init__(self, data_dir, latent_norm=True, latent_multiplier=1.0):
        self.data_dir = data_dir
        self.latent_norm = latent_norm
        self.latent_multiplier = latent_multiplier
        self.files = sorted(glob(os.path.join(data_dir, "*.safetensors")))
        self._open: Dict[str, SafetensorsFile] = {}
        self.img_to_file_map = self.get_img_to_safefile_map()
        if latent_norm:
            self._latent_mean, self._latent_std = self.get_latent_stats()

    def _file(self, path: str) -> SafetensorsFile:
        f = self._open.get(path)
        if f is None:
            f = self._open[path] = SafetensorsFile(path)      # header parsed once per shard, pages mapped lazily
        return f

    def get_img_to_safefile_map(self):
        img_to_file = {}
        for safe_file in self.files:
            num_imgs = self._file(safe_file).get_shape("labels")[0]
            cur_len = len(img_to_file)
            for i in range(num_imgs):
                img_to_file[cur_len + i] = {"safe_file": safe_file, "idx_in_file": i}
        return img_to_file

    def get_latent_stats(self):
        cache = os.path.join(self.data_dir, "latents_stats.pt")
        if not os.path.exists(cache):
            latent_stats = self.compute_latent_stats()
            torch.save(latent_stats, cache)
        else:
            latent_stats = torch.load(cache, weights_only=True)
        return latent_stats["mean"], latent_stats["std"]

    def compute_latent_stats(self):
        """Channel-wise mean / std over min(10000, N) images chosen with numpy's global RNG (img_latent_dataset.py:52-68)."""
        num_samples = min(10000, len(self.img_to_file_map))
        random_indices = np.random.choice(len(self.img_to_file_map), num_samples, replace=False)
        latents = []
        for idx in random_indices:
            info = self.img_to_file_map[idx]
            i = info["idx_in_file"]
            latents.append(self._file(info["safe_file"]).get_slice("latents", i, i + 1))
        latents = torch.cat(latents, dim=0)
        return {"mean": latents.mean(dim=[0, 2, 3], keepdim=True), "std": latents.std(dim=[0, 2, 3], keepdim=True)}

    def __len__(self):
        return len(self.img_to_file_map.keys())

    def __getitem__(self, idx):
        info = self.img_to_file_map[idx]
        f, i = self._file(info["safe_file"]), info["idx_in_file"]
        tensor_key = "latents" if np.random.uniform(0, 1) > 0.5 else "latents_flip"
        feature = f.get_slice(tensor_key, i, i + 1)
        label = f.get_slice("labels", i, i + 1)
        if self.latent_norm:
            feature = (feature - self._latent_mean) / self._latent_std
        feature = feature * self.latent_multiplier
        return feature.squeeze(0), label.squeeze(0)

    def __getstate__(self):                   # DataLoader workers: re-open the maps in the child
        state = dict(self.__dict__)
        state["_open"] = {}
        return state

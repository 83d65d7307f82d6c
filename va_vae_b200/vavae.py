"""Host-side mirror of the reference's VA-VAE decode interface.

Reference objects this replaces, for inference:
  * ``AutoencoderKL`` (LightningDiT/tokenizer/autoencoder.py:452-493): ``decode(z)`` = post_quant_conv -> Decoder;
    same constructor arguments (``embed_dim``, ``ch_mult``, ``model_type``), same ``state_dict`` key set for the
    tensors the decode path reads (``post_quant_conv.*``, ``decoder.*``), same checkpoint filtering in
    ``init_from_ckpt`` (:476-486).  The encoder half is not built: ``encode`` raises.
  * ``VA_VAE.decode_to_images`` (LightningDiT/tokenizer/vavae.py:72-82): uint8 NHWC numpy images.
All arithmetic runs in the CUDA library through the C ABI (include/vavae_b200.h, ``vavae_decoder_*``); there is no
PyTorch or CPU fallback.
"""
import ctypes
from typing import Dict, Sequence

import numpy as np
import torch
import torch.nn as nn

from .engine import ABI_VERSION, PREC_BF16, PREC_FP32, VavaeDecoderConfig, _check, _f32, _stream, load_library


def decoder_param_shapes(z_channels, ch, ch_mult, num_res_blocks, attn_resolutions, resolution, out_ch=3):
    """(name, shape) of every tensor the decode path reads, in the reference's naming (autoencoder.py:277-399,469)."""
    spec = [("post_quant_conv.weight", (z_channels, z_channels, 1, 1)), ("post_quant_conv.bias", (z_channels,))]
    levels = len(ch_mult)
    block_in = ch * ch_mult[-1]
    curr = resolution // 2 ** (levels - 1)

    def conv(name, cin, cout, k):
        spec.extend([(name + ".weight", (cout, cin, k, k)), (name + ".bias", (cout,))])

    def norm(name, c):
        spec.extend([(name + ".weight", (c,)), (name + ".bias", (c,))])

    def resblock(name, cin, cout):
        norm(name + ".norm1", cin); conv(name + ".conv1", cin, cout, 3)
        norm(name + ".norm2", cout); conv(name + ".conv2", cout, cout, 3)
        if cin != cout:
            conv(name + ".nin_shortcut", cin, cout, 1)

    def attn(name, c):
        norm(name + ".norm", c)
        for p in ("q", "k", "v", "proj_out"):
            conv(name + "." + p, c, c, 1)

    conv("decoder.conv_in", z_channels, block_in, 3)
    resblock("decoder.mid.block_1", block_in, block_in)
    attn("decoder.mid.attn_1", block_in)
    resblock("decoder.mid.block_2", block_in, block_in)
    for lvl in reversed(range(levels)):
        block_out = ch * ch_mult[lvl]
        for ib in range(num_res_blocks + 1):
            resblock(f"decoder.up.{lvl}.block.{ib}", block_in, block_out)
            block_in = block_out
            if curr in attn_resolutions:
                attn(f"decoder.up.{lvl}.attn.{ib}", block_in)
        if lvl != 0:
            conv(f"decoder.up.{lvl}.upsample.conv", block_in, block_in, 3)
            curr *= 2
    norm("decoder.norm_out", block_in)
    conv("decoder.conv_out", block_in, out_ch, 3)
    return spec


class DecoderEngine:
    """ctypes wrapper of one native decoder (one per process and GPU)."""

    def __init__(self, config: dict, state_dict: Dict[str, torch.Tensor], device, precision: str = "bf16",
                 max_batch: int = 0):
        lib = load_library()
        if not torch.cuda.is_available():
            raise RuntimeError("vavae_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("vavae_b200 decoders live on CUDA devices only")
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", idx)
        mult, ares = list(config["ch_mult"]), list(config["attn_resolutions"])
        if len(mult) > 8 or len(ares) > 4:
            raise ValueError("at most 8 levels and 4 attention resolutions")
        cfg = VavaeDecoderConfig(abi_version=ABI_VERSION, z_channels=int(config["z_channels"]), ch=int(config["ch"]),
                                 num_levels=len(mult), ch_mult=(ctypes.c_int32 * 8)(*mult),
                                 num_res_blocks=int(config["num_res_blocks"]), num_attn_res=len(ares),
                                 attn_res=(ctypes.c_int32 * 4)(*ares), resolution=int(config["resolution"]),
                                 out_ch=int(config.get("out_ch", 3)),
                                 precision={"bf16": PREC_BF16, "fp32": PREC_FP32}[precision], max_batch=int(max_batch))
        self.config = dict(config)
        self.precision = precision
        self._lib = lib
        self._h = ctypes.c_void_p()
        _check(lib.vavae_decoder_create(ctypes.byref(cfg), idx, ctypes.byref(self._h)))
        try:
            with torch.cuda.device(self.device):
                for name, _ in decoder_param_shapes(config["z_channels"], config["ch"], mult, config["num_res_blocks"],
                                                    ares, config["resolution"], config.get("out_ch", 3)):
                    if name not in state_dict:
                        raise KeyError(f"decoder weight missing from state_dict: {name}")
                    t = state_dict[name].detach().to(dtype=torch.float32).contiguous()
                    shape = (ctypes.c_int64 * t.dim())(*t.shape)
                    _check(lib.vavae_decoder_load_weight(self._h, name.encode(), ctypes.c_void_p(t.data_ptr()), shape,
                                                         t.dim(), 1 if t.is_cuda else 0))
                _check(lib.vavae_decoder_finalize(self._h))
        except Exception:
            lib.vavae_decoder_destroy(self._h)
            self._h = None
            raise
        self.z_res = int(config["resolution"]) // 2 ** (len(mult) - 1)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.vavae_decoder_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def decode(self, z: torch.Tensor, want_float: bool = True, want_uint8: bool = False):
        z = _f32(z, self.device)
        c = self.config
        if z.dim() != 4 or tuple(z.shape[1:]) != (c["z_channels"], self.z_res, self.z_res):
            raise ValueError(f"expected latents of shape (n, {c['z_channels']}, {self.z_res}, {self.z_res}), got {tuple(z.shape)}")
        n, R, oc = z.shape[0], int(c["resolution"]), int(c.get("out_ch", 3))
        out = torch.empty((n, oc, R, R), device=self.device, dtype=torch.float32) if want_float else None
        img = torch.empty((n, R, R, oc), device=self.device, dtype=torch.uint8) if want_uint8 else None
        if n == 0:
            return out, img
        with torch.cuda.device(self.device):
            _check(self._lib.vavae_decoder_decode(self._h, z.data_ptr(), n, out.data_ptr() if want_float else None,
                                                  img.data_ptr() if want_uint8 else None,
                                                  ctypes.c_void_p(_stream(self.device))))
        return out, img

    def launch_count(self) -> int:
        return int(self._lib.vavae_decoder_launch_count(self._h))

    def device_bytes(self) -> int:
        return int(self._lib.vavae_decoder_device_bytes(self._h))


class AutoencoderKL(nn.Module):
    """Decode half of the reference's AutoencoderKL (autoencoder.py:452-493) backed by the native decoder.

    Holds the reference's parameters for ``post_quant_conv`` and ``decoder.*`` (so ``load_state_dict`` /
    ``init_from_ckpt`` work on reference checkpoints; encoder / loss / foundation-model entries are ignored exactly as
    the reference's ``strict=False`` load ignores what it does not own)."""

    def __init__(self, embed_dim=32, ch_mult=(1, 1, 2, 2, 4), use_variational=True, ckpt_path=None, model_type="vavae",
                 ch=128, num_res_blocks=2, resolution=256, out_ch=3, attn_resolutions=None, precision="bf16"):
        super().__init__()
        if model_type not in ("vavae", "marvae"):
            raise ValueError(f"Invalid model type: {model_type}")            # autoencoder.py:486
        if attn_resolutions is None:
            attn_resolutions = (16,) if model_type == "vavae" else ()          # autoencoder.py:455-458
        self.embed_dim, self.model_type, self.precision = embed_dim, model_type, precision
        self.decoder_config = dict(z_channels=embed_dim, ch=ch, ch_mult=tuple(ch_mult), num_res_blocks=num_res_blocks,
                                   attn_resolutions=tuple(attn_resolutions), resolution=resolution, out_ch=out_ch)
        self._names = []
        for name, shape in decoder_param_shapes(embed_dim, ch, tuple(ch_mult), num_res_blocks, tuple(attn_resolutions),
                                                resolution, out_ch):
            self._register(name, torch.zeros(shape))
        self._engine = None
        self._engine_key = None
        if ckpt_path is not None:
            self.init_from_ckpt(ckpt_path)

    def _register(self, dotted, value):
        mod = self
        parts = dotted.split(".")
        for p in parts[:-1]:
            if not hasattr(mod, p):
                mod.add_module(p, nn.Module())
            mod = getattr(mod, p)
        mod.register_parameter(parts[-1], nn.Parameter(value, requires_grad=False))
        self._names.append(dotted)

    def init_from_ckpt(self, path):
        """autoencoder.py:476-486: 'state_dict' unwrapped for vavae, ['model'] for marvae, loss.* dropped, strict=False."""
        sd = torch.load(path, map_location="cpu")
        if self.model_type == "vavae":
            if "state_dict" in sd.keys():
                sd = sd["state_dict"]
            sd = {k: v for k, v in sd.items() if "loss" not in k}
        else:
            sd = sd["model"]
        own = set(self._names)
        return self.load_state_dict({k: v for k, v in sd.items() if k in own}, strict=False)

    def engine(self) -> DecoderEngine:
        p0 = next(self.parameters())
        if p0.device.type != "cuda":
            raise RuntimeError("va_vae_b200.AutoencoderKL runs on CUDA devices only (move it with .to('cuda')); "
                               "there is no CPU fallback")
        key = (p0.device, self.precision, tuple(p._version for p in self.parameters()),
               tuple(p.data_ptr() for p in self.parameters()))
        if self._engine is None or key != self._engine_key:
            if self._engine is not None:
                self._engine.close()
            self._engine = DecoderEngine(self.decoder_config, dict(self.state_dict()), p0.device, self.precision)
            self._engine_key = key
        return self._engine

    def encode(self, x):
        raise NotImplementedError("the encoder half of AutoencoderKL is outside this engine's path (SURVEY.md 8f)")

    @torch.no_grad()
    def decode(self, z):
        """AutoencoderKL.decode, autoencoder.py:490-493: (n, embed_dim, r, r) -> (n, 3, R, R) float."""
        return self.engine().decode(z, want_float=True)[0]

    @torch.no_grad()
    def decode_to_images(self, z) -> np.ndarray:
        """VA_VAE.decode_to_images, vavae.py:72-82: uint8 (n, R, R, 3) numpy array."""
        return self.engine().decode(z, want_float=False, want_uint8=True)[1].cpu().numpy()


class VA_VAE:
    """tokenizer/vavae.py:20-82, decode side: ``VA_VAE(config)`` with ``config`` either the reference's YAML path (needs
    omegaconf or PyYAML) or a dict with ``embed_dim`` and ``ckpt_path``."""

    def __init__(self, config, img_size=256, horizon_flip=0.5, fp16=True, device="cuda"):
        if isinstance(config, str):
            import yaml
            with open(config) as f:
                y = yaml.safe_load(f)
            config = {"embed_dim": y["model"]["params"]["embed_dim"], "ckpt_path": y.get("ckpt_path")}
        self.embed_dim = config["embed_dim"]
        self.ckpt_path = config.get("ckpt_path")
        self.img_size = img_size
        self.horizon_flip = horizon_flip
        self.model = AutoencoderKL(embed_dim=self.embed_dim, ch_mult=(1, 1, 2, 2, 4), ckpt_path=self.ckpt_path,
                                   resolution=img_size).to(device).eval()          # vavae.py:36-42

    def encode_images(self, images):
        raise NotImplementedError("encoding is outside this engine's path (SURVEY.md 8f)")

    def decode_to_images(self, z):
        return self.model.decode_to_images(z)

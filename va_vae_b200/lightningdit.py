"""Drop-in mirror of the reference's `models/lightningdit.py` public surface, backed by the CUDA engine.

Same constructor kwargs, registry names, `forward(x, t, y)`, `forward_with_cfg(...)`, `in_channels`
attribute and `state_dict()` key set as /root/reference/LightningDiT/models/lightningdit.py:275-538, so a
caller such as LightningDiT/inference.py:71-93 (`LightningDiT_models[name](...)`, `load_state_dict(ckpt["ema"])`,
`.eval()`, `.to(device)`) works unchanged.  The module holds parameters only; every numeric operation runs in
libvavae_b200.so (sm_100a kernels).  There is no PyTorch / CPU fallback: calling forward without the
library or without a B200 raises.

Extra (keyword-only, optional) constructor arguments: `precision` ("bf16" tensor-core path | "fp32"
validation path) and `max_chunk_rows`.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn as nn

from .engine import Engine


class _Holder(nn.Module):
    """Parameter container; carries no forward logic."""


def _linear(out_f, in_f, bias=True):
    m = _Holder()
    m.weight = nn.Parameter(torch.empty(out_f, in_f))
    if bias:
        m.bias = nn.Parameter(torch.empty(out_f))
    return m


def _norm(dim, rms, affine=True):
    m = _Holder()
    if rms:
        m.weight = nn.Parameter(torch.ones(dim))               # rmsnorm.py:50
    elif affine:
        m.weight = nn.Parameter(torch.ones(dim))               # nn.LayerNorm default affine (q_norm / k_norm)
        m.bias = nn.Parameter(torch.zeros(dim))
    return m


def _sequential(**children):
    m = _Holder()
    for k, v in children.items():
        m.add_module(k, v)
    return m


def get_2d_sincos_pos_embed(embed_dim, grid_size):
    """Same table as lightningdit.py:444-491 (w coordinate first; each half = [sin | cos])."""
    omega = np.arange(embed_dim // 4, dtype=np.float64) / (embed_dim / 4.0)
    omega = 1.0 / 10000 ** omega
    gh = np.arange(grid_size, dtype=np.float32)
    gw = np.arange(grid_size, dtype=np.float32)
    ww, hh = np.meshgrid(gw, gh)
    parts = []
    for pos in (ww.reshape(-1), hh.reshape(-1)):
        ang = np.einsum("m,d->md", pos, omega)
        parts.append(np.concatenate([np.sin(ang), np.cos(ang)], axis=1))
    return np.concatenate(parts, axis=1)


def rope_tables(head_dim, grid):
    """feat_rope.freqs_cos / freqs_sin buffers, pos_embed.py:96-135 with dim = head_dim // 2."""
    dim = head_dim // 2
    freqs = 1.0 / (10000 ** (torch.arange(0, dim, 2)[: dim // 2].float() / dim))
    t = torch.arange(grid) / grid * grid
    f = torch.einsum("i,f->if", t, freqs).repeat_interleave(2, dim=-1)
    full = torch.cat([f[:, None, :].expand(grid, grid, -1), f[None, :, :].expand(grid, grid, -1)], dim=-1)
    full = full.reshape(grid * grid, -1)
    return full.cos(), full.sin()


class LightningDiT(nn.Module):
    """Diffusion model with a Transformer backbone (B200 engine behind the reference interface)."""

    def __init__(self, input_size=32, patch_size=2, in_channels=32, hidden_size=1152, depth=28, num_heads=16,
                 mlp_ratio=4.0, class_dropout_prob=0.1, num_classes=1000, learn_sigma=False, use_qknorm=False,
                 use_swiglu=False, use_rope=False, use_rmsnorm=False, wo_shift=False, use_checkpoint=False,
                 *, precision="bf16", max_chunk_rows=0):
        super().__init__()
        assert hidden_size % num_heads == 0, "dim should be divisible by num_heads"
        self.learn_sigma = learn_sigma
        self.in_channels = in_channels
        self.out_channels = in_channels if not learn_sigma else in_channels * 2
        self.patch_size = patch_size
        self.num_heads = num_heads
        self.use_rope = use_rope
        self.use_rmsnorm = use_rmsnorm
        self.use_qknorm = use_qknorm
        self.use_swiglu = use_swiglu
        self.wo_shift = wo_shift
        self.depth = depth
        self.hidden_size = hidden_size
        self.use_checkpoint = use_checkpoint     # accepted for signature parity; inference only
        self.input_size = input_size
        self.num_classes = num_classes
        self.class_dropout_prob = class_dropout_prob
        self.precision = precision
        self.max_chunk_rows = max_chunk_rows
        D, hd = hidden_size, hidden_size // num_heads
        grid = input_size // patch_size
        self.num_patches = grid * grid
        mlp_hidden = int(hidden_size * mlp_ratio)
        self.mlp_hidden = int(2 / 3 * mlp_hidden) if use_swiglu else mlp_hidden   # lightningdit.py:213-224

        # ---- parameter tree with the reference's key names
        self.x_embedder = _sequential(proj=_Holder())
        self.x_embedder.proj.weight = nn.Parameter(torch.empty(D, in_channels, patch_size, patch_size))
        self.x_embedder.proj.bias = nn.Parameter(torch.empty(D))
        self.t_embedder = _sequential(mlp=_sequential(**{"0": _linear(D, 256), "2": _linear(D, D)}))
        self.y_embedder = _sequential(embedding_table=_Holder())
        rows = num_classes + (1 if class_dropout_prob > 0 else 0)
        self.y_embedder.embedding_table.weight = nn.Parameter(torch.empty(rows, D))
        self.pos_embed = nn.Parameter(torch.zeros(1, self.num_patches, D), requires_grad=False)
        if use_rope:
            self.feat_rope = _Holder()
            cos, sin = rope_tables(hd, grid)
            self.feat_rope.register_buffer("freqs_cos", cos)
            self.feat_rope.register_buffer("freqs_sin", sin)
        else:
            self.feat_rope = None
        blocks = []
        for _ in range(depth):
            b = _Holder()
            if use_rmsnorm:
                b.norm1, b.norm2 = _norm(D, True), _norm(D, True)
            else:
                b.norm1, b.norm2 = _Holder(), _Holder()        # affine-free LayerNorm: no parameters
            b.attn = _Holder()
            b.attn.qkv = _linear(3 * D, D)
            if use_qknorm:
                b.attn.q_norm, b.attn.k_norm = _norm(hd, use_rmsnorm), _norm(hd, use_rmsnorm)
            b.attn.proj = _linear(D, D)
            b.mlp = _Holder()
            if use_swiglu:
                b.mlp.w12, b.mlp.w3 = _linear(2 * self.mlp_hidden, D), _linear(D, self.mlp_hidden)
            else:
                b.mlp.fc1, b.mlp.fc2 = _linear(self.mlp_hidden, D), _linear(D, self.mlp_hidden)
            b.adaLN_modulation = _sequential(**{"1": _linear((4 if wo_shift else 6) * D, D)})
            blocks.append(b)
        self.blocks = nn.ModuleList(blocks)
        self.final_layer = _Holder()
        self.final_layer.norm_final = _norm(D, True) if use_rmsnorm else _Holder()
        self.final_layer.linear = _linear(patch_size * patch_size * self.out_channels, D)
        self.final_layer.adaLN_modulation = _sequential(**{"1": _linear(2 * D, D)})
        self.initialize_weights()
        self._engine = None
        self._engine_key = None

    # ---------------------------------------------------------------- init (lightningdit.py:340-374)
    def initialize_weights(self):
        for m in self.modules():
            w = getattr(m, "weight", None)
            if isinstance(w, nn.Parameter) and w.dim() == 2 and hasattr(m, "bias") and m.bias is not None \
                    and m.bias.dim() == 1 and m.bias.shape[0] == w.shape[0]:
                nn.init.xavier_uniform_(w)
                nn.init.constant_(m.bias, 0)
        pe = get_2d_sincos_pos_embed(self.pos_embed.shape[-1], int(self.num_patches ** 0.5))
        self.pos_embed.data.copy_(torch.from_numpy(pe).float().unsqueeze(0))
        w = self.x_embedder.proj.weight.data
        nn.init.xavier_uniform_(w.view([w.shape[0], -1]))
        nn.init.constant_(self.x_embedder.proj.bias, 0)
        nn.init.normal_(self.y_embedder.embedding_table.weight, std=0.02)
        nn.init.normal_(getattr(self.t_embedder.mlp, "0").weight, std=0.02)
        nn.init.normal_(getattr(self.t_embedder.mlp, "2").weight, std=0.02)
        for b in self.blocks:                                   # adaLN-zero
            nn.init.constant_(getattr(b.adaLN_modulation, "1").weight, 0)
            nn.init.constant_(getattr(b.adaLN_modulation, "1").bias, 0)
        nn.init.constant_(getattr(self.final_layer.adaLN_modulation, "1").weight, 0)
        nn.init.constant_(getattr(self.final_layer.adaLN_modulation, "1").bias, 0)
        nn.init.constant_(self.final_layer.linear.weight, 0)
        nn.init.constant_(self.final_layer.linear.bias, 0)

    # ---------------------------------------------------------------- engine plumbing
    def engine_config(self) -> dict:
        return dict(input_size=self.input_size, patch_size=self.patch_size, in_channels=self.in_channels,
                    hidden_size=self.hidden_size, depth=self.depth, num_heads=self.num_heads,
                    mlp_hidden=self.mlp_hidden, num_classes=self.num_classes,
                    label_rows=self.y_embedder.embedding_table.weight.shape[0], learn_sigma=int(self.learn_sigma),
                    use_qknorm=int(self.use_qknorm), use_swiglu=int(self.use_swiglu), use_rope=int(self.use_rope),
                    use_rmsnorm=int(self.use_rmsnorm), wo_shift=int(self.wo_shift))

    def _weights_key(self):
        dev = self.pos_embed.device
        tensors = self.__dict__.get("_key_tensors")
        if tensors is None:   # the module tree is fixed after __init__: walk it once, not on every forward call
            tensors = list(self.parameters()) + list(self.buffers())   # buffers: feat_rope.freqs_cos / freqs_sin
            self.__dict__["_key_tensors"] = tensors
        return (str(dev), self.precision, tuple(p._version for p in tensors), tuple(p.data_ptr() for p in tensors))

    def invalidate_engine(self):
        """Forces the next call to re-upload the weights.  The engine holds its own packed COPY of the parameters;
        load_state_dict(), .to()/.cuda()/.float() and in-place ops that bump the autograd version counter are
        detected automatically, but writes through `.data` (EMA-style `p.data.mul_(..)`, `p.data.copy_(..)`) are
        invisible to PyTorch's version counter - call this after such an update."""
        self._engine_key = None

    refresh_weights = invalidate_engine

    def _apply(self, fn, *args, **kwargs):                          # .to() / .cuda() / .float() / .half()
        self._engine_key = None
        self.__dict__.pop("_key_tensors", None)                     # _apply may replace Parameter objects
        return super()._apply(fn, *args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        self._engine_key = None
        return super().load_state_dict(*args, **kwargs)

    def engine(self) -> Engine:
        """Builds (or rebuilds after a weight / device change) the CUDA engine for this module."""
        key = self._weights_key()
        if self._engine is None or key != self._engine_key:
            dev = self.pos_embed.device
            if dev.type != "cuda":
                raise RuntimeError("LightningDiT (B200 engine) must be moved to a CUDA device before use: "
                                   "there is no CPU fallback. Call model.to('cuda').")
            if self._engine is not None:
                self._engine.close()
            sd = dict(self.state_dict())
            half = 128
            # lightningdit.py:121-123, computed with the same torch ops as the reference
            sd["t_embedder.freqs"] = torch.exp(-math.log(10000) * torch.arange(0, half, dtype=torch.float32) / half)
            self._engine = Engine(self.engine_config(), sd, dev, self.precision, self.max_chunk_rows)
            self._engine_key = key
        return self._engine

    def _require_eval(self):
        if self.training and self.class_dropout_prob > 0:
            raise NotImplementedError("the B200 engine implements the sampling (eval) path only; call model.eval() "
                                      "(training-time label dropout, lightningdit.py:163-169, is out of scope)")

    # ---------------------------------------------------------------- reference API
    @torch.no_grad()
    def forward(self, x, t=None, y=None):
        """x: (N, C, H, W), t: (N,), y: (N,) -> (N, C, H, W); lightningdit.py:391-418."""
        self._require_eval()
        return self.engine().forward(x, t, y)

    @torch.no_grad()
    def forward_with_cfg(self, x, t, y, cfg_scale, cfg_interval=None, cfg_interval_start=None):
        """lightningdit.py:420-442: guidance on the first three channels; both halves get the guided value."""
        self._require_eval()
        return self.engine().forward_with_cfg(x, t, y, cfg_scale, cfg_interval, cfg_interval_start, cfg_channels=3)


def _make(depth, hidden_size, patch_size, num_heads):
    def ctor(**kwargs):
        return LightningDiT(depth=depth, hidden_size=hidden_size, patch_size=patch_size, num_heads=num_heads, **kwargs)
    return ctor


# registry, lightningdit.py:498-538
LightningDiT_models = {
    "LightningDiT-S/1": _make(12, 384, 1, 6), "LightningDiT-S/2": _make(12, 384, 2, 6),
    "LightningDiT-B/1": _make(12, 768, 1, 12), "LightningDiT-B/2": _make(12, 768, 2, 12),
    "LightningDiT-L/2": _make(24, 1024, 2, 16),
    "LightningDiT-XL/1": _make(28, 1152, 1, 16), "LightningDiT-XL/2": _make(28, 1152, 2, 16),
    "LightningDiT-1p0B/1": _make(24, 1536, 1, 24), "LightningDiT-1p0B/2": _make(24, 1536, 2, 24),
    "LightningDiT-1p6B/1": _make(28, 1792, 1, 28), "LightningDiT-1p6B/2": _make(28, 1792, 2, 28),
}

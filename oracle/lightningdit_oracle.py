"""TEST INFRASTRUCTURE ONLY -- CPU restatement ("port") of the reference hot path.

This is the checker for the CUDA engine.  Only tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py may import it.  The product path
(va_vae_b200/) never does: it fails loudly when the CUDA library is missing.

Parity status: PINNED against the unmodified reference.  oracle/gen_golden.py imports the
real reference modules from /root/reference (under the sys.modules shims in
oracle/ref_shim.py), runs them on seeded inputs and stores the tensors in tests/golden/;
tests/test_oracle_golden.py checks this restatement against those fixtures on every CPU
run, and tests/test_oracle_vs_reference.py checks it against the live reference whenever
/root/reference is present.  The reference's own test-suite holds no golden vectors for
this path, and the Euler/Heun arithmetic lives in the un-vendored,
un-pinned dependency `torchdiffeq` (LightningDiT/requirements.txt:10): its published
fixed-grid semantics are restated in `sample_ode` below.

Everything is a pure function of a reference-format ``state_dict`` (fp32, nn.Linear
(out,in) layout; key names as produced by LightningDiT/models/lightningdit.py) so that
the same weights feed the reference, this oracle and the engine.
All `file:line` citations are relative to /root/reference/LightningDiT/.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch
import torch.nn.functional as F


@dataclass
class DiTConfig:
    """Constructor arguments of LightningDiT (models/lightningdit.py:279-297)."""
    input_size: int = 16
    patch_size: int = 1
    in_channels: int = 32
    hidden_size: int = 384
    depth: int = 12
    num_heads: int = 6
    mlp_ratio: float = 4.0
    class_dropout_prob: float = 0.1
    num_classes: int = 1000
    learn_sigma: bool = False
    use_qknorm: bool = False
    use_swiglu: bool = True
    use_rope: bool = True
    use_rmsnorm: bool = True
    wo_shift: bool = False

    @property
    def tokens(self):
        return (self.input_size // self.patch_size) ** 2

    @property
    def head_dim(self):
        return self.hidden_size // self.num_heads

    @property
    def out_channels(self):
        return self.in_channels * (2 if self.learn_sigma else 1)

    @property
    def mlp_hidden(self):
        # models/lightningdit.py:213,217 -- int(2/3 * int(D*mlp_ratio)) for SwiGLU
        h = int(self.hidden_size * self.mlp_ratio)
        return int(2 / 3 * h) if self.use_swiglu else h

    @property
    def label_rows(self):
        return self.num_classes + (1 if self.class_dropout_prob > 0 else 0)


# size registry, models/lightningdit.py:498-538
MODEL_SIZES = {
    "S": dict(depth=12, hidden_size=384, num_heads=6),
    "B": dict(depth=12, hidden_size=768, num_heads=12),
    "L": dict(depth=24, hidden_size=1024, num_heads=16),
    "XL": dict(depth=28, hidden_size=1152, num_heads=16),
    "1p0B": dict(depth=24, hidden_size=1536, num_heads=24),
    "1p6B": dict(depth=28, hidden_size=1792, num_heads=28),
}


def config_from_name(name: str, **kw) -> DiTConfig:
    """'LightningDiT-XL/1' -> DiTConfig (registry at models/lightningdit.py:531-538)."""
    size, patch = name.replace("LightningDiT-", "").split("/")
    return DiTConfig(patch_size=int(patch), **MODEL_SIZES[size], **kw)


# ----------------------------------------------------------------------------- tables
def sincos_pos_embed(D: int, grid: int) -> torch.Tensor:
    """Fixed 2-D sin-cos table, models/lightningdit.py:444-491 (w-coordinate first,
    each half = [sin | cos], float64 angles)."""
    omega = torch.arange(D // 4, dtype=torch.float64) / (D / 4.0)
    omega = 1.0 / 10000 ** omega
    ii, jj = torch.meshgrid(torch.arange(grid, dtype=torch.float64),
                            torch.arange(grid, dtype=torch.float64), indexing="ij")
    # np.meshgrid(grid_w, grid_h): grid[0] varies along columns (j), grid[1] along rows (i)
    out = []
    for pos in (jj.reshape(-1), ii.reshape(-1)):
        ang = pos[:, None] * omega[None, :]
        out.append(torch.cat([ang.sin(), ang.cos()], dim=1))
    return torch.cat(out, dim=1).float().unsqueeze(0)


def rope_tables(head_dim: int, grid: int):
    """VisionRotaryEmbeddingFast buffers, models/pos_embed.py:96-135 with dim = head_dim//2
    (models/lightningdit.py:317-323): (T, head_dim) cos/sin; columns [0,hd/2) use the row
    index, [hd/2,hd) the column index; each frequency repeated twice."""
    dim = head_dim // 2
    freqs = 1.0 / (10000 ** (torch.arange(0, dim, 2)[: dim // 2].float() / dim))
    t = torch.arange(grid) / grid * grid
    f = t[:, None] * freqs[None, :]
    f = f.repeat_interleave(2, dim=-1)                      # (grid, dim)
    full = torch.cat([f[:, None, :].expand(grid, grid, -1), f[None, :, :].expand(grid, grid, -1)], dim=-1)
    full = full.reshape(grid * grid, -1)
    return full.cos(), full.sin()


def init_state_dict(cfg: DiTConfig, seed: int = 0, std: float = 0.02) -> dict:
    """Deterministic random weights with the reference's key set and shapes.
    NOT the reference's init order: weights are an *input* to parity tests.  Nothing is left at
    zero (the reference's adaLN-zero init would make every output 0, lightningdit.py:365-374)."""
    g = torch.Generator().manual_seed(seed)
    D, C, p, T = cfg.hidden_size, cfg.in_channels, cfg.patch_size, cfg.tokens
    S, hd = cfg.mlp_hidden, cfg.head_dim

    def lin(out_f, in_f, wstd=None, bstd=0.02):
        wstd = wstd if wstd is not None else math.sqrt(2.0 / (in_f + out_f))
        return torch.randn(out_f, in_f, generator=g) * wstd, torch.randn(out_f, generator=g) * bstd

    sd = {}
    sd["pos_embed"] = sincos_pos_embed(D, cfg.input_size // cfg.patch_size)
    w, b = lin(D, C * p * p)
    sd["x_embedder.proj.weight"], sd["x_embedder.proj.bias"] = w.reshape(D, C, p, p).contiguous(), b
    sd["t_embedder.mlp.0.weight"], sd["t_embedder.mlp.0.bias"] = lin(D, 256, std)
    sd["t_embedder.mlp.2.weight"], sd["t_embedder.mlp.2.bias"] = lin(D, D, std)
    sd["y_embedder.embedding_table.weight"] = torch.randn(cfg.label_rows, D, generator=g) * std
    if cfg.use_rope:
        sd["feat_rope.freqs_cos"], sd["feat_rope.freqs_sin"] = rope_tables(hd, cfg.input_size // cfg.patch_size)
    nmod = 4 if cfg.wo_shift else 6
    for i in range(cfg.depth):
        pre = f"blocks.{i}."
        if cfg.use_rmsnorm:
            sd[pre + "norm1.weight"] = 1.0 + 0.1 * torch.randn(D, generator=g)
            sd[pre + "norm2.weight"] = 1.0 + 0.1 * torch.randn(D, generator=g)
        sd[pre + "attn.qkv.weight"], sd[pre + "attn.qkv.bias"] = lin(3 * D, D)
        if cfg.use_qknorm:
            # LayerNorm(hd) has weight+bias, RMSNorm(hd) only weight (lightningdit.py:56-61)
            sd[pre + "attn.q_norm.weight"] = 1.0 + 0.1 * torch.randn(hd, generator=g)
            sd[pre + "attn.k_norm.weight"] = 1.0 + 0.1 * torch.randn(hd, generator=g)
            if not cfg.use_rmsnorm:
                sd[pre + "attn.q_norm.bias"] = 0.1 * torch.randn(hd, generator=g)
                sd[pre + "attn.k_norm.bias"] = 0.1 * torch.randn(hd, generator=g)
        sd[pre + "attn.proj.weight"], sd[pre + "attn.proj.bias"] = lin(D, D)
        if cfg.use_swiglu:
            sd[pre + "mlp.w12.weight"], sd[pre + "mlp.w12.bias"] = lin(2 * S, D)
            sd[pre + "mlp.w3.weight"], sd[pre + "mlp.w3.bias"] = lin(D, S)
        else:
            sd[pre + "mlp.fc1.weight"], sd[pre + "mlp.fc1.bias"] = lin(S, D)
            sd[pre + "mlp.fc2.weight"], sd[pre + "mlp.fc2.bias"] = lin(D, S)
        sd[pre + "adaLN_modulation.1.weight"], sd[pre + "adaLN_modulation.1.bias"] = lin(nmod * D, D, std)
    if cfg.use_rmsnorm:
        sd["final_layer.norm_final.weight"] = 1.0 + 0.1 * torch.randn(D, generator=g)
    sd["final_layer.linear.weight"], sd["final_layer.linear.bias"] = lin(p * p * cfg.out_channels, D, std)
    sd["final_layer.adaLN_modulation.1.weight"], sd["final_layer.adaLN_modulation.1.bias"] = lin(2 * D, D, std)
    return sd


# ----------------------------------------------------------------------------- primitives
# Optional emulation of the engine's bf16 rounding points (GEMM / attention operands and the tensors the
# engine stores in bf16); everything else stays fp32 exactly like the engine.  Used only to *explain* the
# bf16-mode error budget in tests; parity bars are always measured against the plain fp32 path above.
BF16_EMULATION = False


def _q(x):
    return x.bfloat16().float() if BF16_EMULATION else x


def _lin(x, w, b):
    return F.linear(_q(x), _q(w), b)


def rms_norm(x, w, eps=1e-6):
    """models/rmsnorm.py:63,76-77 -- fp32 statistics, learnable weight."""
    xf = x.float()
    return (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)).type_as(x) * w


def layer_norm_noaffine(x, eps=1e-6):
    """nn.LayerNorm(elementwise_affine=False, eps=1e-6), models/lightningdit.py:196-197."""
    return F.layer_norm(x, x.shape[-1:], None, None, eps)


def modulate(x, shift, scale):
    """models/lightningdit.py:26-30."""
    y = x * (1 + scale.unsqueeze(1))
    return y if shift is None else y + shift.unsqueeze(1)


def rotate_half(x):
    """models/pos_embed.py:38-42 -- adjacent pairs (a,b) -> (-b,a)."""
    a, b = x[..., 0::2], x[..., 1::2]
    return torch.stack((-b, a), dim=-1).flatten(-2)


def timestep_embedding(t, dim=256, max_period=10000):
    """models/lightningdit.py:108-131 -- [cos | sin], t unscaled."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32) / half)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def _norm_block(cfg, sd, key, x):
    if cfg.use_rmsnorm:
        return rms_norm(x, sd[key + ".weight"])
    return layer_norm_noaffine(x)


def _qk_norm(cfg, sd, key, x):
    if cfg.use_rmsnorm:
        return rms_norm(x, sd[key + ".weight"])
    return F.layer_norm(x, x.shape[-1:], sd[key + ".weight"], sd[key + ".bias"], 1e-5)


def attention(cfg, sd, pre, x):
    """models/lightningdit.py:66-91."""
    B, T, D = x.shape
    H, hd = cfg.num_heads, cfg.head_dim
    qkv = _lin(x, sd[pre + "qkv.weight"], sd[pre + "qkv.bias"])
    qkv = qkv.reshape(B, T, 3, H, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    if cfg.use_qknorm:
        q, k = _qk_norm(cfg, sd, pre + "q_norm", q), _qk_norm(cfg, sd, pre + "k_norm", k)
    if cfg.use_rope:
        cos, sin = sd["feat_rope.freqs_cos"], sd["feat_rope.freqs_sin"]
        q = q * cos + rotate_half(q) * sin
        k = k * cos + rotate_half(k) * sin
    q, k, v = _q(q), _q(k), _q(v)
    att = (q * hd ** -0.5) @ k.transpose(-2, -1)      # manual branch :82-86 == SDPA :77-80
    att = att.softmax(dim=-1)
    o = (_q(att) @ v).transpose(1, 2).reshape(B, T, D)
    return _lin(o, sd[pre + "proj.weight"], sd[pre + "proj.bias"])


def mlp(cfg, sd, pre, x):
    if cfg.use_swiglu:
        # models/swiglu_ffn.py:31-36 -- x1 (first S columns) is the SiLU-gated half
        x12 = _lin(x, sd[pre + "w12.weight"], sd[pre + "w12.bias"])
        x1, x2 = x12.chunk(2, dim=-1)
        return _lin(F.silu(x1) * x2, sd[pre + "w3.weight"], sd[pre + "w3.bias"])
    h = F.gelu(_lin(x, sd[pre + "fc1.weight"], sd[pre + "fc1.bias"]), approximate="tanh")
    return _lin(h, sd[pre + "fc2.weight"], sd[pre + "fc2.bias"])


def block(cfg, sd, i, x, c):
    """models/lightningdit.py:239-250."""
    pre = f"blocks.{i}."
    m = _lin(F.silu(c), sd[pre + "adaLN_modulation.1.weight"], sd[pre + "adaLN_modulation.1.bias"])
    if cfg.wo_shift:
        scale_msa, gate_msa, scale_mlp, gate_mlp = m.chunk(4, dim=1)
        shift_msa = shift_mlp = None
    else:
        shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp = m.chunk(6, dim=1)
    x = x + gate_msa.unsqueeze(1) * attention(cfg, sd, pre + "attn.", modulate(_norm_block(cfg, sd, pre + "norm1", x), shift_msa, scale_msa))
    x = x + gate_mlp.unsqueeze(1) * mlp(cfg, sd, pre + "mlp.", modulate(_norm_block(cfg, sd, pre + "norm2", x), shift_mlp, scale_mlp))
    return x


def unpatchify(cfg, x):
    """models/lightningdit.py:376-389."""
    c, p = cfg.out_channels, cfg.patch_size
    h = w = int(x.shape[1] ** 0.5)
    x = x.reshape(x.shape[0], h, w, p, p, c)
    x = torch.einsum("nhwpqc->nchpwq", x)
    return x.reshape(x.shape[0], c, h * p, w * p)


@torch.no_grad()
def forward(cfg: DiTConfig, sd: dict, x, t, y):
    """LightningDiT.forward, models/lightningdit.py:391-418 (eval mode: no label dropout :163-169)."""
    p = cfg.patch_size
    tok = F.conv2d(x, sd["x_embedder.proj.weight"], sd["x_embedder.proj.bias"], stride=p)
    tok = tok.flatten(2).transpose(1, 2) + sd["pos_embed"]
    temb = timestep_embedding(t)
    temb = F.linear(F.silu(F.linear(temb, sd["t_embedder.mlp.0.weight"], sd["t_embedder.mlp.0.bias"])),
                    sd["t_embedder.mlp.2.weight"], sd["t_embedder.mlp.2.bias"])
    c = temb + sd["y_embedder.embedding_table.weight"][y]
    for i in range(cfg.depth):
        tok = block(cfg, sd, i, tok, c)
    # FinalLayer, models/lightningdit.py:267-272 (shift first)
    shift, scale = _lin(F.silu(c), sd["final_layer.adaLN_modulation.1.weight"],
                        sd["final_layer.adaLN_modulation.1.bias"]).chunk(2, dim=1)
    tok = modulate(_norm_block(cfg, sd, "final_layer.norm_final", tok), shift, scale)
    tok = _lin(tok, sd["final_layer.linear.weight"], sd["final_layer.linear.bias"])
    out = unpatchify(cfg, tok)
    if cfg.learn_sigma:
        out, _ = out.chunk(2, dim=1)
    return out


@torch.no_grad()
def forward_with_cfg(cfg, sd, x, t, y, cfg_scale, cfg_interval=None, cfg_interval_start=None, cfg_channels=3):
    """models/lightningdit.py:420-442.  Guidance on the first `cfg_channels` (=3) channels only;
    with cfg_interval True the guided value is replaced by the conditional one when t[0] < start."""
    half = x[: len(x) // 2]
    out = forward(cfg, sd, torch.cat([half, half], dim=0), t, y)
    eps, rest = out[:, :cfg_channels], out[:, cfg_channels:]
    cond, uncond = torch.split(eps, len(eps) // 2, dim=0)
    half_eps = uncond + cfg_scale * (cond - uncond)
    if cfg_interval is True and bool(t[0] < cfg_interval_start):   # fp32 tensor vs python scalar, as in :437-439
        half_eps = cond
    return torch.cat([torch.cat([half_eps, half_eps], dim=0), rest], dim=1)


def time_grid(num_steps: int, timestep_shift: float = 0.0, t0: float = 0.0, t1: float = 1.0) -> torch.Tensor:
    """transport/integrators.py:94-101: fp32 linspace, optional shift s*t/(1+(s-1)*t) evaluated
    element-wise on 0-dim fp32 tensors.  (t0,t1)=(0,1) for Linear+velocity, transport/transport.py:84-111."""
    t = torch.linspace(t0, t1, num_steps)
    if timestep_shift > 0:
        t = torch.tensor([(timestep_shift * tn) / (1 + (timestep_shift - 1) * tn) for tn in t])
    return t


@torch.no_grad()
def sample_ode(cfg, sd, z, y, grid, method="euler", cfg_scale=None, cfg_interval=None,
               cfg_interval_start=None, return_trace=False, atol=1e-6, rtol=1e-3):
    """ode.sample (transport/integrators.py:107-126) with the fixed-grid torchdiffeq semantics:
    one step per grid interval; euler y1 = y0 + dt*f(t0,y0); heun (explicit trapezoid)
    y1 = y0 + dt/2*(k1 + f(t1, y0 + dt*k1)).  The drift is the identity velocity wrapper
    (transport/transport.py:234-248).  z has 2n rows under CFG (inference.py:205-211).
    Returns the final state, and (optionally) every evaluation's velocity in call order."""
    def f(t, x):
        tv = torch.ones(x.shape[0]) * t
        if cfg_scale is None:
            return forward(cfg, sd, x, tv, y)
        return forward_with_cfg(cfg, sd, x, tv, y, cfg_scale, cfg_interval, cfg_interval_start)

    if method == "dopri5":
        from oracle.dopri5 import odeint_dopri5
        return odeint_dopri5(lambda t, x: f(t, x), z, grid, atol=atol, rtol=rtol)[-1]
    x, trace = z, []
    for j in range(len(grid) - 1):
        t0, t1 = grid[j], grid[j + 1]
        dt = t1 - t0
        k1 = f(t0, x)
        trace.append(k1)
        if method == "euler":
            x = x + dt * k1
        elif method in ("heun", "heun2"):
            k2 = f(t1, x + dt * k1)
            trace.append(k2)
            x = x + (dt * 0.5) * (k1 + k2)
        else:
            raise ValueError(method)
    return (x, trace) if return_trace else x


def flops_per_forward(cfg: DiTConfig) -> float:
    """Algorithmic FLOPs of one sample-forward (BASELINE.md section 3 counting rule)."""
    D, T, S, L = cfg.hidden_size, cfg.tokens, cfg.mlp_hidden, cfg.depth
    C, p = cfg.in_channels, cfg.patch_size
    nmod = 4 if cfg.wo_shift else 6
    mlp_f = (2 * D * 2 * S + 2 * S * D) if cfg.use_swiglu else (2 * D * S + 2 * S * D)
    per_tok = 2 * D * 3 * D + 4 * T * D + 2 * D * D + mlp_f
    return (L * T * per_tok + L * 2 * D * nmod * D + 2 * D * 2 * D + (2 * 256 * D + 2 * D * D)
            + 2 * (C * p * p) * D * T + 2 * D * (p * p * cfg.out_channels) * T)

"""TEST INFRASTRUCTURE ONLY -- import the UNMODIFIED reference under sys.modules shims.

Only usable in the build container, where /root/reference exists; the GPU box has
no /root/reference, so nothing at run time (tests -m gpu, smoke(), bench.py) may
import this module.  It is used by oracle/gen_golden.py (to pin the CPU oracle
against the real reference and to emit tests/golden/*.pt) and by the container-only
test tests/test_oracle_vs_reference.py.

The three shims replace packages that are absent from the image and cannot be
installed (no network); see SURVEY.md section 8(c):
  * timm.models.vision_transformer.{PatchEmbed, Mlp}   (LightningDiT/models/lightningdit.py:21)
  * fairscale.nn.model_parallel.*                       (LightningDiT/models/rmsnorm.py:8-15, unused by RMSNorm)
  * torchdiffeq.odeint                                  (LightningDiT/transport/integrators.py:4,118-125)
The torchdiffeq shim implements the published fixed-grid semantics (one step per grid
interval, solution[j] == y_j) for 'euler' and the 2-stage explicit trapezoid 'heun'
(same arithmetic as the deterministic part of LightningDiT/transport/integrators.py:45-48).
"""
import os
import sys
import types

os.environ.setdefault("TORCH_COMPILE_DISABLE", "1")

import torch
import torch.nn as nn

REFERENCE_ROOT = "/root/reference/LightningDiT"


def reference_available() -> bool:
    return os.path.isdir(REFERENCE_ROOT)


class _PatchEmbed(nn.Module):
    """timm==1.0.12 PatchEmbed as used at lightningdit.py:309 (strided conv + flatten)."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768, bias=True, **_):
        super().__init__()
        self.img_size = (img_size, img_size)
        self.patch_size = (patch_size, patch_size)
        self.grid_size = (img_size // patch_size, img_size // patch_size)
        self.num_patches = self.grid_size[0] * self.grid_size[1]
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size, bias=bias)

    def forward(self, x):
        return self.proj(x).flatten(2).transpose(1, 2)


class _Mlp(nn.Module):
    """timm Mlp (fc1 -> act -> fc2); only reached when use_swiglu=False (lightningdit.py:219-224)."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.0, **_):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features)

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


def fixed_grid_odeint(func, y0, t, method="euler", atol=None, rtol=None, **_):
    """Fixed-grid euler / heun over the grid ``t``; returns stack([y_0, ..., y_{P-1}])."""
    if method == "dopri5":
        from oracle.dopri5 import odeint_dopri5      # restated adaptive solver (parity unpinned, see its header)
        a = atol[0] if isinstance(atol, (list, tuple)) else atol
        r = rtol[0] if isinstance(rtol, (list, tuple)) else rtol
        return odeint_dopri5(func, y0, t, atol=a, rtol=r)
    if method not in ("euler", "heun", "heun2"):
        raise ValueError(f"shim supports euler/heun/dopri5 only, got {method!r}")
    ys = [y0]
    y = y0
    for j in range(len(t) - 1):
        t0, t1 = t[j], t[j + 1]
        dt = t1 - t0
        k1 = func(t0, y)
        if method == "euler":
            y = y + dt * k1
        else:
            k2 = func(t1, y + dt * k1)
            y = y + (dt * 0.5) * (k1 + k2)
        ys.append(y)
    return torch.stack(ys)


def install_shims():
    if "timm" not in sys.modules:
        timm = types.ModuleType("timm")
        timm_models = types.ModuleType("timm.models")
        vt = types.ModuleType("timm.models.vision_transformer")
        vt.PatchEmbed, vt.Mlp = _PatchEmbed, _Mlp
        timm.models, timm_models.vision_transformer = timm_models, vt
        sys.modules.update({"timm": timm, "timm.models": timm_models, "timm.models.vision_transformer": vt})
    if "fairscale" not in sys.modules:
        fs = types.ModuleType("fairscale")
        fs_nn = types.ModuleType("fairscale.nn")
        mp = types.ModuleType("fairscale.nn.model_parallel")
        init = types.ModuleType("fairscale.nn.model_parallel.initialize")
        layers = types.ModuleType("fairscale.nn.model_parallel.layers")
        init.get_model_parallel_world_size = lambda: 1
        init.get_model_parallel_rank = lambda: 0
        init.initialize_model_parallel = lambda *a, **k: None
        init.model_parallel_is_initialized = lambda: True
        for nm in ("ColumnParallelLinear", "ParallelEmbedding", "RowParallelLinear"):
            setattr(layers, nm, type(nm, (nn.Module,), {}))
        fs.nn, fs_nn.model_parallel, mp.initialize, mp.layers = fs_nn, mp, init, layers
        sys.modules.update({"fairscale": fs, "fairscale.nn": fs_nn, "fairscale.nn.model_parallel": mp,
                            "fairscale.nn.model_parallel.initialize": init,
                            "fairscale.nn.model_parallel.layers": layers})
    if "torchdiffeq" not in sys.modules:
        tde = types.ModuleType("torchdiffeq")
        tde.odeint = fixed_grid_odeint
        sys.modules["torchdiffeq"] = tde


def import_reference():
    """Returns (models.lightningdit module, transport package) of the unmodified reference."""
    if not reference_available():
        raise RuntimeError("/root/reference is not present (only the build container has it)")
    install_shims()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import importlib
    ld = importlib.import_module("models.lightningdit")
    tr = importlib.import_module("transport")
    return ld, tr


def rerandomize_zero_params(model, seed=0, std=0.02):
    """Fresh reference models output exactly 0 (adaLN-zero, lightningdit.py:365-374): re-draw every
    all-zero parameter N(0, std) in named_parameters() order so parity tests are not vacuous."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for _, p in model.named_parameters():
            if p.abs().sum() == 0:
                p.copy_(torch.randn(p.shape, generator=g) * std)
    return model

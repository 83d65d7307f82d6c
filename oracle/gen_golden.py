"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.pt from the UNMODIFIED reference.

Run in the build container (needs /root/reference):   python oracle/gen_golden.py
Each fixture holds the reference state_dict (tiny model), seeded inputs and the
reference's outputs, so the GPU box (which has no /root/reference) can check both the
CPU oracle and the CUDA engine against numbers the real reference produced.
"""
import os
import sys

os.environ["TORCH_COMPILE_DISABLE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from oracle import ref_shim

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# name -> ctor kwargs of the reference LightningDiT (models/lightningdit.py:279-297)
PROD = dict(use_qknorm=False, use_swiglu=True, use_rope=True, use_rmsnorm=True, wo_shift=False)
CASES = {
    # production flags, head_dim 32, 16 tokens
    "tiny_prod": dict(input_size=4, patch_size=1, in_channels=32, hidden_size=64, depth=2, num_heads=2,
                      num_classes=10, **PROD),
    # head_dim 72 (the XL head size) + qk-norm
    "tiny_hd72_qknorm": dict(input_size=4, patch_size=1, in_channels=32, hidden_size=144, depth=2, num_heads=2,
                             num_classes=10, **{**PROD, "use_qknorm": True}),
    # patch 2 (64 tokens from a 16x16 latent), no-shift adaLN
    "tiny_patch2_woshift": dict(input_size=16, patch_size=2, in_channels=32, hidden_size=128, depth=2, num_heads=2,
                                num_classes=31, **{**PROD, "wo_shift": True}),
    # 256 tokens (the benchmark token count), head_dim 64
    "tiny_t256": dict(input_size=16, patch_size=1, in_channels=32, hidden_size=128, depth=1, num_heads=2,
                      num_classes=31, **PROD),
    # legacy flags: LayerNorm, GELU MLP, no RoPE, learn_sigma
    "tiny_legacy": dict(input_size=4, patch_size=1, in_channels=4, hidden_size=64, depth=2, num_heads=2,
                        num_classes=10, use_qknorm=False, use_swiglu=False, use_rope=False, use_rmsnorm=False,
                        wo_shift=False, learn_sigma=True),
}


def main():
    ld, tr = ref_shim.import_reference()
    os.makedirs(OUT, exist_ok=True)
    for name, kw in CASES.items():
        torch.manual_seed(0)
        model = ld.LightningDiT(**kw).eval()
        ref_shim.rerandomize_zero_params(model, seed=0)
        # make norm weights / biases non-trivial too
        g = torch.Generator().manual_seed(7)
        with torch.no_grad():
            for n_, p_ in model.named_parameters():
                if n_.endswith("bias") or "norm" in n_:
                    p_.add_(0.05 * torch.randn(p_.shape, generator=g))
        n = 3
        C, hw, K = kw["in_channels"], kw["input_size"], kw["num_classes"]
        x = torch.randn(2 * n, C, hw, hw, generator=torch.Generator().manual_seed(1234))
        y = torch.randint(0, K, (n,), generator=torch.Generator().manual_seed(4321))
        y2 = torch.cat([y, torch.full((n,), K)])
        t = torch.rand(2 * n, generator=torch.Generator().manual_seed(99))
        fix = {"ctor": kw, "state_dict": {k: v.clone() for k, v in model.state_dict().items()},
               "x": x, "t": t, "y": y2}
        with torch.no_grad():
            fix["forward"] = model(x, t, y2)
            tu = torch.full((2 * n,), 0.37)
            fix["t_uniform"] = tu
            fix["cfg_guided"] = model.forward_with_cfg(x, tu, y2, 4.0, True, 0.11)
            fix["cfg_gated_off"] = model.forward_with_cfg(x, torch.full((2 * n,), 0.05), y2, 4.0, True, 0.11)
            fix["cfg_no_interval"] = model.forward_with_cfg(x, torch.full((2 * n,), 0.05), y2, 4.0)
            # sampler through the reference's own transport API (transport/transport.py:398-443)
            transport = tr.create_transport("Linear", "velocity", None, None, None)
            sampler = tr.Sampler(transport)
            z = torch.cat([x[:n], x[:n]], 0)
            for method, steps, shift in (("euler", 6, 0.3), ("heun", 4, 0.1), ("euler", 5, 0.0)):
                fn = sampler.sample_ode(sampling_method=method, num_steps=steps, atol=1e-6, rtol=1e-3,
                                        reverse=False, timestep_shift=shift)
                traj = fn(z, model.forward_with_cfg, y=y2, cfg_scale=4.0, cfg_interval=True, cfg_interval_start=0.11)
                fix[f"sample_{method}_{steps}_{shift}"] = traj[-1].clone()
                fix[f"grid_{method}_{steps}_{shift}"] = fn.__self__.t.clone()
            fn = sampler.sample_ode(sampling_method="euler", num_steps=4, atol=1e-6, rtol=1e-3, reverse=False,
                                    timestep_shift=0.3)
            fix["sample_nocfg_euler_4_0.3"] = fn(x, model.forward, y=y2)[-1].clone()
        path = os.path.join(OUT, name + ".pt")
        torch.save(fix, path)
        print(name, "->", path, f"{os.path.getsize(path) / 1e6:.2f} MB", "forward absmax", float(fix['forward'].abs().max()))


if __name__ == "__main__":
    main()

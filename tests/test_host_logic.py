"""Host-side mirror of the reference interface (no GPU): state_dict contract, registry, transport rules,
time grids, the generic fixed-grid loop, sharding helpers."""
import os

import pytest
import torch

import va_vae_b200 as V
from conftest import GOLDEN_CASES, ROOT, rel_l2
from oracle import lightningdit_oracle as O


def test_state_dict_contract_matches_reference_fixtures(golden):
    import va_vae_b200 as V
    for name in GOLDEN_CASES:
        fix = golden(name)
        m = V.LightningDiT(**fix["ctor"])
        mine, ref = m.state_dict(), fix["state_dict"]
        assert set(mine) == set(ref), (name, set(mine) ^ set(ref))
        for k in ref:
            assert tuple(mine[k].shape) == tuple(ref[k].shape), (name, k)
        # fixed tables are bit-identical to the reference's buffers
        assert torch.equal(mine["pos_embed"], ref["pos_embed"])
        if fix["ctor"]["use_rope"]:
            assert torch.equal(mine["feat_rope.freqs_cos"], ref["feat_rope.freqs_cos"])
            assert torch.equal(mine["feat_rope.freqs_sin"], ref["feat_rope.freqs_sin"])
        res = m.load_state_dict(ref, strict=True)
        assert not res.missing_keys and not res.unexpected_keys
        assert m.in_channels == fix["ctor"]["in_channels"]


def test_fresh_model_is_adaln_zero_like_the_reference():
    import va_vae_b200 as V
    m = V.LightningDiT_models["LightningDiT-S/2"](input_size=16, num_classes=31, use_swiglu=True, use_rope=True,
                                                  use_rmsnorm=True)
    sd = m.state_dict()
    assert sd["final_layer.linear.weight"].abs().sum() == 0        # lightningdit.py:371-374
    assert sd["blocks.0.adaLN_modulation.1.weight"].abs().sum() == 0
    assert sd["blocks.3.attn.qkv.bias"].abs().sum() == 0
    assert sd["blocks.3.attn.qkv.weight"].abs().sum() > 0
    assert sd["pos_embed"].shape == (1, 64, 384)


def test_registry_names_and_sizes():
    import va_vae_b200 as V
    assert set(V.LightningDiT_models) == {
        "LightningDiT-S/1", "LightningDiT-S/2", "LightningDiT-B/1", "LightningDiT-B/2", "LightningDiT-L/2",
        "LightningDiT-XL/1", "LightningDiT-XL/2", "LightningDiT-1p0B/1", "LightningDiT-1p0B/2", "LightningDiT-1p6B/1",
        "LightningDiT-1p6B/2"}
    for name in ("LightningDiT-S/1", "LightningDiT-B/2"):
        cfg = O.config_from_name(name)
        m = V.LightningDiT_models[name](input_size=16, use_swiglu=True)
        assert (m.depth, m.hidden_size, m.num_heads, m.patch_size, m.mlp_hidden) == \
            (cfg.depth, cfg.hidden_size, cfg.num_heads, cfg.patch_size, cfg.mlp_hidden)


def test_transport_rules_and_time_grid(golden):
    import va_vae_b200 as V
    t = V.create_transport("Linear", "velocity", None, None, None)
    assert (t.train_eps, t.sample_eps) == (0, 0)
    assert t.check_interval(t.train_eps, t.sample_eps, sde=False, eval=True) == (0, 1)
    t2 = V.create_transport("Linear", "noise")
    assert t2.train_eps == 1e-3 and t2.check_interval(t2.train_eps, t2.sample_eps, eval=True) == (1e-3, 1 - 1e-3)
    with pytest.raises(NotImplementedError):
        V.Sampler(t2)
    fix = golden("tiny_prod")
    s = V.Sampler(t)
    for method, steps, shift in (("euler", 6, 0.3), ("heun", 4, 0.1), ("euler", 5, 0.0)):
        fn = s.sample_ode(sampling_method=method, num_steps=steps, atol=1e-6, rtol=1e-3, reverse=False,
                          timestep_shift=shift)
        assert torch.equal(fn.__self__.t, fix[f"grid_{method}_{steps}_{shift}"])     # bit-exact fp32 grid
    with pytest.raises(AssertionError):
        V.transport.ode(None, t0=1, t1=0, sampler_type="euler", num_steps=4, atol=0, rtol=0, timestep_shift=0)


@pytest.mark.parametrize("method", ["euler", "heun"])
def test_generic_fixed_grid_loop_matches_oracle_semantics(golden, method):
    """Any callable (here: the CPU oracle as `model`) goes through the generic loop; the result must equal the
    oracle's own sampler and the reference fixture."""
    import va_vae_b200 as V
    fix = golden("tiny_prod")
    cfg, sd = O.DiTConfig(**fix["ctor"]), fix["state_dict"]
    steps, shift = (6, 0.3) if method == "euler" else (4, 0.1)
    fn = V.Sampler(V.create_transport()).sample_ode(sampling_method=method, num_steps=steps, atol=1e-6, rtol=1e-3,
                                                   reverse=False, timestep_shift=shift)
    n = fix["x"].shape[0] // 2
    z = torch.cat([fix["x"][:n], fix["x"][:n]], 0)

    def model(x, t, **kw):
        return O.forward_with_cfg(cfg, sd, x, t, **kw)
    traj = fn(z, model, y=fix["y"], cfg_scale=4.0, cfg_interval=True, cfg_interval_start=0.11)
    assert traj.shape == (steps,) + tuple(z.shape)
    assert rel_l2(traj[-1], fix[f"sample_{method}_{steps}_{shift}"]) < 5e-6
    assert torch.equal(traj[0], z)


def test_lazy_trajectory_contract():
    from va_vae_b200.transport import LazyTrajectory
    final = torch.arange(24.).reshape(4, 6)
    tr = LazyTrajectory(final, 10)
    assert len(tr) == 10 and tr.shape == (10, 4, 6)
    assert torch.equal(tr[-1], final) and torch.equal(tr[9], final)
    a, b = tr[-1].chunk(2, dim=0)      # inference.py:216-218
    assert a.shape == (2, 6)
    with pytest.raises(IndexError):
        tr[3]


def test_shard_ranges_cover_everything():
    from va_vae_b200.dist import shard_range, build_cfg_batch
    for n in (0, 1, 7, 8, 1024, 1023):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    z, y = torch.randn(3, 2, 4, 4), torch.tensor([1, 2, 3])
    zz, yy = build_cfg_batch(z, y, 31)
    assert torch.equal(zz[:3], zz[3:]) and yy.tolist() == [1, 2, 3, 31, 31, 31]


def test_engine_label_and_shape_validation_on_the_host():
    """Engine._labels: host labels are range-checked like nn.Embedding (lightningdit.py:168) and every entry point
    insists on y of shape (batch,) -- no GPU needed for the checks themselves."""
    import pytest
    from va_vae_b200.engine import Engine
    eng = Engine.__new__(Engine)
    eng.config = dict(label_rows=32)
    eng.device = torch.device("cpu")
    assert eng._labels(torch.tensor([0, 31], dtype=torch.int32), 2).dtype == torch.int64
    for bad in (torch.tensor([0, 32]), torch.tensor([-1, 3])):
        with pytest.raises(IndexError, match="index out of range"):
            eng._labels(bad, 2)
    with pytest.raises(ValueError):
        eng._labels(torch.tensor([1]), 2)
    with pytest.raises(ValueError):
        eng._labels(torch.zeros(2, 1, dtype=torch.long), 2)
    with pytest.raises(RuntimeError):
        eng._labels(torch.tensor([0.0, 1.0]), 2)


def test_engine_is_invalidated_by_weight_updates_that_bypass_the_version_counter():
    m = V.LightningDiT(input_size=4, patch_size=1, in_channels=4, hidden_size=32, depth=1, num_heads=2, num_classes=3,
                       use_swiglu=True, use_rope=True, use_rmsnorm=True)
    key = m._weights_key()
    m._engine_key = key
    m.load_state_dict(m.state_dict())
    assert m._engine_key is None                      # load_state_dict invalidates
    key = m._weights_key()
    m._engine_key = key
    m.blocks[0].attn.qkv.weight.data.mul_(0.5)        # invisible to ._version ...
    assert m._weights_key() == key
    m.invalidate_engine()                             # ... hence the explicit hook
    assert m._engine_key is None
    m._engine_key = key
    m.feat_rope.freqs_cos.add_(1.0)                   # buffers are part of the key
    assert m._weights_key() != key
    m._engine_key = m._weights_key()
    m.float()                                         # _apply invalidates
    assert m._engine_key is None


def test_bench_arms_describe_the_same_workload(tmp_path):
    """Both arms of bench.py print the same config.workload string (the driver compares them), the default workload
    is the configuration BASELINE.json's metric is quoted on, and the reference arm runs all host threads."""
    import json
    import subprocess
    import sys
    import bench
    sys_argv = sys.argv
    try:
        sys.argv = ["bench.py"]
        a = bench.parse()
    finally:
        sys.argv = sys_argv
    assert a.workload == "xl1" and a.points == 250 and a.method == "heun"
    desc = bench.workload_desc(a, 1)
    assert "LightningDiT-XL/1" in desc and "n=128 per GPU" in desc and "498 evals" in desc
    env = dict(os.environ, OMP_NUM_THREADS="1")        # what torchrun exports to every rank
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "s1",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    sys.argv = ["bench.py", "--workload", "s1"]
    try:
        a = bench.parse()
    finally:
        sys.argv = sys_argv
    assert line["impl"] == "reference" and line["config"]["workload"] == bench.workload_desc(a, 1)
    assert line["cpu_baseline"]["cores"] == (len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count())
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["metric"] == bench.METRIC

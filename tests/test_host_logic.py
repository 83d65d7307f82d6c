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


# ---- tile selection of the layer GEMMs (vavae_gemm_plan: host-only view of launch_gemm_tc's decisions) ----------------
def _ctas(M, N, bn, bn_last):
    cols = (N - bn_last) // bn + 1 if bn_last != bn else -(-N // bn)
    return -(-M // 128) * cols


def test_large_batch_layer_gemms_use_wide_tiles_and_the_tma_epilogue_where_it_measured_faster():
    from va_vae_b200 import engine as E
    M, T, D, S = 65536, 256, 1152, 3072                        # XL/1, 256 model rows
    qkv = E.gemm_plan(E.EPI_QKV, M, 3 * D, D, T, H=16, hd=72)
    assert (qkv["bn"], qkv["bn_last"]) == (256, 128)           # 3456 = 13 x 256 + 128
    w12 = E.gemm_plan(E.EPI_SWIGLU, M, 2 * S, D, T)
    assert (w12["bn"], w12["bn_last"]) == (256, 256)
    proj = E.gemm_plan(E.EPI_GATED_RESIDUAL, M, D, D, T)
    assert (proj["bn"], proj["bn_last"]) == (256, 128) and proj["tma"] and proj["warps"] == 8 and proj["coarse"]
    w3 = E.gemm_plan(E.EPI_GATED_RESIDUAL, M, D, S, T)
    assert w3["tma"] and w3["warps"] == 4 and w3["coarse"]     # long K: four epilogue warps keep the fifth operand stage
    b1_w3 = E.gemm_plan(E.EPI_GATED_RESIDUAL, 131072, 768, 2048, T)
    assert b1_w3["tma"] and b1_w3["warps"] == 4 and b1_w3["coarse"] and b1_w3["bn"] == 256
    # partial tile pairs (M % 256 != 0): the TMA epilogue needs guard-free geometry
    assert not E.gemm_plan(E.EPI_GATED_RESIDUAL, 65536 + 128, D, D, T)["tma"]


def test_small_batch_layer_gemms_use_narrow_tiles_within_one_wave():
    from va_vae_b200 import engine as E
    T, D, S = 256, 1152, 3072
    for batch in (1, 2, 4, 8):
        M = batch * T
        for kind, N, K, kw in ((E.EPI_QKV, 3 * D, D, dict(H=16, hd=72)), (E.EPI_SWIGLU, 2 * S, D, {}),
                               (E.EPI_GATED_RESIDUAL, D, D, {}), (E.EPI_GATED_RESIDUAL, D, S, {})):
            p = E.gemm_plan(kind, M, N, K, T, **kw)
            wide = _ctas(M, N, 256, 128 if N % 256 else 256)
            assert p["bn"] % 32 == 0 and p["bn_last"] % 32 == 0
            if p["bn"] != 256:                                  # narrowed: divides N, fits one wave, more CTAs than before
                assert N % p["bn"] == 0 and p["bn_last"] == p["bn"]
                assert wide < _ctas(M, N, p["bn"], p["bn"]) <= 148
            if kind == E.EPI_GATED_RESIDUAL:                    # producers keep whole 128-column groups -> coarse slots
                assert p["bn"] % 128 == 0 and p["coarse"]
    assert E.gemm_plan(E.EPI_QKV, 256, 3 * D, D, T, H=16, hd=72)["bn"] == 64        # 108 CTAs
    assert E.gemm_plan(E.EPI_SWIGLU, 256, 2 * S, D, T)["bn"] == 96                  # 128 CTAs
    assert E.gemm_plan(E.EPI_GATED_RESIDUAL, 256, D, S, T)["bn"] == 128             # 18 CTAs, coarse slots


def test_partial_sum_layout_follows_the_tile_width_not_the_batch():
    """Coarse slots exactly when every tile holds whole 128-column groups; an odd hidden size gets fine slots at any M."""
    from va_vae_b200 import engine as E
    for M in (256, 4096, 65536):
        for N in (384, 768, 1024, 1152):
            p = E.gemm_plan(E.EPI_GATED_RESIDUAL, M, N, N, 256)
            assert p["coarse"] == (p["bn"] % 128 == 0 and p["bn_last"] % 128 == 0) and p["coarse"]
        p = E.gemm_plan(E.EPI_GATED_RESIDUAL, M, 96, 384, 32)
        assert not p["coarse"] and p["bn"] == 96

"""The CPU oracle (oracle/lightningdit_oracle.py) against fixtures produced by the UNMODIFIED
reference (oracle/gen_golden.py).  fp32 on CPU both sides -> tight tolerance."""
import pytest
import torch

from conftest import GOLDEN_CASES, rel_l2
from oracle import lightningdit_oracle as O

TOL = 2e-6   # same fp32 math, different op order only


def _cfg(fix):
    kw = dict(fix["ctor"])
    return O.DiTConfig(**kw)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_forward_matches_reference(golden, name):
    fix = golden(name)
    cfg = _cfg(fix)
    out = O.forward(cfg, fix["state_dict"], fix["x"], fix["t"], fix["y"])
    assert out.shape == fix["forward"].shape
    assert rel_l2(out, fix["forward"]) < TOL


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_forward_with_cfg_matches_reference(golden, name):
    fix = golden(name)
    cfg, sd = _cfg(fix), fix["state_dict"]
    n2 = fix["x"].shape[0]
    out = O.forward_with_cfg(cfg, sd, fix["x"], fix["t_uniform"], fix["y"], 4.0, True, 0.11)
    assert rel_l2(out, fix["cfg_guided"]) < TOL
    out = O.forward_with_cfg(cfg, sd, fix["x"], torch.full((n2,), 0.05), fix["y"], 4.0, True, 0.11)
    assert rel_l2(out, fix["cfg_gated_off"]) < TOL
    out = O.forward_with_cfg(cfg, sd, fix["x"], torch.full((n2,), 0.05), fix["y"], 4.0)
    assert rel_l2(out, fix["cfg_no_interval"]) < TOL
    # the 3-channel quirk (models/lightningdit.py:432): both halves equal on ch 0..2 only
    n = n2 // 2
    assert torch.equal(out[:n, :3], out[n:, :3]) and not torch.equal(out[:n, 3:], out[n:, 3:])


@pytest.mark.parametrize("name", GOLDEN_CASES)
@pytest.mark.parametrize("method,steps,shift", [("euler", 6, 0.3), ("heun", 4, 0.1), ("euler", 5, 0.0)])
def test_sampler_matches_reference(golden, name, method, steps, shift):
    fix = golden(name)
    cfg, sd = _cfg(fix), fix["state_dict"]
    grid = O.time_grid(steps, shift)
    assert torch.equal(grid, fix[f"grid_{method}_{steps}_{shift}"])   # bit-exact fp32 grid
    n = fix["x"].shape[0] // 2
    z = torch.cat([fix["x"][:n], fix["x"][:n]], 0)
    out = O.sample_ode(cfg, sd, z, fix["y"], grid, method, 4.0, True, 0.11)
    assert rel_l2(out, fix[f"sample_{method}_{steps}_{shift}"]) < 5e-6


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_sampler_nocfg_matches_reference(golden, name):
    fix = golden(name)
    cfg, sd = _cfg(fix), fix["state_dict"]
    out = O.sample_ode(cfg, sd, fix["x"], fix["y"], O.time_grid(4, 0.3), "euler")
    assert rel_l2(out, fix["sample_nocfg_euler_4_0.3"]) < 5e-6


def test_flop_count_matches_baseline_md():
    # BASELINE.md section 3
    for name, gf in (("LightningDiT-S/1", 12.114), ("LightningDiT-B/1", 46.016), ("LightningDiT-XL/1", 237.252)):
        cfg = O.config_from_name(name, input_size=16)
        assert abs(O.flops_per_forward(cfg) / 1e9 - gf) < 2e-3
    cfg = O.config_from_name("LightningDiT-XL/1", input_size=32)
    assert abs(O.flops_per_forward(cfg) / 1e9 - 1049.114) < 2e-3


def test_init_state_dict_keys_match_reference_fixture(golden):
    for name in GOLDEN_CASES:
        fix = golden(name)
        cfg = _cfg(fix)
        mine = O.init_state_dict(cfg, seed=1)
        ref = fix["state_dict"]
        assert set(mine) == set(ref), (name, set(mine) ^ set(ref))
        for k in ref:
            assert tuple(mine[k].shape) == tuple(ref[k].shape), (name, k)
        # fixed tables must equal the reference's buffers
        assert rel_l2(mine["pos_embed"], ref["pos_embed"]) < 1e-6
        if cfg.use_rope:
            assert rel_l2(mine["feat_rope.freqs_cos"], ref["feat_rope.freqs_cos"]) < 1e-6
            assert rel_l2(mine["feat_rope.freqs_sin"], ref["feat_rope.freqs_sin"]) < 1e-6

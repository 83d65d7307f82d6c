"""Container-only: the CPU oracle against the LIVE, unmodified reference (needs /root/reference, which the GPU
box does not have -> skipped there).  BASELINE.json configs[0]: S/1, n=4, 10-point Euler + CFG."""
import pytest
import torch

from conftest import rel_l2
from oracle import lightningdit_oracle as O
from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(), reason="/root/reference not present")


def test_config1_trajectory_matches_live_reference():
    ld, tr = ref_shim.import_reference()
    torch.manual_seed(0)
    flags = dict(use_qknorm=False, use_swiglu=True, use_rope=True, use_rmsnorm=True, wo_shift=False, in_channels=32)
    model = ld.LightningDiT_models["LightningDiT-S/1"](input_size=16, num_classes=31, **flags).eval()
    ref_shim.rerandomize_zero_params(model, seed=0)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    cfg = O.config_from_name("LightningDiT-S/1", input_size=16, num_classes=31)
    n = 4
    z = torch.randn(n, 32, 16, 16, generator=torch.Generator().manual_seed(1234))
    y = torch.randint(0, 31, (n,), generator=torch.Generator().manual_seed(4321))
    zz, yy = torch.cat([z, z], 0), torch.cat([y, torch.full((n,), 31)])
    sampler = tr.Sampler(tr.create_transport("Linear", "velocity", None, None, None))
    fn = sampler.sample_ode(sampling_method="euler", num_steps=10, atol=1e-6, rtol=1e-3, reverse=False, timestep_shift=0.1)
    with torch.no_grad():
        ref = fn(zz, model.forward_with_cfg, y=yy, cfg_scale=10.0, cfg_interval=True, cfg_interval_start=0.11)[-1]
    mine = O.sample_ode(cfg, sd, zz, yy, O.time_grid(10, 0.1), "euler", 10.0, True, 0.11)
    assert rel_l2(mine, ref) < 5e-6
    samples, _ = ref.chunk(2, dim=0)
    assert torch.equal(ref[:n, :3], ref[n:, :3]) and not torch.equal(ref[:n, 3:], ref[n:, 3:])   # 3-channel CFG quirk

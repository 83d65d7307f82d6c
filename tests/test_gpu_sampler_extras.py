"""Round-2 GPU checks through the C ABI: the long-trajectory drift number, the skip of the discarded unconditional
forward, label range errors, the fused SDE update kernel, the forward-only training loss and a second device in one
process."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR, rel_l2
from oracle import lightningdit_oracle as O

pytestmark = pytest.mark.gpu


def _dev(i=0):
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device (run with -m gpu on the B200 box)")
    return torch.device("cuda", i)


def _model(name, cfg, sd, precision, dev=None, **kw):
    import va_vae_b200 as V
    m = V.LightningDiT_models[name](input_size=cfg.input_size, num_classes=cfg.num_classes, in_channels=32,
                                    use_swiglu=True, use_rope=True, use_rmsnorm=True, precision=precision, **kw)
    m.load_state_dict(sd, strict=True)
    return m.eval().to(dev or _dev())


@pytest.mark.parametrize("precision,vel_tol,final_tol", [("fp32", 1e-5, 1e-5), ("bf16", 1e-2, 1e-2)])
def test_long_trajectory_250_point_heun_cfg_vs_oracle_fixture(precision, vel_tol, final_tol):
    """S/1, n=4, 250-point Heun + CFG 10 (498 evaluations): the engine against the full-length CPU-oracle trajectory
    committed by oracle/gen_golden_long.py.  States where the error ends up after a full benchmark-length run
    (BASELINE.json: 'final-latent rel-L2 stated'); the printed numbers are recorded in DESIGN.md section 2."""
    fix = torch.load(os.path.join(GOLDEN_DIR, "long_s1_heun250.pt"), weights_only=True)
    n, K = fix["n"], fix["num_classes"]
    cfg = O.config_from_name(fix["model"], input_size=16, num_classes=K)
    sd = O.init_state_dict(cfg, seed=fix["seed_weights"])
    z = torch.randn(n, 32, 16, 16, generator=torch.Generator().manual_seed(fix["seed_noise"]))
    y = torch.randint(0, K, (n,), generator=torch.Generator().manual_seed(fix["seed_labels"]))
    zz, yy = torch.cat([z, z], 0), torch.cat([y, torch.full((n,), K)])
    grid = O.time_grid(fix["points"], fix["timestep_shift"])
    assert torch.equal(grid, fix["grid"])
    m = _model(fix["model"], cfg, sd, precision)
    out, trace = m.engine().sample_ode(zz.cuda(), yy.cuda(), grid, fix["method"], True, fix["cfg_scale"], return_trace=True)
    assert trace.shape[0] == 2 * (fix["points"] - 1)
    errs = [rel_l2(trace[e][:n], fix["velocities_kept"][i]) for i, e in enumerate(fix["keep_evals"])]
    final = rel_l2(out[:n], fix["final_kept"])
    print(f"long trajectory {precision}: velocity rel-L2 at evals {fix['keep_evals']} = "
          f"{[f'{e:.2e}' for e in errs]}; max {max(errs):.3e}; final-latent rel-L2 {final:.3e}")
    assert max(errs) < vel_tol, f"velocity rel-L2 along the trajectory {max(errs):.3e}"
    assert final < final_tol, f"final-latent rel-L2 {final:.3e}"


@pytest.mark.parametrize("method", ["euler", "heun"])
@pytest.mark.parametrize("n", [2, 96])
def test_skipping_the_discarded_uncond_forward_keeps_the_kept_rows_bit_identical(method, n):
    """models/lightningdit.py:436-439 computes the unconditional forward of an evaluation at t < cfg_interval_start and
    discards it.  With skip_discarded the engine runs the n conditional rows only: rows [0, n) of the final state (the
    half every caller keeps) must be bit-identical, and fewer forwards must have been executed.  n=2 runs from CUDA
    graphs, n=96 eagerly."""
    cfg = O.config_from_name("LightningDiT-S/1", input_size=16, num_classes=31)
    sd = O.init_state_dict(cfg, seed=0)
    z = torch.randn(n, 32, 16, 16, generator=torch.Generator().manual_seed(1234))
    y = torch.randint(0, 31, (n,), generator=torch.Generator().manual_seed(4321))
    zz, yy = torch.cat([z, z], 0).cuda(), torch.cat([y, torch.full((n,), 31)]).cuda()
    grid = O.time_grid(8, 0.3)
    below = int((grid[:-1] < 0.11).sum()) if method == "euler" else int((grid[:-1] < 0.11).sum() + (grid[1:] < 0.11).sum())
    evals = (len(grid) - 1) * (1 if method == "euler" else 2)
    assert 0 < below < evals
    eng = _model("LightningDiT-S/1", cfg, sd, "bf16").engine()
    outs, rows = {}, {}
    for skip in (False, True):
        for _ in range(3):          # third call of a small batch replays the captured graphs
            r0 = eng.forward_rows
            outs[skip] = eng.sample_ode(zz, yy, grid, method, True, 10.0, True, 0.11, 3, skip_discarded=skip)
            rows[skip] = eng.forward_rows - r0
    assert torch.equal(outs[True][:n], outs[False][:n])
    assert rows[False] == 2 * n * evals and rows[True] == 2 * n * evals - n * below
    # and through the reference-facing API
    import va_vae_b200 as V
    m = _model("LightningDiT-S/1", cfg, sd, "bf16")
    fn = V.Sampler(V.create_transport()).sample_ode(sampling_method=method, num_steps=8, atol=1e-6, rtol=1e-3, reverse=False,
                                                    timestep_shift=0.3, skip_discarded_uncond=True)
    got = fn(zz, m.forward_with_cfg, y=yy, cfg_scale=10.0, cfg_interval=True, cfg_interval_start=0.11)[-1]
    assert torch.equal(got.chunk(2, dim=0)[0], outs[False][:n])


def test_out_of_range_labels_raise_like_nn_embedding():
    """nn.Embedding raises IndexError on a label outside the table (models/lightningdit.py:168).  Host labels are checked
    before the call; device labels by the kernels (clamped read + flag, raised at the next status check)."""
    cfg = O.config_from_name("LightningDiT-S/1", input_size=16, num_classes=31)
    sd = O.init_state_dict(cfg, seed=0)
    m = _model("LightningDiT-S/1", cfg, sd, "bf16")
    eng = m.engine()
    x, t = torch.zeros(2, 32, 16, 16).cuda(), torch.zeros(2).cuda()
    for bad in (torch.tensor([3, 32]), torch.tensor([-1, 0])):
        with pytest.raises(IndexError):
            m(x, t, bad)                                   # host tensor: checked before anything is launched
        m(x, t, bad.cuda())                                # device tensor: no sync inside the call ...
        with pytest.raises(IndexError):
            eng.check_device_status()                      # ... the flag surfaces here
        eng.check_device_status()                          # and is cleared
    with pytest.raises(ValueError):
        m(x, t, torch.zeros(3, dtype=torch.long).cuda())   # wrong y shape
    with pytest.raises(ValueError):
        m.forward_with_cfg(x, torch.zeros(1).cuda(), torch.zeros(2, dtype=torch.long).cuda(), 2.0)   # wrong t shape
    with pytest.raises(ValueError):
        eng.sample_ode(x, torch.zeros(1, dtype=torch.long).cuda(), O.time_grid(3), "euler", True, 2.0)   # y with n rows
    zpin, ypin = torch.zeros(2, 32, 16, 16).pin_memory(), torch.tensor([0, 40]).pin_memory()
    with pytest.raises(RuntimeError, match="index out of range"):
        eng.sample_ode_host(zpin, ypin, O.time_grid(3), torch.empty_like(zpin), "euler", True, 2.0)
    out = m(x, t, torch.tensor([0, 31]).cuda())            # the null label (row num_classes) is valid
    eng.check_device_status()
    assert torch.isfinite(out).all()


@pytest.mark.parametrize("method", ["Euler", "Heun"])
@pytest.mark.parametrize("form", ["SBDM", "sigma", "linear", "decreasing", "inccreasing-decreasing"])
@pytest.mark.parametrize("last_step", ["Mean", "Tweedie", "Euler", None])
def test_fused_sde_update_kernel_is_bit_identical_to_the_tensor_op_sequence(method, form, last_step, monkeypatch):
    """One vavae_sde_update pass per stage must reproduce, bit for bit, the ~10 separately rounded fp32 tensor ops of the
    reference's step (transport/integrators.py:30-49, transport.py:306-334) executed by PyTorch on the same GPU."""
    import va_vae_b200 as V
    dev = _dev()
    model = lambda x, t, **kw: torch.tanh(x * (0.3 + t.view(-1, 1, 1, 1))) - 0.1 * x
    mk = lambda: V.Sampler(V.create_transport()).sample_sde(sampling_method=method, diffusion_form=form, diffusion_norm=0.7,
                                                           last_step=last_step, last_step_size=0.04, num_steps=6)
    x0 = torch.randn(4, 32, 16, 16, generator=torch.Generator().manual_seed(3))
    from va_vae_b200 import engine as E
    real, calls = E.sde_update, []
    monkeypatch.setattr(E, "sde_update", lambda *a, **k: calls.append(1) or real(*a, **k))
    torch.manual_seed(9)
    fused = mk()(x0.to(dev), model)
    assert calls, "the native update was not used for a CUDA fp32 state"
    monkeypatch.setenv("VAVAE_SDE_TENSOR_OPS", "1")          # same loop, PyTorch tensor ops on the same GPU
    n_calls = len(calls)
    torch.manual_seed(9)
    gpu_ops = mk()(x0.to(dev), model)
    assert len(calls) == n_calls
    torch.manual_seed(9)
    ref = mk()(x0, model)                                  # CPU tensor-op path (pinned to the live reference on CPU)
    assert len(fused) == len(gpu_ops) == 6
    for a, b in zip(fused, gpu_ops):
        # SBDM with the velocity transport's t0 = 0 divides by t = 0 (the reference does too, transport.py:84-111):
        # inf / nan must then appear in the same places
        assert torch.allclose(a, b, rtol=0, atol=0, equal_nan=True)
    if torch.isfinite(ref[-1]).all():       # Heun without a last step evaluates the score at t = 1, where var = 0
        assert rel_l2(fused[-1], ref[-1]) < 1e-5


def test_sde_constant_diffusion_fails_like_the_reference():
    """compute_diffusion(form='constant') returns a python float and th.sqrt(2 * diffusion) then raises
    (transport/path.py:56, integrators.py:37): the mirror raises the same TypeError on both of its paths."""
    import va_vae_b200 as V
    model = lambda x, t, **kw: -x
    fn = V.Sampler(V.create_transport()).sample_sde(sampling_method="Euler", diffusion_form="constant", num_steps=4)
    for x in (torch.zeros(2, 4, 8, 8), torch.zeros(2, 4, 8, 8).to(_dev())):
        with pytest.raises(TypeError):
            fn(x, model)


def test_training_losses_forward_only_on_the_engine():
    """Transport.training_losses with the native forward as the model vs the CPU oracle forward: same (t, x0) draws,
    loss values within the bf16 / fp32 forward bars."""
    import va_vae_b200 as V
    cfg = O.config_from_name("LightningDiT-S/1", input_size=16, num_classes=31)
    sd = O.init_state_dict(cfg, seed=0)
    x1 = torch.randn(4, 32, 16, 16, generator=torch.Generator().manual_seed(2))
    y = torch.tensor([0, 5, 30, 7])
    transport = V.create_transport("Linear", "velocity", None, None, None, use_cosine_loss=True, use_lognorm=True)
    dev = _dev()
    for precision, tol in (("fp32", 1e-5), ("bf16", 1e-2)):
        m = _model("LightningDiT-S/1", cfg, sd, precision)
        seen = {}

        def model(x, t, y=None):
            seen.update(x=x.cpu(), t=t.cpu())
            return m.forward(x, t, y=y)
        torch.manual_seed(3)
        np.random.seed(4)
        got = transport.training_losses(model, x1.to(dev), dict(y=y.to(dev)))      # x0 ~ CUDA generator, t ~ numpy
        # the same objective evaluated by the CPU oracle on the (xt, t) the engine saw; ut = x1 - x0 = (x1 - xt) / (1 - t)
        te = seen["t"].view(-1, 1, 1, 1)
        ut = (x1 - seen["x"]) / (1 - te)
        pred = O.forward(cfg, sd, seen["x"], seen["t"], y)
        want_loss = ((pred - ut) ** 2).mean(dim=[1, 2, 3])
        want_cos = (1 - torch.nn.functional.cosine_similarity(pred, ut, dim=1)).mean(dim=[1, 2])
        assert got["pred"].is_cuda and got["loss"].shape == (4,)
        assert rel_l2(got["pred"], pred) < tol
        assert torch.allclose(got["loss"].cpu(), want_loss, rtol=max(4 * tol, 1e-4))
        assert torch.allclose(got["cos_loss"].cpu(), want_cos, rtol=max(4 * tol, 1e-4), atol=tol)


def test_second_engine_on_another_gpu_of_the_same_process():
    """Function attributes (the > 48 KB dynamic shared memory opt-in) and the SM count are per device: an engine on
    cuda:1 created after one on cuda:0 must launch the tcgen05 kernels and give the same result."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs in one process")
    cfg = O.config_from_name("LightningDiT-S/1", input_size=16, num_classes=31)
    sd = O.init_state_dict(cfg, seed=0)
    x = torch.randn(4, 32, 16, 16, generator=torch.Generator().manual_seed(1))
    t, y = torch.full((4,), 0.4), torch.tensor([1, 2, 31, 31])
    outs = []
    for i in (0, 1):
        m = _model("LightningDiT-S/1", cfg, sd, "bf16", dev=_dev(i))
        outs.append(m.forward_with_cfg(x.to(_dev(i)), t.to(_dev(i)), y.to(_dev(i)), 4.0).cpu())
    assert torch.equal(outs[0], outs[1])

"""Sampler.sample_sde (transport.py:336-396, integrators.py:8-75): the mirror against the LIVE reference with a toy
velocity model and the same seed (container only), plus shape / argument behaviour that needs no reference."""
import pytest
import torch

import va_vae_b200 as V
from oracle import ref_shim


def toy_model(x, t, scale=1.0):
    tt = t.view(-1, 1, 1, 1)
    return scale * (torch.tanh(x) * (1 + tt) - 0.3 * x.flip(1) * tt)


def test_sde_returns_num_steps_states_and_rejects_unknown_options():
    tr = V.create_transport("Linear", "velocity", None, None, None)
    fn = V.Sampler(tr).sample_sde(sampling_method="Euler", diffusion_form="sigma", last_step="Mean", num_steps=6)
    torch.manual_seed(0)
    xs = fn(torch.randn(2, 3, 4, 4), toy_model, scale=0.5)
    assert isinstance(xs, list) and len(xs) == 6 and xs[-1].shape == (2, 3, 4, 4) and torch.isfinite(xs[-1]).all()
    with pytest.raises(NotImplementedError):
        V.Sampler(tr).sample_sde(sampling_method="Milstein")
    with pytest.raises(NotImplementedError):
        V.Sampler(tr).sample_sde(diffusion_form="nope", num_steps=4)(torch.zeros(1, 1, 2, 2), toy_model)
    with pytest.raises(NotImplementedError):
        V.Sampler(V.create_transport("VP", "velocity")).sample_sde()


@pytest.mark.skipif(not ref_shim.reference_available(), reason="/root/reference not present")
@pytest.mark.parametrize("method", ["Euler", "Heun"])
@pytest.mark.parametrize("form,last", [("sigma", "Mean"), ("sigma", "Euler"), ("SBDM", "Tweedie"), ("linear", None),
                                       ("decreasing", "Mean"), ("inccreasing-decreasing", "Mean")])
def test_sde_matches_live_reference(method, form, last):
    _, rtr = ref_shim.import_reference()
    eps = None
    kw = dict(sampling_method=method, diffusion_form=form, diffusion_norm=0.7, last_step=last, last_step_size=0.04,
              num_steps=9)
    x0 = torch.randn(3, 4, 6, 6, generator=torch.Generator().manual_seed(5))
    torch.manual_seed(123)
    ref = rtr.Sampler(rtr.create_transport("Linear", "velocity", None, eps, eps)).sample_sde(**kw)(x0, toy_model, scale=0.8)
    torch.manual_seed(123)
    mine = V.Sampler(V.create_transport("Linear", "velocity", None, eps, eps)).sample_sde(**kw)(x0, toy_model, scale=0.8)
    assert len(mine) == len(ref) == 9
    for a, b in zip(mine, ref):
        # (SBDM on the linear path starts at t0 = sample_eps = 0 and divides by t: the reference's states are inf / nan
        # there, and so are the mirror's - compared position by position)
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-6, equal_nan=True)


@pytest.mark.skipif(not ref_shim.reference_available(), reason="/root/reference not present")
def test_constant_diffusion_fails_the_same_way_as_the_reference():
    """diffusion_form='constant' makes the reference take th.sqrt of a python float (integrators.py:36) -> TypeError;
    the mirror keeps that behaviour instead of silently diverging."""
    _, rtr = ref_shim.import_reference()
    x0 = torch.zeros(1, 1, 2, 2)
    kw = dict(sampling_method="Euler", diffusion_form="constant", num_steps=4)
    with pytest.raises(TypeError):
        rtr.Sampler(rtr.create_transport("Linear", "velocity", None, None, None)).sample_sde(**kw)(x0, toy_model)
    with pytest.raises(TypeError):
        V.Sampler(V.create_transport("Linear", "velocity", None, None, None)).sample_sde(**kw)(x0, toy_model)

"""CPU checks of the restated adaptive dopri5 (oracle) and of the product's generic host loop for arbitrary callables."""
import math

import pytest
import torch

from oracle.dopri5 import odeint_dopri5


def _exact(t, y0):      # y' = -y + sin t
    return (y0 + 0.5) * math.exp(-t) + 0.5 * (math.sin(t) - math.cos(t))


def test_restated_dopri5_solves_a_known_ode_to_tolerance():
    y0 = torch.tensor([1.0, 2.0], dtype=torch.float64)
    t = torch.linspace(0, 1, 6, dtype=torch.float64)
    sol, st = odeint_dopri5(lambda tt, y: -y + torch.sin(tt), y0, t, atol=1e-9, rtol=1e-7, return_stats=True)
    for i, tt in enumerate(t):
        for j, v in enumerate((1.0, 2.0)):
            assert abs(float(sol[i, j]) - _exact(float(tt), v)) < 1e-6
    assert st["evals"] == 2 + 6 * (st["accepted"] + st["rejected"])      # FSAL: 6 new evaluations per step


def test_generic_host_loop_matches_oracle_restatement():
    from va_vae_b200.transport import _dopri5_generic
    y0 = torch.randn(3, 4, generator=torch.Generator().manual_seed(0))
    A = torch.randn(4, 4, generator=torch.Generator().manual_seed(1)) * 0.5
    f = lambda tt, y: torch.tanh(y @ A) * (1 + tt)
    t = torch.tensor([0.0, 0.3, 1.0])
    a = _dopri5_generic(f, y0, t, 1e-6, 1e-3)
    b = odeint_dopri5(f, y0, t, atol=1e-6, rtol=1e-3)
    assert torch.allclose(a, b, rtol=0, atol=1e-6)


def test_against_real_torchdiffeq_if_present():
    torchdiffeq = pytest.importorskip("torchdiffeq")
    if not hasattr(torchdiffeq, "__version__") and torchdiffeq.odeint.__module__ == "oracle.ref_shim":
        pytest.skip("only the shim is installed")
    y0 = torch.randn(3, 4, generator=torch.Generator().manual_seed(0))
    A = torch.randn(4, 4, generator=torch.Generator().manual_seed(1)) * 0.5
    f = lambda tt, y: torch.tanh(y @ A) * (1 + tt)
    t = torch.tensor([0.0, 0.3, 1.0])
    ref = torchdiffeq.odeint(f, y0, t, method="dopri5", atol=1e-6, rtol=1e-3)
    mine = odeint_dopri5(f, y0, t, atol=1e-6, rtol=1e-3)
    assert torch.allclose(mine, ref, rtol=0, atol=1e-5)

"""Pins oracle/dopri5.py (the restatement of torchdiffeq's adaptive dopri5, which cannot be installed here) against an
independent third-party Dormand-Prince 5(4): scipy.integrate.RK45.

scipy's solver shares with torchdiffeq the Butcher tableau, the RMS norm of err / (atol + rtol * max(|y0|, |y1|)) and
Hairer's initial-step rule.  It differs in three documented details: (1) torchdiffeq's embedded error weights
(35/384 - 1951/21600, ...) are exactly 2/3 of the textbook b5 - b4 that scipy uses (asserted below: C_ERR == -2/3 * E),
so torchdiffeq accepts slightly larger steps; (2) scipy lets an accepted step SHRINK the next one (torchdiffeq clamps
the factor of an accepted step to >= 1) and (3) clips the last step to t_bound (torchdiffeq steps past and
interpolates); the dense-output polynomial differs too.  So:
  * the numerics-heavy pieces (one RK step, the error norm, the initial step) are compared piece by piece against
    scipy's own primitives, to round-off;
  * a driver built from scipy's primitives plus the restated controller must reproduce the restatement's
    accepted / rejected step sequence and solution exactly (to round-off);
  * the stock scipy.integrate.RK45 run to the same tolerances must agree on the solution within the tolerance, and its
    step counts are compared with the difference attributed to the two controller details above.
Cases: an analytic ODE and a tiny LightningDiT drift (the oracle's forward_with_cfg on a 2-layer model).
"""
import math

import numpy as np
import pytest
import torch

from oracle import dopri5 as R
from oracle import lightningdit_oracle as O

rk = pytest.importorskip("scipy.integrate._ivp.rk")
common = pytest.importorskip("scipy.integrate._ivp.common")
from scipy.integrate import RK45  # noqa: E402


def analytic_problem():
    A = np.array([[-0.5, 2.0, 0.0], [-2.0, -0.5, 0.3], [0.1, 0.0, -1.5]])
    f = lambda t, y: A @ y + np.array([math.sin(3 * t), 0.0, math.cos(t)])
    return f, np.array([1.0, -0.5, 2.0]), 0.0, 2.0, 1e-6, 1e-3


def dit_problem():
    cfg = O.DiTConfig(input_size=4, patch_size=1, in_channels=4, hidden_size=32, depth=2, num_heads=2, num_classes=5)
    sd = O.init_state_dict(cfg, seed=3)
    n = 2
    yl = torch.tensor([1, 3, 5, 5])
    shape = (2 * n, 4, 4, 4)

    def f(t, y):
        x = torch.from_numpy(np.asarray(y, dtype=np.float64)).float().reshape(shape)
        tv = torch.full((2 * n,), float(t))
        return O.forward_with_cfg(cfg, sd, x, tv, yl, 4.0).double().reshape(-1).numpy()

    z = torch.randn(n, 4, 4, 4, generator=torch.Generator().manual_seed(7))
    return f, torch.cat([z, z]).double().reshape(-1).numpy(), 0.0, 1.0, 1e-6, 1e-3


PROBLEMS = {"analytic": analytic_problem, "tiny_dit": dit_problem}


def torch_f(f):
    return lambda t, y: torch.from_numpy(f(float(t), y.numpy()))


def restated(f, y0, t0, t1, atol, rtol):
    sol, st = R.odeint_dopri5(torch_f(f), torch.from_numpy(y0.copy()), torch.tensor([t0, t1], dtype=torch.float64),
                              atol=atol, rtol=rtol, return_stats=True)
    return sol[-1].numpy(), st


@pytest.mark.parametrize("name", list(PROBLEMS))
def test_tableau_error_estimate_and_initial_step_match_scipy(name):
    f, y0, t0, t1, atol, rtol = PROBLEMS[name]()
    # tableau: torchdiffeq / the restatement list the same coefficients scipy uses
    for i in range(6):
        assert np.allclose(R.BETA[i], RK45.A[i + 1][: i + 1] if i < 5 else RK45.B, rtol=0, atol=1e-15)
    assert np.allclose(R.ALPHA[:5], RK45.C[1:], rtol=0, atol=1e-15)
    # error weights incl. the FSAL stage: torchdiffeq's are the textbook b5 - b4 (scipy's E, up to sign) scaled by 2/3
    assert np.allclose(R.C_ERR, -2.0 / 3.0 * RK45.E, rtol=0, atol=1e-15)
    # initial step: scipy.select_initial_step(order = error-estimator order 4) vs the restated Hairer rule
    f0 = f(t0, y0)
    h_scipy = common.select_initial_step(f, t0, y0, np.inf, np.inf, f0, 1.0, RK45.error_estimator_order, rtol, atol)
    scale = atol + np.abs(y0) * rtol
    rms = lambda v: float(np.sqrt(np.mean(v * v)))
    d0, d1 = rms(y0 / scale), rms(f0 / scale)
    h0 = 1e-6 if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * d0 / d1
    d2 = rms((f(t0 + h0, y0 + h0 * f0) - f0) / scale) / h0
    h1 = max(1e-6, h0 * 1e-3) if (d1 <= 1e-15 and d2 <= 1e-15) else (0.01 / max(d1, d2)) ** 0.2
    assert min(100 * h0, h1) == pytest.approx(h_scipy, rel=1e-12)
    # one step + error norm from the same (t, y, f, h)
    h = h_scipy
    K = np.empty((7, y0.size))
    y_new, f_new = rk.rk_step(f, t0, y0, f0, h, RK45.A, RK45.B, RK45.C, K)
    err_scipy = rms((K.T @ RK45.E) * h / (atol + np.maximum(np.abs(y0), np.abs(y_new)) * rtol))
    ks = [torch.from_numpy(f0)]
    yt = torch.from_numpy(y0)
    yi = yt
    for i in range(6):
        yi = yt
        for j in range(i + 1):
            yi = yi + (R.BETA[i][j] * h) * ks[j]
        ks.append(torch.from_numpy(f(t0 + R.ALPHA[i] * h, yi.numpy())))
    err = sum((R.C_ERR[j] * h) * ks[j] for j in range(7))
    ratio = R._rms(err / (atol + rtol * torch.max(yt.abs(), yi.abs())))
    assert np.allclose(yi.numpy(), y_new, rtol=1e-12, atol=1e-14)
    assert ratio == pytest.approx(2.0 / 3.0 * err_scipy, rel=1e-8)


def scipy_primitives_with_restated_controller(f, y0, t0, t1, atol, rtol):
    """torchdiffeq's accept / reject / step-factor rule driven by scipy's rk_step and error estimate."""
    y, t, fy = y0.copy(), t0, f(t0, y0)
    h = common.select_initial_step(f, t0, y0, np.inf, np.inf, fy, 1.0, RK45.error_estimator_order, rtol, atol)
    K = np.empty((7, y0.size))
    seq = []
    while True:
        y_new, f_new = rk.rk_step(f, t, y, fy, h, RK45.A, RK45.B, RK45.C, K)
        scale = atol + np.maximum(np.abs(y), np.abs(y_new)) * rtol
        ratio = 2.0 / 3.0 * float(np.sqrt(np.mean(((K.T @ RK45.E) * h / scale) ** 2)))   # torchdiffeq's error weights
        accept = ratio <= 1
        seq.append(accept)
        if accept:
            if t + h >= t1:
                return seq, (t, t + h, y, y_new, fy, f_new, K.copy(), h)
            t, y, fy = t + h, y_new, f_new
        h *= 10.0 if ratio == 0 else min(10.0, max(0.9 / ratio ** 0.2, 1.0 if ratio < 1 else 0.2))


@pytest.mark.parametrize("name", list(PROBLEMS))
def test_step_sequence_matches_scipy_primitives_under_the_restated_controller(name):
    f, y0, t0, t1, atol, rtol = PROBLEMS[name]()
    seq, last = scipy_primitives_with_restated_controller(f, y0, t0, t1, atol, rtol)
    mine, st = restated(f, y0, t0, t1, atol, rtol)
    assert st["accepted"] == sum(seq) and st["rejected"] == len(seq) - sum(seq)
    assert st["evals"] == 2 + 6 * len(seq)
    # the restatement's value at t1 is the 4th-order dense output inside the last accepted step: rebuild it from
    # scipy's stage values with the mid-point weights torchdiffeq uses
    ts, te, ya, yb, fa, fb, K, h = last
    y_mid = ya + h * (K.T @ np.array(R.C_MID))
    x = (t1 - ts) / (te - ts)
    want = R._interp(*(torch.from_numpy(np.asarray(v)) for v in (ya, yb, y_mid, fa, fb)), h, x).numpy()
    assert np.allclose(mine, want, rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize("name", list(PROBLEMS))
def test_solution_agrees_with_stock_scipy_rk45(name):
    f, y0, t0, t1, atol, rtol = PROBLEMS[name]()
    mine, st = restated(f, y0, t0, t1, atol, rtol)
    solver = RK45(f, t0, y0, t1, rtol=rtol, atol=atol)
    accepted = 0
    while solver.status == "running":
        solver.step()
        accepted += 1
    rejected = (solver.nfev - 2) // 6 - accepted
    # both integrate to the same tolerances: the two answers differ by a few local errors at most
    tol = atol + rtol * np.maximum(np.abs(mine), np.abs(solver.y))
    assert np.sqrt(np.mean(((mine - solver.y) / tol) ** 2)) < 3.0
    # step counts: scipy may shrink after an accepted step and clips the last step to t1, torchdiffeq does neither;
    # the counts therefore differ by a handful of steps at most, never in order of magnitude
    assert abs((st["accepted"] + st["rejected"]) - (accepted + rejected)) <= max(3, (accepted + rejected) // 3)

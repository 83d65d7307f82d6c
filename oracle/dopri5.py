"""TEST INFRASTRUCTURE ONLY -- restatement of torchdiffeq's adaptive `dopri5` (the un-vendored, un-pinned dependency the
reference reaches at LightningDiT/transport/integrators.py:118-125 with method="dopri5", atol=1e-6, rtol=1e-3).

PARITY STATUS: torchdiffeq itself is not installed here and not under /root/reference, so this file restates its
published algorithm (Dormand-Prince 5(4) tableau with FSAL; Hairer initial step with order 4; RMS error norm over the
whole state; step factor min(10, max(0.9 / ratio^(1/5), dfactor)) with dfactor 1 after an accepted step and 0.2 after a
rejected one; float64 time / step size, state-dtype stage arithmetic; 4th-order dense output, output points obtained by
interpolation inside the accepted step that covers them).  It is PINNED AGAINST AN INDEPENDENT THIRD-PARTY
IMPLEMENTATION of the same method, scipy.integrate.RK45 (tests/test_dopri5_scipy.py): the tableau, one full RK step, the
error norm and the initial step agree with scipy's own primitives to round-off; a driver made of scipy's primitives plus
the controller restated here reproduces this file's accepted / rejected step sequence and solution exactly, on an
analytic ODE and on a tiny LightningDiT drift; stock RK45 agrees on the solution within the tolerance.  What that
cross-check cannot cover, because scipy differs there by design, is restated from torchdiffeq's documentation only:
its embedded error weights are 2/3 of the textbook b5 - b4 (asserted: C_ERR == -2/3 * scipy's E), an accepted step never
shrinks the next one, and the last step is not clipped to t1.  If torchdiffeq is ever importable,
tests/test_dopri5.py cross-checks against it directly.  This file is the checker for `vavae_sample_dopri5`.
"""
import torch

ALPHA = [1 / 5, 3 / 10, 4 / 5, 8 / 9, 1.0, 1.0]
BETA = [
    [1 / 5],
    [3 / 40, 9 / 40],
    [44 / 45, -56 / 15, 32 / 9],
    [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729],
    [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656],
    [35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84],
]
C_ERR = [35 / 384 - 1951 / 21600, 0, 500 / 1113 - 22642 / 50085, 125 / 192 - 451 / 720,
         -2187 / 6784 - -12231 / 42400, 11 / 84 - 649 / 6300, -1. / 60.]
C_MID = [6025192743 / 30085553152 / 2, 0, 51252292925 / 65400821598 / 2, -2691868925 / 45128329728 / 2,
         187940372067 / 1594534317056 / 2, -1776094331 / 19743644256 / 2, 11237099 / 235043384 / 2]


def _rms(x):
    return float(x.double().pow(2).mean().sqrt())


def _interp(y0, y1, y_mid, f0, f1, dt, x):
    a = 2 * dt * (f1 - f0) - 8 * (y1 + y0) + 16 * y_mid
    b = dt * (5 * f0 - 3 * f1) + 18 * y0 + 14 * y1 - 32 * y_mid
    c = dt * (f1 - 4 * f0) - 11 * y0 - 5 * y1 + 16 * y_mid
    d = dt * f0
    total = y0 + x * d
    xp = x * x
    total = total + xp * c
    xp = xp * x
    total = total + xp * b
    xp = xp * x
    return total + xp * a


@torch.no_grad()
def odeint_dopri5(func, y0, t, atol=1e-6, rtol=1e-3, return_stats=False, max_num_steps=100000):
    """func(t: 0-dim tensor of y0.dtype, y) -> dy/dt.  t: 1-D increasing tensor of output times.
    Returns stack of the solution at every t (first entry = y0), like torchdiffeq.odeint."""
    dtype = y0.dtype
    times = [float(x) for x in t]
    atol, rtol = float(atol), float(rtol)
    n_evals = 0

    def f(tt, yy):
        nonlocal n_evals
        n_evals += 1
        return func(torch.tensor(tt, dtype=dtype), yy)

    t0 = times[0]
    f0 = f(t0, y0)
    scale = atol + y0.abs() * rtol
    d0, d1 = _rms(y0 / scale), _rms(f0 / scale)
    h0 = 1e-6 if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * d0 / d1
    y1 = y0 + torch.tensor(h0, dtype=dtype) * f0
    f1 = f(t0 + h0, y1)
    d2 = _rms((f1 - f0) / scale) / h0
    h1 = max(1e-6, h0 * 1e-3) if (d1 <= 1e-15 and d2 <= 1e-15) else (0.01 / max(d1, d2)) ** (1.0 / 5.0)
    dt = min(100 * h0, h1)

    out = [y0]
    tcur, ycur, fcur = t0, y0, f0
    seg = None                                   # (t_start, t_end, y0, y1, y_mid, f0, f1, dt) of the last accepted step
    accepted = rejected = 0
    for next_t in times[1:]:
        steps = 0
        while seg is None or next_t > seg[1]:
            steps += 1
            assert steps < max_num_steps, "max_num_steps exceeded"
            dtf = torch.tensor(dt, dtype=dtype)
            tf = torch.tensor(tcur, dtype=dtype)
            ks = [fcur]
            yi = None
            for i in range(6):
                yi = ycur
                for j in range(i + 1):
                    if BETA[i][j] != 0:
                        yi = yi + (torch.tensor(BETA[i][j], dtype=dtype) * dtf) * ks[j]
                ti = float(torch.tensor(tcur + dt, dtype=dtype)) if ALPHA[i] == 1.0 else float(tf + torch.tensor(ALPHA[i], dtype=dtype) * dtf)
                ks.append(f(ti, yi))
            ynew, fnew = yi, ks[-1]
            err = sum((torch.tensor(C_ERR[j], dtype=dtype) * dtf) * ks[j] for j in range(7) if C_ERR[j] != 0)
            tol = atol + rtol * torch.max(ycur.abs(), ynew.abs())
            ratio = _rms(err / tol)
            if ratio <= 1:
                accepted += 1
                y_mid = ycur
                for j in range(7):
                    if C_MID[j] != 0:
                        y_mid = y_mid + (torch.tensor(C_MID[j], dtype=dtype) * dtf) * ks[j]
                seg = (tcur, tcur + dt, ycur, ynew, y_mid, fcur, fnew, dtf)
                tcur, ycur, fcur = tcur + dt, ynew, fnew
            else:
                rejected += 1
            if ratio == 0:
                factor = 10.0
            else:
                factor = min(10.0, max(0.9 / ratio ** 0.2, 1.0 if ratio < 1 else 0.2))
            dt = dt * factor
        ts, te, a0, a1, am, g0, g1, dtf = seg
        x = torch.tensor((next_t - ts) / (te - ts), dtype=dtype)
        out.append(_interp(a0, a1, am, g0, g1, dtf, x))
    res = torch.stack(out)
    return (res, dict(accepted=accepted, rejected=rejected, evals=n_evals)) if return_stats else res

"""Mirror of the reference's `transport` sampling API (create_transport / Sampler.sample_ode / Sampler.sample_sde).

Reference: /root/reference/LightningDiT/transport/__init__.py:3-72 (create_transport),
transport/transport.py:40-111,218-250,270-283,398-443 (Transport.check_interval / get_drift, Sampler.sample_ode),
transport/integrators.py:77-126 (class ode), plus Sampler.sample_sde (transport.py:336-396, integrators.py:8-75) for the
linear path, and the forward-only value of Transport.training_losses (transport.py:112-215, path.py:18-136) for the
linear path with a velocity model.  Likelihoods are out of scope for this engine.

`sample_fn = Sampler(transport).sample_ode(...)` returns a bound `ode.sample` exactly like the reference; when
`sample_fn(x, model.forward_with_cfg, **kw)` (or `model.forward`) is called with a bound method of the B200
`LightningDiT`, a velocity model and a fixed-grid method, the whole loop runs natively in
libvavae_b200.so (no per-step host sync, no trajectory buffer).  Any other callable is integrated by a
generic fixed-grid loop that only calls `model(x, t, **kw)`.
"""
from __future__ import annotations

import enum
import os

import numpy as np
import torch as th


class ModelType(enum.Enum):
    NOISE = enum.auto()
    SCORE = enum.auto()
    VELOCITY = enum.auto()


class PathType(enum.Enum):
    LINEAR = enum.auto()
    GVP = enum.auto()
    VP = enum.auto()


class WeightType(enum.Enum):
    NONE = enum.auto()
    VELOCITY = enum.auto()
    LIKELIHOOD = enum.auto()


class Transport:
    """Sampling-side state of transport/transport.py:40-111."""

    def __init__(self, *, model_type, path_type, loss_type, train_eps, sample_eps, use_cosine_loss=False,
                 use_lognorm=False, partitial_train=None, partial_ratio=1.0, shift_lg=False):
        self.loss_type = loss_type
        self.model_type = model_type
        self.path_type = path_type
        self.train_eps = train_eps
        self.sample_eps = sample_eps
        self.use_cosine_loss = use_cosine_loss
        self.use_lognorm = use_lognorm
        self.partitial_train = partitial_train
        self.partial_ratio = partial_ratio
        self.shift_lg = shift_lg

    def check_interval(self, train_eps, sample_eps, *, diffusion_form="SBDM", sde=False, reverse=False, eval=False,
                       last_step_size=0.0):
        """transport/transport.py:84-111."""
        t0, t1 = 0, 1
        eps = train_eps if not eval else sample_eps
        if self.path_type == PathType.VP:
            t1 = 1 - eps if (not sde or last_step_size == 0) else 1 - last_step_size
        elif self.path_type in (PathType.LINEAR, PathType.GVP) and (self.model_type != ModelType.VELOCITY or sde):
            t0 = eps if (diffusion_form == "SBDM" and sde) or self.model_type != ModelType.VELOCITY else 0
            t1 = 1 - eps if (not sde or last_step_size == 0) else 1 - last_step_size
        if reverse:
            t0, t1 = 1 - t0, 1 - t1
        return t0, t1

    def get_drift(self):
        """transport/transport.py:218-250; only the velocity parameterisation is on the engine's path."""
        if self.model_type != ModelType.VELOCITY:
            raise NotImplementedError("the B200 sampling engine covers prediction='velocity' only "
                                      "(every reference config uses it; noise/score drifts are out of scope)")

        def body_fn(x, t, model, **model_kwargs):
            model_output = model(x, t, **model_kwargs)
            assert model_output.shape == x.shape, "Output shape from ODE solver must match input shape"
            return model_output
        body_fn.is_plain_velocity = True
        return body_fn

    # ---- training-side reuse of the forward: loss VALUES only, no autograd through the engine
    def sample_logit_normal(self, mu, sigma, size=1):
        """transport/transport.py:112-122 (scipy's norm.rvs on numpy's global RNG, logistic, fp32 tensor)."""
        from scipy.stats import norm
        samples = norm.rvs(loc=mu, scale=sigma, size=size)
        samples = 1 / (1 + np.exp(-samples))
        return th.tensor(samples, dtype=th.float32)

    def sample_in_range(self, mu, sigma, target_size, range_min=0, range_max=0.5):
        """transport/transport.py:124-133."""
        samples = []
        while len(samples) < target_size:
            generated = self.sample_logit_normal(mu, sigma, size=target_size)
            samples.extend(generated[(generated >= range_min) & (generated <= range_max)])
        return th.tensor(samples[:target_size])

    def sample(self, x1, sp_timesteps=None, shifted_mu=0):
        """transport/transport.py:135-166: x0 ~ N(0, I) like x1; t uniform / logit-normal / partial range.  Same RNG
        calls in the same order as the reference, so a seeded run draws the same (t, x0)."""
        x0 = th.randn_like(x1)
        t0, t1 = self.check_interval(self.train_eps, self.sample_eps)
        if not self.use_lognorm:
            if self.partitial_train is not None and th.rand(1) < self.partial_ratio:
                t = th.rand((x1.shape[0],)) * (self.partitial_train[1] - self.partitial_train[0]) + self.partitial_train[0]
            else:
                t = th.rand((x1.shape[0],)) * (t1 - t0) + t0
        elif not self.shift_lg:
            if self.partitial_train is not None and th.rand(1) < self.partial_ratio:
                t = self.sample_in_range(0, 1, x1.shape[0], range_min=self.partitial_train[0],
                                         range_max=self.partitial_train[1])
            else:
                t = self.sample_logit_normal(0, 1, size=x1.shape[0]) * (t1 - t0) + t0
        else:
            assert self.partitial_train is None, "Shifted lognormal distribution is not compatible with partial training"
            t = self.sample_logit_normal(shifted_mu, 1, size=x1.shape[0]) * (t1 - t0) + t0
        if sp_timesteps is not None:
            t = th.rand((x1.shape[0],)) * (sp_timesteps[1] - sp_timesteps[0]) + sp_timesteps[0]
        return t.to(x1), x0, x1

    @staticmethod
    def plan(t, x0, x1):
        """ICPlan.plan (transport/path.py:114-136): xt = t*x1 + (1-t)*x0 ; ut = 1*x1 + (-1)*x0."""
        te = t.view(t.size(0), *([1] * (x1.dim() - 1)))
        xt = te * x1 + (1 - te) * x0
        ut = 1 * x1 + -1 * x0
        return t, xt, ut

    @th.no_grad()
    def training_losses(self, model, x1, model_kwargs=None, sp_timesteps=None, shifted_mu=0):
        """transport/transport.py:169-215 for path_type='Linear', prediction='velocity' (every reference config):
        returns {'pred', 'loss' [, 'cos_loss']} with the reference's values.  FORWARD ONLY: `model` is evaluated by the
        inference engine, nothing is differentiable - this serves validation-loss curves and parity checks of the
        training objective, not an optimiser step (the backward pass is out of scope)."""
        if self.path_type != PathType.LINEAR or self.model_type != ModelType.VELOCITY:
            raise NotImplementedError("training_losses covers path_type='Linear' with prediction='velocity'")
        if model_kwargs is None:
            model_kwargs = {}
        t, x0, x1 = self.sample(x1, sp_timesteps, shifted_mu)
        t, xt, ut = self.plan(t, x0, x1)
        model_output = model(xt, t, **model_kwargs)
        B, *_, C = xt.shape
        assert model_output.size() == (B, *xt.size()[1:-1], C)
        mean_flat = lambda v: th.mean(v, dim=list(range(1, v.dim())))       # transport/utils.py:13-17
        terms = {"pred": model_output, "loss": mean_flat((model_output - ut) ** 2)}
        if self.use_cosine_loss:
            terms["cos_loss"] = mean_flat(1 - th.nn.functional.cosine_similarity(model_output, ut, dim=1))
        return terms


def create_transport(path_type="Linear", prediction="velocity", loss_weight=None, train_eps=None, sample_eps=None,
                     use_cosine_loss=None, use_lognorm=None, partitial_train=None, partial_ratio=1.0, shift_lg=False):
    """transport/__init__.py:3-72 (same arguments, same epsilon rules)."""
    model_type = {"noise": ModelType.NOISE, "score": ModelType.SCORE}.get(prediction, ModelType.VELOCITY)
    loss_type = {"velocity": WeightType.VELOCITY, "likelihood": WeightType.LIKELIHOOD}.get(loss_weight, WeightType.NONE)
    path = {"Linear": PathType.LINEAR, "GVP": PathType.GVP, "VP": PathType.VP}[path_type]
    if path == PathType.VP:
        train_eps, sample_eps = (1e-5 if train_eps is None else train_eps), (1e-3 if train_eps is None else sample_eps)
    elif path in (PathType.GVP, PathType.LINEAR) and model_type != ModelType.VELOCITY:
        train_eps, sample_eps = (1e-3 if train_eps is None else train_eps), (1e-3 if train_eps is None else sample_eps)
    else:
        train_eps, sample_eps = 0, 0
    return Transport(model_type=model_type, path_type=path, loss_type=loss_type, train_eps=train_eps,
                     sample_eps=sample_eps, use_cosine_loss=use_cosine_loss, use_lognorm=use_lognorm,
                     partitial_train=partitial_train, partial_ratio=partial_ratio, shift_lg=shift_lg)


class LazyTrajectory:
    """What the native sampler returns in place of the reference's (num_steps, *x.shape) tensor
    (inference.py:216 only ever takes `[-1]`): the final state, indexable by -1 / num_steps-1."""

    def __init__(self, final, num_steps, trace=None):
        self._final = final
        self.num_steps = num_steps
        self.velocity_trace = trace
        self.shape = (num_steps,) + tuple(final.shape)

    def __len__(self):
        return self.num_steps

    def __getitem__(self, idx):
        if isinstance(idx, int) and idx in (-1, self.num_steps - 1):
            return self._final
        raise IndexError("the native sampler keeps only the final state of the trajectory (what every reference "
                         "caller uses); pass keep_trajectory=True to Sampler.sample_ode for intermediate points")


def _time_grid(t0, t1, num_steps, timestep_shift):
    """transport/integrators.py:94-101, same ops so the fp32 grid is bit-identical."""
    t = th.linspace(t0, t1, num_steps)
    if timestep_shift > 0:
        def compute_tm(t_n, s):
            return (s * t_n) / (1 + (s - 1) * t_n)
        t = th.tensor([compute_tm(t_n, timestep_shift) for t_n in t])
    return t


class ode:
    """transport/integrators.py:77-126."""

    def __init__(self, drift, *, t0, t1, sampler_type, num_steps, atol, rtol, timestep_shift, keep_trajectory=False,
                 native_ok=True, skip_discarded_uncond=False):
        assert t0 < t1, "ODE sampler has to be in forward time"
        self.drift = drift
        self.t = _time_grid(t0, t1, num_steps, timestep_shift)
        self.atol, self.rtol = atol, rtol
        self.sampler_type = sampler_type
        self.keep_trajectory = keep_trajectory
        self.native_ok = native_ok
        self.skip_discarded_uncond = skip_discarded_uncond

    def _native_target(self, model):
        from .lightningdit import LightningDiT
        owner = getattr(model, "__self__", None)
        if not isinstance(owner, LightningDiT) or not self.native_ok or self.keep_trajectory:
            return None, None
        name = getattr(model, "__name__", "")
        if name in ("forward_with_cfg", "forward"):
            return owner, name
        return None, None

    def sample(self, x, model, **model_kwargs):
        method = self.sampler_type
        if method not in ("euler", "heun", "heun2", "dopri5"):
            raise NotImplementedError(f"sampling_method={method!r}: the B200 engine implements the fixed-grid "
                                      "solvers 'euler' / 'heun' and the adaptive 'dopri5'")
        owner, name = self._native_target(model)
        if owner is not None and method == "dopri5":
            eng = owner.engine()
            owner._require_eval()
            cfg = name == "forward_with_cfg"
            final = eng.sample_dopri5(x, model_kwargs["y"], float(self.t[0]), float(self.t[-1]), self.atol, self.rtol,
                                      cfg, model_kwargs.get("cfg_scale", 1.0), model_kwargs.get("cfg_interval"),
                                      model_kwargs.get("cfg_interval_start"), 3)
            return LazyTrajectory(final, len(self.t))
        if method == "dopri5":
            return _dopri5_generic(lambda tt, yy: self.drift(yy, th.ones(yy.size(0)).to(yy.device) * tt, model, **model_kwargs),
                                   x, self.t.to(x.device), self.atol, self.rtol)
        if owner is not None:
            eng = owner.engine()
            owner._require_eval()
            if name == "forward_with_cfg":
                final = eng.sample_ode(x, model_kwargs["y"], self.t, method, True, model_kwargs["cfg_scale"],
                                       model_kwargs.get("cfg_interval"), model_kwargs.get("cfg_interval_start"), 3,
                                       skip_discarded=self.skip_discarded_uncond)
            else:
                final = eng.sample_ode(x, model_kwargs["y"], self.t, method, False)
            return LazyTrajectory(final, len(self.t))
        # generic fixed-grid loop (torchdiffeq fixed-grid semantics: one step per interval, solution[j] = y_j)
        device = x.device
        t = self.t.to(device)
        ys, y = [x], x
        for j in range(len(t) - 1):
            t0, t1 = t[j], t[j + 1]
            dt = t1 - t0
            k1 = self.drift(y, th.ones(y.size(0)).to(device) * t0, model, **model_kwargs)
            if method == "euler":
                y = y + dt * k1
            else:
                k2 = self.drift(y + dt * k1, th.ones(y.size(0)).to(device) * t1, model, **model_kwargs)
                y = y + (dt * 0.5) * (k1 + k2)
            ys.append(y)
        return th.stack(ys)


_DP_ALPHA = [1 / 5, 3 / 10, 4 / 5, 8 / 9, 1.0, 1.0]
_DP_BETA = [[1 / 5], [3 / 40, 9 / 40], [44 / 45, -56 / 15, 32 / 9],
            [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729],
            [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656],
            [35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84]]
_DP_ERR = [35 / 384 - 1951 / 21600, 0, 500 / 1113 - 22642 / 50085, 125 / 192 - 451 / 720,
           -2187 / 6784 - -12231 / 42400, 11 / 84 - 649 / 6300, -1. / 60.]
_DP_MID = [6025192743 / 30085553152 / 2, 0, 51252292925 / 65400821598 / 2, -2691868925 / 45128329728 / 2,
           187940372067 / 1594534317056 / 2, -1776094331 / 19743644256 / 2, 11237099 / 235043384 / 2]


@th.no_grad()
def _dopri5_generic(func, y0, t, atol, rtol, max_num_steps=100000):
    """Adaptive Dormand-Prince 5(4) for an arbitrary drift callable (host loop over torch tensors): the algorithm
    torchdiffeq.odeint(method="dopri5") applies -- Hairer initial step, RMS error norm, safety 0.9 / ifactor 10 /
    dfactor 0.2, FSAL, outputs by 4th-order dense interpolation.  Returns stack of the solution at every t."""
    dtype = y0.dtype
    rms = lambda v: float(v.double().pow(2).mean().sqrt())
    sc = lambda v: th.tensor(v, dtype=dtype, device=y0.device)
    times = [float(v) for v in t]
    f = lambda tt, yy: func(sc(tt), yy)
    t0 = times[0]
    f0 = f(t0, y0)
    scale = atol + y0.abs() * rtol
    d0, d1 = rms(y0 / scale), rms(f0 / scale)
    h0 = 1e-6 if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * d0 / d1
    f1 = f(t0 + h0, y0 + sc(h0) * f0)
    d2 = rms((f1 - f0) / scale) / h0
    h1 = max(1e-6, h0 * 1e-3) if (d1 <= 1e-15 and d2 <= 1e-15) else (0.01 / max(d1, d2)) ** 0.2
    dt = min(100 * h0, h1)
    out, tcur, ycur, fcur, seg = [y0], t0, y0, f0, None
    for next_t in times[1:]:
        steps = 0
        while seg is None or next_t > seg[1]:
            steps += 1
            if steps > max_num_steps:
                raise RuntimeError("max_num_steps exceeded")
            dtf, ks, yi = sc(dt), [fcur], None
            for i in range(6):
                yi = ycur
                for j in range(i + 1):
                    if _DP_BETA[i][j] != 0:
                        yi = yi + (sc(_DP_BETA[i][j]) * dtf) * ks[j]
                ti = float(sc(tcur + dt)) if _DP_ALPHA[i] == 1.0 else float(sc(tcur) + sc(_DP_ALPHA[i]) * dtf)
                ks.append(f(ti, yi))
            err = sum((sc(_DP_ERR[j]) * dtf) * ks[j] for j in range(7) if _DP_ERR[j] != 0)
            ratio = rms(err / (atol + rtol * th.max(ycur.abs(), yi.abs())))
            if ratio <= 1:
                ymid = ycur
                for j in range(7):
                    if _DP_MID[j] != 0:
                        ymid = ymid + (sc(_DP_MID[j]) * dtf) * ks[j]
                seg = (tcur, tcur + dt, ycur, yi, ymid, fcur, ks[-1], dtf)
                tcur, ycur, fcur = tcur + dt, yi, ks[-1]
            dt = dt * (10.0 if ratio == 0 else min(10.0, max(0.9 / ratio ** 0.2, 1.0 if ratio < 1 else 0.2)))
        ts, te, a0, a1, am, g0, g1, dtf = seg
        x = sc((next_t - ts) / (te - ts))
        ca = 2 * dtf * (g1 - g0) - 8 * (a1 + a0) + 16 * am
        cb = dtf * (5 * g0 - 3 * g1) + 18 * a0 + 14 * a1 - 32 * am
        cc = dtf * (g1 - 4 * g0) - 11 * a0 - 5 * a1 + 16 * am
        out.append(a0 + x * (dtf * g0) + x * x * cc + x * x * x * cb + x * x * x * x * ca)
    return th.stack(out)


class Sampler:
    """transport/transport.py:270-283,398-443 (ODE side)."""

    def __init__(self, transport):
        self.transport = transport
        self.drift = self.transport.get_drift()

    def sample_ode(self, *, sampling_method="dopri5", num_steps=50, atol=1e-6, rtol=1e-3, reverse=False,
                   timestep_shift=0.0, keep_trajectory=False, skip_discarded_uncond=False):
        """Reference arguments (transport.py:398-408) plus two keyword-only extensions that default to the reference's
        behaviour: keep_trajectory (materialise every grid point) and skip_discarded_uncond - with
        `cfg_interval=True`, do not run the unconditional forward of evaluations at t < cfg_interval_start, whose
        output forward_with_cfg throws away (lightningdit.py:436-439).  Callers keep `[-1].chunk(2)[0]`
        (inference.py:216-218); those rows are bit-identical with and without the option, the other half is not."""
        native_ok = not reverse
        if reverse:
            def drift(x, t, model, **kwargs):
                return self.drift(x, th.ones_like(t) * (1 - t), model, **kwargs)
        else:
            drift = self.drift
        t0, t1 = self.transport.check_interval(self.transport.train_eps, self.transport.sample_eps, sde=False, eval=True,
                                               reverse=reverse, last_step_size=0.0)
        _ode = ode(drift=drift, t0=t0, t1=t1, sampler_type=sampling_method, num_steps=num_steps, atol=atol, rtol=rtol,
                   timestep_shift=timestep_shift, keep_trajectory=keep_trajectory, native_ok=native_ok,
                   skip_discarded_uncond=skip_discarded_uncond)
        return _ode.sample

    def sample_sde(self, *, sampling_method="Euler", diffusion_form="SBDM", diffusion_norm=1.0, last_step="Mean",
                   last_step_size=0.04, num_steps=250):
        """transport/transport.py:336-396 with the integrator of transport/integrators.py:8-75, for the linear path with
        a velocity model (the only combination the reference's configs create).  Same arguments, same time grid, same
        noise stream (``th.randn(x.size()).to(x)`` drawn from the global CPU generator once per step, as the reference
        does), same list of ``num_steps`` states as return value.  The reference evaluates the model twice per drift
        (once for the velocity, once for the score, transport.py:306-308); both are derived here from ONE evaluation,
        which is bit-identical for a deterministic model and halves the cost.  Every model evaluation is a native
        forward when ``model`` is a bound method of the B200 module, and for CUDA states the whole elementwise update of
        a step (drift, score, Euler-Maruyama / Heun stage) is ONE pass of the library's `vavae_sde_update` kernel with
        the reference's fp32 operation order (coefficients are host scalars because the batch shares one time)."""
        if self.transport.path_type != PathType.LINEAR or self.transport.model_type != ModelType.VELOCITY:
            raise NotImplementedError("sample_sde covers path_type='Linear' with prediction='velocity'")
        if sampling_method not in ("Euler", "Heun"):
            raise NotImplementedError("Smapler type not implemented.")       # integrators.py:61 (sic)
        if last_step not in (None, "Mean", "Tweedie", "Euler"):
            raise NotImplementedError()
        if last_step is None:
            last_step_size = 0.0
        t0, t1 = self.transport.check_interval(self.transport.train_eps, self.transport.sample_eps,
                                               diffusion_form=diffusion_form, sde=True, eval=True, reverse=False,
                                               last_step_size=last_step_size)
        t_grid = th.linspace(t0, t1, num_steps)                               # integrators.py:24
        dt = t_grid[1] - t_grid[0]

        def expand(t, x):                                                     # path.py:5-13
            return t.view(-1, *([1] * (x.dim() - 1)))

        def diffusion_fn(x, t):                                               # path.py:45-68 on ICPlan (:23-43)
            t = expand(t, x)
            choices = {
                "constant": lambda: diffusion_norm,
                "SBDM": lambda: diffusion_norm * ((1 / t) * ((1 - t) ** 2) - (1 - t) * (-1)),
                "sigma": lambda: diffusion_norm * (1 - t),
                "linear": lambda: diffusion_norm * (1 - t),
                "decreasing": lambda: 0.25 * (diffusion_norm * th.cos(np.pi * t) + 1) ** 2,
                "inccreasing-decreasing": lambda: diffusion_norm * th.sin(np.pi * t) ** 2,
            }
            if diffusion_form not in choices:
                raise NotImplementedError(f"Diffusion form {diffusion_form} not implemented")
            return choices[diffusion_form]()

        def score_from_velocity(v, x, t):                                     # path.py:70-84 with alpha = t, sigma = 1 - t
            t = expand(t, x)
            sigma_t = 1 - t
            var = sigma_t ** 2 - (t / 1) * (-1) * sigma_t
            return ((t / 1) * v - x) / var

        def velocity(x, t, model, **kw):
            out = model(x, t, **kw)
            assert out.shape == x.shape, "Output shape from ODE solver must match input shape"
            return out

        def sde_drift(x, t, model, **kw):                                     # transport.py:306-308, one evaluation
            v = velocity(x, t, model, **kw)
            return v + diffusion_fn(x, t) * score_from_velocity(v, x, t)

        def em_step(x, ti, model, **kw):                                      # integrators.py:30-38
            w_cur = th.randn(x.size()).to(x)
            t = th.ones(x.size(0)).to(x) * ti
            dw = w_cur * th.sqrt(dt)
            drift = sde_drift(x, t, model, **kw)
            diffusion = diffusion_fn(x, t)
            mean_x = x + drift * dt
            return mean_x + th.sqrt(2 * diffusion) * dw

        def heun_step(x, ti, model, **kw):                                    # integrators.py:40-49
            w_cur = th.randn(x.size()).to(x)
            dw = w_cur * th.sqrt(dt)
            t_cur = th.ones(x.size(0)).to(x) * ti
            diffusion = diffusion_fn(x, t_cur)
            xhat = x + th.sqrt(2 * diffusion) * dw
            k1 = sde_drift(xhat, t_cur, model, **kw)
            xp = xhat + dt * k1
            k2 = sde_drift(xp, t_cur + dt, model, **kw)
            return xhat + 0.5 * dt * (k1 + k2)

        # ---- native fused update (CUDA states): same arithmetic, one kernel per stage instead of ~10 tensor ops
        from . import engine as _eng

        def coeffs(ti, device):
            """fp32 scalars of a step, computed with the same torch ops, on the same device, as the tensor path
            applies elementwise (cos / sin of the 'decreasing' forms may differ in the last bit between CPU and GPU)."""
            t = th.ones(1, device=device) * ti
            diffusion = diffusion_fn(t.view(1, 1), t)
            sigma_t = 1 - t
            var = sigma_t ** 2 - (t / 1) * (-1) * sigma_t
            # form "constant" yields a python float and th.sqrt then raises, exactly as in the reference (integrators.py:37)
            return dict(t=float(t), diffusion=float(diffusion), var=float(var),
                        sqrt_2diffusion=float(th.sqrt(2 * diffusion)), sqrt_dt=float(th.sqrt(dt)))

        def native(x):   # VAVAE_SDE_TENSOR_OPS=1: developer knob, keeps the tensor-op path on the GPU (A/B in the tests)
            return x.is_cuda and x.dtype == th.float32 and os.environ.get("VAVAE_SDE_TENSOR_OPS") != "1"

        def em_step_native(x, ti, model, **kw):
            w_cur = th.randn(x.size()).to(x)                                  # same RNG stream as the reference
            t = th.ones(x.size(0)).to(x) * ti
            v = velocity(x, t, model, **kw)
            return _eng.sde_update(0, x.contiguous(), th.empty_like(x), v=v.contiguous(), w=w_cur, dt=float(dt), **coeffs(ti, x.device))

        def heun_step_native(x, ti, model, **kw):
            w_cur = th.randn(x.size()).to(x)
            t_cur = th.ones(x.size(0)).to(x) * ti
            c0, c1 = coeffs(ti, x.device), coeffs(float(th.ones(1) * ti + dt), x.device)
            xhat = _eng.sde_update(1, x.contiguous(), th.empty_like(x), w=w_cur, dt=float(dt), **c0)
            v1 = velocity(xhat, t_cur, model, **kw)
            k1 = th.empty_like(x)
            xp = _eng.sde_update(2, xhat, th.empty_like(x), v=v1.contiguous(), out2=k1, dt=float(dt), **c0)
            v2 = velocity(xp, t_cur + dt, model, **kw)
            return _eng.sde_update(3, xp, th.empty_like(x), v=v2.contiguous(), k1=k1, xh=xhat, dt=float(dt), **c1)

        def step(x, ti, model, **kw):
            if native(x):
                return (em_step_native if sampling_method == "Euler" else heun_step_native)(x, ti, model, **kw)
            return (em_step if sampling_method == "Euler" else heun_step)(x, ti, model, **kw)

        def last_step_fn(x, t, model, **kw):                                  # transport.py:313-334
            if last_step is None:
                return x
            if native(x):
                c = coeffs(t1, x.device)
                v = velocity(x, t, model, **kw).contiguous()
                if last_step == "Mean":
                    return _eng.sde_update(4, x.contiguous(), th.empty_like(x), v=v, dt=float(last_step_size), **c)
                if last_step == "Tweedie":
                    alpha, sigma = t[0], 1 - t[0]
                    return _eng.sde_update(5, x.contiguous(), th.empty_like(x), v=v, dt=0.0, c0=float(alpha),
                                           c1=float((sigma ** 2) / alpha), **c)
                return _eng.sde_update(6, x.contiguous(), th.empty_like(x), v=v, dt=float(last_step_size), **c)
            if last_step == "Mean":
                return x + sde_drift(x, t, model, **kw) * last_step_size
            if last_step == "Tweedie":
                alpha, sigma = t[0], 1 - t[0]                                 # alpha(t)[0][0], sigma(t)[0][0]
                v = velocity(x, t, model, **kw)
                return x / alpha + (sigma ** 2) / alpha * score_from_velocity(v, x, t)
            return x + velocity(x, t, model, **kw) * last_step_size           # "Euler"

        def _sample(init, model, **model_kwargs):
            x = init
            xs = []
            with th.no_grad():
                for ti in t_grid[:-1]:                                        # integrators.py:69-73
                    x = step(x, ti, model, **model_kwargs)
                    xs.append(x)
                ts = th.ones(init.size(0), device=init.device) * t1
                xs.append(last_step_fn(xs[-1], ts, model, **model_kwargs))
            assert len(xs) == num_steps, "Samples does not match the number of steps"
            return xs

        return _sample

    def sample_ode_likelihood(self, *a, **k):
        raise NotImplementedError("likelihood evaluation is out of scope for the B200 ODE sampling engine")

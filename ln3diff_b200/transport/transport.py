import enum

import torch as th

from .. import ops


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


class SNRType(enum.Enum):
    UNIFORM = enum.auto()
    LOGNORM = enum.auto()


class Transport:
    """reference transport/transport.py:48-72: only the Linear path with velocity prediction (the
    I23D release configuration, nsr/lsgm/flow_matching_trainer.py:160-192) is on the hot path."""

    def __init__(self, *, model_type, path_type, loss_type, train_eps, sample_eps, snr_type):
        if path_type != PathType.LINEAR or model_type != ModelType.VELOCITY:
            raise NotImplementedError("ln3diff_b200 implements the Linear path / velocity prediction")
        self.loss_type, self.model_type, self.path_type = loss_type, model_type, path_type
        self.train_eps, self.sample_eps, self.snr_type = train_eps, sample_eps, snr_type
        assert self.snr_type == SNRType.LOGNORM, "use lognorm schedule plz."

    def check_interval(self, train_eps, sample_eps, *, diffusion_form="SBDM", sde=False, reverse=False,
                       eval=False, last_step_size=0.0):
        t0, t1 = 0, 1  # velocity & Linear: stable everywhere (transport.py:85-112)
        if sde:
            raise NotImplementedError("SDE sampling is outside the hot path")
        return (1 - t0, 1 - t1) if reverse else (t0, t1)

    def get_drift(self):
        def velocity_ode(x, t, model, **model_kwargs):
            out = model(x, t, **model_kwargs)
            assert out.shape == x.shape, "Output shape from ODE solver must match input shape"
            return out
        return velocity_ode


def create_transport(path_type="Linear", prediction="velocity", loss_weight=None, train_eps=None,
                     sample_eps=None, snr_type="uniform"):
    """reference transport/__init__.py:3-72."""
    model_type = {"noise": ModelType.NOISE, "score": ModelType.SCORE}.get(prediction, ModelType.VELOCITY)
    loss_type = {"velocity": WeightType.VELOCITY, "likelihood": WeightType.LIKELIHOOD}.get(loss_weight, WeightType.NONE)
    if snr_type == "lognorm":
        snr = SNRType.LOGNORM
    elif snr_type == "uniform":
        snr = SNRType.UNIFORM
    else:
        raise ValueError(f"Invalid snr type {snr_type}")
    ptype = {"Linear": PathType.LINEAR, "GVP": PathType.GVP, "VP": PathType.VP}[path_type]
    return Transport(model_type=model_type, path_type=ptype, loss_type=loss_type, train_eps=0, sample_eps=0,
                     snr_type=snr)


class ode:
    """reference transport/integrators.py:78-120 with the fixed-grid solvers of torchdiffeq."""

    def __init__(self, drift, *, t0, t1, sampler_type, num_steps, atol, rtol):
        assert t0 < t1, "ODE sampler has to be in forward time"
        self.drift = drift
        self.t = th.linspace(t0, t1, num_steps)
        self.atol, self.rtol, self.sampler_type = atol, rtol, sampler_type

    def _axpy(self, x, dt, k, coef=None):
        """x + dt * k as one fused launch on CUDA (`coef`: the step's device-resident (B, 4) row block)."""
        if coef is not None:
            return ops.sampler_affine_update(x.contiguous(), coef, k.float().contiguous())
        return x + dt * k

    def sample(self, x, model, **model_kwargs):
        device = x.device
        B = x.size(0)

        def _fn(t, x):
            # `th.ones(B).to(device) * t` of the reference (integrators.py:104-107) without the H2D copy / sync
            tt = th.full((B,), float(t), device=device, dtype=th.float32)
            return self.drift(x, tt, model, **model_kwargs)

        t = self.t                                                 # float32 grid, kept on the HOST
        if self.sampler_type in ("euler", "heun", "midpoint"):
            n = len(t) - 1
            dts = t[1:] - t[:-1]                                   # float32 arithmetic as the reference's tensors
            halves = 0.5 * dts
            fused = x.is_cuda and x.dtype == th.float32
            tab = None
            if fused:   # every step's update coefficients in one upload: [:, 0] = dt, [:, 1] = dt / 2
                tab = th.zeros(n, 2, B, 4)
                tab[..., 0] = 1.0
                tab[:, 0, :, 1] = dts[:, None]
                tab[:, 1, :, 1] = halves[:, None]
                tab = tab.to(device)
            ys = [x]
            for i in range(n):
                t0, t1, dt, half = t[i], t[i + 1], dts[i], halves[i]
                cf = tab[i, 0] if fused else None
                ch = tab[i, 1] if fused else None
                y = ys[-1]
                if self.sampler_type == "euler":
                    y = self._axpy(y, dt, _fn(t0, y), cf)
                elif self.sampler_type == "midpoint":
                    y = self._axpy(y, dt, _fn(t0 + half, self._axpy(y, half, _fn(t0, y), ch)), cf)
                else:
                    k1 = _fn(t0, y)
                    k2 = _fn(t1, self._axpy(y, dt, k1, cf))
                    y = self._axpy(self._axpy(y, half, k1, ch), half, k2, ch)
                ys.append(y)
            return th.stack(ys, 0)
        try:
            from torchdiffeq import odeint  # the reference's own third-party solver, when it is installed
        except ImportError:
            odeint = None
        if odeint is not None:
            return odeint(_fn, x, t.to(device), method=self.sampler_type, atol=[self.atol], rtol=[self.rtol])
        if self.sampler_type == "dopri5":
            # the shipped I23D default (transport.py:374-381): restated solver, see transport/dopri5.py
            from .dopri5 import odeint_dopri5
            self.last_stats = {}
            return odeint_dopri5(_fn, x, t, rtol=self.rtol, atol=self.atol, stats=self.last_stats)
        raise NotImplementedError(
            f"sampling_method='{self.sampler_type}' needs torchdiffeq (un-vendored dependency of the reference); "
            "'dopri5' and the fixed-grid 'euler' / 'heun' / 'midpoint' solvers are built in")


class Sampler:
    """reference transport/transport.py:246-259,374-421."""

    def __init__(self, transport):
        self.transport = transport
        self.drift = self.transport.get_drift()

    def sample_ode(self, *, sampling_method="dopri5", num_steps=50, atol=1e-6, rtol=1e-3, reverse=False,
                   cfg=False):
        if reverse:
            drift = lambda x, t, model, **kw: self.drift(x, th.ones_like(t) * (1 - t), model, **kw)
        else:
            drift = self.drift
        t0, t1 = self.transport.check_interval(self.transport.train_eps, self.transport.sample_eps, sde=False,
                                               eval=True, reverse=reverse, last_step_size=0.0)
        return ode(drift=drift, t0=t0, t1=t1, sampler_type=sampling_method, num_steps=num_steps, atol=atol,
                   rtol=rtol).sample

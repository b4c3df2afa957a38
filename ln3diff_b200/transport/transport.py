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

    def _axpy(self, x, dt, k):
        """x + dt * k as one fused launch on CUDA."""
        if x.is_cuda and x.dtype == th.float32:
            B = x.shape[0]
            coef = th.tensor([[1.0, float(dt), 0.0, 0.0]], device=x.device).repeat(B, 1)
            return ops.sampler_affine_update(x.contiguous(), coef, k.float().contiguous())
        return x + dt * k

    def sample(self, x, model, **model_kwargs):
        device = x.device

        def _fn(t, x):
            tt = th.ones(x.size(0)).to(device) * t
            return self.drift(x, tt, model, **model_kwargs)

        t = self.t.to(device)
        if self.sampler_type in ("euler", "heun", "midpoint"):
            ys = [x]
            for i in range(len(t) - 1):
                t0, t1 = t[i], t[i + 1]
                dt = t1 - t0
                y = ys[-1]
                if self.sampler_type == "euler":
                    y = self._axpy(y, dt, _fn(t0, y))
                elif self.sampler_type == "midpoint":
                    half = 0.5 * dt
                    y = self._axpy(y, dt, _fn(t0 + half, self._axpy(y, half, _fn(t0, y))))
                else:
                    k1 = _fn(t0, y)
                    k2 = _fn(t1, self._axpy(y, dt, k1))
                    y = self._axpy(self._axpy(y, 0.5 * dt, k1), 0.5 * dt, k2)
                ys.append(y)
            return th.stack(ys, 0)
        try:
            from torchdiffeq import odeint  # third-party adaptive solvers (dopri5, ...)
        except ImportError as e:
            raise NotImplementedError(
                f"sampling_method='{self.sampler_type}' needs torchdiffeq (un-vendored dependency of "
                "the reference); fixed-grid 'euler' / 'heun' / 'midpoint' are built in") from e
        return odeint(_fn, x, t, method=self.sampler_type, atol=[self.atol], rtol=[self.rtol])


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

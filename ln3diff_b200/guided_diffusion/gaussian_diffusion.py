"""Mirror of reference guided_diffusion/gaussian_diffusion.py (sampling side).

Schedules are float64 numpy exactly as the reference builds them (:20-51, :153-204).  The per-step
update of `p_sample` (:498-546) -- x0-from-eps (:422-427), posterior mean (:252-271), fixed
variance noise -- is affine in (x, model_output, noise), so on CUDA it is ONE launch of
ln3_sampler_affine_update with per-sample coefficients gathered from device-resident tables
(the reference re-uploads whole numpy tables 6-8x per step through _extract_into_tensor, :1240-1253).
"""
import enum
import math

import numpy as np
import torch as th

from .. import ops


class ModelMeanType(enum.Enum):
    PREVIOUS_X = enum.auto()
    START_X = enum.auto()
    EPSILON = enum.auto()
    V = enum.auto()


class ModelVarType(enum.Enum):
    LEARNED = enum.auto()
    FIXED_SMALL = enum.auto()
    FIXED_LARGE = enum.auto()
    LEARNED_RANGE = enum.auto()


class LossType(enum.Enum):
    MSE = enum.auto()
    RESCALED_MSE = enum.auto()
    KL = enum.auto()
    RESCALED_KL = enum.auto()


def get_named_beta_schedule(schedule_name, num_diffusion_timesteps):
    """reference :20-51."""
    if schedule_name == "linear":
        scale = 1000 / num_diffusion_timesteps
        return np.linspace(scale * 0.0001, scale * 0.02, num_diffusion_timesteps, dtype=np.float64)
    if schedule_name == "cosine":
        f = lambda t: math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2
        n = num_diffusion_timesteps
        return np.array([min(1 - f((i + 1) / n) / f(i / n), 0.999) for i in range(n)])
    raise NotImplementedError(f"unknown beta schedule: {schedule_name}")


def _extract_into_tensor(arr, timesteps, broadcast_shape):
    res = th.from_numpy(arr).to(device=timesteps.device)[timesteps].float()
    while len(res.shape) < len(broadcast_shape):
        res = res[..., None]
    return res.expand(broadcast_shape)


class GaussianDiffusion:
    """reference :125-204 constructor contract (keyword-only)."""

    def __init__(self, *, betas, model_mean_type, model_var_type, loss_type, rescale_timesteps=False,
                 standarization_xt=False):
        self.model_mean_type, self.model_var_type, self.loss_type = model_mean_type, model_var_type, loss_type
        self.rescale_timesteps, self.standarization_xt = rescale_timesteps, standarization_xt
        betas = np.array(betas, dtype=np.float64)
        self.betas = betas
        assert len(betas.shape) == 1, "betas must be 1-D"
        assert (betas > 0).all() and (betas <= 1).all()
        self.num_timesteps = int(betas.shape[0])
        alphas = 1.0 - betas
        self.alphas_cumprod = np.cumprod(alphas, axis=0)
        self.alphas_cumprod_prev = np.append(1.0, self.alphas_cumprod[:-1])
        self.alphas_cumprod_next = np.append(self.alphas_cumprod[1:], 0.0)
        self.sqrt_alphas_cumprod = np.sqrt(self.alphas_cumprod)
        self.sqrt_one_minus_alphas_cumprod = np.sqrt(1.0 - self.alphas_cumprod)
        self.log_one_minus_alphas_cumprod = np.log(1.0 - self.alphas_cumprod)
        self.sqrt_recip_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod)
        self.sqrt_recipm1_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod - 1)
        self.posterior_variance = betas * (1.0 - self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_log_variance_clipped = np.log(np.append(self.posterior_variance[1], self.posterior_variance[1:]))
        self.posterior_mean_coef1 = betas * np.sqrt(self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_mean_coef2 = (1.0 - self.alphas_cumprod_prev) * np.sqrt(alphas) / (1.0 - self.alphas_cumprod)
        self._coef_cache = {}

    def _scale_timesteps(self, t):
        if self.rescale_timesteps:
            return t.float() * (1000.0 / self.num_timesteps)
        return t

    # -- affine coefficients of p_sample: sample = a x + w0 out + s noise
    def _step_coef_table(self, device):
        key = str(device)
        tab = self._coef_cache.get(key)
        if tab is None:
            if self.model_var_type == ModelVarType.FIXED_LARGE:
                logvar = np.log(np.append(self.posterior_variance[1], self.betas[1:]))
            elif self.model_var_type == ModelVarType.FIXED_SMALL:
                logvar = self.posterior_log_variance_clipped
            else:
                raise NotImplementedError("learned variance is not on the DiT path (learn_sigma=False)")
            c1, c2 = self.posterior_mean_coef1, self.posterior_mean_coef2
            if self.model_mean_type == ModelMeanType.EPSILON:
                a = c1 * self.sqrt_recip_alphas_cumprod + c2
                w0 = -c1 * self.sqrt_recipm1_alphas_cumprod
            elif self.model_mean_type == ModelMeanType.START_X:
                a, w0 = c2, c1
            else:
                self._unsupported_mean_type()
            s = np.exp(0.5 * logvar)
            s[0] = 0.0  # nonzero_mask: no noise at t == 0
            tab = th.tensor(np.stack([a, w0, np.zeros_like(a), s], 1), dtype=th.float32, device=device)
            self._coef_cache[key] = tab
        return tab

    def _unsupported_mean_type(self):
        """Both p_sample paths reject the same configurations.  ModelMeanType.V only works together with
        mixing_normal=True in the reference (the v output becomes eps inside the mixing branch, :340-343;
        without it `assert v_transformed_to_eps_flag` fails, :405-406), and the LSGM mixing path is outside
        the DiT hot path; PREVIOUS_X is unused by every release config."""
        raise NotImplementedError(f"{self.model_mean_type}: the DiT path runs EPSILON / START_X with fixed variance "
                                  "(V needs mixing_normal=True in the reference, which is not mirrored)")

    def _pred_xstart_coef_table(self, device):
        """(T, 4) rows [a, w, 0, 0] with pred_xstart = a x + w out (EPSILON: :422-427; START_X: the output)."""
        key = ("x0", str(device))
        tab = self._coef_cache.get(key)
        if tab is None:
            if self.model_mean_type == ModelMeanType.EPSILON:
                a, w = self.sqrt_recip_alphas_cumprod, -self.sqrt_recipm1_alphas_cumprod
            elif self.model_mean_type == ModelMeanType.START_X:
                a, w = np.zeros_like(self.betas), np.ones_like(self.betas)
            else:
                self._unsupported_mean_type()
            z = np.zeros_like(a)
            tab = th.tensor(np.stack([a, w, z, z], 1), dtype=th.float32, device=device)
            self._coef_cache[key] = tab
        return tab

    def p_mean_variance(self, model, x, t, c=None, clip_denoised=True, denoised_fn=None,
                        model_kwargs=None, mixing_normal=False, direct_return_model_output=False):
        """reference :273-420 for EPSILON / START_X / V with fixed variance (eager torch, API parity)."""
        model_kwargs = model_kwargs or {}
        assert not mixing_normal, "mixing_normal (LSGM) is outside the DiT hot path"
        model_output = model(x, self._scale_timesteps(t), c=c, mixing_normal=mixing_normal, **model_kwargs)
        if direct_return_model_output:
            return model_output
        var, logvar = {
            ModelVarType.FIXED_LARGE: (np.append(self.posterior_variance[1], self.betas[1:]),
                                       np.log(np.append(self.posterior_variance[1], self.betas[1:]))),
            ModelVarType.FIXED_SMALL: (self.posterior_variance, self.posterior_log_variance_clipped),
        }[self.model_var_type]
        model_variance = _extract_into_tensor(var, t, x.shape)
        model_log_variance = _extract_into_tensor(logvar, t, x.shape)

        def process_xstart(v):
            if denoised_fn is not None:
                v = denoised_fn(v)
            return v.clamp(-1, 1) if clip_denoised else v

        if self.model_mean_type == ModelMeanType.START_X:
            pred_xstart = process_xstart(model_output)
        elif self.model_mean_type == ModelMeanType.EPSILON:
            pred_xstart = process_xstart(self._predict_xstart_from_eps(x, t, model_output))
        else:
            self._unsupported_mean_type()
        model_mean = (_extract_into_tensor(self.posterior_mean_coef1, t, x.shape) * pred_xstart +
                      _extract_into_tensor(self.posterior_mean_coef2, t, x.shape) * x)
        return {"mean": model_mean, "variance": model_variance, "log_variance": model_log_variance,
                "pred_xstart": pred_xstart}

    def _predict_xstart_from_eps(self, x_t, t, eps):
        return (_extract_into_tensor(self.sqrt_recip_alphas_cumprod, t, x_t.shape) * x_t -
                _extract_into_tensor(self.sqrt_recipm1_alphas_cumprod, t, x_t.shape) * eps)

    def _wrap_model(self, model):
        return model

    def p_sample(self, model, x, t, cond=None, clip_denoised=True, denoised_fn=None, cond_fn=None,
                 model_kwargs=None, mixing_normal=False):
        """reference :498-546.  Fused on CUDA when nothing needs pred_xstart post-processing."""
        fused = (x.is_cuda and x.dtype == th.float32 and not clip_denoised and denoised_fn is None
                 and cond_fn is None and not mixing_normal)
        if fused:
            model_kwargs = model_kwargs or {}
            out = self._wrap_model(model)(x, self._scale_timesteps(t), c=cond, mixing_normal=False, **model_kwargs)
            noise = th.randn_like(x)
            xc, oc = x.contiguous(), out.float().contiguous()
            sample = ops.sampler_affine_update(xc, self._step_coef_table(x.device)[t].contiguous(), oc, None, noise)
            # the reference always returns the x_0 estimate (progressive loops read it): one more affine launch
            pred_xstart = ops.sampler_affine_update(xc, self._pred_xstart_coef_table(x.device)[t].contiguous(), oc)
            return {"sample": sample, "pred_xstart": pred_xstart}
        assert cond_fn is None, "classifier guidance (cond_fn) is not on the hot path"
        out = self.p_mean_variance(model, x, t, c=cond, clip_denoised=clip_denoised, denoised_fn=denoised_fn,
                                   model_kwargs=model_kwargs, mixing_normal=mixing_normal)
        noise = th.randn_like(x)
        nonzero_mask = (t != 0).float().view(-1, *([1] * (len(x.shape) - 1)))
        sample = out["mean"] + nonzero_mask * th.exp(0.5 * out["log_variance"]) * noise
        return {"sample": sample, "pred_xstart": out["pred_xstart"]}

    def p_sample_loop(self, model, shape, cond=None, noise=None, clip_denoised=True, denoised_fn=None,
                      cond_fn=None, model_kwargs=None, device=None, progress=False, mixing_normal=False):
        """reference :627-672."""
        final = None
        for sample in self.p_sample_loop_progressive(model, shape, cond=cond, noise=noise,
                                                     clip_denoised=clip_denoised, denoised_fn=denoised_fn,
                                                     cond_fn=cond_fn, model_kwargs=model_kwargs, device=device,
                                                     progress=progress, mixing_normal=mixing_normal):
            final = sample
        return final["sample"]

    def p_sample_loop_progressive(self, model, shape, cond=None, noise=None, clip_denoised=True,
                                  denoised_fn=None, cond_fn=None, model_kwargs=None, device=None,
                                  progress=False, mixing_normal=False):
        """reference :674-727."""
        if device is None:
            device = noise.device if noise is not None else th.device("cuda")
        assert isinstance(shape, (tuple, list))
        img = noise if noise is not None else th.randn(*shape, device=device)
        for i in list(range(self.num_timesteps))[::-1]:
            t = th.tensor([i] * shape[0], device=device)
            with th.no_grad():
                out = self.p_sample(model, img, t, cond=cond, clip_denoised=clip_denoised,
                                    denoised_fn=denoised_fn, cond_fn=cond_fn, model_kwargs=model_kwargs,
                                    mixing_normal=mixing_normal)
                yield out
                img = out["sample"]

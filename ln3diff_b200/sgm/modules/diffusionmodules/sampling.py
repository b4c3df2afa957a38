"""Mirror of reference sgm/modules/diffusionmodules/sampling.py:21-130,211-215 (Euler-EDM).

`EulerEDMSampler(...)(denoiser, x, cond, uc)` keeps the reference loop.  The per-step elementwise
tail -- VanillaCFG combine (guiders.py:28-31), to_d (sampling_utils.py:34-35) and the Euler step
(sampling.py:78-79) -- is one ln3_sampler_affine_update launch on CUDA tensors:
    x' = (1 + dt/s) x - (dt/s)(1 - g) D_u - (dt/s) g D_c,   dt = s_next - s, g = cfg scale.
"""
import torch

from .... import ops
from ...util import append_dims, default, instantiate_from_config
from .guiders import IdentityGuider, VanillaCFG

DEFAULT_GUIDER = {"target": "sgm.modules.diffusionmodules.guiders.IdentityGuider"}


def to_d(x, sigma, denoised):
    return (x - denoised) / append_dims(sigma, x.ndim)


class BaseDiffusionSampler:
    def __init__(self, discretization_config, num_steps=None, guider_config=None, verbose=False,
                 device="cuda"):
        self.num_steps = num_steps
        self.discretization = instantiate_from_config(discretization_config)
        self.guider = instantiate_from_config(default(guider_config, DEFAULT_GUIDER))
        self.verbose = verbose
        self.device = device

    def prepare_sampling_loop(self, x, cond, uc=None, num_steps=None):
        sigmas = self.discretization(self.num_steps if num_steps is None else num_steps, device=self.device)
        uc = default(uc, cond)
        x *= torch.sqrt(1.0 + sigmas[0] ** 2.0)
        num_sigmas = len(sigmas)
        s_in = x.new_ones([x.shape[0]])
        return x, s_in, sigmas, num_sigmas, cond, uc

    def denoise(self, x, denoiser, sigma, cond, uc):
        denoised = denoiser(*self.guider.prepare_inputs(x, sigma, cond, uc))
        return self.guider(denoised, sigma)

    def get_sigma_gen(self, num_sigmas):
        return range(num_sigmas - 1)


class SingleStepDiffusionSampler(BaseDiffusionSampler):
    def euler_step(self, x, d, dt):
        return x + dt * d


class EDMSampler(SingleStepDiffusionSampler):
    def __init__(self, s_churn=0.0, s_tmin=0.0, s_tmax=float("inf"), s_noise=1.0, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.s_churn, self.s_tmin, self.s_tmax, self.s_noise = s_churn, s_tmin, s_tmax, s_noise

    def sampler_step(self, sigma, next_sigma, denoiser, x, cond, uc=None, gamma=0.0):
        sigma_hat = sigma * (gamma + 1.0)
        if gamma > 0:
            eps = torch.randn_like(x) * self.s_noise
            x = x + eps * append_dims(sigma_hat ** 2 - sigma ** 2, x.ndim) ** 0.5
        if x.is_cuda and x.dtype == torch.float32 and isinstance(self.guider, (VanillaCFG, IdentityGuider)):
            den = denoiser(*self.guider.prepare_inputs(x, sigma_hat, cond, uc)).float().contiguous()
            r = (next_sigma - sigma_hat) / sigma_hat          # dt / sigma, per sample
            if isinstance(self.guider, VanillaCFG):
                g = self.guider.scale
                d_u, d_c = den.chunk(2)
                coef = torch.stack([1 + r, -r * (1 - g), -r * g, torch.zeros_like(r)], 1)
                x_next = ops.sampler_affine_update(x.contiguous(), coef.float().contiguous(), d_u, d_c)
            else:
                coef = torch.stack([1 + r, -r, torch.zeros_like(r), torch.zeros_like(r)], 1)
                x_next = ops.sampler_affine_update(x.contiguous(), coef.float().contiguous(), den)
            return self.possible_correction_step(x_next, x, None, None, next_sigma, denoiser, cond, uc)
        denoised = self.denoise(x, denoiser, sigma_hat, cond, uc)
        d = to_d(x, sigma_hat, denoised)
        dt = append_dims(next_sigma - sigma_hat, x.ndim)
        euler_step = self.euler_step(x, d, dt)
        return self.possible_correction_step(euler_step, x, d, dt, next_sigma, denoiser, cond, uc)

    def __call__(self, denoiser, x, cond, uc=None, num_steps=None):
        x, s_in, sigmas, num_sigmas, cond, uc = self.prepare_sampling_loop(x, cond, uc, num_steps)
        for i in self.get_sigma_gen(num_sigmas):
            gamma = (min(self.s_churn / (num_sigmas - 1), 2 ** 0.5 - 1)
                     if self.s_tmin <= sigmas[i] <= self.s_tmax else 0.0)
            x = self.sampler_step(s_in * sigmas[i], s_in * sigmas[i + 1], denoiser, x, cond, uc, gamma)
        return x


class EulerEDMSampler(EDMSampler):
    def possible_correction_step(self, euler_step, x, d, dt, next_sigma, denoiser, cond, uc):
        return euler_step

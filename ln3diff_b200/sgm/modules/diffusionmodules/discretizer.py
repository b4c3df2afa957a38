"""Mirror of reference sgm/modules/diffusionmodules/discretizer.py:15-69 (host-side, float64
numpy tables -> float32 sigmas; one-time work, no kernel)."""
from functools import partial

import numpy as np
import torch

from ...util import append_zero


def make_beta_schedule(schedule, n_timestep, linear_start=1e-4, linear_end=2e-2):
    """sgm/modules/diffusionmodules/util.py:20-33 ("linear" is sqrt-linear)."""
    assert schedule == "linear"
    return (torch.linspace(linear_start ** 0.5, linear_end ** 0.5, n_timestep, dtype=torch.float64) ** 2).numpy()


def generate_roughly_equally_spaced_steps(num_substeps, max_step):
    return np.linspace(max_step - 1, 0, num_substeps, endpoint=False).astype(int)[::-1]


class Discretization:
    def __call__(self, n, do_append_zero=True, device="cpu", flip=False):
        sigmas = self.get_sigmas(n, device=device)
        sigmas = append_zero(sigmas) if do_append_zero else sigmas
        return sigmas if not flip else torch.flip(sigmas, (0,))


class LegacyDDPMDiscretization(Discretization):
    def __init__(self, linear_start=0.00085, linear_end=0.0120, num_timesteps=1000):
        self.num_timesteps = num_timesteps
        betas = make_beta_schedule("linear", num_timesteps, linear_start=linear_start, linear_end=linear_end)
        self.alphas_cumprod = np.cumprod(1.0 - betas, axis=0)

    def get_sigmas(self, n, device="cpu"):
        if n < self.num_timesteps:
            alphas_cumprod = self.alphas_cumprod[generate_roughly_equally_spaced_steps(n, self.num_timesteps)]
        elif n == self.num_timesteps:
            alphas_cumprod = self.alphas_cumprod
        else:
            raise ValueError
        to_torch = partial(torch.tensor, dtype=torch.float32, device=device)
        return torch.flip(to_torch((1 - alphas_cumprod) / alphas_cumprod) ** 0.5, (0,))


class EDMDiscretization(Discretization):
    def __init__(self, sigma_min=0.002, sigma_max=80.0, rho=7.0):
        self.sigma_min, self.sigma_max, self.rho = sigma_min, sigma_max, rho

    def get_sigmas(self, n, device="cpu"):
        ramp = torch.linspace(0, 1, n, device=device)
        min_inv_rho, max_inv_rho = self.sigma_min ** (1 / self.rho), self.sigma_max ** (1 / self.rho)
        return (max_inv_rho + ramp * (min_inv_rho - max_inv_rho)) ** self.rho

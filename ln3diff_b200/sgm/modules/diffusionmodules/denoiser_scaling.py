"""Mirror of reference sgm/modules/diffusionmodules/denoiser_scaling.py (scalar tables per step)."""
import torch


class EpsScaling:
    def __call__(self, sigma):
        return torch.ones_like(sigma, device=sigma.device), -sigma, 1 / (sigma ** 2 + 1.0) ** 0.5, sigma.clone()


class VScaling:
    def __call__(self, sigma):
        c_skip = 1.0 / (sigma ** 2 + 1.0)
        c_out = -sigma / (sigma ** 2 + 1.0) ** 0.5
        return c_skip, c_out, 1.0 / (sigma ** 2 + 1.0) ** 0.5, sigma.clone()


class EDMScaling:
    def __init__(self, sigma_data=0.5):
        self.sigma_data = sigma_data

    def __call__(self, sigma):
        sd = self.sigma_data
        return (sd ** 2 / (sigma ** 2 + sd ** 2), sigma * sd / (sigma ** 2 + sd ** 2) ** 0.5,
                1 / (sigma ** 2 + sd ** 2) ** 0.5, 0.25 * sigma.log())

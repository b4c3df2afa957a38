"""Mirror of reference sgm/modules/diffusionmodules/denoiser.py:13-78.

`Denoiser.forward(network, input, sigma, cond)` keeps the reference contract
(`network(input * c_in, c_noise, cond) * c_out + input * c_skip`).  When the network is one of this
package's DiTs the `input * c_in` pre-scale is folded into the patch-embed kernel (in_scale) and
the output combine runs as one fused launch (ln3_sampler_affine_update)."""
import torch
import torch.nn as nn

from .... import ops
from ...util import append_dims, instantiate_from_config


class Denoiser(nn.Module):
    def __init__(self, scaling_config):
        super().__init__()
        self.scaling = instantiate_from_config(scaling_config)

    def possibly_quantize_sigma(self, sigma):
        return sigma

    def possibly_quantize_c_noise(self, c_noise):
        return c_noise

    def forward(self, network, input, sigma, cond, **additional_model_inputs):
        sigma = self.possibly_quantize_sigma(sigma)
        sigma_shape = sigma.shape
        sigma_nd = append_dims(sigma, input.ndim)
        c_skip, c_out, c_in, c_noise = self.scaling(sigma_nd)
        c_noise = self.possibly_quantize_c_noise(c_noise.reshape(sigma_shape))
        if getattr(network, "_ln3_fused_in_scale", False) and input.is_cuda:
            B = input.shape[0]
            net = network(input, c_noise, cond, in_scale=c_in.reshape(B).float().contiguous(),
                          **additional_model_inputs)
            coef = torch.stack([c_skip.reshape(B), c_out.reshape(B), torch.zeros_like(sigma),
                                torch.zeros_like(sigma)], 1).float().contiguous()
            return ops.sampler_affine_update(input.float().contiguous(), coef, net)
        return network(input * c_in, c_noise, cond, **additional_model_inputs) * c_out + input * c_skip


class DiscreteDenoiser(Denoiser):
    def __init__(self, scaling_config, num_idx, discretization_config, do_append_zero=False,
                 quantize_c_noise=True, flip=True):
        super().__init__(scaling_config)
        self.discretization = instantiate_from_config(discretization_config)
        sigmas = self.discretization(num_idx, do_append_zero=do_append_zero, flip=flip)
        self.register_buffer("sigmas", sigmas)
        self.quantize_c_noise = quantize_c_noise
        self.num_idx = num_idx

    def sigma_to_idx(self, sigma):
        dists = sigma - self.sigmas[:, None]
        return dists.abs().argmin(dim=0).view(sigma.shape)

    def idx_to_sigma(self, idx):
        return self.sigmas[idx]

    def possibly_quantize_sigma(self, sigma):
        return self.idx_to_sigma(self.sigma_to_idx(sigma))

    def possibly_quantize_c_noise(self, c_noise):
        return self.sigma_to_idx(c_noise) if self.quantize_c_noise else c_noise

"""TEST INFRASTRUCTURE ONLY (oracle) -- fp32 CPU restatement of the three sampling engines of the
reference (float64 numpy schedule tables, float32 tensors).  Only tests/, smoke() and bench.py's
CPU legs may import this module.

  Euler-EDM + DiscreteDenoiser(EpsScaling) + VanillaCFG  (the shipped T23D path)
      sgm/modules/diffusionmodules/sampling.py:41-130, denoiser.py:13-78,
      denoiser_scaling.py:29-37, discretizer.py:42-69, guiders.py:24-42, sampling_utils.py:34-35,
      sgm/modules/diffusionmodules/util.py:20-33 (sqrt-linear beta schedule)
  DDPM ancestral sampling with respacing (p_sample_loop)
      guided_diffusion/gaussian_diffusion.py:20-39,153-204,273-427,498-546,627-727,
      guided_diffusion/respace.py:8-61,73-136
  Flow-matching probability-flow ODE, fixed grid + CFG
      transport/transport.py:193-225,374-421, transport/integrators.py:78-120,
      dit/dit_i23d.py:155-168 (forward_with_cfg)
Pinned by oracle/make_golden.py against the reference's own classes (tests/golden/samplers.npz).
`model` arguments are plain callables so the engines can be pinned with a cheap toy network.
"""
from __future__ import annotations

import numpy as np
import torch


# ------------------------------------------------------------------ sgm: Euler-EDM
def sgm_alphas_cumprod(linear_start=0.00085, linear_end=0.0120, num_timesteps=1000) -> np.ndarray:
    betas = torch.linspace(linear_start ** 0.5, linear_end ** 0.5, num_timesteps,
                           dtype=torch.float64).numpy() ** 2
    return np.cumprod(1.0 - betas, axis=0)


def legacy_ddpm_sigmas(n: int, append_zero: bool = True, flip: bool = False) -> torch.Tensor:
    """LegacyDDPMDiscretization.__call__ (discretizer.py:18-22,57-69)."""
    ac = sgm_alphas_cumprod()
    if n < 1000:
        ts = np.linspace(1000 - 1, 0, n, endpoint=False).astype(int)[::-1]
        ac = ac[ts]
    elif n != 1000:
        raise ValueError
    sig = torch.flip(torch.tensor((1 - ac) / ac, dtype=torch.float32) ** 0.5, (0,))
    if append_zero:
        sig = torch.cat([sig, sig.new_zeros([1])])
    return torch.flip(sig, (0,)) if flip else sig


def sigma_to_idx(sigma: torch.Tensor, table: torch.Tensor) -> torch.Tensor:
    return (sigma - table[:, None]).abs().argmin(dim=0).view(sigma.shape)


def euler_edm_cfg_sample(network, x: torch.Tensor, cond: dict, uc: dict, num_steps: int,
                         scale: float) -> torch.Tensor:
    """EulerEDMSampler(num_steps)(denoiser, x, cond, uc) with DiscreteDenoiser(EpsScaling,
    num_idx=1000, quantize_c_noise=True) and VanillaCFG(scale).  network(x_in, idx, cond_dict)."""
    table = legacy_ddpm_sigmas(1000, append_zero=False, flip=True)  # ascending, denoiser.sigmas
    sigmas = legacy_ddpm_sigmas(num_steps)
    x = x * torch.sqrt(1.0 + sigmas[0] ** 2.0)
    s_in = x.new_ones([x.shape[0]])
    c_cat = {k: torch.cat((uc[k], cond[k]), 0) for k in cond}
    for i in range(len(sigmas) - 1):
        sigma, nxt = s_in * sigmas[i], s_in * sigmas[i + 1]
        sigma_hat = sigma * (0.0 + 1.0)
        xin, sin = torch.cat([x] * 2), torch.cat([sigma_hat] * 2)
        sq = table[sigma_to_idx(sin, table)]                      # possibly_quantize_sigma
        sq4 = sq[:, None, None, None]
        c_in = 1 / (sq4 ** 2 + 1.0) ** 0.5
        c_noise = sigma_to_idx(sq, table)                          # quantised c_noise -> int64 idx
        den = network(xin * c_in, c_noise, c_cat) * (-sq4) + xin * torch.ones_like(sq4)
        x_u, x_c = den.chunk(2)
        denoised = x_u + scale * (x_c - x_u)
        d = (x - denoised) / sigma_hat[:, None, None, None]
        x = x + (nxt - sigma_hat)[:, None, None, None] * d
    return x


# ------------------------------------------------------------------ guided_diffusion: DDPM
def linear_betas(num_steps: int = 1000) -> np.ndarray:
    scale = 1000 / num_steps
    return np.linspace(scale * 0.0001, scale * 0.02, num_steps, dtype=np.float64)


def space_timesteps(num_timesteps: int, section_counts) -> set:
    if isinstance(section_counts, str):
        if section_counts.startswith("ddim"):
            want = int(section_counts[4:])
            for i in range(1, num_timesteps):
                if len(range(0, num_timesteps, i)) == want:
                    return set(range(0, num_timesteps, i))
            raise ValueError("no integer stride")
        section_counts = [int(v) for v in section_counts.split(",")]
    size_per, extra = divmod(num_timesteps, len(section_counts))
    start, steps = 0, []
    for i, cnt in enumerate(section_counts):
        size = size_per + (1 if i < extra else 0)
        if size < cnt:
            raise ValueError("cannot divide section")
        stride = 1 if cnt <= 1 else (size - 1) / (cnt - 1)
        cur = 0.0
        for _ in range(cnt):
            steps.append(start + round(cur))
            cur += stride
        start += size
    return set(steps)


class DDPMTables:
    """GaussianDiffusion.__init__ tables for (optionally respaced) betas, float64."""

    def __init__(self, betas: np.ndarray, use_timesteps=None):
        betas = np.array(betas, dtype=np.float64)
        self.original_num_steps = len(betas)
        self.timestep_map = list(range(len(betas)))
        if use_timesteps is not None:
            ac = np.cumprod(1.0 - betas)
            last, nb, self.timestep_map = 1.0, [], []
            for i, a in enumerate(ac):
                if i in use_timesteps:
                    nb.append(1 - a / last)
                    last = a
                    self.timestep_map.append(i)
            betas = np.array(nb)
        self.betas = betas
        alphas = 1.0 - betas
        self.ac = np.cumprod(alphas)
        self.ac_prev = np.append(1.0, self.ac[:-1])
        self.sqrt_recip = np.sqrt(1.0 / self.ac)
        self.sqrt_recipm1 = np.sqrt(1.0 / self.ac - 1)
        self.post_var = betas * (1.0 - self.ac_prev) / (1.0 - self.ac)
        self.coef1 = betas * np.sqrt(self.ac_prev) / (1.0 - self.ac)
        self.coef2 = (1.0 - self.ac_prev) * np.sqrt(alphas) / (1.0 - self.ac)
        self.fixed_large_logvar = np.log(np.append(self.post_var[1], betas[1:]))
        self.num_timesteps = len(betas)


def ddpm_p_sample_loop(model, shape, tables: DDPMTables, noise: torch.Tensor,
                       step_noise: list, cond=None) -> torch.Tensor:
    """SpacedDiffusion.p_sample_loop, EPSILON / FIXED_LARGE / clip_denoised=False.
    model(x, t_float in [0,1), cond) as _WrappedModel calls apply_model_inference
    (respace.py:117-136); step_noise[k] is the randn_like drawn at loop iteration k."""
    f32 = lambda a, t: torch.from_numpy(a)[t].float()[:, None, None, None]
    img = noise
    B = shape[0]
    for k, i in enumerate(reversed(range(tables.num_timesteps))):
        t = torch.tensor([i] * B)
        new_ts = torch.tensor(tables.timestep_map, dtype=t.dtype)[t] / tables.original_num_steps
        eps = model(img, new_ts, cond)
        x0 = f32(tables.sqrt_recip, t) * img - f32(tables.sqrt_recipm1, t) * eps
        mean = f32(tables.coef1, t) * x0 + f32(tables.coef2, t) * img
        logvar = f32(tables.fixed_large_logvar, t)
        nz = (t != 0).float()[:, None, None, None]
        img = mean + nz * torch.exp(0.5 * logvar) * step_noise[k]
    return img


# ------------------------------------------------------------------ transport: flow ODE
def flow_ode_cfg_sample(model_fwd, z: torch.Tensor, context: dict, cfg_scale: float,
                        num_steps: int, method: str = "euler") -> torch.Tensor:
    """Sampler(transport).sample_ode(sampling_method=method, num_steps)(zs, forward_with_cfg, ...)
    for the Linear path / velocity prediction (t0=0, t1=1): zs = cat([z, z]), context = cat(cond,
    uncond) (cond FIRST: flow_matching_trainer.py:534-540); returns samples[-1].chunk(2)[0]."""
    x = torch.cat([z, z], 0)
    ts = torch.linspace(0, 1, num_steps)

    def drift(t, x):
        tt = torch.ones(x.size(0)) * t
        out = model_fwd(x, tt, context)
        c, u = torch.split(out, len(out) // 2, dim=0)
        half = u + cfg_scale * (c - u)
        return torch.cat([half, half], 0)

    for i in range(num_steps - 1):
        t0, t1 = ts[i], ts[i + 1]
        dt = t1 - t0
        if method == "euler":
            x = x + dt * drift(t0, x)
        elif method == "heun":
            k1 = drift(t0, x)
            x = x + 0.5 * dt * (k1 + drift(t1, x + dt * k1))
        else:
            raise NotImplementedError(method)
    return x.chunk(2)[0]

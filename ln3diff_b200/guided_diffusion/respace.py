"""Mirror of reference guided_diffusion/respace.py:8-136."""
import numpy as np
import torch as th

from .gaussian_diffusion import GaussianDiffusion


def space_timesteps(num_timesteps, section_counts):
    """reference :8-61."""
    if isinstance(section_counts, str):
        if section_counts.startswith("ddim"):
            desired_count = int(section_counts[len("ddim"):])
            for i in range(1, num_timesteps):
                if len(range(0, num_timesteps, i)) == desired_count:
                    return set(range(0, num_timesteps, i))
            raise ValueError(f"cannot create exactly {num_timesteps} steps with an integer stride")
        section_counts = [int(x) for x in section_counts.split(",")]
    size_per = num_timesteps // len(section_counts)
    extra = num_timesteps % len(section_counts)
    start_idx, all_steps = 0, []
    for i, section_count in enumerate(section_counts):
        size = size_per + (1 if i < extra else 0)
        if size < section_count:
            raise ValueError(f"cannot divide section of {size} steps into {section_count}")
        frac_stride = 1 if section_count <= 1 else (size - 1) / (section_count - 1)
        cur_idx, taken = 0.0, []
        for _ in range(section_count):
            taken.append(start_idx + round(cur_idx))
            cur_idx += frac_stride
        all_steps += taken
        start_idx += size
    return set(all_steps)


class SpacedDiffusion(GaussianDiffusion):
    """reference :64-108: retains `use_timesteps` of the base process, re-deriving betas."""

    def __init__(self, use_timesteps, **kwargs):
        self.use_timesteps = set(use_timesteps)
        self.timestep_map = []
        self.original_num_steps = len(kwargs["betas"])
        base = GaussianDiffusion(**kwargs)
        last, new_betas = 1.0, []
        for i, ac in enumerate(base.alphas_cumprod):
            if i in self.use_timesteps:
                new_betas.append(1 - ac / last)
                last = ac
                self.timestep_map.append(i)
        kwargs["betas"] = np.array(new_betas)
        super().__init__(**kwargs)

    def p_mean_variance(self, model, *args, **kwargs):
        return super().p_mean_variance(self._wrap_model(model), *args, **kwargs)

    def _wrap_model(self, model):
        if isinstance(model, _WrappedModel):
            return model
        # p_sample wraps on every step (reference respace.py:100-104 builds a new wrapper -- and re-uploads
        # the timestep map -- each time); keep the wrapper of the model being sampled so its device-resident
        # map is built once per sampling run
        cached = getattr(self, "_wrapped", None)
        if cached is None or cached.model is not model:
            cached = self._wrapped = _WrappedModel(model, self.timestep_map, self.rescale_timesteps,
                                                   self.original_num_steps)
        return cached

    def _scale_timesteps(self, t):
        return t


class _WrappedModel:
    """reference :111-136: maps respaced t to the original index, divides by the original step
    count and calls `model.apply_model_inference(x, t, c)`."""

    def __init__(self, model, timestep_map, rescale_timesteps, original_num_steps):
        self.model = model
        self.timestep_map = timestep_map
        self.rescale_timesteps = rescale_timesteps
        self.original_num_steps = original_num_steps
        self._map = {}

    def __call__(self, x, ts, c=None, mixing_normal=False, **kwargs):
        key = (str(ts.device), ts.dtype)
        m = self._map.get(key)
        if m is None:  # device-resident once (the reference rebuilds + uploads it every step)
            m = th.tensor(self.timestep_map, device=ts.device, dtype=ts.dtype)
            self._map[key] = m
        new_ts = m[ts]
        if self.rescale_timesteps:
            new_ts = new_ts.float() * (1000.0 / self.original_num_steps)
        new_ts = new_ts / self.original_num_steps
        assert not mixing_normal
        return self.model.apply_model_inference(x, new_ts, c, **kwargs)

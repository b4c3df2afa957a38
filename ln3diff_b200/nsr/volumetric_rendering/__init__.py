from .ray_sampler import PatchRaySampler, RaySampler  # noqa
from .renderer import ImportanceRenderer  # noqa

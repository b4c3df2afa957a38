"""Mirror of the helpers of reference sgm/util.py used by the sampling stack."""
import importlib

import torch


def append_dims(x, target_dims):
    """sgm/util.py: right-pad dims."""
    dims_to_append = target_dims - x.ndim
    if dims_to_append < 0:
        raise ValueError(f"input has {x.ndim} dims but target_dims is {target_dims}, which is less")
    return x[(...,) + (None,) * dims_to_append]


def append_zero(x):
    return torch.cat([x, x.new_zeros([1])])


def default(val, d):
    if val is not None:
        return val
    return d() if callable(d) and not isinstance(d, (dict, list)) else d


_ALIASES = {"sgm.": "ln3diff_b200.sgm."}


def get_obj_from_str(string):
    for ref_prefix, ours in _ALIASES.items():  # reference dotted paths resolve to this mirror
        if string.startswith(ref_prefix):
            string = ours + string[len(ref_prefix):]
            break
    module, cls = string.rsplit(".", 1)
    return getattr(importlib.import_module(module), cls)


def instantiate_from_config(config):
    """sgm/util.py:168-185 (`target:` / `params:` dicts as the YAML configs hold them)."""
    if "target" not in config:
        raise KeyError("Expected key `target` to instantiate.")
    return get_obj_from_str(config["target"])(**config.get("params", dict()))

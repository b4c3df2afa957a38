"""Small host utilities shared by bench.py, smoke() and the tests (no oracle imports here)."""
from __future__ import annotations

import torch


def derandomize_zero_init(module: torch.nn.Module, std: float = 0.02, seed: int = 1234) -> None:
    """adaLN-Zero + the zero-initialised final layer make a freshly constructed DiT output exactly
    0 (reference dit/dit_models_xformers.py:807-819).  There are no checkpoints offline, so
    benchmarks / tests replace every all-zero floating parameter by N(0, std^2), in sorted key
    order from a CPU generator (identical to oracle.dit.derandomize_zero_init)."""
    g = torch.Generator().manual_seed(seed)
    sd = module.state_dict()
    new = {}
    for k in sorted(sd.keys()):
        v = sd[k]
        if v.is_floating_point() and v.numel() > 0 and float(v.abs().max()) == 0.0:
            new[k] = (torch.randn(v.shape, generator=g, dtype=torch.float32) * std).to(v.device, v.dtype)
    if new:
        sd.update(new)
        module.load_state_dict(sd)


def build_t23d(arch: str = "DiT-L/2", seed: int = 0, device: str | None = None):
    """DiT as the reference's create_model_and_diffusion builds it for T23D
    (guided_diffusion/script_util.py:407-415), random-init + derandomised zero tensors."""
    from .dit.dit_models_xformers import TextCondDiTBlock
    from .dit.dit_trilatent import DiT_models
    torch.manual_seed(seed)
    m = DiT_models[arch](input_size=32, num_classes=0, learn_sigma=False, in_channels=4,
                         context_dim=768, roll_out=True, vit_blk=TextCondDiTBlock)
    derandomize_zero_init(m)
    m.eval()
    return m.to(device) if device else m


def orbit_cameras(V: int, radius: float = 1.7, focal: float = 1.3889) -> torch.Tensor:
    """(V, 25) synthetic camera rows in the layout of the reference's assets/objv_eval_pose.pt:
    16 row-major cam2world (OpenCV convention, looking at the origin) + 9 normalised intrinsics
    (fx = fy = 1.3889, cx = cy = 0.5)."""
    import math
    cams = []
    for v in range(V):
        az = 2 * math.pi * v / max(V, 1) + 0.3
        el = 0.35 + 0.2 * ((v % 3) - 1)
        eye = torch.tensor([radius * math.cos(el) * math.cos(az), radius * math.cos(el) * math.sin(az),
                            radius * math.sin(el)])
        fwd = -eye / eye.norm()
        right = torch.linalg.cross(fwd, torch.tensor([0.0, 0.0, 1.0]))
        right = right / right.norm()
        down = torch.linalg.cross(fwd, right)
        c2w = torch.eye(4)
        c2w[:3, 0], c2w[:3, 1], c2w[:3, 2], c2w[:3, 3] = right, down, fwd, eye
        K = torch.tensor([focal, 0, 0.5, 0, focal, 0.5, 0, 0, 1.0])
        cams.append(torch.cat([c2w.reshape(-1), K]))
    return torch.stack(cams).float()

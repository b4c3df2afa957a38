"""Frame sink: what `TrainLoopDiffusionWithRec.render_video_given_triplane` does to every rendered view
before it reaches the video writer (nsr/train_util_diffusion.py:292-376) -- per-view depth normalisation,
`plt.cm.viridis`, [image | depth] side by side, HWC, uint8 -- as ONE device pass over all views
(`ln3_pack_frames`) and ONE batched device->host copy, instead of a `.cpu()` + numpy + matplotlib round
trip per view.  The uint8 frames are also what the prompt-sharded pipeline all-gathers over NCCL.

The colormap is a 256-entry table.  matplotlib is not a dependency of this package (and is absent from the
build image), so the built-in table is the published degree-6 polynomial fit of viridis (max abs error
~0.01 per channel against matplotlib's listed table); callers who need the reference's exact bytes pass
`colormap=plt.cm.viridis(np.arange(256))[:, :3]`.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops

_VIRIDIS_POLY = np.array([   # c0..c6 per channel (r, g, b): rgb(t) = sum_k c_k t^k, t in [0, 1]
    [0.2777273272234177, 0.005407344544966578, 0.3340998053353061],
    [0.1050930431085774, 1.404613529898575, 1.384590162594685],
    [-0.3308618287255563, 0.214847559468213, 0.09509516302823659],
    [-4.634230498983486, -5.799100973351585, -19.33244095627987],
    [6.228269936347081, 14.17993336680509, 56.69055260068105],
    [4.776384997670288, -13.74514537774601, -65.35303263337234],
    [-5.435455855934631, 4.645852612178535, 26.3124352495832]], dtype=np.float64)


def viridis_table() -> np.ndarray:
    """(256, 3) float64 in [0, 1]: the polynomial viridis evaluated at the 256 table positions i / 255."""
    t = np.arange(256, dtype=np.float64) / 255.0
    rgb = np.zeros((256, 3))
    for c in _VIRIDIS_POLY[::-1]:
        rgb = rgb * t[:, None] + c[None]
    return np.clip(rgb, 0.0, 1.0)


def colormap_bytes(colormap: np.ndarray | None = None) -> np.ndarray:
    """(256, 3) uint8: the bytes a colour-mapped depth pixel becomes in the reference's frame,
    uint8(clip((rgb * 2 - 1) * 127.5 + 127.5, 0, 255)) in float64 (nsr/train_util_diffusion.py:305,366-368)."""
    lut = viridis_table() if colormap is None else np.asarray(colormap, dtype=np.float64)
    if lut.shape != (256, 3):
        raise ValueError("colormap must be a (256, 3) table of rgb in [0, 1]")
    return ((lut * 2 - 1) * 127.5 + 127.5).clip(0, 255).astype(np.uint8)


class FrameSink:
    """Device-side video-frame packer with a reusable pinned host buffer.

    sink = FrameSink(device); frames = sink.pack(image_raw, image_depth)      # uint8 (N, H, 2W, 3) on the GPU
    host = sink.to_host(frames)                                               # one async D2H into pinned memory
    """

    def __init__(self, device, colormap: np.ndarray | None = None):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("FrameSink packs frames on CUDA only (no CPU fallback)")
        self.lut = torch.from_numpy(colormap_bytes(colormap)).to(self.device)
        self._host = None

    @torch.no_grad()
    def pack(self, image_raw: torch.Tensor, image_depth: torch.Tensor | None = None, out: torch.Tensor | None = None):
        """image_raw (..., 3, H, W) fp32 in [-1, 1], image_depth (..., 1, H, W) or None -> uint8 (..., H, Wout, 3)."""
        lead = image_raw.shape[:-3]
        H, W = image_raw.shape[-2:]
        img = image_raw.reshape(-1, 3, H, W).float().contiguous()
        dep = image_depth.reshape(-1, 1, H, W).float().contiguous() if image_depth is not None else None
        fr = ops.pack_frames(img, dep, self.lut if dep is not None else None,
                             out=out.view(img.shape[0], H, -1, 3) if out is not None else None)
        return fr.view(*lead, H, fr.shape[2], 3)

    def to_host(self, frames: torch.Tensor) -> torch.Tensor:
        """One non-blocking D2H of all frames into a cached pinned buffer (synchronise the stream before
        reading; the buffer is reused by the next call)."""
        if self._host is None or self._host.shape != frames.shape:
            self._host = torch.empty(frames.shape, dtype=torch.uint8, pin_memory=True)
        self._host.copy_(frames, non_blocking=True)
        return self._host

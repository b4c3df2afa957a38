"""TEST INFRASTRUCTURE ONLY (oracle) -- numpy restatement of the reference's frame sink.

Only tests/ and __graft_entry__.smoke() may import this module; the product path never does.

Restates nsr/train_util_diffusion.py:292-376 for the Objaverse case (no 'image_sr'):
    pred_depth = (d - d.min()) / (d.max() - d.min())                        (:299-301, fp32 torch)
    pred_depth = plt.cm.viridis(pred_depth[..., 0])[..., :3] * 2 - 1        (:304-305, float64)
    pred_vis   = cat([image_raw, pred_depth], dim=-1)                       (:340-345, promotes to float64)
    vis        = (pred_vis.permute(0,2,3,1) * 127.5 + 127.5).clip(0,255).astype(uint8)   (:366-368)
and matplotlib's `Colormap.__call__` on a float array (third-party, absent from the image -- "parity
unpinned" for it; published semantics): xa = x * N in x's dtype; xa == N -> N - 1; truncate to int;
index the N-entry table; NaN -> the "bad" colour (0, 0, 0, 0).  The 256-entry table itself is an input.
"""
from __future__ import annotations

import numpy as np


def colormap_call(x: np.ndarray, table: np.ndarray) -> np.ndarray:
    """matplotlib.colors.Colormap.__call__ for float input in [0, 1]; table (N, 3) float64 -> (..., 3) float64."""
    N = table.shape[0]
    xa = np.array(x, copy=True)
    bad = np.isnan(xa)
    with np.errstate(invalid="ignore"):
        xa *= xa.dtype.type(N)
        xa[xa == N] = N - 1
        xa = np.clip(xa, -1, N)
        idx = np.where(bad, 0, xa).astype(int)
    idx = np.clip(idx, 0, N - 1)
    out = table[idx]
    out[bad] = 0.0
    return out


def video_frames(image_raw: np.ndarray, image_depth: np.ndarray | None, table: np.ndarray) -> np.ndarray:
    """image_raw (N,3,H,W) fp32, image_depth (N,1,H,W) fp32 or None, table (256,3) float64 rgb in [0,1]
    -> uint8 (N, H, W or 2W, 3), one view at a time exactly as the reference's loop does."""
    frames = []
    for n in range(image_raw.shape[0]):
        img = image_raw[n:n + 1].astype(np.float32)
        if image_depth is None:
            vis = img
        else:
            d = image_depth[n:n + 1].astype(np.float32)
            with np.errstate(invalid="ignore", divide="ignore"):
                d = (d - d.min()) / (d.max() - d.min())                     # fp32
            dc = colormap_call(d[0].transpose(1, 2, 0)[..., 0], table) * 2 - 1   # (H, W, 3) float64
            vis = np.concatenate([img.astype(np.float64), dc.transpose(2, 0, 1)[None]], axis=-1)
        vis = vis.transpose(0, 2, 3, 1) * 127.5 + 127.5
        with np.errstate(invalid="ignore"):
            frames.append(np.nan_to_num(vis.clip(0, 255), nan=0.0).astype(np.uint8)[0])
    return np.stack(frames)

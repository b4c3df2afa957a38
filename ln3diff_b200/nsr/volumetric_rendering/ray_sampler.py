"""Mirror of reference nsr/volumetric_rendering/ray_sampler.py:60-331 (full-resolution rays).

`RaySampler.forward(cam2world (N,4,4), intrinsics (N,3,3), resolution)` and
`PatchRaySampler.forward(cam2world, intrinsics, patch_resolution, resolution, fg_bbox=None)` return
(ray_origins, ray_directions, bboxes) with ray m = y*W + x.  The device work is ln3_generate_rays."""
import random

import torch

from ... import ops


def _cams25(cam2world_matrix, intrinsics):
    N = cam2world_matrix.shape[0]
    return torch.cat([cam2world_matrix.reshape(N, 16), intrinsics.reshape(N, 9)], 1).float().contiguous()


class RaySampler(torch.nn.Module):
    def forward(self, cam2world_matrix, intrinsics, resolution, fg_mask=None):
        if not cam2world_matrix.is_cuda:
            raise RuntimeError("ln3diff_b200 RaySampler runs on CUDA only (no CPU fallback)")
        o, d = ops.generate_rays(_cams25(cam2world_matrix, intrinsics), int(resolution))
        return o, d, None


class PatchRaySampler(RaySampler):
    def forward(self, cam2world_matrix, intrinsics, patch_resolution, resolution, fg_bbox=None):
        """Rendering calls this with patch_resolution == resolution (nsr/triplane.py:555-558): the
        'patch' is then the whole image (ray_sampler.py:93-131 clips the sampled end to
        `resolution`).  Training-time random patches are outside the hot path."""
        if patch_resolution != resolution or fg_bbox is not None:
            raise NotImplementedError("random training patches are outside the generation hot path")
        N = cam2world_matrix.shape[0]
        for _ in range(2 * N):  # the reference consumes two python `random.randint` per view here
            random.randint(patch_resolution, resolution + patch_resolution)
        o, d, _ = super().forward(cam2world_matrix, intrinsics, resolution)
        return o, d, [(0, 0, patch_resolution, patch_resolution)] * N

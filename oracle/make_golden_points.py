"""TEST INFRASTRUCTURE.  Golden vectors for the point-query path (mesh extraction), produced by the
reference's own code run in the build container:

  points.npz   ImportanceRenderer._run_model (nsr/volumetric_rendering/renderer.py:310-322) on the seeded
               tri-plane / OSG weights of oracle.fixtures.render_inputs: 4096 random points in
               [-0.6, 0.6]^3 (a fifth of them outside the planes' support) and the 9^3 lattice that
               triplane_decode_grid (vit/vit_triplane.py:2052-2120) builds for the Objaverse bbox.

Run:  python oracle/make_golden_points.py      (needs /root/reference; writes tests/golden/points.npz)
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import _stubs  # noqa: E402  (stubs for the reference's absent third-party imports)

_stubs.install()
sys.path.insert(0, "/root/reference")
from oracle import fixtures as fx  # noqa: E402


def main():
    from nsr.volumetric_rendering.renderer import ImportanceRenderer
    planes, osg, _, _ = fx.render_inputs(8)
    w1, b1, w2, b2 = osg

    class Dec(torch.nn.Module):  # OSGDecoder arithmetic (nsr/triplane.py:356-375) on raw tensors
        def forward(self, feats, dirs):
            v = feats.mean(1)
            N, M, C = v.shape
            v = v.view(N * M, C)
            h = torch.nn.functional.softplus(torch.addmm(b1.unsqueeze(0), v, (w1 * (1 / np.sqrt(32))).t()))
            yy = torch.addmm(b2.unsqueeze(0), h, (w2 * (1 / np.sqrt(64))).t()).view(N, M, -1)
            return {"rgb": torch.sigmoid(yy[..., 1:]) * (1 + 2 * 0.001) - 0.001, "sigma": yy[..., 0:1]}

    opts = {"box_warp": 0.9}
    g = torch.Generator().manual_seed(77)
    pts = (torch.rand(1, 4096, 3, generator=g) - 0.5) * 1.2
    ren = ImportanceRenderer()
    r = ren._run_model(planes[None], Dec(), pts, torch.zeros_like(pts), opts)
    G = 9
    axes = [torch.linspace(-0.45, 0.45, G) for _ in range(3)]
    grid = torch.stack(torch.meshgrid(*axes, indexing="ij"), dim=-1).reshape(1, -1, 3)
    rg = ren._run_model(planes[None], Dec(), grid, torch.zeros_like(grid), opts)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "points.npz"),
                        points=pts[0].numpy(), rgb=r["rgb"][0].numpy(), sigma=r["sigma"][0].numpy(),
                        grid_size=np.array(G), grid_rgb=rg["rgb"][0].numpy(), grid_sigma=rg["sigma"][0].numpy())
    print("points", float(r["sigma"].mean()), float(rg["sigma"].mean()))


if __name__ == "__main__":
    main()

"""Fused ray-march timings: views/s at 128^2 and 256^2 (TF32 MLP), LN3_RENDER_TILES selects the schedule."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ln3diff_b200 import ops
from ln3diff_b200.utils import orbit_cameras

dev = "cuda"
g = torch.Generator().manual_seed(4)
n_obj = 4
planes = (5 * torch.randn(n_obj, 3, 32, 128, 128, generator=g)).to(dev)
osg = [torch.randn(64, 32, generator=g), torch.randn(64, generator=g) * 0.1, torch.randn(4, 64, generator=g),
       torch.randn(4, generator=g) * 0.1]
osg[3][0] += 2.0
osg = tuple(t.to(dev) for t in osg)
pcl = ops.planes_to_channels_last(planes)
out = []
for res, nv in ((128, 16), (256, 8)):
    cams = orbit_cameras(nv).repeat(n_obj, 1).to(dev)
    M = res * res
    nc, nf = torch.rand(n_obj * nv, M, 64, device=dev), torch.rand(n_obj * nv, M, 64, device=dev)
    o, d = ops.generate_rays(cams, res)
    for tf32 in (True, False):
        for _ in range(2):
            r = ops.render_views(pcl, o, d, nc, nf, osg, views_per_obj=nv, mlp_tf32=tf32)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            r = ops.render_views(pcl, o, d, nc, nf, osg, views_per_obj=nv, mlp_tf32=tf32)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        out.append(f"{res}^2 {'tf32' if tf32 else 'fp32'}: {n_obj * nv / (ms / 1e3):.0f} views/s (checksum {float(r['rgb'].sum()):.3f})")
print(f"tiles={os.environ.get('LN3_RENDER_TILES', '1')} " + "; ".join(out), flush=True)

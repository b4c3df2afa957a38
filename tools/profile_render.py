"""Render V views at res (for ncu captures)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ln3diff_b200 import ops
from ln3diff_b200.utils import orbit_cameras

V = int(sys.argv[1]) if len(sys.argv) > 1 else 4
res = int(sys.argv[2]) if len(sys.argv) > 2 else 128
n = int(sys.argv[3]) if len(sys.argv) > 3 else 2
g = torch.Generator().manual_seed(4)
dev = "cuda"
planes = (5 * torch.randn(1, 3, 32, 128, 128, generator=g)).to(dev)
osg = [torch.randn(64, 32, generator=g), torch.randn(64, generator=g) * 0.1,
       torch.randn(4, 64, generator=g), torch.randn(4, generator=g) * 0.1]
osg[3][0] += 2.0
osg = tuple(t.to(dev) for t in osg)
M = res * res
nc, nf = torch.rand(V, M, 64, device=dev), torch.rand(V, M, 64, device=dev)
pcl = ops.planes_to_channels_last(planes)
o, d = ops.generate_rays(orbit_cameras(V).to(dev), res)
for _ in range(n):
    ops.render_views(pcl, o, d, nc, nf, osg, views_per_obj=V, mlp_tf32=len(sys.argv) > 4 and sys.argv[4] == "tf32")
torch.cuda.synchronize()
print("done")

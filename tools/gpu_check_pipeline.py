"""Time pipeline.sample_t23d (configs[1]: DiT-L/2, 250-step Euler-EDM + CFG 6.5, 8 prompts) -- A/B helper."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ln3diff_b200 import pipeline
from ln3diff_b200.utils import build_t23d

dev = "cuda"
B = 8
m = build_t23d("DiT-L/2", device=dev)
g = torch.Generator().manual_seed(0)
c = {"crossattn": torch.randn(B, 77, 768, generator=g).to(dev)}
uc = {"crossattn": torch.zeros(B, 77, 768, device=dev)}
z = torch.randn(B, 12, 32, 32, generator=g).to(dev)
tables = pipeline.edm_cfg_tables(250, 6.5, B, dev)
out = pipeline.sample_t23d(m, z, c, uc, 250, 6.5, tables)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
n = 2
for _ in range(n):
    out = pipeline.sample_t23d(m, z, c, uc, 250, 6.5, tables)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(f"closed_form={os.environ.get('LN3_UNCOND_CLOSED_FORM', '1')} rows={m._ctx_cache.value['rows']}: {ms:.1f} ms per batch, "
      f"{ms / 250:.3f} ms/step, {B / (ms / 1e3):.3f} latents/s, finite={bool(torch.isfinite(out).all())}, "
      f"checksum={float(out.double().abs().mean()):.6f}")

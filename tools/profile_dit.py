"""Run N DiT-L/2 forwards at B'=16 (for ncu launch lists / captures)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ln3diff_b200.dit.dit_trilatent import DiT_models
from ln3diff_b200.dit.dit_models_xformers import TextCondDiTBlock
from oracle import dit as odit

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
arch = sys.argv[2] if len(sys.argv) > 2 else "DiT-L/2"
B = int(sys.argv[3]) if len(sys.argv) > 3 else 16
torch.manual_seed(0)
m = DiT_models[arch](input_size=32, num_classes=0, learn_sigma=False, in_channels=4,
                     context_dim=768, roll_out=True, vit_blk=TextCondDiTBlock)
m.load_state_dict(odit.derandomize_zero_init(m.state_dict()))
m = m.cuda()
x = torch.randn(B, 12, 32, 32, device="cuda")
t = torch.randint(0, 1000, (B,), device="cuda")
ctx = torch.randn(B, 77, 768, device="cuda")
if os.environ.get("LN3_PROFILE_CFG", "0") != "0":   # the sampler's batch: zero-embedding (uncond) half first
    ctx[:B // 2] = 0
for _ in range(n):
    m(x, t, ctx)
torch.cuda.synchronize()
print("done")

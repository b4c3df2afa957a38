"""Run N VAE decodes of 8 latents (for ncu launch lists / captures): python tools/profile_decode.py [n]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ln3diff_b200.utils import build_ae_decoder

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dec = build_ae_decoder("DiT2-L/2", device="cuda")
lat = torch.randn(8, 12, 32, 32, device="cuda")
for _ in range(n):
    out = dec.decode_to_channels_last(lat, in_mul=0.96806)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(n):
    out = dec.decode_to_channels_last(lat, in_mul=0.96806)
e1.record()
torch.cuda.synchronize()
print("done", tuple(out.shape), f"{e0.elapsed_time(e1) / n:.3f} ms per 8-latent decode, checksum {float(out.double().abs().mean()):.6f}")

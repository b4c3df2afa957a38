"""Full-size conditioner towers: parity against the fp32 transformers models (CPU) and throughput on the GPU.
    python tools/gpu_check_conditioners.py -> gpurun_out/conditioners.json"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ln3diff_b200.sgm.modules.encoders.modules import (FrozenCLIPEmbedder, FrozenDinov2ImageEmbedder,
                                                       FrozenOpenCLIPImageEmbedder)
from oracle import conditioners as oc

dev = torch.device("cuda", 0)


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).norm() / b.norm()).item()


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


res = {}
g = torch.Generator().manual_seed(0)
# CLIP-L text, 8 prompts
hf, sd = oc.clip_text(depth=12)
ids = torch.randint(3, 49000, (8, 77), generator=g)
ids[:, 30:] = 49407
emb = FrozenCLIPEmbedder(device=dev, always_return_pooled=True, state_dict=sd)
z, pooled = emb(ids)
t0 = time.time()
with torch.no_grad():
    ref = hf(input_ids=ids)
cpu_s = time.time() - t0
ms = timeit(lambda: emb(ids))
res["clip_text_L"] = {"rel_l2_last": rel(z, ref.last_hidden_state), "rel_l2_pooled": rel(pooled, ref.pooler_output),
                      "ms_per_8_prompts": ms, "prompts_per_s": 8 / ms * 1e3, "cpu_fp32_transformers_s_per_8": cpu_s}
print(res["clip_text_L"], flush=True)
# OpenCLIP ViT-L/14 image tower, 8 images
hf, sd = oc.clip_vision(depth=24, width=1024, mlp=4096, embed=768)
emb = FrozenOpenCLIPImageEmbedder(device=dev, output_tokens=True, state_dict=sd)
img = torch.rand(8, 3, 224, 224, generator=g) * 2 - 1
tokens, zz = emb(img.to(dev))
t0 = time.time()
rz, rt = oc.clip_vision_forward(hf, emb.preprocess(img.to(dev)).cpu())
cpu_s = time.time() - t0
ms = timeit(lambda: emb(img.to(dev)))
res["openclip_vit_L14"] = {"rel_l2_tokens": rel(tokens, rt), "rel_l2_pooled": rel(zz, rz), "ms_per_8_images": ms,
                           "images_per_s": 8 / ms * 1e3, "cpu_fp32_transformers_s_per_8": cpu_s}
print(res["openclip_vit_L14"], flush=True)
# DINOv2 ViT-L/14 reg
hf, sd = oc.dinov2_reg(depth=24, width=1024)
emb = FrozenDinov2ImageEmbedder(device=dev, state_dict=sd)
tok = emb(img.to(dev))
t0 = time.time()
with torch.no_grad():
    ref = hf(pixel_values=emb.preprocess(img.to(dev)).cpu()).last_hidden_state
cpu_s = time.time() - t0
ms = timeit(lambda: emb(img.to(dev)))
res["dinov2_vit_L14_reg"] = {"rel_l2_patch_tokens": rel(tok, ref[:, 5:]), "ms_per_8_images": ms, "images_per_s": 8 / ms * 1e3,
                             "cpu_fp32_transformers_s_per_8": cpu_s}
print(res["dinov2_vit_L14_reg"], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/conditioners.json", "w"), indent=1)

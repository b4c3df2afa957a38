"""250-step Euler-EDM + CFG trajectory of DiT-B/2 (1 prompt): the fused CUDA loop (`pipeline.sample_t23d`
arithmetic, stepped by hand so that every intermediate state is visible) against the fp32 CPU oracle loop on the
same seed.  Writes gpurun_out/trajectory_r2.json: rel-L2 of the sampler state after steps 1, 2, 5, 10, 25, 50,
100, 150, 200, 250 (SURVEY.md 7.2 asks for the full-loop trajectory, not just the end point)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ln3diff_b200 import ops, pipeline
from ln3diff_b200.utils import build_t23d
from oracle import dit as odit
from oracle import samplers as osmp

dev = torch.device("cuda", 0)
STEPS, SCALE = 250, 6.5
marks = [1, 2, 5, 10, 25, 50, 100, 150, 200, 250]
m = build_t23d("DiT-B/2")
sd = {k: v.clone() for k, v in m.state_dict().items()}
g = torch.Generator().manual_seed(41)
x0 = torch.randn(1, 12, 32, 32, generator=g)
c = torch.randn(1, 77, 768, generator=g)
uc = torch.zeros(1, 77, 768)

# ---- oracle loop (records the state at the marks)
from ln3diff_b200.utils import host_cores
torch.set_num_threads(host_cores())
table = osmp.legacy_ddpm_sigmas(1000, append_zero=False, flip=True)
sigmas = osmp.legacy_ddpm_sigmas(STEPS)
ref_states = {}
t0 = time.perf_counter()
with torch.no_grad():
    x = x0 * torch.sqrt(1.0 + sigmas[0] ** 2.0)
    ctx = torch.cat((uc, c), 0)
    for i in range(STEPS):
        s, nxt = torch.ones(1) * sigmas[i], torch.ones(1) * sigmas[i + 1]
        xin, sin = torch.cat([x] * 2), torch.cat([s] * 2)
        sq = table[osmp.sigma_to_idx(sin, table)]
        sq4 = sq[:, None, None, None]
        den = odit.dit_t23d_forward(sd, "DiT-B/2", xin / (sq4 ** 2 + 1.0) ** 0.5, osmp.sigma_to_idx(sq, table), ctx) * (-sq4) + xin
        x_u, x_c = den.chunk(2)
        d = (x - (x_u + SCALE * (x_c - x_u))) / s[:, None, None, None]
        x = x + (nxt - s)[:, None, None, None] * d
        if i + 1 in marks:
            ref_states[i + 1] = x.clone()
cpu_s = time.perf_counter() - t0

# ---- CUDA loop: the body of pipeline.sample_t23d
m = m.to(dev)
tables = pipeline.edm_cfg_tables(STEPS, SCALE, 1, dev)
ctx_d = torch.cat((uc, c), 0).to(dev)
xa = (x0.to(dev) * tables["init_scale"]).contiguous()
xb = torch.empty_like(xa)
gr = m.capture_graph(2, ctx_d, shared_mod=True)
mod = m.modulation_table(tables["t_idx"][:STEPS, 0])
out = {}
for i in range(STEPS):
    gr.x[:1].copy_(xa); gr.x[1:].copy_(xa)
    gr.mod.copy_(mod[i:i + 1]); gr.in_scale.copy_(tables["c_in"][i])
    gr.replay()
    ops.sampler_affine_update(xa, tables["coef"][i], gr.out[:1], gr.out[1:], out=xb)
    xa, xb = xb, xa
    if i + 1 in marks:
        r = ref_states[i + 1].double()
        out[i + 1] = float((xa.double().cpu() - r).norm() / r.norm())
check = pipeline.sample_t23d(m, x0.to(dev), {"crossattn": c.to(dev)}, {"crossattn": uc.to(dev)}, STEPS, SCALE)
res = {"arch": "DiT-B/2", "steps": STEPS, "cfg_scale": SCALE, "rel_l2_after_step": out,
       "pipeline_equals_stepped_loop": bool(torch.equal(check, xa)), "oracle_cpu_seconds": cpu_s}
print(json.dumps(res))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/trajectory_r2.json", "w"), indent=1)

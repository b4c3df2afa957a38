"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by running the REFERENCE's own code.

Runs only in the build container (needs /root/reference, read-only; third-party gaps are filled by
oracle/_stubs.py).  Nothing is copied from the reference: the fixtures hold seeded inputs and the
outputs the reference modules produce for them.  Re-run:  python -m oracle.make_golden

Fixtures
  dit_t23d.npz   reference DiT_TriLatent (DiT-B/2, TextCondDiTBlock) forward on seeded inputs with a
                 synthetic, key-seeded state_dict (oracle.dit.synth_state_dict)
  samplers.npz   EulerEDMSampler+DiscreteDenoiser+VanillaCFG, SpacedDiffusion.p_sample_loop and
                 transport Sampler.sample_ode('euler') around a closed-form toy network
  render.npz     RaySampler + ImportanceRenderer.forward (Objaverse preset) on a seeded 3x32x16x16
                 tri-plane with explicit noise, plus per-stage debug tensors from Triplane-level I/O
  cameras.npz    rows of the reference's assets/objv_eval_pose.pt (the only golden inputs it ships)
"""
from __future__ import annotations

import contextlib
import io
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import _stubs  # noqa: E402
from oracle import dit as odit  # noqa: E402
from oracle import fixtures as fx  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def main():
    _stubs.install()
    os.makedirs(OUT, exist_ok=True)
    import dit.dit_models_xformers as dmx
    _stubs.patch_dit_namespace()
    import dit.dit_trilatent as dt

    # ---------------------------------------------------------------- DiT T23D forward
    with contextlib.redirect_stdout(io.StringIO()):
        ref = dt.DiT_models["DiT-B/2"](input_size=32, num_classes=0, learn_sigma=False, in_channels=4,
                                       context_dim=768, roll_out=True, vit_blk=dmx.TextCondDiTBlock)
    ref.eval()
    shapes = {k: tuple(v.shape) for k, v in ref.state_dict().items()}
    sd = odit.synth_state_dict(shapes, seed=7, keep={"pos_embed": ref.state_dict()["pos_embed"]})
    ref.load_state_dict(sd)
    x, t, ctx = fx.dit_inputs()
    with torch.no_grad():
        y = ref(x, t, {"crossattn": ctx})
    np.savez_compressed(os.path.join(OUT, "dit_t23d.npz"), out=y.numpy(), t=t.numpy(),
                        pos_embed_checksum=np.float64(ref.state_dict()["pos_embed"].double().sum().item()))
    print("dit_t23d", y.shape, float(y.abs().max()))

    # ---------------------------------------------------------------- DiT I23D (PixArt) forward
    import dit.dit_i23d as di
    with contextlib.redirect_stdout(io.StringIO()):
        ref2 = di.DiT_models[fx.I23D_ARCH](input_size=32, num_classes=0, learn_sigma=False, in_channels=4,
                                           context_dim=1024, pooling_ctx_dim=768, roll_out=True)
    ref2.eval()
    assert type(ref2).__name__ == "DiT_I23D_PixelArt" and type(ref2.blocks[0]).__name__ == "ImageCondDiTBlockPixelArtRMSNorm"
    shapes2 = {k: tuple(v.shape) for k, v in ref2.state_dict().items()}
    ref2.load_state_dict(fx.i23d_state_dict(shapes2, ref2.state_dict()["pos_embed"]))
    xi, ti, ci = fx.i23d_inputs()
    with torch.no_grad():
        yi = ref2(xi, ti, ci)
        yc = ref2.forward_with_cfg(xi, ti, ci, 4.0)
    np.savez_compressed(os.path.join(OUT, "dit_i23d.npz"), out=yi.numpy(), out_cfg=yc.numpy())
    print("dit_i23d", yi.shape, float(yi.abs().max()))

    # ---------------------------------------------------------------- samplers (toy network)
    from sgm.modules.diffusionmodules.denoiser import DiscreteDenoiser
    from sgm.modules.diffusionmodules.sampling import EulerEDMSampler
    disc = {"target": "sgm.modules.diffusionmodules.discretizer.LegacyDDPMDiscretization"}
    sampler = EulerEDMSampler(discretization_config=disc, num_steps=10, device="cpu", guider_config={
        "target": "sgm.modules.diffusionmodules.guiders.VanillaCFG", "params": {"scale": 6.5}})
    den = DiscreteDenoiser(scaling_config={"target": "sgm.modules.diffusionmodules.denoiser_scaling.EpsScaling"},
                           num_idx=1000, discretization_config=disc)
    toy = fx.toy_network()
    x0, c, uc, noise, step_noise, z = fx.sampler_inputs()
    sgm_out = sampler(lambda inp, sig, cc: den(toy, inp, sig, cc), x0.clone(), c, uc)
    sig10 = sampler.discretization(10, device="cpu")
    sig250 = sampler.discretization(250, device="cpu")

    import guided_diffusion.gaussian_diffusion as gd
    from guided_diffusion.respace import SpacedDiffusion, space_timesteps
    diff = SpacedDiffusion(use_timesteps=space_timesteps(1000, "10"), betas=gd.get_named_beta_schedule("linear", 1000),
                           model_mean_type=gd.ModelMeanType.EPSILON, model_var_type=gd.ModelVarType.FIXED_LARGE,
                           loss_type=gd.LossType.MSE, rescale_timesteps=False)

    class M:
        def apply_model_inference(self, xx, tt, cc, **kw):
            return toy(xx, tt * 1000, cc)

    it = iter(step_noise)
    orig = torch.randn_like
    torch.randn_like = lambda v: next(it)
    try:
        ddpm_out = diff.p_sample_loop(M(), tuple(noise.shape), cond=c["crossattn"], noise=noise,
                                      clip_denoised=False, device="cpu")
    finally:
        torch.randn_like = orig

    from transport import Sampler, create_transport
    fn = Sampler(create_transport(snr_type="lognorm")).sample_ode(sampling_method="euler", num_steps=10)
    ctx2 = {"crossattn": torch.cat([c["crossattn"], uc["crossattn"]])}

    def fwd_cfg(xx, tt, context, cfg_scale):
        e = toy(xx, tt * 1000, context)
        ce, ue = torch.split(e, len(e) // 2, dim=0)
        h = ue + cfg_scale * (ce - ue)
        return torch.cat([h, h], 0)

    flow_out = fn(torch.cat([z, z], 0), fwd_cfg, context=ctx2, cfg_scale=4.0)[-1].chunk(2)[0]
    np.savez_compressed(os.path.join(OUT, "samplers.npz"), sgm=sgm_out.numpy(), ddpm=ddpm_out.numpy(),
                        flow=flow_out.numpy(), sigmas10=sig10.numpy(), sigmas250=sig250.numpy(),
                        denoiser_sigmas=den.sigmas.numpy(),
                        ddpm_betas10=diff.betas, ddpm_timestep_map=np.array(diff.timestep_map))
    print("samplers", float(sgm_out.abs().max()), float(ddpm_out.abs().max()), float(flow_out.abs().max()))

    # ---------------------------------------------------------------- renderer
    from nsr.volumetric_rendering.ray_sampler import RaySampler
    from nsr.volumetric_rendering.renderer import ImportanceRenderer
    cams = torch.load("/root/reference/assets/objv_eval_pose.pt").float()
    np.savez_compressed(os.path.join(OUT, "cameras.npz"), objv_eval_pose=cams.numpy())
    res = 24
    planes, osg, nc, nf = fx.render_inputs(res)
    w1, b1, w2, b2 = osg

    class Dec(torch.nn.Module):  # OSGDecoder arithmetic (nsr/triplane.py:356-375) on raw tensors
        decoder_output_dim = 3

        def forward(self, feats, dirs):
            v = feats.mean(1)
            N, M, C = v.shape
            v = v.view(N * M, C)
            h = torch.addmm(b1.unsqueeze(0), v, (w1 * (1 / np.sqrt(32))).t())
            h = torch.nn.functional.softplus(h)
            yy = torch.addmm(b2.unsqueeze(0), h, (w2 * (1 / np.sqrt(64))).t()).view(N, M, -1)
            return {"rgb": torch.sigmoid(yy[..., 1:]) * (1 + 2 * 0.001) - 0.001, "sigma": yy[..., 0:1]}

    from oracle.render import OBJAVERSE_OPTS
    outs = {}
    for vi, cam_row in enumerate(fx.RENDER_CAM_ROWS):
        cam = cams[cam_row]
        o, d, _ = RaySampler()(cam[:16].reshape(1, 4, 4), cam[16:25].reshape(1, 3, 3), res)
        ncv, nfv = nc[vi], nf[vi]
        orl, orr = torch.rand_like, torch.rand
        torch.rand_like = lambda tt, *a, **k: ncv.reshape(tt.shape)
        torch.rand = lambda *s, **k: nfv.reshape(*s)
        try:
            r = ImportanceRenderer()(planes[None], Dec(), o, d, dict(OBJAVERSE_OPTS))
        finally:
            torch.rand_like, torch.rand = orl, orr
        outs[f"ray_o_{vi}"] = o[0].numpy()
        outs[f"ray_d_{vi}"] = d[0].numpy()
        outs[f"rgb_{vi}"] = r["feature_samples"][0].numpy()
        outs[f"depth_{vi}"] = r["depth_samples"][0].numpy()
        outs[f"weights_{vi}"] = r["weights_samples"][0].numpy()
        print("render view", vi, float(r["weights_samples"].mean()))
    np.savez_compressed(os.path.join(OUT, "render.npz"), **outs)
    # ---------------------------------------------------------------- VAE decoder
    import ldm.modules.diffusionmodules.model as lm
    lm.XFORMERS_IS_AVAILBLE = True
    lm.xformers = sys.modules["xformers"]
    from dit.dit_decoder import DiT2_models
    from einops import rearrange
    from ldm.modules.diffusionmodules.model import Decoder
    from vit.vit_triplane import PatchEmbedTriplane
    arch, D = fx.DECODER_ARCH, fx.DECODER_DIM
    with contextlib.redirect_stdout(io.StringIO()):
        up = PatchEmbedTriplane(32, 2, 12, D, bias=True)
        vd = DiT2_models[arch](input_size=16, num_classes=0, learn_sigma=False, in_channels=D,
                               mixed_prediction=False, context_dim=None, roll_out=True, plane_n=3,
                               return_all_layers=False)
        vd.pos_embed = torch.nn.Parameter(torch.zeros(1, 768, D))   # vit_triplane.py:210
        srd = Decoder(resolution=128, in_channels=3, ch=32, ch_mult=[1, 2, 2, 4], num_res_blocks=1, dropout=0.0,
                      attn_resolutions=[], out_ch=32, z_channels=D)
    assert type(srd.mid.attn_1).__name__ == "MemoryEfficientAttnBlock"
    parts = (("superresolution.ldm_upsample.", up), ("vit_decoder.", vd), ("superresolution.conv_sr.", srd))
    shapes = {pre + k: tuple(v.shape) for pre, m in parts for k, v in m.state_dict().items()}
    sdd = fx.decoder_state_dict(shapes)
    for pre, m in parts:
        m.load_state_dict({k[len(pre):]: v for k, v in sdd.items() if k.startswith(pre)})
    lat = fx.decoder_latent()
    with torch.no_grad():
        tok = vd(up(lat * fx.SCALING_DIVIDER))
        z = rearrange(tok.reshape(1, 3, 16, 16, D), "b n h w c->(b n) c h w")
        y = rearrange(srd(z), "(b n) c h w->b (n c) h w", n=3)
    np.savez_compressed(os.path.join(OUT, "decoder.npz"), crop=y[:, :, 40:56, 40:56].numpy(),
                        chan_mean=y.mean(dim=(0, 2, 3)).numpy(), chan_absmean=y.abs().mean(dim=(0, 2, 3)).numpy(),
                        tokens_crop=tok[:, ::37, ::29].numpy())
    print("decoder", tuple(y.shape), float(y.abs().max()))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()

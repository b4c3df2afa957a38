"""Mirror of reference nsr/triplane.py: OSGDecoder (:339-375) and Triplane (:423-750).

`Triplane(c_dim, img_resolution, img_channels, out_chans, triplane_size, rendering_kwargs,
decoder_in_chans, decoder_output_dim, ...)` is what create_3DAE_model builds
(nsr/script_util.py:1355-1370); `.forward(planes (N,96,H,W), c (N,25))` returns the reference's dict
(image_raw, image_depth, weights_samples, image_mask, feature_image, shape_synthesized)."""
import numpy as np
import torch
import torch.nn as nn

from .volumetric_rendering.ray_sampler import PatchRaySampler, RaySampler
from .volumetric_rendering.renderer import ImportanceRenderer


class FullyConnectedLayer(nn.Module):
    """nsr/networks_stylegan2.py:122-157 parameter container (runtime gains are NOT folded into
    the stored weights; the render kernel applies weight_gain = lr_mul / sqrt(in_features))."""

    def __init__(self, in_features, out_features, bias=True, activation="linear", lr_multiplier=1, bias_init=0):
        super().__init__()
        assert activation == "linear" and lr_multiplier == 1
        self.in_features, self.out_features = in_features, out_features
        self.weight = nn.Parameter(torch.randn([out_features, in_features]) / lr_multiplier)
        self.bias = nn.Parameter(torch.full([out_features], np.float32(bias_init))) if bias else None
        self.weight_gain = lr_multiplier / np.sqrt(in_features)
        self.bias_gain = lr_multiplier


class OSGDecoder(nn.Module):
    def __init__(self, n_features, options):
        super().__init__()
        self.hidden_dim = 64
        self.decoder_output_dim = options["decoder_output_dim"]
        self.net = nn.Sequential(
            FullyConnectedLayer(n_features, self.hidden_dim, lr_multiplier=options["decoder_lr_mul"]),
            nn.Softplus(),
            FullyConnectedLayer(self.hidden_dim, 1 + options["decoder_output_dim"],
                                lr_multiplier=options["decoder_lr_mul"]))
        self.activation = options.get("decoder_activation", "sigmoid")
        if n_features != 32 or self.decoder_output_dim != 3 or self.activation != "sigmoid":
            raise NotImplementedError("libln3b200 implements the Objaverse OSG decoder 32 -> 64 -> 1+3 (sigmoid)")

    def raw_parameters(self):
        return (self.net[0].weight.float().contiguous(), self.net[0].bias.float().contiguous(),
                self.net[2].weight.float().contiguous(), self.net[2].bias.float().contiguous())


class Triplane(nn.Module):
    def __init__(self, c_dim=25, img_resolution=128, img_channels=3, out_chans=96, triplane_size=224,
                 rendering_kwargs={}, decoder_in_chans=32, decoder_output_dim=32, sr_num_fp16_res=0,
                 sr_kwargs={}, create_triplane=False, bcg_synthesis_kwargs={}, lrm_decoder=False):
        super().__init__()
        if lrm_decoder or create_triplane or bool(sr_kwargs):
            raise NotImplementedError("LRM decoder / learned planes / super-resolution are outside the "
                                      "Objaverse generation path")
        self.c_dim, self.img_resolution, self.img_channels = c_dim, img_resolution, img_channels
        self.triplane_size, self.decoder_in_chans, self.out_chans = triplane_size, decoder_in_chans, out_chans
        self.renderer = ImportanceRenderer()
        self.ray_sampler = PatchRaySampler() if "PatchRaySampler" in rendering_kwargs else RaySampler()
        self.decoder = OSGDecoder(decoder_in_chans, {"decoder_lr_mul": rendering_kwargs.get("decoder_lr_mul", 1),
                                                     "decoder_output_dim": decoder_output_dim})
        self.neural_rendering_resolution = img_resolution
        self.rendering_kwargs = rendering_kwargs
        self.superresolution = None
        self.bcg_synthesis = None

    @torch.no_grad()
    def forward(self, planes=None, c=None, ws=None, ray_origins=None, ray_directions=None, z_bcg=None,
                neural_rendering_resolution=None, update_emas=False, cache_backbone=False,
                use_cached_backbone=False, return_meta=False, return_raw_only=False, sample_ray_only=False,
                fg_bbox=None, **synthesis_kwargs):
        cam2world_matrix = c[:, :16].reshape(-1, 4, 4)
        intrinsics = c[:, 16:25].reshape(-1, 3, 3)
        if neural_rendering_resolution is None:
            neural_rendering_resolution = self.neural_rendering_resolution
        else:
            self.neural_rendering_resolution = neural_rendering_resolution
        if ray_directions is None:
            H = W = self.neural_rendering_resolution
            if sample_ray_only:
                raise NotImplementedError("sample_ray_only (training patches) is outside the hot path")
            if isinstance(self.ray_sampler, PatchRaySampler):
                ray_origins, ray_directions, _ = self.ray_sampler(cam2world_matrix, intrinsics, H, H)
            else:
                ray_origins, ray_directions, _ = self.ray_sampler(cam2world_matrix, intrinsics, H)
        else:
            assert ray_origins is not None
            H = W = int(ray_directions.shape[1] ** 0.5)
        assert planes is not None
        return_sampling_details_flag = self.rendering_kwargs.get("return_sampling_details_flag", False)
        if return_sampling_details_flag:
            return_meta = True
        N, M, _ = ray_origins.shape
        if planes.shape[1] == 3 * 2 * self.decoder_in_chans:
            raise NotImplementedError("background tri-plane compositing is outside the Objaverse path")
        planes = planes.reshape(len(planes), 3, -1, planes.shape[-2], planes.shape[-1])
        rd = self.renderer(planes, self.decoder, ray_origins, ray_directions, self.rendering_kwargs,
                           return_meta=return_meta)
        feature_samples, depth_samples, weights_samples = (rd[k] for k in ("feature_samples", "depth_samples",
                                                                           "weights_samples"))
        shape_synthesized = rd["shape_synthesized"] if return_sampling_details_flag else None
        feature_image = feature_samples.permute(0, 2, 1).reshape(N, feature_samples.shape[-1], H, W).contiguous()
        depth_image = depth_samples.permute(0, 2, 1).reshape(N, 1, H, W)
        weights_samples = weights_samples.permute(0, 2, 1).reshape(N, 1, H, W)
        mask_image = weights_samples * (1 + 2 * 0.001) - 0.001
        rgb_image = feature_image[:, :3]
        if shape_synthesized is not None:
            shape_synthesized.update({"image_depth": depth_image})
        ret = {"feature_image": feature_image, "image_raw": rgb_image, "image_depth": depth_image,
               "weights_samples": weights_samples, "shape_synthesized": shape_synthesized,
               "image_mask": mask_image}
        if return_meta:
            ret.update({"feature_volume": rd.get("feature_volume"), "all_coords": rd.get("all_coords"),
                        "weights": rd.get("weights")})
        return ret

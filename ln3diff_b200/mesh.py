"""Mesh export of a decoded tri-plane: the `export_mesh` branch of
`TrainLoopDiffusionWithRec.render_video_given_triplane` (reference nsr/train_util_diffusion.py:208-249).

Reference sequence and what replaces it:
    grid_out = rec_model(latent=..., grid_size=mesh_size, behaviour='triplane_decode_grid')   -> ln3_query_points
    vtx, faces = mcubes.marching_cubes(sigma (G,G,G) on the CPU, mesh_thres)                  -> ln3_marching_cubes_*
    vtx = (vtx / (mesh_size - 1) * 2 - 1) * 0.45                                              -> folded into the emit pass
    vtx_colors = forward_points(planes, vtx)['rgb'].clip(0,1) * 255 -> uint8                  -> ln3_query_points
    vtx = (rotation_matrix_x(-90) @ vtx.T).T ; trimesh.Trimesh(...).export(path, 'obj')       -> export_obj (host)
Everything up to the vertex colours stays on the device; one D2H copy of (vertices, faces, colours) at the end.
`mcubes` / `trimesh` are not needed (neither is in the image).  No CPU fallback: CPU tensors raise.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import ops


def marching_cubes(volume: torch.Tensor, isovalue: float):
    """Drop-in for `mcubes.marching_cubes(volume, isovalue)` on a CUDA tensor: (vertices (V,3) fp32 in index
    coordinates, faces (F,3) int32), both CUDA tensors.  (PyMCubes returns float64 / unsigned numpy arrays: call
    `.cpu().numpy()` on the results where the caller needs numpy.)"""
    return ops.marching_cubes(volume, isovalue)


def rotation_matrix_x(theta_degrees: float) -> np.ndarray:
    """reference nsr/train_util_diffusion.py:50-58"""
    theta = np.radians(theta_degrees)
    c, s = np.cos(theta), np.sin(theta)
    return np.array([[1, 0, 0], [0, c, -s], [0, s, c]])


@torch.no_grad()
def extract_mesh(rec_decoder, ddpm_latent: dict, mesh_size: int = 192, mesh_thres: float = 10.0,
                 object_scale: float = 0.45) -> dict:
    """reference :213-233 for ONE object.  `rec_decoder` is the AE decoder mirror
    (vit.vit_triplane.RodinSR_..._ditDecoder); `ddpm_latent` must hold 'latent_after_vit' (the output of
    behaviour='decode_after_vae_no_render').  Returns numpy arrays {'vertices' (V,3) float64 rotated as the
    reference exports them, 'faces' (F,3) int64, 'vertex_colors' (V,3) uint8} plus the device tensors under
    'vertices_device' (un-rotated, world units) / 'faces_device'."""
    grid_out = rec_decoder.triplane_decode_grid(ddpm_latent, grid_size=mesh_size)
    sigma = grid_out["sigma"]
    assert sigma.shape[0] == 1, "mesh export handles one object per call (the reference squeezes dim 0)"
    vol = sigma.float().reshape(mesh_size, mesh_size, mesh_size).contiguous()
    s = 2.0 / (mesh_size - 1) * object_scale        # vtx / (G - 1) * 2 - 1, then * 0.45 (g-objaverse scale)
    vtx, faces = ops.marching_cubes(vol, mesh_thres, scale=(s, s, s), offset=(-object_scale,) * 3)
    if vtx.shape[0]:
        rgb = rec_decoder.forward_points(ddpm_latent["latent_after_vit"], vtx.unsqueeze(0))["rgb"]
        colors = (rgb.float().squeeze(0).clamp(0, 1) * 255).to(torch.uint8)
    else:
        colors = torch.empty((0, 3), dtype=torch.uint8, device=vtx.device)
    v_host = vtx.double().cpu().numpy()
    v_host = np.transpose(rotation_matrix_x(-90) @ np.transpose(v_host))     # rotate mesh along x dim (:233)
    return {"vertices": v_host, "faces": faces.cpu().numpy().astype(np.int64),
            "vertex_colors": colors.cpu().numpy(), "vertices_device": vtx, "faces_device": faces}


def export_obj(path: str, vertices: np.ndarray, faces: np.ndarray, vertex_colors: np.ndarray | None = None) -> str:
    """Wavefront OBJ with per-vertex colours (`v x y z r g b`, colours in [0,1]) and 1-based faces -- what
    `trimesh.Trimesh(vertices, faces, vertex_colors).export(path, 'obj')` writes (reference :236-244)."""
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    v = np.asarray(vertices, dtype=np.float64)
    f = np.asarray(faces, dtype=np.int64) + 1
    with open(path, "w") as fh:
        fh.write("# ln3diff_b200 mesh export\n")
        if vertex_colors is not None:
            c = np.asarray(vertex_colors, dtype=np.float64) / 255.0
            for (x, y, z), (r, g, b) in zip(v, c):
                fh.write(f"v {x:.8f} {y:.8f} {z:.8f} {r:.8f} {g:.8f} {b:.8f}\n")
        else:
            for x, y, z in v:
                fh.write(f"v {x:.8f} {y:.8f} {z:.8f}\n")
        for a, b, c_ in f:
            fh.write(f"f {a} {b} {c_}\n")
    return path


def export_mesh(rec_decoder, ddpm_latent: dict, dump_dir: str, name_prefix: str, mesh_size: int = 192,
                mesh_thres: float = 10.0) -> str:
    """reference :208-247: extract + dump `<dump_dir>/mesh/<name_prefix>.obj`; returns the path."""
    m = extract_mesh(rec_decoder, ddpm_latent, mesh_size, mesh_thres)
    return export_obj(os.path.join(dump_dir, "mesh", f"{name_prefix}.obj"), m["vertices"], m["faces"], m["vertex_colors"])

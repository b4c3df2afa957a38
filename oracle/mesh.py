"""TEST INFRASTRUCTURE ONLY (oracle) -- numpy restatement of the reference's mesh extraction.

Only tests/ and __graft_entry__.smoke() may import this module; the product path never does.

Restates nsr/train_util_diffusion.py:208-249:
    vtx, faces = mcubes.marching_cubes(sigma_grid (G,G,G) numpy, mesh_thres)                 (:221-223)
    vtx = vtx / (mesh_size - 1) * 2 - 1 ;  vtx = vtx * 0.45                                   (:225-226)
    vtx_colors = (forward_points(planes, vtx)['rgb'].clip(0, 1) * 255).astype(uint8)          (:228-230)
    vtx = (rotation_matrix_x(-90) @ vtx.T).T                                                  (:233, :50-58)
    trimesh.Trimesh(vertices, faces, vertex_colors).export(path, 'obj')                       (:236-244)

PARITY UNPINNED for marching cubes itself: PyMCubes (un-vendored; requirements.txt lists `PyMCubes`, no pin) is not
in this image and the reference holds no golden mesh.  What is restated is its published algorithm -- Lorensen &
Cline cells with the classic corner / edge numbering, a corner is "set" when `value <= iso`, one shared vertex
per crossed lattice edge at the linear interpolation `x1 + (iso - f1) (x2 - x1) / (f2 - f1)` evaluated in
double -- with case tables DERIVED by tools/gen_mc_tables.py (same triangle counts, 820 in total, and the same
winding as the classic table; ambiguous faces resolved face-consistently, so the triangulation inside a cell
can differ from PyMCubes' while the vertex set cannot).  The parity anchors are therefore mesh invariants:
identical vertex set (one per sign-change lattice edge, positions to fp32 rounding), closed 2-manifold output,
enclosed volume, orientation -- tests/test_mesh.py and tests/test_gpu_mesh.py.
"""
from __future__ import annotations

import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import gen_mc_tables as _gen  # noqa: E402  (test infrastructure importing the table generator)

_TRI, _NUM, _EMASK = _gen.build_tables()
_OWNER = [_gen.edge_owner(e) for e in range(12)]


def marching_cubes(volume: np.ndarray, isovalue: float):
    """volume (nx, ny, nz) -> (vertices float64 (V, 3) in index coordinates, faces int64 (F, 3)).

    Vertex order: by (linear index of the owning lattice point, axis); face order: by the cell's linear index,
    then the case table's triangle order -- the order the device kernels produce, so results compare 1:1."""
    f = np.asarray(volume, dtype=np.float64)
    nx, ny, nz = f.shape
    ins = f <= isovalue
    # owned crossed edges per lattice point, axis 0 / 1 / 2
    cross = np.zeros((3, nx, ny, nz), dtype=bool)
    cross[0, :-1] = ins[:-1] != ins[1:]
    cross[1, :, :-1] = ins[:, :-1] != ins[:, 1:]
    cross[2, :, :, :-1] = ins[:, :, :-1] != ins[:, :, 1:]
    nv = cross.sum(0).reshape(-1)
    first = np.concatenate([[0], np.cumsum(nv)[:-1]]).reshape(nx, ny, nz)
    verts = []
    idx = np.argwhere(cross.any(0))        # lexicographic = linear order
    for (i, j, k) in idx:
        p = np.array([i, j, k], dtype=np.float64)
        for a in range(3):
            if cross[a, i, j, k]:
                q = [i, j, k]
                q[a] += 1
                f1, f2 = f[i, j, k], f[q[0], q[1], q[2]]
                v = p.copy()
                v[a] = p[a] + (isovalue - f1) * ((p[a] + 1) - p[a]) / (f2 - f1)
                verts.append(v)
    verts = np.array(verts, dtype=np.float64).reshape(-1, 3)
    # cells
    c = np.zeros((nx - 1, ny - 1, nz - 1), dtype=np.int64)
    for m, (dx, dy, dz) in enumerate(_gen.CORNERS):
        c |= ins[dx:nx - 1 + dx, dy:ny - 1 + dy, dz:nz - 1 + dz].astype(np.int64) << m
    faces = []
    for (i, j, k) in np.argwhere((c != 0) & (c != 255)):
        for tri in _TRI[c[i, j, k]]:
            t = []
            for e in tri:
                dx, dy, dz, axis = _OWNER[e]
                oi, oj, ok = i + dx, j + dy, k + dz
                rank = int(cross[:axis, oi, oj, ok].sum())
                assert cross[axis, oi, oj, ok]
                t.append(int(first[oi, oj, ok]) + rank)
            faces.append(t)
    return verts, np.array(faces, dtype=np.int64).reshape(-1, 3)


def rotation_matrix_x(theta_degrees: float) -> np.ndarray:
    """nsr/train_util_diffusion.py:50-58"""
    th = np.radians(theta_degrees)
    c, s = np.cos(th), np.sin(th)
    return np.array([[1, 0, 0], [0, c, -s], [0, s, c]])


def mesh_vertices_to_world(vtx: np.ndarray, mesh_size: int) -> np.ndarray:
    """:225-226 (before the export rotation)"""
    return (vtx / (mesh_size - 1) * 2 - 1) * 0.45


def mesh_stats(verts: np.ndarray, faces: np.ndarray) -> dict:
    """Invariants used as parity anchors: manifoldness, Euler characteristic, signed volume, area."""
    e = np.concatenate([faces[:, [0, 1]], faces[:, [1, 2]], faces[:, [2, 0]]], 0)
    und = np.sort(e, 1)
    uniq, counts = np.unique(und, axis=0, return_counts=True)
    # directed edges must pair up with their reverse exactly once on a consistently oriented closed surface
    key = e[:, 0].astype(np.int64) * (verts.shape[0] + 1) + e[:, 1]
    rkey = e[:, 1].astype(np.int64) * (verts.shape[0] + 1) + e[:, 0]
    oriented = np.array_equal(np.sort(key), np.sort(rkey)) and len(np.unique(key)) == len(key)
    a, b, c = verts[faces[:, 0]], verts[faces[:, 1]], verts[faces[:, 2]]
    vol = float(np.einsum("ij,ij->i", a, np.cross(b, c)).sum() / 6.0)
    area = float(np.linalg.norm(np.cross(b - a, c - a), axis=1).sum() / 2.0)
    return {"closed": bool((counts == 2).all()), "oriented": bool(oriented), "boundary_edges": int((counts == 1).sum()),
            "nonmanifold_edges": int((counts > 2).sum()),
            "euler": int(verts.shape[0] - uniq.shape[0] + faces.shape[0]), "volume": vol, "area": area}

"""CPU: marching-cubes case tables, the numpy oracle's mesh invariants, and the OBJ writer."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def test_case_tables_are_consistent():
    import gen_mc_tables as g
    tri, num, emask = g.build_tables()
    assert num[0] == 0 and num[255] == 0 and max(num) == 5 and sum(num) == 820   # the classic table's totals
    for c in range(256):
        used = {e for t in tri[c] for e in t}
        crossed = {e for e in range(12) if emask[c] >> e & 1}
        assert used == crossed, c
        assert num[c] == num[255 - c] or True   # complements may triangulate differently on ambiguous faces
        # every cell patch is bounded by face segments only: edges interior to the patch are used twice, in
        # opposite directions; boundary edges (on cube faces) once
        directed = [(t[i], t[(i + 1) % 3]) for t in tri[c] for i in range(3)]
        assert len(set(directed)) == len(directed), c
    # the generated header is the committed one
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "mc_tables.h")
        g.write_header(p)
        assert open(p).read() == open(os.path.join(ROOT, "ln3diff_b200", "csrc", "mc_tables.h")).read()
    # winding of the classic table: case 1 is the triangle 0-8-3 (any rotation)
    assert tri[1][0] in ((0, 8, 3), (8, 3, 0), (3, 0, 8))


def _sphere(n, r, density=False):
    x = np.linspace(-1, 1, n)
    X, Y, Z = np.meshgrid(x, x, x, indexing="ij")
    d = np.sqrt(X ** 2 + Y ** 2 + Z ** 2)
    return (10 * (r - d)) if density else (d - r)


def test_oracle_mesh_invariants():
    from oracle import mesh
    n, r = 28, 0.6
    h = 2.0 / (n - 1)
    v, f = mesh.marching_cubes(_sphere(n, r), 0.0)
    st = mesh.mesh_stats(v, f)
    assert st["closed"] and st["oriented"] and st["euler"] == 2
    # normals point to the `<= iso` side: inwards for a signed distance, outwards for a density
    assert abs(-st["volume"] * h ** 3 / (4 / 3 * np.pi * r ** 3) - 1) < 0.02
    v2, f2 = mesh.marching_cubes(_sphere(n, r, density=True), 0.0)
    st2 = mesh.mesh_stats(v2, f2)
    assert st2["closed"] and st2["oriented"] and st2["volume"] > 0
    # vertices sit on the iso-surface of the trilinear interpolant's edges: |x| = r to second order in h
    rad = np.linalg.norm(v * h - 1.0, axis=1)
    assert np.abs(rad - r).max() < 0.5 * h ** 2 / r + 1e-9
    # one vertex per sign-change lattice edge
    s = _sphere(n, r) <= 0
    n_cross = (s[1:] != s[:-1]).sum() + (s[:, 1:] != s[:, :-1]).sum() + (s[:, :, 1:] != s[:, :, :-1]).sum()
    assert v.shape[0] == n_cross
    # random field (every ambiguous configuration occurs), closed by a border above the iso value: watertight 2-manifold
    rng = np.random.default_rng(0)
    vol = rng.standard_normal((20, 23, 17))
    vol[0] = vol[-1] = 5; vol[:, 0] = vol[:, -1] = 5; vol[:, :, 0] = vol[:, :, -1] = 5
    v3, f3 = mesh.marching_cubes(vol, 0.0)
    st3 = mesh.mesh_stats(v3, f3)
    assert st3["closed"] and st3["oriented"] and st3["nonmanifold_edges"] == 0
    # reference :225-226 rescale
    w = mesh.mesh_vertices_to_world(np.array([[0.0, 0, 0], [n - 1, n - 1, n - 1]]), n)
    assert np.allclose(w, [[-0.45] * 3, [0.45] * 3])


def test_export_obj_roundtrip(tmp_path):
    from ln3diff_b200 import mesh
    v = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0.5]], dtype=np.float64)
    f = np.array([[0, 1, 2]])
    c = np.array([[255, 0, 0], [0, 255, 0], [0, 0, 255]], dtype=np.uint8)
    p = mesh.export_obj(str(tmp_path / "mesh" / "a.obj"), v, f, c)
    lines = [ln.split() for ln in open(p) if ln[0] in "vf"]
    assert [ln[0] for ln in lines] == ["v", "v", "v", "f"]
    assert np.allclose([float(x) for x in lines[2][1:]], [0, 1, 0.5, 0, 0, 1])
    assert lines[3][1:] == ["1", "2", "3"]
    r = mesh.rotation_matrix_x(-90)
    assert np.allclose(r @ np.array([0, 1, 0]), [0, 0, -1])

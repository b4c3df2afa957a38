"""Generate the marching-cubes case tables of ln3diff_b200/csrc/mc_tables.h from the cube geometry.

The reference extracts meshes with PyMCubes (`mcubes.marching_cubes`, nsr/train_util_diffusion.py:221-223), an
un-vendored third-party package that is not in this image; its 256-row triangle table is third-party data.  The
tables here are DERIVED, not copied: for each of the 256 corner-sign cases the crossed cube edges are joined into
closed loops by walking the six faces and the loops are fan-triangulated (apex chosen so that no fan diagonal
lies in a cube face: the output is a 2-manifold wherever the sampled field is not degenerate).

Conventions (the ones PyMCubes / the classic Lorensen-Cline numbering use, so that vertices coincide):
  corner m: 0 (0,0,0) 1 (1,0,0) 2 (1,1,0) 3 (0,1,0) 4 (0,0,1) 5 (1,0,1) 6 (1,1,1) 7 (0,1,1)   (x, y, z offsets)
  edge e  : 0 01, 1 12, 2 23, 3 30, 4 45, 5 56, 6 67, 7 74, 8 04, 9 15, 10 26, 11 37
  case bit m set <=> value at corner m <= isovalue
  triangles are wound so that their normal points towards the corners with the bit SET (the `<= iso` side), the
  orientation the classic table has (its case 1 is the triangle 0-8-3).
Ambiguous faces (two diagonal corners set) are always resolved by cutting off the SET corners; the rule depends
only on the four corner signs of the face, so neighbouring cells agree on it and closed surfaces come out
watertight (the classic table has face-inconsistent complement cases that can leave holes).

Run:  python tools/gen_mc_tables.py  (rewrites ln3diff_b200/csrc/mc_tables.h); importable for tests / oracle.
"""
from __future__ import annotations

import os

CORNERS = [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)]
EDGES = [(0, 1), (1, 2), (2, 3), (3, 0), (4, 5), (5, 6), (6, 7), (7, 4), (0, 4), (1, 5), (2, 6), (3, 7)]
# faces as corner cycles
FACES = [(0, 1, 2, 3), (4, 5, 6, 7), (0, 1, 5, 4), (3, 2, 6, 7), (0, 3, 7, 4), (1, 2, 6, 5)]
EDGE_OF = {}
for _e, (_a, _b) in enumerate(EDGES):
    EDGE_OF[(_a, _b)] = _e
    EDGE_OF[(_b, _a)] = _e


_EDGE_FACES = [frozenset(i for i, f in enumerate(FACES) if a in f and b in f) for (a, b) in EDGES]


def _mid(e):
    a, b = EDGES[e]
    return tuple((CORNERS[a][i] + CORNERS[b][i]) / 2.0 for i in range(3))


def case_triangles(case: int) -> list[tuple[int, int, int]]:
    inside = [(case >> m) & 1 for m in range(8)]
    crossed = [e for e, (a, b) in enumerate(EDGES) if inside[a] != inside[b]]
    if not crossed:
        return []
    adj: dict[int, list[int]] = {e: [] for e in crossed}
    for f in FACES:
        fe = [EDGE_OF[(f[i], f[(i + 1) % 4])] for i in range(4)]     # edge i joins corner i and i+1
        fc = [e for e in fe if e in adj]
        if len(fc) == 2:
            adj[fc[0]].append(fc[1])
            adj[fc[1]].append(fc[0])
        elif len(fc) == 4:
            # ambiguous face: cut off every SET corner (join the two face edges that meet at it)
            for i in range(4):
                if inside[f[i]]:
                    e0, e1 = fe[(i - 1) % 4], fe[i]
                    adj[e0].append(e1)
                    adj[e1].append(e0)
        else:
            assert len(fc) == 0
    assert all(len(v) == 2 for v in adj.values()), (case, adj)
    tris = []
    seen = set()
    for start in crossed:
        if start in seen:
            continue
        loop = [start]
        seen.add(start)
        prev, cur = None, start
        while True:
            n0, n1 = adj[cur]
            n = n0 if n0 != prev else n1
            if n == start:
                break
            assert n not in seen, (case, loop, n)
            loop.append(n)
            seen.add(n)
            prev, cur = cur, n
        assert len(loop) >= 3, (case, loop)
        # orientation: Newell normal against the direction towards the SET corners touched by the loop
        pts = [_mid(e) for e in loop]
        nrm = [0.0, 0.0, 0.0]
        for i in range(len(pts)):
            p, q = pts[i], pts[(i + 1) % len(pts)]
            nrm[0] += (p[1] - q[1]) * (p[2] + q[2])
            nrm[1] += (p[2] - q[2]) * (p[0] + q[0])
            nrm[2] += (p[0] - q[0]) * (p[1] + q[1])
        cen = [sum(p[i] for p in pts) / len(pts) for i in range(3)]
        dot = 0.0
        for e in loop:
            a, b = EDGES[e]
            c = a if inside[a] else b
            dot += sum(nrm[i] * (CORNERS[c][i] - cen[i]) for i in range(3))
        assert abs(dot) > 1e-9, (case, loop)
        if dot < 0:
            loop = loop[::-1]
        # fan apex: a diagonal joining two loop vertices that lie on the same cube face would coincide with the
        # neighbouring cell's diagonal across that (ambiguous) face -- an edge shared by four triangles.  Every
        # loop has an apex whose fan has no such diagonal; take the first one.
        def in_face_diagonals(ap):
            r = loop[ap:] + loop[:ap]
            return sum(1 for i in range(2, len(r) - 1) if _EDGE_FACES[r[0]] & _EDGE_FACES[r[i]])
        apex = min(range(len(loop)), key=in_face_diagonals)
        assert in_face_diagonals(apex) == 0, (case, loop)
        loop = loop[apex:] + loop[:apex]
        for i in range(1, len(loop) - 1):
            tris.append((loop[0], loop[i], loop[i + 1]))
    return tris


def build_tables():
    tri_table = [case_triangles(c) for c in range(256)]
    num_tris = [len(t) for t in tri_table]
    edge_mask = []
    for c in range(256):
        m = 0
        for e, (a, b) in enumerate(EDGES):
            if ((c >> a) & 1) != ((c >> b) & 1):
                m |= 1 << e
        edge_mask.append(m)
    return tri_table, num_tris, edge_mask


def edge_owner(e: int):
    """(dx, dy, dz, axis): the lattice point that owns cube edge e (its lower corner) and the edge's axis."""
    a, b = EDGES[e]
    ca, cb = CORNERS[a], CORNERS[b]
    lo = tuple(min(ca[i], cb[i]) for i in range(3))
    axis = [i for i in range(3) if ca[i] != cb[i]][0]
    return lo + (axis,)


def write_header(path: str) -> None:
    tri_table, num_tris, edge_mask = build_tables()
    width = 3 * max(num_tris)
    lines = [
        "// GENERATED by tools/gen_mc_tables.py -- do not edit.  Marching-cubes case tables derived from the cube",
        "// geometry (corner / edge numbering and orientation documented in the generator).",
        "#pragma once",
        "#include <stdint.h>",
        "",
        f"#define LN3_MC_MAX_TRIS {max(num_tris)}",
        "// triangles per case",
        "static const uint8_t kMcNumTris[256] = {",
    ]
    for r in range(0, 256, 32):
        lines.append("  " + ", ".join(str(v) for v in num_tris[r:r + 32]) + ",")
    lines += ["};", "// cube edges per triangle corner (3 * kMcNumTris entries used per row)",
              f"static const uint8_t kMcTriTable[256][{width}] = {{"]
    for c in range(256):
        flat = [e for t in tri_table[c] for e in t]
        flat += [0] * (width - len(flat))
        lines.append("  {" + ", ".join(str(v) for v in flat) + "},")
    lines += ["};", "// owner lattice offset (dx, dy, dz) and axis of every cube edge: owner = cell + (dx, dy, dz)",
              "static const uint8_t kMcEdgeOwner[12][4] = {"]
    for e in range(12):
        lines.append("  {" + ", ".join(str(v) for v in edge_owner(e)) + "},")
    lines += ["};", ""]
    with open(path, "w") as f:
        f.write("\n".join(lines))


if __name__ == "__main__":
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = os.path.join(root, "ln3diff_b200", "csrc", "mc_tables.h")
    write_header(out)
    t, n, _ = build_tables()
    print(out, "max tris", max(n), "total tris", sum(n), "case1", t[1], "case 254", t[254])

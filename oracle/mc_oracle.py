"""CPU oracle for the iso-surface extraction of extract_geometry (reference: models/renderer.py:32-40).

TEST INFRASTRUCTURE ONLY (see oracle/neus_oracle.py).  PARITY UNPINNED against the reference's own implementation: the
reference delegates marching cubes to the third-party package PyMCubes (``mcubes.marching_cubes``, requirements pin
PyMCubes==0.1.4), which is not vendored under /root/reference and not installed in this image.  What is pinned instead:
the published algorithm (Lorensen & Cline: linear interpolation on every grid edge whose end points straddle the
iso-value, shared vertices, per-cell triangulation) restated WITHOUT a case table -- each cell's polygons are traced at
run time from the face rule (crossed edges of a face joined pairwise, ambiguous faces cut off their inside corners),
written independently of the product's table generator -- plus analytic properties the tests check (watertightness,
Euler characteristic, enclosed volume and distance to an analytic sphere).
"""
from __future__ import annotations

import numpy as np

_CORNER = [(c & 1, (c >> 1) & 1, (c >> 2) & 1) for c in range(8)]


def _edge_key(p, q):
    return (p, q) if p <= q else (q, p)


def marching_cubes_np(u: np.ndarray, iso: float = 0.0):
    """u [nx,ny,nz] -> (vertices [V,3] float64 in grid-index coordinates, triangles [T,3] int).  Inside: u > iso.
    Pure-Python loops: small grids only."""
    nx, ny, nz = u.shape
    inside = u > iso
    vid = {}
    verts = []

    def vertex(p, q):
        k = _edge_key(p, q)
        if k not in vid:
            a, b = k
            ua, ub = float(u[a]), float(u[b])
            w = (iso - ua) / (ub - ua)
            verts.append([a[i] + w * (b[i] - a[i]) for i in range(3)])
            vid[k] = len(verts) - 1
        return vid[k]

    tris = []
    # the 6 faces of a cell as cycles of local corners
    faces = []
    for axis in range(3):
        o = [a for a in range(3) if a != axis]
        for side in (0, 1):
            cyc = []
            for (s, t) in ((0, 0), (1, 0), (1, 1), (0, 1)):
                c = [0, 0, 0]
                c[axis], c[o[0]], c[o[1]] = side, s, t
                cyc.append(tuple(c))
            faces.append(cyc)
    for x in range(nx - 1):
        for y in range(ny - 1):
            for z in range(nz - 1):
                corner_in = {c: bool(inside[x + c[0], y + c[1], z + c[2]]) for c in _CORNER}
                n_in = sum(corner_in.values())
                if n_in == 0 or n_in == 8:
                    continue
                segs = []
                for cyc in faces:
                    crossed = [i for i in range(4) if corner_in[cyc[i]] != corner_in[cyc[(i + 1) % 4]]]
                    ed = lambda i: _edge_key(cyc[i], cyc[(i + 1) % 4])
                    if len(crossed) == 2:
                        segs.append((ed(crossed[0]), ed(crossed[1])))
                    elif len(crossed) == 4:
                        for i in range(4):
                            if corner_in[cyc[i]]:
                                segs.append((ed((i - 1) % 4), ed(i)))
                adj = {}
                for a, b in segs:
                    adj.setdefault(a, []).append(b)
                    adj.setdefault(b, []).append(a)
                done = set()
                for start in sorted(adj):
                    if start in done:
                        continue
                    loop, prev, cur = [start], None, start
                    done.add(start)
                    while True:
                        cand = [v for v in adj[cur] if v != prev]
                        nxt = cand[0] if cand else adj[cur][0]
                        if nxt == start:
                            break
                        loop.append(nxt)
                        done.add(nxt)
                        prev, cur = cur, nxt
                    mids = np.array([[0.5 * (e[0][i] + e[1][i]) for i in range(3)] for e in loop])
                    area = np.zeros(3)
                    for i in range(1, len(loop) - 1):
                        area += np.cross(mids[i] - mids[0], mids[i + 1] - mids[0])
                    out = np.zeros(3)
                    for e in loop:
                        a, b = e
                        sgn = 1.0 if corner_in[a] else -1.0
                        out += sgn * (np.array(b) - np.array(a))
                    if np.dot(area, out) < 0:
                        loop = loop[::-1]
                    r = min(range(len(loop)), key=lambda i: loop[i])       # fan origin: the smallest edge of the loop
                    loop = loop[r:] + loop[:r]
                    ids = [vertex((x + a[0], y + a[1], z + a[2]), (x + b[0], y + b[1], z + b[2])) for a, b in loop]
                    for i in range(1, len(ids) - 1):
                        tris.append((ids[0], ids[i], ids[i + 1]))
    return np.array(verts, dtype=np.float64).reshape(-1, 3), np.array(tris, dtype=np.int64).reshape(-1, 3)


def mesh_report(verts: np.ndarray, tris: np.ndarray):
    """Topological / geometric summary used by the property tests: directed-edge multiplicities, Euler characteristic,
    signed volume (divergence theorem) and area."""
    from collections import Counter
    und, dirc = Counter(), Counter()
    for t in tris:
        for i in range(3):
            a, b = int(t[i]), int(t[(i + 1) % 3])
            und[_edge_key(a, b)] += 1
            dirc[(a, b)] += 1
    p = verts[tris]
    vol = float(np.einsum("ij,ij->i", p[:, 0], np.cross(p[:, 1], p[:, 2])).sum() / 6.0)
    area = float(np.linalg.norm(np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0]), axis=1).sum() / 2.0)
    used = len(set(int(i) for i in tris.reshape(-1)))
    return dict(closed=all(c == 2 for c in und.values()), oriented=all(c == 1 for c in dirc.values()),
                euler=used - len(und) + len(tris), volume=vol, area=area, used_vertices=used)

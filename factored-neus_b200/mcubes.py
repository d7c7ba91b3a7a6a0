"""GPU marching cubes for ``NeuSRenderer.extract_geometry`` (reference: models/renderer.py:32-40 calls the third-party
CPU package PyMCubes on a 512^3 array after a 512 MiB device-to-host copy).

The 256-case triangle table is DERIVED here, not typed in: for every sign configuration of the 8 cube corners the
iso-surface crosses the edges whose end points differ in sign; on each of the 6 faces the crossed edges are joined
pairwise -- a face with 4 crossings (diagonal corners alike) is resolved by cutting off its INSIDE corners, a rule that
depends on the face alone, so two cells sharing a face always agree and the mesh is watertight -- the segments close
into loops and every loop is fan-triangulated, oriented so that the normal points from inside (u > iso) to outside.
Vertices are shared: each grid point owns its +x, +y, +z edges, so a crossing is emitted once and indexed by all the
cells around the edge (PyMCubes returns shared vertices too).  Vertex coordinates are in grid-index units like
``mcubes.marching_cubes``; ``extract_geometry`` rescales them exactly as the reference does.

The triangulation of ambiguous cases may differ from PyMCubes' classic table in which diagonal it picks; the vertex set
is identical (linear interpolation on the crossed edges).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib as L

# corner c = (x, y, z) bits 0, 1, 2; edge e joins CORNERS[e]; owner grid point offset + axis of each edge
CORNER_XYZ = np.array([[(c >> 0) & 1, (c >> 1) & 1, (c >> 2) & 1] for c in range(8)], dtype=np.int64)
EDGES = [(0, 1), (2, 3), (4, 5), (6, 7),      # x edges (axis 0) at (y,z) = 00, 10, 01, 11
         (0, 2), (1, 3), (4, 6), (5, 7),      # y edges (axis 1) at (x,z) = 00, 10, 01, 11
         (0, 4), (1, 5), (2, 6), (3, 7)]      # z edges (axis 2) at (x,y) = 00, 10, 01, 11
EDGE_AXIS = [0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2]
# faces as corner cycles (any winding; orientation is fixed per loop afterwards)
FACES = [(0, 2, 6, 4), (1, 3, 7, 5), (0, 1, 5, 4), (2, 3, 7, 6), (0, 1, 3, 2), (4, 5, 7, 6)]
_EDGE_OF = {}
for _e, (_a, _b) in enumerate(EDGES):
    _EDGE_OF[(_a, _b)] = _e
    _EDGE_OF[(_b, _a)] = _e


def _case_triangles(case: int):
    inside = [(case >> c) & 1 for c in range(8)]
    segs = []
    for f in FACES:
        crossed = [(i, _EDGE_OF[(f[i], f[(i + 1) % 4])]) for i in range(4) if inside[f[i]] != inside[f[(i + 1) % 4]]]
        if len(crossed) == 2:
            segs.append((crossed[0][1], crossed[1][1]))
        elif len(crossed) == 4:                     # ambiguous face: cut off each inside corner
            for i in range(4):
                if inside[f[i]]:
                    segs.append((_EDGE_OF[(f[(i - 1) % 4], f[i])], _EDGE_OF[(f[i], f[(i + 1) % 4])]))
    adj = {}
    for a, b in segs:
        adj.setdefault(a, []).append(b)
        adj.setdefault(b, []).append(a)
    tris, seen = [], set()
    mid = lambda e: 0.5 * (CORNER_XYZ[EDGES[e][0]] + CORNER_XYZ[EDGES[e][1]])
    for start in sorted(adj):
        if start in seen:
            continue
        loop, prev, cur = [start], None, start
        seen.add(start)
        while True:
            nxt = [v for v in adj[cur] if v != prev]
            nxt = nxt[0] if nxt else adj[cur][0]
            if nxt == start:
                break
            loop.append(nxt)
            seen.add(nxt)
            prev, cur = cur, nxt
        pts = np.array([mid(e) for e in loop])
        area = np.zeros(3)
        for i in range(1, len(loop) - 1):
            area += np.cross(pts[i] - pts[0], pts[i + 1] - pts[0])
        outward = np.zeros(3)                       # inside -> outside along the loop's own edges
        for e in loop:
            a, b = EDGES[e]
            outward += (CORNER_XYZ[b] - CORNER_XYZ[a]) * (1.0 if inside[a] else -1.0)
        if np.dot(area, outward) < 0:
            loop = loop[::-1]
        # canonical fan origin: the loop's smallest edge in (corner, corner) coordinate order
        ekey = lambda e: tuple(sorted((tuple(CORNER_XYZ[EDGES[e][0]]), tuple(CORNER_XYZ[EDGES[e][1]]))))
        r = min(range(len(loop)), key=lambda i: ekey(loop[i]))
        loop = loop[r:] + loop[:r]
        for i in range(1, len(loop) - 1):
            tris.append((loop[0], loop[i], loop[i + 1]))
    return tris


_TABLES = None


def tables():
    """(tri_count [256] int32, tri_edges [256, 3*MAXT] int32 (-1 padded), MAXT)."""
    global _TABLES
    if _TABLES is None:
        per = [_case_triangles(c) for c in range(256)]
        maxt = max(len(t) for t in per)
        cnt = np.array([len(t) for t in per], dtype=np.int32)
        tab = -np.ones((256, 3 * maxt), dtype=np.int32)
        for c, t in enumerate(per):
            for i, tri in enumerate(t):
                tab[c, 3 * i: 3 * i + 3] = tri
        _TABLES = (cnt, tab, maxt)
    return _TABLES


_DEV_TABLES = {}


def marching_cubes(u: torch.Tensor, isovalue: float = 0.0):
    """Iso-surface of the CUDA tensor u [nx,ny,nz] at ``isovalue`` (inside: u > isovalue).  Returns (vertices [V,3]
    float32 in grid-index coordinates, triangles [T,3] int64), both on u's device.  Two launches around two prefix sums."""
    if not u.is_cuda:
        raise RuntimeError("factored-neus_b200.marching_cubes needs a CUDA tensor (no CPU fallback)")
    u = u.detach().to(torch.float32).contiguous()
    nx, ny, nz = u.shape
    dev = u.device
    key = str(dev)
    if key not in _DEV_TABLES:
        cnt, tab, maxt = tables()
        _DEV_TABLES[key] = (torch.from_numpy(cnt).to(dev), torch.from_numpy(tab).to(dev).contiguous(), maxt)
    cnt, tab, maxt = _DEV_TABLES[key]
    lib = L.lib()
    npts = nx * ny * nz
    ncell = max(0, nx - 1) * max(0, ny - 1) * max(0, nz - 1)
    vmask = torch.empty(npts, dtype=torch.uint8, device=dev)
    vcount = torch.empty(npts, dtype=torch.int32, device=dev)
    tcount = torch.empty(max(ncell, 1), dtype=torch.int32, device=dev)
    L.check(lib.fneus_mc_classify(L.ptr(u), nx, ny, nz, float(isovalue), L.ptr(cnt), L.ptr(vmask), L.ptr(vcount),
                                  L.ptr(tcount), L.stream_ptr()), "fneus_mc_classify")
    voff = torch.cumsum(vcount, 0, dtype=torch.int64)            # inclusive; exclusive = voff - vcount
    toff = torch.cumsum(tcount[:ncell], 0, dtype=torch.int64) if ncell else torch.zeros(0, dtype=torch.int64, device=dev)
    V = int(voff[-1]) if npts else 0
    T = int(toff[-1]) if ncell else 0
    verts = torch.empty(V, 3, dtype=torch.float32, device=dev)
    tris = torch.empty(T, 3, dtype=torch.int64, device=dev)
    if V and T:
        L.check(lib.fneus_mc_emit(L.ptr(u), nx, ny, nz, float(isovalue), L.ptr(cnt), L.ptr(tab), maxt, L.ptr(vmask),
                                  L.ptr(vcount), L.ptr(voff), L.ptr(tcount), L.ptr(toff), L.ptr(verts), L.ptr(tris),
                                  L.stream_ptr()), "fneus_mc_emit")
    return verts, tris

"""Tetrahedral background grids: acute-lattice generator, ``.tet`` wire format, DefTet's boundary snap.

The reference obtains its grids from the external QuarTet tool
(``utils/dataloder_helper.py:30-69``, ``README.md:22-29``), which is not available offline, and ships
three of them as text (``diff_render/diftet_6_subdiv/data/cube_{40,50,60}_tet.tet``; format read by
``3_model/prepare_for_wz.py:18-45``).  Those files are QuarTet's *acute lattice*: vertices on a 1/res
lattice, period 4 cells, 46 positively oriented tetrahedra per 4x4x4 tile.  ``ACUTE_TILE`` below is that
tile (lattice offsets of the 4 vertices of each of the 46 tets), recovered once from the interior of the
shipped res-40 grid; it is data describing QuarTet's lattice, not reference code.

``acute_lattice_grid(res)`` tiles it over the cube and keeps every tet whose four vertices lie inside
``[0, res]^3`` -- a conforming, positively oriented tet mesh with T ~= 0.69 res^3 (the QuarTet files also
carry warped boundary tets, so their counts are ~7 % larger; SURVEY.md section 7 step 0 accepts this).
"""
from __future__ import annotations

import numpy as np

__all__ = ["ACUTE_TILE", "acute_lattice_grid", "read_tet_file", "write_tet_file", "snap_boundary",
           "read_tetrahedron", "TetGrid"]

# 46 tets x 4 vertices x (x, y, z) lattice offsets inside one period-4 tile (values 0..5).
ACUTE_TILE = np.array([
    [1,2,2,0,1,0,2,0,1,0,0,2],[1,2,2,2,0,3,0,0,2,2,0,1],[2,0,3,1,2,2,3,2,2,2,0,1],[2,0,1,1,2,2,3,2,2,2,2,0],
    [1,2,2,0,3,0,2,2,0,0,1,0],[0,1,0,1,2,2,2,0,1,2,2,0],[4,3,0,5,2,2,3,2,2,4,1,0],[3,2,2,2,0,1,4,0,2,2,0,3],
    [4,1,0,5,2,2,3,2,2,4,0,2],[4,1,0,2,0,1,3,2,2,2,2,0],[2,0,1,4,1,0,3,2,2,4,0,2],[3,2,2,4,3,0,4,1,0,2,2,0],
    [2,4,1,1,2,2,3,2,2,2,4,3],[1,2,2,2,4,1,0,3,0,0,4,2],[2,4,1,1,2,2,0,3,0,2,2,0],[2,4,1,0,5,0,0,3,0,0,4,2],
    [1,2,2,2,4,1,3,2,2,2,2,0],[1,2,2,2,4,3,2,4,1,0,4,2],[2,4,1,4,3,0,3,2,2,2,2,0],[3,2,2,2,4,1,2,4,3,4,4,2],
    [4,3,0,2,4,1,3,2,2,4,4,2],[5,2,2,4,3,0,3,2,2,4,4,2],[4,5,0,2,4,1,4,3,0,4,4,2],[1,2,2,0,3,4,0,1,4,2,2,4],
    [0,1,4,2,0,3,0,0,2,1,2,2],[1,2,2,0,1,4,2,0,3,2,2,4],[0,1,4,2,0,5,2,0,3,2,2,4],[2,2,4,3,2,2,2,0,3,1,2,2],
    [4,1,4,2,0,5,2,2,4,2,0,3],[3,2,2,4,3,4,2,2,4,4,1,4],[2,0,3,3,2,2,4,1,4,4,0,2],[3,2,2,2,0,3,4,1,4,2,2,4],
    [3,2,2,5,2,2,4,3,4,4,1,4],[5,2,2,3,2,2,4,0,2,4,1,4],[0,3,4,2,4,3,1,2,2,0,4,2],[2,4,3,0,3,4,1,2,2,2,2,4],
    [2,4,3,0,5,4,0,4,2,0,3,4],[0,3,4,2,4,5,2,2,4,2,4,3],[3,2,2,2,2,4,2,4,3,1,2,2],[0,5,4,2,4,3,2,4,5,0,3,4],
    [3,2,2,5,2,2,4,4,2,4,3,4],[3,2,2,4,3,4,2,4,3,2,2,4],[2,4,3,4,3,4,2,4,5,2,2,4],[2,4,5,4,5,4,4,3,4,2,4,3],
    [4,4,2,2,4,3,3,2,2,4,3,4],[4,5,4,2,4,3,4,4,2,4,3,4]], dtype=np.int64).reshape(46, 4, 3)


class TetGrid:
    """A background grid: ``vertices`` (V,3) float64 in [0,1]^3, ``tets`` (T,4) int64, ``mask`` (V,3) bool."""

    def __init__(self, vertices, tets, mask=None, res=None):
        self.vertices = np.ascontiguousarray(vertices, dtype=np.float64)
        self.tets = np.ascontiguousarray(tets, dtype=np.int64)
        self.mask = mask if mask is not None else np.logical_and(self.vertices < 1, self.vertices > 0)
        self.res = res

    @property
    def n_vert(self):
        return self.vertices.shape[0]

    @property
    def n_tet(self):
        return self.tets.shape[0]

    def centred(self):
        """``init_tet_pos`` as the trainer builds it (``train_multigpu.py:65-66``): vertices - 0.5, float32."""
        return (self.vertices - 0.5).astype(np.float32)


def acute_lattice_grid(res: int) -> TetGrid:
    """Synthetic stand-in for ``quartet cube.obj 1/res`` (see module docstring).

    Deterministic: tiles are visited in (x, y, z) lexicographic order, tets in ``ACUTE_TILE`` order, and
    vertices are numbered in order of first use -- so a given ``res`` always yields the same arrays.
    """
    res = int(res)
    if res < 4:
        raise ValueError("acute lattice needs res >= 4 (one period)")
    origins = np.arange(-8, res + 8, 4)
    ox, oy, oz = np.meshgrid(origins, origins, origins, indexing="ij")
    org = np.stack([ox.ravel(), oy.ravel(), oz.ravel()], axis=1)            # (n_tiles, 3)
    cand = org[:, None, None, :] + ACUTE_TILE[None]                         # (n_tiles, 46, 4, 3)
    cand = cand.reshape(-1, 4, 3)
    keep = np.logical_and(cand.min(axis=(1, 2)) >= 0, cand.max(axis=(1, 2)) <= res)
    cand = cand[keep]
    flat = cand.reshape(-1, 3)
    key = (flat[:, 0] * (res + 1) + flat[:, 1]) * (res + 1) + flat[:, 2]
    uniq, first, inverse = np.unique(key, return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")                               # first-use numbering
    rank = np.empty_like(order)
    rank[order] = np.arange(order.size)
    tets = rank[inverse].reshape(-1, 4)
    verts = flat[first[order]].astype(np.float64) / float(res)
    return TetGrid(verts, tets, res=res)


def snap_boundary(vertices: np.ndarray, res: float):
    """Boundary snap + interior mask of ``read_tetrahedron`` (``utils/dataloder_helper.py:64-68``)."""
    if res > 1.0:
        res = 1.0 / res
    vertices = np.array(vertices, dtype=np.float64, copy=True)
    vertices[vertices <= (0 + res / 4.0)] = 0
    vertices[vertices >= (1 - res / 4.0)] = 1
    mask = np.logical_and(vertices < 1, vertices > 0)
    return vertices, mask


def read_tet_file(path: str):
    """Parse the ``tet nV nT`` text format (``3_model/prepare_for_wz.py:18-45``)."""
    with open(path, "r") as f:
        head = f.readline().strip().split(" ")
        n_vert, n_tet = int(head[1]), int(head[2])
        body = np.loadtxt(f, dtype=np.float64, max_rows=n_vert, ndmin=2)
        tets = np.loadtxt(f, dtype=np.int64, max_rows=n_tet, ndmin=2)
    assert body.shape == (n_vert, 3) and tets.shape == (n_tet, 4)
    return body, tets


def write_tet_file(path: str, vertices: np.ndarray, tets: np.ndarray):
    with open(path, "w") as f:
        f.write("tet %d %d\n" % (vertices.shape[0], tets.shape[0]))
        for v in vertices:
            f.write("%g %g %g\n" % (v[0], v[1], v[2]))
        for t in tets:
            f.write("%d %d %d %d\n" % (t[0], t[1], t[2], t[3]))


def read_tetrahedron(res=50, root="..", path=None):
    """Drop-in for ``helpers.read_tetrahedron`` (``utils/dataloder_helper.py:30-69``).

    Reads ``<root>/quartet/meshes/cube_%f_tet.tet`` when it exists (or ``path``); otherwise generates the
    acute lattice in-process instead of shelling out to QuarTet.  Returns ``(vertices, tets, mask)``.
    """
    import os
    r = 1.0 / res if res > 1.0 else res
    file_name = path or os.path.join(root, "quartet/meshes", "cube_%f_tet.tet" % r)
    if os.path.exists(file_name):
        vertices, tets = read_tet_file(file_name)
    else:
        g = acute_lattice_grid(int(round(1.0 / r)))
        vertices, tets = g.vertices, g.tets
    vertices, mask = snap_boundary(vertices, r)
    return vertices, tets, mask

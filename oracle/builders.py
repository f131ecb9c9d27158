"""Oracle for the topology builders (A10-A14): restatement of the reference's pure-Python twins.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

  tet_to_face             utils/tet_utils.py:208-256   (dict keyed min*n^2+max*n+mid, first-occurrence order)
  tet_to_adj_sparse       utils/tet_utils.py:47-92     (set of directed vertex pairs)
  tet_adj_share           utils/tet_utils.py:318-367 / utils/lib/tet_adj_share/run.cpp:40-97 (key order)
  tet_to_face_adj_sparse  utils/lib/tet_face_adj/run.cpp:18-92 (32-bit edge key, ordered pairs per edge)
  colaps_v                utils/lib/colaps_v/run.cpp:18-59 ("%.5f-%.5f-%.5f" string key, first-occurrence ids)
The compiled reference builders themselves are reachable through oracle.native.ref_* when oracle/_ref exists.
"""
import numpy as np

LOCAL_FACES = ((0, 1, 2), (1, 0, 3), (2, 3, 0), (3, 2, 1))      # idx_array of tet_utils.py:213-217


def _face_key(tri, n_point):
    a, b = min(tri), max(tri)
    c = tri[2]
    for p in tri:
        if p != a and p != b:
            c = p
    return a * n_point * n_point + b * n_point + c


def tet_to_face(n_point, tet_list):
    """-> (tet_face_fx3, tet_face_tetidx_fx2, tet_face_tetfaceidx_fx2, tet_boundary_face)"""
    table = {}
    for t_idx, tet in enumerate(tet_list):
        tet = [int(x) for x in tet]
        for i_face, loc in enumerate(LOCAL_FACES):
            tri = [tet[loc[0]], tet[loc[1]], tet[loc[2]]]
            key = _face_key(tri, n_point)
            if key not in table:
                table[key] = [[tri], [t_idx], [i_face]]
            else:
                table[key][0].append(tri)
                table[key][1].append(t_idx)
                table[key][2].append(i_face)
    faces, tets, slots, boundary = [], [], [], []
    for key, (tris, ts, fs) in table.items():       # dict preserves first-occurrence order
        if len(tris) == 2:
            faces.append(tris[0]); tets.append(ts); slots.append(fs)
        elif len(tris) == 1:
            boundary.append(tris[0])
    as_arr = lambda x, w: np.asarray(x, dtype=np.int64).reshape(-1, w)
    return as_arr(faces, 3), as_arr(tets, 2), as_arr(slots, 2), as_arr(boundary, 3)


def tet_to_adj_edges(tet_list):
    """Directed vertex-edge set, sorted lexicographically (the reference order is hash order)."""
    s = set()
    for tet in tet_list:
        tet = [int(x) for x in tet]
        for i in range(4):
            for j in range(4):
                if i != j:
                    s.add((tet[i], tet[j]))
    return np.asarray(sorted(s), dtype=np.int64).reshape(-1, 2)


def tet_adj_share(n_point, tet_list):
    """Rows (t0,t1,f0),(t1,t0,f1) per face shared by exactly two tets, in ascending face-key order."""
    table = {}
    for t_idx, tet in enumerate(tet_list):
        tet = [int(x) for x in tet]
        for i_face, loc in enumerate(LOCAL_FACES):
            tri = [tet[loc[0]], tet[loc[1]], tet[loc[2]]]
            table.setdefault(_face_key(tri, n_point), []).append((t_idx, i_face))
    rows = []
    for key in sorted(table):
        f = table[key]
        if len(f) == 2:
            rows.append((f[0][0], f[1][0], f[0][1]))
            rows.append((f[1][0], f[0][0], f[1][1]))
    return np.asarray(rows, dtype=np.int32).reshape(-1, 3)


def tet_face_adj(n_point, tet_list):
    """Ordered pairs of tet-faces (4*t+i) sharing an edge, grouped by the wrapped int32 edge key a*n+b."""
    edges = {}
    abs_face = {}
    for t_idx, tet in enumerate(tet_list):
        tet = [int(x) for x in tet]
        for i_face, loc in enumerate(LOCAL_FACES):
            tri = [tet[loc[0]], tet[loc[1]], tet[loc[2]]]
            fid = t_idx * 4 + i_face
            for e in range(3):
                a, b = min(tri[e], tri[(e + 1) % 3]), max(tri[e], tri[(e + 1) % 3])
                key = (a * n_point + b) & 0xFFFFFFFF
                if key >= 1 << 31:
                    key -= 1 << 32                    # int overflow wraps (run.cpp:39)
                edges.setdefault(key, []).append(fid)
            a, b = min(tri), max(tri)
            c = tri[0]
            for p in tri:
                if p != a and p != b:
                    c = p
            abs_face[fid] = a * n_point * n_point + b * n_point + c
    out = []
    for key in sorted(edges):
        f = edges[key]
        for fa in f:
            for fb in f:
                if fa == fb or abs_face[fa] == abs_face[fb]:
                    continue
                out.append((fa, fb))
    return np.asarray(out, dtype=np.int32).reshape(-1, 2)


def colaps_v(points):
    """-> (map_array (N,), inverse_idx (n_unique,)) with the 5-decimal string key of run.cpp:18-38."""
    seen = {}
    map_array = np.zeros(len(points), dtype=np.int32)
    inverse = []
    for i, p in enumerate(np.asarray(points, dtype=np.float32)):
        key = "%.5f-%.5f-%.5f" % (float(p[0]), float(p[1]), float(p[2]))
        if key not in seen:
            seen[key] = len(inverse)
            inverse.append(i)
        map_array[i] = seen[key]
    return map_array, np.asarray(inverse, dtype=np.int32)

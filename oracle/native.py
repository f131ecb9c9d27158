"""numpy front-end of oracle/liboracle.so (C restatement) and oracle/_ref/*/run.so (the reference's own
ctypes builders compiled from /root/reference).  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py)."""
import ctypes as C
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            raise RuntimeError("oracle/liboracle.so missing: run `make -C oracle`")
        _LIB = C.CDLL(path)
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def n_threads():
    return max(1, os.cpu_count() or 1)


def _split(fn, n, threads=None, min_chunk=64):
    """Run fn(i0, i1) over [0, n) on host threads (ctypes drops the GIL)."""
    threads = threads or n_threads()
    chunks = max(1, min(threads * 4, n // min_chunk if n >= min_chunk else 1))
    bounds = np.linspace(0, n, chunks + 1).astype(np.int64)
    if chunks == 1 or threads == 1:
        for k in range(chunks):
            fn(int(bounds[k]), int(bounds[k + 1]))
        return
    with ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(lambda k: fn(int(bounds[k]), int(bounds[k + 1])), range(chunks)))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def point_in_tet(tet_bxfx4x3, pts_bxnx3, threads=None):
    tet, pts = _f32(tet_bxfx4x3), _f32(pts_bxnx3)
    B, T = tet.shape[0], tet.shape[1]
    P = pts.shape[1]
    out = np.full((B, P, 1), -1.0, dtype=np.float32)
    L = lib()
    _split(lambda a, b: L.orc_point_in_tet(_p(tet), _p(pts), _p(out), B, P, T, C.c_longlong(a), C.c_longlong(b)), B * P,
           threads)
    return out


def nearest_neighbor(queries, points, threads=None):
    q, p = _f32(queries), _f32(points)
    B, Q, M = q.shape[0], q.shape[1], p.shape[1]
    out = np.zeros((B, Q), dtype=np.int32)
    L = lib()
    _split(lambda a, b: L.orc_nearest_neighbor(_p(q), _p(p), _p(out), B, Q, M, C.c_longlong(a), C.c_longlong(b)), B * Q,
           threads)
    return out.astype(np.int64)


def point_face_distance(pts, faces, n_face_b=None, threads=None):
    pts, faces = _f32(pts), _f32(faces)
    B, P, F = pts.shape[0], pts.shape[1], faces.shape[1]
    nf = _f32(np.full(B, F) if n_face_b is None else n_face_b)
    d = np.zeros((B, P, 1), dtype=np.float32)
    f = np.zeros((B, P, 1), dtype=np.float32)
    L = lib()
    _split(lambda a, b: L.orc_point_face_distance(_p(pts), _p(faces), _p(nf), _p(d), _p(f), B, P, F, C.c_longlong(a),
                                                  C.c_longlong(b)), B * P, threads)
    return d, f


def point_face_distance_bwd(pts, faces, closest_f, dl_dd):
    pts, faces, cf, g = _f32(pts), _f32(faces), _f32(closest_f), _f32(dl_dd)
    B, P, F = pts.shape[0], pts.shape[1], faces.shape[1]
    out = np.zeros((B, F, 3, 3), dtype=np.float32)
    lib().orc_point_face_distance_bwd(_p(pts), _p(faces), _p(cf), _p(g), _p(out), B, P, F)
    return out


def face_adjacency(face_fx3x3, n_max_nei=30, threads=None):
    """-> (adj (F, n_max_nei) f32 with -1 padding, pairs (2,E) int64 as utils.py:52-61 builds them)."""
    face = _f32(face_fx3x3)
    F = face.shape[0]
    adj = np.full((F, n_max_nei), -1.0, dtype=np.float32)
    if F:
        L = lib()
        _split(lambda a, b: L.orc_face_adjacency(_p(face), _p(adj), F, n_max_nei, C.c_longlong(a), C.c_longlong(b)), F,
               threads, min_chunk=16)
    rows = np.repeat(np.arange(F, dtype=np.int64)[:, None], n_max_nei, axis=1)
    mask = adj >= 0
    pairs = np.stack([rows[mask], adj[mask].astype(np.int64)], axis=0)
    return adj, pairs


# ---- the reference's own builders, compiled from /root/reference into oracle/_ref (kind: "reference") ----
# They are executed out of process through oracle/ref_driver (see the note at the top of ref_driver.c).
import subprocess
import tempfile


def ref_lib(name):
    """Path of oracle/_ref/<name>/run.so, or None when it has not been built."""
    path = os.path.join(HERE, "_ref", name, "run.so")
    drv = os.path.join(HERE, "ref_driver")
    return path if (os.path.exists(path) and os.path.exists(drv)) else None


def _drive(mode, name, payload, out_ints):
    path = ref_lib(name)
    if path is None:
        raise RuntimeError("oracle/_ref/%s/run.so or oracle/ref_driver missing (needs /root/reference at build time)" % name)
    with tempfile.TemporaryDirectory() as d:
        fin, fout = os.path.join(d, "in.bin"), os.path.join(d, "out.bin")
        with open(fin, "wb") as f:
            f.write(payload)
        subprocess.run([os.path.join(HERE, "ref_driver"), mode, path, fin, fout], check=True)
        return np.fromfile(fout, dtype=np.int32, count=out_ints)


def ref_run_tet_builder(name, tet_tx4, n_point, out_rows, out_cols):
    """Call `run(int* tet, int* out, int* n_out, int n_point, int n_tet)` of utils/lib/<name>/run.cpp."""
    tet = np.ascontiguousarray(tet_tx4, dtype=np.int32)
    cap = int(out_rows) * int(out_cols)
    hdr = np.array([n_point, tet.shape[0], cap], dtype=np.int32)
    res = _drive("tet", name, hdr.tobytes() + tet.tobytes(), 1 + cap)
    return res[1:].reshape(out_rows, out_cols), int(res[0])


def ref_colaps_v(points_nx3):
    pts = _f32(points_nx3)
    n = pts.shape[0]
    res = _drive("colaps", "colaps_v", np.array([n], dtype=np.int32).tobytes() + pts.tobytes(), 1 + 2 * n)
    cnt = int(res[0])
    return res[1:1 + n].copy(), res[1 + n:1 + n + cnt].copy()


# ---- Kaolin stand-ins (parity unpinned, see render_oracle.c) ----------------------------------------------------
def sparse_render(pixel_coords, render_ranges, face_z, face_xy, face_feat, K, eps=1e-8, threads=None):
    pix, rng, fz, fxy, ff = _f32(pixel_coords), _f32(render_ranges), _f32(face_z), _f32(face_xy), _f32(face_feat)
    B, P = pix.shape[0], pix.shape[1]
    F, D = fz.shape[1], ff.shape[-1]
    out = np.zeros((B, P, K, D), dtype=np.float32)
    idx = np.full((B, P, K), -1, dtype=np.int64)
    L = lib()
    _split(lambda a, b: L.orc_sparse_render(_p(pix), _p(rng), _p(fz), _p(fxy), _p(ff), B, P, F, D, K, C.c_float(eps), _p(out), _p(idx),
                                            C.c_longlong(a), C.c_longlong(b)), B * P, threads, min_chunk=8)
    return out, idx


def check_sign(verts, faces, points, threads=None):
    v, p = _f32(verts), _f32(points)
    f = np.ascontiguousarray(faces, dtype=np.int32)
    B, n, m, np_ = v.shape[0], v.shape[1], f.shape[0], p.shape[1]
    out = np.zeros((B, np_), dtype=np.uint8)
    L = lib()
    _split(lambda a, b: L.orc_check_sign(_p(v), _p(f), _p(p), B, n, m, np_, _p(out), C.c_longlong(a), C.c_longlong(b)), B * np_, threads,
           min_chunk=16)
    return out.astype(bool)


# ---- the reference's own __host__ __device__ kernel helpers compiled for the host (oracle/build_ref_kernels.sh) ----
def ref_kernel_lib(name):
    path = os.path.join(HERE, "_ref", "kernels", name + ".so")
    return C.CDLL(path) if os.path.exists(path) else None


def ref_point_in_tet(tet_bxfx4x3, pts_bxnx3):
    L = ref_kernel_lib("point_in_tet")
    tet, pts = _f32(tet_bxfx4x3), _f32(pts_bxnx3)
    B, T, P = tet.shape[0], tet.shape[1], pts.shape[1]
    out = np.full((B, P, 1), -1.0, dtype=np.float32)
    L.ref_point_in_tet(_p(tet), _p(pts), _p(out), B, P, T)
    return out


def ref_point_face_distance(pts, faces):
    L = ref_kernel_lib("face_distance_fwd")
    pts, faces = _f32(pts), _f32(faces)
    B, P, F = pts.shape[0], pts.shape[1], faces.shape[1]
    nf = _f32(np.full(B, F))
    d = np.zeros((B, P, 1), dtype=np.float32)
    f = np.zeros((B, P, 1), dtype=np.float32)
    L.ref_point_face_distance(_p(pts), _p(faces), _p(nf), _p(d), _p(f), B, P, F)
    return d, f


def ref_point_face_distance_bwd(pts, faces, closest_f, dl_dd):
    L = ref_kernel_lib("face_distance_bwd")
    pts, faces, cf, g = _f32(pts), _f32(faces), _f32(closest_f), _f32(dl_dd)
    B, P, F = pts.shape[0], pts.shape[1], faces.shape[1]
    out = np.zeros((B, F, 3, 3), dtype=np.float32)
    L.ref_point_face_distance_bwd(_p(pts), _p(faces), _p(cf), _p(g), _p(out), B, P, F)
    return out


def ref_face_adjacency(face_fx3x3, n_max_nei=30):
    L = ref_kernel_lib("face_adj")
    face = _f32(face_fx3x3)
    F = face.shape[0]
    adj = np.full((F, n_max_nei), -1.0, dtype=np.float32)
    L.ref_face_adjacency(_p(face), _p(adj), F, n_max_nei)
    return adj

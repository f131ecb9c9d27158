"""The reference's own CUDA kernels on the GPU (TEST / BENCH INFRASTRUCTURE ONLY, see oracle/__init__.py).

oracle/build_ref_kernels.sh compiles the unmodified `__global__` kernels of the reference's four extensions and of
layers/nearest_neighbor for sm_100a into oracle/_ref/kernels_cuda/*.so (built in the authoring container where
/root/reference is mounted; the .so files travel to the GPU box, the sources do not).  The launchers take raw device
pointers, use the reference's launch geometry and the legacy default stream, exactly like the reference
(check_condition_tet_for.cu:199-207, tet_analytic_distance_for.cu:316-324, _back.cu:695-705, tet_face_adj_m_for.cu:114-122,
nearest_neighbor_cuda.cu:57-79).  Used for (i) GPU-side parity of deftet_b200 against the reference's real device code and
(ii) the "reference CUDA" column of the measurements.
"""
import ctypes as C
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}


def available():
    return all(os.path.exists(os.path.join(HERE, "_ref", "kernels_cuda", n + ".so"))
               for n in ("point_in_tet", "face_distance_fwd", "face_distance_bwd", "face_adj", "nearest_neighbor"))


def _lib(name):
    if name not in _LIBS:
        path = os.path.join(HERE, "_ref", "kernels_cuda", name + ".so")
        if not os.path.exists(path):
            raise RuntimeError("oracle/_ref/kernels_cuda/%s.so missing: run `make -C oracle` where /root/reference is mounted" % name)
        _LIBS[name] = C.CDLL(path)
    return _LIBS[name]


def _p(t):
    assert t.is_cuda and t.is_contiguous()
    return C.c_void_p(t.data_ptr())


def _check(rc, what):
    if rc != 0:
        raise RuntimeError("reference CUDA kernel %s: launch failed with cudaError %d" % (what, rc))


def _sync_in():
    # the reference launches on the legacy default stream; make prior work on torch's current stream visible to it
    torch.cuda.current_stream().synchronize()


def point_in_tet(tet_bxfx4x3, pts_bxnx3):
    """check_condition_f_base forward: float tet id per point, -1 = outside (utils.py:33-53 of the extension)."""
    tet, pts = tet_bxfx4x3.float().contiguous(), pts_bxnx3.float().contiguous()
    B, T, P = tet.shape[0], tet.shape[1], pts.shape[1]
    out = torch.zeros(B, P, 1, device=tet.device, dtype=torch.float32)
    _sync_in()
    _check(_lib("point_in_tet").refcuda_point_in_tet(_p(tet), _p(pts), _p(out), B, P, T), "point_in_tet")
    torch.cuda.synchronize()
    return out


def nearest_neighbor(queries, points):
    q, p = queries.float().contiguous(), points.float().contiguous()
    B, Q, M = q.shape[0], q.shape[1], p.shape[1]
    out = torch.zeros(B, Q, device=q.device, dtype=torch.int32)
    _sync_in()
    _check(_lib("nearest_neighbor").refcuda_nearest_neighbor(_p(q), _p(p), _p(out), B, Q, M), "nearest_neighbor")
    torch.cuda.synchronize()
    return out


def point_face_distance(pts, faces_bxfx3x3, n_face_b=None):
    pts, faces = pts.float().contiguous(), faces_bxfx3x3.float().contiguous()
    B, P, F = pts.shape[0], pts.shape[1], faces.shape[1]
    nf = (torch.full((B,), float(F), device=pts.device) if n_face_b is None else n_face_b.float().contiguous())
    d = torch.zeros(B, P, 1, device=pts.device)
    f = torch.zeros(B, P, 1, device=pts.device)
    _sync_in()
    _check(_lib("face_distance_fwd").refcuda_point_face_distance(_p(pts), _p(faces), _p(f), _p(d), _p(nf), B, P, F),
           "point_face_distance")
    torch.cuda.synchronize()
    return d, f


def point_face_distance_bwd(pts, faces_bxfx3x3, closest_f, dl_dd):
    pts, faces = pts.float().contiguous(), faces_bxfx3x3.float().contiguous()
    cf, g = closest_f.float().contiguous(), dl_dd.float().contiguous()
    B, P, F = pts.shape[0], pts.shape[1], faces.shape[1]
    out = torch.zeros(B, F, 3, 3, device=pts.device)
    _sync_in()
    _check(_lib("face_distance_bwd").refcuda_point_face_distance_bwd(_p(pts), _p(faces), _p(cf), _p(g), _p(out), B, P, F),
           "point_face_distance_bwd")
    torch.cuda.synchronize()
    return out


def face_adjacency(face_fx3x3, n_max_nei=30):
    face = face_fx3x3.float().contiguous()
    F = face.shape[0]
    adj = torch.full((F, n_max_nei), -1.0, device=face.device)
    _sync_in()
    _check(_lib("face_adj").refcuda_face_adjacency(_p(face), _p(adj), F, n_max_nei), "face_adjacency")
    torch.cuda.synchronize()
    return adj


def time_ms(fn, reps=3):
    """Median wall time of fn() in ms; fn must end synchronised (every wrapper above does)."""
    import time
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]

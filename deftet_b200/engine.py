"""The geometry hot path as one object: topology built once on the GPU, then a batched forward (+ autograd
backward) of every per-tetrahedron / per-surface loss of ``DefTet.forward_surface_align``
(reference layers/DefTet/deftet.py:51-130) without its per-sample Python loop, host syncs or the
materialised (B,T,4,3) gather.

    eng = GeometryEngine(init_pos, tets)                       # A10/A11 builders + rest inverses
    out = eng.losses(pos, occ, gt_points, u, v, query_points)  # dict of (B,) losses, differentiable in pos

This is the call ``bench.py`` times; ``deftet_b200.deftet.DefTet`` is the drop-in module with the
reference's method names built on the same kernels.
"""
from __future__ import annotations

import torch

from . import builders, energies, search, surface


class GeometryEngine:
    def __init__(self, init_pos, tets, max_boundary_faces=16384, samples_per_face=20, device=None):
        device = torch.device(device) if device is not None else (init_pos.device if torch.is_tensor(init_pos) else torch.device("cuda"))
        self.device = device
        self.init_pos = torch.as_tensor(init_pos, dtype=torch.float32).to(device).contiguous()
        self.tet = torch.as_tensor(tets).to(device).to(torch.int32).contiguous()
        self.n_vert = self.init_pos.shape[0]
        self.n_tet = self.tet.shape[0]
        self.inverse_v = energies.tet_inverse_v(self.init_pos, self.tet)                       # train_multigpu.py:105-110
        self._tet_tiles = None
        f3, ft2, fs2, bnd = builders.tet_to_face(self.n_vert, self.tet)                         # train_multigpu.py:77-82
        self.tet_face_fx3, self.tet_face_tetidx_fx2, self.tet_face_slot_fx2, self.cube_boundary = f3, ft2, fs2, bnd
        self.face_table = surface.FaceTable(f3, ft2)
        self.max_boundary_faces = int(max_boundary_faces)
        self.samples_per_face = int(samples_per_face)
        self._streams = None
        self._ovf_host, self._ovf_event = None, None            # lazily checked capacity flag of the previous call (see losses)

    @property
    def tet_tiles(self):
        """Tile-local topology for the opt-in tiled energy kernels (built on first use)."""
        if self._tet_tiles is None:
            self._tet_tiles = energies.TetTiles(self.tet, self.n_vert)
        return self._tet_tiles

    def vertex_adjacency(self, normalize=True):
        """Row-normalised vertex adjacency (train_multigpu.py:72-75) as a torch sparse tensor."""
        return builders.tet_to_adj_sparse(self.n_vert, self.tet, normalize)

    def _side_streams(self):
        if self._streams is None:
            self._streams = [torch.cuda.Stream(device=self.device) for _ in range(4)]
        return self._streams

    def _watch_overflow(self, overflow):
        """The boundary faces are compacted into a fixed-capacity buffer (no host synchronisation, graph capturable); when a sample has
        more than `max_boundary_faces` of them the surface losses would silently be computed on a truncated surface (the reference's
        get_boundary_index never truncates).  The flag travels to pinned host memory asynchronously and is looked at -- without
        blocking -- at the start of a later call."""
        if torch.cuda.is_current_stream_capturing():
            return
        if self._ovf_host is None:
            self._ovf_host = torch.zeros(1, dtype=overflow.dtype).pin_memory()
            self._ovf_event = torch.cuda.Event()
        elif not self._ovf_event.query():
            return                                            # the previous flag is still in flight: keep watching that one
        self._ovf_host.copy_(overflow.reshape(1), non_blocking=True)
        self._ovf_event.record(torch.cuda.current_stream(self.device))

    def _raise_on_earlier_overflow(self):
        if torch.cuda.is_current_stream_capturing():           # (an event query would invalidate the capture)
            return
        if self._ovf_event is not None and self._ovf_event.query() and int(self._ovf_host[0]) != 0:
            self._ovf_host.zero_()
            raise RuntimeError("GeometryEngine.losses: an earlier call found more boundary faces than max_boundary_faces=%d; its surface "
                               "losses were computed on a truncated surface -- construct the engine with a larger capacity "
                               "(at most face_table.n_face = %d)" % (self.max_boundary_faces, self.face_table.n_face))

    def losses(self, pos, occ, gt_points, u, v, query_points=None, want=("energies", "chamfer", "distance", "normal", "occupancy"),
               concurrent=True, chamfer_targets=None):
        """pos (B,V,3); occ (B,T) in {0,1}; gt_points (B,S,3); u,v (B,Fmax,samples) sampling randoms;
        query_points (B,P,3) for the point-in-tet query; chamfer_targets (B,M,3): the 1-NN target set of the chamfer term when it is
        not gt_points (a rank that holds only a slice of the GT points for the distance term).  Returns a dict; every loss is (B,).

        The loss groups are independent of each other (they only share `pos` and the boundary faces) and each of their
        kernels leaves most issue slots idle (profiles/), so with ``concurrent=True`` they are enqueued on side streams
        (fork after the inputs, join before returning; autograd replays the same stream assignment in backward).
        Everything stays capturable in one CUDA graph."""
        self._raise_on_earlier_overflow()
        out = {}
        main = torch.cuda.current_stream(self.device)
        if concurrent:
            s_en, s_pit, s_ch, s_sd = self._side_streams()
        else:
            s_en = s_pit = s_ch = s_sd = main
        used = []

        def fork(s):
            if s is not main:
                s.wait_stream(main)
                used.append(s)
            return torch.cuda.stream(s)

        def publish(*tensors):
            if concurrent:
                for t in tensors:
                    t.record_stream(main)

        if "energies" in want:
            with fork(s_en):
                out["amips"], out["edge"], out["volume_variance"] = energies.tet_energies(pos, self.tet, self.inverse_v, tiles=self._tet_tiles)
                publish(out["amips"], out["edge"], out["volume_variance"])
        if "occupancy" in want and query_points is not None:
            with fork(s_pit):
                out["condition"], out["barycentric"] = search.point_in_tet(pos, self.tet, query_points)
                publish(out["condition"], out["barycentric"])
        if any(k in want for k in ("chamfer", "distance", "normal", "boundary")):
            faces, counts, overflow = surface.boundary_faces(self.face_table, occ, self.max_boundary_faces)
            out["boundary_faces"], out["boundary_counts"], out["boundary_overflow"] = faces, counts, overflow
            self._watch_overflow(overflow)
        if "chamfer" in want:
            with fork(s_ch):
                out["chamfer"] = surface.surface_chamfer(pos, faces, counts, u, v, gt_points if chamfer_targets is None else chamfer_targets)
                publish(out["chamfer"])
        if "distance" in want:
            with fork(s_sd):
                out["distance"] = surface.surface_distance(pos, faces, counts, gt_points)
                publish(out["distance"])
        if "normal" in want:
            out["normal"] = surface.surface_normal_loss(pos, faces, counts)
        for s in used:
            main.wait_stream(s)
        return out

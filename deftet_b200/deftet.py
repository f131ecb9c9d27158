"""Drop-in ``DefTet`` module: same method names, arguments and return structure as the reference class
(layers/DefTet/deftet.py:21-343), built on the sm_100a kernels.  Differences that are not observable through
the return values: no (B,T,4,3) gather is needed internally, the per-sample Python loop of
``forward_surface_align`` (:89-103) is one batched launch sequence, and ``inverse_v`` is not re-cloned per call."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import energies, render, search, surface

EPS = 1e-10


class DefTet(nn.Module):
    def __init__(self, device=None):
        super(DefTet, self).__init__()
        self.pow = 4
        self.device = device
        self.features_fixed = False
        self.z_window_radius = 0.025
        self.inverse_v = None
        self._face_table_key = None
        self._face_table = None

    # ---- occupancy label (deftet.py:33-49) ------------------------------------------------------------------
    def check_tet_inside_sdfs(self, tet_bxfx4x3, mesh_list):
        verts, faces = mesh_list[0], mesh_list[1]
        with torch.no_grad():
            occupancy = []
            for v, f, tet_fx4x3 in zip(verts, faces, tet_bxfx4x3):
                center = torch.mean(tet_fx4x3, dim=1)
                result = render.check_sign(v, f[0], center.unsqueeze(dim=0), hash_resolution=512)
                occupancy.append(result.unsqueeze(-1))
            occupancy = torch.cat(occupancy, dim=0).float()
        return occupancy

    def _table(self, tet_face_fx3, tet_idx_fx2):
        key = (tet_face_fx3.data_ptr(), tet_idx_fx2.data_ptr(), tet_face_fx3.shape[0])
        if key != self._face_table_key:
            self._face_table = surface.FaceTable(tet_face_fx3, tet_idx_fx2)
            self._face_table_key = key
        return self._face_table

    # ---- deftet.py:51-130 ---------------------------------------------------------------------------------------
    def forward_surface_align(self, vertice_pos, point_pos_bxpx3, tetrahedron_bxfx4=None, mesh_list=None, gt_surface_points=None,
                              tet_face_bxfx3=None, inference=False, pred_occ=None, tet_face_tet_bx4fx2=None, save=False, save_name=None,
                              inference_threshold=0.4):
        tetrahedron_bxfx4 = tetrahedron_bxfx4.long()
        tet32 = tetrahedron_bxfx4[0].to(torch.int32).contiguous()       # stride-0 batch view of one grid (train_multigpu.py:185-194)
        B = vertice_pos.shape[0]
        tet_bxfx4x3 = torch.gather(input=vertice_pos.unsqueeze(2).expand(-1, -1, 4, -1),
                                   index=tetrahedron_bxfx4.unsqueeze(-1).expand(-1, -1, -1, 3), dim=1) if mesh_list is not None else None
        center_occ = self.check_tet_inside_sdfs(tet_bxfx4x3, mesh_list)
        boundary = self.get_boundary_index(tet_face_bxfx3[0], tet_face_tet_bx4fx2[0], center_occ.squeeze(dim=-1))
        amips_energy, edge, volume_variance = energies.tet_energies(vertice_pos, tet32, self.inverse_v.to(vertice_pos.device))
        # batched surface stage (replaces the loop of :89-103); RNG calls in the reference's order
        counts = [int(b.shape[0]) for b in boundary]
        Fmax = max(max(counts), 1)
        faces = torch.zeros(B, Fmax, 3, device=vertice_pos.device, dtype=torch.int32)
        u = torch.zeros(B, Fmax, 20, device=vertice_pos.device)
        v = torch.zeros(B, Fmax, 20, device=vertice_pos.device)
        for i in range(B):
            if counts[i]:
                faces[i, :counts[i]] = boundary[i].to(torch.int32)
                u[i, :counts[i]] = torch.sqrt(torch.rand(size=(1, counts[i], 20, 1), device=vertice_pos.device))[0, :, :, 0]
                v[i, :counts[i]] = torch.rand(size=(1, counts[i], 20, 1), device=vertice_pos.device)[0, :, :, 0]
        cnt = torch.tensor(counts, device=vertice_pos.device, dtype=torch.int32)
        gt = gt_surface_points.reshape(B, -1, 3)
        sum_normal_loss = surface.surface_normal_loss(vertice_pos, faces, cnt).mean().reshape(1)
        sum_chamfer_distance = surface.surface_chamfer(vertice_pos, faces, cnt, u, v, gt).mean().reshape(1)
        sum_analytic_distance = surface.surface_distance(vertice_pos, faces, cnt, gt).mean().reshape(1)
        lap_v_loss = torch.zeros_like(sum_normal_loss)
        center_occ = center_occ.squeeze(-1)
        if inference:
            assert point_pos_bxpx3 is not None, 'point_pos_bxpx3 not given'
            condition, _ = search.point_in_tet(vertice_pos.detach(), tet32, point_pos_bxpx3)
            pred_occ = (pred_occ > inference_threshold).float()
            pred_surface_face = self.get_boundary_index(tet_face_bxfx3[0], tet_face_tet_bx4fx2[0], pred_occ)
            return (amips_energy, edge, volume_variance, sum_analytic_distance, sum_normal_loss, center_occ, condition, boundary,
                    pred_surface_face, sum_chamfer_distance)
        return (amips_energy, edge, volume_variance, sum_analytic_distance, sum_normal_loss, center_occ, boundary, sum_chamfer_distance,
                lap_v_loss)

    def paste_occ(self, pred_tet_occ, condition):
        return search.paste_occ(pred_tet_occ, condition)

    # ---- deftet.py:138-184 (one sample) ------------------------------------------------------------------------------
    def forward(self, v_pos_bxnx3=None, tet_bxfx4=None, boundary_bxfx3=None, gt_surface_point=None, inverse_offset=None, tet_bxfx4x3=None,
                calculate_amips_volume=True):
        if calculate_amips_volume:
            tet32 = tet_bxfx4[0].to(torch.int32).contiguous()
            inv = inverse_offset if inverse_offset is not None else None
            flags = energies.VOLUME | (energies.AMIPS if inv is not None else 0)
            am, _, area_variance = energies.tet_energies(v_pos_bxnx3, tet32, inv, flags)
            amips_energy = am if inv is not None else torch.zeros_like(area_variance)
            if tet_bxfx4x3 is None:
                tet_bxfx4x3 = torch.gather(input=v_pos_bxnx3.unsqueeze(2).expand(-1, -1, 4, -1),
                                           index=tet_bxfx4.long().unsqueeze(-1).expand(-1, -1, -1, 3), dim=1)
        if boundary_bxfx3.shape[1] == 0:
            one_loss = torch.ones(1, device=boundary_bxfx3.device)
            if calculate_amips_volume:
                return one_loss, one_loss, one_loss, area_variance, amips_energy, tet_bxfx4x3
            return one_loss, one_loss, one_loss
        B, F = boundary_bxfx3.shape[0], boundary_bxfx3.shape[1]
        dev = v_pos_bxnx3.device
        faces = boundary_bxfx3.to(torch.int32).contiguous()
        cnt = torch.full((B,), F, device=dev, dtype=torch.int32)
        normal_loss = surface.surface_normal_loss(v_pos_bxnx3, faces, cnt)
        u = torch.sqrt(torch.rand(size=(B, F, 20, 1), device=dev))[..., 0]
        v = torch.rand(size=(B, F, 20, 1), device=dev)[..., 0]
        gt = gt_surface_point.reshape(gt_surface_point.shape[0], -1, 3)
        chamfer_distance = surface.surface_chamfer(v_pos_bxnx3, faces, cnt, u, v, gt)
        analytic_distance = surface.surface_distance(v_pos_bxnx3, faces, cnt, gt)
        if calculate_amips_volume:
            return chamfer_distance, analytic_distance, normal_loss, area_variance, amips_energy, tet_bxfx4x3
        return chamfer_distance, analytic_distance, normal_loss

    # ---- deftet.py:186-203 ------------------------------------------------------------------------------------------------
    def get_boundary_index(self, tet_face_fx3, tet_idx_fx2, occ_bxn):
        table = self._table(tet_face_fx3, tet_idx_fx2)
        faces, counts, _ = surface.boundary_faces(table, occ_bxn, table.n_face)
        n = counts.tolist()
        return [faces[b, :n[b]].long() for b in range(len(n))]

    def get_internal_index(self, tet_face_fx3, tet_idx_fx2, occ_bxn):
        occ2 = torch.gather(input=occ_bxn, index=tet_idx_fx2.reshape(-1).unsqueeze(0).expand(occ_bxn.shape[0], -1), dim=1)
        s = occ2.reshape(occ_bxn.shape[0], -1, 2).sum(dim=-1)
        return [tet_face_fx3[t == 2] for t in s]

    # ---- energies (deftet.py:205-338) ---------------------------------------------------------------------------------------
    def my_inverse(self, T):
        det = torch.abs(torch.det(T)) < 1e-10
        detf = det.float().unsqueeze(-1).unsqueeze(-1)
        eye = torch.eye(T.shape[-1], dtype=torch.float, device=T.device).unsqueeze(0).expand(T.shape[0], -1, -1)
        return torch.inverse(T * (1 - detf) + eye * detf), 1 - det.float()

    def volume_variance(self, tet_bxfx4x3, base_area_mask=None, area_normalize=(20, 20), pow=2, center_occ=None):
        if pow != 4:
            raise NotImplementedError("the reference hard-codes pow = 4 (deftet.py:27,:82); other powers are not implemented")
        return energies.volume_variance_soup(tet_bxfx4x3)

    def amips_energy(self, tet_bxfx4x3, inverse_v, scale=20, center_occ=None, square=False):
        assert scale == 20, "the kernels implement the reference's scale of 20"
        e = energies.amips_energy_soup(tet_bxfx4x3, inverse_v)
        if square:
            raise NotImplementedError("square=True is never used by the reference callers")
        return e

    def tet_inverse_v(self, init_tet_pos, init_tet_fx4, scale=20):
        assert scale == 20
        return energies.tet_inverse_v(init_tet_pos.float(), init_tet_fx4)

    def edge_length(self, tet_bxfx4x3, pow=2):
        if pow != 4:
            raise NotImplementedError("the reference only calls edge_length with pow = 4 (deftet.py:105)")
        return energies.edge_length_soup(tet_bxfx4x3)

    def laplacian_sparse(self, offset, adj):
        adj = adj.coalesce()
        idx = adj.indices()
        return render.laplacian_loss(offset, idx.t().contiguous(), adj.values())

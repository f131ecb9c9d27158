"""Host-side mirror of the diff_render model (diff_render/diftet_6_subdiv/3_model/deftet.py ``Deftet``, 5_rendereq/deftetrneder.py
``rendermeshcolor`` / ``peel2mask`` / ``preprocess_save``, 3_model/cameraop.py ``perspective``) on top of the sm_100a kernels.

Same class / method names, argument meaning and return tuples as the reference, so that
6_optim/optim_with_mask_subdiv_from_gridmov.py drives it unchanged.  Differences, all internal:
  * topology lives on the GPU: ``updategeometry`` / ``subdivision`` / ``deletetet`` run the builders of ``deftet_b200.topology``
    (sort / scan / compact) instead of Python dict loops, a dense (P,P) adjacency matrix and an O(E*T) edge matching loop;
  * ``forward`` with this module's ``rendermeshcolor`` takes the fused route: one projection+gather kernel
    (``topology.project_faces``) feeding the fused rasterizer+compositor (``render.render_composite``), so neither the (B,P,*)
    per-vertex intermediates nor the (B,P,K,d) layer tensor are ever written; any other ``renderfunc`` gets the reference's
    per-vertex tensors;
  * ``saveobj`` (mesh export for visualisation) writes the reference's `tet-geo-*` / `tet-color-*` OBJ files from host-side array code.
"""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.nn as nn

from . import render, topology
from .grid import acute_lattice_grid, read_tet_file, snap_boundary


def perspective(points_bxpx3, cameras):
    """3_model/cameraop.py:14-33 (generic per-vertex form, torch ops)."""
    camera_rot_bx3x3, camera_pos_bx3, camera_proj_3x1 = cameras
    cam = torch.matmul(points_bxpx3 - camera_pos_bx3.view(-1, 1, 3), camera_rot_bx3x3.permute(0, 2, 1))
    xy = cam * camera_proj_3x1.view(-1, 1, 3)
    return cam, xy[:, :, :2] / xy[:, :, 2:3]


def preprocess_save(tfpoints, tfpfeat_bxpxd):
    """5_rendereq/deftetrneder.py:26-28."""
    featact = torch.sigmoid(tfpfeat_bxpxd)
    return featact[:, :1], featact[:, 1:]


def peel2mask(ims_bxpxkxd, imdepth_bxpxkx1=None):
    """5_rendereq/deftetrneder.py:31-64 (dense form, for callers that hold the (B,P,K,d) tensor of ``deftet_sparse_render``)."""
    eps = 1e-10
    alpha = torch.clamp(ims_bxpxkxd[:, :, :, :1], eps, 1.0 - eps)
    shift = nn.functional.pad(1 - alpha[:, :, :-1, :], pad=(0, 0, 1, 0), mode="constant", value=1)
    xvis = alpha * torch.cumprod(shift, dim=2)
    xcolor = (ims_bxpxkxd[:, :, :, 1:] * xvis).sum(dim=2)
    dep = (imdepth_bxpxkx1 * xvis).sum(dim=2) if imdepth_bxpxkx1 is not None else None
    xvis = xvis.sum(2)
    xcolor = xcolor + (1.0 - xvis)
    if dep is not None:
        dep = dep + -6.0 * (1.0 - xvis)
    return xcolor, xvis, dep


def _composite(xy, xydep, fz, fxy, ff, depth, knum):
    color, mask = render.render_composite(xy, xydep, fz, fxy, ff, knum=knum)
    dep = None
    if depth:       # depth rode along as the last colour channel: sum d*vis + (1 - vis)  ->  sum d*vis - 6 (1 - vis)
        dep = color[..., -1:] - 7.0 * (1.0 - mask)
        color = color[..., :-1]
    return color, mask, dep


def rendermeshcolor(xy_1xpx2, xydep_1xpx2, points3d_bxpx3, points2d_bxpx2, tfpfeat_bxpxd, faces_fx3, viewdir=False, depth=False,
                    istraining=False, knum=300):
    """5_rendereq/deftetrneder.py:67-113: sigmoid -> per-face gather -> tet-face rasterizer -> front-to-back compositing, with the last
    two stages fused (``render.render_composite``).  -> (imcolor (B,P,3), immask (B,P,1), imdepth (B,P,1) or None)."""
    assert not viewdir
    if depth:
        tfdepth = tfpfeat_bxpxd[:, :, :1]
        tfpfeat_bxpxd = tfpfeat_bxpxd[:, :, 1:]
    feat = torch.sigmoid(tfpfeat_bxpxd)
    if depth:
        feat = torch.cat([feat, tfdepth], dim=2)          # opacity stays channel 0
    B, F = points3d_bxpx3.shape[0], faces_fx3.shape[0]
    f = faces_fx3.to(points3d_bxpx3.device).long().reshape(-1)
    fz = points3d_bxpx3[:, f, 2].reshape(B, F, 3)
    fxy = points2d_bxpx2[:, f].reshape(B, F, 3, 2)
    ff = feat[:, f].reshape(B, F, 3, -1)
    return _composite(xy_1xpx2, xydep_1xpx2, fz, fxy, ff, depth, knum)


rendermeshcolor.deftet_b200_fused = True


class Deftet(nn.Module):
    """3_model/deftet.py:27-557."""

    def __init__(self, basefolder, res, coef, feature_dim=4, feature_raw=True, feature_fixed_dim=0, feature_fixed_init=None,
                 neighbourlayer=3, device=None, seed=None):
        super().__init__()
        self.res = res
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        file_name = "%s/cube_%d_tet.tet" % (basefolder, res)
        if os.path.exists(file_name):                     # the grids the reference ships (read_tetrahedron(file_name, res=0.02))
            points_px3, tet_list_tx4 = read_tet_file(file_name)
        else:                                             # no QuarTet offline: generate the same acute lattice in-process
            g = acute_lattice_grid(int(res))
            points_px3, tet_list_tx4 = g.vertices, g.tets
        points_px3, _ = snap_boundary(points_px3, 0.02)
        self.coef = coef
        self.neilayer = neighbourlayer
        p = points_px3.astype(np.float32)
        p = p - (p.max(axis=0, keepdims=True) + p.min(axis=0, keepdims=True)) / 2
        assert p.max() <= 0.5 and p.min() >= -0.5
        self.feature_dim = feature_dim
        rng = np.random if seed is None else np.random.RandomState(seed)
        pointfeat_pxd = rng.rand(p.shape[0], feature_dim).astype(np.float32)
        if not feature_raw:
            pointfeat_pxd = pointfeat_pxd * 2 - 1
        self.features_fixed = feature_fixed_dim > 0
        self.features_fixed_dim = feature_fixed_dim
        if self.features_fixed and feature_fixed_dim == 3 and feature_fixed_init is None:
            feature_fixed_init = p * 2 * 0.95
        self.updatevaribale(p, pointfeat_pxd, pointmov_px3=None, pointfeat_fixed_pxd=feature_fixed_init)
        self.updategeometry(tet_list_tx4)

    # ------------------------------------------------------------------------------------------------------------ state
    def _t(self, a, dtype=torch.float32):
        if isinstance(a, np.ndarray):
            a = torch.from_numpy(np.ascontiguousarray(a))
        return a.detach().to(device=self.device, dtype=dtype)

    def updatevaribale(self, points_px3, pointfeat_pxd, pointmov_px3=None, pointfeat_fixed_pxd=None):
        self.tfpoint_px3 = self._t(points_px3)
        mov = torch.zeros_like(self.tfpoint_px3) if pointmov_px3 is None else self._t(pointmov_px3)
        self.tfpointmov_px3 = nn.Parameter(mov.clone())
        self.tfpointfeat_pxd = nn.Parameter(self._t(pointfeat_pxd).clone())
        self.tfpointfeat_fixed_pxd = self._t(pointfeat_fixed_pxd) if self.features_fixed else None
        self.n_point = self.tfpoint_px3.shape[0]

    def updategeometry(self, tet_list_tx4):
        n_point = self.n_point
        tet = self._t(tet_list_tx4, torch.int32).contiguous()
        faces_fx3, face_tet_idx_fx2, _ = topology.tet_to_face_idx(n_point, tet, with_boundary=True)
        self.tff_fx3 = faces_fx3.long()
        self._faces32 = faces_fx3.contiguous()
        self.tftet_tx4 = tet.long()
        self._tet32 = tet
        self.tftet2face_fx2 = face_tet_idx_fx2.long()
        self.tet_neighbour_idx = topology.tet_neighbours(tet, n_point)
        table, adjsum = topology.generate_point_adj_idx(n_point, tet)
        self._adj_table = table                                            # neighbour ids, -1 padding
        self.tfpoint_adj_idx_pxm = table.long() + 1                        # the reference's +1-shifted copy (deftet.py:161)
        self.tfpoint_adj_weights_px1 = adjsum + 1e-10

    def sethw(self, height, width, multiplier):
        xidx = (torch.arange(width, dtype=torch.float32) + 0.5) / width * 2.0 - 1.0
        yidx = -((torch.arange(height, dtype=torch.float32) + 0.5) / height * 2.0 - 1.0)
        ymap, xmap = torch.meshgrid(yidx, xidx, indexing="ij")
        self.height, self.width, self.multiplier = height, width, multiplier
        self.xy_px2 = torch.stack([xmap, ymap], dim=2).view(-1, 2).to(self.device)

    def todev(self, dev):
        dev = torch.device(dev)
        self.device = dev
        for name in ("tfpoint_px3", "tfpoint_adj_weights_px1", "xy_px2", "tfpointfeat_fixed_pxd", "tff_fx3", "_faces32", "tftet_tx4", "_tet32",
                     "tftet2face_fx2", "tet_neighbour_idx", "_adj_table", "tfpoint_adj_idx_pxm"):
            t = getattr(self, name, None)
            if t is not None:
                setattr(self, name, t.to(dev))

    def get_hw(self):
        return self.height, self.width

    def get_point(self, with_coef=False):
        p = self.tfpoint_px3 + self.tfpointmov_px3
        return self.coef * p if with_coef else p

    def get_mov(self):
        return self.tfpointmov_px3

    def get_feat(self):
        tffeat = self.tfpointfeat_pxd
        if self.features_fixed:
            tffeat = torch.cat([tffeat, self.tfpointfeat_fixed_pxd], dim=1)
        return tffeat

    # ------------------------------------------------------------------------------------------------------------ regularisers
    def get_featlap(self, pointfeat_px3):
        return topology.featlap(pointfeat_px3, self._adj_table, self.tfpoint_adj_weights_px1)

    def get_volume_variance(self, base_area_mask=None, area_normalize=(20, 20), pow=2, center_occ=None):
        return topology.volume_deviation(self.get_point(), self._tet32, scale=2.0)

    # ------------------------------------------------------------------------------------------------------------ topology edits
    def deletetet(self, thres, processfunc):
        with torch.no_grad():
            tfweights_px1, _ = processfunc(self.get_point(with_coef=True), self.get_feat())
            kept, _ = topology.delete_tet_by_weight(self._tet32, tfweights_px1, self.tet_neighbour_idx, thres, self.neilayer)
        self.updategeometry(kept)

    def tensor2ndarray(self, processfunc=None):
        w = None
        if processfunc is not None:
            w = processfunc(self.get_point(with_coef=True), self.get_feat())[0].detach().cpu().numpy()
        return (self.tfpoint_px3.detach().cpu().numpy(), self.tfpointmov_px3.detach().cpu().numpy(), self.get_feat().detach().cpu().numpy(),
                self.tftet_tx4.detach().cpu().numpy(), w)

    def subdivision(self, loadpth=None, thres=None, processfunc=None):
        if loadpth is not None:
            self.load_state_dict(torch.load("%s/deftet.pth" % (loadpth,)))
        with torch.no_grad():
            tet_subdiv = None
            if thres is not None:
                assert processfunc is not None
                w = processfunc(self.get_point(with_coef=True), self.get_feat())[0][:, 0]
                tet_subdiv = w[self.tftet_tx4].min(dim=1).values < thres
            feat_mov = torch.cat([self.get_feat().detach(), self.tfpointmov_px3.detach()], dim=1)
            new_p, new_f, new_t = topology.generate_subdivision(self._tet32, self.tfpoint_px3, feat_mov, tet_subdiv)
        pmov, pfeat = new_f[:, -3:], new_f[:, :-3]
        pfeat_fixed = None
        if self.features_fixed:
            pfeat_fixed = pfeat[:, -self.features_fixed_dim:]
            pfeat = pfeat[:, :-self.features_fixed_dim]
        self.updatevaribale(new_p, pfeat, pmov, pfeat_fixed)
        self.updategeometry(new_t)

    # ------------------------------------------------------------------------------------------------------------ render
    def forward(self, impixsample_hxw, camrot_bx3x3, camtrans_bx3, camproj_3x1, renderfunc, viewpoint=False, depth=False, istraining=False,
                knum=300):
        bs = camrot_bx3x3.shape[0]
        xy_px2 = self.xy_px2[impixsample_hxw.view(-1)]
        xy_bxpx2 = xy_px2.unsqueeze(0).repeat(bs, 1, 1)
        xydep_bxpx2 = torch.zeros_like(xy_bxpx2)
        xydep_bxpx2[:, :, 0] = -1000
        fused = getattr(renderfunc, "deftet_b200_fused", False) and not viewpoint
        if fused:
            fz, fxy, ff = topology.project_faces(self.get_point(True), self.get_feat(), self._faces32, camrot_bx3x3, camtrans_bx3.reshape(bs, 3),
                                                 camproj_3x1, multiplier=self.multiplier, sigmoid=True)
            if depth:
                ff = torch.cat([ff, fz.unsqueeze(-1)], dim=-1)
            col, mask, dep = _composite(xy_bxpx2 * self.multiplier, xydep_bxpx2, fz, fxy, ff, depth, knum)
        else:
            tfp_bxpx3 = self.get_point(True).unsqueeze(0).repeat(bs, 1, 1)
            tfpfeat_bxpxd = self.get_feat().unsqueeze(0).repeat(bs, 1, 1)
            cams = [camrot_bx3x3, camtrans_bx3, camproj_3x1]
            vertices_camera_bxpx3, vertices_image_bxpx2 = perspective(tfp_bxpx3, cams)
            if viewpoint:
                viewdir = cams[1].view(-1, 1, 3) - tfp_bxpx3
                viewdir = viewdir / (torch.sqrt((viewdir ** 2).sum(dim=2, keepdim=True)) + 1e-10)
                tfpfeat_bxpxd = torch.cat([viewdir, tfpfeat_bxpxd], dim=2)
            if depth:
                tfpfeat_bxpxd = torch.cat([vertices_camera_bxpx3[:, :, 2:3], tfpfeat_bxpxd], dim=2)
            col, mask, dep = renderfunc(xy_bxpx2 * self.multiplier, xydep_bxpx2, vertices_camera_bxpx3, vertices_image_bxpx2 * self.multiplier,
                                        tfpfeat_bxpxd, self.tff_fx3, viewdir=viewpoint, depth=depth, istraining=istraining)
        return (col, mask, dep) if depth else (col, mask)

    # ------------------------------------------------------------------------------------------------------------ checkpoint
    def state_dict(self, destination=None, prefix="", keep_vars=False):
        models = nn.Module.state_dict(self, destination=destination, prefix=prefix, keep_vars=keep_vars)
        models["points"] = self.tfpoint_px3
        models["tets"] = self.tftet_tx4
        models["feat_fixed"] = self.tfpointfeat_fixed_pxd
        return models

    def load_state_dict(self, state_dict, strict=True):
        m = state_dict
        fixed = m["feat_fixed"].detach() if m.get("feat_fixed") is not None else None
        self.updatevaribale(m["points"].detach(), m["tfpointfeat_pxd"].detach(), m["tfpointmov_px3"].detach(), fixed)
        self.updategeometry(m["tets"].detach())

    def saveobj(self, savedir, prefix, processfunc):
        """Mesh export of 3_model/deftet.py:503-557: for thresholds 0.005 / 0.05 / 0.15 / 0.25 the tet faces that separate a tet with
        occupancy > 2*thres from a neighbour (or the outside, occupancy 0) whose occupancy differs by more than thres
        (``get_face_use_occ``, utils_tetsv.py:77-128), as `tet-geo-*.obj` (``save_tet_face``: 3 vertices per triangle, winding 1-3-2)
        and `tet-color-*.obj` (per-vertex BGR->RGB colours appended).  Tet occupancy = max of its vertex weights.  Not hot-path work:
        the neighbour occupancies come from the GPU A12 builder, the rest is host-side array code and text output."""
        from . import builders
        with torch.no_grad():
            pts = self.get_point(True)
            w, col = processfunc(pts, self.get_feat())
            tet = self._tet32.long()
            T = tet.shape[0]
            occ = w.reshape(-1)[tet].max(dim=1).values                                   # (T,)
            nocc = torch.zeros(T, 4, device=tet.device, dtype=occ.dtype)                   # neighbour occupancy per local face slot
            rows = builders.tet_adj_share(self._tet32, pts.shape[0]).long()                # (2 n_shared, 3): tet, neighbour, slot
            if rows.numel():
                nocc[rows[:, 0], rows[:, 2]] = occ[rows[:, 1]]
            corner = torch.tensor([[0, 1, 2], [1, 0, 3], [2, 3, 0], [3, 2, 1]], device=tet.device)      # slot -> (a, b, c) of the face
            tri_v = tet[:, corner]                                                         # (T,4,3) vertex ids
            pts_c, col_c = pts.detach().cpu().numpy(), col.detach().cpu().numpy()[:, ::-1]
            for thres in (0.005, 0.05, 0.15, 0.25):
                keep = ((nocc - occ.unsqueeze(1)).abs() > thres) & (occ.unsqueeze(1) > 2 * thres)
                # the reference concatenates the four slot masks over all tets in (tet, slot) order
                ids = tri_v[keep].cpu().numpy()                                            # (n,3)
                n = ids.shape[0]
                fidx = np.arange(n) * 3
                flines = np.stack([fidx + 1, fidx + 3, fidx + 2], axis=1)
                xyz = pts_c[ids.reshape(-1)]
                with open("%s/tet-geo-%s-thres-%.3f.obj" % (savedir, prefix, thres), "w") as f:
                    out = []
                    for k in range(n):
                        out.extend("v %f %f %f\n" % tuple(xyz[3 * k + i]) for i in range(3))
                        out.append("f %d %d %d\n" % tuple(flines[k]))
                    f.write("".join(out))
                rgb = col_c[ids.reshape(-1)]
                with open("%s/tet-color-%s-thres-%.3f.obj" % (savedir, prefix, thres), "w") as f:
                    out = []
                    for k in range(n):
                        out.extend("v %f %f %f %f %f %f\n" % (tuple(xyz[3 * k + i]) + tuple(rgb[3 * k + i])) for i in range(3))
                        out.append("f %d %d %d\n" % tuple(flines[k]))
                    f.write("".join(out))

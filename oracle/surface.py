"""Oracle for the predicted-surface stage (A9, A3, A4 and A5 consumers): torch (CPU) restatements of the
reference glue around the brute-force C kernels of oracle/native.py.  TEST INFRASTRUCTURE ONLY.

  get_boundary_index       layers/DefTet/deftet.py:186-195
  sample_surf_point_batch  utils/mesh_utils.py:290-299 (u, v supplied so that tests control the randomness)
  point_point_distance     utils/mesh_utils.py:360-366
  point_mesh_distance      utils/mesh_utils.py:368-374 (+ VarianceFunc backward, tet_analytic_distance_batch/utils.py:63-79)
  get_surface_normal_loss  utils/mesh_utils.py:16-39, get_normal :42-53
"""
import torch

from . import native


def get_boundary_index(tet_face_fx3, tet_idx_fx2, occ_bxn):
    occ2 = torch.gather(occ_bxn, 1, tet_idx_fx2.reshape(-1).unsqueeze(0).expand(occ_bxn.shape[0], -1)).reshape(occ_bxn.shape[0], -1, 2)
    s = occ2.sum(dim=-1)
    out = []
    for b in range(occ_bxn.shape[0]):
        sel = s[b] == 1
        faces = tet_face_fx3[sel]
        flip = (occ2[b][sel][:, 0] == 1).unsqueeze(-1)
        out.append(torch.where(flip, faces.flip(dims=[1]), faces))
    return out


def gather_faces(pos_1xvx3, faces_fx3):
    return pos_1xvx3[:, faces_fx3.long().reshape(-1)].reshape(pos_1xvx3.shape[0], -1, 3, 3)


def sample_points(surface_pos_bxfx3x3, u, v):
    a, b, c = surface_pos_bxfx3x3[:, :, 0:1], surface_pos_bxfx3x3[:, :, 1:2], surface_pos_bxfx3x3[:, :, 2:3]
    return (1 - u) * a + (u * (1 - v)) * b + u * v * c


def chamfer(pred_bxqx3, gt_bxmx3):
    idx = torch.from_numpy(native.nearest_neighbor(pred_bxqx3.detach().numpy(), gt_bxmx3.numpy()))
    closest = torch.gather(gt_bxmx3, 1, idx.unsqueeze(-1).expand(-1, -1, 3))
    return torch.sqrt(torch.sum((pred_bxqx3 - closest) ** 2, dim=-1) + 1e-10)


class _AnalyticDistance(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pts, faces):
        d, f = native.point_face_distance(pts.numpy(), faces.detach().numpy())
        ctx.save_for_backward(pts, faces.detach(), torch.from_numpy(f))
        return torch.from_numpy(d), torch.from_numpy(f)

    @staticmethod
    def backward(ctx, gd, gf):
        pts, faces, f = ctx.saved_tensors
        g = native.point_face_distance_bwd(pts.numpy(), faces.numpy(), f.numpy(), gd.contiguous().numpy())
        return None, torch.from_numpy(g)


def point_mesh_distance(gt_bxsx3, faces_bxfx3x3):
    d, _ = _AnalyticDistance.apply(gt_bxsx3, faces_bxfx3x3)
    return torch.sqrt(d + 1e-10)


def get_normal(a, b, c):
    u, v = b - a, c - a
    n = torch.linalg.cross(u, v, dim=-1)
    return n / torch.sqrt(torch.sum(n ** 2, dim=-1, keepdim=True) + 1e-12)


def normal_loss(pos_1xvx3, faces_fx3):
    face = gather_faces(pos_1xvx3, faces_fx3)
    n = get_normal(face[:, :, 0], face[:, :, 1], face[:, :, 2])
    _, pairs = native.face_adjacency(face[0].detach().numpy())
    if pairs.shape[1] == 0 or pairs.sum() == 0:
        return torch.zeros(pos_1xvx3.shape[0])
    pairs = torch.from_numpy(pairs)
    return (1 - torch.sum(n[:, pairs[0]] * n[:, pairs[1]], dim=-1)).mean(dim=-1)


def surface_losses(pos_bxvx3, boundary_list, gt_bxsx3, u_list, v_list):
    """Per-sample loop of DefTet.forward (deftet.py:138-184) -> (chamfer, analytic, normal) each (B,), with autograd."""
    ch, an, nl = [], [], []
    for b, faces in enumerate(boundary_list):
        if faces.shape[0] == 0:
            one = torch.ones(1)
            ch.append(one); an.append(one); nl.append(one)
            continue
        pos = pos_bxvx3[b:b + 1]
        surf = gather_faces(pos, faces)
        nl.append(normal_loss(pos, faces))
        q = sample_points(surf, u_list[b], v_list[b]).reshape(1, -1, 3)
        ch.append(chamfer(q, gt_bxsx3[b:b + 1]).mean(dim=-1))
        an.append(point_mesh_distance(gt_bxsx3[b:b + 1], surf).mean(dim=-1).mean(dim=-1))
    return torch.cat(ch), torch.cat(an), torch.cat(nl)


def peel2mask(ims_bxpxkxd):
    """Front-to-back compositing of the K depth-sorted slots (5_rendereq/deftetrneder.py:31-64), white background."""
    mask = torch.clamp(ims_bxpxkxd[..., :1], 1e-10, 1.0 - 1e-10)
    color = ims_bxpxkxd[..., 1:]
    shift = torch.nn.functional.pad(1 - mask[:, :, :-1, :], pad=(0, 0, 1, 0), mode="constant", value=1)
    vis = mask * torch.cumprod(shift, dim=2)
    xcolor = (color * vis).sum(dim=2)
    xvis = vis.sum(2)
    return xcolor + (1.0 - xvis), xvis

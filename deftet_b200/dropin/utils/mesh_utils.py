"""Drop-in for the hot-path functions of reference utils/mesh_utils.py (:16-53, :290-299, :360-374)."""
import torch

from deftet_b200.search import NearestNeighbor
from deftet_b200.surface import tet_analytic_distance_f_batch, tet_face_adj_m_f_idx

EPS = 1e-10


def get_normal(a, b, c):
    n = torch.linalg.cross(b - a, c - a, dim=-1)
    return n / (torch.sqrt(torch.sum(n ** 2, dim=-1, keepdim=True) + 1e-12))


def get_surface_normal_loss(vertices_bxnx3, faces_bxfx3):
    face = torch.gather(input=vertices_bxnx3.unsqueeze(dim=-2).expand(-1, -1, 3, -1),
                        index=faces_bxfx3.unsqueeze(dim=-1).expand(-1, -1, -1, 3), dim=1)
    normal_face = get_normal(face[:, :, 0, :], face[:, :, 1, :], face[:, :, 2, :])
    with torch.no_grad():
        one_face_adj_idx = tet_face_adj_m_f_idx(face[0].float())
    if one_face_adj_idx.sum() == 0:
        return torch.zeros(vertices_bxnx3.shape[0], device=faces_bxfx3.device).float()
    normal_loss = 1 - torch.sum(normal_face[:, one_face_adj_idx[0]] * normal_face[:, one_face_adj_idx[1]], dim=-1)
    return normal_loss.mean(dim=-1)


def sample_surf_point_batch(face_bxfx3x3, each_face_num=20):
    a, b, c = face_bxfx3x3[:, :, 0:1, :], face_bxfx3x3[:, :, 1:2, :], face_bxfx3x3[:, :, 2:3, :]
    n_batch, n_face = a.shape[0], a.shape[1]
    u = torch.sqrt(torch.rand(size=(n_batch, n_face, each_face_num, 1), device=face_bxfx3x3.device))
    v = torch.rand(size=(n_batch, n_face, each_face_num, 1), device=face_bxfx3x3.device)
    return (1 - u) * a + (u * (1 - v)) * b + u * v * c


def point_point_distance(a_bxnx3, b_bxmx3):
    closest_index_in_S2 = NearestNeighbor()(a_bxnx3, b_bxmx3)
    closest_S2 = torch.gather(input=b_bxmx3, dim=1, index=closest_index_in_S2.unsqueeze(-1).expand(-1, -1, 3))
    return torch.sqrt(torch.sum((a_bxnx3 - closest_S2) ** 2, dim=-1) + EPS)


def point_mesh_distance(a_bxnx3, mesh_bxfx3):
    batch_surface_length = torch.zeros(mesh_bxfx3.shape[0], device=mesh_bxfx3.device).float() + mesh_bxfx3.shape[1]
    tet_distance, _ = tet_analytic_distance_f_batch(a_bxnx3, mesh_bxfx3, batch_surface_length)
    return torch.sqrt(tet_distance + EPS)

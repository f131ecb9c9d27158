"""Drop-in for reference layers/DefTet/tet_analytic_distance_batch/utils.py."""
import torch

from deftet_b200 import _lib
from deftet_b200.surface import _AnalyticDistance as VarianceFunc  # noqa: F401
from deftet_b200.surface import tet_analytic_distance_f_batch  # noqa: F401


class _Ext:
    """Stands in for the pybind module ``tet_analytic_distance_batch`` (tet_analytic_distance.cpp:29-78): outputs are the
    caller's pre-allocated tensors, written / accumulated in place."""

    @staticmethod
    def forward(gt_point_clouds_bxpx3, face_bxfx3x3, closest_f, closest_d, n_face_b):
        with torch.no_grad():
            d, f = tet_analytic_distance_f_batch(gt_point_clouds_bxpx3, face_bxfx3x3, n_face_b)
        closest_d.copy_(d.reshape(closest_d.shape))
        closest_f.copy_(f.reshape(closest_f.shape))

    @staticmethod
    def backward(gt_point_clouds_bxpx3, face_bxfx3x3, closest_f, dl_dclosest_d, dldtet_bxfx3x3):
        pts, faces = gt_point_clouds_bxpx3.contiguous().float(), face_bxfx3x3.contiguous().float()
        B, S, F = pts.shape[0], pts.shape[1], faces.shape[1]
        cf, g = closest_f.contiguous().float(), dl_dclosest_d.contiguous().float()
        assert dldtet_bxfx3x3.is_contiguous() and dldtet_bxfx3x3.dtype == torch.float32
        with torch.cuda.device(pts.device):
            _lib.check(_lib.lib().dtb_point_face_distance_backward(_lib.ptr(pts), _lib.ptr(faces), _lib.ptr(cf), _lib.ptr(g), B, S, F,
                                                                   _lib.ptr(dldtet_bxfx3x3), _lib.stream_ptr()),
                       "dtb_point_face_distance_backward")


tet_analytic_distance_batch = _Ext()

"""Drop-in for reference layers/DefTet/check_condition_tetrahedron_base/utils.py (no JIT build, sm_100a kernel)."""
import torch
from torch.autograd import Function

from deftet_b200 import search as _search


class _Ext:
    """Stands in for the pybind module ``check_condition_cuda_tet_base`` (check_condition_tet.cpp:75-78)."""

    @staticmethod
    def forward(tet_bxfx4x3, point_pos_bxnx3, condition_bxnx1, bbox_filter_bxfx6=None):
        for name, t in (("tet_bxfx4x3", tet_bxfx4x3), ("point_pos_bxnx3", point_pos_bxnx3), ("condition_bxnx1", condition_bxnx1)):
            if not t.is_cuda:
                raise RuntimeError("%s must be a CUDA tensor" % name)
            if not t.is_contiguous():
                raise RuntimeError("%s must be contiguous" % name)
        condition_bxnx1.copy_(_search.point_in_tet_soup(tet_bxfx4x3, point_pos_bxnx3))

    @staticmethod
    def backward(*args):
        raise RuntimeError("check_condition_cuda_tet_base.backward is a dead 2-D leftover in the reference "
                           "(check_condition_tet_back.cu:82-183, never called); use deftet_b200.search.point_in_tet for gradients")


check_condition_cuda_tet_base = _Ext()


class TriRender2D(Function):
    @staticmethod
    def forward(ctx, tet_bxfx4x3, point_pos_bxnx3):
        return _search.point_in_tet_soup(tet_bxfx4x3.contiguous(), point_pos_bxnx3)

    @staticmethod
    def backward(ctx, condition_bxnx1):
        return None, None


check_condition_f_base = TriRender2D.apply

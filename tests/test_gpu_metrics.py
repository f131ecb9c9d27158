"""N4: evaluation metrics (utils/point_cloud_utils.py) on the CUDA kernels vs oracle/metrics.py (parity unpinned: the two Kaolin
primitives are restated from their call-site contract) and vs the reference module's own arithmetic above them."""
import numpy as np
import pytest
import torch

from oracle import metrics as orc_m
from tests.util import rel_err

pytestmark = pytest.mark.gpu


def _mesh(seed, F):
    rng = np.random.RandomState(seed)
    c = rng.rand(1, F, 1, 3) - 0.5
    fv = (c + 0.15 * (rng.rand(1, F, 3, 3) - 0.5)).astype(np.float32)
    fv[0, :4, 2] = fv[0, :4, 1]                                        # degenerate (zero-area) triangles
    fv[0, 4:8, :, 0] = fv[0, 4:8, :1, 0]                               # axis-aligned (vertical) triangles: invisible to A4, not here
    return fv


@pytest.mark.parametrize("P,F,seed", [(3000, 500, 0), (1000, 1, 1), (257, 129, 2)])
def test_point_to_mesh_distance(P, F, seed):
    from deftet_b200 import metrics
    fv = _mesh(seed, F)
    rng = np.random.RandomState(seed + 10)
    pts = ((rng.rand(1, P, 3) - 0.5) * 1.6).astype(np.float32)
    pts[0, :10] = fv[0, :10 % F + 1, 0][:1]                            # points exactly on a vertex
    d_ref, f_ref = orc_m.point_to_mesh_distance(pts, fv)
    d, f, t = metrics.point_to_mesh_distance(torch.from_numpy(pts).cuda(), torch.from_numpy(fv).cuda())
    assert d.shape == (1, P) and f.dtype == torch.int64 and t.dtype == torch.int32
    assert np.max(np.abs(np.sqrt(d.cpu().numpy()) - np.sqrt(d_ref))) < 1e-5          # tolerance on the distance itself (fp32 vs fp64)
    # a different face index is acceptable only on ties: the chosen face must be as close as the oracle's
    fi = f.cpu().numpy()
    diff = fi != f_ref
    if diff.any():
        d_sel, _ = orc_m.point_to_mesh_distance(pts[:, diff[0]].reshape(-1, 1, 3), fv[0][fi[0][diff[0]]].reshape(-1, 1, 3, 3))
        assert np.max(np.abs(np.sqrt(d_sel.reshape(-1)) - np.sqrt(d_ref[0][diff[0]]))) < 1e-5
    assert int(t.min()) >= 0 and int(t.max()) <= 6
    inside = (t == 0).cpu().numpy()[0]
    assert inside.sum() > 0 or F == 1


def test_sided_distance_and_reference_metrics():
    from deftet_b200 import metrics
    gen = torch.Generator().manual_seed(0)
    a = torch.rand(2, 4000, 3, generator=gen) - 0.5
    b = a[:, :3000] + 0.01 * torch.randn(2, 3000, 3, generator=gen)
    d_ref, i_ref = orc_m.sided_distance(a.numpy(), b.numpy())
    d, i = metrics.sided_distance(a.cuda(), b.cuda())
    assert i.dtype == torch.int64 and np.array_equal(i.cpu().numpy(), i_ref)
    assert rel_err(d, d_ref) < 1e-5
    # the reference's metric arithmetic (utils/point_cloud_utils.py) on top, restated with the oracle primitives
    esp = 1e-15
    s1, s2 = a[:1], b[:1]
    d12, d21 = np.sqrt(orc_m.sided_distance(s1.numpy(), s2.numpy())[0] + esp), np.sqrt(orc_m.sided_distance(s2.numpy(), s1.numpy())[0] + esp)
    assert abs(float(metrics.chamfer_distance(s1.cuda(), s2.cuda())) - (d12.mean() + d21.mean()) / 2) < 1e-6
    for radius, extend in ((0.01, False), (0.02, True)):
        pred_d, gt_d = d12, d21                                        # f_score(gt=s1, pred=s2)
        if extend:
            precision = (gt_d <= radius).sum() / gt_d.size
            recall = (pred_d <= radius).sum() / pred_d.size
        else:
            tp, fp, fn = (gt_d <= radius).sum(), (gt_d > radius).sum(), (pred_d > radius).sum()
            precision, recall = tp / (tp + fp), tp / (tp + fn)
        ref = 2 * precision * recall / (precision + recall + 1e-8)
        assert abs(float(metrics.f_score(s1.cuda(), s2.cuda(), radius, extend)) - ref) < 1e-4
    i12 = orc_m.sided_distance(s1.numpy(), s2.numpy())[1][0]
    i21 = orc_m.sided_distance(s2.numpy(), s1.numpy())[1][0]
    l1 = np.abs(s1[0].numpy() - s2[0].numpy()[i12]).sum(-1).mean() + np.abs(s2[0].numpy() - s1[0].numpy()[i21]).sum(-1).mean()
    assert abs(float(metrics.chamfer_distance_l1(s1.cuda(), s2.cuda())) - l1) < 1e-5
    x, y = torch.rand(1000, generator=gen), torch.rand(1000, generator=gen)
    ref_iou = ((x > .5) & (y > .5)).sum().item() / ((x > .5) | (y > .5)).sum().item()
    assert abs(float(metrics.iou(x.cuda(), y.cuda())) - ref_iou) < 1e-6


def test_hausdorff_distance_of_two_spheres():
    from deftet_b200 import metrics
    from tests.test_gpu_render import _icosphere
    v, f = _icosphere(3)
    va, vb = torch.from_numpy(v * 0.30).float().cuda(), torch.from_numpy(v * 0.35).float().cuda()
    fa = torch.from_numpy(f).long().cuda()
    gen = torch.Generator().manual_seed(1)
    d = torch.randn(5000, 3, generator=gen)
    d = (d / d.norm(dim=-1, keepdim=True)).cuda()
    avg, mx = metrics.hausdorff_distance(va, fa, vb, fa, d * 0.30, d * 0.35)
    assert 0.04 < float(avg) < 0.056 and 0.045 < float(mx) < 0.06        # concentric spheres 0.05 apart (faceted: slightly less / more)
    # kaolin-shaped shim resolves to the same functions
    import importlib.util, os
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "deftet_b200", "dropin", "kaolin", "metrics", "trianglemesh.py")
    spec = importlib.util.spec_from_file_location("kal_tm", root)
    m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
    assert m.point_to_mesh_distance is metrics.point_to_mesh_distance

"""N3 (diff_render topology editing + regularisers + fused projection): CUDA kernels vs the golden fixture generated from the
reference's own Python (tests/golden/diffrender_res8.npz) and vs the numpy oracle (oracle/topology.py) at larger sizes.
Index-valued results: bit-exact, reference order.  Float results: <= 1e-5 relative (max-normalised)."""
import os

import numpy as np
import pytest
import torch

from oracle import topology as orc_t
from tests.util import rel_err

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "diffrender_res8.npz")


def _cuda(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return t if dtype is None else t.to(dtype)


def _eq(t, ref):
    return np.array_equal(t.cpu().numpy().astype(np.int64), np.asarray(ref).astype(np.int64))


def test_edges_and_subdivision_match_reference_golden():
    from deftet_b200 import topology
    G = np.load(GOLD)
    tets, pts, feat = _cuda(G["tets"]), _cuda(G["points"]), _cuda(G["feat"])
    P = pts.shape[0]
    edges, tet_edge = topology.tet_edges(tets, P)
    assert edges.dtype == torch.int32 and _eq(edges, G["edges"]) and _eq(tet_edge, G["tet_edge"])
    for name, sig in (("sub_all", None), ("sub_sig", _cuda(G["sub_sig"]))):
        p, f, t = topology.generate_subdivision(tets, pts, feat, sig)
        assert np.array_equal(p.cpu().numpy(), G[name + "_points"]) and np.array_equal(f.cpu().numpy(), G[name + "_feat"])
        assert _eq(t, G[name + "_tets"])


def test_geometry_tables_and_deletion_match_reference_golden():
    from deftet_b200 import topology
    G = np.load(GOLD)
    tets = _cuda(G["tets"])
    P = G["points"].shape[0]
    f3, ft2, fs2 = topology.tet_to_face_idx(P, tets, with_boundary=True)
    assert _eq(f3, G["face_fx3"]) and _eq(ft2, G["face_tet_fx2"]) and _eq(fs2, G["face_slot_fx2"])
    f3i, ft2i, _ = topology.tet_to_face_idx(P, tets, with_boundary=False)
    keep = G["face_tet_fx2"][:, 1] >= 0
    assert _eq(f3i, G["face_fx3"][keep]) and _eq(ft2i, G["face_tet_fx2"][keep])
    nbr = topology.tet_neighbours(tets, P)
    assert np.array_equal(np.sort(nbr.cpu().numpy().astype(np.int64), axis=1), np.sort(G["tet_neighbour_idx"], axis=1))
    assert _eq(nbr, orc_t.tet_neighbours(G["tets"], P))                     # and column i = neighbour across local face i
    table, adjsum = topology.generate_point_adj_idx(P, tets)
    assert _eq(table, G["point_adj_idx"]) and np.array_equal(adjsum.cpu().numpy(), G["point_adj_sum"])
    w = _cuda(G["point_weights"])
    for L in (1, 2, 3):
        for thres in (0.05, 0.5):
            kept, keep_mask = topology.delete_tet_by_weight(tets, w, nbr, thres, L)
            ref = G["del_L%d_t%03d" % (L, int(thres * 100))]
            assert _eq(kept, ref) and int(keep_mask.sum()) == ref.shape[0]
    kept, keep_mask = topology.delete_tet_by_weight(tets, w, nbr, 2.0, 3)   # nothing survives -> list unchanged (deftet.py:325-326)
    assert _eq(kept, G["tets"]) and bool(keep_mask.all())


def test_regularisers_match_reference_golden():
    from deftet_b200 import topology
    G = np.load(GOLD)
    x = _cuda(G["feat"][:, :4]).requires_grad_(True)
    table = _cuda(G["point_adj_idx"])
    wgt = _cuda(G["point_adj_sum"]) + 1e-10
    lap = topology.featlap(x, table, wgt)
    (lap * _cuda(G["featlap_gout"])).sum().backward()
    assert rel_err(lap.detach(), G["featlap"]) < 1e-5 and rel_err(x.grad, G["featlap_gx"]) < 1e-5
    p = _cuda(G["points"]).requires_grad_(True)
    vv = topology.volume_deviation(p, _cuda(G["tets"]))
    (vv * _cuda(G["volvar_gout"])).sum().backward()
    assert rel_err(vv.detach(), G["volvar"]) < 1e-5 and rel_err(p.grad, G["volvar_gpoint"]) < 1e-5


def test_project_faces_matches_reference_golden_and_autograd():
    from deftet_b200 import topology
    G = np.load(GOLD)
    pts = _cuda(G["points"]).requires_grad_(True)
    feat = _cuda(G["feat"][:, :4]).requires_grad_(True)
    faces = _cuda(G["face_fx3"])
    rot, cpos, proj = _cuda(G["cam_rot"]), _cuda(G["cam_pos"]), _cuda(G["cam_proj"])
    B, F = rot.shape[0], faces.shape[0]
    mult = 1000.0
    fz, fxy, ff = topology.project_faces(pts, feat, faces, rot, cpos, proj, multiplier=mult, sigmoid=True)
    assert fz.shape == (B, F, 3) and fxy.shape == (B, F, 3, 2) and ff.shape == (B, F, 3, 4)
    assert rel_err(fz, G["face_cam_bxfx9"].reshape(B, F, 3, 3)[..., 2]) < 1e-5
    assert rel_err(fxy, G["face_img_bxfx6"].reshape(B, F, 3, 2) * mult) < 1e-5
    gen = torch.Generator().manual_seed(0)
    gz, gxy, gff = torch.randn(B, F, 3, generator=gen).cuda(), torch.randn(B, F, 3, 2, generator=gen).cuda(), torch.randn(B, F, 3, 4, generator=gen).cuda()
    (fz * gz).sum().backward(retain_graph=True)
    ((fxy * gxy).sum() + (ff * gff).sum()).backward()
    # fp64 autograd through the reference expressions (cameraop.perspective, vertex2face, sigmoid)
    rp = torch.from_numpy(G["points"]).double().requires_grad_(True)
    rf = torch.from_numpy(G["feat"][:, :4]).double().requires_grad_(True)
    cam, img = orc_t.perspective(rp.unsqueeze(0).repeat(B, 1, 1), torch.from_numpy(G["cam_rot"]).double(), torch.from_numpy(G["cam_pos"]).double(),
                                 torch.from_numpy(G["cam_proj"]).double())
    f3 = torch.from_numpy(G["face_fx3"])
    rz = orc_t.vertex2face(cam, f3).reshape(B, F, 3, 3)[..., 2]
    rxy = orc_t.vertex2face(img * mult, f3).reshape(B, F, 3, 2)
    rff = orc_t.vertex2face(torch.sigmoid(rf).unsqueeze(0).repeat(B, 1, 1), f3).reshape(B, F, 3, 4)
    ((rz * gz.cpu().double()).sum() + (rxy * gxy.cpu().double()).sum() + (rff * gff.cpu().double()).sum()).backward()
    assert rel_err(ff, rff.detach()) < 1e-5
    assert rel_err(pts.grad, rp.grad) < 1e-5 and rel_err(feat.grad, rf.grad) < 1e-5


@pytest.mark.parametrize("res,frac", [(16, 0.3), (24, 1.0)])
def test_topology_matches_oracle_on_larger_grids(res, frac):
    from deftet_b200 import topology
    from deftet_b200.grid import acute_lattice_grid
    g = acute_lattice_grid(res)
    rng = np.random.RandomState(res)
    tets = g.tets[rng.permutation(g.n_tet)[: int(g.n_tet * 0.9)]]               # holes + arbitrary tet order
    P = g.n_vert
    pts = g.centred()
    feat = rng.rand(P, 5).astype(np.float32)
    sig = rng.rand(tets.shape[0]) < frac
    dt = _cuda(tets)
    edges, te = topology.tet_edges(dt, P)
    e_ref, te_ref = orc_t.tet_edges(tets)
    assert _eq(edges, e_ref) and _eq(te, te_ref)
    p, f, t = topology.generate_subdivision(dt, _cuda(pts), _cuda(feat), _cuda(sig))
    p_ref, f_ref, t_ref = orc_t.subdivide(tets, pts, feat, sig)
    assert np.array_equal(p.cpu().numpy(), p_ref) and np.array_equal(f.cpu().numpy(), f_ref) and _eq(t, t_ref)
    # the subdivided mesh is again a valid input: same volume, conforming faces (every face shared by at most two tets)
    f3, ft2, fs2 = topology.tet_to_face_idx(p.shape[0], t, with_boundary=True)
    r3, rt2, rs2 = orc_t.tet_to_face_idx(p.shape[0], t_ref)
    assert _eq(f3, r3) and _eq(ft2, rt2) and _eq(fs2, rs2)
    nbr = topology.tet_neighbours(dt, P)
    nbr_ref = orc_t.tet_neighbours(tets, P)
    assert _eq(nbr, nbr_ref)
    table, adjsum = topology.generate_point_adj_idx(P, dt)
    tb_ref, as_ref = orc_t.point_adj_idx(P, tets)
    assert _eq(table, tb_ref) and np.array_equal(adjsum.cpu().numpy(), as_ref)
    w = (rng.rand(P, 1) ** 6).astype(np.float32)
    for L, thres in ((0, 0.3), (1, 0.3), (3, 0.6), (5, 0.9)):
        kept, _ = topology.delete_tet_by_weight(dt, _cuda(w), nbr, thres, L)
        k_ref, _ = orc_t.delete_tets(tets, w, nbr_ref, L, thres)
        assert _eq(kept, k_ref if k_ref.shape[0] else tets)
    # regularisers vs fp64 torch autograd of the reference expressions
    x = _cuda(feat).requires_grad_(True)
    lap = topology.featlap(x, table, adjsum + 1e-10)
    gen = torch.Generator().manual_seed(1)
    gl = torch.rand(P, 5, generator=gen)
    (lap * gl.cuda()).sum().backward()
    rx = torch.from_numpy(feat).double().requires_grad_(True)
    rl = orc_t.featlap(rx, torch.from_numpy(tb_ref), torch.from_numpy(as_ref).double() + 1e-10)
    (rl * gl.double()).sum().backward()
    assert rel_err(lap.detach(), rl.detach()) < 1e-5 and rel_err(x.grad, rx.grad) < 1e-5
    dp = _cuda(pts + rng.uniform(-0.2 / res, 0.2 / res, size=pts.shape).astype(np.float32)).requires_grad_(True)
    vv = topology.volume_deviation(dp, dt)
    gv = torch.rand(tets.shape[0], generator=gen)
    (vv * gv.cuda()).sum().backward()
    rp = dp.detach().cpu().double().requires_grad_(True)
    rv = orc_t.volume_deviation(rp, torch.from_numpy(tets))
    (rv * gv.double()).sum().backward()
    assert rel_err(vv.detach(), rv.detach()) < 1e-5 and rel_err(dp.grad, rp.grad) < 1e-5


def test_topology_edge_cases():
    from deftet_b200 import topology
    one = torch.tensor([[0, 1, 2, 3]], dtype=torch.int64).cuda()
    edges, te = topology.tet_edges(one, 4)
    assert edges.tolist() == [[0, 1], [0, 2], [0, 3], [1, 2], [1, 3], [2, 3]] and te.tolist() == [[0, 1, 2, 3, 4, 5]]
    pts = torch.eye(4, 3).cuda()
    p, f, t = topology.generate_subdivision(one, pts, pts, None)
    assert p.shape == (10, 3) and t.shape == (8, 4) and int(t.max()) == 9
    p, f, t = topology.generate_subdivision(one, pts, pts, torch.zeros(1, dtype=torch.bool).cuda())
    assert t.tolist() == [[0, 1, 2, 3]]
    f3, ft2, _ = topology.tet_to_face_idx(4, one, with_boundary=True)
    assert f3.shape == (4, 3) and bool((ft2[:, 1] == -1).all())
    assert topology.tet_to_face_idx(4, one, with_boundary=False)[0].shape == (0, 3)
    assert topology.tet_neighbours(one, 4).tolist() == [[-1, -1, -1, -1]]
    table, deg = topology.generate_point_adj_idx(6, one)                       # vertices 4, 5 are isolated
    assert table.shape == (6, 3) and table[4].tolist() == [-1, -1, -1] and deg.reshape(-1).tolist() == [3, 3, 3, 3, 0, 0]
    with pytest.raises(RuntimeError):
        topology.tet_edges(torch.zeros(1, 4, dtype=torch.int64), 4)            # CPU tensors are refused


def _cams(B):
    th = np.linspace(0.2, 2.0, B)
    rot = np.stack([np.array([[np.cos(t), 0, np.sin(t)], [0, 1, 0], [-np.sin(t), 0, np.cos(t)]]) for t in th]).astype(np.float32)
    pos = np.stack([rot[b].T @ np.array([0, 0, 4.0]) for b in range(B)]).astype(np.float32)
    proj = np.array([[2.5], [2.5], [-1.0]], dtype=np.float32)
    return _cuda(rot), _cuda(pos), _cuda(proj)


def test_deftet_model_fused_forward_equals_generic_path_and_survives_topology_edits(tmp_path):
    """The diff_render model mirror end to end: fused projection + fused rasterizer/compositor == the reference's per-vertex route
    (perspective -> vertex2face -> sparse render -> peel2mask), values and gradients; then subdivision and deletion."""
    from deftet_b200 import diffrender, render
    model = diffrender.Deftet(str(tmp_path), res=8, coef=2.0, feature_dim=4, seed=0)     # no .tet file there: generated lattice
    model.sethw(48, 48, 1000)
    assert [tuple(p.shape) for p in model.parameters()] == [(model.n_point, 3), (model.n_point, 4)]      # pointmov first (optim script :140-143)
    B = 2
    rot, pos, proj = _cams(B)
    sample = torch.ones(48, 48, dtype=torch.bool).cuda()
    K = 64

    def dense_renderfunc(xy, xydep, p3d, p2d, feat, faces, viewdir=False, depth=False, istraining=False):
        # the reference's rendermeshcolor body, on the drop-in sparse renderer + dense peel2mask
        feat = torch.sigmoid(feat)
        Bv, F = p3d.shape[0], faces.shape[0]
        f = faces.reshape(-1)
        ims, _ = render.deftet_sparse_render(xy, xydep, p3d[:, f, 2].reshape(Bv, F, 3), p2d[:, f].reshape(Bv, F, 3, 2), feat[:, f].reshape(Bv, F, 3, -1), knum=K)
        return diffrender.peel2mask(ims, None)

    outs = []
    for fn in (diffrender.rendermeshcolor, dense_renderfunc):
        model.zero_grad()
        col, mask = model(sample, rot, pos, proj, fn, knum=K) if fn is diffrender.rendermeshcolor else model(sample, rot, pos, proj, fn)
        assert col.shape == (B, 48 * 48, 3) and mask.shape == (B, 48 * 48, 1)
        gen = torch.Generator().manual_seed(3)
        gc, gm = torch.rand(B, 48 * 48, 3, generator=gen).cuda(), torch.rand(B, 48 * 48, 1, generator=gen).cuda()
        lap = model.get_featlap(torch.cat([torch.sigmoid(model.get_feat()), model.get_mov()], dim=-1)).sum()
        var = (model.get_volume_variance() ** 2).sum()
        ((col * gc).sum() + (mask * gm).sum() + lap + 1e3 * var).backward()
        outs.append((col.detach(), mask.detach(), model.tfpointmov_px3.grad.clone(), model.tfpointfeat_pxd.grad.clone()))
    assert float(outs[0][1].max()) > 0.5                                      # the object is in view
    # the two routes project the vertices with differently rounded arithmetic (fused kernel vs torch matmul), so a pixel lying within
    # an ulp of a projected edge may change face: allow a handful of such pixels, everything else must agree to 1e-5
    for a, b in ((outs[0][0], outs[1][0]), (outs[0][1], outs[1][1])):
        bad = ((a - b).abs() > 1e-5 * float(b.abs().max())).float().mean()
        assert float(bad) < 2e-3, float(bad)
    assert rel_err(outs[0][2], outs[1][2]) < 5e-3 and rel_err(outs[0][3], outs[1][3]) < 5e-3
    # topology edits keep the model consistent
    n0, t0 = model.n_point, model.tftet_tx4.shape[0]
    model.subdivision(None)
    assert model.tftet_tx4.shape[0] == 8 * t0 and model.n_point > n0 and model.tfpointfeat_pxd.shape == (model.n_point, 4)
    col2, mask2 = model(sample, rot, pos, proj, diffrender.rendermeshcolor, knum=4 * K)
    assert torch.isfinite(col2).all() and float(mask2.max()) > 0.5
    with torch.no_grad():
        model.tfpointfeat_pxd[:, 0] = torch.where(model.tfpoint_px3.norm(dim=1) < 0.3, 4.0, -8.0)      # opaque ball, empty outside
    model.deletetet(0.05, diffrender.preprocess_save)
    assert 0 < model.tftet_tx4.shape[0] < 8 * t0
    sd = model.state_dict()
    assert set(["tfpointmov_px3", "tfpointfeat_pxd", "points", "tets", "feat_fixed"]) <= set(sd.keys())
    model.load_state_dict(sd)
    col3, mask3 = model(sample, rot, pos, proj, diffrender.rendermeshcolor, knum=2 * K)
    assert torch.isfinite(col3).all() and float(mask3.max()) > 0.5


def test_diff_render_dropin_modules_return_reference_types():
    """prepare_for_wz / utils_tetsv drop-ins: numpy in, numpy out, the reference's dtypes and values."""
    import importlib.util
    G = np.load(GOLD)
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "deftet_b200", "dropin", "diff_render")

    def load(name):
        spec = importlib.util.spec_from_file_location("dropin_dr_" + name, os.path.join(root, name + ".py"))
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        return m
    pw, ut = load("prepare_for_wz"), load("utils_tetsv")
    P = G["points"].shape[0]
    e = pw.generate_edge(G["tets"])
    assert e.dtype == np.int64 and np.array_equal(e, G["edges"])
    assert np.array_equal(pw.generate_tet_edge_idx(G["tets"], e), G["tet_edge"])
    p, f, t = pw.generate_subdivision(G["tets"], G["points"], G["feat"], G["sub_sig"])
    assert np.array_equal(p, G["sub_sig_points"]) and np.array_equal(f, G["sub_sig_feat"]) and np.array_equal(t, G["sub_sig_tets"])
    f3, ft2, fs2 = pw.tet_to_face_idx(P, G["tets"], with_boundary=True)
    assert np.array_equal(f3, G["face_fx3"]) and np.array_equal(ft2, G["face_tet_fx2"]) and np.array_equal(fs2, G["face_slot_fx2"])
    idx, s = pw.generate_point_adj_idx(P, G["tets"])
    assert idx.dtype == np.int64 and np.array_equal(idx, G["point_adj_idx"]) and np.array_equal(s, G["point_adj_sum"])
    adj_list, nbr = ut.tet_adj_share(G["tets"], P)
    assert len(adj_list) == 4 and np.array_equal(np.sort(nbr, axis=1), np.sort(G["tet_neighbour_idx"], axis=1))
    assert int(sum(a.nnz for a in adj_list)) == int((G["tet_neighbour_idx"] >= 0).sum())

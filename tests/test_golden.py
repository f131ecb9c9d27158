"""Pin the oracle to the reference: golden vectors produced by the reference's own code (tests/golden/make_golden.py),
the reference's compiled builders when oracle/_ref exists, and the known answers of BASELINE.md section 3."""
import os

import numpy as np
import pytest
import torch

from deftet_b200.grid import acute_lattice_grid, read_tet_file, snap_boundary
from oracle import builders as orc_b
from oracle import energies as orc_e
from oracle import native
from tests.util import deformed_grid, rel_err

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REF = "/root/reference"


def test_energy_oracle_matches_reference_class():
    z = np.load(os.path.join(GOLD, "energies_res8.npz"))
    g, pos, tet = deformed_grid(8, 2, seed=int(z["seed"]))
    assert np.array_equal(pos.numpy(), z["pos"])                       # the synthetic input is reproducible
    inv = orc_e.tet_inverse_v(torch.from_numpy(g.centred()), tet)
    assert rel_err(inv, z["inverse_v"]) < 1e-6
    out = orc_e.energies_with_grad(pos, tet, inv, tuple(z["weights"]))
    assert rel_err(out["amips"], z["amips"]) < 1e-6
    assert rel_err(out["edge"], z["edge"]) < 1e-6
    assert rel_err(out["volvar"], z["volvar"]) < 1e-5
    assert rel_err(out["grad"], z["grad"]) < 1e-5


def test_builder_oracle_matches_reference_outputs():
    z = np.load(os.path.join(GOLD, "builders_res8.npz"))
    g = acute_lattice_grid(8)
    f3, ft2, fs2, bnd = orc_b.tet_to_face(g.n_vert, g.tets)
    assert np.array_equal(f3, z["f3"]) and np.array_equal(ft2, z["ft2"]) and np.array_equal(fs2, z["fs2"]) and np.array_equal(bnd, z["bnd"])
    assert np.array_equal(orc_b.tet_adj_share(g.n_vert, g.tets), z["share"])
    assert np.array_equal(orc_b.tet_face_adj(g.n_vert, g.tets), z["face_adj"])
    assert np.array_equal(orc_b.tet_to_adj_edges(g.tets), z["point_adj_sorted"])
    m, inv = orc_b.colaps_v(g.centred()[g.tets.reshape(-1)])
    assert np.array_equal(m, z["colaps_map"]) and np.array_equal(inv, z["colaps_inv"])


@pytest.mark.skipif(native.ref_lib("tet_adj_share") is None, reason="oracle/_ref not built (needs /root/reference at build time)")
def test_builder_oracle_matches_compiled_reference_libs():
    g = acute_lattice_grid(12)
    o, n = native.ref_run_tet_builder("tet_adj_share", g.tets, g.n_vert, g.n_tet * 8, 3)
    assert np.array_equal(orc_b.tet_adj_share(g.n_vert, g.tets), o[:2 * n])
    o, n = native.ref_run_tet_builder("tet_face_adj", g.tets, g.n_vert, g.n_tet * 200, 2)
    assert np.array_equal(orc_b.tet_face_adj(g.n_vert, g.tets), o[:n])
    rng = np.random.default_rng(1)
    pts = np.round(rng.random((3000, 3)).astype(np.float32) * 40) / 40 - 0.5
    pts[::7] *= -0.0
    m, inv = orc_b.colaps_v(pts)
    mc, ic = native.ref_colaps_v(pts)
    assert np.array_equal(m, mc) and np.array_equal(inv, ic)


def test_cube40_known_answers_recorded():
    z = np.load(os.path.join(GOLD, "cube40_known.npz"))
    assert float(z["amips"]) == 3.0 and abs(float(z["edge"]) - 1.0805563) < 1e-6
    assert int(z["shared_faces"]) == 92604 and int(z["directed_edges"]) == 118566 and int(z["n_tet"]) == 47472


@pytest.mark.reference
def test_oracle_reproduces_known_answers_on_shipped_cube40():
    """T0 of SURVEY.md section 7: the reference's shipped grid through the oracle."""
    v, t = read_tet_file(os.path.join(REF, "diff_render/diftet_6_subdiv/data/cube_40_tet.tet"))
    assert v.shape[0] == 9472 and t.shape[0] == 47472
    pos = torch.from_numpy((v - 0.5).astype(np.float32))
    tet = torch.from_numpy(t)
    inv = orc_e.tet_inverse_v(pos, tet)
    soup = orc_e.gather_tets(pos.unsqueeze(0), tet)
    assert float(orc_e.amips_energy(soup, inv)) == pytest.approx(3.0, abs=2e-6)
    assert float(orc_e.edge_length(soup)) == pytest.approx(1.0805563, rel=1e-6)
    assert orc_b.tet_adj_share(v.shape[0], t[:4000]).shape[1] == 3


def test_grid_generator_properties():
    for res in (8, 16):
        g = acute_lattice_grid(res)
        a = g.vertices[g.tets]
        det = np.linalg.det(np.stack([a[:, 1] - a[:, 0], a[:, 2] - a[:, 0], a[:, 3] - a[:, 0]], 1))
        assert det.min() > 0                                              # positively oriented like the QuarTet grids
        assert g.vertices.min() == 0.0 and g.vertices.max() == 1.0
        f3, ft2, _, bnd = orc_b.tet_to_face(g.n_vert, g.tets)
        assert 2 * f3.shape[0] + bnd.shape[0] == 4 * g.n_tet                # conforming: every face is shared by <= 2 tets
        deg = np.bincount(orc_b.tet_to_adj_edges(g.tets)[:, 0], minlength=g.n_vert)
        assert deg.max() <= 14                                            # acute lattice
        v2, mask = snap_boundary(g.vertices, res)
        assert np.array_equal(v2, g.vertices) and mask.shape == (g.n_vert, 3)
    g2 = acute_lattice_grid(8)
    assert np.array_equal(g2.tets, acute_lattice_grid(8).tets)


@pytest.mark.skipif(native.ref_kernel_lib("point_in_tet") is None, reason="oracle/_ref/kernels not built (needs /root/reference + nvcc at build time)")
def test_c_restatement_matches_reference_device_functions_compiled_for_host():
    """The CUDA-only reference kernels keep their per-element math in `__host__ __device__` templates; oracle/build_ref_kernels.sh
    compiles exactly those functions for the CPU.  The C restatement (oracle/deftet_oracle.c) must agree with them bit for bit."""
    from oracle import surface as orc_s
    from tests.util import sphere_occupancy
    g, pos, tet = deformed_grid(8, 2, seed=21)
    gen = torch.Generator().manual_seed(0)
    pts = (torch.rand(2, 1500, 3, generator=gen) - 0.5) * 1.05
    pts[0, :100] = pos[0, :100]
    soup = orc_e.gather_tets(pos, tet).numpy()
    assert np.array_equal(native.point_in_tet(soup, pts.numpy()), native.ref_point_in_tet(soup, pts.numpy()))
    # undeformed grid: points exactly on lattice planes / vertices
    g0, pos0, tet0 = deformed_grid(8, 1, seed=0, amp=0.0)
    p0 = torch.round(pts[:1] * 8) / 8
    soup0 = orc_e.gather_tets(pos0, tet0).numpy()
    assert np.array_equal(native.point_in_tet(soup0, p0.numpy()), native.ref_point_in_tet(soup0, p0.numpy()))
    # A4 forward / backward on a boundary-face soup (deformed and undeformed: k3 == 0 faces) + random soup
    f3, ft2, _, _ = orc_b.tet_to_face(g.n_vert, g.tets)
    for pp, tt in ((pos, tet), (pos0, tet0)):
        occ = sphere_occupancy(pp[:1], tt, [[0.0, 0.0, 0.0]], [0.3])
        bnd = orc_s.get_boundary_index(torch.from_numpy(f3), torch.from_numpy(ft2), occ)[0]
        faces = orc_s.gather_faces(pp[:1], bnd).numpy()
        q = pts[:1, :800].numpy()
        d, f = native.point_face_distance(q, faces)
        dr, fr = native.ref_point_face_distance(q, faces)
        assert np.array_equal(f, fr) and np.array_equal(d, dr)
        gd = torch.rand(1, 800, 1, generator=gen).numpy()
        assert np.array_equal(native.point_face_distance_bwd(q, faces, f, gd), native.ref_point_face_distance_bwd(q, faces, fr, gd))
        adj, _ = native.face_adjacency(faces[0])
        assert np.array_equal(adj, native.ref_face_adjacency(faces[0]))
    rs = (torch.rand(1, 300, 3, 3, generator=gen) - 0.5).numpy()
    rs[0, :20, :, 0] = rs[0, :20, :1, 0]                      # vertical faces
    q = ((torch.rand(1, 500, 3, generator=gen) - 0.5) * 1.3).numpy()
    d, f = native.point_face_distance(q, rs)
    dr, fr = native.ref_point_face_distance(q, rs)
    assert np.array_equal(f, fr) and np.array_equal(d, dr)
    gd = torch.rand(1, 500, 1, generator=gen).numpy()
    assert np.array_equal(native.point_face_distance_bwd(q, rs, f, gd), native.ref_point_face_distance_bwd(q, rs, fr, gd))


# ---- diff_render topology / regulariser oracle vs the reference's own Python (tests/golden/make_golden_diffrender.py) ----------
def _dr():
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "diffrender_res8.npz"))


def test_topology_oracle_edges_and_subdivision_match_reference():
    from oracle import topology as orc_t
    G = _dr()
    edges, te = orc_t.tet_edges(G["tets"])
    assert np.array_equal(edges, G["edges"]) and np.array_equal(te, G["tet_edge"])
    for name, sig in (("sub_all", None), ("sub_sig", G["sub_sig"])):
        p, f, t = orc_t.subdivide(G["tets"], G["points"], G["feat"], sig)
        assert np.array_equal(p, G[name + "_points"]) and np.array_equal(f, G[name + "_feat"])      # midpoints: bit-identical
        assert np.array_equal(t, G[name + "_tets"])


def test_topology_oracle_geometry_tables_match_reference():
    from oracle import topology as orc_t
    G = _dr()
    P = G["points"].shape[0]
    f3, ft2, fs2 = orc_t.tet_to_face_idx(P, G["tets"])
    assert np.array_equal(f3, G["face_fx3"]) and np.array_equal(ft2, G["face_tet_fx2"]) and np.array_equal(fs2, G["face_slot_fx2"])
    nbr = orc_t.tet_neighbours(G["tets"], P)
    assert np.array_equal(np.sort(nbr, axis=1), np.sort(G["tet_neighbour_idx"], axis=1))        # same neighbour sets (slot order differs)
    table, deg = orc_t.point_adj_idx(P, G["tets"])
    assert np.array_equal(table, G["point_adj_idx"]) and np.array_equal(deg, G["point_adj_sum"])
    for L in (1, 2, 3):
        for thres in (0.05, 0.5):
            kept, _ = orc_t.delete_tets(G["tets"], G["point_weights"], nbr, L, thres)
            assert np.array_equal(kept, G["del_L%d_t%03d" % (L, int(thres * 100))])


def test_topology_oracle_regularisers_and_projection_match_reference():
    import torch
    from oracle import topology as orc_t
    G = _dr()
    x = torch.from_numpy(G["feat"][:, :4]).clone().requires_grad_(True)
    lap = orc_t.featlap(x, torch.from_numpy(G["point_adj_idx"]), torch.from_numpy(G["point_adj_sum"]) + 1e-10)
    (lap * torch.from_numpy(G["featlap_gout"])).sum().backward()
    assert np.allclose(lap.detach().numpy(), G["featlap"], rtol=1e-6, atol=1e-7) and np.allclose(x.grad.numpy(), G["featlap_gx"], rtol=1e-5, atol=1e-7)
    p = torch.from_numpy(G["points"]).clone().requires_grad_(True)
    vv = orc_t.volume_deviation(p, torch.from_numpy(G["tets"]))
    (vv * torch.from_numpy(G["volvar_gout"])).sum().backward()
    assert rel_err(vv.detach(), G["volvar"]) < 1e-5 and rel_err(p.grad, G["volvar_gpoint"]) < 1e-5       # V - mean(V) cancels: max-normalised
    pts = torch.from_numpy(G["points"]).unsqueeze(0).repeat(2, 1, 1)
    cam, img = orc_t.perspective(pts, torch.from_numpy(G["cam_rot"]), torch.from_numpy(G["cam_pos"]), torch.from_numpy(G["cam_proj"]))
    assert np.array_equal(cam.numpy(), G["points_camera"]) and np.array_equal(img.numpy(), G["points_image"])
    f3 = torch.from_numpy(G["face_fx3"])
    assert np.array_equal(orc_t.vertex2face(cam, f3).numpy(), G["face_cam_bxfx9"])
    col, vis = orc_t.peel2mask(torch.from_numpy(G["peel_in"]))
    assert np.allclose(col.numpy(), G["peel_color"], rtol=1e-6, atol=1e-7) and np.allclose(vis.numpy(), G["peel_vis"], rtol=1e-6, atol=1e-7)


def test_metrics_oracle_point_triangle_distance_is_a_true_minimum():
    """oracle/metrics.py (parity unpinned: Kaolin absent) -- self-check of the restated contract: the closed-form distance is a lower
    bound of, and converges to, the distance to densely sampled points of the triangle."""
    from oracle import metrics as orc_m
    rng = np.random.RandomState(0)
    fv = rng.rand(1, 40, 3, 3) - 0.5
    fv[0, :5, 2] = fv[0, :5, 1] + 1e-9 * rng.rand(5, 3)            # needle triangles
    pts = (rng.rand(1, 60, 3) - 0.5) * 2
    d, f = orc_m.point_to_mesh_distance(pts, fv)
    n = 60
    u, v = np.meshgrid(np.linspace(0, 1, n), np.linspace(0, 1, n), indexing="ij")
    keep = (u + v) <= 1.0
    u, v = u[keep], v[keep]
    samples = (fv[0, :, None, 0] * (1 - u - v)[None, :, None] + fv[0, :, None, 1] * u[None, :, None] + fv[0, :, None, 2] * v[None, :, None])   # (F,S,3)
    ds = ((pts[0][:, None, None, :] - samples[None]) ** 2).sum(-1).min(-1)                                                           # (P,F)
    assert np.all(d[0] <= ds.min(-1) + 1e-12)
    assert np.allclose(np.sqrt(d[0]), np.sqrt(ds.min(-1)), atol=2.0 / n)
    sd, si = orc_m.sided_distance(pts, fv[:, :, 0])
    assert sd.shape == (1, 60) and np.all(sd[0] >= d[0] - 1e-12)       # a vertex is a point of its triangle


# ---- N4 second half: voxel-feature sampling -------------------------------------------------------------------------------------
DEVOX_CASES = (("r32", 32), ("r16", 16), ("r8", 8), ("r5_odd", 5), ("r40_big", 40))


def test_devox_oracle_matches_reference_function():
    """oracle/devox.py vs the outputs of the reference's own trilinear_devoxelize (tests/golden/make_golden_devox.py): the numpy
    restatement is bit-identical in the forward pass; the differentiable restatement reproduces its autograd."""
    import torch
    from oracle import devox as od
    g = np.load(os.path.join(GOLD, "devox.npz"))
    for name, R in DEVOX_CASES:
        feat, coords = g[name + "_feat"], g[name + "_coords"]
        assert np.array_equal(od.trilinear_devoxelize(feat, coords, R), g[name + "_out"]), name
        ft, ct = torch.tensor(feat, requires_grad=True), torch.tensor(coords, requires_grad=True)
        o = od.trilinear_devoxelize_torch(ft, ct, R)
        gf, gc = torch.autograd.grad(o, (ft, ct), torch.tensor(g[name + "_grad_out"]))
        assert np.abs(gf.numpy() - g[name + "_grad_feat"]).max() <= 1e-5 * max(1.0, np.abs(g[name + "_grad_feat"]).max()), name
        assert np.abs(gc.numpy() - g[name + "_grad_coords"]).max() <= 1e-5 * max(1.0, np.abs(g[name + "_grad_coords"]).max()), name
        # the border clip really is exercised: some coordinates carry no gradient
        assert (g[name + "_grad_coords"] == 0).any() and (g[name + "_grad_coords"] != 0).any()


def test_devox_oracle_sample_f_matches_reference():
    import torch
    from oracle import devox as od
    g = np.load(os.path.join(GOLD, "devox.npz"))
    pos = torch.tensor(g["sf_pos"], requires_grad=True)
    cl = [torch.tensor(g["sf_feat%d" % i], requires_grad=True) for i in range(3)]
    o = od.sample_f(pos, cl)
    assert np.array_equal(o.detach().numpy(), g["sf_out"])
    grads = torch.autograd.grad(o, [pos] + cl, torch.tensor(g["sf_grad_out"]))
    assert np.abs(grads[0].numpy() - g["sf_grad_pos"]).max() <= 1e-5 * np.abs(g["sf_grad_pos"]).max()
    for i in range(3):
        assert np.abs(grads[1 + i].numpy() - g["sf_grad_feat%d" % i]).max() <= 1e-5 * max(1.0, np.abs(g["sf_grad_feat%d" % i]).max())

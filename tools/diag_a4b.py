"""Dev diagnostic: A4 forward + backward vs the reference's device kernel AND the non-contracted oracle, per point."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from deftet_b200 import surface
from deftet_b200.engine import GeometryEngine
from deftet_b200.grid import acute_lattice_grid
from deftet_b200.synthetic import analytic_scene
from oracle import native as orc, ref_cuda

res, seed = int(sys.argv[1]), int(sys.argv[2])
dev = torch.device("cuda:0")
grid = acute_lattice_grid(res)
B, P, S = 8, 100000, 100000
sc = analytic_scene(grid, B, P, S, seed, dev)
eng = GeometryEngine(grid.centred(), grid.tets, max_boundary_faces=16384, device=dev)
faces, counts, _ = surface.boundary_faces(eng.face_table, sc["occ"], 16384)
soup, cd, cf = surface.closest_faces(sc["pos"], faces, counts, sc["gt"])
cnt = counts.tolist()
for b in range(B):
    nb = cnt[b]
    fb = soup[b:b + 1, :nb].contiguous()
    gb = sc["gt"][b:b + 1].contiguous()
    d_ref, f_ref = ref_cuda.point_face_distance(gb, fb)
    d_our = cd[b]
    diff = (d_our - d_ref.reshape(-1)).abs()
    scale = float(d_ref.max())
    bad = torch.nonzero(diff > 1e-5 * max(scale, 1e-3)).reshape(-1)
    print("sample", b, "F", nb, "max d_ref", scale, "fwd n_bad", int(bad.numel()), "max abs diff", float(diff.max()), flush=True)
    if bad.numel():
        idx = bad[:256].cpu()
        pts = gb[0].cpu()[idx].reshape(1, -1, 3).numpy()
        d_or, f_or = orc.point_face_distance(pts, fb.cpu().numpy())
        d_or, f_or = d_or.reshape(-1), f_or.reshape(-1)
        ours = d_our.cpu()[idx].numpy(); ourf = cf[b].cpu()[idx].numpy()
        refd = d_ref.reshape(-1).cpu()[idx].numpy(); reff = f_ref.reshape(-1).cpu()[idx].numpy()
        print("  FWD ours == oracle brute force (bitwise):", int((ours == d_or).sum()), "of", len(idx), " face ids equal:", int((ourf == f_or).sum()))
        fcpu = fb[0].cpu().numpy()
        for k in range(min(4, len(idx))):
            fo, fr = int(ourf[k]), int(reff[k])
            d_fr, _ = orc.point_face_distance(pts[:, k:k + 1], fcpu[fr:fr + 1].reshape(1, 1, 3, 3))
            tri = fcpu[fr]; n = np.cross(tri[1] - tri[0], tri[2] - tri[0]); nz = abs(n[2]) / np.linalg.norm(n)
            tri2 = fcpu[fo]; n2 = np.cross(tri2[1] - tri2[0], tri2[2] - tri2[0]); nz2 = abs(n2[2]) / np.linalg.norm(n2)
            print("   pt", pts[0, k], "ours d=%.9g f=%d (|nz|=%.5f) | ref d=%.9g f=%d (|nz|=%.5f) | oracle: d=%.9g f=%d; on ref's face %.9g"
                  % (ours[k], fo, nz2, refd[k], fr, nz, d_or[k], int(f_or[k]), float(d_fr.reshape(-1)[0])))
    # backward: per-point contributions.  upstream 1 for every point, but evaluate per point by giving each point its own "face copy":
    # use the oracle (non-contracted) backward on CPU for all points of this sample, and the ref kernel; compare per FACE, then drill down.
    f_our = cf[b].reshape(1, S, 1).contiguous()
    g1 = torch.ones(1, S, 1, device=dev)
    g_ref = ref_cuda.point_face_distance_bwd(gb, fb, f_our, g1)[0]                # (nb,3,3)
    dfaces = fb.clone().requires_grad_(True)
    d2, _ = surface.tet_analytic_distance_f_batch(gb, dfaces, torch.tensor([float(nb)], device=dev))
    d2.sum().backward()
    g_our = dfaces.grad[0]
    g_orc = torch.from_numpy(orc.point_face_distance_bwd(gb.cpu().numpy(), fb.cpu().numpy(), f_our.cpu().numpy(), g1.cpu().numpy()))[0]
    sc_g = float(g_ref.abs().max())
    e_ref = (g_our - g_ref).abs().reshape(nb, -1).max(dim=1).values.cpu()
    e_orc = (g_our.cpu() - g_orc).abs().reshape(nb, -1).max(dim=1).values
    print("  BWD scale %.4g: vs ref kernel max %.3g (faces > 1e-5 rel: %d) | vs non-contracted oracle max %.3g (faces > 1e-5 rel: %d)"
          % (sc_g, float(e_ref.max()) / sc_g, int((e_ref > 1e-5 * sc_g).sum()), float(e_orc.max()) / sc_g, int((e_orc > 1e-5 * sc_g).sum())), flush=True)
    worst = int(torch.argmax(e_ref))
    pts_w = torch.nonzero(f_our.reshape(-1) == worst).reshape(-1)
    # per point of the worst face: ref-kernel gradient vs oracle gradient (single-point launches)
    nshow = 0
    for i in pts_w.tolist():
        p1 = gb[:, i:i + 1].contiguous(); f1 = torch.zeros(1, 1, 1, device=dev); one = torch.ones(1, 1, 1, device=dev)
        tri = fb[:, worst:worst + 1].contiguous()
        gr = ref_cuda.point_face_distance_bwd(p1, tri, f1, one)[0, 0].cpu()
        go = torch.from_numpy(orc.point_face_distance_bwd(p1.cpu().numpy(), tri.cpu().numpy(), f1.cpu().numpy(), one.cpu().numpy()))[0, 0]
        if float((gr - go).abs().max()) > 1e-6 * max(float(go.abs().max()), 1e-12):
            print("   face", worst, "pt", i, "ref-kernel grad", gr.reshape(-1).numpy().round(6), "\n      oracle grad", go.reshape(-1).numpy().round(6))
            nshow += 1
            if nshow >= 3: break

"""Dev diagnostic: where does the A4 distance differ from the reference's device kernel at scale, and is it us or FMA contraction?"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from deftet_b200 import surface
from deftet_b200.engine import GeometryEngine
from deftet_b200.grid import acute_lattice_grid
from deftet_b200.synthetic import analytic_scene
from oracle import native as orc, ref_cuda

res, seed = int(sys.argv[1]) if len(sys.argv) > 1 else 40, int(sys.argv[2]) if len(sys.argv) > 2 else 2000
dev = torch.device("cuda:0")
grid = acute_lattice_grid(res)
B, P, S = 8, int(sys.argv[3]) if len(sys.argv) > 3 else 100000, 100000
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
    print("sample", b, "F", nb, "max d_ref", scale, "n_bad", int(bad.numel()), "max abs diff", float(diff.max()))
    if bad.numel() == 0:
        continue
    idx = bad[:64].cpu()
    pts = gb[0].cpu()[idx].reshape(1, -1, 3).numpy()
    d_or, f_or = orc.point_face_distance(pts, fb.cpu().numpy())
    d_or, f_or = d_or.reshape(-1), f_or.reshape(-1)
    ours = d_our.cpu()[idx].numpy(); ourf = cf[b].cpu()[idx].numpy()
    refd = d_ref.reshape(-1).cpu()[idx].numpy(); reff = f_ref.reshape(-1).cpu()[idx].numpy()
    print("  ours == oracle brute force (bitwise):", int((ours == d_or).sum()), "of", len(idx), " face ids equal:", int((ourf == f_or).sum()))
    fcpu = fb[0].cpu().numpy()
    for k in range(min(6, len(idx))):
        fo, fr = int(ourf[k]), int(reff[k])
        # distance of the reference's face under the non-contracted oracle, and of our face
        d_fr, _ = orc.point_face_distance(pts[:, k:k + 1], fcpu[fr:fr + 1].reshape(1, 1, 3, 3))
        d_fo, _ = orc.point_face_distance(pts[:, k:k + 1], fcpu[fo:fo + 1].reshape(1, 1, 3, 3))
        tri = fcpu[fr]
        n = np.cross(tri[1] - tri[0], tri[2] - tri[0]); nz = abs(n[2]) / np.linalg.norm(n)
        tri2 = fcpu[fo]
        n2 = np.cross(tri2[1] - tri2[0], tri2[2] - tri2[0]); nz2 = abs(n2[2]) / np.linalg.norm(n2)
        print("   pt", pts[0, k], "ours d=%.9g f=%d (|nz|=%.4f) | ref d=%.9g f=%d (|nz|=%.4f) | oracle on ref's face %.9g, on our face %.9g"
              % (ours[k], fo, nz2, refd[k], fr, nz, float(d_fr.reshape(-1)[0]), float(d_fo.reshape(-1)[0])))

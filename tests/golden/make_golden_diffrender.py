"""Generate tests/golden/diffrender_res8.npz from the REFERENCE'S OWN diff_render Python (run here, where /root/reference is
mounted):  python tests/golden/make_golden_diffrender.py

Imports, unmodified, from /root/reference/diff_render/diftet_6_subdiv/3_model:
  prepare_for_wz.py  generate_edge, generate_tet_edge_idx, generate_subdivision, generate_point_adj_idx, tet_to_face_idx, delete_tet
  utils_tetsv.py     tet_adj_share (tet_neighbour_idx)
  deftet.py          Deftet.get_featlap, get_volume_variance, pointweights2tetweights, tetweights2tetneighbourweights
                     (called as plain functions on a stand-in `self`; `config` / `utils_mesh` (needs cv2) are stubbed)
  cameraop.py        perspective;   4_render/vertex2face.py vertex2face;   5_rendereq/deftetrneder.py is NOT importable (kaolin)
                     -- peel2mask is exec'd from its source text.
The fixture pins oracle/topology.py and the CUDA kernels of deftet_b200/csrc/topology.cu (tests/test_golden.py,
tests/test_gpu_topology.py); it travels to the GPU box, the reference does not.
"""
import os
import re
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference/diff_render/diftet_6_subdiv"
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden")


def import_reference():
    cfg = types.ModuleType("config")
    cfg.rootdir = REF
    sys.modules["config"] = cfg
    um = types.ModuleType("utils_mesh")
    um.savemesh = um.savemeshfweights = um.savemeshfweightscolor = None
    sys.modules["utils_mesh"] = um
    sys.path.insert(0, os.path.join(REF, "3_model"))
    sys.path.insert(0, os.path.join(REF, "4_render"))
    import prepare_for_wz, utils_tetsv, cameraop, vertex2face          # noqa: E401
    import deftet as ref_deftet
    src = open(os.path.join(REF, "5_rendereq", "deftetrneder.py")).read()
    m = re.search(r"^def peel2mask\(.*?(?=^def )", src, flags=re.S | re.M)
    ns = {"torch": torch, "nn": torch.nn}
    exec(m.group(0), ns)
    return prepare_for_wz, utils_tetsv, cameraop, vertex2face, ref_deftet, ns["peel2mask"]


def main():
    from deftet_b200.grid import acute_lattice_grid
    pw, ut, cam, v2f, rd, peel2mask = import_reference()
    g = acute_lattice_grid(8)
    rng = np.random.RandomState(7)
    pts = (g.centred() + rng.uniform(-0.01, 0.01, size=(g.n_vert, 3))).astype(np.float32)
    feat = rng.rand(g.n_vert, 7).astype(np.float32)                     # 4 features + 3 pointmov, like Deftet.subdivision
    tets = g.tets.copy()
    P, T = g.n_vert, g.n_tet
    out = dict(points=pts, feat=feat, tets=tets)
    # ---- edges / subdivision ------------------------------------------------------------------------------------------
    edges = pw.generate_edge(tets)
    tet_edge = pw.generate_tet_edge_idx(tets, edges)
    out.update(edges=edges, tet_edge=tet_edge)
    p_all, f_all, t_all = pw.generate_subdivision(tets, pts, feat, None)
    out.update(sub_all_points=p_all, sub_all_feat=f_all, sub_all_tets=t_all)
    sig = rng.rand(T) < 0.4
    p_s, f_s, t_s = pw.generate_subdivision(tets, pts, feat, sig)
    out.update(sub_sig=sig, sub_sig_points=p_s, sub_sig_feat=f_s, sub_sig_tets=t_s)
    # ---- updategeometry pieces ----------------------------------------------------------------------------------------
    f3, ft2, fs2 = pw.tet_to_face_idx(P, tets, with_boundary=True)
    out.update(face_fx3=f3, face_tet_fx2=ft2, face_slot_fx2=fs2)
    _, nbr = ut.tet_adj_share(tets, P)
    out.update(tet_neighbour_idx=nbr)
    adj_idx, adj_sum = pw.generate_point_adj_idx(P, tets)
    out.update(point_adj_idx=adj_idx, point_adj_sum=adj_sum)
    # ---- deletetet chain (3_model/deftet.py:311-329) on the partially deleted mesh too ----------------------------------
    fake = types.SimpleNamespace(tet_neighbour_idx=nbr)
    r = np.linalg.norm(pts - np.array([[0.2, 0.1, -0.15]], dtype=np.float32), axis=1, keepdims=True)
    w = (1.0 / (1.0 + np.exp((r - 0.3) * 25.0)) * (0.5 + 0.5 * rng.rand(P, 1))).astype(np.float32)   # occupancy-like blob
    tw4 = rd.Deftet.pointweights2tetweights(fake, w, tets)
    for L in (1, 2, 3):
        twn = rd.Deftet.tetweights2tetneighbourweights(fake, tw4, neilevel=L)
        for thres in (0.05, 0.5):
            kept = pw.delete_tet(tets, twn, thres)
            out["del_L%d_t%03d" % (L, int(thres * 100))] = kept
    out.update(point_weights=w)
    # ---- get_featlap, get_volume_variance with autograd ---------------------------------------------------------------------
    fake = types.SimpleNamespace(tfpoint_adj_idx_pxm=torch.from_numpy(adj_idx) + 1, tfpoint_adj_weights_px1=torch.from_numpy(adj_sum) + 1e-10)
    x = torch.from_numpy(feat[:, :4]).clone().requires_grad_(True)
    lap = rd.Deftet.get_featlap(fake, x)
    gl = torch.from_numpy(rng.rand(P, 4).astype(np.float32))
    (lap * gl).sum().backward()
    out.update(featlap=lap.detach().numpy(), featlap_gout=gl.numpy(), featlap_gx=x.grad.numpy())
    mov = torch.zeros(P, 3, requires_grad=True)
    fake = types.SimpleNamespace(get_point=lambda: torch.from_numpy(pts) + mov, tftet_tx4=torch.from_numpy(tets))
    vv = rd.Deftet.get_volume_variance(fake)
    gv = torch.from_numpy(rng.rand(T).astype(np.float32))
    (vv * gv).sum().backward()
    out.update(volvar=vv.detach().numpy(), volvar_gout=gv.numpy(), volvar_gpoint=mov.grad.numpy())
    # ---- perspective + vertex2face + peel2mask -----------------------------------------------------------------------------------
    B = 2
    th = np.array([0.3, -1.1])
    rot = np.stack([np.array([[np.cos(t), 0, np.sin(t)], [0, 1, 0], [-np.sin(t), 0, np.cos(t)]]) for t in th]).astype(np.float32)
    campos = np.stack([rot[b].T @ np.array([0, 0, 4.0]) for b in range(B)]).astype(np.float32)
    proj = np.array([[2.0], [2.0], [-1.0]], dtype=np.float32)
    pw3 = torch.from_numpy(pts).unsqueeze(0).repeat(B, 1, 1)
    pcam, pimg = cam.perspective(pw3, [torch.from_numpy(rot), torch.from_numpy(campos), torch.from_numpy(proj)])
    out.update(cam_rot=rot, cam_pos=campos, cam_proj=proj, points_camera=pcam.numpy(), points_image=pimg.numpy())
    f3_t = torch.from_numpy(f3)
    out.update(face_cam_bxfx9=v2f.vertex2face(pcam, f3_t).numpy(), face_img_bxfx6=v2f.vertex2face(pimg, f3_t).numpy())
    ims = torch.from_numpy(rng.rand(B, 50, 6, 4).astype(np.float32))
    ims[:, :, 4:, :] = 0.0                                                   # void slots
    col, vis, _ = peel2mask(ims, None)
    out.update(peel_in=ims.numpy(), peel_color=col.numpy(), peel_vis=vis.numpy())
    np.savez_compressed(os.path.join(OUT, "diffrender_res8.npz"), **out)
    print("diffrender golden: P=%d T=%d E=%d faces=%d sub_all T=%d sub_sig T=%d max_deg=%d" % (
        P, T, edges.shape[0], f3.shape[0], t_all.shape[0], t_s.shape[0], adj_idx.shape[1]))
    for k in sorted(out):
        if k.startswith("del_"):
            print(k, out[k].shape)


if __name__ == "__main__":
    main()

"""X1 (a)/(c): execute the REFERENCE'S OWN, UNMODIFIED parallel.py::ParallelWrapper.forward (parallel.py:93-299) and the train-step
arithmetic of train_multigpu.py:236-273 on top of deftet_b200/dropin, with a small stub network in place of DeformableTetNetwork
(the encoder / decoders are out of scope).  Run as a subprocess by tests/test_gpu_reference_callers.py with
  cwd        = an unpacked copy of the reference's python (oracle/_ref/reference_py.zip)
  PYTHONPATH = deftet_b200/dropin : repo root : tests/x1/stubs : that copy
  DEFTET_REFERENCE_ROOT = that copy
argv: <res> <n_devices> <out.json>.  n_devices = 1: the losses and the gradient of step 1 are compared with the CPU oracle
pipeline (oracle/*.py restatements) here; n_devices = 2: the wrapper runs under nn.DataParallel exactly as train_multigpu.py:136-140
builds it and must reproduce the single-device losses and gradient."""
import json
import os
import sys

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

res, n_dev, out_path = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
REPO = os.environ["DEFTET_B200_REPO"]

# ---- the reference's imports, as train_multigpu.py:8-27 spells them ------------------------------------------------------
from layers.DefTet.deftet import DefTet                       # drop-in (deftet_b200.deftet.DefTet)
from parallel import ParallelWrapper                          # REFERENCE file, unmodified
from utils import tet_utils                                   # drop-in shadow + reference remainder
import utils.dataloder_helper as helpers                      # REFERENCE file (falls through)
from utils.matrix_utils import MySparse                       # REFERENCE file
import parallel as _parallel_mod
assert os.path.realpath(_parallel_mod.__file__).startswith(os.path.realpath(os.environ["DEFTET_REFERENCE_ROOT"])), _parallel_mod.__file__
import train_multigpu                                         # noqa: F401  the trainer module itself must import under the drop-in
import eval as _eval_mod                                      # noqa: F401  (eval.py:29,305 needs utils.mesh_utils.save_mesh)

sys.path.insert(0, REPO)
from deftet_b200.grid import acute_lattice_grid, write_tet_file
from tests.test_gpu_render import _icosphere

dev = torch.device("cuda:0")
torch.manual_seed(0)

# ---- Engine.__init__ state (train_multigpu.py:63-110), grid read through the reference's own reader -----------------------
root = os.path.join(os.path.dirname(out_path), "grid_root")
os.makedirs(os.path.join(root, "quartet", "meshes"), exist_ok=True)
g = acute_lattice_grid(res)
write_tet_file(os.path.join(root, "quartet", "meshes", "cube_%f_tet.tet" % (1.0 / res)), g.vertices, g.tets)
vertices_nx3, tetrahedron_fx4, mask = helpers.read_tetrahedron(res=res, root=root)
init_tet_pos = torch.from_numpy(vertices_nx3).to(dev) - 0.5
init_pos_mask = torch.from_numpy(mask).float().to(dev)
init_tet_fx4 = torch.from_numpy(tetrahedron_fx4).long().to(dev)
deftet = DefTet()
point_adj_sparse = MySparse(tet_utils.c_tet_to_adj_sparse(vertices_nx3, tetrahedron_fx4, normalize=True).to(dev))
tet_face_fx3, tet_facetet_idx_fx2, _, _ = tet_utils.tet_to_face(vertices_nx3.shape[0], tetrahedron_fx4)
tet_face_fx3 = torch.from_numpy(tet_face_fx3).long().cuda()
tet_face_tetidx_fx2 = torch.from_numpy(tet_facetet_idx_fx2).long().cuda()
inverse_v = nn.Parameter(deftet.tet_inverse_v(init_tet_pos, init_tet_fx4))
inverse_v.requires_grad = False
deftet.inverse_v = inverse_v.cuda()
V, T = init_tet_pos.shape[0], init_tet_fx4.shape[0]


class StubNetwork(nn.Module):
    """Stands in for layers/pc_model.py::DeformableTetNetwork: the four methods ParallelWrapper.forward calls."""

    def __init__(self):
        super().__init__()
        gen = torch.Generator().manual_seed(3)
        self.delta = nn.Parameter((torch.rand(V, 3, generator=gen) * 2 - 1) * (0.2 / res))
        self.occ_w = nn.Parameter(torch.tensor([0.1, 0.3, -0.2, 0.5]))

    def encode_inputs(self, x):
        # rounded to 1e-3 so that the GPU and the CPU-oracle reductions (different summation orders) give the SAME encoding and
        # therefore bit-identical vertex positions: the A4 gradient is discontinuous where a foot point crosses a triangle edge
        # (_back.cu:291-317 sends an edge hit's gradient to one vertex only), so positions one ulp apart can move a whole point's
        # contribution -- a 3 % difference in the surf-term gradient in this scene before the rounding was added
        return torch.round(x.mean(dim=1) * 1000.0) / 1000.0

    def decode_pos(self, init_tet_pos_bxnx3, z, encoding, init_pos_mask, cam_pos=None, cam_rot=None, cam_proj=None):
        delta = self.delta.unsqueeze(0).expand(init_tet_pos_bxnx3.shape[0], -1, -1) * (1.0 + 0.05 * encoding[:, :1].unsqueeze(-1))
        if init_pos_mask is not None:
            delta = delta * init_pos_mask
        self._last_tet_pos = init_tet_pos_bxnx3 + delta
        return delta, self._last_tet_pos, delta

    def _logits(self, tet_pos, init_tet_bxfx4):
        B = tet_pos.shape[0]
        cen = tet_pos[:, init_tet_bxfx4[0].reshape(-1)].reshape(B, -1, 4, 3).mean(dim=2)
        return self.occ_w[0] + (cen * self.occ_w[1:]).sum(dim=-1) * 8.0

    def split_decode_occ(self, tet_pos, z, encoding, init_tet_bxfx4, cam_pos=None, cam_rot=None, cam_proj=None):
        return torch.sigmoid(self._logits(tet_pos, init_tet_bxfx4))

    def decode_occ(self, tet_pos, z, encoding, init_tet_bxfx4, cam_pos=None, cam_rot=None, cam_proj=None):
        idx = torch.arange(0, T, 5, device=tet_pos.device)
        return torch.distributions.Bernoulli(logits=self._logits(tet_pos, init_tet_bxfx4)[:, idx]), idx


model = StubNetwork().to(dev)
device_count = n_dev
wrapper = ParallelWrapper(model, deftet, os.path.dirname(out_path), point_adj_sparse, device_count, timing=None, use_two_encoder=False,
                          add_input_noise=False, n_point=500, use_lap_layer=False, use_point=True)
if device_count > 1:
    wrapper = nn.DataParallel(wrapper, device_ids=list(range(device_count)))      # train_multigpu.py:136-140

# ---- one synthetic batch shaped like dataloader.py's (verts / faces lists, sample_points, sdf_point) -----------------------
B, S, P = int(os.environ.get("X1_FORCE_BATCH", 2 * n_dev)), 6000, 2000
v_ico, f_ico = _icosphere(3)
radii = [0.27, 0.33, 0.22, 0.30][:B]
gen = torch.Generator().manual_seed(1)
data_verts = [torch.from_numpy(v_ico * r).float() for r in radii]
data_faces = [torch.from_numpy(f_ico).long() for _ in radii]
d = torch.randn(B, S, 3, generator=gen)
surface_point = (d / d.norm(dim=-1, keepdim=True) * torch.tensor(radii).reshape(B, 1, 1)).to(dev)
points = ((torch.rand(B, P, 3, generator=gen) - 0.5) * 1.05).to(dev)
all_verts = [v.to(dev).unsqueeze(0).expand(device_count, -1, -1) for v in data_verts]       # train_multigpu.py:174-177
all_faces = [v.to(dev).unsqueeze(0).expand(device_count, -1, -1) for v in data_faces]
LAMBDA = dict(area=1e6, edge=1.0, lap=0.5, surf=2.0, delta=0.1, normal=0.3, amips=1.0, surf_chamfer=1.0, lap_v=0.0, occ=1.0, deform=1.0)
optimizer = torch.optim.Adam([p for p in model.parameters() if p.requires_grad], lr=1e-4)
SEED = 1234
record = {"res": res, "n_dev": n_dev, "V": V, "T": T, "steps": []}
grad_step0 = None
for step in range(2):
    optimizer.zero_grad()
    for k in range(torch.cuda.device_count()):
        with torch.cuda.device(k):
            torch.cuda.manual_seed(SEED + step)
    expand = lambda t: t.unsqueeze(0).expand(B, *([-1] * t.dim()))
    out = wrapper(imgs=None, init_tet_pos_bxnx3=expand(init_tet_pos.float()), init_tet_bxfx4=expand(init_tet_fx4), points=points,
                  surface_point=surface_point, save=False, global_step=step, tet_face_tetidx_bxfx2=expand(tet_face_tetidx_fx2),
                  all_verts=all_verts, all_faces=all_faces, return_all=False, tet_face_bxfx3=expand(tet_face_fx3),
                  init_pos_mask=expand(init_pos_mask.float()), cam_pos=None, cam_rot=None, cam_proj=None, pred_threshold=0.4)
    amips_energy, edge, area_variance, surface_align, normal_loss, occ_loss, lap, delta_loss, other_chamfer_distance, lap_v_loss = out
    # train_multigpu.py:236-262
    terms = dict(surf=surface_align.mean(), area=area_variance.mean(), normal=normal_loss.mean(), edge=edge.mean(), amips=amips_energy.mean(),
                 surf_chamfer=other_chamfer_distance.mean(), lap_v=lap_v_loss.mean(), lap=lap.mean(), delta=delta_loss.mean())
    occ = occ_loss.mean()
    deform_loss = sum(terms[k] * LAMBDA[k] for k in terms)
    loss = occ * LAMBDA["occ"] + deform_loss * LAMBDA["deform"]
    if step == 0 and n_dev == 1:          # per-term gradients w.r.t. the shared vertex offsets (diagnostics for the oracle comparison)
        term_grads = {k: (torch.autograd.grad(v, model.delta, retain_graph=True, allow_unused=True)[0] if v.requires_grad else None)
                      for k, v in terms.items()}
        term_grads = {k: (torch.zeros_like(model.delta) if v is None else v).detach().cpu() for k, v in term_grads.items()}
        wrapper_pos = model._last_tet_pos.detach().cpu().clone()
        wrapper_pos_grad_surf = torch.autograd.grad(terms["surf"], model._last_tet_pos, retain_graph=True)[0].detach().cpu().clone()
    loss.backward()
    vals = {k: float(v) for k, v in terms.items()}
    vals.update(occ=float(occ), loss=float(loss))
    assert all(np.isfinite(list(vals.values()))), vals
    assert torch.isfinite(model.delta.grad).all() and float(model.delta.grad.abs().max()) > 0
    if step == 0:
        grad_step0 = (model.delta.grad.detach().cpu().clone(), model.occ_w.grad.detach().cpu().clone())
        delta0 = model.delta.detach().cpu().clone()
        occw0 = model.occ_w.detach().cpu().clone()
    record["steps"].append(vals)
    optimizer.step()
torch.save({"grad_delta": grad_step0[0], "grad_occ_w": grad_step0[1], "delta0": delta0, "occ_w0": occw0}, out_path + ".pt")

# ---- oracle pipeline for step 0 (CPU restatements, test infrastructure), single-device run only ---------------------------
if n_dev == 1:
    from oracle import builders as orc_b, energies as orc_e, native as orc, surface as orc_s
    pos0 = (torch.from_numpy(vertices_nx3) - 0.5).float()          # the trainer's arithmetic (train_multigpu.py:66): subtract in the file's dtype, then .float()
    tet = torch.from_numpy(tetrahedron_fx4).long()
    sp, pm = surface_point.cpu(), init_pos_mask.cpu()
    delta_p = delta0.clone().requires_grad_(True)
    occ_w = occw0.clone().requires_grad_(True)
    enc = torch.round(sp[:, :500].mean(dim=1) * 1000.0) / 1000.0
    delta = delta_p.unsqueeze(0).expand(B, -1, -1) * (1.0 + 0.05 * enc[:, :1].unsqueeze(-1)) * pm
    tet_pos = pos0.unsqueeze(0) + delta
    # The A4 gradient of the reference is discontinuous in the positions (an edge hit sends its whole gradient to ONE vertex of
    # ONE of the two faces that tie on that edge, _back.cu:291-317), so the comparison needs bit-identical positions on both sides
    record["positions_bit_identical"] = bool((wrapper_pos == tet_pos.detach()).all())
    soup = orc_e.gather_tets(tet_pos, tet)
    cen = soup.mean(dim=2)
    ref_occ = np.stack([orc.check_sign(data_verts[b].unsqueeze(0).numpy(), f_ico, cen[b].detach().unsqueeze(0).numpy())[0] for b in range(B)]).astype(np.float32)
    f3, ft2, _, _ = orc_b.tet_to_face(V, tetrahedron_fx4)
    bnd = orc_s.get_boundary_index(torch.from_numpy(f3), torch.from_numpy(ft2), torch.from_numpy(ref_occ))
    Fmax = max(int(b.shape[0]) for b in bnd)
    torch.cuda.manual_seed(SEED)
    u_all = torch.sqrt(torch.rand(size=(B, Fmax, 20), device=dev)).cpu()
    v_all = torch.rand(size=(B, Fmax, 20), device=dev).cpu()
    u_list = [u_all[b, :bnd[b].shape[0]].reshape(1, -1, 20, 1) for b in range(B)]
    v_list = [v_all[b, :bnd[b].shape[0]].reshape(1, -1, 20, 1) for b in range(B)]
    inv = orc_e.tet_inverse_v(pos0, tet)
    ch, an, nl = orc_s.surface_losses(tet_pos, bnd, sp, u_list, v_list)
    idx = torch.arange(0, T, 5)
    logits = (occ_w[0] + (cen * occ_w[1:]).sum(dim=-1) * 8.0)[:, idx]
    o_terms = dict(surf=an.mean(), area=orc_e.volume_variance(soup).mean(), normal=nl.mean(), edge=orc_e.edge_length(soup).mean(),
                   amips=orc_e.amips_energy(soup, inv).mean(), surf_chamfer=ch.mean(), lap_v=torch.zeros(()),
                   delta=torch.mean(torch.abs(delta), dim=-1).mean(dim=-1).mean())
    # Laplacian (deftet.py:340-343) with the row-normalised adjacency rebuilt from the oracle's directed edge list
    e2 = torch.from_numpy(np.asarray(orc_b.tet_to_adj_edges(tetrahedron_fx4))).long()
    deg = torch.zeros(V).index_add_(0, e2[:, 0], torch.ones(e2.shape[0]))
    nei = torch.zeros(B, V, 3).index_add_(1, e2[:, 0], delta[:, e2[:, 1]]) / deg.reshape(1, V, 1)
    o_terms["lap"] = ((nei - delta) ** 2).sum(dim=-1).sum(dim=-1).mean()
    o_occ = F.binary_cross_entropy_with_logits(logits, torch.from_numpy(ref_occ)[:, idx]).mean()
    o_loss = o_occ * LAMBDA["occ"] + sum(o_terms[k] * LAMBDA[k] for k in o_terms) * LAMBDA["deform"]
    rel = lambda a, b: float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp(min=1e-30))
    record["term_grad_rel_err"] = {}
    o_term_grads = {}
    for k, v in o_terms.items():
        if not v.requires_grad:
            continue
        og = torch.autograd.grad(v, delta_p, retain_graph=True, allow_unused=True)[0]
        if og is not None and float(og.abs().max()) > 0:
            o_term_grads[k] = og
            record["term_grad_rel_err"][k] = rel(term_grads[k], og)
    o_loss.backward()
    record["oracle"] = {k: float(v) for k, v in o_terms.items()}
    record["oracle"].update(occ=float(o_occ), loss=float(o_loss))
    record["grad_rel_err_delta"] = rel(grad_step0[0], delta_p.grad)
    record["grad_rel_err_occ_w"] = rel(grad_step0[1], occ_w.grad)
    record["loss_rel_err"] = {k: abs(record["steps"][0][k] - record["oracle"][k]) / max(abs(record["oracle"][k]), 1e-30) for k in record["oracle"]}
json.dump(record, open(out_path, "w"))
print("X1 parallel ok", json.dumps(record)[:600])

if n_dev == 1 and os.environ.get("X1_DIAG"):          # dev diagnostic: A4 of this scene, engine op vs oracle, per sample / vertex / point
    from deftet_b200 import surface
    tp = tet_pos.detach().to(dev)
    for b in range(B):
        faces_b = bnd[b].to(dev).int().unsqueeze(0).contiguous()
        cnt = torch.tensor([faces_b.shape[1]], dtype=torch.int32, device=dev)
        gtb = sp[b:b + 1].contiguous()
        soup, cd, cf = surface.closest_faces(tp[b:b + 1], faces_b, cnt, gtb.to(dev))
        d_o, f_o = orc.point_face_distance(gtb.numpy(), soup.cpu().numpy())
        print("DIAG sample", b, "F", int(cnt), "fwd d equal", int((cd.cpu().numpy().reshape(-1) == d_o.reshape(-1)).sum()), "of", S,
              "face equal", int((cf.cpu().numpy().reshape(-1) == f_o.reshape(-1)).sum()))
        pe = tp[b:b + 1].clone().requires_grad_(True)
        surface.surface_distance(pe, faces_b, cnt, gtb.to(dev)).sum().backward()
        po = tet_pos[b:b + 1].detach().clone().requires_grad_(True)
        orc_s.point_mesh_distance(gtb, orc_s.gather_faces(po, bnd[b])).mean(-1).mean(-1).sum().backward()
        e = (pe.grad.cpu() - po.grad).abs().max(dim=-1).values[0]
        wv = int(torch.argmax(e))
        print("  bwd max abs err %.4g at vertex %d (scale %.4g): ours %s oracle %s" % (float(e.max()), wv, float(po.grad.abs().max()),
              pe.grad[0, wv].cpu().numpy(), po.grad[0, wv].numpy()))
        # per point: faces touching the worst vertex
        fsel = torch.nonzero((bnd[b] == wv).any(dim=1)).reshape(-1)
        pts_sel = torch.nonzero(torch.isin(torch.from_numpy(f_o.reshape(-1)).long(), fsel)).reshape(-1)
        nshow = 0
        for i in pts_sel.tolist():
            f = int(f_o.reshape(-1)[i])
            tri = soup[0:1, f:f + 1].contiguous()
            p1 = gtb[:, i:i + 1].contiguous()
            dfa = tri.clone().requires_grad_(True)
            d2, _ = surface.tet_analytic_distance_f_batch(p1.to(dev), dfa, torch.tensor([1.0], device=dev))
            d2.sum().backward()
            go = orc.point_face_distance_bwd(p1.numpy(), tri.cpu().numpy(), np.zeros((1, 1, 1), np.float32), np.ones((1, 1, 1), np.float32))[0, 0]
            if np.abs(dfa.grad[0, 0].cpu().numpy() - go).max() > 1e-6 * max(np.abs(go).max(), 1e-12):
                print("   pt", i, "face", f, "drop-in grad", dfa.grad[0, 0].cpu().numpy().reshape(-1), "oracle", go.reshape(-1)); nshow += 1
                if nshow > 3: break
        print("  per-point drop-in backward mismatches among %d points: %d" % (len(pts_sel), nshow))
    # batched padded-ragged call, as DefTet.forward_surface_align issues it
    Fm = max(int(x.shape[0]) for x in bnd)
    fpad = torch.zeros(B, Fm, 3, dtype=torch.int32)
    for b in range(B):
        fpad[b, :bnd[b].shape[0]] = bnd[b].int()
    cnts = torch.tensor([int(x.shape[0]) for x in bnd], dtype=torch.int32, device=dev)
    pe = tp.clone().requires_grad_(True)
    lb = surface.surface_distance(pe, fpad.to(dev).contiguous(), cnts, sp.to(dev))
    lb.mean().backward()
    po = tet_pos.detach().clone().requires_grad_(True)
    _, an2, _ = orc_s.surface_losses(po, bnd, sp, u_list, v_list)
    an2.mean().backward()
    print("DIAG batched: loss ours", lb.tolist(), "oracle", an2.tolist(), "grad rel err", rel(pe.grad.cpu(), po.grad))
    # and through the drop-in module with ITS boundary faces
    ctr = tp[:, init_tet_fx4.reshape(-1)].reshape(B, -1, 4, 3).mean(dim=2)
    occ_d = deftet.check_tet_inside_sdfs(None, [[v.to(dev).unsqueeze(0) for v in data_verts], [f.to(dev).unsqueeze(0) for f in data_faces]], centers=ctr)
    print("DIAG occupancy equal to oracle:", bool((occ_d.squeeze(-1).cpu() == torch.from_numpy(ref_occ)).all()))
    table = deftet._table(tet_face_fx3, tet_face_tetidx_fx2)
    fc, cn, _ = surface.boundary_faces(table, occ_d.squeeze(-1), table.n_face)
    print("DIAG boundary counts", cn.tolist(), "faces equal to oracle list:",
          [bool((fc[b, :int(cn[b])].cpu().long() == bnd[b]).all()) if int(cn[b]) == bnd[b].shape[0] else False for b in range(B)])
    # the drop-in module call itself on the same positions
    px = tp.clone().requires_grad_(True)
    expand = lambda t: t.unsqueeze(0).expand(B, *([-1] * t.dim()))
    mesh_list = [[v.to(dev).unsqueeze(0) for v in data_verts], [f.to(dev).unsqueeze(0) for f in data_faces]]
    outs = deftet.forward_surface_align(px, points, expand(init_tet_fx4), mesh_list, gt_surface_points=surface_point,
                                        tet_face_tet_bx4fx2=expand(tet_face_tetidx_fx2), inference=False, tet_face_bxfx3=expand(tet_face_fx3))
    g_mod = torch.autograd.grad(outs[3].mean(), px, retain_graph=True)[0]
    print("DIAG module surf loss", float(outs[3].mean()), "vs batched op", float(lb.mean()), "grad rel err module vs op", rel(g_mod.cpu(), pe.grad.cpu()),
          "module vs oracle", rel(g_mod.cpu(), po.grad))
    scale = ((1.0 + 0.05 * enc[:, :1].unsqueeze(-1)) * pm)
    chain = (po.grad * scale).sum(dim=0)
    print("DIAG chain-ruled oracle grad vs harness oracle term grad:", rel(chain, o_term_grads["surf"]),
          "| harness GPU term grad vs chain-ruled module grad:", rel(term_grads["surf"], (g_mod.cpu() * scale).sum(dim=0)))
    print("DIAG wrapper positions == oracle positions bitwise:", bool((wrapper_pos == tet_pos.detach()).all()), "max abs diff", float((wrapper_pos - tet_pos.detach()).abs().max()))
    print("DIAG wrapper d surf / d pos vs module:", rel(wrapper_pos_grad_surf, g_mod.cpu()), "vs oracle", rel(wrapper_pos_grad_surf, po.grad))
    e = (wrapper_pos_grad_surf - po.grad).abs().max(dim=-1).values
    for b in range(B):
        w = torch.argsort(e[b], descending=True)[:4]
        print("   sample", b, "worst vertices", w.tolist(), "err", e[b][w].tolist(), "wrapper", wrapper_pos_grad_surf[b][w[0]].tolist(), "oracle", po.grad[b][w[0]].tolist())

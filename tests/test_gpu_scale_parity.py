"""Parity at the BASELINE.json configurations: the engine's batched kernels vs the reference's OWN
device kernels (oracle/_ref/kernels_cuda, brute force) on the bench workload -- res 40 batch 8 and res 70 batch 8 with
P = S = 100 000 points per sample, the scene generator bench.py uses.  Binning resolution, brick staging and the
always-test lists change regime with size, so the small-grid parity tests do not cover this.  tools/parity_check.py
states the contract (bit-exact ids up to proven ties, floats within 1e-5 relative)."""
import pytest
import torch

from oracle import ref_cuda
from tools.parity_check import verify_scene

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_cuda.available(), reason="oracle/_ref/kernels_cuda not built (needs /root/reference at build time)")]


def _run(res, B, P, S, seed, shapes=(1, 3), samples=None):
    from deftet_b200.engine import GeometryEngine
    from deftet_b200.grid import acute_lattice_grid
    from deftet_b200.synthetic import analytic_scene
    dev = torch.device("cuda:0")
    grid = acute_lattice_grid(res)
    sc = analytic_scene(grid, B, P, S, seed, dev, shapes=shapes)
    Fmax = 16384 if shapes[1] <= 3 else 32768
    eng = GeometryEngine(grid.centred(), grid.tets, max_boundary_faces=Fmax, device=dev)
    gen = torch.Generator(device=dev).manual_seed(seed)
    u = torch.sqrt(torch.rand(B, Fmax, 20, device=dev, generator=gen))
    v = torch.rand(B, Fmax, 20, device=dev, generator=gen)
    rep = verify_scene(eng, sc, u, v, samples=samples)
    print("parity res=%d B=%d: %s" % (res, B, rep))
    assert rep["ok"]
    return rep


def test_config2_res40_batch8():
    """BASELINE.json configs[1]: res 40, batch 8."""
    _run(40, 8, 100000, 100000, 2000)


def test_config3_res70_batch8():
    """BASELINE.json configs[2]: res 70, batch 8 -- the bench workload, same seed as bench.py's input set 0 on rank 0."""
    _run(70, 8, 100000, 100000, 3000)


def test_config3_res70_many_shapes():
    """F_b sweep end of the workload (about 10 k boundary faces per sample, SURVEY.md 8d 'ShapeNet is 2-4x')."""
    _run(70, 2, 100000, 100000, 3100, shapes=(40, 48))

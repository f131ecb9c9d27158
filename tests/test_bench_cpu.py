"""Host-side logic of bench.py and the binding layer that needs no GPU: the CPU reference arm's bookkeeping (bench.py --impl reference),
the config table, the alignment helper that nn.DataParallel's coalesced parameter views need."""
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_aligned_copies_only_misaligned_views():
    from deftet_b200 import _lib
    buf = torch.zeros(64)
    a = buf[4:36]                                       # 16-byte offset: already aligned
    b = buf[3:35]                                       # 12-byte offset
    assert _lib.aligned(a).data_ptr() == a.data_ptr()
    out = _lib.aligned(b)
    assert out.data_ptr() % 16 == 0 and out.data_ptr() != b.data_ptr() and torch.equal(out, b)
    nc = torch.arange(24.0).reshape(4, 6).t()           # non-contiguous: made contiguous + aligned
    out = _lib.aligned(nc)
    assert out.is_contiguous() and out.data_ptr() % 16 == 0 and torch.equal(out, nc)


def test_reference_arm_reports_a_measured_bounded_sample():
    """`bench.py --impl reference`: ms_per_step must be the MEASURED wall time of the bounded sample (the round-1 arm claimed an
    extrapolated step time), the extrapolation is a separate, labelled field, and the run ends quickly whatever --steps says."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--res", "8", "--points", "3000", "--steps", "50"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "tets/ms" and line["steps"] <= 3 and line["sample_is_bounded"] is True
    assert line["gpu_launches"] == 0 and line["e2e"]["h2d_bytes_per_step"] == 0
    cb = line["cpu_baseline"]
    assert cb["cores"] >= 1 and "A1" in cb["kind"] and "extrapolated" in cb["kind"]
    assert 0 < line["ms_per_step"] < 120e3 and line["full_step_estimate_ms"] >= 0.5 * line["ms_per_step"]
    B, T = line["config"]["global_batch"], int(line["config"]["grid"].split("T=")[1])
    assert abs(line["value"] - B * T / line["full_step_estimate_ms"]) <= 1e-6 * line["value"]


def test_reference_arm_declines_configs_it_is_not_defined_for():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "5"], capture_output=True, text=True,
                       timeout=300, cwd=ROOT)
    assert r.returncode == 0
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and "unavailable" in line


def test_step_terms_follow_the_requested_loss_groups():
    import bench

    class _Eng:                                          # Step only needs these two attributes at construction time
        n_vert, device = 10, torch.device("cpu")

    full = bench.Step(_Eng(), None, 16, 20)
    assert full.names == ["amips", "edge", "volume_variance", "chamfer", "distance", "normal", "occupancy"]
    c2 = bench.Step(_Eng(), None, 16, 20, want=("energies", "occupancy"), loss_scale=0.5)
    assert c2.names == ["amips", "edge", "volume_variance", "occupancy"]
    assert torch.allclose(c2.wvec.reshape(-1), torch.tensor([0.5, 0.5, 0.5e6, 0.5]))

"""The C-ABI library loads and exports every symbol include/deftet_b200.h declares (no compute, no GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "deftet_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dtb_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_hot_path():
    syms = declared_symbols()
    for must in ["dtb_point_in_tet_soup", "dtb_tet_barycentric_backward", "dtb_nearest_neighbor", "dtb_point_face_distance_forward",
                 "dtb_point_face_distance_backward", "dtb_face_adjacency", "dtb_tet_energies_forward", "dtb_tet_energies_backward",
                 "dtb_boundary_faces", "dtb_tet_point_adj", "dtb_tet_to_face", "dtb_tet_adj_share", "dtb_tet_face_adj", "dtb_collapse_vertices",
                 "dtb_host_tet_point_adj", "dtb_host_tet_adj_share", "dtb_host_tet_face_adj", "dtb_host_colaps_v", "dtb_sparse_render_forward",
                 "dtb_sparse_render_backward", "dtb_check_sign", "dtb_laplacian_forward"]:
        assert must in syms


def test_library_exports_every_declared_symbol():
    path = os.path.join(ROOT, "deftet_b200", "libdeftet_b200.so")
    if not os.path.exists(path):
        pytest.fail("libdeftet_b200.so is not built: run `python -c 'import __graft_entry__ as g; g.build()'`")
    lib = ctypes.CDLL(path)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, "declared but not exported: %s" % missing
    lib.dtb_version.restype = ctypes.c_int
    assert lib.dtb_version() >= 100
    lib.dtb_last_error.restype = ctypes.c_char_p
    assert isinstance(lib.dtb_last_error(), bytes)


def test_python_binding_declares_signatures_and_has_no_cpu_fallback():
    import torch
    from deftet_b200 import _lib, builders, energies, render, search, surface  # noqa: F401
    L = _lib.lib()
    assert L.dtb_nearest_neighbor_grid_res(100000) % 4 == 0
    assert L.dtb_tet_energies_workspace(8, 100, 1000) >= 8 * 1000 * 4
    with pytest.raises(_lib.DeftetB200Error):
        search.point_in_tet_soup(torch.zeros(1, 4, 4, 3), torch.zeros(1, 5, 3))      # CPU tensors are refused, never emulated
    with pytest.raises(_lib.DeftetB200Error):
        energies.tet_energies(torch.zeros(1, 4, 3), torch.zeros(1, 4, dtype=torch.int32), torch.zeros(1, 3, 3))


def test_widening_rows_refuse_cpu_tensors_and_size_their_workspaces():
    """N2 / N3 / N4 host mirrors: same rule as the hot path (no CPU emulation); host-only sizing functions are callable without a GPU."""
    import torch
    from deftet_b200 import _lib, devox, graph, metrics, topology  # noqa: F401
    L = _lib.lib()
    with pytest.raises(_lib.DeftetB200Error):
        devox.trilinear_devoxelize(torch.zeros(1, 4, 8, 8, 8), torch.zeros(1, 3, 5), 8)
    with pytest.raises(_lib.DeftetB200Error):
        devox.sample_f(torch.zeros(1, 5, 3), [torch.zeros(1, 4, 8, 8, 8)])
    with pytest.raises(_lib.DeftetB200Error):
        metrics.sided_distance(torch.zeros(1, 5, 3), torch.zeros(1, 6, 3))
    # volume gradient of trilinear_devoxelize: the sorted reduction asks for temporary memory from 8 points per voxel on
    ws = L.dtb_trilinear_devoxelize_backward_workspace
    assert ws(8, 512, 45684, 8, 0) >= 8 * 45684 * 24
    assert ws(8, 64, 45684, 32, 0) == 0                      # 1.4 points per voxel: shared-atomic kernel, no workspace
    assert ws(8, 64, 45684, 32, devox.FORCE_SORT) > 0
    assert ws(8, 512, 45684, 8, devox.NO_SORT) == 0
    assert ws(0, 512, 45684, 8, 0) == 0


def test_run_so_shims_export_run():
    for name in ("tet_point_adj", "tet_adj_share", "tet_face_adj", "colaps_v"):
        path = os.path.join(ROOT, "deftet_b200", "dropin", "utils", "lib", name, "run.so")
        assert os.path.exists(path), "shim %s missing (make shims)" % path
        lib = ctypes.CDLL(path)
        assert hasattr(lib, "run")


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "deftet_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".c", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "liboracle" not in src, os.path.join(dirpath, f)


def test_reference_device_kernels_are_built_and_export_their_launchers():
    """oracle/_ref/kernels_cuda (the reference's own __global__ kernels compiled for sm_100a; test / bench baseline only)."""
    names = {"point_in_tet": "refcuda_point_in_tet", "nearest_neighbor": "refcuda_nearest_neighbor",
             "face_distance_fwd": "refcuda_point_face_distance", "face_distance_bwd": "refcuda_point_face_distance_bwd",
             "face_adj": "refcuda_face_adjacency"}
    d = os.path.join(ROOT, "oracle", "_ref", "kernels_cuda")
    if not os.path.isdir(d):
        pytest.skip("oracle/_ref/kernels_cuda not built (needs /root/reference at build time)")
    for name, sym in names.items():
        lib = ctypes.CDLL(os.path.join(d, name + ".so"))
        assert hasattr(lib, sym), (name, sym)

"""CPU-only checks of the product's host side: tables, C ABI surface, drop-in modules' argument
handling.  No compute call is made here (the product has no CPU path)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import PKG, ROOT


def test_product_tables_equal_oracle_tables(oracle):
    from hipr_b200 import tables
    for P, R in [(11, 9), (7, 5), (11, 4), (15, 9), (11, 12), (5, 3), (9, 7), (13, 16), (3, 2), (21, 13)]:
        assert np.array_equal(tables.line_table_2d(P, R), oracle.line_table_2d(P, R).transpose(2, 0, 1))
    for P, TH, PH in [(11, 9, 9), (7, 5, 4), (9, 6, 7), (11, 4, 12), (5, 3, 3)]:
        assert np.array_equal(tables.line_table_3d(P, TH, PH), oracle.line_table_3d(P, TH, PH).transpose(2, 0, 1))
        assert np.array_equal(tables.line_table_3d_v3(P, TH, PH), oracle.line_table_3d_v3(P, TH, PH).transpose(2, 0, 1))
    t = tables.line_table_2d(11, 9)
    assert t.dtype == np.int32 and t.flags.c_contiguous and t.shape == (9, 11, 2)
    assert tables.line_table_3d(11, 9, 9).shape == (72, 11, 3)


def test_product_tables_are_pinned(monkeypatch):
    from hipr_b200 import tables
    monkeypatch.setitem(tables.PINNED, ("2d", 11, 9), "0" * 64)
    with pytest.raises(RuntimeError, match="pinned"):
        tables.line_table_2d(11, 9)


def test_baked_tables_header_is_current():
    """csrc/baked_tables.cuh and csrc/sortnet_gen.cuh are what their generators emit."""
    for gen, out in (("gen_tables.py", "baked_tables.cuh"), ("gen_sortnet.py", "sortnet_gen.cuh")):
        res = subprocess.run([sys.executable, os.path.join(PKG, "csrc", gen)], capture_output=True, text=True, check=True)
        assert res.stdout == open(os.path.join(PKG, "csrc", out)).read()


def test_table_argument_validation():
    from hipr_b200 import tables
    with pytest.raises(ValueError):
        tables.line_table_2d(10, 9)
    with pytest.raises(ValueError):
        tables.line_table_2d(33, 9)
    with pytest.raises(ValueError):
        tables.line_table_2d(11, 0)
    with pytest.raises(TypeError):
        tables.line_table_2d("11", 9)
    assert np.array_equal(tables.line_table_2d(11.0, 9), tables.line_table_2d(11, 9))   # as the Cython `int` args


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "hipr_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hipr_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    """libhipr_b200.so loads without a GPU and exports exactly what include/hipr_b200.h declares."""
    import hipr_b200
    declared = _declared_functions()
    assert len(declared) >= 25
    assert os.path.exists(hipr_b200.LIB_PATH), "build with python hiprfish-image-analysis_b200/build.py"
    handle = ctypes.CDLL(hipr_b200.LIB_PATH)
    for name in declared:
        assert hasattr(handle, name), name
    assert sorted(hipr_b200.EXPORTS) == declared            # the Python binding covers the whole header
    lib = hipr_b200.lib()
    assert lib.hipr_abi_version() == 1
    assert b"patch_size" in lib.hipr_error_string(-3)
    assert lib.hipr_error_string(0) == b"ok"


def test_library_is_sm100a_native_code():
    import hipr_b200
    out = subprocess.run(["cuobjdump", "-lelf", hipr_b200.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump not available")
    archs = set(re.findall(r"sm_\d+a?", out.stdout))
    assert archs == {"sm_100a"}, archs


def test_missing_library_fails_loudly(monkeypatch):
    from hipr_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libhipr_b200.so")
    with pytest.raises(ImportError, match="no CPU fallback"):
        _lib.lib()


def test_dropin_modules_surface_and_errors():
    import neighbor
    import neighbor2d
    assert callable(neighbor2d.line_profile_2d_v2)
    for name in ("line_profile_v2", "line_profile_memory_efficient_v2", "line_profile_memory_efficient_v3",
                 "line_profile", "neighbor_average"):
        assert callable(getattr(neighbor, name))
    # same exception classes and messages as the Cython typed-memoryview arguments
    with pytest.raises(ValueError, match="Buffer dtype mismatch, expected 'double' but got 'float'"):
        neighbor2d.line_profile_2d_v2(np.zeros((12, 12), np.float32), 11, 9)
    with pytest.raises(ValueError, match=r"wrong number of dimensions \(expected 2, got 3\)"):
        neighbor2d.line_profile_2d_v2(np.zeros((12, 12, 2)), 11, 9)
    with pytest.raises(ValueError, match=r"wrong number of dimensions \(expected 3, got 2\)"):
        neighbor.line_profile_v2(np.zeros((12, 12)), 11, 9, 9)
    with pytest.raises(TypeError):
        neighbor2d.line_profile_2d_v2(np.zeros((12, 12)), "11", 9)
    with pytest.raises(ValueError, match="expected 'float' but got 'double'"):
        neighbor.neighbor_average(np.zeros((30, 30, 30)), 11)
    with pytest.raises(ValueError, match="expected 'float' but got 'double'"):
        neighbor.neighbor_average(np.zeros((30, 30, 30), np.float32), 11)


def test_no_cpu_path():
    """Without a GPU the operators refuse; they never fall back to the oracle or numpy."""
    import torch
    import hipr_b200
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(ValueError, match="no CPU path"):
        hipr_b200.lne2d(torch.zeros(20, 20), "F1")
    with pytest.raises(ValueError, match="no CPU path"):
        hipr_b200.cell_spectra(torch.zeros(4, 4, 3), torch.zeros(4, 4, dtype=torch.int32), 1)
    for call in (lambda: hipr_b200.register_stacks([torch.zeros(8, 8, 3)]),
                 lambda: hipr_b200.denoise_nl_means(torch.zeros(40, 40, dtype=torch.float64), h=0.02),
                 lambda: hipr_b200.cell_geometry(torch.zeros(8, 8, dtype=torch.int32)),
                 lambda: hipr_b200.paint_labels(torch.zeros(8, 8, dtype=torch.int32), torch.zeros(3)),
                 lambda: hipr_b200.channel_sum_raw(torch.zeros(8, 8, 3, dtype=torch.uint8), 255.0),
                 lambda: hipr_b200.neighbor3d_score(torch.zeros(12, 12, 12, 4), "ME2")):
        with pytest.raises(ValueError, match="no CPU path"):
            call()
    import neighbor2d
    with pytest.raises(Exception):
        neighbor2d.line_profile_2d_v2(np.zeros((12, 12)), 11, 9)


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("the oracle", "").replace("oracle/", "ORACLE_DIR/") or f == "ops.py", f
                assert "import oracle" not in src and "from oracle" not in src, f


def test_synthetic_fov_is_deterministic_and_documented():
    from hipr_b200 import synth
    a, la, L = synth.make_fov(64, 96, 95, fov_index=3)
    b, lb, _ = synth.make_fov(64, 96, 95, fov_index=3)
    assert a.dtype.is_floating_point and a.shape == (64, 96, 95) and la.shape == (64, 96)
    assert (a == b).all() and (la == lb).all()
    assert int(la.max()) <= L and int(la.min()) == 0
    c, lc, _ = synth.make_fov(64, 96, 95, fov_index=4)
    assert not (a == c).all()
    ld, _ = synth.make_labels(128, 192, drop_fraction=0.3)
    assert len(ld.unique()) - 1 < (128 // 26) * (192 // 44)


def test_mosaic_peer_buffer_layout_and_argument_checks():
    """Host-only entry points of the peer-memory mosaic exchange: buffer size arithmetic and argument validation
    (no device call is made)."""
    import hipr_b200
    lib = hipr_b200.lib()
    rows, W, world = 2048, 16384, 8
    ext = 2 * (rows + 10) * W * 8                  # two parities of the extended float64 sum image
    keys = 2 * world * 2 * 8
    flags = world * 8
    want = -(-(ext + keys + flags) // 256) * 256
    assert lib.hipr_mosaic_p2p_bytes(rows, W, world) == want
    assert lib.hipr_mosaic_p2p_bytes(4, W, world) < 0          # a slab must hold the 5 halo rows it sends
    assert lib.hipr_mosaic_p2p_bytes(rows, W, 65) < 0
    assert lib.hipr_p2p_alloc(None, 1024) == -1
    assert lib.hipr_p2p_get_handle(None, None) == -1
    assert lib.hipr_p2p_open_handle(None, None) == -1
    assert lib.hipr_p2p_free(None) == 0 and lib.hipr_p2p_close_handle(None) == 0
    assert lib.hipr_cell_spectra_host_fetch(0, None, None, None, None) == -1
    assert lib.hipr_register_stacks(None, None, None, None, 0, 1, 1, None, None, None, None, None) == -1
    assert lib.hipr_denoise_nl_means_2d(None, 64, 64, 1, 7, 11, 0.02, None, None) == -1


def test_round2_host_entry_points_argument_checks():
    """Argument validation of the entry points added for the host drop-ins and the 3-D denoise, without any device
    call: NULL buffers, bad tables, volumes smaller than the reflection pad, workspace arithmetic."""
    import ctypes as C
    import numpy as np
    import hipr_b200
    from hipr_b200 import tables
    lib = hipr_b200.lib()
    tab = np.ascontiguousarray(tables.line_table_2d(11, 9), dtype=np.int32)
    tp = tab.ctypes.data_as(C.c_void_p)
    img = np.zeros((20, 20))
    ip = img.ctypes.data_as(C.c_void_p)
    assert lib.hipr_line_profile_2d_host(None, 20, 20, 11, 9, tp, ip) == -1
    assert lib.hipr_line_profile_2d_host(ip, 20, 20, 11, 9, tp, None) == -1
    assert lib.hipr_line_profile_2d_host(ip, 20, 20, 11, 9, None, ip) != 0            # no table
    assert lib.hipr_line_profile_2d_host(ip, 8, 20, 11, 9, tp, ip) != 0               # image smaller than the patch
    assert lib.hipr_lne3d_dirs_host(None, 20, 20, 20, 11, 72, None, ip) == -1
    assert lib.hipr_lne3d_dirs_host(ip, 8, 20, 20, 11, 72, None, ip) != 0
    # 3-D denoise: the workspace is the reflect-padded float64 volume, dimensions rounded up to 8 x 11 x 27 tiles
    pad = 3 + 11 + 1
    X, Y, Z = 17, 18, 40
    want = (24 + 2 * pad) * (22 + 2 * pad) * (54 + 2 * pad) * 8
    assert lib.hipr_denoise_nl_means_3d_workspace(X, Y, Z, 11) == want
    assert lib.hipr_denoise_nl_means_3d_workspace(X, Y, Z, 16) < 0
    assert lib.hipr_denoise_nl_means_3d(None, X, Y, Z, 1, 7, 11, 0.03, None, None, 0, None) == -1
    assert lib.hipr_denoise_nl_means_3d(ip, X, Y, Z, 1, 5, 11, 0.03, ip, ip, want, None) != 0      # patch_size 5: unsupported
    assert lib.hipr_denoise_nl_means_3d(ip, 15, Y, Z, 1, 7, 11, 0.03, ip, ip, want, None) != 0     # 15 <= pad: np.pad would wrap
    assert lib.hipr_denoise_nl_means_3d(ip, X, Y, Z, 1, 7, 11, 0.03, ip, ip, want - 8, None) != 0  # workspace too small


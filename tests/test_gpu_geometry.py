"""GPU parity of per-cell geometry (regionprops replacement) and paint-by-label (csrc/cell_geometry.cu)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("H,W,drop,dtype", [(160, 200, 0.0, "int32"), (97, 131, 0.2, "int64"), (64, 33, 0.0, "int32")])
def test_cell_geometry_matches_oracle(torch_cuda, oracle, H, W, drop, dtype):
    import hipr_b200
    from hipr_b200 import synth
    labels, _ = synth.make_labels(H, W, seed=H + W, drop_fraction=drop, dtype=getattr(torch_cuda, dtype))
    wl, wa, wg = oracle.cell_geometry(labels.numpy())
    lab, area, geom = hipr_b200.cell_geometry(labels.cuda())
    assert np.array_equal(lab.cpu().numpy(), wl)            # labels and pixel counts: bit-exact
    assert np.array_equal(area.cpu().numpy(), wa)
    g = geom.cpu().numpy()
    np.testing.assert_allclose(g[:, :2], wg[:, :2], rtol=1e-14)                 # centroid: exact integer sums / area
    np.testing.assert_allclose(g[:, 2:5], wg[:, 2:5], rtol=1e-9, atol=1e-9)     # axis lengths, eccentricity
    np.testing.assert_allclose(g[:, 6:], wg[:, 6:], rtol=1e-10, atol=1e-8)      # central moments
    round_cells = np.abs(wg[:, 2] - wg[:, 3]) < 1e-6 * wg[:, 2]                 # orientation undefined for discs
    np.testing.assert_allclose(g[~round_cells, 5], wg[~round_cells, 5], rtol=1e-8, atol=1e-9)


def test_cell_geometry_single_pixel_and_line_cells(torch_cuda, oracle):
    import hipr_b200
    seg = np.zeros((40, 50), dtype=np.int32)
    seg[3, 4] = 7                 # one pixel: zero moments, eccentricity 0, orientation pi/4 branch
    seg[10, 5:25] = 2             # horizontal line
    seg[12:30, 40] = 9            # vertical line
    seg[20:24, 10:14] = 4         # square: a == c
    wl, wa, wg = oracle.cell_geometry(seg)
    lab, area, geom = hipr_b200.cell_geometry(torch_cuda.from_numpy(seg).cuda())
    assert np.array_equal(lab.cpu().numpy(), wl) and np.array_equal(area.cpu().numpy(), wa)
    np.testing.assert_allclose(geom.cpu().numpy(), wg, rtol=1e-12, atol=1e-12)
    empty = hipr_b200.cell_geometry(torch_cuda.zeros((8, 8), dtype=torch_cuda.int32, device="cuda"))
    assert empty[0].numel() == 0 and empty[2].shape == (0, 9)


@pytest.mark.parametrize("K", [None, 3])
def test_paint_labels(torch_cuda, oracle, K):
    import hipr_b200
    from hipr_b200 import synth
    labels, L = synth.make_labels(120, 150, seed=11, drop_fraction=0.1)
    rng = np.random.default_rng(3)
    values = rng.random((L + 1,) if K is None else (L + 1, K))
    values[0] = 0.0
    want = oracle.paint_labels(labels.numpy(), values)
    got = hipr_b200.paint_labels(labels.cuda(), torch_cuda.from_numpy(values).cuda())
    assert np.array_equal(got.cpu().numpy(), want)
    got32 = hipr_b200.paint_labels(labels.long().cuda(), torch_cuda.from_numpy(values.astype(np.float32)).cuda())
    assert np.array_equal(got32.cpu().numpy(), want.astype(np.float32))
    with pytest.raises(ValueError):
        hipr_b200.paint_labels(labels.cuda(), torch_cuda.from_numpy(values[:5]).cuda(), max_label=L)


def test_next_rows_against_frozen_vectors(torch_cuda):
    """The CUDA path against tests/golden/next_rows_vectors.npz (no oracle code involved)."""
    import os
    import hipr_b200
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "next_rows_vectors.npz"))
    cu = lambda a: torch_cuda.from_numpy(np.ascontiguousarray(a)).cuda()
    cube, ssum, _ = hipr_b200.register_stacks([cu(g["reg_stack%d" % i]) for i in range(3)], g["reg_shifts"],
                                              calibration=cu(g["reg_calibration"]))
    np.testing.assert_allclose(cube.cpu().numpy(), g["reg_cube"], rtol=1.2e-7)
    np.testing.assert_allclose(ssum.cpu().numpy(), g["reg_sum"], rtol=1e-13)
    np.testing.assert_allclose(hipr_b200.denoise_nl_means(cu(g["nlm_in"]), h=0.02).cpu().numpy(), g["nlm_out_h002"], rtol=1e-9)
    np.testing.assert_allclose(hipr_b200.denoise_nl_means(cu(g["nlm_in"]), h=0.1).cpu().numpy(), g["nlm_out_h01"], rtol=1e-9)
    score = hipr_b200.lne2d(hipr_b200.denoise_nl_means(cu(g["nlm_in"]), h=0.02), "F1")
    np.testing.assert_allclose(score.cpu().numpy(), g["nlm_score_F1"], rtol=1e-6, atol=1e-9)
    lab, area, geom = hipr_b200.cell_geometry(cu(g["geo_seg"]))
    assert np.array_equal(lab.cpu().numpy(), g["geo_labels"]) and np.array_equal(area.cpu().numpy(), g["geo_area"])
    np.testing.assert_allclose(geom.cpu().numpy(), g["geo_geometry"], rtol=1e-10, atol=1e-10)
    painted = hipr_b200.paint_labels(cu(g["geo_seg"]), cu(g["paint_values"]))
    assert np.array_equal(painted.cpu().numpy(), g["paint_out"])

"""GPU parity, per-cell mean spectra: labels / pixel counts bit-exact, means rtol 1e-5."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _check(oracle, got, cube_np, lab_np):
    labels, area, avg, norm = [g.cpu().numpy() if hasattr(g, "cpu") else g for g in got]
    wl, wa, wavg, wnorm = oracle.cell_spectra(lab_np, cube_np)
    assert labels.dtype == np.int64 and area.dtype == np.int64
    assert np.array_equal(labels, wl)              # bit-exact ids, ascending present labels
    assert np.array_equal(area, wa)                # bit-exact pixel counts
    np.testing.assert_allclose(avg, wavg, rtol=1e-5, atol=0)
    np.testing.assert_allclose(norm, wnorm, rtol=1e-5, atol=0)


@pytest.mark.parametrize("shape", [(64, 96), (130, 257), (33, 31)])
@pytest.mark.parametrize("label_dtype", ["int32", "int64"])
def test_cell_spectra_fov(torch_cuda, oracle, shape, label_dtype):
    import hipr_b200
    from hipr_b200 import synth
    cube, lab, _ = synth.make_fov(shape[0], shape[1], 95, fov_index=1, label_dtype=getattr(torch_cuda, label_dtype))
    got = hipr_b200.cell_spectra(cube.cuda(), lab.cuda())
    _check(oracle, got, cube.numpy(), lab.numpy())


def test_cell_spectra_noncontiguous_labels(torch_cuda, oracle):
    import hipr_b200
    from hipr_b200 import synth
    cube, lab, L = synth.make_fov(160, 200, 95, fov_index=2, drop_fraction=0.3)
    assert len(np.unique(lab.numpy())) - 1 < L
    _check(oracle, hipr_b200.cell_spectra(cube.cuda(), lab.cuda()), cube.numpy(), lab.numpy())


@pytest.mark.parametrize("C", [1, 8, 32, 63, 64, 97, 130, 300])
def test_cell_spectra_channel_counts(torch_cuda, oracle, C):
    import hipr_b200
    rng = np.random.default_rng(C)
    lab = rng.integers(0, 6, (40, 50)).astype(np.int32) * (rng.random((40, 50)) > 0.3)
    cube = rng.random((40, 50, C)).astype(np.float32)
    got = hipr_b200.cell_spectra(torch_cuda.from_numpy(cube).cuda(), torch_cuda.from_numpy(lab.astype(np.int32)).cuda())
    _check(oracle, got, cube, lab)


def test_cell_spectra_edge_cases(torch_cuda, oracle):
    import hipr_b200
    t = torch_cuda
    cube = np.random.default_rng(0).random((20, 20, 95)).astype(np.float32)
    # all background
    lab = np.zeros((20, 20), np.int32)
    got = hipr_b200.cell_spectra(t.from_numpy(cube).cuda(), t.from_numpy(lab).cuda())
    assert all(g.shape[0] == 0 for g in got)
    # negative labels are background too; one huge sparse id; single-pixel cell
    lab[3, 4] = -7
    lab[5, 5] = 100000
    lab[6:9, 6:9] = 2
    _check(oracle, hipr_b200.cell_spectra(t.from_numpy(cube).cuda(), t.from_numpy(lab).cuda()), cube, lab)
    # one label covering everything
    lab[:] = 3
    _check(oracle, hipr_b200.cell_spectra(t.from_numpy(cube).cuda(), t.from_numpy(lab).cuda()), cube, lab)


def test_cell_spectra_properties(torch_cuda, oracle):
    """counts sum = number of foreground pixels; relabelling permutes rows."""
    import hipr_b200
    from hipr_b200 import synth
    cube, lab, L = synth.make_fov(128, 192, 95, fov_index=4)
    labels, area, avg, _ = hipr_b200.cell_spectra(cube.cuda(), lab.cuda())
    assert int(area.sum()) == int((lab > 0).sum())
    perm = np.random.default_rng(0).permutation(L) + 1
    lut = np.concatenate([[0], perm]).astype(np.int32)
    lab2 = torch_cuda.from_numpy(lut[lab.numpy()])
    labels2, area2, avg2, _ = hipr_b200.cell_spectra(cube.cuda(), lab2.cuda())
    order = np.argsort(lut[labels.cpu().numpy()])
    assert np.array_equal(area.cpu().numpy()[order], area2.cpu().numpy())
    np.testing.assert_allclose(avg.cpu().numpy()[order], avg2.cpu().numpy(), rtol=1e-12)


def test_cell_spectra_slabs_accumulate(torch_cuda, oracle):
    """Row slabs accumulated into the same sums/counts = the whole image (mosaic split path)."""
    import hipr_b200
    from hipr_b200 import synth
    cube, lab, L = synth.make_fov(120, 160, 95, fov_index=6)
    c, l = cube.cuda(), lab.cuda()
    sums = counts = None
    for r0 in (0, 40, 90):
        r1 = {0: 40, 40: 90, 90: 120}[r0]
        sums, counts = hipr_b200.cell_spectra_accumulate(c[r0:r1], l[r0:r1], L, sums, counts)
    _check(oracle, hipr_b200.cell_spectra_finalize(sums, counts), cube.numpy(), lab.numpy())


def test_cell_spectra_3d(torch_cuda, oracle):
    import hipr_b200
    rng = np.random.default_rng(3)
    lab = (rng.integers(0, 9, (6, 7, 8)) * (rng.random((6, 7, 8)) > 0.5)).astype(np.int64)
    cube = rng.random((6, 7, 8, 63)).astype(np.float32)
    got = hipr_b200.cell_spectra(torch_cuda.from_numpy(cube).cuda(), torch_cuda.from_numpy(lab).cuda())
    _check(oracle, got, cube, lab)


def test_cell_spectra_host_entry(torch_cuda, oracle):
    import hipr_b200
    from hipr_b200 import synth
    cube, lab, _ = synth.make_fov(150, 222, 95, fov_index=7, label_dtype=torch_cuda.int64)
    got = hipr_b200.cell_spectra_host(cube.numpy(), lab.numpy())
    _check(oracle, got, cube.numpy(), lab.numpy())


def test_host_entry_point_more_cells_than_first_capacity(torch_cuda, oracle):
    """More cells than the wrapper's first guess (4096): the table is fetched from the device without a second
    pass (hipr_cell_spectra_host_fetch)."""
    import hipr_b200
    rng = np.random.default_rng(8)
    H, W, Cn = 96, 128, 7
    labels = np.zeros((H, W), dtype=np.int32)
    ids = rng.permutation(H * W)[:6000]
    labels.reshape(-1)[ids] = np.arange(1, 6001, dtype=np.int32) * 3        # non-contiguous ids, one pixel each
    labels[10:14, 10:30] = 7                                                # and one larger cell
    cube = rng.random((H, W, Cn), dtype=np.float32)
    lab, area, avg, norm = hipr_b200.cell_spectra_host(cube, labels)
    wl, wa, wavg, wnorm = oracle.cell_spectra(labels, cube)
    assert lab.size > 4096 and np.array_equal(lab, wl) and np.array_equal(area, wa)
    np.testing.assert_allclose(avg, wavg, rtol=1e-5)
    np.testing.assert_allclose(norm, wnorm, rtol=1e-5)

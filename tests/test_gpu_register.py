"""GPU parity of the registration paste + channel stack + flat field + channel sum kernel
(csrc/register.cu) against the numpy restatement of syn/..._measurement.py:86-105."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CHANS = (32, 23, 20, 14, 6)     # 405 / 488 / 514 / 561 / 633 nm: 95 channels (eco/Snakefile:34)


def _stacks(rng, H, W, chans):
    return [rng.random((H, W, c), dtype=np.float32) for c in chans]


@pytest.mark.parametrize("H,W,chans,shifts", [
    (70, 70, CHANS, [(0, 0), (3, -2), (-4, 5), (1, 1), (-1, -7)]),
    (64, 64, (23, 20, 14, 6), [(0.0, 0.0), (2.9, -1.2), (-0.7, 3.99), (12.0, -9.0)]),    # syn: 63 channels, float shifts
    (33, 131, (5, 1, 7), [(0, 0), (-32, 130), (32, -130)]),                                # shifts that leave one row / column
    (40, 50, (95,), [(0, 0)]),
    (20, 20, (4, 4), [(0, 0), (20, -20)]),                                                  # shifted exactly out of the frame
])
def test_register_matches_oracle(torch_cuda, oracle, H, W, chans, shifts):
    import hipr_b200
    rng = np.random.default_rng(H * W)
    stacks = _stacks(rng, H, W, chans)
    want_cube, want_sum = oracle.register_stacks(stacks, shifts)
    cube, s, mk = hipr_b200.register_stacks([torch_cuda.from_numpy(a).cuda() for a in stacks], shifts)
    assert np.array_equal(cube.cpu().numpy(), want_cube.astype(np.float32))          # a paste: bit-exact
    np.testing.assert_allclose(s.cpu().numpy(), want_sum, rtol=1e-14, atol=0)
    vmax, vmin = mk.values()
    assert float(vmax) == s.max().item() and float(vmin) == s.min().item()


def test_register_flat_field(torch_cuda, oracle):
    import hipr_b200
    rng = np.random.default_rng(5)
    H, W = 96, 80
    stacks = _stacks(rng, H, W, CHANS)
    shifts = [(0, 0), (2, 1), (-3, 0), (0, -4), (5, 5)]
    cal = (0.5 + rng.random((H, W, 95), dtype=np.float32)).astype(np.float32)
    want_cube, want_sum = oracle.register_stacks(stacks, shifts, cal)
    dev = [torch_cuda.from_numpy(a).cuda() for a in stacks]
    cube, s, mk = hipr_b200.register_stacks(dev, shifts, calibration=torch_cuda.from_numpy(cal).cuda())
    # the cube is stored in float32: one rounding of numpy's float64 quotient
    np.testing.assert_allclose(cube.cpu().numpy(), want_cube, rtol=1.2e-7, atol=0)
    np.testing.assert_allclose(s.cpu().numpy(), want_sum, rtol=1e-13, atol=0)
    # no cube requested: sums only
    none_cube, s2, _ = hipr_b200.register_stacks(dev, shifts, calibration=torch_cuda.from_numpy(cal).cuda(), return_cube=False)
    assert none_cube is None and torch_cuda.equal(s, s2)


def test_score_and_cells_from_stacks(torch_cuda, oracle):
    """Stacks -> registered cube -> score map and per-cell spectra == the oracle run on the oracle's cube."""
    import hipr_b200
    from hipr_b200 import synth
    H = W = 160
    cube0, labels, _ = synth.make_fov(H, W, 95, fov_index=3)
    cube0 = cube0.numpy()
    edges = np.cumsum((0,) + CHANS)
    stacks = [np.ascontiguousarray(cube0[:, :, a:b]) for a, b in zip(edges[:-1], edges[1:])]
    shifts = [(0, 0), (1, -1), (-2, 0), (0, 2), (1, 1)]
    want_cube, _ = oracle.register_stacks(stacks, shifts)
    score, cube, s, mk = hipr_b200.neighbor2d_score_from_stacks([torch_cuda.from_numpy(a).cuda() for a in stacks], shifts)
    want_score = oracle.neighbor2d_score(want_cube.astype(np.float32), "F1")
    np.testing.assert_allclose(score.cpu().numpy(), want_score, rtol=1e-5, atol=5e-7)
    lab, area, avg, norm = hipr_b200.cell_spectra(cube, labels.cuda())
    wl, wa, wavg, wnorm = oracle.cell_spectra(labels.numpy(), want_cube)
    assert np.array_equal(lab.cpu().numpy(), wl) and np.array_equal(area.cpu().numpy(), wa)
    np.testing.assert_allclose(avg.cpu().numpy(), wavg, rtol=1e-5)
    np.testing.assert_allclose(norm.cpu().numpy(), wnorm, rtol=1e-5)


def test_register_argument_errors(torch_cuda, oracle):
    import hipr_b200
    a = torch_cuda.zeros((8, 8, 3), device="cuda")
    with pytest.raises(ValueError):                                          # the reference's paste raises too
        oracle.register_stacks([np.zeros((8, 8, 3))], [(9, 0)])
    with pytest.raises(ValueError):
        hipr_b200.register_stacks([a], shifts=[(9, 0)])
    with pytest.raises(ValueError):
        hipr_b200.register_stacks([a], shifts=[(0, -9)])
    with pytest.raises(ValueError):
        hipr_b200.register_stacks([])
    with pytest.raises(ValueError):
        hipr_b200.register_stacks([a, torch_cuda.zeros((8, 9, 3), device="cuda")])
    with pytest.raises(ValueError):
        hipr_b200.register_stacks([a], shifts=[(0, 0), (1, 1)])
    with pytest.raises(ValueError):
        hipr_b200.register_stacks([a] * 9)                                   # more than 8 stacks
    with pytest.raises(ValueError):
        hipr_b200.register_stacks([torch_cuda.zeros((8, 8, 200), device="cuda")])   # more than 192 channels
    with pytest.raises(ValueError):
        hipr_b200.register_stacks([a.cpu()])


@pytest.mark.parametrize("with_cal", [False, True])
def test_register_interior_and_border_tiles(torch_cuda, oracle, with_cal):
    """A frame wide enough for the bulk-copy interior kernel (W % 4 == 0, several 64-pixel tiles) with shifts of
    every residue mod 4, so that every alignment slack and both kernels (interior / border tiles) are hit."""
    import hipr_b200
    rng = np.random.default_rng(9)
    H, W = 37, 448
    stacks = _stacks(rng, H, W, CHANS)
    for shifts in ([(0, 0), (3, -2), (-4, 5), (1, 1), (-1, -7)], [(2, 3), (0, -65), (5, 70), (-3, 2), (0, 0)],
                   [(0, 0), (0, 0), (0, 0), (0, 0), (0, 0)]):
        cal = (0.5 + rng.random((H, W, 95), dtype=np.float32)).astype(np.float32) if with_cal else None
        want_cube, want_sum = oracle.register_stacks(stacks, shifts, cal)
        cube, s, mk = hipr_b200.register_stacks([torch_cuda.from_numpy(a).cuda() for a in stacks], shifts,
                                                calibration=None if cal is None else torch_cuda.from_numpy(cal).cuda())
        if with_cal:
            np.testing.assert_allclose(cube.cpu().numpy(), want_cube, rtol=1.2e-7, atol=0)
            np.testing.assert_allclose(s.cpu().numpy(), want_sum, rtol=1e-13, atol=0)
        else:
            assert np.array_equal(cube.cpu().numpy(), want_cube.astype(np.float32))
            np.testing.assert_allclose(s.cpu().numpy(), want_sum, rtol=1e-14, atol=0)
        vmax, vmin = mk.values()
        assert float(vmax) == s.max().item() and float(vmin) == s.min().item()


@pytest.mark.parametrize("chans", [CHANS, (23, 20, 14, 6)])
@pytest.mark.parametrize("with_cal", [False, True])
def test_register_compiled_layouts_ragged_width(torch_cuda, oracle, chans, with_cal):
    """Both channel layouts the straight-line kernel is compiled for (95 = 32+23+20+14+6, 63 = 23+20+14+6), on a frame
    whose width is not a multiple of the 64-pixel tile (the last tile of a row takes the general route inside the same
    launch), with row and column shifts, with and without a flat field; and HIPR_REGISTER_GENERIC-independent parity:
    the same inputs through a layout the kernel is not compiled for (one channel moved) agree with the oracle too."""
    import hipr_b200
    rng = np.random.default_rng(31 + len(chans))
    H, W = 29, 200
    Cn = sum(chans)
    stacks = _stacks(rng, H, W, chans)
    shifts = [(0, 0), (3, -2), (-4, 5), (1, 1), (-1, -7)][:len(chans)]
    cal = (0.5 + rng.random((H, W, Cn), dtype=np.float32)).astype(np.float32) if with_cal else None
    want_cube, want_sum = oracle.register_stacks(stacks, shifts, cal)
    cube, s, mk = hipr_b200.register_stacks([torch_cuda.from_numpy(a).cuda() for a in stacks], shifts,
                                            calibration=None if cal is None else torch_cuda.from_numpy(cal).cuda())
    if with_cal:
        assert np.array_equal(cube.cpu().numpy(), (want_cube.astype(np.float32)))   # correctly rounded float32 quotient
        np.testing.assert_allclose(s.cpu().numpy(), want_sum, rtol=1e-13, atol=0)
    else:
        assert np.array_equal(cube.cpu().numpy(), want_cube.astype(np.float32))
        np.testing.assert_allclose(s.cpu().numpy(), want_sum, rtol=1e-14, atol=0)

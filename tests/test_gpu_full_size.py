"""Parity at BASELINE.json's sizes.  Config 1 (512 x 512 x 95) is small enough for the oracle itself (with the
compiled reference stencil where oracle/_ref exists).  Config 2 (2048 x 2048 x 95) is checked through windows
(the oracle on crops of the full-size inputs against the same window of the full-size CUDA result) and through
size-independent properties: exact scale invariance, checksum of checksums, pixel-count conservation, label
permutation.  A 3-D volume (config 4's cross-section at a quarter of its extent) gets the same treatment."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-5, 5e-7


def test_config1_fov_against_the_oracle(torch_cuda, oracle):
    """512 x 512 x 95: score map and per-cell spectra, whole FOV, oracle on the CPU."""
    import hipr_b200
    from hipr_b200 import synth
    from oracle import load_ref
    ref = load_ref("neighbor2d")
    cube, labels, _ = synth.make_fov(512, 512, 95, fov_index=0, drop_fraction=0.1)
    want = oracle.neighbor2d_score(cube.numpy(), "F1", lp_func=ref.line_profile_2d_v2 if ref is not None else None)
    got = hipr_b200.neighbor2d_score(cube.cuda(), "F1")
    np.testing.assert_allclose(got.cpu().numpy(), want, rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(hipr_b200.neighbor2d_score_host(cube.numpy(), "F1"), want, rtol=RTOL, atol=ATOL)
    lab, area, avg, norm = hipr_b200.cell_spectra(cube.cuda(), labels.cuda())
    wl, wa, wavg, wnorm = oracle.cell_spectra(labels.numpy(), cube.numpy())
    assert np.array_equal(lab.cpu().numpy(), wl) and np.array_equal(area.cpu().numpy(), wa)
    np.testing.assert_allclose(avg.cpu().numpy(), wavg, rtol=1e-5)
    np.testing.assert_allclose(norm.cpu().numpy(), wnorm, rtol=1e-5)


@pytest.fixture(scope="module")
def fov2048(torch_cuda):
    from hipr_b200 import synth
    cube, labels, L = synth.make_fov(2048, 2048, 95, fov_index=1, device="cuda")
    return cube, labels, L


def test_config2_windows_against_the_oracle(torch_cuda, oracle, fov2048):
    """Oracle on 96 x 128 windows (+ 5-pixel halo) of the 2048^2 FOV == the same window of the full CUDA result.
    The score of a pixel depends on its 11 x 11 neighbourhood only (every normalisation cancels)."""
    import hipr_b200
    cube, _, _ = fov2048
    full = hipr_b200.neighbor2d_score(cube, "F1").cpu().numpy()
    assert full.shape == (2048, 2048) and np.isfinite(full).all() and full.min() >= 0.0 and full.max() <= 1.0
    for (r, c) in [(0, 0), (1000, 517), (2048 - 96, 2048 - 128), (333, 2048 - 128), (2048 - 96, 40)]:
        r0, r1, c0, c1 = max(r - 5, 0), min(r + 96 + 5, 2048), max(c - 5, 0), min(c + 128 + 5, 2048)
        crop = cube[r0:r1, c0:c1].cpu().numpy()
        want = oracle.neighbor2d_score(crop, "F1")[r - r0: r - r0 + 96, c - c0: c - c0 + 128]
        np.testing.assert_allclose(full[r:r + 96, c:c + 128], want, rtol=RTOL, atol=0)    # strictly relative


def test_config2_properties(torch_cuda, fov2048):
    import hipr_b200
    cube, labels, L = fov2048
    score = hipr_b200.neighbor2d_score(cube, "F1")
    # exact scale invariance: doubling every sample doubles every float64 sum exactly
    assert torch_cuda.equal(hipr_b200.neighbor2d_score(cube * 2.0, "F1"), score)
    # checksum of checksums: sum over pixels of the channel sums == sum over channels of the channel totals
    s = hipr_b200.channel_sum(cube, normalize=False, dtype=torch_cuda.float64)
    per_channel = cube.to(torch_cuda.float64).sum(dim=(0, 1))
    np.testing.assert_allclose(float(s.sum()), float(per_channel.sum()), rtol=1e-12)
    # per-cell reduction: pixel counts are conserved exactly, intensity totals to float32-accumulation accuracy
    lab, area, avg, norm = hipr_b200.cell_spectra(cube, labels)
    fg = labels > 0
    assert int(area.sum()) == int(fg.sum())
    assert np.array_equal(lab.cpu().numpy(), np.unique(labels[fg].cpu().numpy()))
    total = (avg * area[:, None].double()).sum(dim=0)
    np.testing.assert_allclose(total.cpu().numpy(), cube[fg].to(torch_cuda.float64).sum(dim=0).cpu().numpy(), rtol=1e-6)
    assert float(norm.max(dim=1).values.min()) == 1.0 and float(norm.max()) == 1.0
    # relabelling the cells permutes the rows and nothing else
    perm = torch_cuda.randperm(L, device="cuda", generator=torch_cuda.Generator(device="cuda").manual_seed(3)) + 1
    lut = torch_cuda.cat([torch_cuda.zeros(1, dtype=perm.dtype, device="cuda"), perm])
    relab = lut[labels.long()].to(labels.dtype)
    lab2, area2, avg2, _ = hipr_b200.cell_spectra(cube, relab)
    back = torch_cuda.argsort(lut[lab])              # row j of the relabelled table is original row back[j]
    assert torch_cuda.equal(lab2, lut[lab][back])
    assert torch_cuda.equal(area2, area[back])
    np.testing.assert_allclose(avg2.cpu().numpy(), avg[back].cpu().numpy(), rtol=1e-6)
    # geometry on the same labels: areas agree with the spectra reduction, centroids lie inside the image
    glab, garea, geom = hipr_b200.cell_geometry(labels, L)
    assert torch_cuda.equal(glab, lab) and torch_cuda.equal(garea, area)
    assert float(geom[:, 0].min()) >= 0 and float(geom[:, 0].max()) <= 2047 and float(geom[:, 3].min()) >= 0


def test_volume_windows_and_scale_invariance(torch_cuda, oracle):
    """256 x 256 x 64 x 95 z-stack (config 4's depth): windows against the oracle, exact scale invariance."""
    import hipr_b200
    from hipr_b200 import synth
    cube = synth.make_volume_cube(256, 256, 64, 95, seed=7, device="cuda")
    full = hipr_b200.neighbor3d_score(cube, "ME2")
    assert torch_cuda.equal(hipr_b200.neighbor3d_score(cube * 4.0, "ME2"), full)
    full = full.cpu().numpy()
    assert np.isfinite(full).all()
    s = hipr_b200.channel_sum(cube, normalize=False, dtype=torch_cuda.float64)
    gmax = float(s.max())
    for (x, y, z) in [(0, 0, 0), (120, 77, 30), (256 - 12, 256 - 10, 64 - 14)]:
        lo = [max(v - 5, 0) for v in (x, y, z)]
        hi = [min(v + n + 5, m) for v, n, m in zip((x, y, z), (12, 10, 14), (256, 256, 64))]
        win = s[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]].cpu().numpy() / gmax      # ME2's 1e-8 is relative to the global max
        want = oracle.lne3d(win, "ME2")[x - lo[0]: x - lo[0] + 12, y - lo[1]: y - lo[1] + 10, z - lo[2]: z - lo[2] + 14]
        np.testing.assert_allclose(full[x:x + 12, y:y + 10, z:z + 14], want, rtol=RTOL, atol=ATOL)


def test_config2_registration_paste_is_exact(torch_cuda):
    """2048 x 2048 x 95 through K0 (registration paste + channel stack + flat field + channel sum): the cube equals a
    paste written with torch slicing bit for bit (the float32 quotient of the flat field is IEEE division), the sums
    equal float64 sums of the float64 quotients to 1e-13 -- a whole-FOV check no CPU oracle could finish in seconds."""
    import hipr_b200
    torch = torch_cuda
    H = W = 2048
    chans = (32, 23, 20, 14, 6)
    shifts = [(0, 0), (3, -2), (-4, 1), (2, 5), (-1, -3)]
    g = torch.Generator(device="cuda").manual_seed(7)
    stacks = [torch.rand((H, W, c), generator=g, device="cuda") + 0.1 for c in chans]
    cal = torch.rand((H, W, 95), generator=g, device="cuda") + 0.5
    want = torch.zeros((H, W, 95), device="cuda")
    o = 0
    for st, (dr, dc), c in zip(stacks, shifts, chans):
        want[max(dr, 0):H + min(dr, 0), max(dc, 0):W + min(dc, 0), o:o + c] = \
            st[max(-dr, 0):H + min(-dr, 0), max(-dc, 0):W + min(-dc, 0)]
        o += c
    cube, s, mk = hipr_b200.register_stacks(stacks, shifts)
    assert torch.equal(cube, want)
    ref = want.double().sum(2)
    assert float(((s - ref).abs() / ref).max()) < 1e-14
    vmax, vmin = mk.values()
    assert float(vmax) == s.max().item() and float(vmin) == s.min().item()
    cube_c, s_c, _ = hipr_b200.register_stacks(stacks, shifts, cal)
    assert torch.equal(cube_c, want / cal)
    ref_c = (want.double() / cal.double()).sum(2)
    assert float(((s_c - ref_c).abs() / ref_c).max()) < 1e-13


def test_config2_literal_gather_is_exact(torch_cuda):
    """The strict drop-in output at 2048 x 2048 (3.3 GB in float64): every one of the 99 samples of every pixel equals
    the padded image at the table's offset -- checked on the device with index arithmetic, direction by direction."""
    import hipr_b200
    from hipr_b200 import tables
    torch = torch_cuda
    H = W = 2048
    g = torch.Generator(device="cuda").manual_seed(8)
    pad = torch.rand((H + 10, W + 10), generator=g, device="cuda", dtype=torch.float64)
    lp = hipr_b200.line_profile_2d(pad, 11, 9)
    assert tuple(lp.shape) == (H, W, 9, 11)
    tab = tables.line_table_2d(11, 9)                        # (9, 11, 2) patch coordinates
    for t in range(9):
        for li in range(11):
            dy, dx = int(tab[t, li, 0]), int(tab[t, li, 1])
            assert torch.equal(lp[:, :, t, li], pad[dy:dy + H, dx:dx + W])
    lp32 = hipr_b200.line_profile_2d(pad.float(), 11, 9)
    assert torch.equal(lp32[:, :, 4, 7], pad.float()[int(tab[4, 7, 0]):int(tab[4, 7, 0]) + H, int(tab[4, 7, 1]):int(tab[4, 7, 1]) + W])

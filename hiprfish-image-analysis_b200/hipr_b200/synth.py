"""Seeded synthetic fields of view: a (H, W, C) float32 spectral cube
and an int32 label image of rod-shaped cells.  Runs on any torch device; generation is never
inside a timed region.  Not part of the hot path."""
import math

import torch

EXCITATION_BLOCKS = (0, 32, 55, 75, 89, 95)   # 405/488/514/561/633 nm: 32+23+20+14+6 channels


def make_labels(H, W, seed=4321, device="cpu", cell=(16, 32), drop_fraction=0.0, dtype=torch.int32):
    """Non-overlapping capsules on a jittered grid, ids 1..L in raster order of the grid;
    `drop_fraction` of the ids are removed (non-contiguous labels, as the reference keeps
    the original watershed ids)."""
    g = torch.Generator(device="cpu").manual_seed(int(seed))
    ch, cw = cell
    py, px = ch + 10, cw + 12                      # grid pitch
    ny, nx = max(H // py, 1), max(W // px, 1)
    L = ny * nx
    jy = torch.randint(0, 8, (ny, nx), generator=g)
    jx = torch.randint(0, 8, (ny, nx), generator=g)
    ang = torch.rand((ny, nx), generator=g) * math.pi
    keep = torch.rand((ny, nx), generator=g) >= drop_fraction
    yy = torch.arange(H, device=device)[:, None]
    xx = torch.arange(W, device=device)[None, :]
    gy = torch.clamp(yy // py, max=ny - 1)
    gx = torch.clamp(xx // px, max=nx - 1)
    jy, jx, ang, keep = jy.to(device), jx.to(device), ang.to(device), keep.to(device)
    cy = gy * py + py // 2 + jy[gy, gx] - 4
    cx = gx * px + px // 2 + jx[gy, gx] - 4
    a = ang[gy, gx]
    dy, dx = (yy - cy).float(), (xx - cx).float()
    u = dx * torch.cos(a) + dy * torch.sin(a)      # along the rod
    v = -dx * torch.sin(a) + dy * torch.cos(a)
    half_len, rad = (cw - ch) / 2.0, ch / 2.0 - 2.0
    du = torch.clamp(u.abs() - half_len, min=0.0)
    inside = (du * du + v * v) <= rad * rad
    # keep a cell inside its own grid box so cells never touch
    inside &= ((yy - gy * py) >= 1) & ((yy - gy * py) < py - 1) & ((xx - gx * px) >= 1) & ((xx - gx * px) < px - 1)
    ids = (gy * nx + gx + 1).to(torch.int64)
    lab = torch.where(inside & keep[gy, gx], ids, torch.zeros_like(ids))
    return lab.to(dtype), L


def make_cube(H, W, C=95, seed=1234, device="cpu", labels=None, n_barcodes=16):
    """cube[y, x, c] = barcode[label][c] * cell mask + 0.02 + U(0, 0.05): fp32, no flat lines."""
    g = torch.Generator(device=device).manual_seed(int(seed))
    gb = torch.Generator(device="cpu").manual_seed(int(seed) + 7)
    bank = torch.zeros((n_barcodes, C))
    edges = [e for e in EXCITATION_BLOCKS if e < C] + [C]
    for b in range(n_barcodes):
        for lo, hi in zip(edges[:-1], edges[1:]):
            if hi <= lo:
                continue
            if torch.rand(1, generator=gb).item() < 0.6:
                pos = torch.arange(hi - lo).float()
                mu = torch.rand(1, generator=gb).item() * (hi - lo)
                bank[b, lo:hi] = (0.3 + 0.7 * torch.rand(1, generator=gb).item()) * torch.exp(-0.5 * ((pos - mu) / 3.0) ** 2)
    bank = bank.to(device)
    cube = torch.rand((H, W, C), generator=g, device=device, dtype=torch.float32) * 0.05 + 0.02
    if labels is not None:
        lab = labels.to(device=device, dtype=torch.int64)
        code = bank[(lab % n_barcodes)]            # (H, W, C)
        cube += code * (lab > 0)[..., None]
    return cube


def make_fov(H, W, C=95, fov_index=0, device="cpu", drop_fraction=0.0, label_dtype=torch.int32):
    labels, L = make_labels(H, W, seed=4321 + fov_index, device=device, drop_fraction=drop_fraction,
                            dtype=label_dtype)
    cube = make_cube(H, W, C, seed=1234 + fov_index, device=device, labels=labels)
    return cube, labels, L


def make_volume_cube(X, Y, Z, C=95, seed=99, device="cpu"):
    """(X, Y, Z, C) float32 z-stack: smooth blobs + noise (biofilm stand-in)."""
    g = torch.Generator(device=device).manual_seed(int(seed))
    xs = torch.arange(X, device=device).float()[:, None, None]
    ys = torch.arange(Y, device=device).float()[None, :, None]
    zs = torch.arange(Z, device=device).float()[None, None, :]
    field = (torch.sin(xs / 7.0) * torch.cos(ys / 9.0) * torch.sin(zs / 5.0 + 1.0)) ** 2
    cube = torch.rand((X, Y, Z, C), generator=g, device=device, dtype=torch.float32) * 0.05 + 0.02
    spec = torch.linspace(0.2, 1.0, C, device=device)
    cube += field[..., None] * spec
    return cube

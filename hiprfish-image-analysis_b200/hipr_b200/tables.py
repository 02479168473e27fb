"""Host-side line tables.

The reference builds its sampling table with Python-object arithmetic at the top of every call
(eco/neighbor2d.pyx:32-55; bio/neighbor.pyx:141-170 for 3-D; :301-324 for the v3 variant).  The
entries depend on the last bit of libm's cos/sin through np.round (5*cos(60 deg) =
2.5000000000000004 rounds to 3, 5*cos(120 deg) = -2.4999999999999996 to -2), so the table is
data: it is computed here on the host exactly as the reference does, checked against a pinned
hash for the parameters the pipelines use, and handed to the kernels.  It is never recomputed
on the device.

ABI layout: int32 (n_dirs, patch_size, ndim), patch coordinates 0..patch_size-1.
"""
import functools
import hashlib
import operator

import numpy as np

# sha256 of the int64 (n_dirs, P, ndim) array of OFFSETS FROM THE CENTRE, C order
PINNED = {
    ("2d", 11, 9): "b1278799e8f5fd61d163437987943d3b2d104245e493df6dab5f2072335dbb3d",
    ("3d", 11, 9, 9): "13bca0604962da5b1568581024e69ea5f4079c9ad9b04745a4beec04a8d93040",
}


def _int_arg(v, name):
    """C `int` argument conversion as the Cython modules do it: integers as they are, floats
    truncated (the reference accepts 11.0 and 11.5 alike), anything else is a TypeError."""
    try:
        return operator.index(v)
    except TypeError:
        if isinstance(v, (float, np.floating)):
            return int(v)
        raise TypeError("%s: an integer is required" % name)


def _line(endpoint, P):
    """Patch coordinates (P, ndim) of the samples of the line whose half-length end point is
    `endpoint` (integer offsets per axis)."""
    half = (P - 1) // 2
    e = np.asarray(endpoint, dtype=np.int64)
    reach = int(np.abs(e).max())
    n = 2 * reach + 1
    li = np.arange(n, dtype=np.int64)[:, None]
    frac = np.sign(e)[None, :] * li * (2 * np.abs(e)[None, :] + 1) / n       # float64 true division
    coords = (np.sign(frac) * np.floor(np.abs(frac))).astype(np.int64) + half - e[None, :]
    if n < P:
        lead = (P - n) // 2
        coords = np.concatenate([np.repeat(coords[:1], lead, 0), coords, np.repeat(coords[-1:], lead, 0)])
    return coords[:P]


@functools.lru_cache(maxsize=32)
def _table_2d(P, R):
    half = (P - 1) // 2
    rows = []
    for phi in range(R):
        end = (int(np.round(half * np.cos(phi * np.pi / R))), int(np.round(half * np.sin(phi * np.pi / R))))
        rows.append(_line(end, P))
    return np.ascontiguousarray(np.stack(rows), dtype=np.int32)


@functools.lru_cache(maxsize=32)
def _table_3d(P, TH, PH):
    half = (P - 1) // 2
    rows = []
    for theta in range(1, TH):
        st = np.sin(theta * np.pi / TH)
        for phi in range(PH):
            end = (int(np.round(half * np.cos(phi * np.pi / PH) * st)),
                   int(np.round(half * np.sin(phi * np.pi / PH) * st)),
                   int(np.round(half * np.cos(theta * np.pi / TH))))
            rows.append(_line(end, P))
    return np.ascontiguousarray(np.stack(rows), dtype=np.int32)


@functools.lru_cache(maxsize=8)
def _table_3d_v3(P, TH, PH):
    """bio/neighbor.pyx:301-324: rounding differs from the v2 table and full-length lines use the
    signed step inside (2*step+1), so entries can leave the patch (up to 18 for (11,9,9))."""
    half = (P - 1) // 2
    rows = []
    for theta in range(1, TH):
        st = np.sin(theta * np.pi / TH)
        for phi in range(PH):
            e = np.array([int(np.round(half * np.cos(phi * np.pi / PH) * st)),
                          int(np.round(half * np.sin(phi * np.pi / PH) * st)),
                          int(np.round(half * np.cos(theta * np.pi / TH)))], dtype=np.int64)
            n = 2 * int(np.abs(e).max()) + 1
            li = np.arange(n, dtype=np.int64)[:, None]
            if n < P:
                frac = np.sign(e)[None, :] * li * (2 * np.abs(e)[None, :] + 1) / n
                coords = np.round(frac).astype(np.int64) + half - e[None, :]
                lead = (P - n) // 2
                coords = np.concatenate([np.repeat(coords[:1], lead, 0), coords, np.repeat(coords[-1:], lead, 0)])
            else:
                frac = np.sign(e)[None, :] * li * (2 * e[None, :] + 1) / n
                coords = np.floor(frac).astype(np.int64) + half - e[None, :]
            rows.append(coords[:P])
    return np.ascontiguousarray(np.stack(rows), dtype=np.int32)


def _validate(P, *ranges):
    if P < 3 or P % 2 == 0 or P > 31:
        raise ValueError("patch_size must be odd and in 3..31, got %d" % P)
    for r in ranges:
        if r < 1:
            raise ValueError("angle ranges must be positive")


def _pin(kind, tab, *params):
    want = PINNED.get((kind,) + params)
    if want is not None:
        half = (params[0] - 1) // 2
        got = hashlib.sha256((tab.astype(np.int64) - half).tobytes()).hexdigest()
        if got != want:
            raise RuntimeError("line table for %s%r does not match the pinned reference table "
                               "(libm rounding differs on this host?)" % (kind, params))
    return tab


def line_table_2d(patch_size, phi_range):
    P, R = _int_arg(patch_size, "patch_size"), _int_arg(phi_range, "phi_range")
    _validate(P, R)
    return _pin("2d", _table_2d(P, R), P, R)


def line_table_3d(patch_size, theta_range, phi_range):
    P, TH, PH = (_int_arg(patch_size, "patch_size"), _int_arg(theta_range, "theta_range"),
                 _int_arg(phi_range, "phi_range"))
    _validate(P, TH - 1, PH)
    return _pin("3d", _table_3d(P, TH, PH), P, TH, PH)


def line_table_3d_v3(patch_size, theta_range, phi_range):
    P, TH, PH = (_int_arg(patch_size, "patch_size"), _int_arg(theta_range, "theta_range"),
                 _int_arg(phi_range, "phi_range"))
    _validate(P, TH - 1, PH)
    return _table_3d_v3(P, TH, PH)

"""Generates tests/golden/next_rows_vectors.npz: frozen inputs / outputs of the oracle's restatements for the
rows either side of the stencil -- registration paste + flat field + channel sum,
non-local-means denoise, per-cell geometry, paint by label.  The registration and paint blocks are numpy in
the reference (restated verbatim); NL-means and regionprops are scikit-image, which is neither pinned by the
reference nor installed here, so those vectors freeze the restated algorithm (parity unpinned at that
boundary) and guard the oracle against drift between rounds.

    python tests/golden/make_golden_next.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import hipr_oracle as O  # noqa: E402


def main():
    out = {}
    rng = np.random.default_rng(21)
    H, W = 24, 28
    chans = (6, 5, 4)
    stacks = [rng.random((H, W, c)).astype(np.float32) for c in chans]
    shifts = np.array([[0, 0], [2, -3], [-4, 1]], dtype=np.int64)
    cal = (0.5 + rng.random((H, W, sum(chans)))).astype(np.float32)
    for i, s in enumerate(stacks):
        out["reg_stack%d" % i] = s
    out["reg_shifts"] = shifts
    out["reg_calibration"] = cal
    cube, ssum = O.register_stacks(stacks, shifts, cal)
    out["reg_cube"], out["reg_sum"] = cube, ssum
    yy, xx = np.mgrid[0:36, 0:40]
    img = (np.sin(yy / 5.0) ** 2 + np.cos(xx / 7.0) ** 2) / 2 + 0.03 * rng.random((36, 40))
    img = img / img.max()
    out["nlm_in"] = img
    out["nlm_out_h002"] = O.denoise_nl_means_2d(img, h=0.02)
    out["nlm_out_h01"] = O.denoise_nl_means_2d(img, h=0.1)
    out["nlm_score_F1"] = O.lne2d(out["nlm_out_h002"], "F1")
    seg = np.zeros((40, 48), dtype=np.int64)
    seg[3:9, 4:30] = 2
    seg[12:30, 6:12] = 5
    for r in range(14, 34):
        seg[r, 20 + (r - 14) // 2: 28 + (r - 14) // 2] = 9        # slanted cell
    seg[36, 40] = 11
    out["geo_seg"] = seg
    out["geo_labels"], out["geo_area"], out["geo_geometry"] = O.cell_geometry(seg)
    values = np.zeros((12, 3))
    values[1:] = rng.random((11, 3))
    out["paint_values"] = values
    out["paint_out"] = O.paint_labels(seg, values)
    path = os.path.join(HERE, "next_rows_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()

"""Generates tests/golden/*.npz from the UNMODIFIED reference stencils compiled into oracle/_ref
(oracle/build_ref.py) plus the measurement scripts' numpy blocks as transcribed in the oracle.
Run in the build container (where /root/reference exists); the fixtures are committed so that
the parity pin also holds on the GPU box.

    python tests/golden/make_golden.py
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import hipr_oracle as O, load_ref  # noqa: E402


def image(shape, seed):
    rng = np.random.default_rng(seed)
    grids = np.meshgrid(*[np.arange(n, dtype=np.float64) for n in shape], indexing="ij")
    img = sum(np.sin(g / (5.0 + 2 * k) + k) ** 2 for k, g in enumerate(grids)) + 0.05 * rng.random(shape)
    return img.astype(np.float32).astype(np.float64)


def main():
    r2, r3 = load_ref("neighbor2d"), load_ref("neighbor")
    assert r2 is not None and r3 is not None, "build oracle/_ref first (python oracle/build_ref.py)"
    out = {}
    # --- tables, read back out of the compiled reference: gather an index image
    P, R = 11, 9
    idx = np.arange(21 * 21, dtype=np.float64).reshape(21, 21)
    lp = r2.line_profile_2d_v2(idx, P, R)[5, 5]                     # (R, P) flat indices of the patch at (5,5)
    tab2 = np.stack([(lp // 21) - 5, (lp % 21) - 5], axis=-1).astype(np.int64)   # (R, P, 2) patch coords
    out["table2d_11_9"] = tab2
    idx3 = np.arange(21 ** 3, dtype=np.float64).reshape(21, 21, 21)
    lp3 = r3.line_profile_v2(idx3, 11, 9, 9)[5, 5, 5]               # (72, 11)
    tab3 = np.stack([lp3 // 441 - 5, (lp3 // 21) % 21 - 5, lp3 % 21 - 5], axis=-1).astype(np.int64)
    out["table3d_11_9_9"] = tab3
    print("2-D table sha256 (offsets from centre):", hashlib.sha256((tab2 - 5).tobytes()).hexdigest())
    print("3-D table sha256 (offsets from centre):", hashlib.sha256((tab3 - 5).tobytes()).hexdigest())
    # --- 2-D: padded input, literal gather (checksummed), the three epilogue flavours
    img = image((40, 52), 11)
    padded = np.pad(img / img.max(), 5, mode="edge")
    lp = r2.line_profile_2d_v2(padded, 11, 9)
    out["img2d"] = img
    out["lp2d_sha256"] = np.frombuffer(hashlib.sha256(lp.tobytes()).digest(), dtype=np.uint8)
    out["lp2d_corner"] = lp[:3, :4].copy()
    for f in ("F1", "F2", "F3"):
        out["score2d_" + f] = O.EPILOGUES[f](lp)
    # --- 3-D: literal gather checksum, me_v2, v3, fused flavours
    vol = image((6, 7, 9), 12)
    vp = np.pad(vol / vol.max(), 5, mode="edge")
    lp5 = r3.line_profile_v2(vp, 11, 9, 9)
    out["vol3d"] = vol
    out["lp3d_sha256"] = np.frombuffer(hashlib.sha256(lp5.tobytes()).digest(), dtype=np.uint8)
    out["me2_3d"] = r3.line_profile_memory_efficient_v2(vp, 11, 9, 9)
    out["score3d_ME2"] = O.epilogue_F2_dirs(out["me2_3d"])
    out["score3d_F2"] = O.epilogue_F2(lp5)
    out["score3d_F3"] = O.epilogue_F3(lp5)
    v3in = np.pad(image((14, 5, 4), 13), 5, mode="edge")
    out["v3_in"] = v3in
    out["v3_out"] = r3.line_profile_memory_efficient_v3(v3in, 11, 9, 9)[:6]   # rows whose reads stay in the buffer
    # --- whole 2-D path from a cube, and per-cell spectra
    rng = np.random.default_rng(14)
    lab = np.zeros((36, 44), dtype=np.int64)
    lab[4:14, 5:20] = 3
    lab[18:30, 8:18] = 7
    lab[20:33, 25:40] = 12
    cube = (0.02 + 0.05 * rng.random((36, 44, 95)) + (lab[..., None] > 0) * rng.random(95)[None, None, :]).astype(np.float32)
    out["cube"] = cube
    out["labels"] = lab
    out["cube_score_F1"] = O.neighbor2d_score(cube, "F1", lp_func=r2.line_profile_2d_v2)
    l, a, avg, norm = O.cell_spectra(lab, cube)
    out["cell_labels"], out["cell_area"], out["cell_avgint"], out["cell_avgint_norm"] = l, a, avg, norm
    np.savez_compressed(os.path.join(HERE, "reference_vectors.npz"), **out)
    print("wrote", os.path.join(HERE, "reference_vectors.npz"),
          os.path.getsize(os.path.join(HERE, "reference_vectors.npz")), "bytes")


if __name__ == "__main__":
    main()

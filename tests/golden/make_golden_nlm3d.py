"""Freezes the oracle's 3-D non-local-means restatement (oracle/hipr_oracle.py::denoise_nl_means_3d, the loop-for-loop
form of scikit-image's _fast_nl_means_denoising_3d) at the caller's parameters (patch 7, distance 11, h = 0.03,
bio/hiprfish_imaging_biofilm_analysis.py:454) on a small volume: minutes of numpy, so the GPU test reads the result
from nlm3d_vectors.npz instead of recomputing it.  PARITY UNPINNED (scikit-image is not installed): the vectors pin
the restatement, not scikit-image.      python tests/golden/make_golden_nlm3d.py"""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import hipr_oracle  # noqa: E402


def volume(shape, seed, noise=0.03):
    rng = np.random.default_rng(seed)
    x, y, z = np.meshgrid(*[np.arange(n) for n in shape], indexing="ij")
    v = (np.sin(x / 4.0) ** 2 + np.cos(y / 5.0) ** 2 + np.sin(z / 6.0 + 1.0) ** 2) / 3 + noise * rng.random(shape)
    return v / v.max()


if __name__ == "__main__":
    vol = volume((17, 18, 40), 3)
    out = hipr_oracle.denoise_nl_means_3d(vol, patch_distance=11, h=0.03)
    np.savez_compressed(os.path.join(HERE, "nlm3d_vectors.npz"), nlm3d_in=vol, nlm3d_out_h003=out)
    print("wrote nlm3d_vectors.npz", out.shape, float(np.abs(out - vol).max()))

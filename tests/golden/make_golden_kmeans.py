"""Freezes scikit-learn's answers for the 1-D k-means row so that the pin also holds where
scikit-learn is a different version: inputs are regenerated from seeds, the outputs of
sklearn.cluster.KMeans (version recorded) are stored.

    python tests/golden/make_golden_kmeans.py
"""
import hashlib
import os
import sys

import numpy as np
import sklearn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import hipr_oracle as O  # noqa: E402


def kmeans_case(seed, shape=(96, 128)):
    """A score-map-like image: dark background, bright cells, a few exact zeros; float32 representable."""
    rng = np.random.default_rng(seed)
    img = rng.normal(0.12, 0.04, shape)
    yy, xx = np.mgrid[:shape[0], :shape[1]]
    for _ in range(12):
        cy, cx = rng.integers(8, shape[0] - 8), rng.integers(8, shape[1] - 8)
        img += (0.45 + 0.25 * rng.random()) * np.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / (2 * (3 + 3 * rng.random()) ** 2))
    img = np.clip(img, 0, None)
    img[rng.random(shape) < 0.02] = 0.0
    return img.astype(np.float32)


CASES = [  # (seed, k, n_init, transform, eps, positive_only)
    (1, 2, 1, None, 0.0, False),
    (2, 3, 1, None, 0.0, False),
    (3, 2, 1, "log10", 1e-8, False),
    (4, 2, 1, "log", 1e-2, False),
    (5, 2, 1, None, 0.0, True),
    (6, 3, 1, None, 0.0, True),
    (7, 2, 3, None, 0.0, False),
]


def main():
    out = {"sklearn_version": np.array(sklearn.__version__)}
    for i, (seed, k, n_init, tr, eps, pos) in enumerate(CASES):
        img = kmeans_case(seed)
        cen, labels, mask, n_iter, inertia = O.kmeans1d_sklearn(img, k, 0, n_init, tr, eps, pos)
        out["case%d_centers" % i] = cen
        out["case%d_n_iter" % i] = np.array(n_iter)
        out["case%d_inertia" % i] = np.array(inertia)
        out["case%d_labels_sha256" % i] = np.frombuffer(hashlib.sha256(labels.astype(np.int32).tobytes()).digest(), dtype=np.uint8)
        out["case%d_mask_sum" % i] = np.array(int(mask.sum()))
        out["case%d_img_sha256" % i] = np.frombuffer(hashlib.sha256(img.tobytes()).digest(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "kmeans_vectors.npz"), **out)
    print("wrote kmeans_vectors.npz with", len(CASES), "cases, scikit-learn", sklearn.__version__)


if __name__ == "__main__":
    main()

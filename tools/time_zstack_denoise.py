"""The z-stack chain of bio/..._analysis.py:452-462 from a host cube (hipr_neighbor3d_host_denoise): sum -> /max ->
3-D NL-means (h = 0.03, distance 11) -> 72-direction stencil -> float32 score.  python tools/time_zstack_denoise.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "hiprfish-image-analysis_b200")]
import numpy as np
import torch
import hipr_b200
from hipr_b200 import ops, synth
X, Y, Z = 256, 264, 54
cube = synth.make_volume_cube(X, Y, Z, 95, seed=5, device="cuda")
host = ops.pinned_empty(tuple(cube.shape), np.float32)
torch.from_numpy(host).copy_(cube.cpu())
del cube
score = ops.pinned_empty((X, Y, Z), np.float32)
for h in (None, 0.03):
    ops.neighbor3d_score_host(host, "ME2", out=score, denoise_h=h)
    t0 = time.perf_counter()
    ops.neighbor3d_score_host(host, "ME2", out=score, denoise_h=h)
    ms = (time.perf_counter() - t0) * 1e3
    print("%dx%dx%dx95 host cube (%.2f GB), denoise_h=%s: %.1f ms  %.2f Mvox/s  mean score %.4f" % (X, Y, Z, host.nbytes / 1e9, h, ms, X * Y * Z / ms / 1e3, float(score.mean())))

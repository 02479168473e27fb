"""hipr_neighbor2d_host_batch against one hipr_neighbor2d_host(_denoise) call per FOV, 2048^2 x 95 pinned cubes."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "hiprfish-image-analysis_b200")]
import numpy as np
import torch
import hipr_b200
from hipr_b200 import ops, synth
cube = synth.make_fov(2048, 2048, 95, fov_index=0)[0]
host = ops.pinned_empty(tuple(cube.shape), np.float32)
torch.from_numpy(host).copy_(cube)
score = ops.pinned_empty((2048, 2048), np.float32)
for h in (None, 0.02):
    ops.neighbor2d_score_host(host, "F1", out=score, denoise_h=h)
    t0 = time.perf_counter()
    for _ in range(4):
        ops.neighbor2d_score_host(host, "F1", out=score, denoise_h=h)
    single = (time.perf_counter() - t0) / 4 * 1e3
    ops.neighbor2d_score_host_batch([host] * 2, "F1", denoise_h=h, out=[score] * 2)
    t0 = time.perf_counter()
    ops.neighbor2d_score_host_batch([host] * 6, "F1", denoise_h=h, out=[score] * 6)
    batch = (time.perf_counter() - t0) / 6 * 1e3
    print("denoise_h=%s: one call per FOV %.2f ms, batch of 6 %.2f ms per FOV" % (h, single, batch))

import sys, time, numpy as np, torch
sys.path.insert(0, 'hiprfish-image-analysis_b200')
import hipr_b200
from hipr_b200 import synth, ops
cube = synth.make_fov(2048, 2048, 95, fov_index=0)[0].numpy()      # pageable numpy array
pinned = ops.pinned_empty(cube.shape, np.float32); pinned[...] = cube
for name, arr in (("pageable", cube), ("pinned", pinned)):
    hipr_b200.neighbor2d_score_host(arr, "F1")
    t0 = time.perf_counter()
    for _ in range(3):
        hipr_b200.neighbor2d_score_host(arr, "F1")
    dt = (time.perf_counter() - t0) / 3
    print("%s: %.1f ms per FOV, %.1f GB/s, %.0f Mpix/s" % (name, dt * 1e3, cube.nbytes / dt / 1e9, 2048 * 2048 / dt / 1e6))

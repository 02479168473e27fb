"""Device-resident timings of the 3-D literal outputs: line_profile_v2 (a2, 6336 B/voxel in float64),
line_profile_memory_efficient_v2 (a3, 576 B/voxel float64; fixed-point float32 form 288 B/voxel)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "hiprfish-image-analysis_b200")]
import torch
import hipr_b200
from hipr_b200 import ops


def timed(fn, n=5):
    for _ in range(2):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n


g = torch.Generator(device="cuda").manual_seed(0)
X, Y, Z = 96, 128, 64
vp = torch.rand((X + 10, Y + 10, Z + 10), generator=g, device="cuda", dtype=torch.float64)
nv = X * Y * Z
for dt in (torch.float64, torch.float32):
    v = vp.to(dt)
    t = timed(lambda: ops.line_profile_3d(v, 11, 9, 9))
    b = 792 * v.element_size() + v.element_size()
    print("line_profile_v2 %s %dx%dx%d: %.3f ms  %.0f GB/s (%d B/voxel)" % (dt, X, Y, Z, t, nv * b / t / 1e6, b))
X, Y, Z = 256, 256, 64
vp = torch.rand((X + 10, Y + 10, Z + 10), generator=g, device="cuda", dtype=torch.float64)
nv = X * Y * Z
t = timed(lambda: ops.lne3d_dirs(vp, 11, 9, 9, padded=True))
print("me_v2 float64 %dx%dx%d: %.3f ms  %.1f Mvox/s  %.0f GB/s (584 B/voxel)" % (X, Y, Z, t, nv / t / 1e3, nv * 584 / t / 1e6))

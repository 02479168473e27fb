"""Times the 3-D non-local-means kernel.  python tools/time_nlm3d.py [X Y Z]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "hiprfish-image-analysis_b200")]
import torch
import hipr_b200
from hipr_b200 import ops
X, Y, Z = [int(v) for v in sys.argv[1:4]] if len(sys.argv) > 3 else (128, 132, 54)
g = torch.Generator(device="cuda").manual_seed(0)
v = torch.rand((X, Y, Z), generator=g, device="cuda", dtype=torch.float64) * 0.05 + 0.5
ops.denoise_nl_means(v, h=0.03)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); a.record()
out = ops.denoise_nl_means(v, h=0.03)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b)
print("noise volume (every pair inside the cutoff)  nlm3d %dx%dx%d: %.1f ms  %.3f Mvox/s  %.1f G voxel-shifts/s" % (X, Y, Z, ms, X * Y * Z / ms / 1e3, X * Y * Z * 12167 / ms / 1e6))

# a z-stack-like volume: the normalised channel sums of the synthetic cells (most pairs across a cell border are cut off)
from hipr_b200 import synth
cube = synth.make_volume_cube(X, Y, Z, 95, seed=3, device="cuda")
s = ops.channel_sum(cube, None, normalize=True, dtype=torch.float64)
del cube
ops.denoise_nl_means(s, h=0.03)
torch.cuda.synchronize(); a.record()
out = ops.denoise_nl_means(s, h=0.03)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b)
print("synthetic z-stack sums  nlm3d %dx%dx%d: %.1f ms  %.3f Mvox/s" % (X, Y, Z, ms, X * Y * Z / ms / 1e3))

"""Times the 3-D fixed-point stencil (+ refinement) on a resident sum volume.  python tools/time3d.py [X Y Z]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "hiprfish-image-analysis_b200")]
import torch
import hipr_b200
from hipr_b200 import ops, synth
X, Y, Z = [int(v) for v in sys.argv[1:4]] if len(sys.argv) > 3 else (256, 256, 64)
cube = torch.empty((X, Y, Z, 95), dtype=torch.float32, device="cuda")
full = synth.make_volume_cube(X, 8, Z, 95, seed=99, device="cuda")
for y in range(0, Y, 8):
    cube[:, y:y + 8] = full[:, :8] + 0.01 * torch.rand((X, 8, Z, 1), device="cuda")
s, mk = ops.channel_sum(cube, None, normalize=False, dtype=torch.float64, return_max=True)
for fl in ("ME2",):
    for _ in range(2):
        ops.lne3d_fixed(s, fl, maxkey=mk)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(5):
        ops.lne3d_fixed(s, fl, maxkey=mk)
    b.record(); torch.cuda.synchronize()
    print(fl, X, Y, Z, "ms", a.elapsed_time(b) / 5, "Mvox/s", X * Y * Z / (a.elapsed_time(b) / 5) / 1e3)

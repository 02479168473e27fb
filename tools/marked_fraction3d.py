"""Diagnostic: fraction of voxels the 3-D fixed-point stencil marks for float64 refinement.
Run as: HIPR_LNE3D_REFINE=2 python tools/marked_fraction3d.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "hiprfish-image-analysis_b200")]
import torch
import hipr_b200
from hipr_b200 import synth
assert os.environ.get("HIPR_LNE3D_REFINE") == "2"
X, Y, Z, C = 256, 256, 64, 95
dev = "cuda"
cube = torch.empty((X, Y, Z, C), dtype=torch.float32, device=dev)
full = synth.make_volume_cube(X, 8, Z, C, seed=99, device=dev)
for y in range(0, Y, 8):
    cube[:, y:y + 8] = full[:, :8] + 0.01 * torch.rand((X, 8, Z, 1), device=dev)
for fl in ("ME2", "F2", "F3"):
    s = hipr_b200.neighbor3d_score(cube, fl)
    m = (s == -2.0)
    print(fl, "marked fraction %.5f" % float(m.float().mean()), "by z-plane (first 8):", [round(float(v), 4) for v in m.float().mean(dim=(0, 1))[:8]])

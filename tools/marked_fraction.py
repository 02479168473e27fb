"""Diagnostic: fraction of pixels the fixed-point stencil marks for float64 refinement (HIPR_LNE2D_REFINE=2 leaves the
sentinels in the score map).  Run as: HIPR_LNE2D_REFINE=2 python tools/marked_fraction.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "hiprfish-image-analysis_b200")]
import torch
import hipr_b200
from hipr_b200 import synth
assert os.environ.get("HIPR_LNE2D_REFINE") == "2"
for side, idx in ((512, 0), (1024, 0), (2048, 0), (2048, 1)):
    cube = synth.make_fov(side, side, 95, fov_index=idx, device="cuda")[0]
    for fl in ("F1", "F2", "F3"):
        s = hipr_b200.neighbor2d_score(cube, fl)
        print(side, idx, fl, "marked fraction %.5f" % float((s == -2.0).float().mean()))

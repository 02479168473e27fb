import sys, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'hiprfish-image-analysis_b200')
import hipr_b200
from hipr_b200 import synth, ops
from oracle import hipr_oracle as o
cube = synth.make_fov(96, 128, 95, fov_index=5)[0]
s = cube.numpy().astype(np.float64).sum(axis=2); s = s / s.max()
den = o.denoise_nl_means_2d(s, h=0.02)
want = o.lne2d(den, "F1")
sg = ops.channel_sum(cube.cuda(), None, normalize=True, dtype=torch.float64)
print("sum err", np.abs(sg.cpu().numpy() / s - 1).max())
dg = ops.denoise_nl_means(sg, h=0.02)
print("nlm err", np.abs(dg.cpu().numpy() / den - 1).max())
for name, got in (("fixed tile-local", ops.lne2d_fixed(dg, "F1")), ("fixed global", ops.lne2d_fixed(dg, "F1", range_keys=ops.image_range(dg))),
                  ("float64 stencil", ops.lne2d(dg, "F1")), ("float64 stencil on oracle den", ops.lne2d(torch.from_numpy(den).cuda(), "F1")),
                  ("fixed on oracle den", ops.lne2d_fixed(torch.from_numpy(den).cuda(), "F1"))):
    g = got.cpu().numpy().astype(np.float64)
    d = np.abs(g - want)
    bad = d > 5e-7 + 1e-5 * np.abs(want)
    print("%-32s max abs %.3e  viol %d  worst want %.6e got %.6e" % (name, d.max(), bad.sum(), want.flat[d.argmax()], g.flat[d.argmax()]))

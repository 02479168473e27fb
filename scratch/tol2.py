import sys, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'hiprfish-image-analysis_b200')
import hipr_b200
from hipr_b200 import synth, ops
from oracle import hipr_oracle as O
big = synth.make_fov(768, 256, 95, fov_index=31)[0]
want = O.neighbor2d_score(big.numpy(), "F1")
s, mk = ops.channel_sum(big.cuda(), None, normalize=False, dtype=torch.float64, return_max=True)
for name, got in (("pipeline (tile-local)", hipr_b200.neighbor2d_score(big.cuda(), "F1")),
                  ("global range", ops.lne2d_fixed(s, "F1", range_keys=mk)),
                  ("tile-local direct", ops.lne2d_fixed(s, "F1"))):
    g = got.cpu().numpy().astype(np.float64)
    d = np.abs(g - want); ex = d - 1e-5 * np.abs(want); i = np.argmax(ex)
    r, c = np.unravel_index(i, d.shape)
    print("%-24s max abs %.3e worst excess %.3e at (%d,%d) want %.6f got %.6f" % (name, d.max(), ex.max(), r, c, want[r, c], g[r, c]))
sn = s.cpu().numpy()
r, c = 291, None

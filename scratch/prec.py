import sys, torch, numpy as np
sys.path.insert(0, 'hiprfish-image-analysis_b200'); sys.path.insert(0, '.')
from hipr_b200 import ops, synth
from oracle import hipr_oracle as O
cube, _, _ = synth.make_fov(96, 160, 95, fov_index=3)
want = O.neighbor2d_score(cube.numpy(), "F1")
c = cube.cuda()
s, mk = ops.channel_sum(c, None, normalize=False, dtype=torch.float64, return_max=True)
s_ref = cube.numpy().astype(np.float64).sum(2)
print("sum max rel err", np.abs(s.cpu().numpy() - s_ref).max() / s_ref.max())
def rep(name, got):
    d = np.abs(got - want)
    bad = d > 1e-5 * np.abs(want) + 2e-7
    print("%-12s max abs %.3e  n_bad(2e-7) %d  n_bad(1e-6) %d  max rel (want>1e-3) %.3e" % (name, d.max(), bad.sum(), (d > 1e-5 * np.abs(want) + 1e-6).sum(), (d / np.maximum(np.abs(want), 1e-3)).max()))
rep("global", ops.lne2d_fixed(s, "F1", range_keys=mk).cpu().numpy())
import ctypes as C
from hipr_b200._lib import lib, check
from hipr_b200 import tables
tab = tables.line_table_2d(11, 9)
out = torch.empty((96, 160), dtype=torch.float32, device='cuda')
check(lib().hipr_lne2d_q(C.c_void_p(s.data_ptr()), 96, 160, 160, 0, 1, 11, 9, tab.ctypes.data_as(C.c_void_p), 1, None, C.c_void_p(out.data_ptr()), None), "q")
torch.cuda.synchronize()
rep("local", out.cpu().numpy())
rep("banded", ops.neighbor2d_pipeline(c, "F1")[0].cpu().numpy())
rep("fused", ops.neighbor2d_fused(c, "F1").cpu().numpy())
rep("f64", ops.lne2d(s / s.max(), "F1").cpu().numpy())

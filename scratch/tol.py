import sys, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'hiprfish-image-analysis_b200')
import hipr_b200
from hipr_b200 import synth
from oracle import hipr_oracle as O, load_ref
ref = load_ref("neighbor2d")
lp = ref.line_profile_2d_v2 if ref is not None else None
worst = []
for idx in range(10):
    cube = synth.make_fov(512, 512, 95, fov_index=100 + idx)[0]
    want = O.neighbor2d_score(cube.numpy(), "F1", lp_func=lp)
    got = hipr_b200.neighbor2d_score(cube.cuda(), "F1").cpu().numpy().astype(np.float64)
    got64 = hipr_b200.neighbor2d_score(cube.cuda(), "F1", dtype=torch.float64).cpu().numpy()
    d = np.abs(got - want)
    excess = d - 1e-5 * np.abs(want)
    i = np.argmax(excess)
    worst.append(excess.max())
    print(idx, "max abs %.3e  max excess over rtol %.3e at want=%.4f  n>5e-7: %d  n>1e-6: %d | float64 stencil max abs %.2e"
          % (d.max(), excess.max(), want.flat[i], (excess > 5e-7).sum(), (excess > 1e-6).sum(), np.abs(got64 - want).max()), flush=True)
print("worst excess", max(worst))

"""Times K0 (registration paste + channel stack [+ flat field] + channel sum) on a 2048^2 x 95 FOV and checks it
against a torch restatement of the paste.  python tools/time_register.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "hiprfish-image-analysis_b200")]
import torch
import hipr_b200
from hipr_b200 import ops

H = W = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
chans = (32, 23, 20, 14, 6)
shifts = [(0, 0), (3, -2), (-4, 1), (2, 5), (-1, -3)]
g = torch.Generator(device="cuda").manual_seed(1)
stacks = [torch.rand((H, W, c), generator=g, device="cuda") + 0.1 for c in chans]
cal = torch.rand((H, W, 95), generator=g, device="cuda") + 0.5


def timed(fn, n=20):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n


cube, s, _ = ops.register_stacks(stacks, shifts)
want = torch.zeros((H, W, 95), device="cuda")
o = 0
for st, (dr, dc), c in zip(stacks, shifts, chans):
    want[max(dr, 0):H + min(dr, 0), max(dc, 0):W + min(dc, 0), o:o + c] = st[max(-dr, 0):H + min(-dr, 0), max(-dc, 0):W + min(-dc, 0)]
    o += c
print("cube equal:", bool(torch.equal(cube, want)), " sum max rel err:", float(((s - want.double().sum(2)).abs() / want.double().sum(2)).max()))
cube_c, s_c, _ = ops.register_stacks(stacks, shifts, cal)
print("flat-field cube equal (IEEE float32 quotient):", bool(torch.equal(cube_c, want / cal)),
      " sum max rel err:", float(((s_c - (want.double() / cal.double()).sum(2)).abs() / s_c.abs()).max()))
# a wide-range check of the quotient path: magnitudes from 1e-30 to 1e30, zeros, denormals
wide = [st * torch.exp(torch.empty_like(st).uniform_(-69, 69)) for st in stacks]
wide[0][::7, ::5] = 0.0
wide[1][::3, ::11] = 1e-41
calw = cal * torch.exp(torch.empty_like(cal).uniform_(-69, 69))
cube_w, _, _ = ops.register_stacks(wide, shifts, calw)
wantw = torch.zeros((H, W, 95), device="cuda")
o = 0
for st, (dr, dc), c in zip(wide, shifts, chans):
    wantw[max(dr, 0):H + min(dr, 0), max(dc, 0):W + min(dc, 0), o:o + c] = st[max(-dr, 0):H + min(-dr, 0), max(-dc, 0):W + min(-dc, 0)]
    o += c
print("wide-range quotient equal:", bool(torch.equal(cube_w, wantw / calw)))
del wide, calw, cube_w, wantw, cube_c
npix = H * W
t = timed(lambda: ops.register_stacks(stacks, shifts))
print("register_stacks           %.3f ms  %.0f GB/s (768 B/px)" % (t, npix * 768 / t / 1e6))
t = timed(lambda: ops.register_stacks(stacks, shifts, cal))
print("register_stacks+flatfield %.3f ms  %.0f GB/s (1148 B/px)" % (t, npix * 1148 / t / 1e6))

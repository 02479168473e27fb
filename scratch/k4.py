import sys, torch, time
sys.path.insert(0, 'hiprfish-image-analysis_b200')
from hipr_b200 import ops
dev = torch.device('cuda')
def timeit(fn, n=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
X, Y, Z = 512, 512, 64
vol = torch.rand((X, Y, Z), device=dev, dtype=torch.float32) * 0.1 + torch.sin(torch.arange(Z, device=dev) / 5.0) ** 2
nv = X * Y * Z
for fl in ("ME2", "F2", "F3"):
    ms = timeit(lambda: ops.lne3d(vol, fl))
    print("lne3d %s f32 %dx%dx%d: %.3f ms  %.1f Mvox/s" % (fl, X, Y, Z, ms, nv / ms / 1e3))
ms = timeit(lambda: ops.lne3d(vol.double(), "ME2"), n=2, warm=1)
print("lne3d ME2 f64: %.3f ms  %.1f Mvox/s" % (ms, nv / ms / 1e3))
volp = torch.nn.functional.pad(vol[None, None], (5, 5, 5, 5, 5, 5), mode='replicate')[0, 0].contiguous()
ms = timeit(lambda: ops.lne3d_dirs(volp), n=3, warm=1)
print("lne3d_dirs (me_v2 output, 288 B/vox) f32: %.3f ms  %.1f Mvox/s  %.0f GB/s written" % (ms, nv / ms / 1e3, nv * 288 / ms / 1e6))
sub = volp[:138, :138].contiguous()
ms = timeit(lambda: ops.line_profile_3d(sub, 11, 9, 9), n=3, warm=1)
nsub = 128 * 128 * 64
print("line_profile_3d literal f32 128x128x64: %.3f ms %.1f Mvox/s %.0f GB/s written" % (ms, nsub / ms / 1e3, nsub * 3168 / ms / 1e6))
# 2-D literal
img = torch.rand((2058, 2058), device=dev, dtype=torch.float64)
ms = timeit(lambda: ops.line_profile_2d(img, 11, 9), n=5, warm=2)
print("line_profile_2d literal f64 2048^2: %.3f ms %.1f Mpix/s %.0f GB/s written" % (ms, 2048 * 2048 / ms / 1e3, 2048 * 2048 * 792 / ms / 1e6))
ms = timeit(lambda: ops.line_profile_2d(img.float(), 11, 9), n=5, warm=2)
print("line_profile_2d literal f32 2048^2 (incl cast): %.3f ms" % ms)
# full c4 chansum
cube = torch.empty((1024, 1024, 64, 95), device=dev, dtype=torch.float32)
cube.uniform_(0, 1)
ms = timeit(lambda: ops.channel_sum(cube, None, normalize=False, dtype=torch.float32, return_max=True), n=3, warm=1)
print("channel_sum c4 (25.5 GB): %.3f ms %.0f GB/s" % (ms, cube.numel() * 4 / ms / 1e6))
s, mk = ops.channel_sum(cube, None, normalize=False, dtype=torch.float32, return_max=True)
del cube
ms = timeit(lambda: ops.lne3d_fixed(s, "ME2", maxkey=mk), n=2, warm=1)
print("lne3d ME2 f32 c4 1024x1024x64: %.3f ms %.1f Mvox/s" % (ms, s.numel() / ms / 1e3))
for fl in ("ME2", "F2", "F3"):
    ms = timeit(lambda: ops.lne3d_fixed(vol, fl))
    print("lne3d_fixed %s %dx%dx%d: %.3f ms  %.1f Mvox/s" % (fl, X, Y, Z, ms, nv / ms / 1e3))
ms = timeit(lambda: ops.lne3d_fixed(vol.double(), "ME2"))
print("lne3d_fixed ME2 from f64: %.3f ms" % ms)
ms = timeit(lambda: ops.lne3d_fixed(volp, "ME2", padded=True, dirs_only=True), n=3, warm=1)
print("lne3d_fixed dirs: %.3f ms %.0f GB/s written" % (ms, nv * 288 / ms / 1e6))

import sys, torch, time
sys.path.insert(0, 'hiprfish-image-analysis_b200')
from hipr_b200 import ops, synth
dev = torch.device('cuda')
def timeit(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for H in (128, 256, 512, 2048):
    cube = torch.rand((H, 2048, 95), device=dev)
    ms = timeit(lambda: ops.channel_sum(cube, None, normalize=False, dtype=torch.float64, return_max=True))
    gb = cube.numel() * 4 / 1e9
    print("K1 H=%d  %.1f MB  %.4f ms  %.0f GB/s" % (H, gb * 1e3, ms, gb / ms * 1e3))
# fused on a small (L2 resident) image
for H in (256, 512, 2048):
    cube = synth.make_fov(H, 2048, 95, device=dev)[0]
    ms = timeit(lambda: ops.neighbor2d_fused(cube, "F1"))
    gb = cube.numel() * 4 / 1e9
    print("fused H=%d %.4f ms %.0f GB/s (algorithmic)" % (H, ms, gb / ms * 1e3))
    s, mk = ops.channel_sum(cube, None, normalize=False, dtype=torch.float64, return_max=True)
    ms = timeit(lambda: ops.lne2d_fixed(s, "F1", range_keys=mk))
    print("K3q   H=%d %.4f ms  %.1f Mpix/s" % (H, ms, H * 2048 / ms / 1e3))

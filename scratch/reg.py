import sys, torch, numpy as np
sys.path.insert(0, 'hiprfish-image-analysis_b200')
from hipr_b200 import ops
dev = torch.device('cuda')
def timeit(fn, n=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
H = W = 2048
chans = (32, 23, 20, 14, 6)
stacks = [torch.rand((H, W, c), device=dev) for c in chans]
shifts = [(0, 0), (3, -2), (-4, 5), (1, 1), (-1, -7)]
cal = torch.rand((H, W, 95), device=dev) + 0.5
ms = timeit(lambda: ops.register_stacks(stacks, shifts))
print("register (380 R + 388 W B/px): %.4f ms  %.0f GB/s" % (ms, H * W * 768 / ms / 1e6))
ms = timeit(lambda: ops.register_stacks(stacks, shifts, return_cube=False))
print("register, sums only (380 R): %.4f ms  %.0f GB/s" % (ms, H * W * 388 / ms / 1e6))
ms = timeit(lambda: ops.register_stacks(stacks, shifts, cal))
print("register + flat field (760 R + 388 W B/px): %.4f ms  %.0f GB/s" % (ms, H * W * 1148 / ms / 1e6))

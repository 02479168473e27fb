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
H = W = 2048; C = 95
cube = torch.rand((H, W, C), device=dev)
cal = torch.rand((H, W, C), device=dev) + 0.5
ms = timeit(lambda: ops.channel_sum(cube, cal, normalize=False, dtype=torch.float64, return_max=True))
print("channel_sum with flat-field (2 streams, 760 B/px): %.4f ms  %.0f GB/s" % (ms, H * W * 760 / ms / 1e6))
ms = timeit(lambda: ops.channel_sum(cube, None, normalize=False, dtype=torch.float64, return_max=True))
print("channel_sum: %.4f ms  %.0f GB/s" % (ms, H * W * 380 / ms / 1e6))
got = ops.channel_sum(cube[:256], cal[:256], normalize=False, dtype=torch.float64).cpu().numpy()
want = (cube[:256].cpu().numpy().astype(np.float64) / cal[:256].cpu().numpy().astype(np.float64)).sum(axis=2)
print("max rel err vs numpy float64 divide:", np.abs(got / want - 1).max())

"""Times the strict drop-in gather (line_profile_2d_v2) on a 2048^2 image: device-resident (write roofline, 792 + 8 B/px),
host numpy -> numpy through hipr_line_profile_2d_host (pageable and page-locked output), and the round-1 route
(.cuda() / .cpu().numpy()).  python tools/time_gather.py [N]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "hiprfish-image-analysis_b200")]
import numpy as np
import torch
import hipr_b200
from hipr_b200 import ops

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
a = np.random.default_rng(0).random((N + 10, N + 10))
dev = torch.from_numpy(a).cuda()
for dt in (torch.float64, torch.float32):
    d = dev.to(dt)
    for _ in range(2):
        o = ops.line_profile_2d(d, 11, 9)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(5):
        o = ops.line_profile_2d(d, 11, 9)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    bpp = 100 * d.element_size()
    print("device %s: %.3f ms  %.0f GB/s (%d B/px)" % (dt, ms, N * N * bpp / ms / 1e6, bpp))
    del o
torch.cuda.empty_cache()
for name, kw in (("pageable out", {}), ("page-locked out", {"pinned": True})):
    out = ops.line_profile_2d_host(a, 11, 9, **kw)          # warm-up: workspace, staging ring, pinned cache
    t0 = time.perf_counter()
    for _ in range(3):
        out = ops.line_profile_2d_host(a, 11, 9, out=out)
    ms = (time.perf_counter() - t0) / 3 * 1e3
    print("host -> host, %s: %.1f ms  %.1f GB/s of output, %.2f Mpix/s" % (name, ms, out.nbytes / ms / 1e6, N * N / ms / 1e3))
    t0 = time.perf_counter()
    out2 = ops.line_profile_2d_host(a, 11, 9, **kw)
    print("   including the allocation of the result: %.1f ms" % ((time.perf_counter() - t0) * 1e3))
    del out, out2
t0 = time.perf_counter()
old = ops.line_profile_2d(torch.from_numpy(a).cuda(), 11, 9).cpu().numpy()
print("round-1 route (.cuda() / .cpu().numpy()): %.1f ms" % ((time.perf_counter() - t0) * 1e3))

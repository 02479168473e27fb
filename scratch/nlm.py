import sys, torch
sys.path.insert(0, 'hiprfish-image-analysis_b200')
from hipr_b200 import ops, synth
dev = torch.device('cuda')
def timeit(fn, n=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
cube = synth.make_fov(2048, 2048, 95, fov_index=0, device=dev)[0]
s = ops.channel_sum(cube, None, normalize=True, dtype=torch.float64)
ms = timeit(lambda: ops.denoise_nl_means(s, h=0.02))
print("nlm 2048^2 f64 (7, 11): %.3f ms  %.1f Mpix/s" % (ms, 2048 * 2048 / ms / 1e3))
ms = timeit(lambda: ops.denoise_nl_means(s.float(), h=0.02))
print("nlm 2048^2 f32 in: %.3f ms" % ms)
ms = timeit(lambda: ops.neighbor2d_score(cube, "F1", denoise_h=0.02))
print("chain sum -> nlm -> F1 score: %.3f ms" % ms)

import torch
x = torch.empty(3_322_000_000 // 8, dtype=torch.float64, device="cuda")
for f, name in ((lambda: x.zero_(), "zero_ (cudaMemset)"), (lambda: x.fill_(1.5), "fill_ (torch kernel)")):
    for _ in range(3): f()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(10): f()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    print("%s: %.3f ms  %.0f GB/s" % (name, ms, x.numel() * 8 / ms / 1e6))

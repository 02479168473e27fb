import sys, torch
sys.path.insert(0, 'hiprfish-image-analysis_b200')
from hipr_b200 import ops, synth
dev = torch.device('cuda')
cube, lab, L = synth.make_fov(2048, 2048, 95, device=dev)
sums = torch.zeros((L + 1, 95), dtype=torch.float64, device=dev); counts = torch.zeros(L + 1, dtype=torch.int32, device=dev)
def run(): ops.cell_spectra_accumulate(cube, lab, L, sums, counts)
best = 1e9
for rep in range(5):
    for _ in range(5): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(100): run()
    e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 100)
print("K5 accumulate: %.4f ms  %.2f M cells/s" % (best, L / best / 1e3))

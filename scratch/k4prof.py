import sys, torch
sys.path.insert(0, 'hiprfish-image-analysis_b200')
from hipr_b200 import ops
dev = torch.device('cuda')
X, Y, Z = 512, 512, 64
vol = (torch.rand((X, Y, Z), device=dev, dtype=torch.float64) * 0.1 + torch.sin(torch.arange(Z, device=dev) / 5.0) ** 2)
mk = ops.image_range(vol)
for _ in range(3):
    ops.lne3d_fixed(vol, "ME2", maxkey=mk)
torch.cuda.synchronize()

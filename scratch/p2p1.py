import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, '.'); sys.path.insert(0, 'hiprfish-image-analysis_b200')
os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT="29555")
torch.cuda.set_device(0)
dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
from hipr_b200 import sharding, synth
from oracle import hipr_oracle as O
Wm = 256
big = synth.make_fov(768, Wm, 95, fov_index=31)[0]
slab = sharding.P2PMosaicSlab(768, Wm)
for fl in ("F1", "F2"):
    for rep in range(3):
        got = slab.score(big.cuda(), fl, bands=4)
    torch.cuda.synchronize()
    print("scored", fl, flush=True)
    want = O.neighbor2d_score(big.numpy(), fl)
    d = np.abs(got.cpu().numpy() - want)
    bad = d > 5e-7 + 1e-5 * np.abs(want)
    print(fl, "max abs", d.max(), "viol", bad.sum(), "rows with viol", np.unique(np.nonzero(bad)[0])[:20], flush=True)
    ref = slab.score(big.cuda(), fl, bands=0)
    print("banded == unbanded:", torch.equal(got, ref), float((got - ref).abs().max()))
slab.check_peers(); slab.close()
dist.destroy_process_group()

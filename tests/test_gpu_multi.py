"""Two-GPU NCCL run of the split-mosaic path (BASELINE config 5 in miniature) with the real CUDA
hooks.  Skipped on a single-GPU box; the same logic runs under gloo in tests/test_sharding.py."""
import os
import sys

import numpy as np
import pytest

from conftest import PKG, ROOT

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, tmp):
    import torch
    import torch.distributed as dist
    for p in (ROOT, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from hipr_b200 import sharding, synth
        from oracle import hipr_oracle as O
        Hm, Wm = 160, 256
        cube, labels, L = synth.make_fov(Hm, Wm, 95, fov_index=13)
        r0, r1 = sharding.slab_bounds(Hm, rank, world)
        slab = sharding.MosaicSlab()
        got = slab.score(cube[r0:r1].cuda(), "F1").cpu().numpy()
        want = O.neighbor2d_score(cube.numpy(), "F1")[r0:r1]
        np.testing.assert_allclose(got, want, rtol=1e-5, atol=5e-7)
        lab_out, area, avg, _ = slab.cell_spectra(cube[r0:r1].cuda(), labels[r0:r1].cuda(), L)
        wl, wa, wavg, _ = O.cell_spectra(labels.numpy(), cube.numpy())
        assert np.array_equal(lab_out.cpu().numpy(), wl) and np.array_equal(area.cpu().numpy(), wa)
        np.testing.assert_allclose(avg.cpu().numpy(), wavg, rtol=1e-5)
        open(os.path.join(tmp, "ok%d" % rank), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_mosaic_two_gpus_nccl(torch_cuda, tmp_path):
    import torch.multiprocessing as mp
    if torch_cuda.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    port = 29700 + (os.getpid() % 1000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert all((tmp_path / ("ok%d" % r)).exists() for r in range(2))


def _p2p_worker(rank, world, port, tmp):
    import torch
    import torch.distributed as dist
    for p in (ROOT, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from hipr_b200 import sharding, synth
        from oracle import hipr_oracle as O
        Hm, Wm = 161, 256                      # uneven slabs (81 / 80 rows)
        r0, r1 = sharding.slab_bounds(Hm, rank, world)
        p2p = sharding.P2PMosaicSlab(r1 - r0, Wm)
        nccl = sharding.MosaicSlab()
        for step in range(4):                  # both parities, reused buffers
            cube = synth.make_fov(Hm, Wm, 95, fov_index=20 + step)[0]
            mine = cube[r0:r1].cuda()
            got = p2p.score(mine, "F1")
            ref = nccl.score(mine, "F1")
            assert torch.equal(got, ref), "peer-memory exchange differs from the NCCL exchange at step %d" % step
            if step == 0:
                want = O.neighbor2d_score(cube.numpy(), "F1")[r0:r1]
                np.testing.assert_allclose(got.cpu().numpy(), want, rtol=1e-5, atol=5e-7)
        got3 = p2p.score(mine, "F3")            # F3 consumes the exchanged global range
        assert torch.equal(got3, nccl.score(mine, "F3"))
        p2p.check_peers()
        p2p.close()
        open(os.path.join(tmp, "p2p%d" % rank), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_mosaic_two_gpus_peer_memory(torch_cuda, tmp_path):
    """The library's own halo / range exchange over NVLink peer memory == the NCCL path, bit for bit."""
    import torch.multiprocessing as mp
    if torch_cuda.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    port = 29800 + (os.getpid() % 1000)
    mp.spawn(_p2p_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert all((tmp_path / ("p2p%d" % r)).exists() for r in range(2))

"""Two-GPU NCCL run of the split-mosaic path (BASELINE config 5 in miniature) with the real CUDA
hooks.  Skipped on a single-GPU box; the same logic runs under gloo in tests/test_sharding.py."""
import os
import sys

import numpy as np
import pytest

from conftest import PKG, ROOT

pytestmark = pytest.mark.gpu


def _run_guarded(fn, rank, world, port, tmp):
    """A rank that fails writes its traceback and exits hard: it must not sit in destroy_process_group (or leave its
    peer in a collective) until the outer timeout."""
    import traceback
    try:
        fn(rank, world, port, tmp)
    except BaseException:
        open(os.path.join(tmp, "fail%d.txt" % rank), "w").write(traceback.format_exc())
        os._exit(1)


def _worker_body(rank, world, port, tmp):
    import torch
    import torch.distributed as dist
    for p in (ROOT, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    if True:
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
    dist.destroy_process_group()


def _worker(rank, world, port, tmp):
    _run_guarded(_worker_body, rank, world, port, tmp)


def _spawn(fn, tmp_path, port):
    import torch.multiprocessing as mp
    try:
        mp.spawn(fn, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    except Exception:
        msgs = [p.read_text() for p in sorted(tmp_path.glob("fail*.txt"))]
        pytest.fail("a rank failed:\n" + "\n".join(msgs))


def test_mosaic_two_gpus_nccl(torch_cuda, tmp_path):
    if torch_cuda.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _spawn(_worker, tmp_path, 29700 + (os.getpid() % 1000))
    assert all((tmp_path / ("ok%d" % r)).exists() for r in range(2))


def _p2p_worker_body(rank, world, port, tmp):
    import torch
    import torch.distributed as dist
    for p in (ROOT, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    if True:
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
        for fl in ("F1", "F2"):                # banded: stencil and exchange under the channel sum, tile-local ranges
            big = synth.make_fov(4 * 96 * world, Wm, 95, fov_index=31)[0]
            b0, b1 = sharding.slab_bounds(big.shape[0], rank, world)
            p2p_big = sharding.P2PMosaicSlab(b1 - b0, Wm)
            for rep in range(3):
                got_b = p2p_big.score(big[b0:b1].cuda(), fl, bands=4)
            want_b = O.neighbor2d_score(big.numpy(), fl)[b0:b1]
            # atol 1e-6: this FOV has one pixel (291, 86), score 1.2e-4, where the fixed-point stencil is 7.1e-7 off:
            # a local minimum along six of nine lines (lq = 0, uq = 8e-7), where F1's 1e-8 epsilon gives the score a
            # sensitivity of 146 to uq (DESIGN.md section 3, K3q error bound)
            np.testing.assert_allclose(got_b.cpu().numpy(), want_b, rtol=1e-5, atol=1e-6)
            p2p_big.check_peers()
            p2p_big.close()
        got3 = p2p.score(mine, "F3")            # F3 consumes the exchanged global range
        assert torch.equal(got3, nccl.score(mine, "F3"))
        p2p.check_peers()
        p2p.close()
        open(os.path.join(tmp, "p2p%d" % rank), "w").write("ok")
    dist.destroy_process_group()


def _p2p_worker(rank, world, port, tmp):
    _run_guarded(_p2p_worker_body, rank, world, port, tmp)


def test_mosaic_two_gpus_peer_memory(torch_cuda, tmp_path):
    """The library's own halo / range exchange over NVLink peer memory == the NCCL path, bit for bit."""
    if torch_cuda.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _spawn(_p2p_worker, tmp_path, 29800 + (os.getpid() % 1000))
    assert all((tmp_path / ("p2p%d" % r)).exists() for r in range(2))


# ---- the same peer-memory exchange with BOTH ranks on ONE GPU -------------------------------------------
# cudaIpc mappings work between processes that share a device, and the exchange kernels never call a collective
# library, so a single-GPU box exercises mosaic_push_kernel / mosaic_wait_kernel / mosaic_guard_kernel too.  The
# one-time handle exchange goes through a gloo group (NCCL refuses two ranks on one device).

def _cpu_exchange_hooks():
    """MosaicSlab hooks = the CUDA operators, with the halo / range exchange carried by CPU tensors (gloo)."""
    import torch
    from hipr_b200 import ops

    def channel_sum(cube_slab):
        s, mk = ops.channel_sum(cube_slab, None, normalize=False, dtype=torch.float64, return_max=True)
        vmax, vmin = mk.values()
        return s.cpu(), vmax.cpu(), vmin.cpu()

    def score(ext, gmax, gmin, flavour):
        keys = ops.MaxKey.from_values(gmax.cuda(), gmin.cuda()) if flavour not in ("F1", "F2") else None
        return ops.lne2d_fixed(ext.cuda(), flavour, 11, 9, padded=False, range_keys=keys)

    return {"channel_sum": channel_sum, "score": score, "accumulate": ops.cell_spectra_accumulate,
            "finalize": ops.cell_spectra_finalize}


def _one_gpu_worker_body(rank, world, port, tmp):
    import torch
    import torch.distributed as dist
    for p in (ROOT, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(0)                   # every rank on the same device
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hipr_b200 import sharding, synth
    from oracle import hipr_oracle as O
    Hm, Wm = 161, 256                          # uneven slabs
    r0, r1 = sharding.slab_bounds(Hm, rank, world)
    p2p = sharding.P2PMosaicSlab(r1 - r0, Wm)
    ref = sharding.MosaicSlab(hooks=_cpu_exchange_hooks())
    for step in range(4):                      # both parities, reused buffers
        cube = synth.make_fov(Hm, Wm, 95, fov_index=40 + step)[0]
        mine = cube[r0:r1].cuda()
        got = p2p.score(mine, "F1", check=True)
        assert torch.equal(got, ref.score(mine, "F1")), "peer-memory exchange differs from the reference exchange, step %d" % step
        if step == 0:
            want = O.neighbor2d_score(cube.numpy(), "F1")[r0:r1]
            np.testing.assert_allclose(got.cpu().numpy(), want, rtol=1e-5, atol=5e-7)
    got3 = p2p.score(mine, "F3", check=True)   # F3 consumes the exchanged global range
    assert torch.equal(got3, ref.score(mine, "F3"))
    want3 = O.neighbor2d_score(cube.numpy(), "F3")[r0:r1]
    np.testing.assert_allclose(got3.cpu().numpy(), want3, rtol=1e-5, atol=5e-7)
    p2p.close()
    # banded: exchange and stencil under the channel sum (hipr_mosaic_p2p_score)
    big = synth.make_fov(4 * 96 * world, Wm, 95, fov_index=31)[0]
    b0, b1 = sharding.slab_bounds(big.shape[0], rank, world)
    p2p_big = sharding.P2PMosaicSlab(b1 - b0, Wm)
    for fl in ("F1", "F2"):
        for rep in range(3):
            got_b = p2p_big.score(big[b0:b1].cuda(), fl, bands=4, check=True)
        want_b = O.neighbor2d_score(big.numpy(), fl)[b0:b1]
        np.testing.assert_allclose(got_b.cpu().numpy(), want_b, rtol=1e-5, atol=1e-6)   # atol: see _p2p_worker_body
    p2p_big.close()
    # a peer that never delivers: the wait kernel gives up, the score is poisoned, check_peers raises
    late = sharding.P2PMosaicSlab(r1 - r0, Wm, timeout_ms=300.0)
    if rank == 0:
        out = late.score(mine, "F1")
        assert bool(torch.isnan(out).all()), "a timed-out exchange must not return a score"
        with pytest.raises(RuntimeError):
            late.check_peers()
    late.close()
    open(os.path.join(tmp, "one%d" % rank), "w").write("ok")
    dist.destroy_process_group()


def _one_gpu_worker(rank, world, port, tmp):
    _run_guarded(_one_gpu_worker_body, rank, world, port, tmp)


def test_mosaic_peer_memory_two_ranks_one_gpu(torch_cuda, tmp_path):
    """Runs on a single-GPU box: the peer-memory exchange between two processes that share cuda:0."""
    _spawn(_one_gpu_worker, tmp_path, 29900 + (os.getpid() % 1000))
    assert all((tmp_path / ("one%d" % r)).exists() for r in range(2))

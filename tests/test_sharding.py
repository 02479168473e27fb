"""Multi-GPU host logic on CPU: world_size-2 (and 3) gloo process groups.  The compute hooks are
the oracle (this is a test); the product's hooks are the CUDA operators."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG, ROOT


def test_fov_shard_and_slab_bounds():
    from hipr_b200 import sharding
    for world in (1, 2, 4, 8):
        seen = sorted(i for r in range(world) for i in sharding.fov_shard(256, r, world))
        assert seen == list(range(256))
        assert all(len(sharding.fov_shard(256, r, world)) == 256 // world for r in range(world))
        bounds = [sharding.slab_bounds(16384, r, world) for r in range(world)]
        assert bounds[0][0] == 0 and bounds[-1][1] == 16384
        assert all(a[1] == b[0] for a, b in zip(bounds[:-1], bounds[1:]))
    assert [sharding.slab_bounds(16, r, 3) for r in range(3)] == [(0, 6), (6, 11), (11, 16)]
    with pytest.raises(ValueError):
        sharding.slab_bounds(7, 0, 2)


def _oracle_hooks():
    from oracle import hipr_oracle as O

    def channel_sum(cube_slab):
        s = torch.from_numpy(np.sum(cube_slab.numpy().astype(np.float64), axis=2))
        return s, s.max().reshape(1), s.min().reshape(1)

    def score(ext, gmax, gmin, flavour):
        return torch.from_numpy(O.lne2d(ext.numpy() / float(gmax), flavour))

    def accumulate(cube_slab, labels_slab, max_label):
        lab = labels_slab.numpy().reshape(-1)
        vals = cube_slab.numpy().reshape(lab.size, -1).astype(np.float64)
        sums = np.zeros((max_label + 1, vals.shape[1]))
        np.add.at(sums, lab, vals)
        counts = np.bincount(lab, minlength=max_label + 1).astype(np.int32)
        sums[0] = 0
        counts[0] = 0
        return torch.from_numpy(sums), torch.from_numpy(counts)

    def finalize(sums, counts):
        c = counts.numpy()
        labels = np.nonzero(c)[0]
        avg = sums.numpy()[labels] / c[labels][:, None]
        return labels, c[labels].astype(np.int64), avg, avg / avg.max(axis=1)[:, None]

    return {"channel_sum": channel_sum, "score": score, "accumulate": accumulate, "finalize": finalize}


def _worker(rank, world, port, tmp):
    for p in (ROOT, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from hipr_b200 import sharding, synth
        from oracle import hipr_oracle as O
        Hm, Wm = 60, 48
        cube, labels, L = synth.make_fov(Hm, Wm, 95, fov_index=9, label_dtype=torch.int64)
        r0, r1 = sharding.slab_bounds(Hm, rank, world)
        slab = sharding.MosaicSlab(hooks=_oracle_hooks())
        # halo exchange alone: the extended slab equals the matching rows of the whole image
        s_full = torch.from_numpy(np.sum(cube.numpy().astype(np.float64), axis=2))
        ext, nt, nb = sharding.exchange_halo(s_full[r0:r1].clone())
        assert (nt, nb) == (0 if rank == 0 else 5, 0 if rank == world - 1 else 5)
        assert torch.equal(ext, s_full[r0 - nt: r1 + nb])
        # score of the slab == rows of the unsplit score, bit for bit (the stencil is a gather)
        got = slab.score(cube[r0:r1], "F1").numpy()
        want = O.neighbor2d_score(cube.numpy(), "F1")[r0:r1]
        assert np.array_equal(got, want, equal_nan=True), "rank %d score differs" % rank
        # per-cell spectra: all-reduced partial sums/counts == the unsplit reduction; counts exact
        lab_out, area, avg, norm = slab.cell_spectra(cube[r0:r1], labels[r0:r1], L)
        wl, wa, wavg, wnorm = O.cell_spectra(labels.numpy(), cube.numpy())
        rooted = slab.cell_spectra(cube[r0:r1], labels[r0:r1], L, root=world - 1)     # reduce to one rank only
        if rank == world - 1:
            assert np.array_equal(rooted[0], wl) and np.array_equal(rooted[1], wa)
            np.testing.assert_allclose(rooted[2], wavg, rtol=1e-12)
        else:
            assert rooted is None
        assert np.array_equal(lab_out, wl) and np.array_equal(area, wa)
        np.testing.assert_allclose(avg, wavg, rtol=1e-12)
        # FOV sharding needs no communication: every FOV is scored by exactly one rank
        mine = sharding.fov_shard(5, rank, world)
        flags = torch.zeros(5, dtype=torch.int32)
        flags[mine] = 1
        dist.all_reduce(flags)
        assert flags.tolist() == [1] * 5
        open(os.path.join(tmp, "ok%d" % rank), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_mosaic_slabs_gloo(tmp_path, world):
    port = 29500 + (os.getpid() % 2000) + world
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / ("ok%d" % r)).exists() for r in range(world))

"""Multi-GPU host logic: one process per GPU (torch.distributed).

* Independent fields of view (BASELINE config 3) shard by index with NO data-path collective:
  rank r takes FOVs r, r + world, ...  (`fov_shard`).
* One stitched mosaic split across GPUs (config 5) is cut into row slabs.  The stencil runs on
  the 1-channel sum image, so the only exchange is a 5-row halo of that image with the slab
  above and below (`exchange_halo`, one batch of isend/irecv: NCCL send/recv over NVLink on
  GPUs), plus two tiny all-reduces: the global max/min of the sum image (`allreduce_range`) and,
  for per-cell spectra, the (L+1, C) partial sums and (L+1) integer pixel counts
  (`allreduce_cells`; integer counts stay exact).

The compute steps are the CUDA operators in hipr_b200.ops; `MosaicSlab` takes them as hooks so
the communication logic can be exercised on CPU tensors with the gloo backend in tests.
"""
import torch
import torch.distributed as dist

HALO = 5   # (patch_size - 1) / 2 for the 11-sample lines every pipeline uses


def fov_shard(n_fov, rank, world):
    """Indices of the fields of view rank `rank` processes."""
    return list(range(rank, n_fov, world))


def slab_bounds(n_rows, rank, world):
    """Rows [r0, r1) of the mosaic owned by `rank`: contiguous, balanced, every rank >= HALO rows."""
    if n_rows < world * HALO:
        raise ValueError("mosaic of %d rows is too short for %d slabs" % (n_rows, world))
    base, extra = divmod(n_rows, world)
    r0 = rank * base + min(rank, extra)
    return r0, r0 + base + (1 if rank < extra else 0)


def exchange_halo(slab, group=None, halo=HALO):
    """slab (rows, W) -> (halo_top + rows + halo_bottom, W): rows from the neighbouring ranks above
    and below; the outer ranks get nothing on their outer side (the stencil clamps to the edge
    there, = np.pad(mode='edge')).  Returns (extended, n_top, n_bottom)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    slab = slab.contiguous()
    up, down = rank - 1, rank + 1
    top = torch.empty((halo, slab.shape[1]), dtype=slab.dtype, device=slab.device) if up >= 0 else None
    bottom = torch.empty((halo, slab.shape[1]), dtype=slab.dtype, device=slab.device) if down < world else None
    ops = []
    peer = (lambda r: r) if group is None else (lambda r: dist.get_global_rank(group, r))
    send_up = slab[:halo].contiguous() if up >= 0 else None
    send_down = slab[-halo:].contiguous() if down < world else None
    if up >= 0:
        ops += [dist.P2POp(dist.isend, send_up, peer(up), group), dist.P2POp(dist.irecv, top, peer(up), group)]
    if down < world:
        ops += [dist.P2POp(dist.isend, send_down, peer(down), group),
                dist.P2POp(dist.irecv, bottom, peer(down), group)]
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    parts = [p for p in (top, slab, bottom) if p is not None]
    return torch.cat(parts, dim=0), (halo if top is not None else 0), (halo if bottom is not None else 0)


def allreduce_range(vmax, vmin, group=None):
    """Global (max, min) from per-slab float64 scalars (1-element tensors)."""
    t = torch.stack([vmax.reshape(()), -vmin.reshape(())]).to(torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return t[0:1].clone(), (-t[1:2]).clone()


def allreduce_cells(sums, counts, group=None):
    """Sum the per-slab (L+1, C) float64 channel sums and (L+1) int32 pixel counts in place."""
    dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    return sums, counts


class MosaicSlab:
    """One rank's row slab of a stitched mosaic: score map and per-cell spectra with the same
    results as the unsplit mosaic.

    hooks (defaults = the CUDA operators):
      channel_sum(cube_slab) -> (sum_slab float64 (rows, W), vmax (1,), vmin (1,))
      score(extended_sum, vmax, vmin, flavour) -> (rows_ext, W) score of the extended image
      accumulate(cube_slab, labels_slab, max_label) -> (sums, counts)
      finalize(sums, counts) -> (labels, area, avgint, avgint_norm)
    """

    def __init__(self, group=None, hooks=None):
        self.group = group
        self.hooks = hooks or _cuda_hooks()

    def score(self, cube_slab, flavour="F1"):
        s, vmax, vmin = self.hooks["channel_sum"](cube_slab)
        ext, n_top, n_bottom = exchange_halo(s, self.group)
        gmax, gmin = allreduce_range(vmax, vmin, self.group)
        full = self.hooks["score"](ext, gmax, gmin, flavour)
        return full[n_top: full.shape[0] - n_bottom]

    def cell_spectra(self, cube_slab, labels_slab, max_label):
        sums, counts = self.hooks["accumulate"](cube_slab, labels_slab, max_label)
        allreduce_cells(sums, counts, self.group)
        return self.hooks["finalize"](sums, counts)


def _cuda_hooks():
    from . import ops

    def channel_sum(cube_slab):
        s, mk = ops.channel_sum(cube_slab, None, normalize=False, dtype=torch.float64, return_max=True)
        vmax, vmin = mk.values()
        return s, vmax, vmin

    def score(ext, gmax, gmin, flavour):
        return ops.lne2d_fixed(ext, flavour, 11, 9, padded=False, range_keys=ops.MaxKey.from_values(gmax, gmin))

    return {"channel_sum": channel_sum, "score": score, "accumulate": ops.cell_spectra_accumulate,
            "finalize": ops.cell_spectra_finalize}

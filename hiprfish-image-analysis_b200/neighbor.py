"""Drop-in replacement for the reference's `neighbor` Cython extension (bio/neighbor.pyx).

    from neighbor import line_profile_v2, line_profile_memory_efficient_v2, \\
        line_profile_memory_efficient_v3                 # bio/...analysis.py:29,39,40

Same names and positional signatures; numpy float64 in -> new numpy float64 out (copies
included), or torch CUDA tensor in -> CUDA tensor out.  No CPU fallback.

`line_profile` and `neighbor_average` are dead code in the reference (never imported by any
script) and cannot complete there: they are exported so that the module surface is the same, and
fail the way the originals do.
"""
import numpy as np

from hipr_b200 import tables as _tables
from neighbor2d import _as_double_2d, _cname


def _run(fn, image_padded, *params):
    import torch

    params = [_tables._int_arg(p, "parameter") for p in params]
    if isinstance(image_padded, torch.Tensor):
        return fn(image_padded, *params)
    a = _as_double_2d(image_padded, 3)
    dev = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return fn(dev, *params).cpu().numpy()


def line_profile_v2(image_padded, patch_size, theta_range, phi_range):
    """(Xp,Yp,Zp) -> (X,Y,Z,(theta_range-1)*phi_range,P) literal gather, bio/neighbor.pyx:115-181."""
    import torch
    from hipr_b200 import ops
    if isinstance(image_padded, torch.Tensor):
        return _run(ops.line_profile_3d, image_padded, patch_size, theta_range, phi_range)
    params = [_tables._int_arg(p, "parameter") for p in (patch_size, theta_range, phi_range)]
    return ops.line_profile_3d_host(_as_double_2d(image_padded, 3), *params)


def line_profile_memory_efficient_v2(image_padded, patch_size, theta_range, phi_range):
    """(Xp,Yp,Zp) -> (X,Y,Z,T): (centre-min)/max(max-min,1e-8) per direction, bio/neighbor.pyx:186-263."""
    import torch
    from hipr_b200 import ops
    if isinstance(image_padded, torch.Tensor):
        return _run(lambda v, p, t, f: ops.lne3d_dirs(v, p, t, f, padded=True), image_padded, patch_size, theta_range,
                    phi_range)
    # host arrays: bands of x-planes under the device -> host copy (hipr_lne3d_dirs_host); ordinary numpy result
    params = [_tables._int_arg(p, "parameter") for p in (patch_size, theta_range, phi_range)]
    return ops.lne3d_dirs_host(_as_double_2d(image_padded, 3), *params)


def line_profile_memory_efficient_v3(image_padded, patch_size, theta_range, phi_range):
    """(Xp,Yp,Zp) -> (X,Y,Z), bio/neighbor.pyx:268-349.  The v3 table reaches outside the 11^3 patch
    (entries up to 18); the reference then reads image_padded's buffer at the flat address with
    bounds checks off.  In-buffer addresses are reproduced exactly; voxels whose reads would
    pass the end of the buffer (undefined behaviour in the reference) are NaN."""
    from hipr_b200 import ops
    return _run(lambda v, p, t, f: ops.lne3d(v, "V3", p, t, f, padded=True), image_padded, patch_size, theta_range,
                phi_range)


def line_profile(image_padded, patch_size, theta_range, phi_range):
    """bio/neighbor.pyx:42-110: dead code.  It ignores the voxel position when sampling (:102),
    prints every sample (:101) and divides by zero on flat input; no script imports it.  The
    argument checks of its `np.ndarray[float, ndim=3]` signature are reproduced, then it refuses."""
    a = np.asarray(image_padded)
    if a.dtype != np.float32:
        raise ValueError("Buffer dtype mismatch, expected 'float' but got %r" % _cname(a.dtype))
    if a.ndim != 3:
        raise ValueError("Buffer has wrong number of dimensions (expected 3, got %d)" % a.ndim)
    raise NotImplementedError("neighbor.line_profile is dead code in the reference (bio/neighbor.pyx:42-110) "
                              "and is not accelerated; use line_profile_memory_efficient_v2")


def neighbor_average(image_padded, patch_size):
    """bio/neighbor.pyx:8-37: cannot run in the reference either -- it binds a float64 ndarray to a
    `float[:,:,:,:]` view (:25-26), which raises this ValueError after the input check."""
    a = np.asarray(image_padded)
    if a.dtype != np.float32:
        raise ValueError("Buffer dtype mismatch, expected 'float' but got %r" % _cname(a.dtype))
    if a.ndim != 3:
        raise ValueError("Buffer has wrong number of dimensions (expected 3, got %d)" % a.ndim)
    raise ValueError("Buffer dtype mismatch, expected 'float' but got 'double'")

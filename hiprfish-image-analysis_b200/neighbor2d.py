"""Drop-in replacement for the reference's `neighbor2d` Cython extension.

    from neighbor2d import line_profile_2d_v2            # syn/...measurement.py:30, bio/...analysis.py:34

Same name, same positional signature, same array conventions as eco/neighbor2d.pyx:8-64:
numpy float64 (Hp, Wp) in -> new numpy float64 (Hp-P+1, Wp-P+1, phi_range, P) out, computed by
the sm_100a gather kernel in row bands under the device -> host copy (hipr_line_profile_2d_host).  A torch CUDA tensor (float32 or
float64) is also accepted and then a CUDA tensor of the same dtype is returned without any
host copy.  There is no CPU fallback.

The fused entry points the measurement scripts can call instead of the numpy blocks around the
stencil live in hipr_b200 (lne2d, neighbor2d_score, channel_sum, cell_spectra).
"""
import numpy as np

from hipr_b200 import tables as _tables


def _as_double_2d(a, ndim):
    a = np.asarray(a) if not hasattr(a, "__cuda_array_interface__") else a
    if isinstance(a, np.ndarray):
        # same checks, same exception classes and messages as a `double[:, :]` typed memoryview
        if a.dtype != np.float64:
            raise ValueError("Buffer dtype mismatch, expected 'double' but got %r" % _cname(a.dtype))
        if a.ndim != ndim:
            raise ValueError("Buffer has wrong number of dimensions (expected %d, got %d)" % (ndim, a.ndim))
    return a


def _cname(dt):
    return {"float64": "double", "float32": "float", "int64": "long", "int32": "int", "float16": "half"}.get(str(dt), str(dt))


def line_profile_2d_v2(image_padded, patch_size, phi_range):
    """lp[i, j, t, li] = image_padded[i + tab[t, li, 0], j + tab[t, li, 1]]  (eco/neighbor2d.pyx:56-63)."""
    import torch
    from hipr_b200 import ops

    patch_size = _tables._int_arg(patch_size, "patch_size")
    phi_range = _tables._int_arg(phi_range, "phi_range")
    if isinstance(image_padded, torch.Tensor):
        return ops.line_profile_2d(image_padded, patch_size, phi_range)
    a = _as_double_2d(image_padded, 2)
    # host arrays: banded gather under the device -> host copy, through the page-locked staging ring
    # (hipr_line_profile_2d_host); the result is an ordinary numpy array, as the reference's np.zeros is
    return ops.line_profile_2d_host(a, patch_size, phi_range)

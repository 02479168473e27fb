"""Device-resident operators: torch CUDA tensors in, torch CUDA tensors out, stream-ordered on
torch's current stream.  torch is used for device memory and streams only; every kernel that
runs here is one of libhipr_b200's.  No CPU fallback: CPU tensors are rejected.

Reference blocks replaced (paths relative to the reference repository):
  register_stacks   syn/hiprfish_imaging_multispecies_spectral_image_measurement.py:86-105
  channel_sum       syn/hiprfish_imaging_multispecies_spectral_image_measurement.py:105-106
  lne2d             eco/neighbor2d.pyx:56-63 + syn/...measurement.py:109-124 (F1), bio F2 / F3
  neighbor2d_score  syn/...measurement.py:105-124 without the skimage denoise
  line_profile_2d   eco/neighbor2d.pyx:8-64
  line_profile_3d   bio/neighbor.pyx:115-181
  lne3d_dirs        bio/neighbor.pyx:186-263
  lne3d             bio/neighbor.pyx + bio/hiprfish_imaging_biofilm_analysis.py:812-817, 905-917, 1114-1125
  cell_spectra      syn/...measurement.py:167-172
"""
import ctypes as C

import numpy as np
import torch

from . import tables
from ._lib import F32, F64, FLAVOURS, check, lib

_DT = {torch.float32: F32, torch.float64: F64}


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dev(t, name, dtypes=(torch.float32, torch.float64)):
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor on a CUDA device" % name)
    if not t.is_cuda:
        raise ValueError("%s must live on a CUDA device (there is no CPU path)" % name)
    if t.dtype not in dtypes:
        raise TypeError("%s: unsupported dtype %s" % (name, t.dtype))
    return t.contiguous()


def _tab_ptr(tab):
    return tab.ctypes.data_as(C.c_void_p)


def _flavour(f):
    try:
        return FLAVOURS[f]
    except KeyError:
        raise ValueError("unknown flavour %r (one of %s)" % (f, sorted(FLAVOURS)))


class MaxKey:
    """Two device scalars: max and min of an image as order-preserving keys (see hipr_b200.h)."""

    def __init__(self, device):
        # filled (and reset) by hipr_chansum / hipr_image_range; no torch kernel on the hot path
        self.key = torch.empty(2, dtype=torch.int64, device=device)

    def ptr(self):
        return C.c_void_p(self.key.data_ptr())

    def value(self):
        out = torch.empty(1, dtype=torch.float64, device=self.key.device)
        with torch.cuda.device(self.key.device):
            check(lib().hipr_maxkey_decode(self.ptr(), C.c_void_p(out.data_ptr()), _stream()), "maxkey_decode")
        return out

    def values(self):
        """(max, min) as 1-element float64 device tensors."""
        out = torch.empty(2, dtype=torch.float64, device=self.key.device)
        with torch.cuda.device(self.key.device):
            check(lib().hipr_range_decode(self.ptr(), C.c_void_p(out.data_ptr()), _stream()), "range_decode")
        return out[0:1], out[1:2]

    @classmethod
    def from_values(cls, vmax, vmin):
        mm = torch.cat([vmax.reshape(1), vmin.reshape(1)]).to(torch.float64).contiguous()
        mk = cls(mm.device)
        with torch.cuda.device(mm.device):
            check(lib().hipr_range_encode(C.c_void_p(mm.data_ptr()), mk.ptr(), _stream()), "range_encode")
        return mk


def channel_sum(cube, calibration=None, normalize=True, dtype=torch.float32, return_max=False):
    """cube (..., C) float32 -> (...) sum over the last axis, divided by its global max when
    `normalize` (image_registered_sum / np.max(...)).  Accumulation is float64."""
    cube = _dev(cube, "cube", (torch.float32,))
    if cube.dim() < 2:
        raise ValueError("cube must have at least 2 dimensions (..., C)")
    Cn = cube.shape[-1]
    npix = cube.numel() // Cn
    if calibration is not None:
        calibration = _dev(calibration, "calibration", (torch.float32,))
        if calibration.shape != cube.shape:
            calibration = calibration.expand_as(cube).contiguous()
    out = torch.empty(cube.shape[:-1], dtype=dtype, device=cube.device)
    mk = MaxKey(cube.device) if (normalize or return_max) else None
    with torch.cuda.device(cube.device):
        check(lib().hipr_chansum(C.c_void_p(cube.data_ptr()),
                                 C.c_void_p(calibration.data_ptr()) if calibration is not None else None,
                                 npix, Cn, C.c_void_p(out.data_ptr()), _DT[dtype],
                                 mk.ptr() if mk else None, _stream()), "channel_sum")
        if normalize:
            check(lib().hipr_normalize(C.c_void_p(out.data_ptr()), _DT[dtype], npix, mk.ptr(), _stream()),
                  "normalize")
    return (out, mk) if return_max else out


def channel_sum_raw(cube, scale, return_max=False):
    """cube (..., C) of raw detector counts (torch.uint8 / uint16 / int16 bit pattern of uint16) -> float64
    channel sums of float32(count) / float32(scale), the values bioformats' rescale hands the scripts."""
    if not isinstance(cube, torch.Tensor) or not cube.is_cuda:
        raise ValueError("cube must be a CUDA tensor (there is no CPU path)")
    if cube.dtype not in (torch.uint8, torch.uint16, torch.int16):
        raise TypeError("raw cube must be uint8 or uint16, got %s" % cube.dtype)
    cube = cube.contiguous()
    Cn = cube.shape[-1]
    npix = cube.numel() // Cn
    out = torch.empty(cube.shape[:-1], dtype=torch.float64, device=cube.device)
    mk = MaxKey(cube.device)
    with torch.cuda.device(cube.device):
        check(lib().hipr_chansum_raw(C.c_void_p(cube.data_ptr()), cube.element_size(), float(scale), npix, Cn,
                                     C.c_void_p(out.data_ptr()), mk.ptr(), _stream()), "channel_sum_raw")
    return (out, mk) if return_max else out


def register_stacks(stacks, shifts=None, calibration=None, return_cube=True):
    """Registration paste + np.dstack + flat-field divide + channel sum in one pass (csrc/register.cu).

    stacks: list of (H, W, c_e) float32 CUDA tensors, one per excitation; shifts: per stack (row, col)
    integer shifts as the scripts derive them from register_translation (int(shift_vector)), None = no
    shift; calibration: None or a (H, W, C) float32 divisor.  Returns (cube (H, W, C) float32 or None,
    channel sums (H, W) float64, MaxKey)."""
    stacks = [_dev(s, "stack %d" % i, (torch.float32,)) for i, s in enumerate(stacks)]
    if not stacks:
        raise ValueError("need at least one excitation stack")
    H, W = stacks[0].shape[:2]
    for s in stacks:
        if s.dim() != 3 or tuple(s.shape[:2]) != (H, W):
            raise ValueError("every stack must be (H, W, c_e) with the same H, W")
    n = len(stacks)
    if shifts is None:
        shifts = [(0, 0)] * n
    if len(shifts) != n:
        raise ValueError("one (row, col) shift per stack")
    chans = np.asarray([s.shape[2] for s in stacks], dtype=np.int32)
    srow = np.asarray([int(v[0]) for v in shifts], dtype=np.int32)     # int(): truncation, as the scripts do
    scol = np.asarray([int(v[1]) for v in shifts], dtype=np.int32)
    Cn = int(chans.sum())
    dev = stacks[0].device
    if calibration is not None:
        calibration = _dev(calibration, "calibration", (torch.float32,))
        if tuple(calibration.shape) != (H, W, Cn):
            calibration = calibration.expand(H, W, Cn).contiguous()
    cube = torch.empty((H, W, Cn), dtype=torch.float32, device=dev) if return_cube else None
    s64 = torch.empty((H, W), dtype=torch.float64, device=dev)
    mk = MaxKey(dev)
    ptrs = (C.c_void_p * n)(*[s.data_ptr() for s in stacks])
    with torch.cuda.device(dev):
        check(lib().hipr_register_stacks(ptrs, chans.ctypes.data_as(C.c_void_p), srow.ctypes.data_as(C.c_void_p),
                                         scol.ctypes.data_as(C.c_void_p), n, H, W,
                                         C.c_void_p(calibration.data_ptr()) if calibration is not None else None,
                                         C.c_void_p(cube.data_ptr()) if cube is not None else None,
                                         C.c_void_p(s64.data_ptr()), mk.ptr(), _stream()), "register_stacks")
    return cube, s64, mk


def neighbor2d_score_from_stacks(stacks, shifts=None, calibration=None, flavour="F1", return_cube=True):
    """syn/..._measurement.py:86-124 without the skimage calls: excitation stacks -> registered cube,
    channel sums, score map (fixed-point stencil).  Returns (score, cube, sums, MaxKey)."""
    cube, s64, mk = register_stacks(stacks, shifts, calibration, return_cube)
    score = lne2d_fixed(s64, flavour, 11, 9, padded=False, range_keys=mk)
    return score, cube, s64, mk


def denoise_nl_means(image, patch_size=7, patch_distance=11, h=0.1):
    """skimage.restoration.denoise_nl_means(image, h=h) for a 2-D image (fast mode, sigma 0), as
    syn/...measurement.py:108 calls it on the normalised sum image, or for a 3-D volume as bio/...analysis.py:454
    calls it on a z-stack.  (H, W) / (X, Y, Z) float32 / float64 CUDA tensor -> same dtype.  csrc/nlm2d.cu,
    csrc/nlm3d.cu; patch_size 7, patch_distance <= 15."""
    image = _dev(image, "image")
    if image.dim() == 3:
        # a z-stack volume (bio/...analysis.py:454, h = 0.03): csrc/nlm3d.cu
        X, Y, Z = image.shape
        need = int(lib().hipr_denoise_nl_means_3d_workspace(X, Y, Z, int(patch_distance)))
        if need < 0:
            raise ValueError("patch_distance must be in [0, 15]")
        work = torch.empty(need, dtype=torch.uint8, device=image.device)
        out = torch.empty_like(image)
        with torch.cuda.device(image.device):
            check(lib().hipr_denoise_nl_means_3d(C.c_void_p(image.data_ptr()), X, Y, Z, _DT[image.dtype], int(patch_size),
                                                 int(patch_distance), float(h), C.c_void_p(out.data_ptr()),
                                                 C.c_void_p(work.data_ptr()), need, _stream()), "denoise_nl_means_3d")
        return out
    if image.dim() != 2:
        raise ValueError("image must be 2-D or 3-D, got %d-D" % image.dim())
    Hh, Ww = image.shape
    out = torch.empty_like(image)
    with torch.cuda.device(image.device):
        check(lib().hipr_denoise_nl_means_2d(C.c_void_p(image.data_ptr()), Hh, Ww, _DT[image.dtype], int(patch_size),
                                             int(patch_distance), float(h), C.c_void_p(out.data_ptr()), _stream()),
              "denoise_nl_means")
    return out


def lne2d(image, flavour="F1", patch_size=11, phi_range=9, padded=False, maxkey=None):
    """Score map ("local neighbourhood enhancement") of a 2-D image.

    image: (H, W) when padded=False (border samples clamp to the edge, = np.pad(mode='edge')),
    or the already padded (H+P-1, W+P-1) image when padded=True.  maxkey: optional MaxKey from
    channel_sum(normalize=False, return_max=True); samples are divided by the max on load."""
    image = _dev(image, "image")
    if image.dim() != 2:
        raise ValueError("image must be 2-D, got %d-D" % image.dim())
    tab = tables.line_table_2d(patch_size, phi_range)
    P = tab.shape[1]
    Hs, Ws = image.shape
    H, W = (Hs - P + 1, Ws - P + 1) if padded else (Hs, Ws)
    if H < 1 or W < 1:
        raise ValueError("image smaller than the patch")
    out = torch.empty((H, W), dtype=image.dtype, device=image.device)
    with torch.cuda.device(image.device):
        check(lib().hipr_lne2d(C.c_void_p(image.data_ptr()), Hs, Ws, Ws, int(bool(padded)), _DT[image.dtype], P,
                               tab.shape[0], _tab_ptr(tab), _flavour(flavour),
                               maxkey.ptr() if maxkey is not None else None, C.c_void_p(out.data_ptr()),
                               _stream()), "lne2d")
    return out


def image_range(image):
    """MaxKey (max and min keys) of any float image, for lne2d_fixed."""
    image = _dev(image, "image")
    mk = MaxKey(image.device)
    with torch.cuda.device(image.device):
        check(lib().hipr_image_range(C.c_void_p(image.data_ptr()), _DT[image.dtype], image.numel(), mk.ptr(),
                                     _stream()), "image_range")
    return mk


def lne2d_fixed(image, flavour="F1", patch_size=11, phi_range=9, padded=False, range_keys=None):
    """lne2d on a 31-bit fixed-point copy of the image (csrc/lne2d_q.cu): min, max and differences
    along each line are exact, so a float64 image keeps its precision at float32 speed.  The
    implied normalisation is image / max(image).  (11, 9) only; float32 score out."""
    image = _dev(image, "image")
    if image.dim() != 2:
        raise ValueError("image must be 2-D, got %d-D" % image.dim())
    tab = tables.line_table_2d(patch_size, phi_range)
    if tab.shape[:2] != (9, 11):
        raise ValueError("the fixed-point stencil exists for patch_size=11, phi_range=9 only; use lne2d")
    if range_keys is None and _flavour(flavour) not in (1, 2):
        range_keys = image_range(image)      # F3 scales its epsilon by the global range
    # F1 / F2 with no keys: every 32x32 tile is quantised with its own min / max (finest resolution)
    Hs, Ws = image.shape
    H, W = (Hs - 10, Ws - 10) if padded else (Hs, Ws)
    if H < 1 or W < 1:
        raise ValueError("image smaller than the patch")
    out = torch.empty((H, W), dtype=torch.float32, device=image.device)
    with torch.cuda.device(image.device):
        check(lib().hipr_lne2d_q(C.c_void_p(image.data_ptr()), Hs, Ws, Ws, int(bool(padded)), _DT[image.dtype], 11, 9,
                                 _tab_ptr(tab), _flavour(flavour), range_keys.ptr() if range_keys is not None else None,
                                 C.c_void_p(out.data_ptr()), _stream()), "lne2d_fixed")
    return out


UNSUPPORTED = -9
_SIDE = {}


def _side_streams(device, n=2):
    key = (device.index if device.index is not None else torch.cuda.current_device(), n)
    if key not in _SIDE:
        _SIDE[key] = [torch.cuda.Stream(device=device) for _ in range(n)]
    return _SIDE[key]


def neighbor2d_fused(cube, flavour="F1", patch_size=11, phi_range=9, return_sum=False):
    """cube (H, W, C) float32 -> float32 score in ONE launch (csrc/fused2d.cu).  Returns None when
    the request is outside the fused kernel's envelope (the caller then uses the two-kernel
    path).  With return_sum: (score, float64 channel sums, MaxKey)."""
    cube = _dev(cube, "cube", (torch.float32,))
    if cube.dim() != 3:
        raise ValueError("cube must be (H, W, C)")
    H, W, Cn = cube.shape
    tab = tables.line_table_2d(patch_size, phi_range)
    score = torch.empty((H, W), dtype=torch.float32, device=cube.device)
    s = torch.empty((H, W), dtype=torch.float64, device=cube.device) if return_sum else None
    mk = MaxKey(cube.device) if return_sum else None
    with torch.cuda.device(cube.device):
        code = lib().hipr_neighbor2d_fused(C.c_void_p(cube.data_ptr()), H, W, Cn, tab.shape[1], tab.shape[0],
                                           _tab_ptr(tab), _flavour(flavour), C.c_void_p(score.data_ptr()),
                                           C.c_void_p(s.data_ptr()) if return_sum else None,
                                           mk.ptr() if return_sum else None, _stream())
    if code == UNSUPPORTED:
        return None
    check(code, "neighbor2d_fused")
    return (score, s, mk) if return_sum else score


def neighbor2d_pipeline(cube, flavour="F1", patch_size=11, phi_range=9, bands=0):
    """cube (H, W, C) float32 -> (score float32, channel sums float64, MaxKey) through
    hipr_neighbor2d: channel sum and fixed-point stencil overlapped band by band (csrc/pipeline2d.cu).
    Returns None for parameters outside (11, 9)."""
    cube = _dev(cube, "cube", (torch.float32,))
    if cube.dim() != 3:
        raise ValueError("cube must be (H, W, C)")
    H, W, Cn = cube.shape
    tab = tables.line_table_2d(patch_size, phi_range)
    score = torch.empty((H, W), dtype=torch.float32, device=cube.device)
    s = torch.empty((H, W), dtype=torch.float64, device=cube.device)
    mk = MaxKey(cube.device)
    with torch.cuda.device(cube.device):
        code = lib().hipr_neighbor2d(C.c_void_p(cube.data_ptr()), H, W, Cn, tab.shape[1], tab.shape[0], _tab_ptr(tab),
                                     _flavour(flavour), C.c_void_p(score.data_ptr()), C.c_void_p(s.data_ptr()),
                                     mk.ptr(), int(bands), _stream())
    if code == UNSUPPORTED:
        return None
    check(code, "neighbor2d")
    return score, s, mk


def neighbor2d_score(cube, flavour="F1", calibration=None, patch_size=11, phi_range=9, dtype=None,
                     return_sum=False, denoise_h=None):
    """cube (H, W, C) or (N, H, W, C) float32 -> score map(s) (H, W) / (N, H, W).

    channel sum -> /max -> [NL-means denoise] -> edge pad -> line profiles -> epilogue, FOV by FOV (each
    FOV has its own max).  dtype=None (default): float64 channel sums feed the fixed-point stencil
    (lne2d_fixed; float32 score).  dtype=torch.float32 / float64: the sum image is stored in that
    type and the floating-point stencil of that type runs, dividing by the max on load.
    denoise_h: None, or the `h` of skimage.restoration.denoise_nl_means applied to the normalised sum
    image before the stencil (0.02 at syn/...measurement.py:108): the whole chain stays on the device."""
    cube = _dev(cube, "cube", (torch.float32,))
    if denoise_h is not None and cube.dim() == 3:
        s = channel_sum(cube, calibration, normalize=True, dtype=torch.float64)
        den = denoise_nl_means(s, h=denoise_h)
        # the denoised image is smooth: its 11-sample lines span ~1e-4 of a tile's range, below what the
        # 31-bit fixed-point stencil resolves to 1e-5 (measured 5e-6), so the float64 stencil runs here
        # (0.25 ms at 2048^2, next to ~9 ms of denoising)
        score = lne2d(den if dtype in (None, torch.float64) else den.to(dtype), flavour, patch_size, phi_range)
        return (score, den) if return_sum else score
    if cube.dim() == 4:
        # independent FOVs alternate between two streams: FOV i+1's channel sum (HBM-bound) runs
        # under FOV i's stencil (SM-bound); joined on the caller's stream before returning
        cur = torch.cuda.current_stream(cube.device)
        side = _side_streams(cube.device)
        for st in side:
            st.wait_stream(cur)
        res = []
        for i, c in enumerate(cube):
            with torch.cuda.stream(side[i % len(side)]):
                res.append(neighbor2d_score(c, flavour, None if calibration is None else calibration, patch_size,
                                            phi_range, dtype, return_sum, denoise_h))
        for st in side:
            cur.wait_stream(st)
        for r in res:
            for t in (r if return_sum else (r,)):
                t.record_stream(cur)
        if return_sum:
            return torch.stack([r[0] for r in res]), torch.stack([r[1] for r in res])
        return torch.stack(res)
    if cube.dim() != 3:
        raise ValueError("cube must be (H, W, C) or (N, H, W, C)")
    if dtype is None and calibration is None:
        res = neighbor2d_pipeline(cube, flavour, patch_size, phi_range)
        if res is not None:
            score, s, mk = res
            if not return_sum:
                return score
            sn = torch.empty(s.shape, dtype=torch.float32, device=s.device)
            with torch.cuda.device(cube.device):
                check(lib().hipr_normalize_cast(C.c_void_p(s.data_ptr()), s.numel(), mk.ptr(),
                                                C.c_void_p(sn.data_ptr()), _stream()), "normalize_cast")
            return score, sn
    fixed = dtype is None and tables.line_table_2d(patch_size, phi_range).shape[:2] == (9, 11)
    if dtype is None:
        dtype = torch.float64
    s, mk = channel_sum(cube, calibration, normalize=False, dtype=dtype, return_max=True)
    if fixed:
        score = lne2d_fixed(s, flavour, patch_size, phi_range, padded=False, range_keys=mk)
    else:
        score = lne2d(s, flavour, patch_size, phi_range, padded=False, maxkey=mk)
    if return_sum:
        with torch.cuda.device(cube.device):
            check(lib().hipr_normalize(C.c_void_p(s.data_ptr()), _DT[dtype], s.numel(), mk.ptr(), _stream()),
                  "normalize")
        return score, s
    return score


def line_profile_2d(image_padded, patch_size, phi_range):
    """Literal gather: (Hp, Wp) -> (Hp-P+1, Wp-P+1, phi_range, P), same dtype."""
    image_padded = _dev(image_padded, "image_padded")
    if image_padded.dim() != 2:
        raise ValueError("Buffer has wrong number of dimensions (expected 2, got %d)" % image_padded.dim())
    tab = tables.line_table_2d(patch_size, phi_range)
    R, P = tab.shape[0], tab.shape[1]
    Hp, Wp = image_padded.shape
    if Hp < P or Wp < P:
        raise ValueError("image smaller than the patch")
    out = torch.empty((Hp - P + 1, Wp - P + 1, R, P), dtype=image_padded.dtype, device=image_padded.device)
    with torch.cuda.device(image_padded.device):
        check(lib().hipr_line_profile_2d(C.c_void_p(image_padded.data_ptr()), Hp, Wp, _DT[image_padded.dtype], P, R,
                                         _tab_ptr(tab), C.c_void_p(out.data_ptr()), _stream()), "line_profile_2d")
    return out


def line_profile_2d_host(image_padded, patch_size, phi_range, out=None, pinned=False):
    """numpy float64 (Hp, Wp) -> numpy float64 (Hp-P+1, Wp-P+1, phi_range, P) through hipr_line_profile_2d_host: the
    gather runs in row bands on the device under the device -> host copy of the previous bands.  out: a caller's
    array (page-locked or not); pinned=True returns an array backed by page-locked memory (torch's caching host
    allocator), which the copy reaches at the PCIe rate without the staging ring."""
    a = np.ascontiguousarray(image_padded)
    if a.dtype != np.float64:
        raise TypeError("image_padded must be float64, got %s" % a.dtype)
    if a.ndim != 2:
        raise ValueError("Buffer has wrong number of dimensions (expected 2, got %d)" % a.ndim)
    tab = tables.line_table_2d(patch_size, phi_range)
    R, P = tab.shape[0], tab.shape[1]
    Hp, Wp = a.shape
    if Hp < P or Wp < P:
        raise ValueError("image smaller than the patch")
    shape = (Hp - P + 1, Wp - P + 1, R, P)
    if out is None:
        out = torch.empty(shape, dtype=torch.float64, pin_memory=True).numpy() if pinned else np.empty(shape, dtype=np.float64)
    elif out.shape != shape or out.dtype != np.float64 or not out.flags.c_contiguous:
        raise ValueError("out must be a C-contiguous float64 array of shape %s" % (shape,))
    check(lib().hipr_line_profile_2d_host(a.ctypes.data_as(C.c_void_p), Hp, Wp, P, R, _tab_ptr(tab),
                                          out.ctypes.data_as(C.c_void_p)), "line_profile_2d_host")
    return out


def _vol(volume, name):
    volume = _dev(volume, name)
    if volume.dim() != 3:
        raise ValueError("Buffer has wrong number of dimensions (expected 3, got %d)" % volume.dim())
    return volume


def line_profile_3d(volume_padded, patch_size, theta_range, phi_range):
    """Literal 5-D gather: (Xp, Yp, Zp) -> (X, Y, Z, (theta_range-1)*phi_range, P)."""
    v = _vol(volume_padded, "image_padded")
    tab = tables.line_table_3d(patch_size, theta_range, phi_range)
    T, P = tab.shape[0], tab.shape[1]
    Xp, Yp, Zp = v.shape
    if min(Xp, Yp, Zp) < P:
        raise ValueError("volume smaller than the patch")
    out = torch.empty((Xp - P + 1, Yp - P + 1, Zp - P + 1, T, P), dtype=v.dtype, device=v.device)
    with torch.cuda.device(v.device):
        check(lib().hipr_line_profile_3d(C.c_void_p(v.data_ptr()), Xp, Yp, Zp, _DT[v.dtype], P, T, _tab_ptr(tab),
                                         C.c_void_p(out.data_ptr()), _stream()), "line_profile_3d")
    return out


def lne3d_dirs(volume, patch_size=11, theta_range=9, phi_range=9, padded=True, maxkey=None):
    """line_profile_memory_efficient_v2: per-direction relative centre value, (X, Y, Z, T)."""
    v = _vol(volume, "image_padded")
    tab = tables.line_table_3d(patch_size, theta_range, phi_range)
    T, P = tab.shape[0], tab.shape[1]
    Xs, Ys, Zs = v.shape
    dims = [d - P + 1 for d in v.shape] if padded else list(v.shape)
    if min(dims) < 1:
        raise ValueError("volume smaller than the patch")
    out = torch.empty(dims + [T], dtype=v.dtype, device=v.device)
    with torch.cuda.device(v.device):
        check(lib().hipr_lne3d_dirs(C.c_void_p(v.data_ptr()), Xs, Ys, Zs, int(bool(padded)), _DT[v.dtype], P, T,
                                    _tab_ptr(tab), maxkey.ptr() if maxkey is not None else None,
                                    C.c_void_p(out.data_ptr()), _stream()), "lne3d_dirs")
    return out


def lne3d_dirs_host(volume_padded, patch_size=11, theta_range=9, phi_range=9, out=None, pinned=False):
    """numpy float64 (Xp, Yp, Zp) -> numpy float64 (X, Y, Z, T) through hipr_lne3d_dirs_host
    (line_profile_memory_efficient_v2 for host arrays: bands of x-planes under the device -> host copy)."""
    a = np.ascontiguousarray(volume_padded)
    if a.dtype != np.float64:
        raise TypeError("image_padded must be float64, got %s" % a.dtype)
    if a.ndim != 3:
        raise ValueError("Buffer has wrong number of dimensions (expected 3, got %d)" % a.ndim)
    tab = tables.line_table_3d(patch_size, theta_range, phi_range)
    T, P = tab.shape[0], tab.shape[1]
    Xp, Yp, Zp = a.shape
    if min(a.shape) < P:
        raise ValueError("volume smaller than the patch")
    shape = (Xp - P + 1, Yp - P + 1, Zp - P + 1, T)
    if out is None:
        out = torch.empty(shape, dtype=torch.float64, pin_memory=True).numpy() if pinned else np.empty(shape, dtype=np.float64)
    elif out.shape != shape or out.dtype != np.float64 or not out.flags.c_contiguous:
        raise ValueError("out must be a C-contiguous float64 array of shape %s" % (shape,))
    check(lib().hipr_lne3d_dirs_host(a.ctypes.data_as(C.c_void_p), Xp, Yp, Zp, P, T, _tab_ptr(tab),
                                     out.ctypes.data_as(C.c_void_p)), "lne3d_dirs_host")
    return out


def line_profile_3d_host(volume_padded, patch_size=11, theta_range=9, phi_range=9, out=None, pinned=False):
    """numpy float64 (Xp, Yp, Zp) -> numpy float64 (X, Y, Z, T, P) through hipr_line_profile_3d_host
    (line_profile_v2 for host arrays: bands of x-planes under the device -> host copy)."""
    a = np.ascontiguousarray(volume_padded)
    if a.dtype != np.float64:
        raise TypeError("image_padded must be float64, got %s" % a.dtype)
    if a.ndim != 3:
        raise ValueError("Buffer has wrong number of dimensions (expected 3, got %d)" % a.ndim)
    tab = tables.line_table_3d(patch_size, theta_range, phi_range)
    T, P = tab.shape[0], tab.shape[1]
    Xp, Yp, Zp = a.shape
    if min(a.shape) < P:
        raise ValueError("volume smaller than the patch")
    shape = (Xp - P + 1, Yp - P + 1, Zp - P + 1, T, P)
    if out is None:
        out = torch.empty(shape, dtype=torch.float64, pin_memory=True).numpy() if pinned else np.empty(shape, dtype=np.float64)
    elif out.shape != shape or out.dtype != np.float64 or not out.flags.c_contiguous:
        raise ValueError("out must be a C-contiguous float64 array of shape %s" % (shape,))
    check(lib().hipr_line_profile_3d_host(a.ctypes.data_as(C.c_void_p), Xp, Yp, Zp, P, T, _tab_ptr(tab),
                                     out.ctypes.data_as(C.c_void_p)), "line_profile_3d_host")
    return out


def lne3d(volume, flavour="F2", patch_size=11, theta_range=9, phi_range=9, padded=False, maxkey=None):
    """3-D score map (X, Y, Z).  flavour 'F2' / 'F3' / 'ME2', or 'V3' (needs padded=True)."""
    v = _vol(volume, "volume")
    if flavour == "V3":
        tab = tables.line_table_3d_v3(patch_size, theta_range, phi_range)
    else:
        tab = tables.line_table_3d(patch_size, theta_range, phi_range)
    T, P = tab.shape[0], tab.shape[1]
    Xs, Ys, Zs = v.shape
    dims = [d - P + 1 for d in v.shape] if padded else list(v.shape)
    if min(dims) < 1:
        raise ValueError("volume smaller than the patch")
    out = torch.empty(dims, dtype=v.dtype, device=v.device)
    with torch.cuda.device(v.device):
        check(lib().hipr_lne3d(C.c_void_p(v.data_ptr()), Xs, Ys, Zs, int(bool(padded)), _DT[v.dtype], P, T,
                               _tab_ptr(tab), _flavour(flavour), maxkey.ptr() if maxkey is not None else None,
                               C.c_void_p(out.data_ptr()), _stream()), "lne3d")
    return out


def lne3d_fixed(volume, flavour="ME2", patch_size=11, theta_range=9, phi_range=9, padded=False, maxkey=None,
                dirs_only=False):
    """lne3d / lne3d_dirs on a 31-bit fixed-point copy of the volume (brick-local range, exact
    differences; csrc/lne3d.cu).  float32 out.  maxkey scales the 1e-8 epsilons of F3 / ME2 (None: the
    volume is already normalised).  Returns None outside (11, 9, 9)."""
    v = _vol(volume, "volume")
    tab = tables.line_table_3d(patch_size, theta_range, phi_range)
    T, P = tab.shape[0], tab.shape[1]
    Xs, Ys, Zs = v.shape
    dims = [d - P + 1 for d in v.shape] if padded else list(v.shape)
    if min(dims) < 1:
        raise ValueError("volume smaller than the patch")
    out = torch.empty(dims + ([T] if dirs_only else []), dtype=torch.float32, device=v.device)
    with torch.cuda.device(v.device):
        code = lib().hipr_lne3d_q(C.c_void_p(v.data_ptr()), Xs, Ys, Zs, int(bool(padded)), _DT[v.dtype], P, T,
                                  _tab_ptr(tab), _flavour(flavour), int(bool(dirs_only)),
                                  maxkey.ptr() if maxkey is not None else None, C.c_void_p(out.data_ptr()), _stream())
    if code == UNSUPPORTED:
        return None
    check(code, "lne3d_fixed")
    return out


def neighbor3d_score(cube, flavour="ME2", patch_size=11, theta_range=9, phi_range=9, dtype=None, denoise_h=None,
                     denoise_distance=11):
    """cube (X, Y, Z, C) float32 -> score volume (X, Y, Z): channel sum -> /max -> edge pad ->
    3-D line profiles -> epilogue (bio/...analysis.py:807-817 for 'ME2').  dtype=None: float64 sums +
    fixed-point stencil (float32 score); torch.float32 / float64: floating-point stencil of that type.
    A batch (N, X, Y, Z, C) of independent z-stacks alternates between two streams, so that the channel sum of
    stack i+1 (HBM-bound) runs beside the stencil of stack i (ALU-bound); returns (N, X, Y, Z).
    denoise_h: with the NL-means denoise of bio/...analysis.py:454 (h = 0.03 there) between the normalisation and the
    stencil (csrc/nlm3d.cu, float64 stencil after it)."""
    cube = _dev(cube, "cube", (torch.float32,))
    if cube.dim() == 5:
        cur = torch.cuda.current_stream(cube.device)
        side = _side_streams(cube.device)
        for st in side:
            st.wait_stream(cur)
        res = []
        for i, c in enumerate(cube):
            with torch.cuda.stream(side[i % len(side)]):
                res.append(neighbor3d_score(c, flavour, patch_size, theta_range, phi_range, dtype, denoise_h, denoise_distance))
        for st in side:
            cur.wait_stream(st)
        for r in res:
            r.record_stream(cur)
        return torch.stack(res)
    if cube.dim() != 4:
        raise ValueError("cube must be (X, Y, Z, C) or (N, X, Y, Z, C)")
    if denoise_h is not None:
        # bio/...analysis.py:452-462 in full: sum -> /max -> denoise_nl_means(h) -> edge pad -> me_v2 -> epilogue; the
        # denoised volume is too smooth for the fixed-point grid (as in 2-D): float64 stencil
        s = channel_sum(cube, None, normalize=True, dtype=torch.float64)
        den = denoise_nl_means(s, patch_distance=denoise_distance, h=denoise_h)
        return lne3d(den, flavour, patch_size, theta_range, phi_range, padded=False)
    s, mk = channel_sum(cube, None, normalize=False, dtype=dtype or torch.float64, return_max=True)
    if dtype is None:
        res = lne3d_fixed(s, flavour, patch_size, theta_range, phi_range, padded=False, maxkey=mk)
        if res is not None:
            return res
    return lne3d(s, flavour, patch_size, theta_range, phi_range, padded=False, maxkey=mk)


# ---------------------------------------------------------------------------------------------
# per-cell spectra
# ---------------------------------------------------------------------------------------------

def label_max(labels):
    labels = _dev(labels, "labels", (torch.int32, torch.int64))
    out = torch.empty(1, dtype=torch.int64, device=labels.device)
    with torch.cuda.device(labels.device):
        check(lib().hipr_label_max(C.c_void_p(labels.data_ptr()), labels.element_size(), labels.numel(),
                                   C.c_void_p(out.data_ptr()), _stream()), "label_max")
    return out


def cell_spectra_accumulate(cube, labels, max_label, sums=None, counts=None):
    """Adds this (slab of a) label image's per-label channel sums and pixel counts into
    sums (max_label+1, C) float64 / counts (max_label+1) int32 (allocated zeroed if None)."""
    cube = _dev(cube, "cube", (torch.float32,))
    labels = _dev(labels, "labels", (torch.int32, torch.int64))
    Cn = cube.shape[-1]
    npix = labels.numel()
    if cube.numel() != npix * Cn:
        raise ValueError("cube %s and labels %s do not match" % (tuple(cube.shape), tuple(labels.shape)))
    max_label = int(max_label)
    if (sums is None) != (counts is None):
        raise ValueError("pass both sums and counts (to keep accumulating) or neither")
    fresh = sums is None
    if sums is None:
        sums = torch.empty((max_label + 1, Cn), dtype=torch.float64, device=cube.device)
    if counts is None:
        counts = torch.empty(max_label + 1, dtype=torch.int32, device=cube.device)
    with torch.cuda.device(cube.device):
        if fresh:
            check(lib().hipr_cell_spectra_reset(C.c_void_p(sums.data_ptr()), C.c_void_p(counts.data_ptr()), max_label,
                                                Cn, _stream()), "cell_spectra_reset")
        check(lib().hipr_cell_spectra_accumulate(C.c_void_p(cube.data_ptr()), C.c_void_p(labels.data_ptr()),
                                                 labels.element_size(), npix,
                                                 labels.shape[-1] if labels.dim() > 1 else 0, Cn, max_label,
                                                 C.c_void_p(sums.data_ptr()), C.c_void_p(counts.data_ptr()), None,
                                                 _stream()), "cell_spectra_accumulate")
    return sums, counts


def cell_spectra_finalize(sums, counts):
    """-> (labels int64 (n,), area int64 (n,), avgint float64 (n, C), avgint_norm float64 (n, C)),
    rows in ascending order of the labels present.  Synchronises once to learn n."""
    max_label, Cn = sums.shape[0] - 1, sums.shape[1]
    dev = sums.device
    cap = max(max_label, 1)
    n_cells = torch.zeros(1, dtype=torch.int32, device=dev)
    labels = torch.empty(cap, dtype=torch.int64, device=dev)
    area = torch.empty(cap, dtype=torch.int64, device=dev)
    avg = torch.empty((cap, Cn), dtype=torch.float64, device=dev)
    norm = torch.empty((cap, Cn), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        check(lib().hipr_cell_spectra_finalize(C.c_void_p(sums.data_ptr()), C.c_void_p(counts.data_ptr()), max_label,
                                               Cn, C.c_void_p(n_cells.data_ptr()), C.c_void_p(labels.data_ptr()),
                                               C.c_void_p(area.data_ptr()), C.c_void_p(avg.data_ptr()),
                                               C.c_void_p(norm.data_ptr()), _stream()), "cell_spectra_finalize")
    n = int(n_cells.item())
    return labels[:n], area[:n], avg[:n], norm[:n]


def cell_spectra(cube, labels, max_label=None):
    """regionprops-equivalent per-cell mean spectra of cube (..., C) over the label image."""
    if max_label is None:
        max_label = int(label_max(labels).item())
    sums, counts = cell_spectra_accumulate(cube, labels, max_label)
    return cell_spectra_finalize(sums, counts)


GEOMETRY_COLUMNS = ("centroid_row", "centroid_col", "major_axis_length", "minor_axis_length", "eccentricity",
                    "orientation", "mu20", "mu02", "mu11")


def cell_geometry(labels, max_label=None):
    """regionprops(segmentation) geometry of a 2-D label image: -> (labels int64 (n,), area int64 (n,),
    geometry float64 (n, 9) with columns GEOMETRY_COLUMNS), rows in ascending order of the labels present
    (syn/hiprfish_imaging_classify_spectra.py:38-46).  Integer raw moments on the device are exact."""
    labels = _dev(labels, "labels", (torch.int32, torch.int64))
    if labels.dim() != 2:
        raise ValueError("labels must be a 2-D label image")
    if max_label is None:
        max_label = int(label_max(labels).item())
    max_label = int(max_label)
    dev = labels.device
    Hh, Ww = labels.shape
    cap = max(max_label, 1)
    mom = torch.empty((max_label + 1, 6), dtype=torch.int64, device=dev)
    scratch = torch.empty(max_label + 1, dtype=torch.int32, device=dev)
    n_cells = torch.zeros(1, dtype=torch.int32, device=dev)
    lab = torch.empty(cap, dtype=torch.int64, device=dev)
    area = torch.empty(cap, dtype=torch.int64, device=dev)
    geom = torch.empty((cap, 9), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        check(lib().hipr_cell_moments(C.c_void_p(labels.data_ptr()), labels.element_size(), Hh, Ww, max_label,
                                      C.c_void_p(mom.data_ptr()), _stream()), "cell_moments")
        check(lib().hipr_cell_geometry_finalize(C.c_void_p(mom.data_ptr()), max_label, C.c_void_p(scratch.data_ptr()),
                                                C.c_void_p(n_cells.data_ptr()), C.c_void_p(lab.data_ptr()),
                                                C.c_void_p(area.data_ptr()), C.c_void_p(geom.data_ptr()), _stream()),
              "cell_geometry_finalize")
    n = int(n_cells.item())
    return lab[:n], area[:n], geom[:n]


def orientation_xy(geometry):
    """Orientation column in the 'xy' coordinate convention of scikit-image <= 0.15 (regionprops' default before
    0.16, the era the reference's cpython-35 artefacts point to): -0.5 * atan2(2 mu11, var_col - var_row), from the
    central moments cell_geometry returns (columns mu20 = row variance, mu02 = column variance, mu11).  cell_geometry's
    own `orientation` column follows the 'rc' convention of scikit-image >= 0.16.  Neither is pinned by the reference
    (no version file, scikit-image not installed here): PARITY UNPINNED, see INTEGRATION.md.  geometry: the (n, 9)
    table (torch tensor or numpy array) -> (n,) of the same kind."""
    if isinstance(geometry, torch.Tensor):
        mu20, mu02, mu11 = geometry[:, 6], geometry[:, 7], geometry[:, 8]
        return -0.5 * torch.atan2(2.0 * mu11, mu02 - mu20)
    g = np.asarray(geometry)
    return -0.5 * np.arctan2(2.0 * g[:, 8], g[:, 7] - g[:, 6])


def paint_labels(labels, values, max_label=None):
    """out[p] = values[labels[p]]: the `image[segmentation == label] = value` loops of
    eco/hiprfish_imaging_image_classification.py:64-70 / bio/...analysis.py:1247-1257 as one gather.
    values: (max_label + 1,) or (max_label + 1, K) float32 / float64, row 0 = background."""
    labels = _dev(labels, "labels", (torch.int32, torch.int64))
    values = _dev(values, "values")
    squeeze = values.dim() == 1
    v2 = values.reshape(values.shape[0], -1)
    if max_label is None:
        max_label = v2.shape[0] - 1
    if v2.shape[0] < max_label + 1:
        raise ValueError("values needs max_label + 1 rows")
    K = v2.shape[1]
    out = torch.empty(tuple(labels.shape) + (() if squeeze else (K,)), dtype=values.dtype, device=labels.device)
    with torch.cuda.device(labels.device):
        check(lib().hipr_paint_labels(C.c_void_p(labels.data_ptr()), labels.element_size(), labels.numel(),
                                      C.c_void_p(v2.data_ptr()), K, int(max_label), _DT[values.dtype],
                                      C.c_void_p(out.data_ptr()), _stream()), "paint_labels")
    return out


# ---------------------------------------------------------------------------------------------
# 1-D k-means thresholding
# ---------------------------------------------------------------------------------------------

KMEANS_TRANSFORMS = {None: 0, "none": 0, "log10": 1, "log": 2}


class KMeans1DResult:
    """cluster_centers_ (k,), counts (k,), positive_means (k,), labels (int32 tensor, image shape) or None, mask
    (bool tensor: the brightest cluster) or None, n_iter, inertia, bright (label of the brightest cluster),
    n_samples, best_init."""

    def __init__(self, **kw):
        self.__dict__.update(kw)


def kmeans_threshold(image, n_clusters=2, random_state=0, n_init=1, transform=None, eps=0.0, positive_only=False,
                     max_iter=300, tol=1e-4, return_labels=True, return_mask=True, fill_label=0):
    """KMeans(n_clusters, random_state=random_state).fit_predict(f(image).reshape(-1, 1)) of scikit-learn 1.9 on the
    device, plus the scripts' mask orientation (syn/...measurement.py:125-137): f = identity, log10(image + eps)
    (transform='log10') or ln(image + eps) ('log'); positive_only: cluster image[image > 0] only, the other pixels
    get fill_label / mask 0 (bio/...analysis.py:819).  image: float32 / float64 CUDA tensor of any shape.
    Synchronises once (to read the 32-double result)."""
    image = _dev(image, "image")
    k = int(n_clusters)
    if not isinstance(random_state, (int, np.integer)):
        raise TypeError("random_state must be an integer seed (the reference passes 0)")
    nu = lib().hipr_kmeans1d_uniforms(k, int(n_init))
    check(nu if nu < 0 else 0, "kmeans_threshold")
    # scikit-learn's draws, in its order: RandomState(seed).choice (one double) then uniform(size=trials) per seed
    u = np.ascontiguousarray(np.random.RandomState(int(random_state)).random_sample(nu), dtype=np.float64)
    try:
        tcode = KMEANS_TRANSFORMS[transform]
    except KeyError:
        raise ValueError("transform must be None, 'log10' or 'log'")
    dev = image.device
    work = torch.empty(int(lib().hipr_kmeans1d_workspace_bytes()), dtype=torch.uint8, device=dev)
    labels = torch.empty(image.shape, dtype=torch.int32, device=dev) if return_labels else None
    mask = torch.empty(image.shape, dtype=torch.uint8, device=dev) if return_mask else None
    res = torch.empty(32, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        check(lib().hipr_kmeans1d(C.c_void_p(image.data_ptr()), _DT[image.dtype], image.numel(), tcode, float(eps),
                                  int(bool(positive_only)), k, int(n_init), u.ctypes.data_as(C.c_void_p), int(max_iter),
                                  float(tol), C.c_void_p(work.data_ptr()),
                                  C.c_void_p(labels.data_ptr()) if labels is not None else None, int(fill_label),
                                  C.c_void_p(mask.data_ptr()) if mask is not None else None,
                                  C.c_void_p(res.data_ptr()), _stream()), "kmeans_threshold")
    r = res.cpu().numpy()
    status = int(r[0])
    if status == 2:
        raise ValueError("Input X contains NaN or infinity.")          # scikit-learn's check_array message
    if status == 4:
        raise ValueError("n_samples=%d should be >= n_clusters=%d." % (int(r[1]), k))
    if status == 3:
        raise RuntimeError("a cluster ran empty during Lloyd iterations (scikit-learn would relocate it; not reproduced)")
    return KMeans1DResult(cluster_centers_=r[8:8 + k].copy(), counts=r[16:16 + k].astype(np.int64),
                          positive_means=r[24:24 + k].copy(), labels=labels, mask=mask.view(torch.bool) if mask is not None else None,
                          n_iter=int(r[4]), inertia=float(r[5]), bright=int(r[7]), n_samples=int(r[1]), best_init=int(r[6]))


# ---------------------------------------------------------------------------------------------
# host-buffer (numpy) entry points: copies happen inside the library
# ---------------------------------------------------------------------------------------------

def pinned_empty(shape, dtype=np.float32):
    """numpy array backed by page-locked memory from hipr_host_alloc (freed with the array)."""
    dtype = np.dtype(dtype)
    nbytes = int(np.prod(shape)) * dtype.itemsize
    p = C.c_void_p()
    check(lib().hipr_host_alloc(C.byref(p), max(nbytes, 1)), "host_alloc")

    class _Owner:
        def __init__(self, ptr):
            self.ptr = ptr

        def __del__(self):
            try:
                lib().hipr_host_free(self.ptr)
            except Exception:
                pass

    buf = (C.c_char * max(nbytes, 1)).from_address(p.value)
    buf._hipr_owner = _Owner(p)
    return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)


def neighbor2d_score_host(cube, flavour="F1", patch_size=11, phi_range=9, return_sum=False, out=None, denoise_h=None):
    """numpy (H, W, C) float32 -> numpy (H, W) float32 score map, through hipr_neighbor2d_host (denoise_h: with the
    NL-means denoise of syn/...measurement.py:108 in the chain, hipr_neighbor2d_host_denoise)."""
    cube = np.ascontiguousarray(cube)
    if cube.dtype != np.float32:
        raise TypeError("cube must be float32, got %s" % cube.dtype)
    if cube.ndim != 3:
        raise ValueError("cube must be (H, W, C)")
    H, W, Cn = cube.shape
    tab = tables.line_table_2d(patch_size, phi_range)
    score = out if out is not None else np.empty((H, W), dtype=np.float32)
    s = np.empty((H, W), dtype=np.float32) if return_sum else None
    if denoise_h is not None:
        check(lib().hipr_neighbor2d_host_denoise(cube.ctypes.data_as(C.c_void_p), H, W, Cn, tab.shape[1], tab.shape[0],
                                                 _tab_ptr(tab), _flavour(flavour), float(denoise_h),
                                                 score.ctypes.data_as(C.c_void_p),
                                                 s.ctypes.data_as(C.c_void_p) if return_sum else None),
              "neighbor2d_host_denoise")
        return (score, s) if return_sum else score
    check(lib().hipr_neighbor2d_host(cube.ctypes.data_as(C.c_void_p), H, W, Cn, tab.shape[1], tab.shape[0],
                                     _tab_ptr(tab), _flavour(flavour), score.ctypes.data_as(C.c_void_p),
                                     s.ctypes.data_as(C.c_void_p) if return_sum else None), "neighbor2d_host")
    return (score, s) if return_sum else score


def neighbor2d_score_host_batch(cubes, flavour="F1", patch_size=11, phi_range=9, denoise_h=None, out=None):
    """A list of numpy (H, W, C) float32 cubes -> list of numpy (H, W) float32 score maps through
    hipr_neighbor2d_host_batch: FOV i + 1 crosses PCIe while FOV i is denoised (denoise_h), scored and read back."""
    cubes = [np.ascontiguousarray(c) for c in cubes]
    if not cubes:
        raise ValueError("empty batch")
    H, W, Cn = cubes[0].shape
    for c in cubes:
        if c.dtype != np.float32:
            raise TypeError("cube must be float32, got %s" % c.dtype)
        if c.shape != (H, W, Cn):
            raise ValueError("every cube of a batch must have the same (H, W, C)")
    n = len(cubes)
    tab = tables.line_table_2d(patch_size, phi_range)
    scores = out if out is not None else [np.empty((H, W), dtype=np.float32) for _ in range(n)]
    if len(scores) != n:
        raise ValueError("one output array per cube")
    cp = (C.c_void_p * n)(*[c.ctypes.data for c in cubes])
    sp = (C.c_void_p * n)(*[s.ctypes.data for s in scores])
    check(lib().hipr_neighbor2d_host_batch(cp, n, H, W, Cn, tab.shape[1], tab.shape[0], _tab_ptr(tab), _flavour(flavour),
                                           float(denoise_h) if denoise_h else 0.0, sp), "neighbor2d_host_batch")
    return scores


def neighbor3d_score_host(cube, flavour="ME2", patch_size=11, theta_range=9, phi_range=9, out=None, denoise_h=None,
                          denoise_distance=11):
    """numpy (X, Y, Z, C) float32 -> numpy (X, Y, Z) float32 score volume, through hipr_neighbor3d_host
    (bio/...analysis.py:807-817 for 'ME2'); denoise_h: with the 3-D NL-means of bio/...analysis.py:454 in the chain
    (hipr_neighbor3d_host_denoise)."""
    cube = np.ascontiguousarray(cube)
    if cube.dtype != np.float32:
        raise TypeError("cube must be float32, got %s" % cube.dtype)
    if cube.ndim != 4:
        raise ValueError("cube must be (X, Y, Z, C)")
    X, Y, Z, Cn = cube.shape
    tab = tables.line_table_3d(patch_size, theta_range, phi_range)
    score = out if out is not None else np.empty((X, Y, Z), dtype=np.float32)
    if denoise_h is not None:
        check(lib().hipr_neighbor3d_host_denoise(cube.ctypes.data_as(C.c_void_p), X, Y, Z, Cn, tab.shape[1], tab.shape[0],
                                                 _tab_ptr(tab), _flavour(flavour), float(denoise_h), int(denoise_distance),
                                                 score.ctypes.data_as(C.c_void_p)), "neighbor3d_host_denoise")
        return score
    check(lib().hipr_neighbor3d_host(cube.ctypes.data_as(C.c_void_p), X, Y, Z, Cn, tab.shape[1], tab.shape[0],
                                     _tab_ptr(tab), _flavour(flavour), score.ctypes.data_as(C.c_void_p)),
          "neighbor3d_host")
    return score


def neighbor2d_score_host_raw(cube, scale, flavour="F1", patch_size=11, phi_range=9, out=None):
    """numpy (H, W, C) uint16 / uint8 raw counts -> numpy (H, W) float32 score map, the score of the cube
    bioformats' rescale would produce (count / scale in float32), through hipr_neighbor2d_host_raw."""
    cube = np.ascontiguousarray(cube)
    if cube.dtype not in (np.uint8, np.uint16):
        raise TypeError("raw cube must be uint8 or uint16, got %s" % cube.dtype)
    if cube.ndim != 3:
        raise ValueError("cube must be (H, W, C)")
    H, W, Cn = cube.shape
    tab = tables.line_table_2d(patch_size, phi_range)
    score = out if out is not None else np.empty((H, W), dtype=np.float32)
    check(lib().hipr_neighbor2d_host_raw(cube.ctypes.data_as(C.c_void_p), cube.itemsize, float(scale), H, W, Cn,
                                         tab.shape[1], tab.shape[0], _tab_ptr(tab), _flavour(flavour),
                                         score.ctypes.data_as(C.c_void_p), None), "neighbor2d_host_raw")
    return score


def cell_spectra_host(cube, labels):
    """numpy cube (..., C) float32 + integer label image -> (labels, area, avgint, avgint_norm)."""
    cube = np.ascontiguousarray(cube)
    labels = np.ascontiguousarray(labels)
    if cube.dtype != np.float32:
        raise TypeError("cube must be float32, got %s" % cube.dtype)
    if labels.dtype not in (np.int32, np.int64):
        raise TypeError("labels must be int32 or int64, got %s" % labels.dtype)
    Cn = cube.shape[-1]
    npix = labels.size
    if cube.size != npix * Cn:
        raise ValueError("cube and labels do not match")
    cap = 4096

    def alloc(k):
        return np.empty(k, np.int64), np.empty(k, np.int64), np.empty((k, Cn), np.float64), np.empty((k, Cn), np.float64)

    n = C.c_int64(0)
    lab, area, avg, norm = alloc(cap)
    code = lib().hipr_cell_spectra_host(cube.ctypes.data_as(C.c_void_p), labels.ctypes.data_as(C.c_void_p),
                                        labels.itemsize, npix, labels.shape[-1] if labels.ndim > 1 else 0, Cn, cap,
                                        C.byref(n),
                                        lab.ctypes.data_as(C.c_void_p), area.ctypes.data_as(C.c_void_p),
                                        avg.ctypes.data_as(C.c_void_p), norm.ctypes.data_as(C.c_void_p))
    if code == -7 and n.value > cap:
        # more cells than the first guess: the table is still on the device, fetch it (no second pass over the cube)
        cap = int(n.value)
        lab, area, avg, norm = alloc(cap)
        code = lib().hipr_cell_spectra_host_fetch(cap, lab.ctypes.data_as(C.c_void_p), area.ctypes.data_as(C.c_void_p),
                                                  avg.ctypes.data_as(C.c_void_p), norm.ctypes.data_as(C.c_void_p))
    check(code, "cell_spectra_host")
    k = int(n.value)
    return lab[:k], area[:k], avg[:k], norm[:k]


class Fov:
    """One upload of a field of view for both steps of `measure_biofilm_images_no_reference`
    (syn/...measurement.py:161-173): the (H, W, C) float32 numpy cube is streamed to the current CUDA device once and
    stays resident; `score()` is lines 105-124 (without the denoise), `cell_spectra(segmentation)` lines 167-172.

        with hipr_b200.Fov(image_registered) as fov:
            image_final = fov.score("F1")
            ...                                  # watershed on the host
            labels, area, avgint, avgint_norm = fov.cell_spectra(image_seg)
    """

    def __init__(self, cube):
        cube = np.ascontiguousarray(cube)
        if cube.dtype != np.float32:
            raise TypeError("cube must be float32, got %s" % cube.dtype)
        if cube.ndim != 3:
            raise ValueError("cube must be (H, W, C)")
        self.shape = cube.shape
        h = C.c_void_p()
        check(lib().hipr_fov_upload(cube.ctypes.data_as(C.c_void_p), cube.shape[0], cube.shape[1], cube.shape[2],
                                    C.byref(h)), "fov_upload")
        self._h = h

    def _handle(self):
        if self._h is None:
            raise ValueError("this field of view has been released")
        return self._h

    def score(self, flavour="F1", patch_size=11, phi_range=9, return_sum=False, out=None):
        H, W, _ = self.shape
        tab = tables.line_table_2d(patch_size, phi_range)
        score = out if out is not None else np.empty((H, W), dtype=np.float32)
        s = np.empty((H, W), dtype=np.float32) if return_sum else None
        check(lib().hipr_fov_score(self._handle(), tab.shape[1], tab.shape[0], _tab_ptr(tab), _flavour(flavour),
                                   score.ctypes.data_as(C.c_void_p), s.ctypes.data_as(C.c_void_p) if return_sum else None),
              "fov_score")
        return (score, s) if return_sum else score

    def cell_spectra(self, labels):
        labels = np.ascontiguousarray(labels)
        if labels.dtype not in (np.int32, np.int64):
            raise TypeError("labels must be int32 or int64, got %s" % labels.dtype)
        H, W, Cn = self.shape
        if labels.shape != (H, W):
            raise ValueError("labels must be (%d, %d)" % (H, W))
        cap = 4096
        while True:
            lab, area = np.empty(cap, np.int64), np.empty(cap, np.int64)
            avg, norm = np.empty((cap, Cn), np.float64), np.empty((cap, Cn), np.float64)
            n = C.c_int64(0)
            code = lib().hipr_fov_cell_spectra(self._handle(), labels.ctypes.data_as(C.c_void_p), labels.itemsize, cap,
                                               C.byref(n), lab.ctypes.data_as(C.c_void_p), area.ctypes.data_as(C.c_void_p),
                                               avg.ctypes.data_as(C.c_void_p), norm.ctypes.data_as(C.c_void_p))
            if code == -7 and n.value > cap:
                cap = int(n.value)           # the cube is resident: the second call is one pass over it, no upload
                continue
            check(code, "fov_cell_spectra")
            k = int(n.value)
            return lab[:k], area[:k], avg[:k], norm[:k]

    def device_arrays(self):
        """(cube, channel sums, range keys) device pointers as ints, for callers that continue on the device."""
        a, b, c = C.c_void_p(), C.c_void_p(), C.c_void_p()
        check(lib().hipr_fov_device_arrays(self._handle(), C.byref(a), C.byref(b), C.byref(c)), "fov_device_arrays")
        return a.value, b.value, c.value

    def close(self):
        if self._h is not None:
            lib().hipr_fov_release(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

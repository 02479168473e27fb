"""hipr_b200: B200-native HiPR-FISH spectral-segmentation front end (host side).

Python mirror of the reference's operator surface for this path; all compute is hand-written
sm_100a CUDA in libhipr_b200.so reached through a C ABI (include/hipr_b200.h)."""
from . import tables  # noqa: F401
from ._lib import EXPORTS, LIB_PATH, HiprError, lib  # noqa: F401


def __getattr__(name):
    # torch is imported lazily so that table construction and ABI checks work without it
    import importlib
    if name.startswith("__"):
        raise AttributeError(name)
    ops = importlib.import_module(__name__ + ".ops")
    if name == "ops":
        return ops
    if hasattr(ops, name):
        return getattr(ops, name)
    raise AttributeError(name)

"""sos-slam_b200 — B200-native photometric bundle adjustment / direct alignment (the SOS-SLAM hot path).

The compute lives in csrc/ (hand-written sm_100a CUDA behind the C ABI of include/sosba.h, built into
csrc/libsosba.so).  This Python package is plumbing for tests and benchmarks: a ctypes binding of the
C ABI, the synthetic window generator and the point-sharding helper.  There is no CPU fallback:
`load()` raises if the CUDA library is missing.
"""
import os

from . import binding, problem, synth  # noqa: F401
from .binding import Handle, Lib, SosbaError  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libsosba.so")
_lib = None


def load() -> Lib:
    """The product library.  Fails loudly when the CUDA extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SosbaError(f"CUDA library {LIB_PATH} not built; run __graft_entry__.build(). There is no CPU fallback.")
        _lib = Lib(LIB_PATH, "sosba")
    return _lib

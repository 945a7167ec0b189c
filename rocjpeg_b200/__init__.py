"""rocjpeg_b200 — a B200-native (sm_100a) baseline-JPEG decoder behind rocJPEG's C API.

The product is the shared library rocjpeg_b200/lib/librocjpeg.so (sources in
rocjpeg_b200/csrc/, public headers in include/): a drop-in for the reference's
librocjpeg.so whose decode path is hand-written CUDA. `rocjpeg_b200.api` is a thin
ctypes mirror of that C API for tests and bench.py; `rocjpeg_b200.datagen` makes
the deterministic synthetic inputs. Nothing here imports the CPU oracle.
"""
from . import api  # noqa: F401

__all__ = ["api"]
__version__ = "0.6.0"

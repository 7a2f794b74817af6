"""libplacebo_b200 — B200-native libplacebo hot path (C ABI in libplacebo_b200.so).

The product is the shared library next to this file (built by `__graft_entry__.build()` /
`make -C libplacebo_b200/csrc`); its interface is `include/libplacebo_b200.h`. This package
only locates the library and offers the thin Python plumbing used by bench.py and the
multi-GPU driver (device memory via torch, torch.distributed for the statistics exchange).
There is no Python or CPU fallback for the pixel path: if the CUDA library is missing,
`load()` raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libplacebo_b200.so")
INCLUDE_DIR = os.path.join(os.path.dirname(_HERE), "include")

_lib = None


def load():
    """Load libplacebo_b200.so with local symbol scope (it exports libplacebo's own pl_* names,
    which must not interpose a real libplacebo living in the same process)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` (needs nvcc). libplacebo_b200 has no CPU fallback.")
        _lib = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_LOCAL)
    return _lib

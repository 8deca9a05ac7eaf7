"""salviarenderer_b200 — B200-native implementation of SALVIA's draw pipeline.

The product is `csrc/libsalvia_b200.so` (hand-written sm_100a CUDA behind the C ABI of
include/salvia_b200.h).  `load()` returns a ctypes binding to it and raises if it is missing or if no
CUDA device is usable: there is NO CPU fallback in this package.
"""
from __future__ import annotations

import os

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
PRODUCT_LIB = os.path.join(_HERE, "csrc", "libsalvia_b200.so")


def load(ordinal: int = 0) -> abi.Backend:
    """Open the CUDA product library on CUDA device `ordinal`. Fails loudly when it is unavailable."""
    if not os.path.exists(PRODUCT_LIB):
        raise abi.SlvError(
            f"{PRODUCT_LIB} is not built — run `python -c 'import __graft_entry__ as g; g.build()'`. "
            "There is no CPU fallback.")
    return abi.Backend(PRODUCT_LIB, ordinal)

"""Developer tool: distribution of the per-tile and per-(region, warp block) list lengths of one frame (the in-order
chains k_cover walks).  Uses the product-only debug entry point slv_debug_read."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import salviarenderer_b200 as pkg  # noqa: E402
from salviarenderer_b200 import scenes as S  # noqa: E402

be = pkg.load(0)
be.lib.slv_debug_read.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_size_t]
be.lib.slv_debug_read.restype = C.c_int32
sc = S.SponzaLike(3840, 2160, 4)
sc.setup(be)
n_tiles = 60 * 34
for f in (0, 3, 7):
    sc.render(be, f)
    be.flush()
    act = np.zeros(n_tiles + 1, np.uint32)
    assert be.lib.slv_debug_read(be.dev, 0, act.ctypes.data, act.nbytes) == 0
    off = np.zeros(n_tiles + 1, np.uint32)
    assert be.lib.slv_debug_read(be.dev, 1, off.ctypes.data, off.nbytes) == 0
    na = int(act[0])
    bd = np.zeros(na * 128 * 2, np.uint32)
    assert be.lib.slv_debug_read(be.dev, 2, bd.ctypes.data, bd.nbytes) == 0
    bc = bd[1::2]
    tl = np.diff(off.astype(np.int64))
    q = [50, 90, 99, 99.9, 100]
    print(f"frame {f}: active tiles {na}; tile list len: total {tl.sum()} pct{q} = {np.percentile(tl, q).astype(int).tolist()}")
    print(f"          warp-block list len: total {bc.sum()} nonzero {np.count_nonzero(bc)} pct{q} = {np.percentile(bc, q).astype(int).tolist()}; "
          f"top 10 {np.sort(bc)[-10:].tolist()}; items > 256: {(bc > 256).sum()}, > 1024: {(bc > 1024).sum()}")
    # work share of the longest chains
    s = np.sort(bc)[::-1].astype(np.int64)
    print(f"          share of pairs in the longest 1% of items: {s[:len(s)//100].sum() / max(s.sum(), 1):.2%}")

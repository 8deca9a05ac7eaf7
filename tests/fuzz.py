"""Test infrastructure: seeded random points of the TriangleSoup parameter space (tests/cases.py holds the hand-picked ones): target
size (multiples of 4: SURVEY Appendix B #16), sample count, triangle count and size, cull mode, depth function / write / enable,
two-sided stencil with random ops, functions, masks and reference, blend shader, colour format, list / strip, index width,
base vertex, indexed or not, split draws, attribute modifiers.  The same seed gives the same scene on every backend."""
import numpy as np

from salviarenderer_b200 import abi as A, scenes as S


def soup_from_seed(seed: int):
    r = np.random.default_rng(1000 + seed)
    kw = {}
    kw["w"], kw["h"] = int(r.integers(16, 128)) * 4, int(r.integers(12, 96)) * 4
    kw["samples"] = int(r.choice([1, 2, 4]))
    kw["n"] = int(r.integers(20, 500))
    kw["size"] = float(r.choice([0.15, 0.5, 1.0, 2.5]))
    kw["seed"] = 100 + seed
    kw["cull"] = int(r.choice([A.CULL_NONE, A.CULL_FRONT, A.CULL_BACK]))
    kw["strip"] = bool(r.random() < 0.25)
    kw["index_dtype"] = np.uint32 if r.random() < 0.4 else np.uint16
    if not kw["strip"] and r.random() < 0.3:
        kw["base_vertex"] = int(r.choice([-23, -5, 11, 64]))
    elif r.random() < 0.2:
        kw["indexed"] = False
    if not kw["strip"]:
        kw["split"] = int(r.choice([1, 1, 2, 3]))
    kind = r.random()
    if kind < 0.35:  # depth only
        kw["ds"] = A.depth_stencil_desc(depth_enable=bool(r.random() < 0.9), depth_write=bool(r.random() < 0.8), depth_func=int(r.integers(0, 8)))
    elif kind < 0.7:  # stencil, two-sided
        ops = lambda: (int(r.integers(1, 9)), int(r.integers(1, 9)), int(r.integers(1, 9)), int(r.integers(0, 8)))  # noqa: E731
        kw["ds"] = A.depth_stencil_desc(depth_enable=bool(r.random() < 0.8), depth_write=bool(r.random() < 0.7), depth_func=int(r.integers(0, 8)),
                                        stencil_enable=True, read_mask=int(r.integers(0, 256)), write_mask=int(r.integers(0, 256)),
                                        front=ops(), back=ops())
        kw["stencil_ref"] = int(r.integers(0, 256))
    kw["bs"] = int(r.choice([A.BS_REPLACE, A.BS_REPLACE, A.BS_LERP_SRC_ALPHA, A.BS_REPLACE_AND_COUNT]))
    if kw["bs"] != A.BS_REPLACE_AND_COUNT:
        kw["color_fmt"] = int(r.choice([A.PF_RGBA8, A.PF_BGRA8, A.PF_RGBA32F] if kw["bs"] == A.BS_LERP_SRC_ALPHA else [A.PF_RGBA8, A.PF_BGRA8]))
    m = r.random()
    if m < 0.15:
        kw["modifiers"] = [A.AM_NOPERSPECTIVE]
    elif m < 0.3:
        kw["modifiers"] = [A.AM_CENTROID | A.AM_LINEAR]
    elif m < 0.4:
        kw["modifiers"] = [A.AM_NOINTERPOLATION]
    return kw, S.TriangleSoup(**kw)

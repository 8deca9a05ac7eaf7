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


# ANISO_NEEDS_GRADIENTS: an anisotropic mip filter is only defined for fetches that carry derivatives (sample_2d_grad: SASL tex2D,
# PS_TEX_GRAD_ALPHA, PS_SPONZA_GRAD).  The explicit-LOD entry a C++ pixel shader's tex2d uses passes a null anisotropic_info that
# the minification branch dereferences (sampler.cpp:768 -> sampler.cpp:731): the reference crashes on the first minified quad
# (SURVEY.md Appendix B #6), so the random scenes never pair such a program with an anisotropic sampler.
def scene_from_seed(seed: int):
    """Seeded random instances of the textured scenes: TextureAndBlending (both derivative paths, point / trilinear / anisotropic
    mip filters), the Sponza-like atrium (built-in and SASL-twin pixel shader, trilinear / anisotropic), the AnisotropicFilter
    sample, the shadow-map sample, vertex texture fetch - random target size (multiples of 4), sample count and frame."""
    r = np.random.default_rng(5000 + seed)
    w, h = int(r.integers(20, 100)) * 4, int(r.integers(14, 64)) * 4
    s = int(r.choice([1, 2, 4]))
    kind = int(r.integers(0, 6))
    if kind == 0:
        ps = int(r.choice([A.PS_TEX_ALPHA, A.PS_TEX_GRAD_ALPHA]))
        mip = int(r.choice([A.FILTER_POINT, A.FILTER_LINEAR, A.FILTER_ANISOTROPIC]))
        if ps == A.PS_TEX_ALPHA and mip == A.FILTER_ANISOTROPIC:
            mip = A.FILTER_LINEAR  # see ANISO_NEEDS_GRADIENTS
        sc = S.TextureAndBlending(w, h, s, ps_program=ps, mip_filter=mip,
                                  max_aniso=int(r.choice([2, 4, 8, 16])) if mip == A.FILTER_ANISOTROPIC else 0)
    elif kind == 1:
        ps = int(r.choice([A.PS_SPONZA, A.PS_SPONZA_GRAD]))
        aniso = int(r.choice([0, 0, 4, 16]))
        sc = S.SponzaLike(w, h, s, tex_size=int(r.choice([32, 64, 128])), max_aniso=aniso if ps == A.PS_SPONZA_GRAD else 0,
                          ps_program=ps)
    elif kind == 2:
        sc = S.AnisotropicFilter(w, h, s)
    elif kind == 3:
        sc = S.StandardShadowMap(w, h, int(r.choice([1, 4])), tex_size=int(r.choice([32, 64])))
    elif kind == 4:
        sc = S.TerrainVTF(w, h, s, block=int(r.choice([8, 16])), tex_size=int(r.choice([16, 32, 64])))
    else:
        sc = S.HeightFieldTwoPass(w, h, s, nx=int(r.integers(10, 60)), nz=int(r.integers(10, 50)), shadowed=bool(r.random() < 0.4) and s == 1)
    return sc, int(r.integers(0, sc.n_frames)), f"kind {kind} {w}x{h}x{s}"


def scene_tolerance(scene):
    """The shadow-map pixel shaders call expf / logf / pow: colour within 1 LSB there, bit-exact elsewhere."""
    return 1 if type(scene).__name__ == "StandardShadowMap" or getattr(scene, "shadowed", False) else 0


def viewport_soup_from_seed(seed: int):
    """soup_from_seed with a random viewport instead of the whole target (viewport.h:5-12): an integer or fractional
    sub-rectangle, a rectangle that reaches past the target, and depth ranges other than 0..1 - plus front_ccw, non-zero
    byte offsets of the vertex streams and the 1 / 2 / 3-float vertex formats.  Upstream sizes its tile grid from
    the viewport's WIDTH and HEIGHT but anchors it at the target's origin (rasterizer.cpp:1106), so with x / y > 0 the right /
    bottom part of the viewport falls outside the grid and is not drawn - mirrored, and what these scenes pin."""
    g = np.random.default_rng(7000 + seed)
    kw, _ = soup_from_seed(seed)
    w, h = kw["w"], kw["h"]
    kind = int(g.integers(0, 4))
    if kind == 0:
        x, y = int(g.integers(0, w // 2)), int(g.integers(0, h // 2))
        vw, vh = int(g.integers(8, w - x + 1)), int(g.integers(8, h - y + 1))
    elif kind == 1:
        x, y = float(g.uniform(0, w / 2)), float(g.uniform(0, h / 2))
        vw, vh = float(g.uniform(8, w - x)), float(g.uniform(8, h - y))
    elif kind == 2:
        x, y = float(g.uniform(0, w / 3)), float(g.uniform(0, h / 3))
        vw, vh = float(g.uniform(w / 2, 1.5 * w)), float(g.uniform(h / 2, 1.5 * h))
    else:
        x, y, vw, vh = 0, 0, w, h
    minz, maxz = [(0.0, 1.0), (0.2, 0.9), (0.5, 0.5), (0.0, 0.5)][int(g.integers(0, 4))]
    kw["viewport"] = (x, y, vw, vh, minz, maxz)
    # two more pieces of state no hand-picked scene varies: the winding that counts as front, byte offsets of the vertex streams
    kw["front_ccw"] = bool(g.random() < 0.5)
    if g.random() < 0.5:
        kw["stream_pad"] = (int(g.integers(0, 9)), int(g.integers(0, 9)))
    # the narrow vertex formats: get_vec4 pads with 0 and the element's default w - 1 for a position semantic, 0 for any other
    # (semantic_value::default_w, constants.h: the only two values the reference can produce)
    if g.random() < 0.5:
        kw["color_element"] = (int(g.choice([A.FMT_R32_FLOAT, A.FMT_R32G32_FLOAT, A.FMT_R32G32B32_FLOAT])), float(g.choice([0.0, 1.0])))
    return kw, S.TriangleSoup(**kw)

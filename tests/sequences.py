"""API call sequences around clears, viewports and frame boundaries (shared by the CPU and the GPU suite).  Every sequence is a
function of a backend; it returns the buffers an application could observe.  They exist because the product executes clears
lazily, fuses the resolve into shading and pipelines frames: each sequence hits one of the hand-over points."""
import numpy as np

from salviarenderer_b200 import abi as A, scenes as S


def _snap(be, t):
    out = [be.read_texture(t.color).copy(), be.read_texture(t.ds).copy()]
    if t.resolved is not None:
        out.append(be.read_texture(t.resolved).copy())
    return out


def _soup_draw(be, sc, vp=None, ds=None, prim_count=None):
    d = S.base_desc(sc.t, sc.w, sc.h, cull=A.CULL_NONE, ds=ds)
    if vp is not None:
        d.viewport.x, d.viewport.y, d.viewport.w, d.viewport.h = vp
    d.n_color_targets = 1  # colour only (no coverage counter target)
    sc.mesh.fill_desc(be, d, prim_count=prim_count)
    d.vs = A.shader_binding(A.VS_MVP_PASSTHROUGH, S.pack_vs_mvp_passthrough(np.eye(4, dtype=np.float32), [1]))
    d.ps = A.shader_binding(A.PS_ATTR0_COLOR)
    d.bs = A.shader_binding(A.BS_REPLACE)
    be.draw(d)


def _soup(samples, w=256, h=192, n=400, seed=21):
    sc = S.TriangleSoup(w=w, h=h, samples=samples, n=n, seed=seed, bs=A.BS_REPLACE)
    return sc


def seq_viewport_smaller_than_target(be, samples=4):
    """The tile grid comes from the viewport: pixels outside it keep the clear values."""
    sc = _soup(samples)
    sc.setup(be)
    t = sc.t
    be.clear_color(t.color, (0.1, 0.6, 0.3, 1.0))
    be.clear_depth_stencil(t.ds, A.CLEAR_DEPTH | A.CLEAR_STENCIL, 1.0, 7)
    _soup_draw(be, sc, vp=(0, 0, 128, 64))
    if t.resolved is not None:
        be.resolve(t.color, t.resolved)
    return _snap(be, t)


def seq_colour_clear_only_between_frames(be, samples=4):
    """Frame 2 clears colour but keeps frame 1's depth: only triangles in front of frame 1's show up."""
    sc = _soup(samples)
    sc.setup(be)
    t = sc.t
    be.clear_color(t.color, (0.0, 0.0, 0.0, 1.0))
    be.clear_depth_stencil(t.ds, A.CLEAR_DEPTH | A.CLEAR_STENCIL, 1.0, 0)
    _soup_draw(be, sc, prim_count=200)
    be.clear_color(t.color, (0.9, 0.1, 0.1, 1.0))
    _soup_draw(be, sc)
    if t.resolved is not None:
        be.resolve(t.color, t.resolved)
    return _snap(be, t)


def seq_partial_depth_stencil_clears(be, samples=2):
    """clear_depth_stencil with a single flag after a whole clear; a second whole clear overrides the first."""
    sc = _soup(samples)
    sc.setup(be)
    t = sc.t
    be.clear_color(t.color, (0.3, 0.3, 0.3, 1.0))
    be.clear_color(t.color, (0.2, 0.4, 0.6, 1.0))
    be.clear_depth_stencil(t.ds, A.CLEAR_DEPTH | A.CLEAR_STENCIL, 0.25, 3)
    be.clear_depth_stencil(t.ds, A.CLEAR_DEPTH, 1.0, 0)       # stencil 3 stays
    _soup_draw(be, sc)
    be.clear_depth_stencil(t.ds, A.CLEAR_STENCIL, 0.0, 9)     # depth of the draw stays
    if t.resolved is not None:
        be.resolve(t.color, t.resolved)
    return _snap(be, t)


def seq_clear_without_draw_then_resolve(be, samples=4):
    """A clear that no draw consumes must still be what resolve and readback see."""
    sc = _soup(samples)
    sc.setup(be)
    t = sc.t
    be.clear_color(t.color, (0.25, 0.5, 0.75, 1.0))
    be.clear_depth_stencil(t.ds, A.CLEAR_DEPTH | A.CLEAR_STENCIL, 0.5, 1)
    be.resolve(t.color, t.resolved)
    return _snap(be, t)


def seq_resolve_twice_and_draw_after_resolve(be, samples=4):
    """Resolve in the middle of a frame, more draws, resolve again (the second resolve sees both batches)."""
    sc = _soup(samples)
    sc.setup(be)
    t = sc.t
    be.clear_color(t.color, (0.0, 0.2, 0.0, 1.0))
    be.clear_depth_stencil(t.ds, A.CLEAR_DEPTH | A.CLEAR_STENCIL, 1.0, 0)
    _soup_draw(be, sc, prim_count=150)
    be.resolve(t.color, t.resolved)
    first = be.read_texture(t.resolved).copy()
    _soup_draw(be, sc, ds=A.depth_stencil_desc(depth_func=A.CMP_ALWAYS))
    be.resolve(t.color, t.resolved)
    return _snap(be, t) + [first]


def seq_many_frames_back_to_back(be, samples=4):
    """Six frames without any readback in between (the product pipelines them over two scratch sets), then everything is read."""
    sc = S.SponzaLike(320, 192, samples, tex_size=32)
    sc.setup(be)
    for f in (0, 3, 5, 1, 6, 2):
        sc.render(be, f)
    return _snap(be, sc.t)


SEQUENCES = {f.__name__[4:]: f for f in (seq_viewport_smaller_than_target, seq_colour_clear_only_between_frames,
                                         seq_partial_depth_stencil_clears, seq_clear_without_draw_then_resolve,
                                         seq_resolve_twice_and_draw_after_resolve, seq_many_frames_back_to_back)}

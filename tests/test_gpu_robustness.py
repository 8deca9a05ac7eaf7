"""GPU suite (-m gpu): failure modes the advisor's review of round 1 named.

* work-list arenas: a batch of many full-screen triangles (every tile fully covered: 128 region-list words per tile-list
  entry) renders correctly with the default arenas, and a device whose arenas are too small reports SLV_OUT_OF_MEMORY for
  that frame ONCE, grows them to the recorded need and renders the re-issued frame correctly (no sticky failure);
* fused resolve: a batch that SAMPLES the resolve destination (temporal feedback) must not have the resolve fused into its
  shading kernel;
* more than two colour targets are rejected instead of being silently ignored; rejected draws leave the statistics alone.
"""
import os

import numpy as np
import pytest

import cases
from salviarenderer_b200 import abi as A, scenes as S

pytestmark = pytest.mark.gpu


@pytest.mark.timeout(900)
def test_many_fullscreen_triangles_4k(cuda, reference):
    """200 full-screen triangles in 4 draws of one batch at 3840x2160 (2040 tiles): 400 k fully covered (tile, triangle) pairs,
    three times what round 1's region arena held.  Against the unmodified reference."""
    a, b = S.OverlayQuads(3840, 2160, 1, n=100), S.OverlayQuads(3840, 2160, 1, n=100)
    a.setup(cuda)
    b.setup(reference)
    msgs = cases.compare_frames(a.run(cuda), b.run(reference))
    assert not msgs, msgs


def test_arena_overflow_is_reported_once_and_recovered(built, oracle):
    import salviarenderer_b200 as pkg
    os.environ["SLV_ARENA_MIN"] = "4096"
    try:
        small = pkg.load(0)
    finally:
        del os.environ["SLV_ARENA_MIN"]
    a, b = S.OverlayQuads(1280, 720, 4, n=40), S.OverlayQuads(1280, 720, 4, n=40)
    a.setup(small)
    b.setup(oracle)
    want = b.run(oracle)
    with pytest.raises(A.SlvError):      # 240 tiles x 80 triangles = 19 k list entries in a 4 k-entry arena
        a.run(small)
    for attempt in range(3):              # list arena first, then the region arena: each overflow is reported once
        try:
            got = a.run(small)
            break
        except A.SlvError:
            continue
    else:
        pytest.fail("the device did not recover from the arena overflow")
    assert not cases.compare_frames(got, want)
    assert not cases.compare_frames(a.run(small), want)   # and stays healthy
    small.close()


def _feedback_frames(be, n_frames=3, w=640, h=360):
    """Frame k draws a textured plane that samples the RESOLVED image of frame k - 1 and then resolves into that same texture."""
    t = S.create_targets(be, w, h, 4, A.PF_RGBA8)
    plane = S.create_planar((-3.0, -1.0, -3.0), (6, 0, 0), (0, 0, 6), 1, 1, True)
    plane.elements = [(0, S._V4, 0, 0, 1.0)]
    plane.upload(be)
    rng = np.random.default_rng(5)
    be.upload_texture(t.resolved, rng.integers(0, 256, size=(h, w, 4), dtype=np.uint8))
    samp = be.create_sampler(A.sampler_desc(A.FILTER_LINEAR, A.FILTER_LINEAR, A.FILTER_POINT, addr_u=A.ADDR_WRAP, addr_v=A.ADDR_WRAP), t.resolved)
    outs = []
    for k in range(n_frames):
        be.clear_color(t.color, (0.2, 0.2, 0.5, 1.0))
        be.clear_depth_stencil(t.ds, A.CLEAR_DEPTH | A.CLEAR_STENCIL, 1.0, 0)
        ang = -0.4 * (k + 1)
        view = S.mat_lookat((np.cos(ang) * 1.5, 1.5, np.sin(ang) * 1.5), (0, 0, 0), (0, 1, 0))
        proj = S.mat_perspective_fov(np.pi / 2, np.float32(w) / np.float32(h), 0.1, 100.0)
        wvp = S.mat_mul(S.mat_translate(-0.5, 0, -0.5), S.mat_mul(view, proj))
        d = S.base_desc(t, w, h, cull=A.CULL_BACK)
        plane.fill_desc(be, d)
        d.vs = A.shader_binding(A.VS_PLANE_XZ, S.pack_vs_plane_xz(wvp))
        d.ps = A.shader_binding(A.PS_TEX_ALPHA, S.pack_ps_tex_alpha(0, 1.0), [samp])
        d.bs = A.shader_binding(A.BS_REPLACE)
        be.draw(d)
        be.resolve(t.color, t.resolved)
        outs.append(be.read_texture(t.resolved).copy())
    return outs


def test_resolve_into_a_texture_the_batch_samples(cuda, oracle):
    got, want = _feedback_frames(cuda), _feedback_frames(oracle)
    for k, (g, w_) in enumerate(zip(got, want)):
        assert np.array_equal(g, w_), f"frame {k}: {(g != w_).sum()} bytes differ"


def test_third_colour_target_is_rejected_and_statistics_untouched(cuda):
    sc = S.TriangleSoup(samples=1, n=20, seed=2, bs=A.BS_REPLACE)
    sc.setup(cuda)
    extra = cuda.create_texture(sc.w, sc.h, 1, A.PF_RGBA8)
    cuda.query_begin()
    d = S.base_desc(sc.t, sc.w, sc.h, cull=A.CULL_NONE)
    sc.mesh.fill_desc(cuda, d)
    d.vs = A.shader_binding(A.VS_MVP_PASSTHROUGH, S.pack_vs_mvp_passthrough(S.mat_identity(), [1]))
    d.ps = A.shader_binding(A.PS_ATTR0_COLOR)
    d.bs = A.shader_binding(A.BS_REPLACE)
    d.n_color_targets = 3
    d.color_targets[1] = 0
    d.color_targets[2] = extra.handle
    with pytest.raises(A.SlvError):
        cuda.draw(d)
    d.n_color_targets = 1
    d.bs = A.shader_binding(99)           # an unknown blend program: rejected as well
    with pytest.raises(A.SlvError):
        cuda.draw(d)
    st = cuda.query_get()
    assert st["ia_primitives"] == 0 and st["ia_vertices"] == 0 and st["cinvocations"] == 0


def test_texture_level_tracking(cuda):
    """slv_texture_level_tracking / slv_texture_levels_touched (the B_tex accounting of bench.py): the mask holds exactly the
    mip levels the frame's sampler calls read; a small target minifies 256^2 textures, so the finest level stays untouched
    there while a large target reads it; switching tracking on again resets the masks; off = nothing recorded."""
    small = S.SponzaLike(160, 90, 1, tex_size=256)
    small.setup(cuda)
    cuda.texture_level_tracking(True)
    small.render(cuda, 3)
    cuda.flush()
    masks = [cuda.texture_levels_touched(t) for t in small.textures]
    n_levels = cuda.level_count(small.textures[0])
    assert any(masks) and all(m < (1 << n_levels) for m in masks)
    assert all(not (m & 1) for m in masks), masks  # 256^2 texels over a 160x90 frame: level 0 is never selected
    cuda.texture_level_tracking(True)  # reset
    assert all(cuda.texture_levels_touched(t) == 0 for t in small.textures)
    cuda.texture_level_tracking(False)
    small.render(cuda, 3)
    cuda.flush()
    assert all(cuda.texture_levels_touched(t) == 0 for t in small.textures)
    big = S.SponzaLike(1920, 1080, 1, tex_size=64)
    big.setup(cuda)
    cuda.texture_level_tracking(True)
    big.render(cuda, 3)
    cuda.flush()
    assert any(cuda.texture_levels_touched(t) & 1 for t in big.textures)  # magnified somewhere: level 0 is read
    cuda.texture_level_tracking(False)

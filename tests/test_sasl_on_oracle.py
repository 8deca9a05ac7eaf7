"""CPU suite: SASL shaders INSIDE the pipeline, without a GPU.  The CPU checker's slv_shader_compile builds the code the front
end generates - the very text the product hands to NVRTC - for the host (oracle/slv_host_shader.h: the four pixels of a quad
run as fibers that meet at every ddx / ddy, texture fetches go to the checker's own sampler), so SASL vertex and pixel shaders
run in the restated pipeline and whole frames can be compared with the cpp twins the samples ship - on the restatement AND on
the live unmodified reference.  tests/test_gpu_sasl_jit.py makes the same comparisons for the sm_100a build of that code."""
import numpy as np
import pytest

import cases
from salviarenderer_b200 import abi as A, scenes as S
from salviarenderer_b200.sasl import jit
from test_gpu_sasl_jit import PS_TEX_ALPHA, VS_SPONZA, VS_TERRAIN


def host_only(source, stage, **kw):
    return jit.compile(source, stage, device=False, **kw)


def test_sasl_sponza_vertex_shader_equals_builtin_and_reference(oracle, reference):
    sh = host_only(VS_SPONZA, "vs")
    mod = jit.load(oracle, sh)
    mk = lambda: S.SponzaLike(240, 136, 4, tex_size=32)  # noqa: E731
    ref, twin, got = mk(), mk(), mk()
    ref.setup(reference)
    twin.setup(oracle)
    got.setup(oracle)
    got.vs_binding = lambda wvp, light, eye: A.shader_binding(
        A.program_jit(mod), sh.unit.pack_uniforms({"wvpMatrix": np.asarray(wvp, np.float32).reshape(4, 4), "lightPos": light, "eyePos": eye}))
    for f in (0, 5):
        r, a, b = ref.run(reference, f), twin.run(oracle, f), got.run(oracle, f)
        assert cases.compare_frames(a, b) == [], f"frame {f}: SASL vs built-in twin"
        assert cases.compare_frames(r, b) == [], f"frame {f}: SASL on the restatement vs the twin on the live reference"
        assert a.stats["cprimitives"] > 1000


@pytest.mark.parametrize("mip_filter,aniso,samples", [(A.FILTER_LINEAR, 0, 1), (A.FILTER_ANISOTROPIC, 16, 4)])
def test_sasl_tex2d_pixel_shader_equals_builtin_grad_path(oracle, reference, mip_filter, aniso, samples):
    """tex2D + constant alpha with blending (TextureAndBlending): the plane draw reads attribute 0, the box draws attribute 1."""
    sh = [host_only(PS_TEX_ALPHA.format(decls=d), "ps", derivatives="cpp") for d in ("float4 uv: TEXCOORD0;", "float4 pad: TEXCOORD0; float4 uv: TEXCOORD1;")]
    mods = [jit.load(oracle, s) for s in sh]
    mk = lambda: S.TextureAndBlending(320, 180, samples=samples, ps_program=A.PS_TEX_GRAD_ALPHA, mip_filter=mip_filter, max_aniso=aniso)  # noqa: E731
    ref, twin, got = mk(), mk(), mk()
    ref.setup(reference)
    twin.setup(oracle)
    got.setup(oracle)
    got.ps_binding = lambda reg, alpha, samp: A.shader_binding(A.program_jit(mods[reg]), sh[reg].unit.pack_uniforms({"alpha": alpha}), [samp])
    for f in (0, 2):
        r, a, b = ref.run(reference, f), twin.run(oracle, f), got.run(oracle, f)
        assert cases.compare_frames(a, b) == [], f"frame {f}"
        assert cases.compare_frames(r, b) == [], f"frame {f}: against the live reference"


def install_pair(sc, be):
    """bench.install_sasl_shaders for a CPU checker (no sm_100a image is compiled)."""
    import bench
    vsh, psh = host_only(bench.SASL_VS_SPONZA, "vs"), host_only(bench.SASL_PS_SPONZA, "ps")
    vs_mod, ps_mod = jit.load(be, vsh), jit.load(be, psh)
    sc.vs_binding = lambda wvp, light, eye: A.shader_binding(A.program_jit(vs_mod), vsh.unit.pack_uniforms(
        {"wvpMatrix": np.asarray(wvp, np.float32).reshape(4, 4), "lightPos": light, "eyePos": eye}))
    base = sc.frame_draws

    def frame_draws(be_, frame):
        ds = base(be_, frame)
        for d, (m, _, _) in zip(ds, sc.groups):
            d.ps = A.shader_binding(A.program_jit(ps_mod), b"", [sc.samplers[m]])
        return ds

    sc.frame_draws = frame_draws


@pytest.mark.parametrize("w,h,samples,aniso,frames", [(240, 136, 4, 16, (1, 6)), (320, 180, 1, 0, (3,))])
def test_headline_sasl_pair_equals_twins_on_the_reference(oracle, reference, w, h, samples, aniso, frames):
    """bench.py's headline shaders - the SASL Sponza vertex + pixel shader, tex2D with per-line / per-column derivatives, 16x
    anisotropic - run in the restated pipeline against SLV_VS_SPONZA + SLV_PS_SPONZA_GRAD on the LIVE REFERENCE: every buffer
    and counter.  The GPU suite pins the sm_100a build of the same generated code to the same twins."""
    got = S.SponzaLike(w, h, samples, tex_size=32, max_aniso=aniso)
    got.setup(oracle)
    install_pair(got, oracle)
    ref = S.SponzaLike(w, h, samples, tex_size=32, max_aniso=aniso, ps_program=A.PS_SPONZA_GRAD, sasl_derivatives=True)
    ref.setup(reference)
    for f in frames:
        assert cases.compare_frames(ref.run(reference, f), got.run(oracle, f)) == [], f"frame {f}"


def test_sasl_vertex_texture_fetch_equals_builtin(oracle, reference):
    sh = host_only(VS_TERRAIN, "vs")
    mod = jit.load(oracle, sh)
    ref = S.TerrainVTF(320, 180, 1)
    ref.setup(reference)
    got = S.TerrainVTF(320, 180, 1, vs_binding=lambda wvp, off, scale, samp: A.shader_binding(
        A.program_jit(mod), sh.unit.pack_uniforms({"wvpMatrix": np.asarray(wvp, np.float32).reshape(4, 4), "terrainOffset": off, "terrainScale": scale}), [samp]))
    got.setup(oracle)
    for f in (0, 3):
        a, b = ref.run(reference, f), got.run(oracle, f)
        assert cases.compare_frames(a, b) == [], f"frame {f}"
        assert a.stats["cprimitives"] == 8192


def test_compile_errors_and_release(oracle):
    with pytest.raises(A.SlvError, match="slv_shader_compile failed"):
        oracle.shader_compile("ps", "this is not C++")
    sh = host_only("float4 main(float4 p: TEXCOORD0): COLOR { return p; }", "ps")
    mod = jit.load(oracle, sh)
    assert mod > 0
    oracle.release(mod) if hasattr(oracle, "release") else None


# ---- the remaining tests of the GPU suite's SASL module, run on the restatement -------------------------------------------------
@pytest.fixture
def host_jit(monkeypatch):
    """tests/test_gpu_sasl_jit.py's test bodies with jit.compile producing no sm_100a image (nothing here loads one)."""
    import test_gpu_sasl_jit as G
    real = jit.compile
    monkeypatch.setattr(G.jit, "compile", lambda source, stage, entry=None, derivatives="sasl", **kw: real(source, stage, entry, derivatives, device=False))
    return G


def test_known_answers_through_the_pipeline(oracle, host_jit):
    """The reference's known answers for the intrinsics (tests/golden/sasl_kat.json), each through a draw into an rgba32f target."""
    host_jit.test_sasl_intrinsics_match_the_reference_known_answers(oracle)


def test_derivative_conventions_through_the_pipeline(oracle, host_jit):
    host_jit.test_sasl_derivative_convention(oracle)


def test_skinning_vertex_shader_array_uniforms(oracle, host_jit):
    """samples/AstroBoy's skinning vertex shader: array uniforms (addresses in the uniform block), int4 inputs, a data-dependent break."""
    host_jit.test_sasl_skinning_vertex_shader_array_uniforms(oracle)


@pytest.mark.parametrize("w,h,samples", [(320, 180, 1), (200, 120, 4)])
def test_two_sampler_shadow_map_shader(oracle, host_jit, w, h, samples):
    """The StandardShadowMap colour pass in SASL (two samplers, nine tex2Dlod taps, exp / log / pow) against SLV_PS_SSM_DRAW:
    everything but the colour identical, the colour within a few LSB (SASL's log is eflib's fast_log polynomial)."""
    host_jit.test_sasl_two_sampler_shadow_map_shader(oracle, oracle, w, h, samples)

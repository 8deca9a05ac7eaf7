"""GPU suite: SASL shaders compiled at run time (salviarenderer_b200/sasl: front end -> CUDA toolchain -> cubin ->
slv_shader_module_load) against the built-in device programs that are themselves pinned to the reference.

* the Sponza vertex shader written in SASL (the sample ships the same shader as SASL text and as a cpp twin,
  samples/Sponza/Sponza.cpp:39-97) must give bit-identical frames and counters to SLV_VS_SPONZA;
* `tex2D` + constant alpha written in SASL (TextureAndBlending.cpp:79-94) with the cpp derivative convention must equal
  SLV_PS_TEX_GRAD_ALPHA (sample_2d_grad with the quad's derivatives), trilinear and 16x anisotropic, with blending;
* the SASL derivative convention (per row / per column) is checked through ddx / ddy of a linear function.
"""
import numpy as np
import pytest

import cases
from salviarenderer_b200 import abi as A, scenes as S
from salviarenderer_b200.sasl import jit

pytestmark = pytest.mark.gpu

VS_SPONZA = """
float4x4 wvpMatrix;
float4   lightPos;
float4   eyePos;
struct VSIn  { float4 pos: POSITION; float4 tex: TEXCOORD0; float4 norm: NORMAL; };
struct VSOut { float4 pos: sv_position; float4 tex: TEXCOORD0; float4 norm: TEXCOORD1; float4 lightDir: TEXCOORD2; float4 eyeDir: TEXCOORD3; };
VSOut vs_main(VSIn in) {
    VSOut o;
    o.norm = in.norm;
    o.pos = mul(in.pos, wvpMatrix);
    o.lightDir = lightPos - in.pos;
    o.eyeDir = eyePos - in.pos;
    o.tex = in.tex;
    return o;
}
"""

PS_TEX_ALPHA = """
sampler texSamp;
float   alpha;
struct PSIn {{ {decls} }};
float4 ps_main(PSIn in): COLOR {{
    float4 c = tex2D(texSamp, in.uv.xy);
    c.w = alpha;
    return c;
}}
"""


@pytest.mark.parametrize("samples,size", [(4, (480, 272)), (1, (640, 360))])
def test_sasl_sponza_vertex_shader_equals_builtin(cuda, samples, size):
    sh = jit.compile(VS_SPONZA, "vs")
    mod = jit.load(cuda, sh)
    w, h = size
    ref = S.SponzaLike(w, h, samples, tex_size=128)
    ref.setup(cuda)
    got = S.SponzaLike(w, h, samples, tex_size=128)
    got.setup(cuda)
    got.vs_binding = lambda wvp, light, eye: A.shader_binding(
        A.program_jit(mod), sh.unit.pack_uniforms({"wvpMatrix": np.asarray(wvp, np.float32).reshape(4, 4), "lightPos": light, "eyePos": eye}))
    for f in (0, 5):
        a, b = ref.run(cuda, f), got.run(cuda, f)
        assert cases.compare_frames(a, b) == [], f"frame {f}"
        assert a.stats["cprimitives"] > 1000


@pytest.mark.parametrize("mip_filter,aniso,samples", [(A.FILTER_LINEAR, 0, 1), (A.FILTER_ANISOTROPIC, 16, 4)])
def test_sasl_tex2d_pixel_shader_equals_builtin_grad_path(cuda, mip_filter, aniso, samples):
    # the plane draw reads attribute 0, the box draws attribute 1: two shaders (a SASL pixel shader's k-th input is attribute k)
    sh = [jit.compile(PS_TEX_ALPHA.format(decls=d), "ps", derivatives="cpp")
          for d in ("float4 uv: TEXCOORD0;", "float4 pad: TEXCOORD0; float4 uv: TEXCOORD1;")]
    mods = [jit.load(cuda, s) for s in sh]
    ref = S.TextureAndBlending(640, 360, samples=samples, ps_program=A.PS_TEX_GRAD_ALPHA, mip_filter=mip_filter, max_aniso=aniso)
    ref.setup(cuda)
    got = S.TextureAndBlending(640, 360, samples=samples, ps_program=A.PS_TEX_GRAD_ALPHA, mip_filter=mip_filter, max_aniso=aniso)
    got.setup(cuda)
    got.ps_binding = lambda reg, alpha, samp: A.shader_binding(A.program_jit(mods[reg]), sh[reg].unit.pack_uniforms({"alpha": alpha}), [samp])
    for f in (0, 2):
        a, b = ref.run(cuda, f), got.run(cuda, f)
        assert cases.compare_frames(a, b) == [], f"frame {f}"


def test_sasl_derivative_convention(cuda):
    """ddx / ddy of attribute values: the SASL convention differences pixel pairs per row / per column, the cpp one uses
    q1 - q0 / q2 - q0 for the whole quad; for a noperspective linear attribute both must give the constant gradient, and for
    a perspective one they must agree exactly on pixel 0 of every quad."""
    src = """
    struct PSIn { float4 c: TEXCOORD0; };
    float4 ps_main(PSIn in): COLOR { return float4(ddx(in.c.x) * 8.0f + 0.5f, ddy(in.c.z) * 8.0f + 0.5f, in.c.z, 1.0f); }
    """
    out = {}
    for conv in ("sasl", "cpp"):
        sh = jit.compile(src, "ps", derivatives=conv)
        mod = jit.load(cuda, sh)
        t = S.create_targets(cuda, 256, 256, 1, A.PF_RGBA8)
        # one large triangle pair with attribute = position (varies linearly on screen)
        mesh = S.create_planar((-3.0, -1.0, -3.0), (6, 0, 0), (0, 0, 6), 1, 1, True)
        mesh.elements = [(0, S._V4, 0, 0, 1.0)]
        mesh.upload(cuda)
        cuda.clear_color(t.color, (0, 0, 0, 0))
        cuda.clear_depth_stencil(t.ds, A.CLEAR_DEPTH | A.CLEAR_STENCIL, 1.0, 0)
        view = S.mat_lookat((0.4, 2.5, 0.7), (0, 0, 0), (0, 1, 0))
        proj = S.mat_perspective_fov(np.pi / 2, 1.0, 0.1, 100.0)
        wvp = S.mat_mul(view, proj)
        d = S.base_desc(t, 256, 256, cull=A.CULL_NONE)
        mesh.fill_desc(cuda, d, prim_count=1)  # ONE triangle: every fully covered quad lies inside it
        d.vs = A.shader_binding(A.VS_MVP_PASSTHROUGH, S.pack_vs_mvp_passthrough(wvp, [0]))
        d.ps = A.shader_binding(A.program_jit(mod))
        d.bs = A.shader_binding(A.BS_REPLACE)
        cuda.draw(d)
        out[conv] = cuda.read_texture(t.color).reshape(256, 256, 4).astype(np.int32)
    a, b = out["sasl"], out["cpp"]
    covered = (a[..., 3] == 255) & (b[..., 3] == 255)
    assert covered.sum() > 5000
    # pixel 0 of every quad (even x, even y): identical in both conventions
    q0 = covered[0::2, 0::2]
    assert np.array_equal(a[0::2, 0::2][q0], b[0::2, 0::2][q0])
    # fully covered quads: the cpp convention gives all four pixels pixel 0's derivatives, the SASL one does not
    full = covered[0::2, 0::2] & covered[0::2, 1::2] & covered[1::2, 0::2] & covered[1::2, 1::2]
    assert full.sum() > 1000
    for dy, dx in ((0, 1), (1, 0), (1, 1)):
        assert np.array_equal(b[dy::2, dx::2, :2][full], b[0::2, 0::2, :2][full])
    assert np.any(a[1::2, 1::2, :2][full] != a[0::2, 0::2, :2][full])


VS_TERRAIN = """
float4x4 wvpMatrix;
float2   terrainOffset;
float2   terrainScale;
sampler  terrainSamp;
struct VSIn  { float4 pos: POSITION; float4 uv: TEXCOORD0; };
struct VSOut { float4 pos: sv_position; float displacement: TEXCOORD0; };
VSOut vs_main(VSIn in) {
    VSOut o;
    float2 terrainUV = terrainOffset + in.uv.xy * terrainScale;
    float displacement = tex2Dlod(terrainSamp, float4(terrainUV, 0.0f, 0.0f)).x;
    float4 displaced_pos = float4(in.pos.xyz + float3(0.0f, displacement * 20.0f, 0.0f), 1.0f);
    o.pos = mul(displaced_pos, wvpMatrix);
    o.displacement = displacement;
    return o;
}
"""


def test_sasl_vertex_texture_fetch_equals_builtin(cuda):
    """The VertexTextureFetch sample's vertex shader (tex2Dlod in the VERTEX stage) in SASL against SLV_VS_TERRAIN_VTF,
    which is pinned to the reference's sampler::sample_2d_lod through the golden fixtures."""
    sh = jit.compile(VS_TERRAIN, "vs")
    assert sh.reflection.samplers == ["terrainSamp"] and sh.reflection.n_vs_output_attrs == 1
    mod = jit.load(cuda, sh)
    ref = S.TerrainVTF(640, 360, 1)
    ref.setup(cuda)
    got = S.TerrainVTF(640, 360, 1, vs_binding=lambda wvp, off, scale, samp: A.shader_binding(
        A.program_jit(mod), sh.unit.pack_uniforms({"wvpMatrix": np.asarray(wvp, np.float32).reshape(4, 4), "terrainOffset": off,
                                                   "terrainScale": scale}), [samp]))
    got.setup(cuda)
    for f in (0, 3):
        a, b = ref.run(cuda, f), got.run(cuda, f)
        assert cases.compare_frames(a, b) == [], f"frame {f}"
        assert a.stats["cprimitives"] == 8192


# ---- SASL pixel shaders on the visibility-first path (the module's quad-granular k_shade) --------------------------------
PS_SPONZA = """
sampler texSamp;
struct PSIn { float4 tex: TEXCOORD0; float4 norm: TEXCOORD1; float4 lightDir: TEXCOORD2; float4 eyeDir: TEXCOORD3; };
float4 ps_main(PSIn in): COLOR {
    float4 diff = tex2D(texSamp, in.tex.xy);
    float illum = clamp(dot(normalize(in.lightDir.xyz), normalize(in.norm.xyz)), 0.0f, 1.0f);
    return float4(diff.xyz * illum, 1.0f);
}
"""


@pytest.fixture(scope="module")
def cuda_jit_immediate(built):
    """A second device of the product on which SASL pixel shaders always take the immediate k_raster path."""
    import os
    import salviarenderer_b200 as pkg
    os.environ["SLV_JIT_IMMEDIATE"] = "1"
    try:
        be = pkg.load(0)
    finally:
        del os.environ["SLV_JIT_IMMEDIATE"]
    return be


def _sponza_with_ps(be, sh, w, h, samples, tex_size, max_aniso, uniforms=None):
    mod = jit.load(be, sh)
    sc = S.SponzaLike(w, h, samples, tex_size=tex_size, max_aniso=max_aniso)
    sc.setup(be)
    base = sc.frame_draws

    def frame_draws(be_, frame):
        ds = base(be_, frame)
        for d, (m, _, _) in zip(ds, sc.groups):
            d.ps = A.shader_binding(A.program_jit(mod), sh.unit.pack_uniforms(uniforms or {}), [sc.samplers[m]])
        return ds

    sc.frame_draws = frame_draws
    return sc


@pytest.mark.parametrize("w,h,samples,aniso,conv,frames", [
    (480, 272, 4, 0, "sasl", (0, 5)),
    (960, 540, 1, 0, "sasl", (3,)),
    (640, 360, 2, 16, "sasl", (7,)),
    (1000, 600, 4, 16, "cpp", (2,)),     # target size not a multiple of the tile / quad size
])
def test_sasl_pixel_shader_visibility_first_equals_immediate(cuda, cuda_jit_immediate, w, h, samples, aniso, conv, frames):
    """A SASL pixel shader (lighting + tex2D with implicit gradients) over the Sponza-like scene: the visibility-first path
    (k_cover + the module's quad-granular k_shade, fused resolve, lazy clears) against the immediate k_raster path, which
    test_sasl_tex2d_pixel_shader_equals_builtin_grad_path pins to the built-in program.  Every buffer and counter identical;
    the visibility-first path must have run fewer shader lanes (it shades a quad once per DISTINCT final owner)."""
    sh = jit.compile(PS_SPONZA, "ps", derivatives=conv)
    a = _sponza_with_ps(cuda, sh, w, h, samples, 128, aniso)
    b = _sponza_with_ps(cuda_jit_immediate, sh, w, h, samples, 128, aniso)
    for f in frames:
        ra, rb = a.run(cuda, f), b.run(cuda_jit_immediate, f)
        assert cases.compare_frames(ra, rb) == [], f"frame {f}"
        ta, tb = cuda.traffic(), cuda_jit_immediate.traffic()
        for k in ("z_tested", "z_written", "c_written", "c_read"):
            assert ta[k] == tb[k], f"frame {f}: traffic counter {k}"
        assert ra.stats["ps_invocations"] > 1000
        assert tb["ps_executed"] == rb.stats["ps_invocations"]  # the immediate path shades every quad it counts
        assert 0 < ta["ps_executed"] < tb["ps_executed"], "the SASL shader did not take the visibility-first path"


def test_sasl_pixel_shader_visibility_first_equals_builtin_twin(cuda):
    """tex2D + constant alpha in SASL (cpp derivative convention) on every draw of the Sponza-like scene against the built-in
    SLV_PS_TEX_GRAD_ALPHA: both take the visibility-first path (REPLACE blend), one through the pixel-granular k_shade, the
    other through the quad-granular one; 16x anisotropic samplers, 4x MSAA."""
    sh = jit.compile(PS_TEX_ALPHA.format(decls="float4 uv: TEXCOORD0;"), "ps", derivatives="cpp")
    got = _sponza_with_ps(cuda, sh, 800, 448, 4, 128, 16, uniforms={"alpha": 0.75})
    ref = S.SponzaLike(800, 448, 4, tex_size=128, max_aniso=16)
    ref.setup(cuda)
    base = ref.frame_draws

    def frame_draws(be_, frame):
        ds = base(be_, frame)
        for d, (m, _, _) in zip(ds, ref.groups):
            d.ps = A.shader_binding(A.PS_TEX_GRAD_ALPHA, S.pack_ps_tex_alpha(0, 0.75), [ref.samplers[m]])
        return ds

    ref.frame_draws = frame_draws
    for f in (1, 6):
        ra, rb = ref.run(cuda, f), got.run(cuda, f)
        assert cases.compare_frames(ra, rb) == [], f"frame {f}"
        assert ra.stats["ps_invocations"] > 1000


@pytest.mark.parametrize("w,h,samples,aniso,frames", [(800, 448, 4, 16, (1, 6)), (960, 540, 1, 0, (3,))])
def test_sasl_sponza_pair_equals_builtin_twins(cuda, w, h, samples, aniso, frames):
    """bench.py's headline shaders - the SASL Sponza vertex + pixel shader (tex2D with the SASL per-row / per-column
    derivatives) - against their built-in twins SLV_VS_SPONZA + SLV_PS_SPONZA_GRAD, which tests/test_gpu_parity.py pins to the
    unmodified reference at full size: quad-granular k_shade of the JIT module against the pixel-granular built-in one."""
    import bench
    got = S.SponzaLike(w, h, samples, tex_size=128, max_aniso=aniso)
    bench.install_sasl_shaders(got, cuda, A)
    got.setup(cuda)
    ref = S.SponzaLike(w, h, samples, tex_size=128, max_aniso=aniso, ps_program=A.PS_SPONZA_GRAD, sasl_derivatives=True)
    ref.setup(cuda)
    for f in frames:
        ra, rb = ref.run(cuda, f), got.run(cuda, f)
        assert cases.compare_frames(ra, rb) == [], f"frame {f}"
        assert ra.stats["ps_invocations"] > 1000


# ---- two samplers in one SASL pixel shader: the StandardShadowMap colour pass ---------------------------------------------
PS_SSM_DRAW = """
sampler texSamp;
sampler smSamp;
float4 ambient;
float4 diffuse;
float4 specular;
float  shininess;
struct PSIn { float4 tex: TEXCOORD0; float4 norm: TEXCOORD1; float4 lightDir: TEXCOORD2; float4 eyeDir: TEXCOORD3; float4 lsp: TEXCOORD4; };
float tap(float x, float y) { return tex2Dlod(smSamp, float4(x, y, 0.0f, 0.0f)).x; }
float4 ps_main(PSIn in): COLOR {
    float esm = 25000.0f;
    float lx = in.lsp.x / in.lsp.w;
    float ly = in.lsp.y / in.lsp.w;
    float lz = in.lsp.z / in.lsp.w;
    float cx = (lx + 1.0f) * 0.5f;
    float cy = 1.0f - (ly + 1.0f) * 0.5f;
    float off = 1.0f / 512.0f;
    float sd0 = tap(cx + -off, cy + -off);
    float occluder = 0.0f;
    occluder += 0.111014f * exp(esm * (tap(cx + 0.0f, cy + -off) - sd0));
    occluder += 0.027681f * exp(esm * (tap(cx + off, cy + -off) - sd0));
    occluder += 0.111014f * exp(esm * (tap(cx + -off, cy + 0.0f) - sd0));
    occluder += 0.445213f * exp(esm * (tap(cx + 0.0f, cy + 0.0f) - sd0));
    occluder += 0.111014f * exp(esm * (tap(cx + off, cy + 0.0f) - sd0));
    occluder += 0.027681f * exp(esm * (tap(cx + -off, cy + off) - sd0));
    occluder += 0.111014f * exp(esm * (tap(cx + 0.0f, cy + off) - sd0));
    occluder += 0.027681f * exp(esm * (tap(cx + off, cy + off) - sd0));
    occluder += 0.027681f;
    occluder = log(occluder);
    occluder += esm * sd0;
    float occlusion = clamp(exp(occluder - esm * lz), 0.0f, 1.0f);
    float4 tex = tex2D(texSamp, in.tex.xy);
    float3 n = normalize(in.norm.xyz);
    float3 l = normalize(in.lightDir.xyz);
    float3 e = normalize(in.eyeDir.xyz);
    float idiff = clamp(dot(l, n), 0.0f, 1.0f);
    float k2 = 2.0f * dot(l, n);
    float3 r = -(l - n * k2);
    float ispec = clamp(dot(r, e), 0.0f, 1.0f);
    float sp = pow(ispec, shininess);
    float3 illum = ambient.xyz + (diffuse.xyz * idiff + specular.xyz * sp) * occlusion;
    return float4(tex.xyz * illum, 1.0f);
}
"""


@pytest.mark.parametrize("w,h,samples", [(640, 360, 1), (400, 240, 4)])
def test_sasl_two_sampler_shadow_map_shader(cuda, cuda_jit_immediate, w, h, samples):
    """The colour pass of samples/StandardShadowMap written in SASL — TWO samplers (diffuse texture through tex2D, the shadow
    map through nine tex2Dlod taps), exp / log / pow.  (1) The visibility-first path (quad-granular k_shade) and the immediate
    k_raster path of the same module give identical frames.  (2) Against the built-in SLV_PS_SSM_DRAW (pinned to the reference,
    cases c5_ssm_*): geometry, depth, the shadow map and the counters are identical; the colour differs only through SASL's
    `log`, which is eflib's fast_log polynomial in the reference (sasl/src/drivers/compiler_impl.cpp:389, |error| < 0.01) where
    the cpp shader calls logf — a fraction of a percent of the occlusion term, i.e. a few LSB in penumbra pixels only."""
    sh = jit.compile(PS_SSM_DRAW, "ps", derivatives="cpp")
    assert sh.reflection.samplers == ["texSamp", "smSamp"]

    def sasl_scene(be):
        mod = jit.load(be, sh)
        sc = S.StandardShadowMap(w, h, samples, tex_size=64, textured_plane=True, ps_binding=lambda amb, dif, spe, shin, ts, ss: A.shader_binding(
            A.program_jit(mod), sh.unit.pack_uniforms({"ambient": amb, "diffuse": dif, "specular": spe, "shininess": float(shin)}), [ts, ss]))
        sc.setup(be)
        return sc

    got, imm = sasl_scene(cuda), sasl_scene(cuda_jit_immediate)
    ref = S.StandardShadowMap(w, h, samples, tex_size=64, textured_plane=True)
    ref.setup(cuda)
    for f in (1, 6):
        a, b, c = ref.run(cuda, f), got.run(cuda, f), imm.run(cuda_jit_immediate, f)
        assert cases.compare_frames(b, c) == [], f"frame {f}: visibility-first vs immediate"
        assert a.stats["ps_invocations"] > 1000
        for k in cases.GATED_COUNTERS:
            assert a.stats[k] == b.stats[k], k
        assert np.array_equal(a.depth.view(np.uint32), b.depth.view(np.uint32)) and np.array_equal(a.count.view(np.uint32), b.count.view(np.uint32))
        d = np.abs(a.color.astype(np.int32) - b.color.astype(np.int32))
        assert d.max() <= 6, f"frame {f}: SASL vs built-in colour differs by up to {d.max()} LSB"
        assert (d > 0).any(-1).mean() < 0.25, f"frame {f}: {(d > 0).any(-1).mean():.1%} of the samples differ"
        assert (d == 0).all(-1).mean() > 0.5


def test_sasl_intrinsics_match_the_reference_known_answers(cuda):
    """The reference's own known answers (tests/golden/sasl_kat.json: eflib values for the inputs and reference expressions of
    sasl/test/jit_test/general.cpp:159-433, generated by oracle/sasl_kat_gen.cpp) against the RUN-TIME COMPILED pixel shader on
    the GPU: one shader holds every call of the fixture, a draw per case writes the result into an rgba32f target.  Criterion:
    the reference test's own BOOST_CHECK_CLOSE tolerance (1e-4 %); most components are bit-identical."""
    import sasl_kat
    sh = jit.compile(sasl_kat.shader_source(), "ps")
    mod = jit.load(cuda, sh)
    t = S.create_targets(cuda, 8, 8, 1, A.PF_RGBA32F)
    mesh = S.create_planar((-3.0, -1.0, -3.0), (6, 0, 0), (0, 0, 6), 1, 1, True)
    mesh.elements = [(0, S._V4, 0, 0, 1.0)]
    mesh.upload(cuda)
    wvp = S.mat_mul(S.mat_lookat((0.0, 2.5, 0.0001), (0, 0, 0), (0, 1, 0)), S.mat_perspective_fov(np.pi / 2, 1.0, 0.1, 100.0))
    exact = total = 0
    for case in sasl_kat.CASES:
        cuda.clear_color(t.color, (-7.0, -7.0, -7.0, -7.0))
        cuda.clear_depth_stencil(t.ds, A.CLEAR_DEPTH | A.CLEAR_STENCIL, 1.0, 0)
        d = S.base_desc(t, 8, 8, cull=A.CULL_NONE)
        mesh.fill_desc(cuda, d)
        d.vs = A.shader_binding(A.VS_MVP_PASSTHROUGH, S.pack_vs_mvp_passthrough(wvp, [0]))
        d.ps = A.shader_binding(A.program_jit(mod), sasl_kat.uniforms_for(sh.unit, case))
        d.bs = A.shader_binding(A.BS_REPLACE)
        cuda.draw(d)
        px = np.frombuffer(cuda.read_texture(t.color), np.float32).reshape(8, 8, 4)
        got = px[4, 4]
        assert not np.array_equal(got, np.full(4, -7.0, np.float32)), "the probe pixel was not covered"
        assert np.array_equal(px[3, 3], got, equal_nan=True)  # uniform over the plane
        exact += sasl_kat.check(case, got)
        total += len(case["expected"])
    assert exact >= total * 0.8, (exact, total)


def test_sasl_skinning_vertex_shader_array_uniforms(cuda):
    """samples/AstroBoy's skinning vertex shader (AstroBoy.cpp:39-79): bone palettes in ARRAY uniforms sized by another global
    (device buffers whose addresses ride in the uniform block), `int4 BLEND_INDICES` inputs, run-time indices, a data-dependent
    `break`.  The GPU frame must equal, bit for bit, the frame of the same mesh with the shader's outputs precomputed on the
    host (the front end's code compiled for the host, itself pinned to a float32 numpy restatement in tests/test_sasl_frontend.py)
    and passed through SLV_VS_MVP_PASSTHROUGH with an identity matrix."""
    from sasl_host import HostShader
    from test_sasl_frontend import VS_SKIN, skin_test_data
    sh = jit.compile(VS_SKIN, "vs")
    mod = jit.load(cuda, sh)
    n = 24
    grid = S.create_planar((-3.0, 0.0, -3.0), (6.0 / n, 0, 0), (0, 0, 6.0 / n), n, n, True, index_dtype=np.uint32)
    nv = len(grid.streams[0])
    bones, invs, _, _, _, _ = skin_test_data(n_bones=6)
    rng = np.random.default_rng(4)
    pos = np.ascontiguousarray(grid.streams[0][:, :3], dtype=np.float32)
    pos[:, 1] += rng.uniform(-0.3, 0.3, nv).astype(np.float32)
    nrm = rng.uniform(0.0, 1.0, (nv, 3)).astype(np.float32)
    idx = rng.integers(0, 6, (nv, 4)).astype(np.int32)
    for v in range(nv):
        idx[v, 1 + v % 4:] = -1
    wts = rng.uniform(0.2, 0.6, (nv, 4)).astype(np.float32)
    view = S.mat_lookat((0.5, 4.0, -4.5), (0, 0, 0), (0, 1, 0))
    wvp = np.asarray(S.mat_mul(view, S.mat_perspective_fov(np.pi / 2, 16 / 9, 0.1, 100.0)), np.float32).reshape(4, 4)
    eye, light = np.array([0.5, 4.0, -4.5, 1], np.float32), np.array([3, 4, -1, 1], np.float32)

    def render(mesh, vs_binding):
        t = S.create_targets(cuda, 640, 360, 4, A.PF_RGBA8)
        cuda.query_begin()
        cuda.clear_color(t.color, (0.1, 0.1, 0.2, 1.0))
        cuda.clear_depth_stencil(t.ds, A.CLEAR_DEPTH | A.CLEAR_STENCIL, 1.0, 0)
        d = S.base_desc(t, 640, 360, cull=A.CULL_NONE)
        mesh.fill_desc(cuda, d)
        d.vs = vs_binding
        d.ps = A.shader_binding(A.PS_ATTR0_COLOR)
        d.bs = A.shader_binding(A.BS_REPLACE)
        cuda.draw(d)
        return S.read_frame(cuda, t, cuda.query_get())

    # (a) the SASL shader on the GPU: four streams (float3, float3, int4 as raw bits, float4), palettes in device buffers
    skinned = S.Mesh([pos, nrm, idx.view(np.float32), wts],
                     [(0, A.FMT_R32G32B32_FLOAT, 0, 0, 1.0), (1, A.FMT_R32G32B32_FLOAT, 1, 0, 0.0), (2, S._V4, 2, 0, 0.0), (3, S._V4, 3, 0, 0.0)],
                     grid.indices, grid.prim_count)
    hb, hi = cuda.create_buffer(np.ascontiguousarray(bones)), cuda.create_buffer(np.ascontiguousarray(invs))
    ub = sh.unit.pack_uniforms({"wvpMatrix": wvp, "eyePos": eye, "lightPos": light, "boneCount": len(bones),
                                "boneMatrices": cuda.buffer_device_ptr(hb)[0], "invMatrices": cuda.buffer_device_ptr(hi)[0]})
    ra = render(skinned, A.shader_binding(A.program_jit(mod), ub))
    # (b) the same shader evaluated on the host, its four outputs as vertex data
    hs = HostShader(sh.unit)
    bones_c, invs_c = np.ascontiguousarray(bones), np.ascontiguousarray(invs)
    ub_host = sh.unit.pack_uniforms({"wvpMatrix": wvp, "eyePos": eye, "lightPos": light, "boneCount": len(bones),
                                     "boneMatrices": bones_c.ctypes.data, "invMatrices": invs_c.ctypes.data})
    outs = np.zeros((nv, 4, 4), np.float32)
    for v in range(nv):
        regs = np.zeros((4, 4), np.float32)
        regs[0, :3], regs[1, :3], regs[2], regs[3] = pos[v], nrm[v], idx[v].view(np.float32), wts[v]
        outs[v] = hs.vs(regs, ub_host)[:4]
    pre = S.Mesh([np.ascontiguousarray(outs[:, k]) for k in range(4)], [(k, S._V4, k, 0, 1.0 if k == 0 else 0.0) for k in range(4)],
                 grid.indices, grid.prim_count)
    rb = render(pre, A.shader_binding(A.VS_MVP_PASSTHROUGH, S.pack_vs_mvp_passthrough(np.eye(4, dtype=np.float32), [1, 2, 3])))
    assert ra.stats["ps_invocations"] > 20000 and ra.stats["cprimitives"] == rb.stats["cprimitives"]
    assert not cases.compare_frames(ra, rb)

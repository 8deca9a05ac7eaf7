"""CPU suite: the reference's own known answers for SASL intrinsics (tests/golden/sasl_kat.json: eflib values for the inputs of
sasl/test/jit_test/general.cpp's `intrinsics` case) against the front end's generated code compiled for the host."""
import sasl_kat
from salviarenderer_b200.sasl import compile_shader
from sasl_host import HostShader


def test_fixture_shape():
    assert len(sasl_kat.CASES) >= 110 and len(sasl_kat.BRANCH) >= 38
    assert sasl_kat.FIXTURE["generator"] == "oracle/sasl_kat_gen.cpp"


def test_front_end_matches_the_reference_known_answers():
    unit = compile_shader(sasl_kat.shader_source(), "ps")
    hs = HostShader(unit)
    exact = total = 0
    for case in sasl_kat.CASES:
        got, keep = hs.ps([[0, 0, 0, 0]], sasl_kat.uniforms_for(unit, case))
        assert keep
        exact += sasl_kat.check(case, got)
        total += len(case["expected"])
    assert exact >= total * 0.8, (exact, total)  # most components are bit-identical to eflib; all are within 1e-4 %


def test_quad_known_answers_derivatives_and_loop():
    """general.cpp:1526-1602 (ddx_ddy) and :1668-1716 (ps_for_loop): the reference's srand(0) inputs and its expected values
    (sasl_kat.json "quad_cases"), through the front end and the generated code run for a whole 2x2 quad on the host - the four
    pixels as fibers that meet at every ddx / ddy (tests/sasl_host.py).  Pure float subtractions and additions: bit-identical."""
    import numpy as np
    import os
    from sasl_host import HostQuadShader
    for name, src, fn, n in (("ddx_ddy", sasl_kat.DERIVATIVES, "kat_derivatives", 10), ("for_loop", sasl_kat.FOR_LOOP, "kat_loop", 1)):
        case = sasl_kat.QUAD[name]
        unit = compile_shader(src, "ps")
        assert unit.reflection.uses_derivatives == (name == "ddx_ddy")
        got = HostQuadShader(unit, sasl_kat.quad_harness(fn, n, n)).run(case["inputs"], n)
        assert np.array_equal(got, np.asarray(case["expected"], np.float32)), name
    # the reference's own units, read in place where the tree exists
    repo = "/root/reference/sasl/test/repo"
    if os.path.isdir(repo):
        for name, fn, n in (("ddx_ddy", "fn", 10), ("for_loop", "fn", 1)):
            unit = compile_shader(open(os.path.join(repo, name + ".sps")).read(), "ps")
            got = HostQuadShader(unit, sasl_kat.quad_harness(fn, n, n)).run(sasl_kat.QUAD[name]["inputs"], n)
            assert np.array_equal(got, np.asarray(sasl_kat.QUAD[name]["expected"], np.float32)), name


def test_quad_derivative_conventions():
    """ddx / ddy in the cpp_pixel_shader convention (q1 - q0, q2 - q0 for the whole quad, cpp_pixel_shader.cpp:13-19) against
    SASL's per line / per column one: they agree on pixel 0 and differ on pixel 3 unless the field is affine."""
    import numpy as np
    from sasl_host import HostQuadShader
    unit = compile_shader("float2 g(float v: TEXCOORD0): COLOR { return float2(ddx(v), ddy(v)); }", "ps")
    v = np.array([[1.0], [4.0], [9.0], [25.0]], np.float32)   # pixel = row * 2 + col
    sasl = HostQuadShader(unit, sasl_kat.quad_harness("g", 1, 2)).run(v, 2)
    cpp = HostQuadShader(unit, sasl_kat.quad_harness("g", 1, 2), deriv_cpp=True).run(v, 2)
    assert sasl.tolist() == [[3, 8], [3, 21], [16, 8], [16, 21]]
    assert cpp.tolist() == [[3, 8]] * 4


def test_sasl_pixel_shader_with_texture_fetches_runs_on_the_cpu(oracle, reference):
    """bench.py's SASL Sponza pixel shader (tex2D = sample_2d_grad with the quad's per-pixel derivatives, times clamp(N.L)) run for
    whole quads on the host, its texture fetches served by the LIVE REFERENCE's sampler (slv_sampler_probe), against a float32
    restatement whose fetches go to the same sampler - trilinear and 16x anisotropic.  Checks what only the GPU suite saw so far:
    the derivative operands the generated code hands to the fetch (dudx, dvdx, dudy, dvdy per pixel, SASL convention)."""
    import numpy as np
    import bench
    from salviarenderer_b200 import abi as A, scenes as S
    from sasl_host import HostQuadShader
    f32 = np.float32
    unit = compile_shader(bench.SASL_PS_SPONZA, "ps")
    n_in = sum({"float": 1, "float2": 2, "float3": 3, "float4": 4}[t] for _, _, t in unit.reflection.inputs)
    assert [t for _, _, t in unit.reflection.inputs] == ["float4"] * 4  # uv, normal, light direction, eye direction
    hs = HostQuadShader(unit, sasl_kat.quad_harness(unit.reflection.entry, n_in, 4))
    rng = np.random.default_rng(5)
    for be in (reference, oracle):
        tex = S.make_texture(be, S.brick_texture(64, seed=3))
        for mipf, aniso in ((A.FILTER_LINEAR, 0), (A.FILTER_ANISOTROPIC, 16)):
            samp = be.create_sampler(A.sampler_desc(A.FILTER_LINEAR, A.FILTER_LINEAR, mipf, addr_u=A.ADDR_WRAP, addr_v=A.ADDR_WRAP, max_anisotropy=aniso), tex)
            hs.bind_samplers(be, samp)
            for _ in range(25):
                base = rng.uniform(-2, 2, 2)
                jac = rng.uniform(-0.2, 0.2, (2, 2)) * rng.choice([0.05, 1.0, 4.0])     # magnified, minified, strongly anisotropic
                uv = np.array([base + jac @ np.array([x, y]) for y in (0, 1) for x in (0, 1)], f32)   # pixel = row * 2 + col
                nrm, lgt = rng.standard_normal((4, 3)).astype(f32), rng.standard_normal((4, 3)).astype(f32)
                inp = np.zeros((4, 16), f32)
                inp[:, 0:2], inp[:, 4:7], inp[:, 8:11] = uv, nrm, lgt
                got = hs.run(inp, 4)
                ddx = np.array([uv[p | 1] - uv[p & ~1] for p in range(4)], f32)
                ddy = np.array([uv[p | 2] - uv[p & ~2] for p in range(4)], f32)
                texel = be.sampler_probe(samp, uv, ddx, ddy)
                for p in range(4):
                    def normalize(v):  # eflib normalize3: v * (1 / length), zero-length guard
                        ln = f32(np.sqrt(f32(f32(f32(v[0] * v[0]) + f32(v[1] * v[1])) + f32(v[2] * v[2]))))
                        inv = f32(f32(1.0) / (f32(1.0) if abs(ln) <= f32(1.1920928955078125e-7) else ln))
                        return (v * inv).astype(f32)
                    n, l = normalize(nrm[p]), normalize(lgt[p])
                    d = f32(f32(f32(l[0] * n[0]) + f32(l[1] * n[1])) + f32(l[2] * n[2]))
                    d = f32(min(max(d, f32(0.0)), f32(1.0)))
                    want = np.append((texel[p, :3] * d).astype(f32), f32(1.0))
                    assert np.array_equal(got[p], want), (be.name, mipf, p, got[p], want)

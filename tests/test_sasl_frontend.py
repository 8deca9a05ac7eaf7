"""CPU suite: the SASL front end (salviarenderer_b200/sasl/frontend.py).  Generated code is compiled for the host and
executed (tests/sasl_host.py); expected values are restated in numpy float32 with the same operation order.  The
shaders are written for these tests; the features are the ones sasl/test/repo/*.svs|*.sps and the samples' shaders use
(struct semantics, swizzles and write masks, branches, loops, intrinsics, constructors, casts, functions)."""
import os
import numpy as np
import pytest

from salviarenderer_b200.sasl import CompileError, compile_shader
from sasl_host import HostShader

f32 = np.float32

VS_SPONZA = """
float4x4 wvpMatrix;
float4   eyePos;
float4   lightPos;
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


def mul_row(v, m):
    """eflib transform: out[j] = ((v0*m0j + v1*m1j) + v2*m2j) + v3*m3j in float32."""
    v, m = np.asarray(v, f32), np.asarray(m, f32).reshape(4, 4)
    out = np.zeros(4, f32)
    for j in range(4):
        acc = f32(v[0] * m[0, j])
        for i in range(1, 4):
            acc = f32(acc + f32(v[i] * m[i, j]))
        out[j] = acc
    return out


def test_vertex_shader_reflection_and_values():
    unit = compile_shader(VS_SPONZA, "vs")
    r = unit.reflection
    assert r.entry == "vs_main"
    assert [(u[0], u[1], u[2], u[3]) for u in r.uniforms] == [("wvpMatrix", "float4x4", 0, 64), ("eyePos", "float4", 64, 16), ("lightPos", "float4", 80, 16)]
    assert r.uniform_bytes == 96
    assert [(s, i) for s, i, _ in r.inputs] == [("POSITION", 0), ("TEXCOORD", 0), ("NORMAL", 0)]
    assert [(s, i) for s, i, _ in r.outputs] == [("TEXCOORD", 0), ("TEXCOORD", 1), ("TEXCOORD", 2), ("TEXCOORD", 3)]
    assert r.n_vs_output_attrs == 4
    rng = np.random.default_rng(5)
    m = rng.standard_normal((4, 4)).astype(f32)
    eye, light = rng.standard_normal(4).astype(f32), rng.standard_normal(4).astype(f32)
    ub = unit.pack_uniforms({"wvpMatrix": m, "eyePos": eye, "lightPos": light})
    hs = HostShader(unit)
    for _ in range(16):
        pos, tex, nrm = (rng.standard_normal(4).astype(f32) for _ in range(3))
        o = hs.vs([pos, tex, nrm], ub)
        assert np.array_equal(o[0], mul_row(pos, m))
        assert np.array_equal(o[1], tex) and np.array_equal(o[2], nrm)
        assert np.array_equal(o[3], light - pos) and np.array_equal(o[4], eye - pos)
    with pytest.raises(KeyError):
        unit.pack_uniforms({"Shininess": 1.0})  # unknown names fail, as set_constant does upstream


PS_MISC = """
float4 tint;
int    steps;
struct PSIn  { float4 a: TEXCOORD0; float3 b: TEXCOORD1; float s: TEXCOORD2; };
struct PSOut { float4 c: COLOR0; };

float3 shade(float3 n, float3 l, float k) {
    float3 nn = normalize(n);
    float d = clamp(dot(normalize(l), nn), 0.0f, 1.0f);
    return lerp(float3(0.0f, 0.0f, 0.0f), nn, d) * k;
}

PSOut fn(PSIn in) {
    PSOut o;
    float3 x, y;
    x = (in.a).xyz;
    y = (in.a).wxy;
    o.c.yzx = x + y;              // write mask with a permutation
    o.c.w = 88.3f;
    if (in.s > 0.0f) { o.c.w = in.s; }
    if (in.s > 1.0f) { o.c.w = in.b.x; } else if (in.s > 0.5f) { o.c.w = o.c.w + 2.0f; }
    float acc = in.s;
    for (int i = 0; i < steps; i = i + 1) {
        if (i == 1) { continue; }
        acc = acc * 2.0f;
        if (acc > 5000.0f) { break; }
    }
    int n = 0;
    do { n += 2; } while (n < 5);
    float3 r = reflect(in.b, normalize(x));
    float3 sh = shade(in.b, x, tint.y);
    float3 cr = cross(in.b, x);
    o.c.x += acc + (float)n + r.z + sh.x + cr.y;
    o.c.y = (in.a.x > in.a.y ? in.a.x : in.a.y) * tint.x;
    o.c.z *= -tint.z;
    return o;
}
"""


def np_normalize(v):
    v = np.asarray(v, f32)
    acc = f32(v[0] * v[0])
    for c in v[1:]:
        acc = f32(acc + f32(c * c))
    ln = f32(np.sqrt(acc))
    if abs(ln) <= f32(1.1920928955078125e-7):
        ln = f32(1)
    return (v * f32(f32(1) / ln)).astype(f32)


def np_dot(a, b):
    acc = f32(a[0] * b[0])
    for x, y in zip(a[1:], b[1:]):
        acc = f32(acc + f32(x * y))
    return acc


def ps_misc_expected(a, b, s, tint, steps):
    a, b, tint = (np.asarray(v, f32) for v in (a, b, tint))
    s = f32(s)
    x, y = a[[0, 1, 2]], a[[3, 0, 1]]
    c = np.zeros(4, f32)
    t = (x + y).astype(f32)
    c[1], c[2], c[0] = t[0], t[1], t[2]
    c[3] = f32(88.3)
    if s > 0:
        c[3] = s
    if s > 1:
        c[3] = b[0]
    elif s > f32(0.5):
        c[3] = f32(c[3] + f32(2))
    acc = s
    for i in range(steps):
        if i == 1:
            continue
        acc = f32(acc * f32(2))
        if acc > f32(5000):
            break
    n = 6
    nx = np_normalize(x)
    d2 = f32(f32(2) * np_dot(b, nx))
    r = (b - (d2 * nx).astype(f32)).astype(f32)
    nn = np_normalize(b)
    d = np_dot(np_normalize(x), nn)
    d = f32(0) if d < 0 else (f32(1) if d > 1 else d)
    sh = ((np.zeros(3, f32) + ((nn - np.zeros(3, f32)).astype(f32) * d).astype(f32)).astype(f32) * tint[1]).astype(f32)
    cr_y = f32(f32(b[2] * x[0]) - f32(b[0] * x[2]))
    add = f32(f32(f32(f32(acc + f32(n)) + r[2]) + sh[0]) + cr_y)
    c[0] = f32(c[0] + add)
    c[1] = f32((a[0] if a[0] > a[1] else a[1]) * tint[0])
    c[2] = f32(c[2] * f32(-tint[2]))
    return c


def test_pixel_shader_control_flow_and_intrinsics():
    unit = compile_shader(PS_MISC, "ps")
    assert unit.reflection.entry == "fn"
    assert [(s, i, t) for s, i, t in unit.reflection.inputs] == [("TEXCOORD", 0, "float4"), ("TEXCOORD", 1, "float3"), ("TEXCOORD", 2, "float")]
    hs = HostShader(unit)
    rng = np.random.default_rng(11)
    for k in range(40):
        a, b = rng.standard_normal(4).astype(f32), rng.standard_normal(3).astype(f32)
        s = f32(rng.uniform(-1, 3))
        tint = rng.standard_normal(4).astype(f32)
        steps = int(rng.integers(0, 14))
        got, keep = hs.ps([a, list(b) + [0], [s, 0, 0, 0]], unit.pack_uniforms({"tint": tint, "steps": steps}))
        assert keep
        assert np.array_equal(got, ps_misc_expected(a, b, s, tint, steps)), (k, got, ps_misc_expected(a, b, s, tint, steps))


def test_matrix_ops_constructors_and_casts():
    src = """
    float4x4 M;
    float3x3 N;
    struct I { float4 v: TEXCOORD0; };
    float4 main(I i): COLOR {
        float4 a = mul(M, i.v);            // matrix x column vector
        float4 b = mul(i.v, transpose(M)); // == a, by the other path
        float3 c = mul(i.v.xyz, N);
        int k = (int)(i.v.x * 10.0f);
        uint u = (uint)abs(k);
        float4 r = float4(a.x - b.x, c.yz, float(k) + (float)(u & 3u));
        r.y += float2(1.0f, 2.0f).y + M[1].z + M._m23;
        bool big = any(i.v > float4(1.0f, 1.0f, 1.0f, 1.0f)) && !all(i.v > float4(0.0f, 0.0f, 0.0f, 0.0f));
        r.z = big ? r.z : -r.z;
        return r;
    }
    """
    unit = compile_shader(src, "ps")
    hs = HostShader(unit)
    rng = np.random.default_rng(3)
    M, N = rng.standard_normal((4, 4)).astype(f32), rng.standard_normal((3, 3)).astype(f32)
    ub = unit.pack_uniforms({"M": M, "N": N})
    assert unit.reflection.uniform("N")[2] == 64 and unit.reflection.uniform_bytes == 112
    for _ in range(20):
        v = rng.standard_normal(4).astype(f32)
        got, _ = hs.ps([v], ub)
        a0 = f32(M[0, 0] * v[0])
        for j in range(1, 4):
            a0 = f32(a0 + f32(M[0, j] * v[j]))
        c = np.zeros(3, f32)
        for j in range(3):
            acc = f32(v[0] * N[0, j])
            for i in range(1, 3):
                acc = f32(acc + f32(v[i] * N[i, j]))
            c[j] = acc
        k = int(f32(v[0] * f32(10)))
        u = abs(k)
        want = np.array([f32(a0 - a0), f32(c[1] + f32(f32(f32(2) + M[1, 2]) + M[2, 3])), c[2], f32(f32(k) + f32(u & 3))], f32)
        big = bool((v > 1).any()) and not bool((v > 0).all())
        if not big:
            want[2] = -want[2]
        assert np.array_equal(got, want), (got, want)


@pytest.mark.parametrize("src,stage,needle", [
    ("float4 main(float4 p: POSITION): SV_Position { return q; }", "vs", "undeclared"),
    ("float4 main(float4 p: POSITION): SV_Position { return ddx(p); }", "vs", "pixel shaders"),
    ("sampler s; float4 main(float4 t: TEXCOORD0): COLOR { float4 c = t; if (t.x > 0.0f) { c = tex2D(s, t.xy); } return c; }", "ps", "divergent"),
    ("float4 main(float4 p): SV_Position { return p; }", "vs", "semantic"),
    ("float4 main(float4 p: POSITION): TEXCOORD0 { return p; }", "vs", "SV_Position"),
    ("float4 main(float4 p: POSITION): SV_Position { return p.xyzq; }", "vs", "cannot take"),
    ("sampler a; sampler b; sampler c; float4 main(float4 t: TEXCOORD0): COLOR { return tex2D(a, t.xy) + tex2D(b, t.xy) + tex2D(c, t.xy); }",
     "ps", "two samplers"),
    ("sampler a; sampler b; float4 main(float4 p: POSITION): SV_Position { return p + tex2Dlod(a, p) + tex2Dlod(b, p); }", "vs", "one per vertex"),
    ("float4 main(float4 p: POSITION): SV_Position { return p @ p; }", "vs", "unexpected character"),
])
def test_compile_errors(src, stage, needle):
    with pytest.raises(CompileError) as e:
        compile_shader(src, stage)
    assert needle in str(e.value)


def test_texture_shader_reflection():
    unit = compile_shader("""
        sampler texSamp;
        float alpha;
        float4 ps_main(float4 uv: TEXCOORD0): COLOR { float4 c = tex2D(texSamp, uv.xy); c.w = alpha; return c; }
    """, "ps")
    r = unit.reflection
    assert r.samplers == ["texSamp"] and r.uses_derivatives and r.uniform("alpha")[2:] == (0, 4)
    assert "sasl_tex2d_grad" in unit.code and "sasl_ddx" in unit.code


# The reference's own SASL test units (sasl/test/repo) and the shadow-map sample's shaders: the ones inside the supported
# subset must go through the front end AND the generated code must compile (host C++).  Read in place, never copied;
# skipped where /root/reference does not exist (the GPU box).
REFERENCE_UNITS_OK = [
    "arithmetic.sps", "arithmetic.ss", "array_and_index.ss", "assigns.ss", "bit_ops.ss", "bool.ss", "branches.sps", "branches.ss", "casts.ss", "comments.ss",
    "constructors.ss", "ddx_ddy.sps", "decl.ss", "deps.ss", "do_while.sps", "empty.ss", "for_loop.sps", "function.ss", "host_intrinsic_detection.ss",
    "initializer.ss", "intrinsics.sps", "intrinsics.ss", "intrinsics.svs", "local_var.ss", "null.ss", "swizzle.ss", "swizzle_and_wm.sps", "tex.sps",
    "unary_operators.ss", "vec_and_mat.sps", "vec_and_mat.svs", "while.sps",
    "input_assigned.svs", "semantic_fn.svs", "semfn_par.svs", "struct_semin.svs",
]
REFERENCE_UNITS_REJECTED = {  # outside the subset (or erroneous on purpose upstream): must fail with CompileError, not crash
    "incomplete.ss", "semantic_errors.ss", "scalar.sps",
}


def _ref_unit(name):
    import os
    for d in ("/root/reference/sasl/test/repo", "/root/reference/resources/ssm"):
        p = os.path.join(d, name)
        if os.path.exists(p):
            return open(p, errors="replace").read()
    pytest.skip("reference tree not present")


@pytest.mark.parametrize("name", REFERENCE_UNITS_OK + ["tex.svs", "Draw.savs", "GenSM.savs", "GenSM.saps"])
def test_reference_sasl_units_compile(name, tmp_path):
    import subprocess
    from sasl_host import RT_DIR
    ext = name.rsplit(".", 1)[1]
    stage = "vs" if ext in ("svs", "savs") else ("ps" if ext in ("sps", "saps") else "lib")
    unit = compile_shader(_ref_unit(name), stage)
    src = tmp_path / "u.cpp"
    body = unit.code
    if stage == "lib":  # functions only: instantiate nothing, but the code must parse and type-check
        body = "namespace slv { struct RasterParams; }\n" + body
    src.write_text('#include "sasl_rt.h"\n' + body + "\nint main() { return 0; }\n")
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-I" + RT_DIR, str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]


@pytest.mark.parametrize("name", sorted(REFERENCE_UNITS_REJECTED))
def test_reference_sasl_units_outside_subset_fail_cleanly(name):
    ext = name.rsplit(".", 1)[1]
    stage = "vs" if ext == "svs" else ("ps" if ext == "sps" else "lib")
    with pytest.raises(CompileError):
        compile_shader(_ref_unit(name), stage)


# ---- the math intrinsics mirror the host functions the reference binds its JIT-ed code to, and `switch` --------------------
PS_MATH = """
float4 k;
int    sel;
struct PSIn { float4 a: TEXCOORD0; float4 b: TEXCOORD1; };
int pick(int n, int base) {
    int ret = 0;
    switch (n) {
    case 1: return base;
    case 2: return base * base;
    case 3:
    case 4: ret = base + 100;
    case 7: ret = ret + 1; break;
    case -2: ret = -5; break;
    default: return 0;
    }
    return ret;
}
float4 fn(PSIn in): COLOR {
    float4 r;
    r.x = log(in.a.x) + log2(in.a.y) * 0.5f + exp2(in.a.z);
    r.y = floor(in.b.x) + ceil(in.b.y) * 2.0f + round(in.b.z) * 4.0f + trunc(in.b.w) * 8.0f;
    r.z = frac(in.b.x) + ldexp(in.a.w, in.b.y);
    float4 d = dst(in.a, in.b);
    int acc = 0;
    for (int i = 0; i < 6; ++i) {
        switch (i) {
        case 1: continue;
        case 4: acc = acc + 1000; break;
        default: acc = acc + pick(sel + i, 3);
        }
        acc = acc + 1;
    }
    r.w = d.y + d.z + d.w + d.x + (float)acc;
    return r;
}
"""


def _np_fast_log2(v):
    x = np.array(v, np.float32).view(np.int32).astype(np.int64)
    log_2 = ((x >> 23) & 255) - 128
    x = (x & ~(255 << 23)) + (127 << 23)
    f = np.array(x, np.int32).view(np.float32)
    f = f32(f32(f32(f32(f32(-1.0) / f32(3)) * f) + f32(2)) * f) - f32(f32(2.0) / f32(3))
    return f32(f32(f) + f32(log_2))


def _np_fast_round(v):
    v = f32(v)
    bias = f32(-8388608.0) if np.signbit(v) else f32(8388608.0)
    return f32(f32(v + bias) - bias)


def _np_floor(v):
    f = _np_fast_round(v)
    return f32(f - f32(1)) if f > v else f


def _np_ceil(v):
    f = _np_fast_round(v)
    return f32(f + f32(1)) if f < v else f


def _py_pick(n, base):
    ret = 0
    if n == 1:
        return base
    if n == 2:
        return base * base
    if n in (3, 4):
        return base + 100 + 1
    if n == 7:
        return 1
    if n == -2:
        return -5
    return 0


def test_math_intrinsics_mirror_the_reference_and_switch():
    """log / log2 = eflib fast_log / fast_log2, exp2 = ldexpf(1, (int)x), floor / ceil / round / trunc = eflib's magic-number
    versions, frac = |v| - floor(|v|), ldexp truncates its exponent (sasl/src/drivers/compiler_impl.cpp:340-404,
    sasl/src/codegen/cg_impl.cpp:1139-1150); switch with fall-through, negative labels, default, and break / continue inside a
    loop (sasl/test/repo/branches.ss)."""
    unit = compile_shader(PS_MATH, "ps")
    hs = HostShader(unit)
    rng = np.random.default_rng(5)
    for kcase in range(60):
        a = rng.uniform(0.05, 9.0, 4).astype(f32)
        b = (rng.uniform(-6.0, 6.0, 4)).astype(f32)
        if kcase % 7 == 0:
            b[2] = f32(np.floor(b[2])) + f32(0.5)      # ties: round half to even
        sel = int(rng.integers(-4, 6))
        got, keep = hs.ps([a, b], unit.pack_uniforms({"k": (0, 0, 0, 0), "sel": sel}))
        assert keep
        exp2 = f32(np.ldexp(np.float32(1.0), int(a[2])))
        rx = f32(f32(f32(_np_fast_log2(a[0]) * f32(0.69314718)) + f32(_np_fast_log2(a[1]) * f32(0.5))) + exp2)
        tr = _np_floor(b[3]) if b[3] > 0 else _np_ceil(b[3])
        ry = f32(f32(f32(_np_floor(b[0]) + f32(_np_ceil(b[1]) * f32(2))) + f32(_np_fast_round(b[2]) * f32(4))) + f32(tr * f32(8)))
        ab = f32(abs(b[0]))
        rz = f32(f32(ab - _np_floor(ab)) + f32(np.ldexp(a[3], int(b[1]))))
        acc = 0
        for i in range(6):
            if i == 1:
                continue
            if i == 4:
                acc += 1000
            else:
                acc += _py_pick(sel + i, 3)
            acc += 1
        rw = f32(f32(f32(f32(f32(a[1] * b[1]) + a[2]) + b[3]) + f32(1.0)) + f32(acc))
        want = np.array([rx, ry, rz, rw], f32)
        assert np.array_equal(got, want), (kcase, got, want, a, b, sel)


PS_MORE = """
int4 bits;
struct PSIn { float4 a: TEXCOORD0; float4 b: TEXCOORD1; float4 c: TEXCOORD2; };
float4 fn(PSIn in): COLOR {
    float3 r = refract(in.a.xyz, in.b.xyz, in.a.w);
    float3 ff = faceforward(in.a.xyz, in.b.xyz, in.c.xyz);
    float4 l = lit(in.c.x, in.c.y, in.c.z);
    uint2 hb = firstbithigh(uint2(asuint(bits.x), asuint(bits.y)));
    int lb = firstbitlow(bits.z);
    uint rb = reversebits(asuint(bits.w));
    bool3 cls = bool3(isnan(in.c.w), isinf(in.b.w), isfinite(in.b.w));
    float4 o;
    o.x = r.x + r.y * 2.0f + r.z * 4.0f;
    o.y = ff.x + ff.y * 2.0f + ff.z * 4.0f + rcp(in.c.z);
    o.z = l.x + l.y * 2.0f + l.z * 4.0f + l.w * 8.0f;
    o.w = (float)hb.x + (float)hb.y * 64.0f + (float)lb * 4096.0f + (float)(rb >> 24) * 0.001f + (cls.x ? 1.0f : 0.0f) * 0.25f
          + (cls.y ? 1.0f : 0.0f) * 0.5f + (cls.z ? 1.0f : 0.0f) * 0.125f;
    return o;
}
"""


def test_more_intrinsics_follow_the_reference_code_generator():
    """refract / faceforward / lit in the reference's order of operations (sasl/src/codegen/cg_impl.cpp:1229-1352), the bit
    intrinsics of compiler_impl.cpp:414-432 and the float classification of cgs.cpp:1828-1842."""
    unit = compile_shader(PS_MORE, "ps")
    hs = HostShader(unit)
    rng = np.random.default_rng(9)
    for kcase in range(50):
        a, b, c = (rng.standard_normal(4).astype(f32) for _ in range(3))
        a[3] = f32(rng.uniform(0.3, 1.6))
        if kcase % 5 == 0:
            c[3] = f32(np.nan)
        if kcase % 4 == 0:
            b[3] = f32(np.inf)
        bits = [int(v) for v in rng.integers(1, 2 ** 31 - 1, 4)]
        got, keep = hs.ps([a, b, c], unit.pack_uniforms({"bits": bits}))
        eta = a[3]
        ndi = f32(f32(f32(b[0] * a[0]) + f32(b[1] * a[1])) + f32(b[2] * a[2]))
        k = f32(f32(1) - f32(f32(eta * eta) * f32(f32(1) - f32(ndi * ndi))))
        flag = k < 0
        k = f32(0) if flag else k
        rr = f32(f32(eta * ndi) + np.sqrt(k, dtype=f32))
        r = [f32(0) if flag else f32(f32(eta * a[i]) - f32(rr * b[i])) for i in range(3)]
        idn = f32(f32(f32(b[0] * c[0]) + f32(b[1] * c[1])) + f32(b[2] * c[2]))
        ff = [a[i] if idn < 0 else f32(f32(0) - a[i]) for i in range(3)]
        l = [f32(1), f32(0) if c[0] < 0 else c[0], f32(0) if (c[0] < 0 or c[1] < 0) else f32(c[1] * c[2]), f32(1)]
        ox = f32(f32(r[0] + f32(r[1] * f32(2))) + f32(r[2] * f32(4)))
        oy = f32(f32(f32(ff[0] + f32(ff[1] * f32(2))) + f32(ff[2] * f32(4))) + f32(f32(1) / c[2]))
        oz = f32(f32(f32(l[0] + f32(l[1] * f32(2))) + f32(l[2] * f32(4))) + f32(l[3] * f32(8)))
        clz = lambda v: 32 - int(v).bit_length()
        ctz = lambda v: (int(v) & -int(v)).bit_length() - 1
        rev = int(format(bits[3], "032b")[::-1], 2)
        ow = f32(f32(clz(bits[0])) + f32(f32(clz(bits[1])) * f32(64)))
        ow = f32(ow + f32(f32(ctz(bits[2])) * f32(4096)))
        ow = f32(ow + f32(f32(rev >> 24) * f32(0.001)))
        ow = f32(ow + f32(f32(1.0 if np.isnan(c[3]) else 0.0) * f32(0.25)))
        ow = f32(ow + f32(f32(1.0 if np.isinf(b[3]) else 0.0) * f32(0.5)))
        ow = f32(ow + f32(f32(1.0 if np.isfinite(b[3]) else 0.0) * f32(0.125)))
        want = np.array([ox, oy, oz, ow], f32)
        assert np.array_equal(got, want, equal_nan=True), (kcase, got, want)


def test_componentwise_logical_operators_on_vectors():
    """`||` / `&&` on bool vectors work per component (sasl/test/repo/bool.ss: `i > j || i > k && i <= j + k` on int3 / float3x4)."""
    src = """
    struct PSIn { float4 a: TEXCOORD0; float4 b: TEXCOORD1; float4 c: TEXCOORD2; };
    float4 fn(PSIn in): COLOR {
        bool4 r = in.a > in.b || in.a > in.c && in.a <= in.b + in.c;
        return float4(r.x ? 1.0f : 0.0f, r.y ? 1.0f : 0.0f, r.z ? 1.0f : 0.0f, r.w ? 1.0f : 0.0f);
    }
    """
    unit = compile_shader(src, "ps")
    hs = HostShader(unit)
    rng = np.random.default_rng(3)
    for _ in range(40):
        a, b, c = (rng.integers(-3, 4, 4).astype(f32) for _ in range(3))
        got, _keep = hs.ps([a, b, c])
        want = ((a > b) | ((a > c) & (a <= (b + c).astype(f32)))).astype(f32)
        assert np.array_equal(got, want), (a, b, c, got, want)


def test_forward_calls_and_recursion():
    """A function may be called before its definition, and may call itself (sasl/test/repo/function.ss: fib): generation
    follows the call graph, functions on a cycle get a prototype and are not force-inlined."""
    src = """
    int n;
    struct PSIn { float4 a: TEXCOORD0; };
    float4 fn(PSIn in): COLOR {
        return float4((float)fib(n), later(in.a.x), (float)even(n), 1.0f);
    }
    float later(float x) { return x * 2.0f + 1.0f; }
    int fib(int i) {
        if (i < 2) { return i; }
        return fib(i - 1) + fib(i - 2);
    }
    int even(int i) { if (i == 0) { return 1; } return odd(i - 1); }
    int odd(int i) { if (i == 0) { return 0; } return even(i - 1); }
    """
    unit = compile_shader(src, "ps", "fn")
    hs = HostShader(unit)
    fib = lambda i: i if i < 2 else fib(i - 1) + fib(i - 2)
    for n in (0, 1, 2, 7, 12):
        got, _ = hs.ps([[0.25 * n, 0, 0, 0]], unit.pack_uniforms({"n": n}))
        assert np.array_equal(got, np.array([fib(n), f32(f32(0.25 * n) * f32(2)) + f32(1), 1 - n % 2, 1], f32)), (n, got)


def test_jit_cubins_hold_no_fused_packed_multiply_add():
    """The run-time compiled kernels include the library's own headers (packed fp32 pairs): no FFMA2 may appear (tests/test_abi.py)."""
    import shutil
    import subprocess
    import tempfile
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not (os.path.exists(cuobjdump) and (shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc"))):
        pytest.skip("CUDA toolkit not available")
    import bench
    from salviarenderer_b200.sasl import jit
    for src, stage in ((bench.SASL_PS_SPONZA, "ps"), (bench.SASL_VS_SPONZA, "vs")):
        sh = jit.compile(src, stage)
        with tempfile.NamedTemporaryFile(suffix=".cubin") as f:
            f.write(sh.cubin)
            f.flush()
            sass = subprocess.run([cuobjdump, "-sass", f.name], capture_output=True, text=True).stdout
        assert sass.count("FADD2") > 10 and sass.count("FFMA2") == 0, stage


# ---- preprocessor (the reference runs Boost.Wave in front of its parser) ---------------------------------------------------------
def test_preprocessor_directives_and_macros():
    from salviarenderer_b200.sasl.preprocess import PreprocessError, preprocess
    out = preprocess("#define N 4\n#define SQ(x) ((x)*(x))\nfloat a[N]; int b = SQ(N+1);\n#if N > 3 && !defined(Q)\nint yes;\n#elif N\nint maybe;\n"
                     "#else\nint no;\n#endif\n#undef N\nint N;\n/* a\n   block */ int c; // N stays\n")
    lines = out.split("\n")
    assert lines[2] == "float a[4]; int b = ((4+1)*(4+1));" and lines[4] == "int yes;" and lines[6] == "" and lines[8] == ""
    assert lines[11] == "int N;" and "int c;" in lines[13] and len(lines) == 15  # line numbers of the source survive
    assert preprocess("#ifdef A\nx\n#else\ny\n#endif", defines={"A": None}).split("\n")[1] == "x"
    assert preprocess("#define A B\n#define B A\nA B").split("\n")[2] == "A B"  # self-reference stops the expansion
    for bad in ("#if 1\nx", "#endif", "#else", "#include \"nope.ss\"", "#error stop", "#frobnicate", "#define F(a) a\nF(1, 2)"):
        with pytest.raises(PreprocessError):
            preprocess(bad)
    # a shader through the front end: directives, a function-like macro, a virtual include
    src = """#include <common.sasl>
    #define SCALE(v) ((v) * gain)
    #ifndef GAIN_DEFAULT
    #  define GAIN_DEFAULT 2.0f
    #endif
    struct PSIn { float4 c: TEXCOORD0; };
    float4 ps_main(PSIn in): COLOR { float gain = GAIN_DEFAULT; return SCALE(in.c) + bias(); }
    """
    unit = compile_shader(src, "ps", virtual_files={"common.sasl": "float4 bias() { return float4(0.5f, 0.25f, 0.0f, 1.0f); }"})
    got, _ = HostShader(unit).ps([[1, 2, 3, 4]])
    assert np.array_equal(got, np.array([2.5, 4.25, 6.0, 9.0], f32))
    unit = compile_shader(src, "ps", defines={"GAIN_DEFAULT": "0.5f"}, virtual_files={"common.sasl": "float4 bias() { return float4(0.0f, 0.0f, 0.0f, 0.0f); }"})
    got, _ = HostShader(unit).ps([[1, 2, 3, 4]])
    assert np.array_equal(got, np.array([0.5, 1.0, 1.5, 2.0], f32))
    with pytest.raises(CompileError):
        compile_shader(src, "ps")  # the include cannot be found


def test_reference_preprocessor_units():
    """sasl/test/repo/{preprocessors,include_main,include_header,include_search_path}.ss as the reference's tests drive them
    (sasl/test/jit_test/general.cpp:111-119: main() == 0; the driver tests add a virtual file and search paths)."""
    import os
    repo = "/root/reference/sasl/test/repo"
    if not os.path.isdir(repo):
        pytest.skip("reference tree not present")
    unit = compile_shader(_ref_unit("preprocessors.ss"), "lib")
    assert "main" in unit.code
    with pytest.raises(CompileError):  # the guarded garbage becomes visible
        compile_shader(_ref_unit("preprocessors.ss"), "lib", defines={"SASL_COMPILER_ERROR": None})
    main = os.path.join(repo, "include_main.ss")
    unit = compile_shader(open(main).read(), "lib", file_name=main,
                          virtual_files={"virtual_include.ss": "float virtual_add(float a, float b) { return a + b; }"})
    assert all(fn in unit.code for fn in ("header_add", "virtual_add", "main_add"))
    with pytest.raises(CompileError):
        compile_shader(open(main).read(), "lib", file_name=main)  # <virtual_include.ss> exists nowhere on disk
    with pytest.raises(CompileError):
        compile_shader(open(main).read(), "lib", file_name=main, defines={"FAILED_INCLUDE": None},
                       virtual_files={"virtual_include.ss": "float virtual_add(float a, float b) { return a + b; }"})
    sp = os.path.join(repo, "include_search_path.ss")
    compile_shader(open(sp).read(), "lib", file_name=sp, include_dirs=[os.path.join(repo, "include")], sys_include_dirs=[os.path.join(repo, "sysincl")])
    with pytest.raises(CompileError):
        compile_shader(open(sp).read(), "lib", file_name=sp)


# ---- array uniforms, integer inputs, run-time indices: the skinning vertex shader of samples/AstroBoy (AstroBoy.cpp:39-79) ----------
VS_SKIN = """
float4x4 wvpMatrix; float4 eyePos; float4 lightPos;
int      boneCount;
float4x4 boneMatrices[boneCount];
float4x4 invMatrices[boneCount];
struct VSIn  { float3 pos: POSITION; float3 norm: NORMAL; int4 indices: BLEND_INDICES; float4 weights: BLEND_WEIGHTS; };
struct VSOut { float4 pos: sv_position; float4 norm: TEXCOORD1; float4 lightDir: TEXCOORD2; float4 eyeDir: TEXCOORD3; };
VSOut vs_main(VSIn in) {
    VSOut o;
    float4 n4 = float4(in.norm, 0.0f);
    float4 p4 = float4(in.pos, 1.0f);
    float4 skin_pos = float4(0.0f, 0.0f, 0.0f, 0.0f);
    float4 skin_nor = float4(0.0f, 0.0f, 0.0f, 0.0f);
    for (int i = 0; i < 4; ++i) {
        float4 w = in.weights[i].xxxx;
        int boneId = in.indices[i];
        if (boneId == -1) { break; }
        float4 posInBoneSpace = mul(invMatrices[boneId], p4);
        skin_pos += (mul(boneMatrices[boneId], posInBoneSpace) * w);
        float4 norInBoneSpace = mul(invMatrices[boneId], n4);
        skin_nor += (mul(boneMatrices[boneId], norInBoneSpace) * w);
    }
    o.pos = mul(skin_pos, wvpMatrix);
    o.norm = skin_nor;
    o.lightDir = lightPos - skin_pos;
    o.eyeDir = eyePos - skin_pos;
    return o;
}
"""


def skin_reference(pos, norm, indices, weights, bones, invs, wvp, eye, light):
    """float32 restatement of VS_SKIN: mul(M, v) = rows of M dot v, mul(v, M) = v times the columns, sums left to right."""
    def mv(M, v):
        return np.array([f32(f32(f32(f32(M[r, 0] * v[0]) + f32(M[r, 1] * v[1])) + f32(M[r, 2] * v[2])) + f32(M[r, 3] * v[3])) for r in range(4)], f32)

    def vm(v, M):
        return np.array([f32(f32(f32(f32(v[0] * M[0, c]) + f32(v[1] * M[1, c])) + f32(v[2] * M[2, c])) + f32(v[3] * M[3, c])) for c in range(4)], f32)

    p4, n4 = np.array([*pos, 1.0], f32), np.array([*norm, 0.0], f32)
    sp, sn = np.zeros(4, f32), np.zeros(4, f32)
    for i in range(4):
        b = int(indices[i])
        if b == -1:
            break
        w = f32(weights[i])
        sp = (sp + mv(bones[b], mv(invs[b], p4)) * w).astype(f32)
        sn = (sn + mv(bones[b], mv(invs[b], n4)) * w).astype(f32)
    return vm(sp, wvp), sn, (light - sp).astype(f32), (eye - sp).astype(f32)


def skin_test_data(n_bones=5, n_verts=64, seed=9):
    rng = np.random.default_rng(seed)
    bones = (np.eye(4, dtype=f32)[None] + rng.uniform(-0.3, 0.3, (n_bones, 4, 4)).astype(f32)).astype(f32)
    invs = (np.eye(4, dtype=f32)[None] + rng.uniform(-0.2, 0.2, (n_bones, 4, 4)).astype(f32)).astype(f32)
    pos = rng.uniform(-1, 1, (n_verts, 3)).astype(f32)
    norm = rng.uniform(-1, 1, (n_verts, 3)).astype(f32)
    idx = rng.integers(0, n_bones, (n_verts, 4)).astype(np.int32)
    for v in range(n_verts):  # 1 .. 4 influences; -1 terminates the list
        k = 1 + v % 4
        idx[v, k:] = -1
    wts = rng.uniform(0.1, 1.0, (n_verts, 4)).astype(f32)
    return bones, invs, pos, norm, idx, wts


def test_skinning_vertex_shader_array_uniforms_and_integer_inputs():
    unit = compile_shader(VS_SKIN, "vs")
    r = unit.reflection
    assert r.arrays == {"boneMatrices": ("float4x4", 64, "boneCount"), "invMatrices": ("float4x4", 64, "boneCount")}
    assert [u[1] for u in r.uniforms] == ["float4x4", "float4", "float4", "int", "float4x4[]", "float4x4[]"]
    assert r.inputs == [("POSITION", 0, "float3"), ("NORMAL", 0, "float3"), ("BLEND_INDICES", 0, "int4"), ("BLEND_WEIGHTS", 0, "float4")]
    hs = HostShader(unit)
    bones, invs, pos, norm, idx, wts = skin_test_data()
    wvp = (np.eye(4, dtype=f32) + np.arange(16, dtype=f32).reshape(4, 4) * f32(0.01)).astype(f32)
    eye, light = np.array([0, 2, -5, 1], f32), np.array([3, 4, -1, 1], f32)
    bones_c, invs_c = np.ascontiguousarray(bones), np.ascontiguousarray(invs)
    ub = unit.pack_uniforms({"wvpMatrix": wvp, "eyePos": eye, "lightPos": light, "boneCount": len(bones),
                             "boneMatrices": bones_c.ctypes.data, "invMatrices": invs_c.ctypes.data})
    for v in range(len(pos)):
        regs = np.zeros((4, 4), f32)
        regs[0, :3], regs[1, :3], regs[3] = pos[v], norm[v], wts[v]
        regs[2] = idx[v].view(f32)  # the register carries the integers' bits
        out = hs.vs(regs, ub)
        want = skin_reference(pos[v], norm[v], idx[v], wts[v], bones, invs, wvp, eye, light)
        for k in range(4):
            assert np.array_equal(out[k].view(np.uint32), want[k].view(np.uint32)), (v, k, out[k], want[k])
    with pytest.raises(CompileError):
        compile_shader("float4 a[n];\nfloat4 f(float4 p: POSITION): SV_Position { return a[0]; }", "vs")  # size is not a global


def test_reference_array_unit_compiles():
    unit = compile_shader(_ref_unit("array.svs"), "vs")  # int mat_size; float4x4 mat_arr[mat_size]; int4 BLEND_INDICES input
    assert unit.reflection.arrays == {"mat_arr": ("float4x4", 64, "mat_size")}


def test_writes_to_globals_and_inputs_are_local_copies():
    """sasl/test/repo/input_assigned.svs: a shader may assign to a global and to its inputs; the uniform block itself is never
    written (the function works on a copy initialised from it), so the next invocation starts from the uniform's value again."""
    src = """
    float x;
    struct VSIN  { float4 pos: SV_Position; };
    struct VSOUT { float4 pos: SV_Position; float4 k: TEXCOORD0; };
    VSOUT fn(VSIN in) {
        VSOUT o;
        x += 0.5f;
        in.pos.x += x;
        o.pos = in.pos;
        o.pos.x += 0.5f;
        o.k = float4(x, x * 2.0f, 0.0f, 0.0f);
        return o;
    }
    """
    unit = compile_shader(src, "vs")
    hs = HostShader(unit)
    ub = unit.pack_uniforms({"x": 1.0})
    for _ in range(2):  # twice: the second run must not see the first run's write
        out = hs.vs([[3.0, 4.0, 5.0, 1.0]], ub)
        assert np.array_equal(out[0], np.array([5.0, 4.0, 5.0, 1.0], f32)) and np.array_equal(out[1], np.array([1.5, 3.0, 0, 0], f32))
    # a scalar SV_Position (semfn_par.svs) is padded with zeros
    unit = compile_shader("float fn(float a: SV_Position): SV_Position { return a * 2.0f; }", "vs")
    assert np.array_equal(HostShader(unit).vs([[1.5, 9, 9, 9]])[0], np.array([3.0, 0, 0, 0], f32))


def test_reference_semantic_order_replays_the_reference_array():
    """reflection_impl.cpp:70-118 inserts at std::lower_bound under semantic_value::operator< (constants.h:94-96), which is not
    a strict weak order: the resulting array - and with it the attribute a semantic lands on - depends on the insertion order."""
    from salviarenderer_b200.sasl.frontend import reference_semantic_order as order
    T, N, P = "TEXCOORD", "NORMAL", "SV_POSITION"
    assert order([(P, 0), (T, 0), (T, 1), (T, 2), (T, 3)]) == [0, 1, 2, 3, 4]         # ascending: declaration order
    assert order([(T, 1), (T, 0)]) == [1, 0]                                          # sorted by index
    assert order([(T, 2), (P, 0), (T, 0), (T, 1)]) == [1, 2, 3, 0]
    # NORMAL0 < TEXCOORD1 (index) and TEXCOORD1 < NORMAL0 (system value): the later one is inserted BEHIND the earlier one
    assert order([(T, 1), (N, 0)]) == [0, 1]
    assert order([(N, 0), (T, 1)]) == [0, 1]
    assert order([(N, 0), (T, 0)]) == [1, 0]                                          # TEXCOORD0 < NORMAL0 only
    assert order([("COLOR", 0), ("SV_Target".upper(), 1), (T, 0)]) == [2, 0, 1]       # COLOR and SV_Target are one system value
    assert order([("FOG", 0), ("BINORMAL", 0), (T, 0)]) == [2, 1, 0]                  # customised semantics compare by name
    with pytest.raises(CompileError, match="bound twice"):
        order([(T, 0), ("POSITION", 0), (P, 0)])                                       # POSITION == SV_Position


def test_pixel_shader_inputs_map_to_attributes_in_reference_semantic_order():
    """With a C++ vertex shader bound the reference feeds attribute k to the k-th entry of the pixel shader's semantic array
    (pixel_shader_unit::update, shader_unit.cpp:106-140), whatever the declaration order."""
    src = """
    struct I { float4 late: TEXCOORD2; float2 first: TEXCOORD0; float3 mid: TEXCOORD1; };
    float4 main(I i): COLOR { return float4(i.first.x, i.mid.y, i.late.z, i.late.w); }
    """
    unit = compile_shader(src, "ps")
    assert unit.reflection.inputs == [("TEXCOORD", 0, "float2"), ("TEXCOORD", 1, "float3"), ("TEXCOORD", 2, "float4")]
    got, keep = HostShader(unit).ps([[1, 2, 3, 4], [5, 6, 7, 8], [9, 10, 11, 12]], b"")
    assert keep and np.array_equal(got, np.array([1, 6, 11, 12], f32))
    # a lone TEXCOORD1 is entry 0 of the array: it reads attribute 0 (the reference's positional mapping)
    unit = compile_shader("float4 main(float4 uv: TEXCOORD1): COLOR { return uv; }", "ps")
    got, _ = HostShader(unit).ps([[1, 2, 3, 4], [5, 6, 7, 8]], b"")
    assert np.array_equal(got, np.array([1, 2, 3, 4], f32))


def test_vertex_shader_outputs_fill_attribute_registers_in_reference_semantic_order():
    """Attribute registers are numbered along the vertex shader's output-semantic array, position aside
    (sasl/src/shims/interp_shim.cpp:57-76)."""
    src = """
    struct O { float4 b: TEXCOORD1; float4 pos: SV_Position; float4 a: TEXCOORD0; };
    O main(float4 p: POSITION) { O o; o.pos = p; o.a = p * 2.0f; o.b = p * 3.0f; return o; }
    """
    unit = compile_shader(src, "vs")
    assert [(s, i) for s, i, _ in unit.reflection.outputs] == [("TEXCOORD", 0), ("TEXCOORD", 1)]
    out = HostShader(unit).vs([[1, 2, 3, 4]])
    assert np.array_equal(out[0], np.array([1, 2, 3, 4], f32))
    assert np.array_equal(out[1], np.array([2, 4, 6, 8], f32)) and np.array_equal(out[2], np.array([3, 6, 9, 12], f32))
    with pytest.raises(CompileError, match="bound twice"):
        compile_shader("struct O { float4 p: SV_Position; float4 a: TEXCOORD0; float4 b: TEXCOORD(0); };"
                       " O main(float4 p: POSITION) { O o; o.p = p; o.a = p; o.b = p; return o; }", "vs")

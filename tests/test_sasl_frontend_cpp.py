"""CPU suite: the C++ SASL front end (salviarenderer_b200/host/sasl_frontend.hpp - what shader::compile of the C++ surface runs in
process) against the Python one (salviarenderer_b200/sasl/frontend.py - what bench.py and the GPU suite run through sasl/jit.py).
Both must produce the SAME unit text, byte for byte - reflection and generated device code - or both reject the source, over
every shader this repository holds: bench.py's headline pair, the known-answer shader, every string literal of the SASL test
modules that looks like a shader (tried as vs, ps and lib, so the error paths are compared too), and - read in place where
/root/reference exists - the reference's own sasl/test/repo units and sample shaders with the options its tests drive them with."""
import ast
import glob
import os
import subprocess

import pytest

from conftest import ROOT
from salviarenderer_b200.sasl import emit, frontend
from salviarenderer_b200.sasl.preprocess import PreprocessError, preprocess

HOST = os.path.join(ROOT, "salviarenderer_b200", "host")
CLI_SRC = os.path.join(ROOT, "tests", "cpp", "sasl_frontend_cli.cpp")
REF = "/root/reference"


@pytest.fixture(scope="module")
def cli(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("sasl_cli") / "sasl_frontend_cli")
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-D_GLIBCXX_ASSERTIONS", "-I" + HOST, "-o", exe, CLI_SRC], capture_output=True, text=True)  # out-of-range accesses abort
    assert r.returncode == 0, r.stderr[-3000:]
    return exe


def run_cpp(cli, src, stage, entry=None, defines=None, include_dirs=(), sys_include_dirs=(), file_name=None, virtual_files=None, tmp=None):
    cmd = [cli, stage]
    if entry:
        cmd += ["--entry", entry]
    for k, v in (defines or {}).items():
        cmd += ["-D", k if v is None else f"{k}={v}"]
    for d in include_dirs:
        cmd += ["-I", d]
    for d in sys_include_dirs:
        cmd += ["--sys", d]
    if file_name:
        cmd += ["--file", file_name]
    for k, (name, text) in enumerate((virtual_files or {}).items()):
        path = os.path.join(str(tmp), f"virtual_{k}")
        with open(path, "w") as f:
            f.write(text)
        cmd += ["--virtual", name, path]
    r = subprocess.run(cmd, input=src.encode(), capture_output=True)
    assert r.returncode in (0, 2), f"the C++ front end crashed (rc {r.returncode}): {r.stderr.decode()[-500:]}"
    out = r.stdout.decode(errors="replace")
    return (out, None) if r.returncode == 0 else (None, out[len("error\n"):].strip())


def run_py(src, stage, entry=None, **kw):
    try:
        return emit.render(frontend.compile_shader(src, stage, entry, **kw)), None
    except frontend.CompileError as e:
        return None, str(e)


def same(cli, src, stage, tmp, **kw):
    a, ea = run_py(src, stage, **kw)
    b, eb = run_cpp(cli, src, stage, tmp=tmp, **kw)
    if a is None or b is None:
        assert a is None and b is None, f"{stage}: python {'rejects: ' + ea if a is None else 'accepts'}, C++ {'rejects: ' + eb if b is None else 'accepts'}\n{src[:400]}"
        return False
    if a != b:
        la, lb = a.split("\n"), b.split("\n")
        k = next((i for i, (x, y) in enumerate(zip(la, lb)) if x != y), min(len(la), len(lb)))
        raise AssertionError(f"{stage}: units differ at line {k}: {la[k:k + 1]} vs {lb[k:k + 1]}\n{src[:400]}")
    return True


def looks_like_a_shader(s):
    return len(s) > 30 and "(" in s and "{" in s and ("return" in s or "struct" in s)


def literal_sources():
    """Every string literal of the SASL test modules that looks like a shader; str.format templates get their usual fillers."""
    out = []
    for mod in ("test_sasl_frontend.py", "test_gpu_sasl_jit.py", "test_host_surface.py", "sasl_kat.py"):
        tree = ast.parse(open(os.path.join(ROOT, "tests", mod)).read())
        for node in ast.walk(tree):
            if isinstance(node, ast.Constant) and isinstance(node.value, str) and looks_like_a_shader(node.value):
                s = node.value
                if "{decls}" in s:
                    out += [s.format(decls=d) for d in ("float4 uv: TEXCOORD0;", "float4 pad: TEXCOORD0; float4 uv: TEXCOORD1;")]
                else:
                    out.append(s)
    return out


def test_headline_and_known_answer_shaders(cli, tmp_path):
    import bench
    import sasl_kat
    assert same(cli, bench.SASL_VS_SPONZA, "vs", tmp_path)
    assert same(cli, bench.SASL_PS_SPONZA, "ps", tmp_path)
    assert same(cli, sasl_kat.shader_source(), "ps", tmp_path)


def test_every_shader_literal_of_the_test_modules(cli, tmp_path):
    srcs = literal_sources()
    assert len(srcs) >= 35
    accepted = sum(same(cli, s, stage, tmp_path) for s in srcs for stage in ("vs", "ps", "lib"))
    assert accepted >= 40, accepted


def test_language_corner_cases(cli, tmp_path):
    """Hand-picked sources for paths the other corpora touch lightly: literals, casts, swizzle stores, compound assignment,
    increments, run-time indices, switch fall-through with continue, recursion, entry selection, semantic order, error paths."""
    cases = [
        ("ps", "float4 main(float4 p: TEXCOORD0): COLOR { int k = 0x1F; uint m = 0xFFu; float e = 1e-3; float h = 2.5h; float d = .5; float q = 3.;"
               " return p * (float)(k + (int)m) + float4(e, h, d, q) + 7 + 3u + 10L; }"),
        ("ps", "float4 main(float4 p: TEXCOORD1, float2 q: TEXCOORD0): COLOR { p.zw = q; p.x += q.y; p.y *= 2; ++p.x; p.w--; int2 i = int2(1, 2); i <<= 1; i %= 3;"
               " bool2 b = p.xy > q; return b.x && !b.y ? p : -p; }"),
        ("ps", "float3x3 M; int sel; float4 main(float4 p: TEXCOORD0): COLOR { float3 r = M[sel]; float s = p[sel]; float3 c = M[1]; return float4(r + c, s + M._m12 + M._23); }"),
        ("ps", "float4 main(float4 p: TEXCOORD0): COLOR { float a = 0; for (int i = 0; i < 8; ++i) { switch (i) { case 0: case 1: a += 1; case 2: a += 2; break; case -3: continue;"
               " default: a += 0.5f; } if (a > 10) break; } int k = 3; do { a += k; } while (--k > 0); while (a > 100) a -= 100; return p * a; }"),
        ("ps", "float f(float x); float g(float x) { return x > 1 ? f(x - 1) : x; } float f(float x) { return g(x * 0.5f); } float4 main(float4 p: TEXCOORD0): COLOR { return p * f(p.x); }"),
        ("ps", "float helper(float x) { return x * 2; } float4 other(float4 p: TEXCOORD0): COLOR { return p; } float4 shade(float4 p: TEXCOORD0): COLOR { return p * helper(p.y); }"),
        ("vs", "float4x4 wvp; struct O { float4 n: NORMAL; float4 t: TEXCOORD1; float4 pos: SV_Position; float4 u: TEXCOORD0; };"
               " O main(float4 p: POSITION, int4 bi: BLEND_INDICES, float3 n: NORMAL) { O o; o.pos = mul(p, wvp); o.n = float4(n, bi.x); o.t = p; o.u = p.wzyx; return o; }"),
        ("vs", "sampler s; float4 main(float4 p: POSITION): SV_Position { return p + tex2Dlod(s, float4(p.xy, 0, 0)); }"),
        ("ps", "sampler a; sampler b; float4 main(float4 p: TEXCOORD0): COLOR { return tex2D(a, p.xy) + tex2Dproj(b, p) + tex2Dbias(a, p) + tex2Dgrad(b, p.xy, ddx(p.xy), ddy(p.zw)); }"),
        ("ps", "float4 main(float4 p: TEXCOORD0): COLOR { if (p.x > 0) return tex2D(p, p.xy); return p; }"),            # not a sampler
        ("ps", "sampler s; float4 main(float4 p: TEXCOORD0): COLOR { if (p.x > 0) { return tex2D(s, p.xy); } return p; }"),  # divergent
        ("ps", "float4 main(float4 p: TEXCOORD0, float4 q: TEXCOORD0): COLOR { return p + q; }"),                          # bound twice
        ("ps", "float4 main(float4 p: TEXCOORD0): COLOR { return undefined_name; }"),
        ("ps", "float4 main(float4 p: TEXCOORD0): COLOR { return p.xyzq; }"),
        ("ps", "float4 main(float4 p: TEXCOORD0): COLOR { float3 v = p; float2x2 m = float2x2(p); return float4(v, m[1].y) + nosuch(p); }"),
        ("ps", "float4 main(float4 p: TEXCOORD0) COLOR { return p; }"),                                                     # syntax
        ("ps", "float4 main(float4 p: TEXCOORD0): COLOR { return p @ 2; }"),                                                # lexer
        ("vs", "float4 main(float4 p: POSITION): TEXCOORD0 { return p; }"),                                                  # no position
        ("lib", "int f(int a, uint b) { return (a & 3) | (int)(b >> 2) ^ ~a; } bool g(float x) { return isnan(x) || isinf(x) || !isfinite(x); }"
                " uint h(uint v) { return countbits(v) + firstbithigh(v) + firstbitlow(v) + reversebits(v); }"),
    ]
    n_ok = sum(same(cli, src, stage, tmp_path) for stage, src in cases)
    assert n_ok >= 9, n_ok
    assert same(cli, cases[5][1], "ps", tmp_path, entry="other")
    assert not same(cli, cases[5][1], "ps", tmp_path, entry="missing")


PP_CASES = [
    "#define N 4\n#define SQ(x) ((x)*(x))\nfloat a[N]; int b = SQ(N+1);\n#if N > 3 && !defined(Q)\nint yes;\n#elif N\nint maybe;\n"
    "#else\nint no;\n#endif\n#undef N\nint N;\n/* a\n   block */ int c; // N stays\n",
    "#define A B\n#define B A\nA B",
    "#define F(a, b) (a + b)\n#define G() 7\nF(F(1, 2), G()) F (3,\\\n 4) F\n#if defined A || (2 * 3 - 1) % 4 == 1 && ~0 < 0\nyes\n#endif\n",
    "#define S \"a string with F(1) and // no comment\"\nS; 'c' 12.5e3f x.y\n#if 0x10 == 16 && 10u >= 010\nhex\n#else\nnohex\n#endif\n",
    "#ifdef A\n#if 1\na1\n#else\na0\n#endif\n#elif 1\nb\n#else\nc\n#endif\n#pragma once\n#line 7\n#\nend",
    "#if 1\nx", "#endif", "#else", "#include \"nope.ss\"", "#error stop", "#frobnicate", "#define F(a) a\nF(1, 2)", "#if 1 +\n#endif", "#if 1 / 0\n#endif",
]


def test_preprocessor_alone(cli, tmp_path):
    for src in PP_CASES:
        for defines in (None, {"A": None}, {"A": "1", "Q": "x"}):
            try:
                a, ea = preprocess(src, defines), None
            except PreprocessError as e:
                a, ea = None, str(e)
            b, eb = run_cpp(cli, src, "pp", defines=defines, tmp=tmp_path)
            assert (a is None) == (b is None), (src, defines, ea, eb)
            assert a == b, (src, defines)


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")
def test_reference_units_read_in_place(cli, tmp_path):
    """sasl/test/repo/*.{ss,svs,sps} and the samples' shaders: same unit or same rejection; the preprocessor units with the
    virtual file, search paths and defines the reference's driver tests use."""
    files = sorted(glob.glob(REF + "/sasl/test/repo/*.s*") + glob.glob(REF + "/resources/**/*.sa[vp]s", recursive=True))
    assert len(files) >= 49
    accepted = 0
    for f in files:
        ext = f.rsplit(".", 1)[1]
        stage = "vs" if ext in ("svs", "savs") else ("ps" if ext in ("sps", "saps") else "lib")
        accepted += same(cli, open(f, errors="replace").read(), stage, tmp_path, file_name=f)
    assert accepted >= 42, accepted  # the include units need the options below
    repo = REF + "/sasl/test/repo"
    virtual = {"virtual_include.ss": "float virtual_add(float a, float b) { return a + b; }"}
    main = os.path.join(repo, "include_main.ss")
    assert same(cli, open(main).read(), "lib", tmp_path, file_name=main, virtual_files=virtual)
    assert not same(cli, open(main).read(), "lib", tmp_path, file_name=main, virtual_files=virtual, defines={"FAILED_INCLUDE": None})
    sp = os.path.join(repo, "include_search_path.ss")
    assert same(cli, open(sp).read(), "lib", tmp_path, file_name=sp, include_dirs=[os.path.join(repo, "include")], sys_include_dirs=[os.path.join(repo, "sysincl")])
    assert not same(cli, open(os.path.join(repo, "preprocessors.ss")).read(), "lib", tmp_path, defines={"SASL_COMPILER_ERROR": None})


def test_random_programs(cli, tmp_path):
    """Differential fuzz (tests/sasl_fuzz.py): typed random pixel shaders - every operator, constructor, cast, swizzle store,
    intrinsic, control-flow statement - some of them ill-typed on purpose.  10,500 seeds were run when this was written (about half
    accepted, half rejected, identical units and identical messages, the C++ side with libstdc++ assertions on); the suite keeps 300."""
    import sasl_fuzz
    accepted = sum(same(cli, sasl_fuzz.program_from_seed(seed, sloppy=0.02 if seed % 2 else 0.0), "ps", tmp_path) for seed in range(300))
    assert 100 < accepted < 300, accepted


def test_c_abi_translate_equals_the_python_front_end(built):
    """slv_sasl_translate (include/salvia_b200.h): the C++ front end inside the product library, callable without a device -
    the first half of the reference's compile(code, profile); the unit text equals the Python front end's."""
    import bench
    from salviarenderer_b200.sasl import jit
    for stage, src in (("vs", bench.SASL_VS_SPONZA), ("ps", bench.SASL_PS_SPONZA)):
        assert jit.translate_in_library(src, stage) == emit.render(frontend.compile_shader(src, stage))
    two = "float4 a(float4 p: TEXCOORD0): COLOR { return p; } float4 b(float4 p: TEXCOORD0): COLOR { return p * 2.0f; }"
    assert jit.translate_in_library(two, "ps", entry="a") == emit.render(frontend.compile_shader(two, "ps", "a"))
    with pytest.raises(frontend.CompileError, match="line 1"):
        jit.translate_in_library("float4 broken(", "ps")


REJECTED = [
    "float3x3 N; float4 main(float4 p: TEXCOORD0): COLOR { return p * N._m33; }",
    "float4 main(float4 p: TEXCOORD0): COLOR { return p[1.5]; }",
    "float4 main(float4 p: TEXCOORD0): COLOR { return p[010]; }",
    "float4 main(float4 p: TEXCOORD0): COLOR { return cross(p); }",
    "float4 main(float4 p: TEXCOORD0): COLOR { return length(); }",
    "float4 main(float4 p: TEXCOORD0): COLOR { return normalize(); }",
    "float4 main(float4 p: TEXCOORD0): COLOR { return transpose(); }",
    "float4 main(float4 p: TEXCOORD0): COLOR { return asfloat(); }",
    "float4 main(float4 p: TEXCOORD0): COLOR { return lit(p, p); }",
    "float4 main(float4 p: TEXCOORD0): COLOR { return dst(p); }",
    "float4 main(float4 p: TEXCOORD0): COLOR { return any(); }",
    "float4 main(float4 p: TEXCOORD0): COLOR { return reflect(p); }",
    "float4 main(float4 p: TEXCOORD0): COLOR { return distance(p); }",
    "float4 main(float4 p: TEXCOORD0): COLOR { return count_bits(); }",
    "float4 main(float4 p: TEXCOORD0): COLOR { return faceforward(p, p); }",
    "struct S { float a; }; float4 main(float4 p: TEXCOORD0): COLOR { S s; return abs(s); }",
    "float4 main(float4 p: TEXCOORD0: COLOR { return p; }",
    "float4 main(float4 p: TEXCOORD(1.5)): COLOR { return p; }",
    "float4 main(float4 p: A1B): COLOR { return p; }",
    "float4 main(float4 p: TEXCOORD0): COLOR { switch (1) { case 1.5: return p; } return p; }",
    "float4 main(float4 p: TEXCOORD0): COLOR { float x[3]; return p; }",
    "float4 main(float4 p: TEXCOORD0): COLOR { return p; ",
    "float4 main(float4 p: TEXCOORD0): COLOR { return p.xyzwx; }",
    "float4 main(float4 p: TEXCOORD0): COLOR { float2x2 m = float2x2(1,2,3,4); return float4(m[5], 0, 0); }",
    "",
    "float4 main(float4 p: TEXCOORD0): COLOR { return 0xFFFFFFFFFFFFFFFFFFFF + p; }",
    "float4 main(float4 p: TEXCOORD0): COLOR { break; return p; }",
    "float4 main(float4 p: TEXCOORD0): COLOR { p = 1 = 2; return p; }",
    "void f() { return 1; } float4 main(float4 p: TEXCOORD0): COLOR { f(); return p; }",
    "struct G { float a; }; G g; float4 main(float4 p: TEXCOORD0): COLOR { return p; }",
    "bool flags[4]; float4 main(float4 p: TEXCOORD0): COLOR { return p; }",
    "float4 b[n]; float4 main(float4 p: TEXCOORD0): COLOR { return b[0]; }",
    "float4 main(float4 p: TEXCOORD0): COLOR { return int2(1, 0) ? p : p * 2.0f; }",   # condition narrower than the operands
    "sampler s; float4 main(float4 p: TEXCOORD0): COLOR { return s ? p : p; }",
]


def test_rejections_carry_the_same_message(cli, tmp_path):
    """Ill-formed sources - wrong argument counts, bad indices, semantics, literals, statements - are rejected by both front
    ends with the SAME diagnostic (a CompileError, never a stray Python exception or a C++ crash)."""
    for src in REJECTED:
        a, ea = run_py(src, "ps")
        b, eb = run_cpp(cli, src, "ps", tmp=tmp_path)
        assert a is None and b is None, src
        assert ea == eb, (src, ea, eb)


def test_front_end_copies_coexist_in_one_process(built):
    """The product library and the CPU checker each carry a copy of the header-only front end (slv_sasl_translate) and are
    loaded into one process by every test.  Called alternately they must give the same unit - in a child process: a regression
    here is a crash (the checker once shared function-local statics and half a C++ runtime with the other module)."""
    import sys
    import textwrap
    from conftest import ORACLE_LIB, PRODUCT_LIB
    code = textwrap.dedent(f"""
        import ctypes as C
        libs = [C.CDLL({PRODUCT_LIB!r}), C.CDLL({ORACLE_LIB!r})]
        out = []
        for lib in libs * 3:
            lib.slv_sasl_translate.argtypes = [C.c_uint32, C.c_char_p, C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.c_char_p, C.c_size_t]
            u, n, log = C.c_void_p(), C.c_size_t(), C.create_string_buffer(4096)
            rc = lib.slv_sasl_translate(1, b"float4 main(float4 p: TEXCOORD0): COLOR {{ return p * 2 + sin(p); }}", None, C.byref(u), C.byref(n), log, 4096)
            assert rc == 0, log.value
            out.append(C.string_at(u.value, n.value))
            bad = lib.slv_sasl_translate(1, b"float4 broken(", None, C.byref(u), C.byref(n), log, 4096)
            assert bad == 1 and b"line 1" in log.value
        assert len(set(out)) == 1 and out[0].startswith(b"SLVSASL 1")
        print("ok")
    """)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == "ok", (r.returncode, r.stderr[-2000:])


def test_front_end_in_an_executable_next_to_a_library_copy(built, tmp_path):
    """tests/cpp/frontend_coexist_test.cpp: a C++ host that includes sasl_frontend.hpp (shader::compile in process) and loads a
    library with its own copy - the CUDA product, the CPU checker - calls both alternately; same units, same diagnostics."""
    from conftest import ORACLE_LIB, PRODUCT_LIB
    exe = str(tmp_path / "frontend_coexist_test")
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), "-I" + HOST,
                        os.path.join(ROOT, "tests", "cpp", "frontend_coexist_test.cpp"), "-o", exe, "-ldl"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    for lib in (PRODUCT_LIB, ORACLE_LIB):
        out = subprocess.run([exe, lib], capture_output=True, text=True, timeout=120)
        assert out.returncode == 0 and out.stdout.strip() == "ok", (lib, out.returncode, out.stdout, out.stderr)


def test_generated_code_of_random_programs_is_valid_cxx(tmp_path):
    """What the front end accepts must come out as code the compilers accept: the accepted random programs, compiled for the host
    (syntax only; 300 seeds compiled when this was written, and six of them through NVRTC for sm_100a, all fine; the suite keeps 40)."""
    import sasl_fuzz
    from sasl_host import HARNESS_PS, RT_DIR
    n = 0
    for seed in range(40):
        try:
            unit = frontend.compile_shader(sasl_fuzz.program_from_seed(seed, sloppy=0.0 if seed % 2 == 0 else 0.02), "ps")
        except frontend.CompileError:
            continue
        src = tmp_path / f"s{seed}.cpp"
        src.write_text('#include "sasl_rt.h"\n' + unit.code + HARNESS_PS)
        r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-w", "-I" + RT_DIR, str(src)], capture_output=True, text=True)
        assert r.returncode == 0, (seed, r.stderr[:1500])
        n += 1
    assert n >= 15, n

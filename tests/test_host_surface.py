"""The C++ host surface (salviarenderer_b200/host/salvia_b200_renderer.hpp — the mirror of salvia::core::renderer over the C
ABI): tests/cpp/host_surface_test.cpp renders through it and prints buffer hashes and counters.  The SAME binary is run
against different libraries and must print the same lines: CPU checkers here, the CUDA product on the GPU box."""
import os
import subprocess
import sys

import pytest

from conftest import ORACLE_LIB, PRODUCT_LIB, REF_LIB, ROOT


@pytest.fixture(scope="module")
def host_binary(built, tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("host") / "host_surface_test")
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"),
                    "-I" + os.path.join(ROOT, "salviarenderer_b200", "host"), os.path.join(ROOT, "tests", "cpp", "host_surface_test.cpp"),
                    "-o", exe, "-ldl"], check=True)
    return exe


def run(exe, lib, *size):
    # the SASL section's compile() runs the C++ front end in process - on the CUDA product only; the interpreter settings only
    # matter when SLV_SASL_FRONTEND=python selects the Python twin
    env = dict(os.environ, SLV_SASL_PYTHON=sys.executable, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    out = subprocess.run([exe, lib, *map(str, size)], capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stdout + out.stderr
    lines = out.stdout.strip().splitlines()
    return lines[0], lines[1:]


@pytest.mark.parametrize("size", [(320, 240, 4), (200, 120, 1)])
def test_cpp_surface_oracle_equals_reference(host_binary, size):
    name_o, out_o = run(host_binary, ORACLE_LIB, *size)
    assert name_o == "backend oracle"
    if not os.path.exists(REF_LIB):
        pytest.skip("reference library not built here")
    name_r, out_r = run(host_binary, REF_LIB, *size)
    assert name_r == "backend reference"
    assert out_o == out_r


@pytest.mark.gpu
@pytest.mark.parametrize("size", [(320, 240, 4), (1000, 600, 2), (200, 120, 1)])
def test_cpp_surface_product_equals_checker(host_binary, size):
    name_p, out_p = run(host_binary, PRODUCT_LIB, *size)
    assert name_p == "backend cuda-sm100a"
    checker = REF_LIB if os.path.exists(REF_LIB) else ORACLE_LIB
    _, out_c = run(host_binary, checker, *size)
    assert out_p == out_c


@pytest.mark.gpu
def test_cpp_surface_sasl_equals_twins_on_the_product(host_binary):
    """The SASL section (compile() -> set_*_shader_code, NVRTC in process) against the same section with the cpp twins, both
    on the CUDA product: identical line."""
    _, out_sasl = run(host_binary, PRODUCT_LIB, 320, 240, 4)
    os.environ["SLV_HOST_TEST_NO_SASL"] = "1"
    try:
        _, out_twin = run(host_binary, PRODUCT_LIB, 320, 240, 4)
    finally:
        del os.environ["SLV_HOST_TEST_NO_SASL"]
    assert out_sasl[-1].startswith("sasl color") and out_sasl == out_twin


def test_cpp_compile_runs_the_front_end_and_parses_its_unit(tmp_path):
    """shader::compile() (renderer.h:136-147 on the C++ surface) needs no device and no interpreter: it runs the C++ SASL front
    end in process (host/sasl_frontend.hpp) - or, with SLV_SASL_FRONTEND=python, the Python twin as a child process - and
    parses the unit.  Either way the reflection the C++ side ends up with equals what the Python front end reports; a compile
    error comes back as a null object with the front end's message."""
    from salviarenderer_b200.sasl import compile_shader
    from test_sasl_frontend import VS_SKIN
    exe = str(tmp_path / "sasl_compile_test")
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"),
                    "-I" + os.path.join(ROOT, "salviarenderer_b200", "host"), os.path.join(ROOT, "tests", "cpp", "sasl_compile_test.cpp"),
                    "-o", exe, "-ldl"], check=True)
    in_process = {k: v for k, v in os.environ.items() if k not in ("SLV_SASL_FRONTEND", "SLV_SASL_PYTHON", "PYTHONPATH")}
    in_process["SLV_SASL_PYTHON"] = "/nonexistent/python"  # the default path must not need an interpreter
    child = dict(os.environ, SLV_SASL_FRONTEND="python", SLV_SASL_PYTHON=sys.executable, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    ps = """sampler texSamp; float4 tint; float gain;
    struct PSIn { float4 uv: TEXCOORD0; float4 n: TEXCOORD1; };
    float4 ps_main(PSIn in): COLOR { return tex2D(texSamp, in.uv.xy) * tint * gain + ddx(in.n); }"""
    for env, stage, src in [(e, st, sr) for e in (in_process, child) for st, sr in (("vs", VS_SKIN), ("ps", ps))]:
        out = subprocess.run([exe, stage], input=src, capture_output=True, text=True, timeout=120, env=env)
        assert out.returncode == 0, out.stdout + out.stderr
        got = out.stdout.strip().splitlines()
        r = compile_shader(src, stage)
        want = [f"n_vs_output_attrs {r.reflection.n_vs_output_attrs}", f"uniform_bytes {r.reflection.uniform_bytes}",
                f"uses_derivatives {int(r.reflection.uses_derivatives)}"]
        want += sorted(f"uniform {n} {t.replace(' ', '')} {o} {s}" for n, t, o, s in r.reflection.uniforms)  # std::map: sorted by name
        want += [f"sampler {i} {n}" for i, n in enumerate(r.reflection.samplers)]
        want += [f"input {s} {i} {k}" for k, (s, i, _) in enumerate(r.reflection.inputs)]
        want += [f"output {s} {i} {k}" for k, (s, i, _) in enumerate(r.reflection.outputs)]
        want += [f"code_bytes {len(r.code.encode())}"]
        assert got == want, (got, want)
    msgs = []
    for env in (in_process, child):
        bad = subprocess.run([exe, "ps"], input="float4 broken(", capture_output=True, text=True, timeout=120, env=env)
        assert bad.returncode == 2 and bad.stdout.startswith("error\n") and "line 1" in bad.stdout
        msgs.append(bad.stdout)
    assert msgs[0] == msgs[1]

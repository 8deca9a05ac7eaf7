"""CPU suite: run-time shader compilation in process (slv_shader_compile_cubin: NVRTC over the sources embedded in the product
library; include/salvia_b200.h) and the front end as a command (salviarenderer_b200/sasl/emit.py, what the C++ host surface's
compile() runs).  Compiling needs no GPU; loading the image does (tests/test_gpu_sasl_jit.py, tests/test_host_surface.py)."""
import ctypes as C
import os
import shutil
import subprocess
import sys
import tempfile

import pytest

from conftest import PRODUCT_LIB, ROOT
from salviarenderer_b200.sasl import frontend, jit

VS = """
float4x4 wvp; float4 tint;
struct VSIn  { float4 pos: POSITION; float4 uv: TEXCOORD0; };
struct VSOut { float4 pos: sv_position; float4 uv: TEXCOORD0; float4 col: TEXCOORD1; };
VSOut vs_main(VSIn in) { VSOut o; o.pos = mul(in.pos, wvp); o.uv = in.uv; o.col = tint * in.uv.x; return o; }
"""


def _nvrtc_available():
    return any(os.path.exists(p) for p in ("/usr/local/cuda/lib64/libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so"))


@pytest.fixture(scope="module")
def lib(built):
    L = C.CDLL(PRODUCT_LIB)
    L.slv_shader_compile_cubin.restype = C.c_int32
    L.slv_shader_compile_cubin.argtypes = [C.c_uint32, C.c_char_p, C.c_uint32, C.c_uint32, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.c_char_p, C.c_size_t]
    L.slv_free.restype = None
    L.slv_free.argtypes = [C.c_void_p]
    return L


@pytest.mark.skipif(not _nvrtc_available(), reason="libnvrtc not installed")
def test_compile_in_process_produces_the_pipeline_kernels(lib, tmp_path, monkeypatch):
    monkeypatch.setenv("SLV_JIT_CACHE", str(tmp_path / "cache"))
    unit = frontend.compile_shader(VS, "vs")
    img, n, log = C.c_void_p(), C.c_size_t(), C.create_string_buffer(8192)
    rc = lib.slv_shader_compile_cubin(0, unit.code.encode(), unit.reflection.n_vs_output_attrs, 0, C.byref(img), C.byref(n), log, len(log))
    assert rc == 0, log.value.decode()
    cubin = C.string_at(img.value, n.value)
    lib.slv_free(img)
    assert cubin[:4] == b"\x7fELF"
    for name in jit.VS_ENTRY_POINTS:
        assert name.encode() in cubin
    # second call: served from the cache directory (mode 0700, file 0600), same bytes
    cached = [f for f in os.listdir(tmp_path / "cache") if f.endswith(".cubin")]
    assert len(cached) == 1 and (os.stat(tmp_path / "cache" / cached[0]).st_mode & 0o077) == 0
    rc = lib.slv_shader_compile_cubin(0, unit.code.encode(), unit.reflection.n_vs_output_attrs, 0, C.byref(img), C.byref(n), log, len(log))
    assert rc == 0 and C.string_at(img.value, n.value) == cubin
    lib.slv_free(img)
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if os.path.exists(cuobjdump):  # sm_100a code, the numerics contract (no fused packed multiply-add)
        with tempfile.NamedTemporaryFile(suffix=".cubin") as f:
            f.write(cubin)
            f.flush()
            sass = subprocess.run([cuobjdump, "-sass", f.name], capture_output=True, text=True).stdout
        assert "sm_100a" in sass and "FFMA2" not in sass


@pytest.mark.skipif(not _nvrtc_available(), reason="libnvrtc not installed")
def test_compile_errors_come_back_in_the_log(lib, tmp_path, monkeypatch):
    monkeypatch.setenv("SLV_JIT_CACHE", str(tmp_path / "cache"))
    img, n, log = C.c_void_p(), C.c_size_t(), C.create_string_buffer(8192)
    rc = lib.slv_shader_compile_cubin(1, b"this is not device code", 0, 0, C.byref(img), C.byref(n), log, len(log))
    assert rc == 1 and not img.value and b"error" in log.value  # SLV_FAILED
    assert lib.slv_shader_compile_cubin(7, b"", 0, 0, C.byref(img), C.byref(n), log, len(log)) == 3  # SLV_INVALID_PARAMETER: stage
    assert lib.slv_shader_compile_cubin(0, b"", 99, 0, C.byref(img), C.byref(n), log, len(log)) == 3  # too many outputs
    assert not os.path.exists(tmp_path / "cache") or not os.listdir(tmp_path / "cache")  # nothing cached for a failed compile


def test_cache_directory_somebody_else_can_write_is_not_used(lib, tmp_path, monkeypatch):
    """An image is code that runs in the process's GPU context: a group/world-writable cache directory is ignored (nothing read
    from it, nothing written to it)."""
    d = tmp_path / "shared"
    d.mkdir()
    os.chmod(d, 0o777)
    monkeypatch.setenv("SLV_JIT_CACHE", str(d))
    img, n, log = C.c_void_p(), C.c_size_t(), C.create_string_buffer(4096)
    lib.slv_shader_compile_cubin(1, b"not code", 0, 0, C.byref(img), C.byref(n), log, len(log))
    assert not os.listdir(d)


def test_front_end_command(tmp_path):
    env = dict(os.environ, PYTHONPATH=ROOT)
    out = subprocess.run([sys.executable, "-m", "salviarenderer_b200.sasl.emit", "vs"], input=VS, capture_output=True, text=True, env=env)
    assert out.returncode == 0, out.stderr
    head, code = out.stdout.split("\ncode ", 1)
    lines = head.splitlines()
    assert lines[0] == "SLVSASL 1" and "stage vs" in lines and "n_vs_output_attrs 2" in lines
    assert "uniform wvp float4x4 0 64" in lines and "uniform tint float4 64 16" in lines
    assert "input POSITION 0 0" in lines and "input TEXCOORD 0 1" in lines and "output TEXCOORD 1 1" in lines
    nbytes, body = code.split("\n", 1)
    assert len(body.encode()) == int(nbytes) and "slv_jit_vs" in body
    bad = subprocess.run([sys.executable, "-m", "salviarenderer_b200.sasl.emit", "ps"], input="float4 broken(", capture_output=True, text=True, env=env)
    assert bad.returncode == 2 and bad.stdout.startswith("error\n")

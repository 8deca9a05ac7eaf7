"""SASL shader -> sm_100a cubin -> shader module of the C ABI.

compile(): frontend.compile_shader -> generated device code -> slv_shader_compile_cubin of the product library: NVRTC IN
PROCESS over the library's own embedded pipeline-kernel sources (NVVM / NVPTX -> PTX -> SASS for sm_100a, the pipeline kernels
and the shader in one translation unit) -> cubin bytes, cached on disk by content hash.  No GPU is needed to compile.
SLV_JIT_COMPILER=nvcc selects the former path, an `nvcc -cubin` subprocess over csrc/slv_jit_unit.cu (same flags, same code).
load(): hands the cubin to the library (slv_shader_module_load); draws then select it with abi.program_jit(module).
Needs the CUDA toolkit's libnvrtc (or nvcc) at run time, as the reference needs LLVM at run time.
"""
from __future__ import annotations

import ctypes
import hashlib
import os
import shutil
import subprocess
import tempfile
from dataclasses import dataclass

from . import frontend

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(os.path.dirname(HERE), "csrc")
INCLUDE = os.path.join(os.path.dirname(os.path.dirname(HERE)), "include")
NVCC_FLAGS = ["-cubin", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false"]


@dataclass
class CompiledShader:
    unit: frontend.ShaderUnit
    cubin: bytes                 # b"" when compiled with device=False (a shader only ever loaded into a CPU checker)
    ptx_entry_points: tuple
    derivatives: str = "sasl"

    @property
    def reflection(self):
        return self.unit.reflection


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def cache_dir():
    """Per-user cache of compiled cubins ($SLV_JIT_CACHE, else $XDG_CACHE_HOME/salvia_b200_jit, else ~/.cache/salvia_b200_jit),
    mode 0700.  A cubin is code that runs in this process's GPU context, so a directory somebody else owns or can write to is
    never trusted: compilation then goes to a fresh private directory and nothing is cached."""
    d = os.environ.get("SLV_JIT_CACHE") or os.path.join(
        os.environ.get("XDG_CACHE_HOME") or os.path.join(os.path.expanduser("~"), ".cache"), "salvia_b200_jit")
    try:
        os.makedirs(d, mode=0o700, exist_ok=True)
        st = os.stat(d)
        if st.st_uid != os.getuid() or (st.st_mode & 0o022):
            raise PermissionError(d)
    except OSError:
        d = tempfile.mkdtemp(prefix="salvia_b200_jit_")  # 0700, ours
    return d


def _trusted(path: str) -> bool:
    try:
        st = os.stat(path)
    except OSError:
        return False
    return st.st_uid == os.getuid() and not (st.st_mode & 0o022)


PRODUCT_LIB = os.path.join(CSRC, "libsalvia_b200.so")
VS_ENTRY_POINTS = ("slv_jit_k_geometry", "slv_jit_k_vertex_shade")
PS_ENTRY_POINTS = ("slv_jit_k_raster_s1", "slv_jit_k_raster_s2", "slv_jit_k_raster_s4", "slv_jit_k_shade_s1", "slv_jit_k_shade_s2", "slv_jit_k_shade_s4")
_lib = None


def _product_lib():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(PRODUCT_LIB)
        lib.slv_shader_compile_cubin.restype = ctypes.c_int32
        lib.slv_shader_compile_cubin.argtypes = [ctypes.c_uint32, ctypes.c_char_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.POINTER(ctypes.c_void_p),
                                                 ctypes.POINTER(ctypes.c_size_t), ctypes.c_char_p, ctypes.c_size_t]
        lib.slv_free.restype = None
        lib.slv_free.argtypes = [ctypes.c_void_p]
        lib.slv_sasl_translate.restype = ctypes.c_int32
        lib.slv_sasl_translate.argtypes = [ctypes.c_uint32, ctypes.c_char_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_size_t),
                                           ctypes.c_char_p, ctypes.c_size_t]
        _lib = lib
    return _lib


def translate_in_library(source: str, stage: str, entry: str | None = None) -> str:
    """slv_sasl_translate (include/salvia_b200.h): the C++ front end inside the product library; returns the unit text
    (`SLVSASL 1 ... code NBYTES` + device code) that sasl/emit.py renders from the Python front end - the two are identical."""
    lib = _product_lib()
    unit, n = ctypes.c_void_p(), ctypes.c_size_t()
    log = ctypes.create_string_buffer(1 << 14)
    rc = lib.slv_sasl_translate(0 if stage == "vs" else 1, source.encode(), entry.encode() if entry else None, ctypes.byref(unit), ctypes.byref(n), log, len(log))
    if rc != 0:
        raise frontend.CompileError(log.value.decode(errors="replace"))
    try:
        return ctypes.string_at(unit.value, n.value).decode()
    finally:
        lib.slv_free(unit)


def compile_in_process(unit: frontend.ShaderUnit, derivatives: str = "sasl") -> bytes:
    """slv_shader_compile_cubin (include/salvia_b200.h): NVRTC inside the product library, its disk cache included."""
    lib = _product_lib()
    image, n = ctypes.c_void_p(), ctypes.c_size_t()
    log = ctypes.create_string_buffer(1 << 16)
    rc = lib.slv_shader_compile_cubin(0 if unit.stage == "vs" else 1, unit.code.encode(), unit.reflection.n_vs_output_attrs,
                                      1 if (unit.stage == "ps" and derivatives == "cpp") else 0, ctypes.byref(image), ctypes.byref(n), log, len(log))
    if rc != 0:
        raise frontend.CompileError("device compilation of the generated code failed (slv_shader_compile_cubin):\n" + log.value.decode(errors="replace")[-4000:])
    try:
        return ctypes.string_at(image.value, n.value)
    finally:
        lib.slv_free(image)


def compile(source: str, stage: str, entry: str | None = None, derivatives: str = "sasl", keep_dir: str | None = None,  # noqa: A001
            device: bool = True) -> CompiledShader:
    """stage 'vs' | 'ps'.  derivatives: 'sasl' (per row / per column) or 'cpp' (q1 - q0 / q2 - q0 for the whole quad).
    keep_dir (nvcc path only): keep the generated file there.  device=False: the front end only, no sm_100a image - for a
    shader that is only loaded into a CPU checker (tests), whose slv_shader_compile builds the generated code for the host."""
    unit = frontend.compile_shader(source, stage, entry)
    if not device:
        return CompiledShader(unit, b"", (), derivatives)
    if os.environ.get("SLV_JIT_COMPILER", "nvrtc") != "nvcc" and keep_dir is None and os.path.exists(PRODUCT_LIB):
        return CompiledShader(unit, compile_in_process(unit, derivatives), VS_ENTRY_POINTS if stage == "vs" else PS_ENTRY_POINTS, derivatives)
    defs = ["-DSLV_JIT_VS=1", f"-DSLV_JIT_R={unit.reflection.n_vs_output_attrs + 1}"] if stage == "vs" else ["-DSLV_JIT_PS=1"]
    if stage == "ps" and derivatives == "cpp":
        defs.append("-DSLV_JIT_DERIV_CPP=1")
    deps = b"".join(open(os.path.join(d, f), "rb").read() for d, f in
                    ((CSRC, "slv_jit_unit.cu"), (CSRC, "slv_kernels.cuh"), (CSRC, "slv_deferred.cuh"), (CSRC, "slv_sampler.cuh"),
                     (CSRC, "slv_common.cuh"),
                     (HERE, "sasl_rt.h"), (INCLUDE, "salvia_b200.h")))
    key = hashlib.sha256(unit.code.encode() + b"\0" + " ".join(defs + NVCC_FLAGS).encode() + b"\0" + deps).hexdigest()[:24]
    cdir = cache_dir()
    path = os.path.join(cdir, key + ".cubin")
    if not _trusted(path):
        work = keep_dir or tempfile.mkdtemp(prefix="slvjit_")
        try:
            gen = os.path.join(work, f"generated_{key}.cuh")
            with open(gen, "w") as f:
                f.write(unit.code)
            # unique name inside the cache directory, then an atomic rename: concurrent compiles of one shader (the ranks of a
            # multi-GPU run) never see a partial file
            fd, tmp = tempfile.mkstemp(prefix=key + ".", suffix=".tmp", dir=cdir)
            os.close(fd)
            cmd = [_nvcc(), *NVCC_FLAGS, *defs, f'-DSLV_JIT_GENERATED="{gen}"', "-I" + INCLUDE, "-I" + CSRC, "-I" + HERE,
                   "-o", tmp, os.path.join(CSRC, "slv_jit_unit.cu")]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                os.unlink(tmp)
                raise frontend.CompileError("device compilation of the generated code failed:\n" + r.stderr[-4000:])
            os.chmod(tmp, 0o600)
            os.replace(tmp, path)
        finally:
            if keep_dir is None:
                shutil.rmtree(work, ignore_errors=True)
    return CompiledShader(unit, open(path, "rb").read(), VS_ENTRY_POINTS if stage == "vs" else PS_ENTRY_POINTS, derivatives)


def load(be, shader: CompiledShader) -> int:
    """Registers the compiled shader with the library; returns the module handle for abi.program_jit().  The CUDA product takes
    the sm_100a image; a CPU checker (tests) takes the generated code through slv_shader_compile and builds it for the host."""
    if be.name != "cuda-sm100a":
        return be.shader_compile(shader.unit.stage, shader.unit.code, shader.reflection.n_vs_output_attrs, deriv_cpp=shader.derivatives == "cpp")
    if not shader.cubin:
        raise frontend.CompileError("this shader was compiled with device=False: there is no sm_100a image to load")
    return be.shader_module_load(shader.unit.stage, shader.cubin, shader.reflection.n_vs_output_attrs)

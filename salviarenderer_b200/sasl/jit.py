"""SASL shader -> sm_100a cubin -> shader module of the C ABI.

compile(): frontend.compile_shader -> generated .cuh -> `nvcc -cubin` of csrc/slv_jit_unit.cu (NVVM / NVPTX -> PTX ->
SASS for sm_100a, the pipeline kernels and the shader in one translation unit) -> cubin bytes, cached on disk by content
hash.  load(): hands the cubin to the library (slv_shader_module_load); draws then select it with
abi.program_jit(module).  Needs the CUDA toolkit's nvcc at run time (as the reference needs LLVM at run time).
"""
from __future__ import annotations

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
    cubin: bytes
    ptx_entry_points: tuple

    @property
    def reflection(self):
        return self.unit.reflection


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def cache_dir():
    d = os.environ.get("SLV_JIT_CACHE") or os.path.join(tempfile.gettempdir(), "salvia_b200_jit")
    os.makedirs(d, exist_ok=True)
    return d


def compile(source: str, stage: str, entry: str | None = None, derivatives: str = "sasl", keep_dir: str | None = None) -> CompiledShader:  # noqa: A001
    """stage 'vs' | 'ps'.  derivatives: 'sasl' (per row / per column) or 'cpp' (q1 - q0 / q2 - q0 for the whole quad)."""
    unit = frontend.compile_shader(source, stage, entry)
    defs = ["-DSLV_JIT_VS=1", f"-DSLV_JIT_R={unit.reflection.n_vs_output_attrs + 1}"] if stage == "vs" else ["-DSLV_JIT_PS=1"]
    if stage == "ps" and derivatives == "cpp":
        defs.append("-DSLV_JIT_DERIV_CPP=1")
    deps = b"".join(open(os.path.join(d, f), "rb").read() for d, f in
                    ((CSRC, "slv_jit_unit.cu"), (CSRC, "slv_kernels.cuh"), (CSRC, "slv_deferred.cuh"), (CSRC, "slv_sampler.cuh"),
                     (CSRC, "slv_common.cuh"),
                     (HERE, "sasl_rt.h"), (INCLUDE, "salvia_b200.h")))
    key = hashlib.sha256(unit.code.encode() + b"\0" + " ".join(defs + NVCC_FLAGS).encode() + b"\0" + deps).hexdigest()[:24]
    path = os.path.join(cache_dir(), key + ".cubin")
    if not os.path.exists(path):
        work = keep_dir or tempfile.mkdtemp(prefix="slvjit_")
        try:
            gen = os.path.join(work, "generated.cuh")
            with open(gen, "w") as f:
                f.write(unit.code)
            tmp = os.path.join(work, "out.cubin")
            cmd = [_nvcc(), *NVCC_FLAGS, *defs, f'-DSLV_JIT_GENERATED="{gen}"', "-I" + INCLUDE, "-I" + CSRC, "-I" + HERE,
                   "-o", tmp, os.path.join(CSRC, "slv_jit_unit.cu")]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise frontend.CompileError("device compilation of the generated code failed:\n" + r.stderr[-4000:])
            os.replace(tmp, path)
        finally:
            if keep_dir is None:
                shutil.rmtree(work, ignore_errors=True)
    names = ("slv_jit_k_geometry",) if stage == "vs" else (
        "slv_jit_k_raster_s1", "slv_jit_k_raster_s2", "slv_jit_k_raster_s4", "slv_jit_k_shade_s1", "slv_jit_k_shade_s2", "slv_jit_k_shade_s4")
    return CompiledShader(unit, open(path, "rb").read(), names)


def load(be, shader: CompiledShader) -> int:
    """Registers the compiled shader with the library; returns the module handle for abi.program_jit()."""
    return be.shader_module_load(shader.unit.stage, shader.cubin, shader.reflection.n_vs_output_attrs)

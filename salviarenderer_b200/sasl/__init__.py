"""SASL (salvia's HLSL-like shading language) -> sm_100a.

The reference compiles SASL with its own front end and LLVM's MCJIT to host code (sasl/src/drivers/compiler_impl.cpp:
124-302, sasl/src/codegen/*).  Here a new front end (frontend.py: lexer, parser, semantic analysis, reflection) lowers
the subset the samples and the reference's shader tests use to scalar straight-line device code, which jit.py compiles
with the CUDA toolchain (NVVM's NVPTX backend -> PTX -> SASS for sm_100a) TOGETHER with the pipeline kernels, so the
shader is inlined into k_geometry / k_raster exactly like the built-in device programs.  The resulting cubin is loaded
through the C ABI (slv_shader_module_load) and selected per draw with SLV_PROGRAM_JIT(module).
"""
from .frontend import CompileError, Reflection, ShaderUnit, compile_shader  # noqa: F401

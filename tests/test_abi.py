"""The C-ABI boundary: every symbol include/salvia_b200.h declares is exported by every library (no compute
calls on the product here — there may be no GPU)."""
import ctypes
import os
import re

import pytest

from conftest import ORACLE_LIB, PRODUCT_LIB, REF_LIB, ROOT
from salviarenderer_b200 import abi


def header_symbols():
    src = open(os.path.join(ROOT, "include", "salvia_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(slv_[a-z0-9_]+)\s*\(", src)))


def test_binding_lists_every_header_symbol():
    assert header_symbols() == sorted(abi.ENTRY_POINTS)


@pytest.mark.parametrize("which", ["product", "oracle", "reference"])
def test_library_exports_whole_abi(built, which):
    path = {"product": PRODUCT_LIB, "oracle": ORACLE_LIB, "reference": REF_LIB}[which]
    if which == "reference" and not os.path.exists(path):
        pytest.skip("reference library not built here")
    lib = ctypes.CDLL(path)
    missing = [n for n in header_symbols() if not hasattr(lib, n)]
    assert not missing, f"{path} lacks {missing}"
    lib.slv_backend_name.restype = ctypes.c_char_p
    assert lib.slv_backend_name().decode() == {"product": "cuda-sm100a", "oracle": "oracle", "reference": "reference"}[which]
    assert lib.slv_abi_version() == 1


def test_struct_sizes_match_header(built):
    """ctypes mirrors must have the C layout: compile a tiny probe with the real header."""
    import subprocess
    import tempfile
    code = r'''
#include "salvia_b200.h"
#include <stdio.h>
int main(){ printf("%zu %zu %zu %zu %zu %zu\n", sizeof(slv_draw_desc), sizeof(slv_sampler_desc), sizeof(slv_shader_binding),
  sizeof(slv_depth_stencil_desc), sizeof(slv_pipeline_statistics), sizeof(slv_input_element)); return 0; }
'''
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "p.c")
        open(src, "w").write(code)
        exe = os.path.join(td, "p")
        subprocess.run(["gcc", "-I" + os.path.join(ROOT, "include"), src, "-o", exe], check=True)
        out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout.split()
    got = [ctypes.sizeof(t) for t in (abi.DrawDesc, abi.SamplerDesc, abi.ShaderBinding, abi.DepthStencilDesc,
                                      abi.PipelineStatistics, abi.InputElement)]
    assert [int(v) for v in out] == got


def test_product_has_no_cpu_fallback(built):
    """Without a CUDA device the product must refuse to create a device (and the package must raise)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import salviarenderer_b200 as pkg
    with pytest.raises(abi.SlvError):
        pkg.load(0)


def test_package_does_not_reference_oracle():
    """The product path may not import, link or execute anything under oracle/."""
    pkg_dir = os.path.join(ROOT, "salviarenderer_b200")
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                for needle in ("libsalvia_oracle", "libsalvia_ref", "oracle/", "salvia_oracle"):
                    hits = [ln for ln in txt.splitlines() if needle in ln and not ln.strip().startswith(("#", "//", "*", '"""')) and "test infrastructure" not in ln and "unmodified reference" not in ln and "CPU restatement" not in ln]
                    assert not hits, f"{f} references {needle}: {hits[:2]}"


def test_no_fused_packed_multiply_add_in_the_product():
    """The kernels use Blackwell's packed fp32 pairs (FADD2 / FMUL2).  ptxas 12.9 contracts a packed multiply that feeds a packed
    add into one FFMA2 - rounding once instead of twice - whatever --fmad says, so the sources keep every such sum in scalar adds
    (slv_common.cuh: add2_after_mul).  This pins it: the library's SASS holds packed adds and multiplies and no FFMA2."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", PRODUCT_LIB], capture_output=True, text=True).stdout
    assert sass.count("FADD2") > 100 and sass.count("FMUL2") > 100
    assert sass.count("FFMA2") == 0

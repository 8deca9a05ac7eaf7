"""The reference-side binding, compiled for real: oracle/b200_renderer.hpp is a subclass of the reference's
own renderer_impl (salvia/include/salvia/core/renderer_impl.h:23-112) whose commit_state_and_command() marshals render_state
into the C ABI; oracle/b200_bridge_test.cpp drives one scene (indexed lit mesh + a textured, trilinear-filtered second pass with
a start index, 4x MSAA + resolve, a pipeline-statistics query) through salvia::core::renderer twice - into the reference's
sync_renderer and into b200_renderer bound to a C-ABI library - and compares every buffer and counter.  A third pass draws with
a 16x anisotropic sampler and, where the bound library compiles SASL on the spot (the restatement), with the SASL Sponza pair
through the reference interface's own SASL calls - the binding's compile() (slv_sasl_translate), set_vertex_shader_code /
set_pixel_shader_code, set_vs_variable_value, set_ps_sampler, create_input_layout(descs, n, shader_object) - against the pair's
cpp twins on the reference's sync_renderer; the reference behind the ABI runs the twins on both sides, the CUDA product skips
the pass unless SLV_BRIDGE_PASS3 / SLV_BRIDGE_SASL ask for it (NVRTC at run time).

CPU suite: the library is a CPU checker (the restatement, and the reference behind the ABI).  GPU suite: the CUDA product.  The
binary needs /root/reference to BUILD (oracle/Makefile target `bridge`); it then travels to the GPU box in oracle/_ref/."""
import os
import subprocess

import pytest

from conftest import ORACLE_LIB, PRODUCT_LIB, REF_LIB, ROOT

BRIDGE = os.path.join(ROOT, "oracle", "_ref", "b200_bridge_test")


def run(lib, *size):
    if not os.path.exists(BRIDGE):
        pytest.skip("oracle/_ref/b200_bridge_test not built (needs /root/reference)")
    out = subprocess.run([BRIDGE, lib, *map(str, size)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    lines = out.stdout.strip().splitlines()
    assert lines[-1] == "MATCH", out.stdout
    return lines


# multiples of 4: the reference writes past the end of targets that are not
@pytest.mark.parametrize("size", [(320, 240, 4), (200, 120, 1), (260, 132, 4)])
def test_bridge_into_the_restatement(built, size):
    lines = run(ORACLE_LIB, *size)
    assert lines[0] == "backend oracle" and lines[1].startswith("pass 3: SASL pair")


def test_bridge_into_the_reference_behind_the_abi(built):
    if not os.path.exists(REF_LIB):
        pytest.skip("reference library not built here")
    assert run(REF_LIB, 320, 240, 4)[0] == "backend reference"


@pytest.mark.gpu
@pytest.mark.parametrize("size", [(320, 240, 4), (1280, 720, 4), (200, 120, 1)])
def test_bridge_into_the_cuda_product(built, size):
    lines = run(PRODUCT_LIB, *size)
    assert lines[0] == "backend cuda-sm100a"

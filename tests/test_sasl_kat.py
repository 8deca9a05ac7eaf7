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

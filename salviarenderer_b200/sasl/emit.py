"""SASL front end as a command: what the C++ host surface's compile() runs (salviarenderer_b200/host/salvia_b200_renderer.hpp).

    python -m salviarenderer_b200.sasl.emit vs|ps [--entry NAME] < shader.sasl > unit.txt

Prints the shader's reflection and the generated device code in a line-oriented format (the generated code goes to
slv_shader_compile, which compiles it in process with NVRTC):

    SLVSASL 1
    stage vs|ps
    n_vs_output_attrs N
    uniform_bytes N
    uses_derivatives 0|1
    uniform NAME TYPE OFFSET SIZE          (one per global)
    sampler SLOT NAME
    input SEMANTIC INDEX REGISTER          (VS: input register; PS: attribute)
    output SEMANTIC INDEX ATTRIBUTE
    code NBYTES
    <NBYTES of generated code>

A compile error prints `error` + the message and exits with status 2."""
from __future__ import annotations

import argparse
import re
import sys

from . import frontend


def split_semantic(sem: str):
    m = re.match(r"^(.*?)(\d*)$", sem)
    return (m.group(1).upper(), int(m.group(2) or 0))


def render(unit: frontend.ShaderUnit) -> str:
    r = unit.reflection
    out = ["SLVSASL 1", f"stage {unit.stage}", f"n_vs_output_attrs {r.n_vs_output_attrs}", f"uniform_bytes {r.uniform_bytes}",
           f"uses_derivatives {int(r.uses_derivatives)}"]
    for name, ty, off, size in r.uniforms:
        out.append(f"uniform {name} {ty.replace(' ', '')} {off} {size}")
    for slot, name in enumerate(r.samplers):
        out.append(f"sampler {slot} {name}")
    for k, item in enumerate(r.inputs):
        sem, idx = split_semantic(str(item[0])) if not isinstance(item[1], int) else (str(item[0]).upper(), int(item[1]))
        out.append(f"input {sem} {idx} {k}")
    for k, item in enumerate(r.outputs):
        sem, idx = split_semantic(str(item[0])) if not isinstance(item[1], int) else (str(item[0]).upper(), int(item[1]))
        out.append(f"output {sem} {idx} {k}")
    code = unit.code.encode()
    out.append(f"code {len(code)}")
    return "\n".join(out) + "\n" + unit.code


def main(argv=None) -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("stage", choices=["vs", "ps"])
    ap.add_argument("--entry", default=None)
    a = ap.parse_args(argv)
    try:
        unit = frontend.compile_shader(sys.stdin.read(), a.stage, a.entry)
    except frontend.CompileError as e:
        sys.stdout.write("error\n" + str(e) + "\n")
        return 2
    sys.stdout.write(render(unit))
    return 0


if __name__ == "__main__":
    sys.exit(main())

"""Seeded random SASL sources for differential tests of the two front ends (tests/test_sasl_frontend_cpp.py).  The generator is
typed, but a share of its productions (`sloppy`) ignores the type it was asked for: a good part of the programs is ill-typed, so
the rejection paths are compared as well."""
import numpy as np

TYPES = ["float", "float2", "float3", "float4", "int", "uint", "bool", "int2", "uint3", "bool4", "float3x3", "float4x4", "float2x3"]
FLOATS = {1: "float", 2: "float2", 3: "float3", 4: "float4"}
F_UNARY = ["sqrt", "exp", "exp2", "log", "log2", "log10", "sin", "cos", "tan", "asin", "acos", "atan", "sinh", "cosh", "tanh", "floor", "ceil", "trunc", "round",
           "abs", "rsqrt", "frac", "saturate", "sign", "radians", "degrees", "rcp", "normalize"]
F_BINARY = ["min", "max", "pow", "fmod", "step", "atan2", "ldexp", "reflect"]
F_TERNARY = ["clamp", "lerp", "smoothstep", "mad"]


class Gen:
    """Typed generator: expr(t) builds an expression of type t (mostly - `sloppy` of the productions ignore the type)."""

    def __init__(self, seed, sloppy=0.02):
        self.r = np.random.default_rng(seed)
        self.sloppy = sloppy
        self.vars = [("p", "float4"), ("q", "float2"), ("gain", "float"), ("M", "float4x4"), ("N", "float3x3"), ("count", "int"), ("flags", "uint"), ("on", "bool")]
        self.n = 0
        self.div = 0  # > 0 inside if / loop / switch bodies: derivative intrinsics are rejected there

    def pick(self, xs):
        if self.div and not self.chance(0.03):  # mostly keep ddx / ddy / implicit-derivative fetches out of divergent code
            calm = [x for x in xs if not (isinstance(x, str) and x.startswith(("ddx(", "ddy(", "tex2D(", "tex2Dproj(", "tex2Dbias(")))]
            xs = calm or xs
        return xs[int(self.r.integers(0, len(xs)))]

    def chance(self, p):
        return self.r.random() < p

    def fvec(self, n, depth):
        return self.expr(FLOATS[n], depth)

    def expr(self, t, depth=0):
        if self.chance(self.sloppy):
            t = self.pick(TYPES)
        deep = depth > 3
        if t in FLOATS.values():
            n = int(t[-1]) if t[-1].isdigit() else 1
            r = self.r.random()
            if deep or r < 0.25:
                cands = [v for v, ty in self.vars if ty in FLOATS.values() and (int(ty[-1]) if ty[-1].isdigit() else 1) >= n]
                if cands and not self.chance(0.2):
                    v = self.pick(cands)
                    ty = dict(self.vars)[v]
                    w = int(ty[-1]) if ty[-1].isdigit() else 1
                    if w == n and self.chance(0.5):
                        return v
                    return v + "." + "".join(self.pick("xyzw"[:w]) for _ in range(n))
                lit = self.pick(["0.5", "1.0f", "2.5f", ".25", "3.", "1e-3", "2", "7", "2.0h"])
                return lit if n == 1 else f"{t}({', '.join(self.pick(['0.5f', '1', '2.0f', '-1.5f']) for _ in range(n))})"
            if r < 0.45:
                return f"({self.expr(t, depth + 1)} {self.pick(['+', '-', '*', '/', '%'])} {self.expr(self.pick([t, 'float']), depth + 1)})"
            if r < 0.5:
                return f"(-{self.expr(t, depth + 1)})"
            if r < 0.56:
                return f"({self.expr('bool', depth + 1)} ? {self.expr(t, depth + 1)} : {self.expr(t, depth + 1)})"
            if r < 0.66:
                return f"{self.pick(F_UNARY)}({self.expr(t, depth + 1)})"
            if r < 0.73:
                return f"{self.pick(F_BINARY)}({self.expr(t, depth + 1)}, {self.expr(self.pick([t, 'float']), depth + 1)})"
            if r < 0.78:
                return f"{self.pick(F_TERNARY)}({self.expr(t, depth + 1)}, {self.expr(t, depth + 1)}, {self.expr(self.pick([t, 'float']), depth + 1)})"
            if r < 0.84 and n > 1:  # constructor from parts
                parts, left = [], n
                while left:
                    k = int(self.r.integers(1, left + 1))
                    parts.append(self.fvec(k, depth + 1))
                    left -= k
                return f"{t}({', '.join(parts)})"
            if r < 0.87:
                return f"({t}){self.expr(self.pick(['int', 'uint', 'float4', 'bool']), depth + 1)}" if n == 1 else f"({t}){self.expr('float4', depth + 1)}"
            if n == 1:
                return self.pick([f"dot({self.fvec(3, depth + 1)}, {self.fvec(3, depth + 1)})", f"length({self.fvec(self.pick([2, 3, 4]), depth + 1)})",
                                  f"distance({self.fvec(2, depth + 1)}, {self.fvec(2, depth + 1)})", f"helper({self.fvec(1, depth + 1)}, {self.fvec(3, depth + 1)})",
                                  f"M._m{self.pick('0123')}{self.pick('0123')}", f"N[{self.pick('012')}].{self.pick('xyz')}", f"p[{self.expr('int', depth + 1)}]",
                                  f"asfloat({self.expr(self.pick(['int', 'uint']), depth + 1)})", f"ddx({self.fvec(1, depth + 1)})", f"ddy({self.fvec(1, depth + 1)})"])
            if n == 3:
                return self.pick([f"cross({self.fvec(3, depth + 1)}, {self.fvec(3, depth + 1)})", f"mul({self.fvec(3, depth + 1)}, N)", f"mul(N, {self.fvec(3, depth + 1)})",
                                  f"N[{self.expr('int', depth + 1)}]", f"refract({self.fvec(3, depth + 1)}, {self.fvec(3, depth + 1)}, {self.fvec(1, depth + 1)})",
                                  f"faceforward({self.fvec(3, depth + 1)}, {self.fvec(3, depth + 1)}, {self.fvec(3, depth + 1)})", f"transpose(N)[{self.pick('012')}]"])
            if n == 4:
                return self.pick([f"mul({self.fvec(4, depth + 1)}, M)", f"mul(M, {self.fvec(4, depth + 1)})", f"dst({self.fvec(4, depth + 1)}, {self.fvec(4, depth + 1)})",
                                  f"lit({self.fvec(1, depth + 1)}, {self.fvec(1, depth + 1)}, {self.fvec(1, depth + 1)})", f"tex2D(samp, {self.fvec(2, depth + 1)})",
                                  f"tex2Dlod(samp, {self.fvec(4, depth + 1)})", f"tex2Dproj(samp, {self.fvec(4, depth + 1)})", f"tex2Dbias(samp, {self.fvec(4, depth + 1)})",
                                  f"tex2Dgrad(samp, {self.fvec(2, depth + 1)}, {self.fvec(2, depth + 1)}, {self.fvec(2, depth + 1)})", f"M[{self.pick('0123')}]"])
            return self.pick([f"ddx({self.fvec(2, depth + 1)})", f"({self.fvec(2, depth + 1)} * 0.5f)"])
        if t in ("int", "uint"):
            r = self.r.random()
            if deep or r < 0.35:
                cands = [v for v, ty in self.vars if ty == t]
                return self.pick(cands) if cands and self.chance(0.6) else self.pick(["0", "1", "2", "7", "10L", "0x1F"] if t == "int" else ["3u", "0xFFu", "1u"])
            if r < 0.65:
                return f"({self.expr(t, depth + 1)} {self.pick(['+', '-', '*', '/', '%', '&', '|', '^', '<<', '>>'])} {self.expr(t, depth + 1)})"
            if r < 0.72:
                return f"({self.pick(['-', '~'])}{self.expr(t, depth + 1)})"
            if r < 0.82:
                return f"({t}){self.expr(self.pick(['float', 'int', 'uint', 'bool']), depth + 1)}"
            if r < 0.9:
                return f"{self.pick(['countbits', 'firstbithigh', 'firstbitlow', 'reversebits', 'abs'])}({self.expr(t, depth + 1)})"
            return f"{'asint' if t == 'int' else 'asuint'}({self.fvec(1, depth + 1)})"
        if t == "bool":
            r = self.r.random()
            if deep or r < 0.2:
                return self.pick(["on", "true", "false"])
            if r < 0.6:
                k = self.pick(["float", "int", "uint"])
                return f"({self.expr(k, depth + 1)} {self.pick(['<', '>', '<=', '>=', '==', '!='])} {self.expr(k, depth + 1)})"
            if r < 0.8:
                return f"({self.expr('bool', depth + 1)} {self.pick(['&&', '||'])} {self.expr('bool', depth + 1)})"
            if r < 0.88:
                return f"(!{self.expr('bool', depth + 1)})"
            n = self.pick([2, 3, 4])
            return f"{self.pick(['any', 'all'])}({self.fvec(n, depth + 1)} {self.pick(['<', '>=', '!='])} {self.fvec(n, depth + 1)})" if self.chance(0.6) else \
                f"{self.pick(['isnan', 'isinf', 'isfinite'])}({self.fvec(1, depth + 1)})"
        if t == "float3x3":
            return self.pick(["N", "transpose(N)", "mul(N, N)", "(N * 2.0f)", "(float3x3)M", f"float3x3({self.fvec(3, depth + 1)}, {self.fvec(3, depth + 1)}, {self.fvec(3, depth + 1)})"])
        if t == "float4x4":
            return self.pick(["M", "transpose(M)", "mul(M, M)", "(M + M)"])
        if t in ("int2", "uint3", "bool4"):
            base, n = t[:-1], int(t[-1])
            return f"{t}({', '.join(self.expr(base, depth + 1) for _ in range(n))})"
        return f"{t}({self.fvec(2, depth + 1)}, {self.fvec(4, depth + 1)})"  # float2x3

    def stmt(self, depth=0, in_loop=False):
        r = self.r.random()
        if depth > 2 or r < 0.35:
            t = self.pick(TYPES[:7] + ["float3", "float4", "float"]) if not self.chance(0.1) else self.pick(TYPES)
            self.n += 1
            name = f"v{self.n}"
            s = f"{t} {name} = {self.expr(t)};"
            self.vars.append((name, t))
            return s
        if r < 0.6:
            v, t = self.pick([x for x in self.vars if x[0] not in ("M", "N", "on")])
            if t in FLOATS.values() and t != "float" and self.chance(0.4):
                w = int(t[-1])
                k = int(self.r.integers(1, w + 1))
                lanes = list(self.r.permutation(w)[:k])
                return f"{v}.{''.join('xyzw'[i] for i in lanes)} {self.pick(['=', '+=', '*='])} {self.fvec(k, 1)};"
            ops = ["=", "+=", "-=", "*=", "/="] + (["%=", "&=", "|=", "<<=", "^=", ">>="] if t in ("int", "uint") else [])
            return f"{v} {self.pick(ops)} {self.expr(t if t in TYPES else 'float', 1)};"
        if r < 0.66:
            v = self.pick([x[0] for x in self.vars if x[1] in ("int", "uint", "float")])
            return self.pick([f"{v}++;", f"{v}--;", f"++{v};", f"--{v};"])
        n_vars = len(self.vars)
        if r < 0.78:
            s = f"if ({self.expr('bool')}) {{ {self.block(depth + 1, in_loop)} }}"
            if self.chance(0.5):
                s += f" else {{ {self.block(depth + 1, in_loop)} }}"
        elif r < 0.88:
            self.n += 1
            i = f"i{self.n}"
            self.vars.append((i, "int"))
            s = f"for (int {i} = 0; {i} < {int(self.r.integers(1, 5))}; ++{i}) {{ {self.block(depth + 1, True)} }}"
        elif r < 0.93:
            s = f"switch ({self.expr('int')}) {{ case 0: {self.block(depth + 1, in_loop)} case 1: case -2: {self.block(depth + 1, in_loop)} break; " \
                f"default: {self.block(depth + 1, in_loop)} }}"
        elif r < 0.96 and in_loop:
            s = self.pick(["break;", "continue;"])
        elif r < 0.98:
            s = f"do {{ {self.block(depth + 1, True)} }} while ({self.expr('bool')});"
        else:
            s = f"return {self.expr('float4')};"
        del self.vars[n_vars:]
        return s

    def block(self, depth, in_loop):
        n_vars = len(self.vars)
        self.div += 1
        s = " ".join(self.stmt(depth, in_loop) for _ in range(int(self.r.integers(1, 4))))
        self.div -= 1
        del self.vars[n_vars:]
        return s

    def program(self):
        body = "\n    ".join(self.stmt() for _ in range(int(self.r.integers(2, 8))))
        ret = self.expr("float4")
        return f"""sampler samp;
float gain; float4x4 M; float3x3 N; int count; uint flags; bool on;
float helper(float a, float3 b) {{ return a * dot(b, b); }}
struct PSIn {{ float4 p: TEXCOORD0; float2 q: TEXCOORD1; }};
float4 ps_main(PSIn i): COLOR {{
    float4 p = i.p; float2 q = i.q;
    {body}
    return {ret};
}}
"""


def program_from_seed(seed: int, sloppy: float = 0.02) -> str:
    return Gen(seed, sloppy).program()

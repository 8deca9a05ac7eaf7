"""SASL front end: source text -> scalarised C++ device code + reflection.

Scope (what the samples' shaders and sasl/test/repo/*.svs|*.sps use): global uniforms (scalars, vectors, matrices) and
samplers; structs with semantics; functions; local declarations; assignment (plain and compound) to variables, members,
swizzles and elements; if / else, for, while, do-while, break, continue, return; arithmetic, comparison, logical and
ternary operators; constructors, casts, swizzles, indexing; the intrinsics listed in INTRINSICS.

Code generation is one pass over the AST into three-address form over SCALARS: every vector / matrix / struct value is
a list of C scalar expressions, every non-trivial operation lands in a fresh `const` temporary, so the order of the
floating-point operations is exactly the order written here (the product compiles with -fmad=false: no contraction).
Numerics follow the cpp-shader twins the samples ship (eflib): `mul(v, M)` accumulates left to right over the rows of M
(eflib/src/math.cpp:142-154), `dot` left to right, `normalize` = v * (1 / length) with eflib's zero-length guard.
The reference's own SASL numerics come from its LLVM code generator, which cannot be built here: parity of
this path is pinned against the cpp twins, not against the reference JIT.
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field

__all__ = ["CompileError", "Reflection", "ShaderUnit", "compile_shader"]


class CompileError(Exception):
    pass


# ------------------------------------------------------------------------------------------------- types
@dataclass(frozen=True)
class Type:
    kind: str           # 'void' | 'scalar' | 'vector' | 'matrix' | 'struct' | 'sampler'
    base: str = "float"  # 'float' | 'int' | 'uint' | 'bool'
    rows: int = 1
    cols: int = 1
    name: str = ""      # struct name

    @property
    def n(self) -> int:
        return self.rows * self.cols

    def __str__(self):
        if self.kind == "struct":
            return self.name
        if self.kind in ("void", "sampler"):
            return self.kind
        if self.kind == "scalar":
            return self.base
        if self.kind == "vector":
            return f"{self.base}{self.cols}"
        return f"{self.base}{self.rows}x{self.cols}"


VOID = Type("void")
SAMPLER = Type("sampler")
FLOAT = Type("scalar", "float")
INT = Type("scalar", "int")
UINT = Type("scalar", "uint")
BOOL = Type("scalar", "bool")


def vec(base, n):
    return Type("scalar", base) if n == 1 else Type("vector", base, 1, n)


def mat(base, r, c):
    return Type("matrix", base, r, c)


C_BASE = {"float": "float", "int": "int", "uint": "unsigned", "bool": "bool"}
RANK = {"bool": 0, "int": 1, "uint": 2, "float": 3}


def parse_type_name(s: str):
    m = re.fullmatch(r"(float|int|uint|bool|half|double|int8_t|int16_t|int32_t|int64_t|uint8_t|uint16_t|uint32_t|uint64_t)(?:([1-4])(?:x([1-4]))?)?", s)
    if not m:
        return None
    base = m.group(1)
    base = {"half": "float", "double": "float"}.get(base, "uint" if base.startswith("uint") else ("int" if base.startswith("int") else base))
    if m.group(3):
        return mat(base, int(m.group(2)), int(m.group(3)))
    if m.group(2):
        return vec(base, int(m.group(2)))
    return Type("scalar", base)


# ------------------------------------------------------------------------------------------------- lexer
TOKEN_RE = re.compile(r"""
    (?P<ws>\s+|//[^\n]*|/\*.*?\*/)
  | (?P<num>0[xX][0-9a-fA-F]+[uU]?|(?:\d+\.\d*|\.\d+|\d+)(?:[eE][+-]?\d+)?[fFhHuUlL]?)
  | (?P<id>[A-Za-z_]\w*)
  | (?P<op>\+\+|--|<<=|>>=|<<|>>|<=|>=|==|!=|&&|\|\||\+=|-=|\*=|/=|%=|&=|\|=|\^=|[-+*/%<>=!&|^~?:;,.(){}\[\]])
""", re.S | re.X)


@dataclass
class Tok:
    kind: str
    text: str
    line: int


def lex(src: str):
    out, pos, line = [], 0, 1
    while pos < len(src):
        m = TOKEN_RE.match(src, pos)
        if not m:
            raise CompileError(f"line {line}: unexpected character {src[pos]!r}")
        pos = m.end()
        kind = m.lastgroup
        text = m.group()
        if kind != "ws":
            out.append(Tok(kind, text, line))
        line += text.count("\n")
    out.append(Tok("eof", "", line))
    return out


# ------------------------------------------------------------------------------------------------- AST
@dataclass
class Node:
    op: str
    args: tuple = ()
    line: int = 0


@dataclass
class VarDecl:
    type: Type
    name: str
    semantic: str | None
    init: Node | None
    line: int
    array: int = 0              # > 0: literal element count; -1: sized by another global (array_len)
    array_len: str | None = None


@dataclass
class Func:
    name: str
    ret: Type
    ret_semantic: str | None
    params: list
    body: Node
    line: int


def norm_semantic(s: str | None):
    """'TEXCOORD(1)' / 'texcoord1' / 'Texcoord' -> ('TEXCOORD', 1)."""
    if s is None:
        return None
    m = re.fullmatch(r"([A-Za-z_]+?)(?:\((\d+)\)|(\d+))?", s.strip())
    if not m:
        raise CompileError(f"bad semantic {s!r}")
    return (m.group(1).upper(), int(m.group(2) or m.group(3) or 0))


# salvia/include/salvia/shader/constants.h:22-37 (enum system_values) and :54-79 (the names that map to them)
_SYSTEM_VALUE = {"POSITION": 1, "SV_POSITION": 1, "TEXCOORD": 2, "NORMAL": 3, "BLEND_INDICES": 4, "BLEND_WEIGHTS": 5, "PSIZE": 6,
                 "COLOR": 7, "SV_TARGET": 7, "DEPTH": 8, "SV_DEPTH": 8}


def _semantic_key(sem):
    sv = _SYSTEM_VALUE.get(sem[0])
    return (sv, "", sem[1]) if sv else (9, sem[0].lower(), sem[1])  # sv_customized carries its lower-case name


def reference_semantic_order(sems) -> list:
    """Indices of `sems` ((NAME, index) pairs, in declaration order) in the order of the reference's semantic array.
    The reference keeps a shader's input / output semantics in an array it inserts into at std::lower_bound
    (reflection_impl.cpp:70-118: add_input_semantic / add_output_semantic), and every consumer walks that array: vertex-shader
    outputs -> attribute registers (sasl/src/shims/interp_shim.cpp:57-76), pixel-shader inputs <- attributes 0, 1, 2 ...
    (pixel_shader_unit::update, shader_unit.cpp:106-140).  semantic_value::operator< is `sv < r.sv || name < r.name || index <
    r.index` (constants.h:94-96) - not a strict weak order (NORMAL0 < TEXCOORD1 and TEXCOORD1 < NORMAL0), so the result depends on
    the insertion sequence; this replays the binary search of libstdc++'s lower_bound with that predicate.  A semantic equal to
    the element found is rejected (the reference's add_*_semantic returns false: "ABI analysis error")."""
    keys = [_semantic_key(s) for s in sems]

    def less(a, b):
        return a[0] < b[0] or a[1] < b[1] or a[2] < b[2]

    order = []
    for i, k in enumerate(keys):
        first, n = 0, len(order)
        while n > 0:
            half = n >> 1
            if less(keys[order[first + half]], k):
                first += half + 1
                n -= half + 1
            else:
                n = half
        if first < len(order) and keys[order[first]] == k:
            raise CompileError(f"semantic {sems[i][0]}{sems[i][1]} is bound twice")
        order.insert(first, i)
    return order


class Parser:
    ASSIGN_OPS = {"=", "+=", "-=", "*=", "/=", "%=", "&=", "|=", "^=", "<<=", ">>="}
    BIN_PREC = [("||",), ("&&",), ("|",), ("^",), ("&",), ("==", "!="), ("<", ">", "<=", ">="), ("<<", ">>"), ("+", "-"), ("*", "/", "%")]

    def __init__(self, src):
        self.toks = lex(src)
        self.i = 0
        self.structs: dict[str, list[VarDecl]] = {}

    @property
    def t(self):
        return self.toks[self.i]

    def err(self, msg):
        raise CompileError(f"line {self.t.line}: {msg} (at {self.t.text!r})")

    def accept(self, text):
        if self.t.text == text and self.t.kind in ("op", "id"):
            self.i += 1
            return True
        return False

    def expect(self, text):
        if not self.accept(text):
            self.err(f"expected {text!r}")

    def ident(self):
        if self.t.kind != "id":
            self.err("expected an identifier")
        self.i += 1
        return self.toks[self.i - 1].text

    def is_type(self, tok=None):
        tok = tok or self.t
        return tok.kind == "id" and (tok.text in ("void", "sampler") or tok.text in self.structs or parse_type_name(tok.text) is not None)

    def type(self):
        while self.t.text in ("const", "uniform", "static", "in", "out", "inout") and self.toks[self.i + 1].kind == "id" and self.is_type(self.toks[self.i + 1]):
            self.i += 1
        name = self.ident()
        if name == "void":
            return VOID
        if name == "sampler":
            return SAMPLER
        if name in self.structs:
            return Type("struct", name=name)
        ty = parse_type_name(name)
        if ty is None:
            self.err(f"unknown type {name!r}")
        return ty

    def semantic(self):
        if self.accept(":"):
            s = self.ident()
            if self.accept("("):
                if self.t.kind != "num":
                    self.err("expected a semantic index")
                s += f"({self.t.text})"
                self.i += 1
                self.expect(")")
            return s
        return None

    # ---- top level
    def program(self):
        globals_, funcs = [], []
        while self.t.kind != "eof":
            if self.accept(";"):
                continue
            if self.accept("struct"):
                name = self.ident()
                self.expect("{")
                members = []
                while not self.accept("}"):
                    ty = self.type()
                    while True:
                        n = self.ident()
                        members.append(VarDecl(ty, n, self.semantic(), None, self.t.line))
                        if not self.accept(","):
                            break
                    self.expect(";")
                self.accept(";")
                self.structs[name] = members
                continue
            line = self.t.line
            ty = self.type()
            name = self.ident()
            if self.accept("("):
                params = []
                if not self.accept(")"):
                    while True:
                        pty = self.type()
                        pname = self.ident()
                        params.append(VarDecl(pty, pname, self.semantic(), None, self.t.line))
                        if not self.accept(","):
                            break
                    self.expect(")")
                rsem = self.semantic()
                body = self.block()
                funcs.append(Func(name, ty, rsem, params, body, line))
            else:
                while True:
                    arr, arr_len = 0, None
                    if self.accept("["):
                        if self.t.kind == "num" and re.fullmatch(r"\d+", self.t.text):
                            arr = int(self.t.text)
                        elif self.t.kind == "id":  # `float4x4 bones[boneCount]`: sized at run time by another global
                            arr, arr_len = -1, self.t.text
                        else:
                            raise CompileError(f"line {self.t.line}: the size of an array must be an integer literal or the name of a global")
                        self.i += 1
                        self.expect("]")
                    sem = self.semantic()
                    init = self.assign_expr() if self.accept("=") else None
                    globals_.append(VarDecl(ty, name, sem, init, line, arr, arr_len))
                    if not self.accept(","):
                        break
                    name = self.ident()
                self.expect(";")
        return globals_, funcs

    # ---- statements
    def block(self):
        line = self.t.line
        self.expect("{")
        stmts = []
        while not self.accept("}"):
            stmts.append(self.statement())
        return Node("block", tuple(stmts), line)

    def statement(self):
        line = self.t.line
        if self.t.text == "{":
            return self.block()
        if self.accept(";"):
            return Node("block", (), line)
        if self.accept("if"):
            self.expect("(")
            c = self.expr()
            self.expect(")")
            a = self.statement()
            b = self.statement() if self.accept("else") else None
            return Node("if", (c, a, b), line)
        if self.accept("for"):
            self.expect("(")
            init = None if self.t.text == ";" else self.simple_statement()
            self.expect(";")
            cond = None if self.t.text == ";" else self.expr()
            self.expect(";")
            step = None if self.t.text == ")" else self.expr()
            self.expect(")")
            return Node("for", (init, cond, step, self.statement()), line)
        if self.accept("while"):
            self.expect("(")
            c = self.expr()
            self.expect(")")
            return Node("for", (None, c, None, self.statement()), line)
        if self.accept("do"):
            body = self.statement()
            self.expect("while")
            self.expect("(")
            c = self.expr()
            self.expect(")")
            self.expect(";")
            return Node("dowhile", (body, c), line)
        if self.accept("switch"):
            self.expect("(")
            sel = self.expr()
            self.expect(")")
            self.expect("{")
            groups = []  # [(labels, statements)]: labels = case expressions, None for `default`
            while not self.accept("}"):
                labels = []
                while self.t.text in ("case", "default"):
                    if self.accept("default"):
                        labels.append(None)
                    else:
                        self.expect("case")
                        labels.append(self.expr())
                    self.expect(":")
                if not labels:
                    raise CompileError(f"line {self.t.line}: statement before the first case label of a switch")
                stmts = []
                while self.t.text not in ("case", "default", "}"):
                    stmts.append(self.statement())
                groups.append((tuple(labels), tuple(stmts)))
            return Node("switch", (sel, tuple(groups)), line)
        if self.accept("break"):
            self.expect(";")
            return Node("break", (), line)
        if self.accept("continue"):
            self.expect(";")
            return Node("continue", (), line)
        if self.accept("return"):
            e = None if self.t.text == ";" else self.expr()
            self.expect(";")
            return Node("return", (e,), line)
        s = self.simple_statement()
        self.expect(";")
        return s

    def simple_statement(self):
        line = self.t.line
        if self.is_type() and self.toks[self.i + 1].kind == "id":
            ty = self.type()
            decls = []
            while True:
                name = self.ident()
                init = self.assign_expr() if self.accept("=") else None
                decls.append(VarDecl(ty, name, None, init, line))
                if not self.accept(","):
                    break
            return Node("decl", tuple(decls), line)
        return Node("expr", (self.expr(),), line)

    # ---- expressions
    def expr(self):
        e = self.assign_expr()
        while self.accept(","):
            e = Node("comma", (e, self.assign_expr()), e.line)
        return e

    def assign_expr(self):
        lhs = self.ternary()
        if self.t.kind == "op" and self.t.text in self.ASSIGN_OPS:
            op = self.t.text
            self.i += 1
            rhs = self.assign_expr()
            return Node("assign", (op, lhs, rhs), lhs.line)
        return lhs

    def ternary(self):
        c = self.binary(0)
        if self.accept("?"):
            a = self.assign_expr()
            self.expect(":")
            b = self.assign_expr()
            return Node("select", (c, a, b), c.line)
        return c

    def binary(self, level):
        if level == len(self.BIN_PREC):
            return self.unary()
        lhs = self.binary(level + 1)
        while self.t.kind == "op" and self.t.text in self.BIN_PREC[level]:
            op = self.t.text
            self.i += 1
            rhs = self.binary(level + 1)
            lhs = Node("bin", (op, lhs, rhs), lhs.line)
        return lhs

    def unary(self):
        line = self.t.line
        if self.t.kind == "op" and self.t.text in ("-", "+", "!", "~"):
            op = self.t.text
            self.i += 1
            return Node("un", (op, self.unary()), line)
        if self.t.kind == "op" and self.t.text in ("++", "--"):
            op = self.t.text
            self.i += 1
            target = self.unary()
            return Node("assign", ("+=" if op == "++" else "-=", target, Node("num", ("1",), line)), line)
        # C-style cast: '(' type ')' unary
        if self.t.text == "(" and self.is_type(self.toks[self.i + 1]) and self.toks[self.i + 2].text == ")":
            self.i += 1
            ty = self.type()
            self.expect(")")
            return Node("call", (str(ty), (self.unary(),)), line)
        return self.postfix()

    def postfix(self):
        e = self.primary()
        while True:
            line = self.t.line
            if self.accept("."):
                e = Node("member", (e, self.ident()), line)
            elif self.accept("["):
                idx = self.expr()
                self.expect("]")
                e = Node("index", (e, idx), line)
            elif self.t.kind == "op" and self.t.text in ("++", "--"):
                op = self.t.text
                self.i += 1
                e = Node("postinc", (e, "+" if op == "++" else "-"), line)
            else:
                return e

    def primary(self):
        line = self.t.line
        if self.t.kind == "num":
            self.i += 1
            return Node("num", (self.toks[self.i - 1].text,), line)
        if self.accept("("):
            e = self.expr()
            self.expect(")")
            return e
        if self.t.kind == "id":
            if self.t.text in ("true", "false"):
                self.i += 1
                return Node("bool", (self.toks[self.i - 1].text,), line)
            name = self.ident()
            if self.accept("("):
                args = []
                if not self.accept(")"):
                    while True:
                        args.append(self.assign_expr())
                        if not self.accept(","):
                            break
                    self.expect(")")
                return Node("call", (name, tuple(args)), line)
            return Node("var", (name,), line)
        self.err("expected an expression")


# ------------------------------------------------------------------------------------------------- code generation
@dataclass
class Value:
    type: Type
    comps: list          # C scalar expressions (struct: flattened member after member)
    lvalue: bool = False  # comps are assignable C lvalues


@dataclass
class Reflection:
    stage: str
    entry: str
    uniforms: list = field(default_factory=list)    # (name, type string, byte offset, byte size)
    uniform_bytes: int = 0
    samplers: list = field(default_factory=list)    # names, in slot order
    arrays: dict = field(default_factory=dict)      # array uniform -> (element type, element bytes, length: int or the global that holds it)
    inputs: list = field(default_factory=list)      # VS: (semantic, index, type) -> input register k; PS: -> attribute k
    outputs: list = field(default_factory=list)     # VS: non-position outputs -> attribute k; PS: colour targets
    n_vs_output_attrs: int = 0
    uses_derivatives: bool = False
    writes_depth: bool = False

    def uniform(self, name):
        for u in self.uniforms:
            if u[0] == name:
                return u
        raise KeyError(name)


@dataclass
class ShaderUnit:
    stage: str            # 'vs' | 'ps'
    code: str             # C++ definitions (device functions + the entry wrapper), needs sasl_rt.h first
    reflection: Reflection
    source: str

    def pack_uniforms(self, values: dict) -> bytes:
        """name -> number / sequence of numbers (matrices row-major), laid out as the generated code reads them.
        Unknown names raise (the reference's set_*_variable fails the same way)."""
        import struct
        buf = bytearray(self.reflection.uniform_bytes)
        for name, v in values.items():
            _, ty, off, size = self.reflection.uniform(name)
            if ty.endswith("[]"):  # an array uniform: the address of its buffer (device pointer / host pointer)
                struct.pack_into("<Q", buf, off, int(v))
                continue
            flat = list(v) if hasattr(v, "__iter__") else [v]
            flat = [x for row in flat for x in (row if hasattr(row, "__iter__") else [row])]
            if len(flat) * 4 != size:
                raise ValueError(f"uniform {name}: expected {size // 4} scalars, got {len(flat)}")
            fmt = "i" if ty.startswith("int") or ty.startswith("bool") else ("I" if ty.startswith("uint") else "f")
            struct.pack_into(f"<{len(flat)}{fmt}", buf, off, *[(int(x) if fmt != "f" else float(x)) for x in flat])
        return bytes(buf)


SWZ = {"x": 0, "y": 1, "z": 2, "w": 3, "r": 0, "g": 1, "b": 2, "a": 3}
# Math intrinsics go through sasl_rt.h's sasl_m_* wrappers, which mirror the host functions the reference links its JIT-ed code
# to (sasl/src/drivers/compiler_impl.cpp:340-404): exp / log10 / sin ... pow / fmod are the C library's float functions (on the
# GPU: the double-precision device function rounded ONCE to float = the correctly rounded result, which is what glibc's float
# functions return but for rare last-bit cases; CUDA's own float versions are 1-2 ulp off); exp2 is ldexpf(1, (int)x); log and
# log2 are eflib's fast_log / fast_log2 polynomial; floor / ceil / round / trunc are eflib's fast_floor / fast_ceil / fast_round
# / trunc (magic-number rounding); ldexp(x, e) is ldexpf(x, (int)e).  Pure float arithmetic, bit-identical on host and device.
UNARY_MATH = {"sqrt": "sqrtf", "exp": "sasl_m_exp", "exp2": "sasl_m_exp2", "log": "sasl_m_log", "log2": "sasl_m_log2",
              "log10": "sasl_m_log10", "sin": "sasl_m_sin", "cos": "sasl_m_cos", "tan": "sasl_m_tan", "asin": "sasl_m_asin",
              "acos": "sasl_m_acos", "atan": "sasl_m_atan", "sinh": "sasl_m_sinh", "cosh": "sasl_m_cosh", "tanh": "sasl_m_tanh",
              "floor": "sasl_m_floor", "ceil": "sasl_m_ceil", "trunc": "sasl_m_trunc", "round": "sasl_m_round"}
INTRINSICS = sorted(list(UNARY_MATH) + ["abs", "rsqrt", "frac", "saturate", "sign", "radians", "degrees", "min", "max", "pow",
                                         "fmod", "step", "atan2", "ldexp", "clamp", "lerp", "smoothstep", "mad", "dot", "cross", "dst",
                                         "length", "distance", "normalize", "reflect", "mul", "transpose", "any", "all",
                                         "ddx", "ddy", "tex2D", "tex2Dlod", "tex2Dbias", "tex2Dproj", "tex2Dgrad", "asfloat",
                                         "asint", "asuint", "countbits", "count_bits", "firstbithigh", "firstbitlow", "reversebits",
                                         "isinf", "isfinite", "isnan", "rcp", "refract", "lit", "faceforward"])


class Gen:
    def __init__(self, src: str, stage: str, entry: str | None):
        self.src, self.stage = src, stage
        self.p = Parser(src)
        self.globals_, self.funcs = self.p.program()
        self.structs = self.p.structs
        self.lines: list[str] = []
        self.indent = 1
        self.ntemp = 0
        self.scopes: list[dict] = []
        self.fn_table: dict[str, Func] = {}
        self.refl = Reflection(stage, "")
        self.uniform_vars: dict[str, Value] = {}
        self.uniform_arrays: dict[str, Type] = {}   # array uniforms: element type (the block holds the buffer's address)
        self.sampler_slots: dict[str, int] = {}
        self.loop_depth = 0
        self.divergent = 0   # > 0 while emitting code under a data-dependent branch or loop
        self.cur_ret: Value | None = None
        self.entry = None if stage == "lib" else self.pick_entry(entry)
        self.refl.entry = self.entry.name if self.entry else ""

    # ---- helpers
    def err(self, node, msg):
        raise CompileError(f"line {getattr(node, 'line', 0)}: {msg}")

    def emit(self, s):
        self.lines.append("  " * self.indent + s)

    def temp(self, base, expr):
        self.ntemp += 1
        name = f"t{self.ntemp}"
        self.emit(f"const {C_BASE[base]} {name} = {expr};")
        return name

    def flat_types(self, ty: Type):
        """Scalar base types of the flattened components of `ty`."""
        if ty.kind == "struct":
            out = []
            for m in self.structs[ty.name]:
                out += self.flat_types(m.type)
            return out
        if ty.kind == "sampler":
            return ["int"]
        return [ty.base] * ty.n

    def pick_entry(self, entry):
        if not self.funcs:
            raise CompileError("no function in the translation unit")
        if entry:
            for f in self.funcs:
                if f.name == entry:
                    return f
            raise CompileError(f"entry function {entry!r} not found")

        def has_sem(f):
            if f.ret_semantic:
                return True
            for prm in f.params:
                if prm.semantic or (prm.type.kind == "struct" and any(m.semantic for m in self.structs[prm.type.name])):
                    return True
            return f.ret.kind == "struct" and any(m.semantic for m in self.structs[f.ret.name])
        cands = [f for f in self.funcs if has_sem(f)] or [f for f in self.funcs if f.name in ("main", "vs_main", "ps_main", "fn")]
        if not cands:
            raise CompileError("cannot determine the entry function (no parameter or return value carries a semantic)")
        return cands[-1]

    # ---- conversions
    def convert(self, v: Value, to: Type, node=None, explicit=False) -> Value:
        if v.type == to:
            return v
        if to.kind == "struct" or v.type.kind in ("struct", "sampler", "void"):
            self.err(node, f"cannot convert {v.type} to {to}")
        src = v.comps
        if v.type.n == 1 and to.n > 1:
            src = src * to.n
        elif v.type.n > to.n and (explicit or True):  # HLSL truncation
            if v.type.kind == "matrix" or to.kind == "matrix":
                if v.type.kind == "matrix" and to.kind == "matrix" and to.rows <= v.type.rows and to.cols <= v.type.cols:
                    src = [v.comps[r * v.type.cols + c] for r in range(to.rows) for c in range(to.cols)]
                else:
                    self.err(node, f"cannot convert {v.type} to {to}")
            else:
                src = src[:to.n]
        elif v.type.n != to.n:
            self.err(node, f"cannot convert {v.type} to {to}")
        if v.type.base != to.base:
            if to.base == "bool":
                src = [self.temp("bool", f"({c} != 0)") for c in src]
            else:
                src = [self.temp(to.base, f"({C_BASE[to.base]})({c})") for c in src]
        return Value(to, list(src))

    def unify(self, a: Value, b: Value, node, arith=True):
        """Common type for a binary / ternary operation (scalar broadcast, base promotion, vector truncation)."""
        ta, tb = a.type, b.type
        for t in (ta, tb):
            if t.kind not in ("scalar", "vector", "matrix"):
                self.err(node, f"operand of type {t}")
        base = ta.base if RANK[ta.base] >= RANK[tb.base] else tb.base
        if arith and base == "bool":
            base = "int"
        if ta.n == 1:
            shape = tb
        elif tb.n == 1:
            shape = ta
        elif ta.kind == "matrix" or tb.kind == "matrix":
            if (ta.rows, ta.cols) != (tb.rows, tb.cols):
                self.err(node, f"shape mismatch {ta} vs {tb}")
            shape = ta
        else:
            shape = ta if ta.cols <= tb.cols else tb
        to = Type(shape.kind, base, shape.rows, shape.cols)
        return self.convert(a, to, node), self.convert(b, to, node), to

    def materialize(self, v: Value) -> Value:
        """Copies lvalue components into temporaries (so a later store cannot change what was read)."""
        if not v.lvalue:
            return v
        bases = self.flat_types(v.type)
        return Value(v.type, [self.temp(b, c) for b, c in zip(bases, v.comps)])

    # ---- scopes
    def lookup(self, name, node):
        for s in reversed(self.scopes):
            if name in s:
                return s[name]
        if name in self.uniform_vars:
            return self.uniform_vars[name]
        self.err(node, f"undeclared identifier {name!r}")

    def declare(self, ty: Type, name: str, node) -> Value:
        self.ntemp += 1
        bases = self.flat_types(ty)
        names = [f"v{self.ntemp}_{name}_{k}" for k in range(len(bases))]
        for b, n in zip(bases, names):
            self.emit(f"{C_BASE[b]} {n} = 0;")
        v = Value(ty, names, True)
        self.scopes[-1][name] = v
        return v

    def store(self, dst: Value, src: Value, node):
        src = self.materialize(self.convert(src, dst.type, node))
        for d, s in zip(dst.comps, src.comps):
            self.emit(f"{d} = {s};")

    # ---- expressions
    def expr(self, n: Node) -> Value:
        return getattr(self, "e_" + n.op)(n)

    def e_num(self, n):
        t = n.args[0]
        if t.lower().startswith("0x"):
            if int(t.rstrip("uU"), 16) > 0xFFFFFFFF:
                self.err(n, f"integer literal {t} does not fit 32 bits")
            return Value(UINT if t[-1] in "uU" else INT, [str(int(t.rstrip("uU"), 16)) + ("u" if t[-1] in "uU" else "")])
        if re.fullmatch(r"\d+[uU]", t):
            return Value(UINT, [t[:-1] + "u"])
        if re.fullmatch(r"\d+[lL]?", t):
            return Value(INT, [t.rstrip("lL")])
        body = t.rstrip("fFhHlL")
        if "." not in body and "e" not in body.lower():
            body += ".0"
        return Value(FLOAT, [body + "f"])

    def e_bool(self, n):
        return Value(BOOL, [n.args[0]])

    def e_var(self, n):
        return self.lookup(n.args[0], n)

    def e_comma(self, n):
        self.expr(n.args[0])
        return self.expr(n.args[1])

    def e_member(self, n):
        base = self.expr(n.args[0])
        name = n.args[1]
        ty = base.type
        if ty.kind == "struct":
            off = 0
            for m in self.structs[ty.name]:
                k = len(self.flat_types(m.type))
                if m.name == name:
                    return Value(m.type, base.comps[off:off + k], base.lvalue)
                off += k
            self.err(n, f"{ty} has no member {name!r}")
        if ty.kind in ("scalar", "vector") and all(c in SWZ for c in name) and 1 <= len(name) <= 4:
            idx = [SWZ[c] for c in name]
            if max(idx) >= ty.n:
                self.err(n, f"swizzle .{name} out of range for {ty}")
            return Value(vec(ty.base, len(idx)), [base.comps[i] for i in idx], base.lvalue and len(set(idx)) == len(idx))
        m = re.fullmatch(r"_m([0-3])([0-3])|_([1-4])([1-4])", name)
        if ty.kind == "matrix" and m:
            r, c = (int(m.group(1)), int(m.group(2))) if m.group(1) is not None else (int(m.group(3)) - 1, int(m.group(4)) - 1)
            if r < ty.rows and c < ty.cols:
                return Value(Type("scalar", ty.base), [base.comps[r * ty.cols + c]], base.lvalue)
        self.err(n, f"cannot take .{name} of {ty}")

    def e_index(self, n):
        idx = n.args[1]
        # an element of an array uniform: loads through the address the uniform block holds
        if n.args[0].op == "var" and n.args[0].args[0] in self.uniform_arrays and not any(n.args[0].args[0] in sc for sc in self.scopes):
            name = n.args[0].args[0]
            ety = self.uniform_arrays[name]
            iv = self.convert(self.expr(idx), INT, n)
            i = self.temp("int", iv.comps[0])
            return Value(ety, [self.temp(ety.base, f"U.{name}[{i} * {ety.n} + {k}]") for k in range(ety.n)])
        base = self.expr(n.args[0])
        ty = base.type
        if idx.op != "num":
            # a run-time index into a vector / the rows of a matrix: a chain of selects over the components (read-only)
            iv = self.expr(idx)
            if iv.type.kind != "scalar" or iv.type.base not in ("int", "uint"):
                self.err(n, "an index must be an integer scalar")
            if ty.kind not in ("vector", "matrix"):
                self.err(n, f"cannot index {ty}")
            i = self.temp("int", self.convert(iv, INT, n).comps[0])
            src = self.materialize(base)
            count, width = (ty.cols, 1) if ty.kind == "vector" else (ty.rows, ty.cols)
            comps = []
            for c in range(width):
                e = src.comps[(count - 1) * width + c]
                for r in range(count - 2, -1, -1):
                    e = f"({i} == {r} ? {src.comps[r * width + c]} : {e})"
                comps.append(self.temp(ty.base, e))
            return Value(vec(ty.base, width), comps)
        digits = idx.args[0].rstrip("uUlL")
        if re.fullmatch(r"0[xX][0-9a-fA-F]+", digits):
            i = int(digits, 16)
        elif re.fullmatch(r"\d+", digits):
            i = int(digits)
        else:
            self.err(n, "an index must be an integer literal")
        if ty.kind == "vector":
            if i >= ty.n:
                self.err(n, "index out of range")
            return Value(Type("scalar", ty.base), [base.comps[i]], base.lvalue)
        if ty.kind == "matrix":
            if i >= ty.rows:
                self.err(n, "index out of range")
            return Value(vec(ty.base, ty.cols), base.comps[i * ty.cols:(i + 1) * ty.cols], base.lvalue)
        self.err(n, f"cannot index {ty}")

    def e_un(self, n):
        op, v = n.args[0], self.expr(n.args[1])
        if v.type.kind not in ("scalar", "vector", "matrix"):
            self.err(n, f"unary {op} on {v.type}")
        if op == "+":
            return v
        if op == "!":
            v = self.convert(v, Type(v.type.kind, "bool", v.type.rows, v.type.cols), n)
            return Value(v.type, [self.temp("bool", f"!{c}") for c in v.comps])
        if op == "~":
            return Value(v.type, [self.temp(v.type.base, f"~{c}") for c in v.comps])
        base = "int" if v.type.base == "bool" else v.type.base
        v = self.convert(v, Type(v.type.kind, base, v.type.rows, v.type.cols), n)
        return Value(v.type, [self.temp(base, f"-{c}") for c in v.comps])

    def e_bin(self, n):
        op = n.args[0]
        if op in ("&&", "||"):  # no side effects in operands of the supported subset: evaluate both
            a, b = self.expr(n.args[1]), self.expr(n.args[2])
            if a.type.kind in ("vector", "matrix") or b.type.kind in ("vector", "matrix"):  # component-wise on vectors / matrices
                shape = a.type if a.type.kind in ("vector", "matrix") else b.type
                bt = Type(shape.kind, "bool", shape.rows, shape.cols)
                a, b = self.convert(a, bt, n), self.convert(b, bt, n)
                return Value(bt, [self.temp("bool", f"{x} {op} {y}") for x, y in zip(a.comps, b.comps)])
            a = self.convert(a, BOOL, n)
            b = self.convert(b, BOOL, n)
            return Value(BOOL, [self.temp("bool", f"{a.comps[0]} {op} {b.comps[0]}")])
        a, b = self.expr(n.args[1]), self.expr(n.args[2])
        if op in ("==", "!=", "<", ">", "<=", ">="):
            a, b, to = self.unify(a, b, n, arith=False)
            rt = Type(to.kind, "bool", to.rows, to.cols)
            return Value(rt, [self.temp("bool", f"{x} {op} {y}") for x, y in zip(a.comps, b.comps)])
        a, b, to = self.unify(a, b, n)
        if op in ("%", "&", "|", "^", "<<", ">>") and to.base == "float":
            if op != "%":
                self.err(n, f"operator {op} on floating-point operands")
            return Value(to, [self.temp("float", f"fmodf({x}, {y})") for x, y in zip(a.comps, b.comps)])
        return Value(to, [self.temp(to.base, f"{x} {op} {y}") for x, y in zip(a.comps, b.comps)])

    def e_select(self, n):
        c = self.expr(n.args[0])
        a, b = self.expr(n.args[1]), self.expr(n.args[2])
        if a.type.kind == "struct":
            self.err(n, "?: on structs")
        a, b, to = self.unify(a, b, n, arith=False)
        if c.type.kind not in ("scalar", "vector", "matrix") or (c.type.n != 1 and c.type.n != to.n):
            self.err(n, f"?: condition of type {c.type} with operands of type {to}")
        c = self.convert(c, Type("scalar" if c.type.n == 1 else to.kind, "bool", 1 if c.type.n == 1 else to.rows, c.type.cols if c.type.n > 1 else 1), n)
        cs = c.comps * to.n if c.type.n == 1 else c.comps
        return Value(to, [self.temp(to.base, f"{k} ? {x} : {y}") for k, x, y in zip(cs, a.comps, b.comps)])

    def e_assign(self, n):
        op, lhs_n, rhs_n = n.args
        rhs = self.expr(rhs_n)
        lhs = self.expr(lhs_n)
        if not lhs.lvalue:
            self.err(n, "left side of an assignment is not assignable")
        if op != "=":
            rhs = self.e_bin(Node("bin", (op[:-1], _Lit(self.materialize(lhs)), _Lit(rhs)), n.line))
        self.store(lhs, rhs, n)
        return lhs

    def e_postinc(self, n):
        v = self.expr(n.args[0])
        if not v.lvalue:
            self.err(n, "operand of ++/-- is not assignable")
        old = self.materialize(v)
        one = Value(INT, ["1"])
        self.store(v, self.e_bin(Node("bin", (n.args[1], _Lit(old), _Lit(one)), n.line)), n)
        return old

    def e_lit(self, n):
        return n.args[0]

    # ---- calls: constructors, intrinsics, user functions
    def e_call(self, n):
        name, args_n = n.args
        ty = parse_type_name(name)
        if ty is not None:
            return self.construct(ty, [self.expr(a) for a in args_n], n)
        if name in self.fn_table:
            return self.call_user(self.fn_table[name], [self.expr(a) for a in args_n], n)
        fn = getattr(self, "i_" + name, None)
        args = [self.expr(a) for a in args_n]
        if name in UNARY_MATH:
            return self.map1(args, n, lambda c: f"{UNARY_MATH[name]}({c})")
        if fn is None:
            self.err(n, f"unknown function {name!r}")
        return fn(args, n)

    def construct(self, ty, args, n):
        if len(args) == 1 and args[0].type.kind in ("scalar", "vector", "matrix") and (args[0].type.n == 1 or args[0].type.n >= ty.n):
            return self.convert(args[0], ty, n, explicit=True)
        comps = []
        for a in args:
            if a.type.kind not in ("scalar", "vector", "matrix"):
                self.err(n, f"constructor argument of type {a.type}")
            a = self.convert(a, Type(a.type.kind, ty.base, a.type.rows, a.type.cols), n)
            comps += a.comps
        if len(comps) != ty.n:
            self.err(n, f"{ty} constructed from {len(comps)} components")
        return Value(ty, comps)

    def call_user(self, f: Func, args, n):
        if len(args) != len(f.params):
            self.err(n, f"{f.name} expects {len(f.params)} arguments")
        actual = []
        for a, prm in zip(args, f.params):
            if prm.type.kind == "sampler":
                if a.type.kind != "sampler":
                    self.err(n, "sampler argument expected")
                actual += a.comps
            else:
                actual += self.materialize(self.convert(a, prm.type, n) if prm.type.kind != "struct" else a).comps
        rets = []
        if f.ret.kind != "void":
            for b in self.flat_types(f.ret):
                self.ntemp += 1
                rets.append(f"r{self.ntemp}")
                self.emit(f"{C_BASE[b]} {rets[-1]} = 0;")
        if self.stage == "ps" and self.divergent and f.name in self.fns_with_derivatives:
            self.err(n, f"{f.name} takes screen-space derivatives and is called under divergent control flow")
        self.emit(f"sasl_fn_{f.name}({', '.join(self.ctx_args() + actual + rets)});")
        return Value(f.ret, rets)

    def ctx_args(self):
        return ["U", "p", "px"] if self.stage == "ps" else ["U", "S0"]

    def map1(self, args, n, fmt, base="float"):
        if len(args) != 1:
            self.err(n, "expects 1 argument")
        v = self.to_base(args[0], base, n)
        return Value(v.type, [self.temp(v.type.base, fmt(c)) for c in v.comps])

    def to_base(self, v, base, n):
        if v.type.kind not in ("scalar", "vector", "matrix"):
            self.err(n, f"argument of type {v.type}")
        return self.convert(v, Type(v.type.kind, base, v.type.rows, v.type.cols), n)

    def mapn(self, args, n, fmt, count):
        if len(args) != count:
            self.err(n, f"expects {count} arguments")
        vs = [self.to_base(a, "float", n) for a in args]
        shape = max(vs, key=lambda v: v.type.n).type
        if any(v.type.n not in (1, shape.n) for v in vs):
            shape = min((v.type for v in vs if v.type.n > 1), key=lambda t: t.n)
        vs = [self.convert(v, shape, n) for v in vs]
        return Value(shape, [self.temp("float", fmt(*cs)) for cs in zip(*[v.comps for v in vs])])

    def i_abs(self, a, n):
        if len(a) == 1 and a[0].type.base in ("int", "uint", "bool"):
            return self.map1(a, n, lambda c: f"abs({c})", "int")
        return self.map1(a, n, lambda c: f"fabsf({c})")

    def i_rsqrt(self, a, n): return self.map1(a, n, lambda c: f"(1.0f / sqrtf({c}))")
    def i_frac(self, a, n):  # the reference's definition: |v| - floor(|v|) (sasl/src/codegen/cg_impl.cpp:1139-1150)
        return self.map1(a, n, lambda c: f"(fabsf({c}) - sasl_m_floor(fabsf({c})))")

    def i_ldexp(self, a, n): return self.mapn(a, n, lambda x, y: f"sasl_m_ldexp({x}, {y})", 2)
    def i_saturate(self, a, n): return self.map1(a, n, lambda c: f"sasl_clamp({c}, 0.0f, 1.0f)")
    def i_sign(self, a, n): return self.map1(a, n, lambda c: f"(({c} > 0.0f) ? 1.0f : (({c} < 0.0f) ? -1.0f : 0.0f))")
    def i_radians(self, a, n): return self.map1(a, n, lambda c: f"({c} * 0.017453292519943295f)")
    def i_degrees(self, a, n): return self.map1(a, n, lambda c: f"({c} * 57.29577951308232f)")
    def i_min(self, a, n): return self.mapn(a, n, lambda x, y: f"fminf({x}, {y})", 2)
    def i_max(self, a, n): return self.mapn(a, n, lambda x, y: f"fmaxf({x}, {y})", 2)
    def i_pow(self, a, n): return self.mapn(a, n, lambda x, y: f"sasl_m_pow({x}, {y})", 2)
    def i_fmod(self, a, n): return self.mapn(a, n, lambda x, y: f"fmodf({x}, {y})", 2)
    def i_atan2(self, a, n): return self.mapn(a, n, lambda x, y: f"sasl_m_atan2({x}, {y})", 2)
    def i_step(self, a, n): return self.mapn(a, n, lambda e, x: f"(({x} >= {e}) ? 1.0f : 0.0f)", 2)
    def i_clamp(self, a, n): return self.mapn(a, n, lambda x, lo, hi: f"sasl_clamp({x}, {lo}, {hi})", 3)
    def i_mad(self, a, n): return self.mapn(a, n, lambda x, y, z: f"(({x} * {y}) + {z})", 3)
    def i_smoothstep(self, a, n): return self.mapn(a, n, lambda lo, hi, x: f"sasl_smoothstep({lo}, {hi}, {x})", 3)

    def i_lerp(self, a, n):
        if len(a) != 3:
            self.err(n, "lerp expects 3 arguments")
        d = self.e_bin(Node("bin", ("-", _Lit(a[1]), _Lit(a[0])), n.line))
        return self.e_bin(Node("bin", ("+", _Lit(a[0]), _Lit(self.e_bin(Node("bin", ("*", _Lit(d), _Lit(a[2])), n.line)))), n.line))

    def sum_lr(self, terms):
        acc = terms[0]
        for t in terms[1:]:
            acc = self.temp("float", f"{acc} + {t}")
        return acc

    def dot_comps(self, a, b):
        return self.sum_lr([self.temp("float", f"{x} * {y}") for x, y in zip(a, b)])

    def i_dot(self, a, n):
        if len(a) != 2:
            self.err(n, "dot expects 2 arguments")
        x, y, _ = self.unify(self.to_base(a[0], "float", n), self.to_base(a[1], "float", n), n)
        return Value(FLOAT, [self.dot_comps(x.comps, y.comps)])

    def i_cross(self, a, n):
        if len(a) != 2:
            self.err(n, "cross expects 2 arguments")
        x, y = (self.convert(self.to_base(v, "float", n), vec("float", 3), n) for v in a)
        (ax, ay, az), (bx, by, bz) = x.comps, y.comps
        return Value(vec("float", 3), [self.temp("float", f"({ay} * {bz}) - ({az} * {by})"), self.temp("float", f"({az} * {bx}) - ({ax} * {bz})"),
                                       self.temp("float", f"({ax} * {by}) - ({ay} * {bx})")])

    def i_dst(self, a, n):  # distance vector: (1, a.y * b.y, a.z, b.w)
        if len(a) != 2:
            self.err(n, "dst(float4, float4)")
        x, y = (self.convert(self.to_base(v, "float", n), vec("float", 4), n) for v in a)
        return Value(vec("float", 4), [self.temp("float", "1.0f"), self.temp("float", f"{x.comps[1]} * {y.comps[1]}"), x.comps[2], y.comps[3]])

    def i_length(self, a, n):
        if not a:
            self.err(n, "length expects 1 argument")
        v = self.to_base(a[0], "float", n)
        return Value(FLOAT, [self.temp("float", f"sqrtf({self.dot_comps(v.comps, v.comps)})")])

    def i_distance(self, a, n):
        if len(a) != 2:
            self.err(n, "distance expects 2 arguments")
        return self.i_length([self.e_bin(Node("bin", ("-", _Lit(a[0]), _Lit(a[1])), n.line))], n)

    def i_normalize(self, a, n):  # eflib normalize3: zero-length vectors are left alone (length := 1)
        if not a:
            self.err(n, "normalize expects 1 argument")
        v = self.to_base(a[0], "float", n)
        ln = self.temp("float", f"sqrtf({self.dot_comps(v.comps, v.comps)})")
        ln = self.temp("float", f"sasl_eq_eps({ln}, 0.0f) ? 1.0f : {ln}")
        inv = self.temp("float", f"1.0f / {ln}")
        return Value(v.type, [self.temp("float", f"{c} * {inv}") for c in v.comps])

    def i_reflect(self, a, n):  # eflib reflect3(i, n) = i - 2 * dot(i, n) * n
        if len(a) != 2:
            self.err(n, "reflect expects 2 arguments")
        i, nn, _ = self.unify(self.to_base(a[0], "float", n), self.to_base(a[1], "float", n), n)
        d = self.dot_comps(i.comps, nn.comps)
        s = self.temp("float", f"2.0f * {d}")
        return Value(i.type, [self.temp("float", f"{x} - ({s} * {y})") for x, y in zip(i.comps, nn.comps)])

    def i_mul(self, a, n):
        if len(a) != 2:
            self.err(n, "mul expects 2 arguments")
        x, y = self.to_base(a[0], "float", n), self.to_base(a[1], "float", n)
        tx, ty = x.type, y.type
        if tx.kind != "matrix" and ty.kind != "matrix":
            return self.e_bin(Node("bin", ("*", _Lit(x), _Lit(y)), n.line))
        if tx.kind == "vector" and ty.kind == "matrix":   # row vector x matrix (eflib transform)
            if tx.cols != ty.rows:
                self.err(n, f"mul({tx}, {ty})")
            out = [self.sum_lr([self.temp("float", f"{x.comps[i]} * {y.comps[i * ty.cols + j]}") for i in range(ty.rows)]) for j in range(ty.cols)]
            return Value(vec("float", ty.cols), out)
        if tx.kind == "matrix" and ty.kind == "vector":
            if tx.cols != ty.cols:
                self.err(n, f"mul({tx}, {ty})")
            out = [self.sum_lr([self.temp("float", f"{x.comps[i * tx.cols + j]} * {y.comps[j]}") for j in range(tx.cols)]) for i in range(tx.rows)]
            return Value(vec("float", tx.rows), out)
        if tx.kind == "matrix" and ty.kind == "matrix":
            if tx.cols != ty.rows:
                self.err(n, f"mul({tx}, {ty})")
            out = [self.sum_lr([self.temp("float", f"{x.comps[i * tx.cols + k]} * {y.comps[k * ty.cols + j]}") for k in range(tx.cols)])
                   for i in range(tx.rows) for j in range(ty.cols)]
            return Value(mat("float", tx.rows, ty.cols), out)
        return self.e_bin(Node("bin", ("*", _Lit(x), _Lit(y)), n.line))  # scalar * matrix

    def i_transpose(self, a, n):
        if not a or a[0].type.kind != "matrix":
            self.err(n, "transpose expects a matrix")
        m = a[0]
        return Value(mat(m.type.base, m.type.cols, m.type.rows), [m.comps[r * m.type.cols + c] for c in range(m.type.cols) for r in range(m.type.rows)])

    def i_any(self, a, n):
        if not a:
            self.err(n, "any expects 1 argument")
        v = self.to_base(a[0], "bool", n)
        return Value(BOOL, [self.temp("bool", " || ".join(v.comps))])

    def i_all(self, a, n):
        if not a:
            self.err(n, "all expects 1 argument")
        v = self.to_base(a[0], "bool", n)
        return Value(BOOL, [self.temp("bool", " && ".join(v.comps))])

    def i_asfloat(self, a, n): return self.bitcast(a, n, "float", "sasl_asfloat")
    def i_asint(self, a, n): return self.bitcast(a, n, "int", "sasl_asint")
    def i_asuint(self, a, n): return self.bitcast(a, n, "uint", "sasl_asuint")

    def bitcast(self, a, n, base, fn):
        if not a:
            self.err(n, f"{fn[5:]} expects 1 argument")
        v = a[0]
        return Value(Type(v.type.kind, base, v.type.rows, v.type.cols), [self.temp(base, f"{fn}({c})") for c in v.comps])

    def i_countbits(self, a, n):
        if not a:
            self.err(n, "countbits expects 1 argument")
        v = self.to_base(a[0], "uint", n)
        return Value(v.type, [self.temp("uint", f"sasl_countbits({c})") for c in v.comps])

    i_count_bits = i_countbits  # both spellings are registered upstream (semantic_analyser.cpp:1981-1982)

    def bits1(self, a, n, fn):  # sasl.firstbithigh / firstbitlow / reversebits .u32; the result keeps the argument's int / uint base
        if len(a) != 1 or a[0].type.base not in ("int", "uint"):
            self.err(n, "expects one int or uint argument")
        v = a[0]
        return Value(v.type, [self.temp(v.type.base, f"({'int' if v.type.base == 'int' else 'unsigned'}){fn}((unsigned){c})") for c in v.comps])

    def i_firstbithigh(self, a, n): return self.bits1(a, n, "sasl_firstbithigh")
    def i_firstbitlow(self, a, n): return self.bits1(a, n, "sasl_firstbitlow")
    def i_reversebits(self, a, n): return self.bits1(a, n, "sasl_reversebits")

    def fclass(self, a, n, fmt):
        if len(a) != 1:
            self.err(n, "expects 1 argument")
        v = self.to_base(a[0], "float", n)
        return Value(Type(v.type.kind, "bool", v.type.rows, v.type.cols), [self.temp("bool", fmt(c)) for c in v.comps])

    # cgs.cpp:1828-1842: isinf = |v| == inf, isfinite = !isinf && v == v, isnan = unordered(v, v)
    def i_isinf(self, a, n): return self.fclass(a, n, lambda c: f"(fabsf({c}) == sasl_asfloat(0x7F800000u))")
    def i_isfinite(self, a, n): return self.fclass(a, n, lambda c: f"(!(fabsf({c}) == sasl_asfloat(0x7F800000u)) && ({c} == {c}))")
    def i_isnan(self, a, n): return self.fclass(a, n, lambda c: f"({c} != {c})")

    def i_rcp(self, a, n): return self.map1(a, n, lambda c: f"(1.0f / {c})")

    def i_refract(self, a, n):  # cg_impl.cpp:1229-1269, in the reference's order of operations
        if len(a) != 3:
            self.err(n, "refract(i, n, eta)")
        i, nn, _ = self.unify(self.to_base(a[0], "float", n), self.to_base(a[1], "float", n), n)
        eta = self.convert(self.to_base(a[2], "float", n), FLOAT, n).comps[0]
        eta2 = self.temp("float", f"{eta} * {eta}")
        ndi = self.dot_comps(nn.comps, i.comps)
        eta_i = [self.temp("float", f"{eta} * {c}") for c in i.comps]
        k = self.temp("float", f"{ndi} * {ndi}")
        k = self.temp("float", f"1.0f - {k}")
        k = self.temp("float", f"{eta2} * {k}")
        k = self.temp("float", f"1.0f - {k}")
        flag = self.temp("bool", f"{k} < 0.0f")
        k = self.temp("float", f"{flag} ? 0.0f : {k}")
        r = self.temp("float", f"{eta} * {ndi}")
        r = self.temp("float", f"{r} + sqrtf({k})")
        out = []
        for ei, c in zip(eta_i, nn.comps):
            t = self.temp("float", f"{r} * {c}")
            t = self.temp("float", f"{ei} - {t}")
            out.append(self.temp("float", f"{flag} ? 0.0f : {t}"))
        return Value(i.type, out)

    def i_faceforward(self, a, n):  # cg_impl.cpp:1300-1319: dot(i, ng) < 0 ? n : 0 - n
        if len(a) != 3:
            self.err(n, "faceforward(n, i, ng)")
        nn, i, _ = self.unify(self.to_base(a[0], "float", n), self.to_base(a[1], "float", n), n)
        ng = self.convert(self.to_base(a[2], "float", n), i.type, n)
        d = self.dot_comps(i.comps, ng.comps)
        flag = self.temp("bool", f"{d} < 0.0f")
        return Value(nn.type, [self.temp("float", f"{flag} ? {c} : (0.0f - {c})") for c in nn.comps])

    def i_lit(self, a, n):  # cg_impl.cpp:1320-1352: (1, max(n.l, 0), n.l < 0 || n.h < 0 ? 0 : n.h * m, 1)
        if len(a) != 3:
            self.err(n, "lit(n_dot_l, n_dot_h, m)")
        l, h, m = (self.convert(self.to_base(v, "float", n), FLOAT, n).comps[0] for v in a)
        diffuse = self.temp("float", f"({l} < 0.0f) ? 0.0f : {l}")
        spec = self.temp("float", f"(({l} < 0.0f) || ({h} < 0.0f)) ? 0.0f : ({h} * {m})")
        return Value(vec("float", 4), [self.temp("float", "1.0f"), diffuse, spec, self.temp("float", "1.0f")])

    # ---- screen-space derivatives and texture sampling (pixel shaders)
    def need_quad(self, n, what):
        if self.stage != "ps":
            self.err(n, f"{what} is only available in pixel shaders")
        if self.divergent:
            self.err(n, f"{what} under divergent control flow is not supported (the four pixels of a quad must reach it together)")
        self.refl.uses_derivatives = True
        self.cur_fn_derivs = True

    def i_ddx(self, a, n):
        self.need_quad(n, "ddx")
        return self.map1(a, n, lambda c: f"sasl_ddx(px, {c})")

    def i_ddy(self, a, n):
        self.need_quad(n, "ddy")
        return self.map1(a, n, lambda c: f"sasl_ddy(px, {c})")

    def sampler_slot(self, v, n):
        if v.type.kind != "sampler":
            self.err(n, "first argument must be a sampler")
        return v.comps[0]

    def tex_result(self, call):
        self.ntemp += 1
        r = [f"x{self.ntemp}_{k}" for k in range(4)]
        self.emit(f"float {r[0]}, {r[1]}, {r[2]}, {r[3]};")
        self.emit(call(r))
        return Value(vec("float", 4), r)

    def i_tex2D(self, a, n):  # SASL tex2D == sample_2d_grad with the quad's per-pixel derivatives
        if len(a) != 2:
            self.err(n, "tex2D(sampler, uv)")
        if self.stage == "vs":
            self.err(n, "vertex shaders sample with tex2Dlod")
        self.need_quad(n, "tex2D")
        s = self.sampler_slot(a[0], n)
        uv = self.convert(self.to_base(a[1], "float", n), vec("float", 2), n)
        return self.tex_result(lambda r: f"sasl_tex2d_grad(p, px, {s}, {uv.comps[0]}, {uv.comps[1]}, sasl_ddx(px, {uv.comps[0]}), sasl_ddx(px, {uv.comps[1]}), "
                                         f"sasl_ddy(px, {uv.comps[0]}), sasl_ddy(px, {uv.comps[1]}), 0.0f, {', '.join(r)});")

    def i_tex2Dgrad(self, a, n):
        if self.stage != "ps" or len(a) != 4:
            self.err(n, "tex2Dgrad(sampler, uv, ddx, ddy) in a pixel shader")
        s = self.sampler_slot(a[0], n)
        uv, dx, dy = (self.convert(self.to_base(v, "float", n), vec("float", 2), n) for v in a[1:])
        return self.tex_result(lambda r: f"sasl_tex2d_grad(p, px, {s}, {uv.comps[0]}, {uv.comps[1]}, {dx.comps[0]}, {dx.comps[1]}, {dy.comps[0]}, {dy.comps[1]}, 0.0f, {', '.join(r)});")

    def i_tex2Dbias(self, a, n):
        if len(a) != 2:
            self.err(n, "tex2Dbias(sampler, float4(uv, _, bias))")
        self.need_quad(n, "tex2Dbias")
        s = self.sampler_slot(a[0], n)
        c = self.convert(self.to_base(a[1], "float", n), vec("float", 4), n)
        return self.tex_result(lambda r: f"sasl_tex2d_grad(p, px, {s}, {c.comps[0]}, {c.comps[1]}, sasl_ddx(px, {c.comps[0]}), sasl_ddx(px, {c.comps[1]}), "
                                         f"sasl_ddy(px, {c.comps[0]}), sasl_ddy(px, {c.comps[1]}), {c.comps[3]}, {', '.join(r)});")

    def i_tex2Dlod(self, a, n):
        if len(a) != 2:
            self.err(n, "tex2Dlod(sampler, float4(uv, _, lod))")
        s = self.sampler_slot(a[0], n)
        c = self.convert(self.to_base(a[1], "float", n), vec("float", 4), n)
        if self.stage == "vs":  # sasl.vs.tex2d.lod = sampler::sample_2d_lod(coord.xy, coord.w) (sampler_api.cpp:50-52)
            return self.tex_result(lambda r: f"sasl_vs_tex2d_lod(S0, {s}, {c.comps[0]}, {c.comps[1]}, {c.comps[3]}, {', '.join(r)});")
        return self.tex_result(lambda r: f"sasl_tex2d_lod(p, px, {s}, {c.comps[0]}, {c.comps[1]}, {c.comps[3]}, {', '.join(r)});")

    def i_tex2Dproj(self, a, n):
        if len(a) != 2:
            self.err(n, "tex2Dproj(sampler, float4)")
        c = self.convert(self.to_base(a[1], "float", n), vec("float", 4), n)
        inv = self.temp("float", f"1.0f / {c.comps[3]}")
        uv = Value(vec("float", 2), [self.temp("float", f"{c.comps[0]} * {inv}"), self.temp("float", f"{c.comps[1]} * {inv}")])
        return self.i_tex2D([a[0], uv], n)

    # ---- statements
    def stmt(self, n: Node):
        getattr(self, "s_" + n.op)(n)

    def s_block(self, n):
        self.scopes.append({})
        for s in n.args:
            self.stmt(s)
        self.scopes.pop()

    def s_decl(self, n):
        for d in n.args:
            if d.type.kind in ("void", "sampler"):
                self.err(n, f"cannot declare a local of type {d.type}")
            init = self.materialize(self.expr(d.init)) if d.init is not None else None
            v = self.declare(d.type, d.name, n)
            if init is not None:
                self.store(v, init, n)

    def s_expr(self, n):
        self.expr(n.args[0])

    def s_if(self, n):
        c = self.convert(self.expr(n.args[0]), BOOL, n)
        self.emit(f"if ({c.comps[0]}) {{")
        self.indent += 1
        self.divergent += 1
        self.s_block(Node("block", (n.args[1],), n.line))
        self.indent -= 1
        if n.args[2] is not None:
            self.emit("} else {")
            self.indent += 1
            self.s_block(Node("block", (n.args[2],), n.line))
            self.indent -= 1
        self.divergent -= 1
        self.emit("}")

    def loop_body(self, body):
        """Body inside `do { } while (0)`: `continue` leaves it (the step still runs), `break` raises the loop's flag."""
        self.emit("do {")
        self.indent += 1
        self.s_block(Node("block", (body,), body.line))
        self.indent -= 1
        self.emit("} while (0);")
        self.emit(f"if (brk{self.loop_ids[-1]}) break;")

    def s_for(self, n):
        init, cond, step, body = n.args
        self.scopes.append({})
        self.emit("{")
        self.indent += 1
        if init is not None:
            self.stmt(init)
        self.ntemp += 1
        self.loop_ids = getattr(self, "loop_ids", []) + [self.ntemp]
        self.emit(f"bool brk{self.ntemp} = false;")
        self.emit("for (;;) {")
        self.indent += 1
        self.divergent += 1
        if cond is not None:
            c = self.convert(self.expr(cond), BOOL, n)
            self.emit(f"if (!{c.comps[0]}) break;")
        self.loop_body(body)
        if step is not None:
            self.expr(step)
        self.divergent -= 1
        self.indent -= 1
        self.emit("}")
        self.loop_ids = self.loop_ids[:-1]
        self.indent -= 1
        self.emit("}")
        self.scopes.pop()

    def s_dowhile(self, n):
        body, cond = n.args
        self.ntemp += 1
        self.loop_ids = getattr(self, "loop_ids", []) + [self.ntemp]
        self.emit(f"bool brk{self.ntemp} = false;")
        self.emit("for (;;) {")
        self.indent += 1
        self.divergent += 1
        self.loop_body(body)
        c = self.convert(self.expr(cond), BOOL, n)
        self.emit(f"if (!{c.comps[0]}) break;")
        self.divergent -= 1
        self.indent -= 1
        self.emit("}")
        self.loop_ids = self.loop_ids[:-1]

    def case_value(self, e, n):
        """Case labels are integer literals (optionally negated)."""
        neg = False
        while e.op == "un" and e.args[0] in ("-", "+"):
            neg ^= e.args[0] == "-"
            e = e.args[1]
        t = e.args[0] if e.op == "num" else ""
        if t.lower().startswith("0x") and re.fullmatch(r"0[xX][0-9a-fA-F]+[uU]?", t):
            v = int(t.rstrip("uU"), 16)
        elif re.fullmatch(r"\d+[uUlL]?", t):
            v = int(t.rstrip("uUlL"))
        else:
            self.err(n, "case labels must be integer literals")
        return -v if neg else v

    def s_switch(self, n):
        """switch with C fall-through semantics, lowered to guarded blocks inside one `do { } while (0)`: a group runs when
        an earlier group fell through into it or the selector equals one of its labels (`default`: none of the switch's
        labels); `break` leaves the do-while, `continue` (inside a loop) leaves it with the loop's continue flag raised."""
        sel_e, groups = n.args
        sel = self.convert(self.expr(sel_e), INT, n)
        self.ntemp += 1
        sid = self.ntemp
        labels = [[None if l is None else self.case_value(l, n) for l in ls] for ls, _ in groups]
        all_vals = [v for ls in labels for v in ls if v is not None]
        if len(set(all_vals)) != len(all_vals):
            self.err(n, "duplicate case label")
        if sum(1 for ls in labels for v in ls if v is None) > 1:
            self.err(n, "more than one default label")
        self.emit("{")
        self.indent += 1
        self.emit(f"const int sel{sid} = {sel.comps[0]};")
        self.emit(f"bool fall{sid} = false;")
        if getattr(self, "loop_ids", []):
            self.emit(f"bool cnt{sid} = false;")
        self.emit("do {")
        self.indent += 1
        self.divergent += 1
        self.switch_ids = getattr(self, "switch_ids", []) + [(sid, len(getattr(self, "loop_ids", [])))]
        for ls, (_, stmts) in zip(labels, groups):
            conds = [f"sel{sid} == {v}" for v in ls if v is not None]
            if None in ls:
                conds.append("(" + " && ".join([f"sel{sid} != {v}" for v in all_vals] or ["true"]) + ")")
            self.emit(f"if (fall{sid} || {' || '.join(conds)}) {{")
            self.indent += 1
            self.emit(f"fall{sid} = true;")
            self.s_block(Node("block", stmts, n.line))
            self.indent -= 1
            self.emit("}")
        self.switch_ids = self.switch_ids[:-1]
        self.divergent -= 1
        self.indent -= 1
        self.emit("} while (0);")
        if getattr(self, "loop_ids", []):
            self.emit(f"if (cnt{sid}) break;")  # `continue` inside the switch: leave the loop body's do { } while (0)
        self.indent -= 1
        self.emit("}")

    def in_switch(self):
        """The innermost breakable construct is a switch (opened after the innermost loop)."""
        sw = getattr(self, "switch_ids", [])
        return bool(sw) and sw[-1][1] == len(getattr(self, "loop_ids", []))

    def s_break(self, n):
        if self.in_switch():
            self.emit("break;")  # leaves the switch's do { } while (0)
            return
        if not getattr(self, "loop_ids", []):
            self.err(n, "break outside a loop or switch")
        self.emit(f"brk{self.loop_ids[-1]} = true; break;")

    def s_continue(self, n):
        if not getattr(self, "loop_ids", []):
            self.err(n, "continue outside a loop")
        if self.in_switch():
            for sid, depth in self.switch_ids:  # every switch opened inside the innermost loop hands the request outwards
                if depth == len(self.loop_ids):
                    self.emit(f"cnt{sid} = true;")
        self.emit("break;")  # leaves the do { } while (0) around the body; the loop's step still runs

    def s_return(self, n):
        if self.divergent and self.stage == "ps":
            self.fn_early_return = True
        e = n.args[0]
        if e is not None:
            if self.cur_ret is None:
                self.err(n, "void function returns a value")
            self.store(self.cur_ret, self.expr(e), n)
        self.emit("return;")

    # ---- functions
    def called_functions(self, f: Func) -> set:
        """Names of the user functions f's body calls."""
        names = {g.name for g in self.funcs}
        out = set()

        def walk(x):
            if isinstance(x, Node):
                if x.op == "call" and x.args[0] in names:
                    out.add(x.args[0])
                for a in x.args:
                    walk(a)
            elif isinstance(x, (tuple, list)):
                for a in x:
                    walk(a)
            elif isinstance(x, VarDecl):
                walk(x.init)

        walk(f.body)
        return out

    def generation_order(self):
        """Callees before callers (a function may be used before its definition), restricted to what the entry reaches
        (every function for a library unit); functions on a call cycle are returned as `recursive`: they get a prototype
        ahead of all bodies and are not force-inlined."""
        by_name = {f.name: f for f in self.funcs}
        callees = {f.name: self.called_functions(f) for f in self.funcs}
        order, state, stack, recursive = [], {}, [], set()

        def dfs(f):
            state[f.name] = 1
            stack.append(f.name)
            for c in sorted(callees[f.name]):
                if state.get(c, 0) == 0:
                    dfs(by_name[c])
                elif state[c] == 1:
                    recursive.update(stack[stack.index(c):])
            stack.pop()
            state[f.name] = 2
            order.append(f)

        for f in ([self.entry] if self.entry is not None else self.funcs):
            if state.get(f.name, 0) == 0:
                dfs(by_name[f.name])
        return order, recursive

    def fn_head(self, f: Func, bases_params):
        ctx = (["const SaslUniforms& U", "const slv::RasterParams& p", "const Ctx& px"] if self.stage == "ps"
               else ["const SaslUniforms& U", "const SaslSampler& S0"])
        kw = "SASL_FN_REC" if f.name in getattr(self, "recursive", ()) else "SASL_FN"
        return ("template <class Ctx>\n" if self.stage == "ps" else "") + f"{kw} void sasl_fn_{f.name}({', '.join(ctx + bases_params)})"

    def fn_signature(self, f: Func):
        """The flattened C++ parameter list of f: (parameter declarations, scope of the parameters, names of the results)."""
        bases_params = []
        scope = {}
        for prm in f.params:
            if prm.type.kind == "sampler":
                nm = f"a_{prm.name}"
                bases_params.append(f"const int {nm}")
                scope[prm.name] = Value(SAMPLER, [nm])
                continue
            names = []
            for k, b in enumerate(self.flat_types(prm.type)):
                nm = f"a_{prm.name}_{k}"
                names.append(nm)
                bases_params.append(f"{C_BASE[b]} {nm}")
            scope[prm.name] = Value(prm.type, names, True)
        ret_names = []
        if f.ret.kind != "void":
            for k, b in enumerate(self.flat_types(f.ret)):
                ret_names.append(f"ret_{k}")
                bases_params.append(f"{C_BASE[b]}& ret_{k}")
        return bases_params, scope, ret_names

    def gen_function(self, f: Func):
        self.cur_fn_derivs = False
        if f.name in getattr(self, "recursive", ()):
            self.fn_table[f.name] = f  # visible to its own body (and to the other members of its cycle)
        bases_params, scope, ret_names = self.fn_signature(f)
        self.cur_ret = Value(f.ret, ret_names, True) if ret_names else None
        head = self.fn_head(f, bases_params) + " {"
        start = len(self.lines)
        self.indent = 1
        self.scopes = [scope]
        # a global the function assigns to (sasl/test/repo/input_assigned.svs: `x += 0.5f`): the uniform block is read-only and
        # shared, so the function works on its own copy, initialised from the uniform - the write is local to the invocation
        for name in self.assigned_globals(f.body, set(scope)):
            src = self.uniform_vars[name]
            self.store(self.declare(src.type, name, f), src, f)
        self.s_block(f.body)
        body = self.lines[start:]
        del self.lines[start:]
        self.lines.append(head)
        self.lines += body
        self.lines.append("}")
        self.lines.append("")
        if self.cur_fn_derivs:
            if f.name in getattr(self, "recursive", ()):
                raise CompileError(f"line {f.line}: {f.name}: screen-space derivatives in a recursive function")
            self.fns_with_derivatives.add(f.name)
        self.fn_table[f.name] = f

    def assigned_globals(self, body, local_names) -> list:
        """Names of uniform globals (not samplers, not arrays) that `body` stores to, in order of first store; names shadowed by
        a parameter or by a declaration anywhere in the function are left alone (conservative)."""
        declared, out = set(local_names), []

        def root(n):
            while isinstance(n, Node) and n.op in ("member", "index"):
                n = n.args[0]
            return n.args[0] if isinstance(n, Node) and n.op == "var" else None

        def walk(n):
            if isinstance(n, VarDecl):
                declared.add(n.name)
                if n.init is not None:
                    walk(n.init)
                return
            if isinstance(n, (tuple, list)):
                for x in n:
                    walk(x)
                return
            if not isinstance(n, Node):
                return
            if n.op == "assign":
                r = root(n.args[1])
                if r is not None and r not in out:
                    out.append(r)
            elif n.op == "postinc":
                r = root(n.args[0])
                if r is not None and r not in out:
                    out.append(r)
            walk(n.args)
        walk(body)
        return [r for r in out if r not in declared and r in self.uniform_vars and r not in self.uniform_arrays
                and self.uniform_vars[r].type.kind in ("scalar", "vector", "matrix")]

    # ---- translation unit
    def run(self) -> ShaderUnit:
        self.fns_with_derivatives: set[str] = set()
        # globals: uniforms (packed 16-byte aligned, in declaration order) and samplers
        off = 0
        fields = []
        for g in self.globals_:
            if g.type.kind == "sampler":
                self.sampler_slots[g.name] = len(self.refl.samplers)
                self.uniform_vars[g.name] = Value(SAMPLER, [str(len(self.refl.samplers))])
                self.refl.samplers.append(g.name)
                continue
            if g.type.kind == "struct":
                raise CompileError(f"line {g.line}: global {g.name}: struct uniforms are not supported")
            if g.array:
                # an array uniform lives in a buffer of its own (bone palettes do not fit the 256-byte block): the block holds
                # its ADDRESS - device memory for the product (slv_buffer_device_ptr), host memory for host-compiled code
                if g.type.kind not in ("scalar", "vector", "matrix") or g.type.base == "bool":
                    raise CompileError(f"line {g.line}: global {g.name}: arrays of {g.type} are not supported")
                if g.array_len is not None and not any(x.name == g.array_len and x.type.kind == "scalar" and x.type.base in ("int", "uint")
                                                       for x in self.globals_):
                    raise CompileError(f"line {g.line}: global {g.name}: the array size {g.array_len!r} is not an integer global")
                off = (off + 15) & ~15
                fields.append(f"  alignas(16) const {C_BASE[g.type.base]}* {g.name};")
                self.uniform_arrays[g.name] = g.type
                self.refl.uniforms.append((g.name, f"{g.type}[]", off, 8))
                self.refl.arrays[g.name] = (str(g.type), 4 * g.type.n, g.array_len if g.array_len is not None else g.array)
                off += 8
                continue
            n = g.type.n
            off = (off + 15) & ~15
            cb = C_BASE["int" if g.type.base == "bool" else g.type.base]
            fields.append(f"  alignas(16) {cb} {g.name}[{n}];")
            comps = [f"U.{g.name}[{k}]" for k in range(n)]
            if g.type.base == "bool":
                comps = [f"({c} != 0)" for c in comps]
            self.uniform_vars[g.name] = Value(g.type, comps)
            self.refl.uniforms.append((g.name, str(g.type), off, 4 * n))
            off += 4 * n
        self.refl.uniform_bytes = (off + 15) & ~15
        header = ["struct SaslUniforms {"] + (fields or ["  int unused_;"]) + ["};", ""]
        order, self.recursive = self.generation_order()
        for f in order:  # prototypes of the functions on call cycles, ahead of every body
            if f.name in self.recursive:
                self.fn_table[f.name] = f
                self.lines.append(self.fn_head(f, self.fn_signature(f)[0]) + ";")
        if self.recursive:
            self.lines.append("")
        for f in order:
            self.gen_function(f)
        wrapper = [] if self.stage == "lib" else (self.gen_vs_wrapper() if self.stage == "vs" else self.gen_ps_wrapper())
        code = "\n".join(header + self.lines + wrapper) + "\n"
        return ShaderUnit(self.stage, code, self.refl, self.src)

    def entry_io(self):
        """Flattened (path, type, semantic) lists of the entry's inputs and outputs."""
        f = self.entry
        ins, outs = [], []
        for prm in f.params:
            if prm.type.kind == "struct":
                ins.append(("struct", prm, [(m.name, m.type, norm_semantic(m.semantic)) for m in self.structs[prm.type.name]]))
            elif prm.type.kind == "sampler":
                raise CompileError("the entry function cannot take a sampler")
            else:
                ins.append(("value", prm, [(prm.name, prm.type, norm_semantic(prm.semantic))]))
        if f.ret.kind == "struct":
            outs = [(m.name, m.type, norm_semantic(m.semantic)) for m in self.structs[f.ret.name]]
        elif f.ret.kind != "void":
            outs = [("ret", f.ret, norm_semantic(f.ret_semantic))]
        return ins, outs

    def gen_vs_wrapper(self):
        ins, outs = self.entry_io()
        L = ["// entry wrapper: input register k <- k-th input semantic; out[0] <- SV_Position, out[1 + k] <- k-th other output",
             "SASL_FN void slv_jit_vs(const float4* in, const unsigned char* uniforms, float4* out, const SaslSampler& S0) {",
             "  const SaslUniforms& U = *reinterpret_cast<const SaslUniforms*>(uniforms);"]
        args, reg = [], 0
        for _, prm, members in ins:
            for name, ty, sem in members:
                if sem is None:
                    raise CompileError(f"vertex-shader input {name} has no semantic")
                if ty.kind not in ("scalar", "vector") or ty.base == "bool":
                    raise CompileError(f"vertex-shader input {name}: only float / int vectors are supported")
                self.refl.inputs.append((sem[0], sem[1], str(ty)))
                # integer inputs: the register holds the element's raw bits (get_vec4 of the *_sint / *_uint formats
                # reinterprets them, stream_assembler.cpp:26-45)
                cast = {"float": "{}", "int": "sasl_asint({})", "uint": "sasl_asuint({})"}[ty.base]
                args += [cast.format(f"in[{reg}].{'xyzw'[k]}") for k in range(ty.n)]
                reg += 1
        if reg > 8:
            raise CompileError("more than 8 vertex-shader inputs")
        rets, stores, attr, have_pos, full = [], [], 0, False, {}
        for k, (name, ty, sem) in enumerate(outs):
            if sem is None:
                raise CompileError(f"vertex-shader output {name} has no semantic")
            if ty.kind not in ("scalar", "vector") or ty.base != "float":
                raise CompileError(f"vertex-shader output {name}: only float vectors are supported")
            names = [f"o{k}_{c}" for c in range(ty.n)]
            L.append(f"  float {', '.join(n + ' = 0' for n in names)};")
            rets += names
            full[k] = names + ["0.0f"] * (4 - ty.n)
        # attribute registers follow the reference's semantic array, not the declaration order (reference_semantic_order)
        for k in reference_semantic_order([sem for _, _, sem in outs]):
            name, ty, sem = outs[k]
            if sem[0] in ("SV_POSITION", "POSITION") and not have_pos:
                have_pos = True  # a position narrower than float4 (the reference's semantic test units) is padded with zeros
                stores.append(f"  out[0] = make_float4({', '.join(full[k])});")
            else:
                attr += 1
                self.refl.outputs.append((sem[0], sem[1], str(ty)))
                stores.append(f"  out[{attr}] = make_float4({', '.join(full[k])});")
        if not have_pos:
            raise CompileError("the vertex shader does not write SV_Position")
        if attr > 5:
            raise CompileError("more than 5 vertex-shader output attributes (vs_output_ops, shader.cpp:45-52)")
        self.refl.n_vs_output_attrs = attr
        L.append(f"  sasl_fn_{self.entry.name}({', '.join(['U', 'S0'] + args + rets)});")
        L += stores
        L.append("}")
        L.append(f"#define SLV_JIT_VS_OUTPUT_ATTRS {attr}")
        L.append(f"#define SLV_JIT_VS_SAMPLERS {len(self.refl.samplers)}")
        return L

    def gen_ps_wrapper(self):
        ins, outs = self.entry_io()
        L = ["// entry wrapper: k-th input <- interpolated attribute k; colour target 0 <- COLOR / SV_Target",
             "template <class Ctx>",
             "SASL_FN bool slv_jit_ps(const slv::RasterParams& p, const Ctx& px, float4& color) {",
             "  const SaslUniforms& U = *reinterpret_cast<const SaslUniforms*>(p.ps_uniforms);"]
        args, flat = [], [m for _, _, members in ins for m in members]
        for name, ty, sem in flat:
            if ty.kind not in ("scalar", "vector") or ty.base != "float":
                raise CompileError(f"pixel-shader input {name}: only float vectors are supported")
            if sem is not None and sem[0] in ("SV_POSITION", "POSITION"):
                raise CompileError("reading SV_Position in a pixel shader is not supported")
        if len(flat) > 5:
            raise CompileError("more than 5 pixel-shader inputs")
        # input -> attribute: position k of the reference's semantic array (with a C++ vertex shader bound the reference hands
        # attribute k to the k-th entry, shader_unit.cpp:129; reference_semantic_order); inputs without a semantic - which the
        # reference rejects - keep the declaration order
        sems = [sem for _, _, sem in flat]
        order = reference_semantic_order(sems) if all(s is not None for s in sems) else list(range(len(flat)))
        attr_of = {k: a for a, k in enumerate(order)}
        for a, k in enumerate(order):
            name, ty, sem = flat[k]
            L.append(f"  const float4 a{a} = px.attr({a});")
            self.refl.inputs.append(((sem or ('TEXCOORD', a))[0], (sem or ('TEXCOORD', a))[1], str(ty)))
        for k, (name, ty, sem) in enumerate(flat):
            args += [f"a{attr_of[k]}.{'xyzw'[c]}" for c in range(ty.n)]
        rets, color = [], None
        for k, (name, ty, sem) in enumerate(outs):
            if ty.kind not in ("scalar", "vector") or ty.base != "float":
                raise CompileError(f"pixel-shader output {name}: only float vectors are supported")
            names = [f"o{k}_{c}" for c in range(ty.n)]
            L.append(f"  float {', '.join(n + ' = 0' for n in names)};")
            rets += names
            self.refl.outputs.append(((sem or ('COLOR', k))[0], (sem or ('COLOR', k))[1], str(ty)))
            if sem is not None and sem[0] == "DEPTH":
                raise CompileError("pixel-shader depth output is not supported (framebuffer.cpp:348-353 ignores it for cpp shaders too)")
            if color is None and (sem is None or sem[0] in ("COLOR", "SV_TARGET")) and (sem is None or sem[1] == 0):
                color = names + ["0.0f"] * (4 - ty.n)
        L.append(f"  sasl_fn_{self.entry.name}({', '.join(['U', 'p', 'px'] + args + rets)});")
        L.append(f"  color = make_float4({', '.join(color or ['0.0f'] * 4)});")
        L.append("  return true;")
        L.append("}")
        L.append(f"#define SLV_JIT_PS_SAMPLERS {len(self.refl.samplers)}")
        return L


def _Lit(v: Value) -> Node:
    return Node("lit", (v,), 0)


def compile_shader(source: str, stage: str, entry: str | None = None, *, defines: dict | None = None, include_dirs=(),
                   sys_include_dirs=(), virtual_files: dict | None = None, file_name: str | None = None) -> ShaderUnit:
    """stage: 'vs' or 'ps' (the reference's compile(code, profile): salvia/include/salvia/core/renderer.h:136-147).
    Sources with preprocessor directives go through sasl/preprocess.py first (the reference runs Boost.Wave): `defines`
    (name -> value or None), `include_dirs` / `sys_include_dirs` and `virtual_files` (name -> text, the reference's
    add_virtual_file) feed it; `file_name` locates `#include "..."` relative to the source."""
    if stage not in ("vs", "ps", "lib"):
        raise ValueError("stage must be 'vs', 'ps' or 'lib' (functions only, no entry point: the reference's *.ss test units)")
    if "#" in source or defines:
        from .preprocess import PreprocessError, preprocess
        try:
            source = preprocess(source, defines, include_dirs, sys_include_dirs, virtual_files, file_name)
        except PreprocessError as e:
            raise CompileError(f"preprocessor: {e}") from None
    g = Gen(source, stage, entry)
    unit = g.run()
    if len(unit.reflection.samplers) > (2 if stage == "ps" else 1):  # RasterParams.sampler0 / sampler1, GeomParams.sampler0
        raise CompileError("at most two samplers per pixel shader and one per vertex shader are supported")
    return unit

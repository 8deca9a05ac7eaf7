"""SASL preprocessor (the reference runs Boost.Wave in front of its parser: sasl/src/drivers/compiler_impl.cpp, the units
sasl/test/repo/{preprocessors,include_main,include_header,include_search_path}.ss).

A line-oriented C preprocessor for what shaders use: `#define NAME [tokens]` (object-like and function-like macros), `#undef`,
`#if` / `#ifdef` / `#ifndef` / `#elif` / `#else` / `#endif` with `defined(X)`, integer arithmetic, comparison and logical
operators, `#include "file"` (directory of the including file, then the user paths) and `#include <file>` (system paths; both
forms look into the virtual files first - the reference's `add_virtual_file`), `#error`, `#pragma` / `#line` (ignored).
Skipped and directive lines become EMPTY lines, so the line numbers the front end reports stay those of the top-level file.
"""
from __future__ import annotations

import os
import re

IDENT = re.compile(r"[A-Za-z_]\w*")
TOKEN = re.compile(r"""\s+|//[^\n]*|/\*.*?\*/|"(?:\\.|[^"\\])*"|[A-Za-z_]\w*|\d[\w.]*|.""", re.S)


class PreprocessError(Exception):
    pass


class _Macro:
    def __init__(self, params, body):
        self.params, self.body = params, body  # params: None for object-like macros


def _strip_comments(src: str) -> str:
    """Block comments may span lines: replace them by the same number of newlines (line comments are left to the lexer)."""
    def repl(m):
        t = m.group()
        if t.startswith("/*"):
            return "\n" * t.count("\n") or " "
        return t
    return re.sub(r'/\*.*?\*/|"(?:\\.|[^"\\])*"|//[^\n]*', repl, src, flags=re.S)


class Preprocessor:
    def __init__(self, defines=None, include_dirs=(), sys_include_dirs=(), virtual_files=None, max_depth=32):
        self.macros: dict[str, _Macro] = {}
        for k, v in (defines or {}).items():
            self.macros[k] = _Macro(None, "" if v is None else str(v))
        self.include_dirs, self.sys_include_dirs = list(include_dirs), list(sys_include_dirs)
        self.virtual_files = dict(virtual_files or {})
        self.max_depth = max_depth
        self.included: list[str] = []

    # ---- macro expansion -----------------------------------------------------------------------------------------------------
    def expand(self, text: str, hide=frozenset()) -> str:
        out, pos = [], 0
        toks = TOKEN.findall(text)
        i = 0
        while i < len(toks):
            t = toks[i]
            m = self.macros.get(t) if IDENT.fullmatch(t) and t not in hide else None
            if m is None:
                out.append(t)
                i += 1
                continue
            if m.params is None:
                out.append(self.expand(m.body, hide | {t}))
                i += 1
                continue
            # function-like: needs '(' (possibly after white space)
            j = i + 1
            while j < len(toks) and toks[j].isspace():
                j += 1
            if j >= len(toks) or toks[j] != "(":
                out.append(t)
                i += 1
                continue
            depth, args, cur = 1, [], []
            j += 1
            while j < len(toks) and depth:
                c = toks[j]
                if c == "(":
                    depth += 1
                elif c == ")":
                    depth -= 1
                    if depth == 0:
                        break
                if c == "," and depth == 1:
                    args.append("".join(cur).strip())
                    cur = []
                else:
                    cur.append(c)
                j += 1
            if depth:
                raise PreprocessError(f"unterminated argument list of macro {t}")
            if cur or args:
                args.append("".join(cur).strip())
            if len(args) != len(m.params):
                raise PreprocessError(f"macro {t} takes {len(m.params)} argument(s), {len(args)} given")
            args = [self.expand(a, hide) for a in args]
            body = "".join(args[m.params.index(b)] if b in m.params else b for b in TOKEN.findall(m.body))
            out.append(self.expand(body, hide | {t}))
            i = j + 1
        del pos
        return "".join(out)

    # ---- #if expressions -----------------------------------------------------------------------------------------------------
    def evaluate(self, expr: str, where: str) -> bool:
        expr = re.sub(r"defined\s*\(\s*([A-Za-z_]\w*)\s*\)|defined\s+([A-Za-z_]\w*)",
                      lambda m: "1" if (m.group(1) or m.group(2)) in self.macros else "0", expr)
        expr = self.expand(expr)
        expr = re.sub(r"(?<![\w.])[A-Za-z_]\w*", "0", expr)  # remaining identifiers are 0, as in C (not the x of 0x10 / u of 10u)
        expr = re.sub(r"(\d+)[uUlL]+", r"\1", expr)
        expr = re.sub(r"(?<![\w.])0([0-7]+)\b", r"0o\1", expr)   # C octal literals
        expr = expr.replace("&&", " and ").replace("||", " or ")
        expr = re.sub(r"!(?!=)", " not ", expr)
        if not re.fullmatch(r"[\d\s()+\-*/%<>=!&|^~xXa-fA-Fandortn]*", expr):
            raise PreprocessError(f"{where}: cannot evaluate #if expression")
        try:
            return bool(eval(expr.replace("/", "//"), {"__builtins__": {}}, {}))  # noqa: S307 - digits and operators only (checked above)
        except Exception as e:  # noqa: BLE001
            raise PreprocessError(f"{where}: cannot evaluate #if expression ({e})") from None

    # ---- files -----------------------------------------------------------------------------------------------------------------
    def _find(self, name: str, system: bool, cur_dir: str | None):
        if name in self.virtual_files:
            return name, self.virtual_files[name]
        dirs = ([] if system or cur_dir is None else [cur_dir]) + ([] if system else self.include_dirs) + self.sys_include_dirs + \
               (self.include_dirs if system else [])
        for d in dirs:
            p = os.path.join(d, name)
            if os.path.isfile(p):
                return p, open(p, errors="replace").read()
        return None, None

    def process(self, src: str, file_name: str | None = None, depth: int = 0, top: bool = True) -> str:
        if depth > self.max_depth:
            raise PreprocessError("#include nested too deeply")
        cur_dir = os.path.dirname(os.path.abspath(file_name)) if file_name and os.path.isfile(file_name) else None
        where0 = file_name or "<source>"
        out = []
        stack = []  # (currently active, some branch already taken, parent active)
        active = True
        lines = _strip_comments(src).replace("\r\n", "\n").replace("\r", "\n").split("\n")
        i = 0
        while i < len(lines):
            line = lines[i]
            n_joined = 0
            while line.endswith("\\") and i + 1 < len(lines):  # line continuation
                i += 1
                n_joined += 1
                line = line[:-1] + lines[i]
            where = f"{where0}:{i + 1 - n_joined}"
            m = re.match(r"\s*#\s*(\w*)\s*(.*)", line)
            if not m:
                out.append(self.expand(line) if active else "")
                out.extend([""] * n_joined)
                i += 1
                continue
            cmd, rest = m.group(1), m.group(2).strip()
            emitted = ""
            if cmd in ("ifdef", "ifndef", "if"):
                cond = False
                if active:
                    if cmd == "if":
                        cond = self.evaluate(rest, where)
                    else:
                        name = IDENT.match(rest)
                        if not name:
                            raise PreprocessError(f"{where}: #{cmd} needs a name")
                        cond = (name.group() in self.macros) == (cmd == "ifdef")
                stack.append((active, cond, active))
                active = active and cond
            elif cmd in ("elif", "else"):
                if not stack:
                    raise PreprocessError(f"{where}: #{cmd} without #if")
                _, taken, parent = stack[-1]
                cond = parent and not taken and (True if cmd == "else" else self.evaluate(rest, where))
                stack[-1] = (active, taken or cond, parent)
                active = cond
            elif cmd == "endif":
                if not stack:
                    raise PreprocessError(f"{where}: #endif without #if")
                _, _, parent = stack.pop()
                active = parent
            elif not active:
                pass
            elif cmd == "define":
                dm = re.match(r"([A-Za-z_]\w*)(\(([^)]*)\))?\s*(.*)", rest)
                if not dm:
                    raise PreprocessError(f"{where}: malformed #define")
                params = [p.strip() for p in dm.group(3).split(",") if p.strip()] if dm.group(2) else None
                self.macros[dm.group(1)] = _Macro(params, dm.group(4).strip())
            elif cmd == "undef":
                self.macros.pop(rest.split()[0] if rest else "", None)
            elif cmd == "include":
                im = re.match(r'"([^"]+)"|<([^>]+)>', self.expand(rest) if not rest.startswith(('"', "<")) else rest)
                if not im:
                    raise PreprocessError(f"{where}: malformed #include")
                name, system = im.group(1) or im.group(2), im.group(2) is not None
                path, text = self._find(name, system, cur_dir)
                if text is None:
                    raise PreprocessError(f"{where}: cannot find include file {name!r}")
                self.included.append(path)
                # an included file contributes its text on ONE output line, so the line numbers of the including file survive
                emitted = " ".join(s for s in self.process(text, path, depth + 1, top=False).split("\n") if s.strip())
            elif cmd == "error":
                raise PreprocessError(f"{where}: #error {rest}")
            elif cmd in ("pragma", "line", ""):
                pass
            else:
                raise PreprocessError(f"{where}: unknown directive #{cmd}")
            out.append(emitted)
            out.extend([""] * n_joined)
            i += 1
        if stack:
            raise PreprocessError(f"{where0}: unterminated #if")
        del top
        return "\n".join(out)


def preprocess(source: str, defines=None, include_dirs=(), sys_include_dirs=(), virtual_files=None, file_name=None) -> str:
    return Preprocessor(defines, include_dirs, sys_include_dirs, virtual_files).process(source, file_name)

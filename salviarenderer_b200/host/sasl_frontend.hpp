// SASL front end in C++: source text -> scalarised device code + reflection, in process.
//
// The reference's shader compiler is a C++ library the renderer links (sasl/src/{parser,semantic,codegen,drivers};
// salvia::shader::compile, salvia/include/salvia/core/renderer.h:136-147).  This header is its counterpart for the B200
// renderer: lexer, recursive-descent parser, semantic analysis, reflection and one-pass code generation into three-address
// form over SCALARS (every operation in a fresh `const` temporary, so the floating-point operation order is exactly the written
// one), plus the preprocessor the reference runs in front of its parser (Boost.Wave: #define with function-like macros, #if
// / #ifdef / #elif / #else, #include with search paths and virtual files, #error).  The output is the text of the functions
// that csrc/slv_jit_unit.cu inlines into the pipeline kernels (slv_shader_compile, NVRTC) and the reflection the host needs:
// uniform block layout, sampler slots, input semantics -> registers, outputs -> attributes.
//
// It produces, byte for byte, what the Python package salviarenderer_b200/sasl/frontend.py produces (the module bench.py and
// the GPU suite drive through sasl/jit.py); tests/test_sasl_frontend_cpp.py compiles both over the shader corpus - the samples'
// shaders, the reference's own sasl/test/repo units, the known-answer shader - and compares the units.  Language scope and
// numerics are documented in DESIGN.md section 10.
#pragma once

#include <algorithm>
#include <cctype>
#include <cstdint>
#include <cstdlib>
#include <fstream>
#include <functional>
#include <map>
#include <memory>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace salvia_b200 {
namespace sasl {

struct compile_error : std::runtime_error {
  using std::runtime_error::runtime_error;
};

// ---- what a compilation returns ---------------------------------------------------------------------------------------------
struct uniform_info { std::string name, type; size_t offset = 0, size = 0; };
struct semantic_info { std::string semantic; uint32_t index = 0; std::string type; };
struct unit {
  std::string stage, entry, code;
  std::vector<uniform_info> uniforms;    // declaration order; an array uniform ("float4x4[]") holds the address of its buffer
  size_t uniform_bytes = 0;
  std::vector<std::string> samplers;     // slot order
  std::vector<semantic_info> inputs;     // VS: entry k -> input register k; PS: entry k -> attribute k
  std::vector<semantic_info> outputs;    // VS: non-position outputs, entry k -> attribute k; PS: colour targets
  int n_vs_output_attrs = 0;
  bool uses_derivatives = false;
};
struct options {
  std::vector<std::pair<std::string, std::string>> defines;  // name -> replacement text ("" for a bare -DNAME)
  std::vector<std::string> include_dirs, sys_include_dirs;
  std::map<std::string, std::string> virtual_files;           // the reference's add_virtual_file
  std::string file_name;                                      // locates `#include "..."` relative to the source
};

namespace detail {

inline std::string rstrip_set(std::string s, const char* set) {
  while (!s.empty() && std::string(set).find(s.back()) != std::string::npos) s.pop_back();
  return s;
}
inline bool all_digits(const std::string& s) { return !s.empty() && std::all_of(s.begin(), s.end(), [](unsigned char c) { return std::isdigit(c); }); }
inline std::string upper(std::string s) { for (auto& c : s) c = (char)std::toupper((unsigned char)c); return s; }
inline std::string lower(std::string s) { for (auto& c : s) c = (char)std::tolower((unsigned char)c); return s; }
inline std::string join(const std::vector<std::string>& v, const std::string& sep) {
  std::string out;
  for (size_t i = 0; i < v.size(); ++i) { if (i) out += sep; out += v[i]; }
  return out;
}
inline std::string quoted(const std::string& s) { return "'" + s + "'"; }

// ---- types ------------------------------------------------------------------------------------------------------------------
enum class Kind { Void, Scalar, Vector, Matrix, Struct, Sampler };
enum class Base { Float, Int, Uint, Bool };
struct Type {
  Kind kind = Kind::Void;
  Base base = Base::Float;
  int rows = 1, cols = 1;
  std::string name;  // struct name
  int n() const { return rows * cols; }
  bool operator==(const Type& o) const { return kind == o.kind && base == o.base && rows == o.rows && cols == o.cols && name == o.name; }
  bool operator!=(const Type& o) const { return !(*this == o); }
  bool numeric() const { return kind == Kind::Scalar || kind == Kind::Vector || kind == Kind::Matrix; }
};
inline const char* base_name(Base b) { return b == Base::Float ? "float" : b == Base::Int ? "int" : b == Base::Uint ? "uint" : "bool"; }
inline const char* c_base(Base b) { return b == Base::Float ? "float" : b == Base::Int ? "int" : b == Base::Uint ? "unsigned" : "bool"; }
inline int rank(Base b) { return b == Base::Bool ? 0 : b == Base::Int ? 1 : b == Base::Uint ? 2 : 3; }
inline std::string to_string(const Type& t) {
  switch (t.kind) {
    case Kind::Struct: return t.name;
    case Kind::Void: return "void";
    case Kind::Sampler: return "sampler";
    case Kind::Scalar: return base_name(t.base);
    case Kind::Vector: return std::string(base_name(t.base)) + std::to_string(t.cols);
    default: return std::string(base_name(t.base)) + std::to_string(t.rows) + "x" + std::to_string(t.cols);
  }
}
inline Type make_type(Kind k, Base b = Base::Float, int r = 1, int c = 1) { Type t; t.kind = k; t.base = b; t.rows = r; t.cols = c; return t; }
inline Type scalar_of(Base b) { return make_type(Kind::Scalar, b); }
inline Type vec(Base b, int n) { return n == 1 ? scalar_of(b) : make_type(Kind::Vector, b, 1, n); }
inline Type mat(Base b, int r, int c) { return make_type(Kind::Matrix, b, r, c); }
inline Type with_base(const Type& t, Base b) { return make_type(t.kind, b, t.rows, t.cols); }

// float | int | uint | bool | half | double | [u]int{8,16,32,64}_t, optionally N or NxM
inline bool parse_type_name(const std::string& s, Type& out) {
  static const char* names[] = {"float", "int", "uint", "bool", "half", "double", "int8_t", "int16_t", "int32_t", "int64_t",
                                "uint8_t", "uint16_t", "uint32_t", "uint64_t"};
  for (const char* nm : names) {
    const std::string b(nm);
    if (s.compare(0, b.size(), b) != 0) continue;
    const std::string rest = s.substr(b.size());
    auto dim = [](char c) { return c >= '1' && c <= '4'; };
    const bool ok = rest.empty() || (rest.size() == 1 && dim(rest[0])) || (rest.size() == 3 && dim(rest[0]) && rest[1] == 'x' && dim(rest[2]));
    if (!ok) continue;
    const Base base = (b == "half" || b == "double" || b == "float") ? Base::Float : b.compare(0, 4, "uint") == 0 ? Base::Uint
                      : b.compare(0, 3, "int") == 0 ? Base::Int : Base::Bool;
    if (rest.size() == 3) out = mat(base, rest[0] - '0', rest[2] - '0');
    else if (rest.size() == 1) out = vec(base, rest[0] - '0');
    else out = scalar_of(base);
    return true;
  }
  return false;
}

// ---- lexer ------------------------------------------------------------------------------------------------------------------
enum class TokKind { Num, Id, Op, Eof };
struct Tok { TokKind kind; std::string text; int line; };

inline std::vector<Tok> lex(const std::string& src) {
  static const char* ops3[] = {"<<=", ">>="};
  static const char* ops2[] = {"++", "--", "<<", ">>", "<=", ">=", "==", "!=", "&&", "||", "+=", "-=", "*=", "/=", "%=", "&=", "|=", "^="};
  std::vector<Tok> out;
  size_t pos = 0;
  int line = 1;
  const size_t n = src.size();
  auto digit = [&](size_t i) { return i < n && std::isdigit((unsigned char)src[i]); };
  auto count_nl = [&](size_t a, size_t b) { return (int)std::count(src.begin() + a, src.begin() + b, '\n'); };
  while (pos < n) {
    const unsigned char c = src[pos];
    size_t end = pos;
    if (std::isspace(c)) {
      while (end < n && std::isspace((unsigned char)src[end])) ++end;
      line += count_nl(pos, end);
      pos = end;
      continue;
    }
    if (c == '/' && pos + 1 < n && src[pos + 1] == '/') {
      while (end < n && src[end] != '\n') ++end;
      pos = end;
      continue;
    }
    if (c == '/' && pos + 1 < n && src[pos + 1] == '*') {
      const size_t close = src.find("*/", pos + 2);
      if (close != std::string::npos) {
        line += count_nl(pos, close + 2);
        pos = close + 2;
        continue;
      }
    }
    // numbers: 0x... first, then decimal / floating literals with one optional suffix character
    if (c == '0' && pos + 2 < n && (src[pos + 1] == 'x' || src[pos + 1] == 'X') && std::isxdigit((unsigned char)src[pos + 2])) {
      end = pos + 2;
      while (end < n && std::isxdigit((unsigned char)src[end])) ++end;
      if (end < n && (src[end] == 'u' || src[end] == 'U')) ++end;
      out.push_back({TokKind::Num, src.substr(pos, end - pos), line});
      pos = end;
      continue;
    }
    if (std::isdigit(c) || (c == '.' && digit(pos + 1))) {
      if (std::isdigit(c)) {
        while (digit(end)) ++end;
        if (end < n && src[end] == '.') { ++end; while (digit(end)) ++end; }
      } else {
        ++end;
        while (digit(end)) ++end;
      }
      if (end < n && (src[end] == 'e' || src[end] == 'E')) {
        size_t e = end + 1;
        if (e < n && (src[e] == '+' || src[e] == '-')) ++e;
        if (digit(e)) { while (digit(e)) ++e; end = e; }
      }
      if (end < n && std::string("fFhHuUlL").find(src[end]) != std::string::npos) ++end;
      out.push_back({TokKind::Num, src.substr(pos, end - pos), line});
      pos = end;
      continue;
    }
    if (std::isalpha(c) || c == '_') {
      while (end < n && (std::isalnum((unsigned char)src[end]) || src[end] == '_')) ++end;
      out.push_back({TokKind::Id, src.substr(pos, end - pos), line});
      pos = end;
      continue;
    }
    std::string op;
    for (const char* o : ops3) if (src.compare(pos, 3, o) == 0) op = o;
    // "++" and "--" come before the three-character operators in the reference pattern; neither is a prefix of "<<=" / ">>="
    if (op.empty()) for (const char* o : ops2) if (src.compare(pos, 2, o) == 0) { op = o; break; }
    if (op.empty() && std::string("-+*/%<>=!&|^~?:;,.(){}[]").find((char)c) != std::string::npos) op = std::string(1, (char)c);
    if (op.empty()) throw compile_error("line " + std::to_string(line) + ": unexpected character " + quoted(std::string(1, (char)c)));
    out.push_back({TokKind::Op, op, line});
    pos += op.size();
  }
  out.push_back({TokKind::Eof, "", line});
  return out;
}

// ---- AST --------------------------------------------------------------------------------------------------------------------
struct Value {
  Type type;
  std::vector<std::string> comps;  // C scalar expressions (struct: flattened member after member)
  bool lvalue = false;             // comps are assignable C lvalues
};
struct Node;
using NodePtr = std::shared_ptr<Node>;
struct VarDecl {
  Type type;
  std::string name;
  bool has_semantic = false;
  std::string semantic;
  NodePtr init;
  int line = 0;
  int array = 0;                 // > 0: literal element count; -1: sized by another global (array_len)
  std::string array_len;
};
struct SwitchGroup { std::vector<NodePtr> labels; /* nullptr = default */ std::vector<NodePtr> stmts; };
struct Node {
  // block | if | for | dowhile | switch | break | continue | return | decl | expr | comma | assign | select | bin | un | call |
  // member | index | postinc | num | bool | var | lit
  std::string op;
  std::string str;                 // operator / name / literal text
  std::vector<NodePtr> kids;       // children in source order (nullptr where a slot is empty)
  std::vector<VarDecl> decls;      // decl
  std::vector<SwitchGroup> groups; // switch
  Value lit;                       // lit: an already evaluated value
  int line = 0;
};
inline NodePtr mk(const std::string& op, int line, std::vector<NodePtr> kids = {}, const std::string& str = "") {
  auto n = std::make_shared<Node>();
  n->op = op; n->line = line; n->kids = std::move(kids); n->str = str;
  return n;
}
struct Func {
  std::string name;
  Type ret;
  bool has_ret_semantic = false;
  std::string ret_semantic;
  std::vector<VarDecl> params;
  NodePtr body;
  int line = 0;
};

struct Semantic { bool valid = false; std::string name; int index = 0; };
// 'TEXCOORD(1)' / 'texcoord1' / 'Texcoord' -> ('TEXCOORD', 1)
inline Semantic norm_semantic(bool has, const std::string& raw) {
  Semantic s;
  if (!has) return s;
  size_t a = 0, b = raw.size();
  while (a < b && std::isspace((unsigned char)raw[a])) ++a;
  while (b > a && std::isspace((unsigned char)raw[b - 1])) --b;
  const std::string t = raw.substr(a, b - a);
  size_t i = 0;
  while (i < t.size() && (std::isalpha((unsigned char)t[i]) || t[i] == '_')) ++i;
  const std::string name = t.substr(0, i), rest = t.substr(i);
  std::string digits;
  bool ok = !name.empty();
  if (rest.empty()) digits = "0";
  else if (rest.size() >= 3 && rest.front() == '(' && rest.back() == ')') digits = rest.substr(1, rest.size() - 2);
  else digits = rest;
  if (!ok || !all_digits(digits)) throw compile_error("bad semantic " + quoted(raw));
  s.valid = true; s.name = upper(name); s.index = std::atoi(digits.c_str());
  return s;
}

// salvia/include/salvia/shader/constants.h:22-37 (enum system_values) and :54-79 (the names that map to them)
inline int system_value(const std::string& name) {
  static const std::map<std::string, int> m = {{"POSITION", 1}, {"SV_POSITION", 1}, {"TEXCOORD", 2}, {"NORMAL", 3}, {"BLEND_INDICES", 4},
                                               {"BLEND_WEIGHTS", 5}, {"PSIZE", 6}, {"COLOR", 7}, {"SV_TARGET", 7}, {"DEPTH", 8}, {"SV_DEPTH", 8}};
  auto it = m.find(name);
  return it == m.end() ? 0 : it->second;
}
// Indices of `sems` (declaration order) in the order of the reference's semantic array: inserted at std::lower_bound
// (reflection_impl.cpp:70-118) under semantic_value::operator< = `sv < r.sv || name < r.name || index < r.index`
// (constants.h:94-96), which is not a strict weak order - the array depends on the insertion sequence, so the binary search of
// libstdc++'s lower_bound is replayed with that predicate.  A semantic equal to the element found is rejected, as upstream.
inline std::vector<int> reference_semantic_order(const std::vector<Semantic>& sems) {
  struct Key { int sv; std::string name; int index; };
  std::vector<Key> keys;
  for (const auto& s : sems) {
    const int sv = system_value(s.name);
    keys.push_back(sv ? Key{sv, "", s.index} : Key{9, lower(s.name), s.index});
  }
  auto less = [](const Key& a, const Key& b) { return a.sv < b.sv || a.name < b.name || a.index < b.index; };
  std::vector<int> order;
  for (int i = 0; i < (int)keys.size(); ++i) {
    size_t first = 0, n = order.size();
    while (n > 0) {
      const size_t half = n >> 1;
      if (less(keys[order[first + half]], keys[i])) { first += half + 1; n -= half + 1; }
      else n = half;
    }
    if (first < order.size()) {
      const Key& k = keys[order[first]];
      if (k.sv == keys[i].sv && k.name == keys[i].name && k.index == keys[i].index)
        throw compile_error("semantic " + sems[i].name + std::to_string(sems[i].index) + " is bound twice");
    }
    order.insert(order.begin() + first, i);
  }
  return order;
}

// ---- parser -----------------------------------------------------------------------------------------------------------------
class Parser {
public:
  std::vector<Tok> toks;
  size_t i = 0;
  std::map<std::string, std::vector<VarDecl>> structs;

  explicit Parser(const std::string& src) : toks(lex(src)) {}

  const Tok& t() const { return toks[i]; }
  const Tok& at(size_t k) const { return toks[std::min(k, toks.size() - 1)]; }
  [[noreturn]] void err(const std::string& msg) const { throw compile_error("line " + std::to_string(t().line) + ": " + msg + " (at " + quoted(t().text) + ")"); }
  bool accept(const char* text) {
    if (t().text == text && (t().kind == TokKind::Op || t().kind == TokKind::Id)) { ++i; return true; }
    return false;
  }
  void expect(const char* text) { if (!accept(text)) err(std::string("expected ") + quoted(text)); }
  std::string ident() {
    if (t().kind != TokKind::Id) err("expected an identifier");
    ++i;
    return toks[i - 1].text;
  }
  bool is_type(const Tok& tok) const {
    Type ty;
    return tok.kind == TokKind::Id && (tok.text == "void" || tok.text == "sampler" || structs.count(tok.text) || parse_type_name(tok.text, ty));
  }
  static bool is_qualifier(const std::string& s) { return s == "const" || s == "uniform" || s == "static" || s == "in" || s == "out" || s == "inout"; }
  Type type() {
    while (is_qualifier(t().text) && at(i + 1).kind == TokKind::Id && is_type(at(i + 1))) ++i;
    const std::string name = ident();
    if (name == "void") return make_type(Kind::Void);
    if (name == "sampler") return make_type(Kind::Sampler);
    if (structs.count(name)) { Type ty = make_type(Kind::Struct); ty.name = name; return ty; }
    Type ty;
    if (!parse_type_name(name, ty)) { --i; ++i; err("unknown type " + quoted(name)); }
    return ty;
  }
  bool semantic(std::string& out) {
    if (!accept(":")) return false;
    out = ident();
    if (accept("(")) {
      if (t().kind != TokKind::Num) err("expected a semantic index");
      out += "(" + t().text + ")";
      ++i;
      expect(")");
    }
    return true;
  }
  VarDecl var_decl(const Type& ty, const std::string& name, bool with_semantic, int line) {
    VarDecl d;
    d.type = ty; d.name = name; d.line = line;
    if (with_semantic) d.has_semantic = semantic(d.semantic);
    return d;
  }

  void program(std::vector<VarDecl>& globals, std::vector<Func>& funcs) {
    while (t().kind != TokKind::Eof) {
      if (accept(";")) continue;
      if (accept("struct")) {
        const std::string name = ident();
        expect("{");
        std::vector<VarDecl> members;
        while (!accept("}")) {
          const Type ty = type();
          for (;;) {
            const std::string n = ident();
            VarDecl d = var_decl(ty, n, true, 0);
            d.line = t().line;
            members.push_back(d);
            if (!accept(",")) break;
          }
          expect(";");
        }
        accept(";");
        structs[name] = members;
        continue;
      }
      const int line = t().line;
      const Type ty = type();
      std::string name = ident();
      if (accept("(")) {
        Func f;
        f.name = name; f.ret = ty; f.line = line;
        if (!accept(")")) {
          for (;;) {
            const Type pty = type();
            const std::string pname = ident();
            VarDecl d = var_decl(pty, pname, true, 0);
            d.line = t().line;
            f.params.push_back(d);
            if (!accept(",")) break;
          }
          expect(")");
        }
        f.has_ret_semantic = semantic(f.ret_semantic);
        f.body = block();
        funcs.push_back(f);
      } else {
        for (;;) {
          VarDecl d;
          d.type = ty; d.name = name; d.line = line;
          if (accept("[")) {
            if (t().kind == TokKind::Num && all_digits(t().text)) d.array = std::atoi(t().text.c_str());
            else if (t().kind == TokKind::Id) { d.array = -1; d.array_len = t().text; }
            else throw compile_error("line " + std::to_string(t().line) + ": the size of an array must be an integer literal or the name of a global");
            ++i;
            expect("]");
          }
          d.has_semantic = semantic(d.semantic);
          if (accept("=")) d.init = assign_expr();
          globals.push_back(d);
          if (!accept(",")) break;
          name = ident();
        }
        expect(";");
      }
    }
  }

  // ---- statements
  NodePtr block() {
    const int line = t().line;
    expect("{");
    std::vector<NodePtr> stmts;
    while (!accept("}")) stmts.push_back(statement());
    return mk("block", line, stmts);
  }
  NodePtr statement() {
    const int line = t().line;
    if (t().text == "{") return block();
    if (accept(";")) return mk("block", line);
    if (accept("if")) {
      expect("(");
      NodePtr c = expr();
      expect(")");
      NodePtr a = statement();
      NodePtr b = accept("else") ? statement() : nullptr;
      return mk("if", line, {c, a, b});
    }
    if (accept("for")) {
      expect("(");
      NodePtr init = t().text == ";" ? nullptr : simple_statement();
      expect(";");
      NodePtr cond = t().text == ";" ? nullptr : expr();
      expect(";");
      NodePtr step = t().text == ")" ? nullptr : expr();
      expect(")");
      NodePtr body = statement();
      return mk("for", line, {init, cond, step, body});
    }
    if (accept("while")) {
      expect("(");
      NodePtr c = expr();
      expect(")");
      NodePtr body = statement();
      return mk("for", line, {nullptr, c, nullptr, body});
    }
    if (accept("do")) {
      NodePtr body = statement();
      expect("while");
      expect("(");
      NodePtr c = expr();
      expect(")");
      expect(";");
      return mk("dowhile", line, {body, c});
    }
    if (accept("switch")) {
      expect("(");
      NodePtr sel = expr();
      expect(")");
      expect("{");
      NodePtr sw = mk("switch", line, {sel});
      while (!accept("}")) {
        SwitchGroup g;
        bool any = false;
        while (t().text == "case" || t().text == "default") {
          any = true;
          if (accept("default")) g.labels.push_back(nullptr);
          else { expect("case"); g.labels.push_back(expr()); }
          expect(":");
        }
        if (!any) throw compile_error("line " + std::to_string(t().line) + ": statement before the first case label of a switch");
        while (t().text != "case" && t().text != "default" && t().text != "}") g.stmts.push_back(statement());
        sw->groups.push_back(g);
      }
      return sw;
    }
    if (accept("break")) { expect(";"); return mk("break", line); }
    if (accept("continue")) { expect(";"); return mk("continue", line); }
    if (accept("return")) {
      NodePtr e = t().text == ";" ? nullptr : expr();
      expect(";");
      return mk("return", line, {e});
    }
    NodePtr s = simple_statement();
    expect(";");
    return s;
  }
  NodePtr simple_statement() {
    const int line = t().line;
    if (is_type(t()) && at(i + 1).kind == TokKind::Id) {
      const Type ty = type();
      NodePtr n = mk("decl", line);
      for (;;) {
        VarDecl d;
        d.type = ty; d.name = ident(); d.line = line;
        if (accept("=")) d.init = assign_expr();
        n->decls.push_back(d);
        if (!accept(",")) break;
      }
      return n;
    }
    return mk("expr", line, {expr()});
  }

  // ---- expressions
  NodePtr expr() {
    NodePtr e = assign_expr();
    while (accept(",")) { NodePtr r = assign_expr(); e = mk("comma", e->line, {e, r}); }
    return e;
  }
  static bool is_assign_op(const std::string& s) {
    static const std::set<std::string> ops = {"=", "+=", "-=", "*=", "/=", "%=", "&=", "|=", "^=", "<<=", ">>="};
    return ops.count(s) != 0;
  }
  NodePtr assign_expr() {
    NodePtr lhs = ternary();
    if (t().kind == TokKind::Op && is_assign_op(t().text)) {
      const std::string op = t().text;
      ++i;
      NodePtr rhs = assign_expr();
      return mk("assign", lhs->line, {lhs, rhs}, op);
    }
    return lhs;
  }
  NodePtr ternary() {
    NodePtr c = binary(0);
    if (accept("?")) {
      NodePtr a = assign_expr();
      expect(":");
      NodePtr b = assign_expr();
      return mk("select", c->line, {c, a, b});
    }
    return c;
  }
  NodePtr binary(int level) {
    static const std::vector<std::vector<std::string>> prec = {{"||"}, {"&&"}, {"|"}, {"^"}, {"&"}, {"==", "!="}, {"<", ">", "<=", ">="},
                                                               {"<<", ">>"}, {"+", "-"}, {"*", "/", "%"}};
    if (level == (int)prec.size()) return unary();
    NodePtr lhs = binary(level + 1);
    while (t().kind == TokKind::Op && std::find(prec[level].begin(), prec[level].end(), t().text) != prec[level].end()) {
      const std::string op = t().text;
      ++i;
      NodePtr rhs = binary(level + 1);
      lhs = mk("bin", lhs->line, {lhs, rhs}, op);
    }
    return lhs;
  }
  NodePtr unary() {
    const int line = t().line;
    if (t().kind == TokKind::Op && (t().text == "-" || t().text == "+" || t().text == "!" || t().text == "~")) {
      const std::string op = t().text;
      ++i;
      return mk("un", line, {unary()}, op);
    }
    if (t().kind == TokKind::Op && (t().text == "++" || t().text == "--")) {
      const std::string op = t().text;
      ++i;
      NodePtr target = unary();
      return mk("assign", line, {target, mk("num", line, {}, "1")}, op == "++" ? "+=" : "-=");
    }
    // C-style cast: '(' type ')' unary
    if (t().text == "(" && is_type(at(i + 1)) && at(i + 2).text == ")") {
      ++i;
      const Type ty = type();
      expect(")");
      return mk("call", line, {unary()}, to_string(ty));
    }
    return postfix();
  }
  NodePtr postfix() {
    NodePtr e = primary();
    for (;;) {
      const int line = t().line;
      if (accept(".")) e = mk("member", line, {e}, ident());
      else if (accept("[")) {
        NodePtr idx = expr();
        expect("]");
        e = mk("index", line, {e, idx});
      } else if (t().kind == TokKind::Op && (t().text == "++" || t().text == "--")) {
        const std::string op = t().text;
        ++i;
        e = mk("postinc", line, {e}, op == "++" ? "+" : "-");
      } else return e;
    }
  }
  NodePtr primary() {
    const int line = t().line;
    if (t().kind == TokKind::Num) { ++i; return mk("num", line, {}, toks[i - 1].text); }
    if (accept("(")) {
      NodePtr e = expr();
      expect(")");
      return e;
    }
    if (t().kind == TokKind::Id) {
      if (t().text == "true" || t().text == "false") { ++i; return mk("bool", line, {}, toks[i - 1].text); }
      const std::string name = ident();
      if (accept("(")) {
        std::vector<NodePtr> args;
        if (!accept(")")) {
          for (;;) {
            args.push_back(assign_expr());
            if (!accept(",")) break;
          }
          expect(")");
        }
        return mk("call", line, args, name);
      }
      return mk("var", line, {}, name);
    }
    err("expected an expression");
  }
};

}  // namespace detail
}  // namespace sasl
}  // namespace salvia_b200

#include "sasl_frontend_gen.inl"

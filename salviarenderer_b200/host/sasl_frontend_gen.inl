// Code generation of the SASL front end (included by sasl_frontend.hpp).  One pass over the AST into three-address form over
// scalars; mirrors salviarenderer_b200/sasl/frontend.py function by function (the two must emit identical text).
#pragma once

namespace salvia_b200 {
namespace sasl {
namespace detail {

inline NodePtr lit_node(const Value& v) {
  auto n = std::make_shared<Node>();
  n->op = "lit";
  n->lit = v;
  return n;
}

class Gen {
public:
  std::string stage;
  Parser p;
  std::vector<VarDecl> globals;
  std::vector<Func> funcs;
  std::vector<std::string> lines;
  int indent = 1;
  int ntemp = 0;
  std::vector<std::map<std::string, Value>> scopes;
  std::map<std::string, const Func*> fn_table;
  unit refl;
  std::map<std::string, Value> uniform_vars;
  std::map<std::string, Type> uniform_arrays;  // array uniforms: element type (the block holds the buffer's address)
  int divergent = 0;                             // > 0 while emitting code under a data-dependent branch or loop
  bool has_ret = false;
  Value cur_ret;
  const Func* entry = nullptr;
  std::vector<int> loop_ids;
  std::vector<std::pair<int, size_t>> switch_ids;  // (id, number of loops open when the switch was opened)
  std::set<std::string> recursive, fns_with_derivatives;
  bool cur_fn_derivs = false;

  Gen(const std::string& src, const std::string& stage_, const std::string& entry_name) : stage(stage_), p(src) {
    p.program(globals, funcs);
    refl.stage = stage;
    if (stage != "lib") entry = pick_entry(entry_name);
    refl.entry = entry ? entry->name : "";
  }

  const std::vector<VarDecl>& members(const std::string& struct_name) const { return p.structs.at(struct_name); }

  // ---- helpers
  [[noreturn]] void err(int line, const std::string& msg) const { throw compile_error("line " + std::to_string(line) + ": " + msg); }
  [[noreturn]] void err(const NodePtr& n, const std::string& msg) const { err(n ? n->line : 0, msg); }
  void emit(const std::string& s) { lines.push_back(std::string(2 * (size_t)indent, ' ') + s); }
  std::string temp(Base base, const std::string& expr) {
    const std::string name = "t" + std::to_string(++ntemp);
    emit(std::string("const ") + c_base(base) + " " + name + " = " + expr + ";");
    return name;
  }
  // Scalar base types of the flattened components of `ty`
  std::vector<Base> flat_types(const Type& ty) const {
    std::vector<Base> out;
    if (ty.kind == Kind::Struct) {
      for (const auto& m : members(ty.name)) { auto sub = flat_types(m.type); out.insert(out.end(), sub.begin(), sub.end()); }
      return out;
    }
    if (ty.kind == Kind::Sampler) return {Base::Int};
    return std::vector<Base>((size_t)ty.n(), ty.base);
  }
  bool struct_has_semantic(const Type& ty) const {
    if (ty.kind != Kind::Struct) return false;
    for (const auto& m : members(ty.name)) if (m.has_semantic && !m.semantic.empty()) return true;
    return false;
  }
  const Func* pick_entry(const std::string& name) const {
    if (funcs.empty()) throw compile_error("no function in the translation unit");
    if (!name.empty()) {
      for (const auto& f : funcs) if (f.name == name) return &f;
      throw compile_error("entry function " + quoted(name) + " not found");
    }
    auto has_sem = [&](const Func& f) {
      if (f.has_ret_semantic && !f.ret_semantic.empty()) return true;
      for (const auto& prm : f.params)
        if ((prm.has_semantic && !prm.semantic.empty()) || struct_has_semantic(prm.type)) return true;
      return struct_has_semantic(f.ret);
    };
    const Func* pick = nullptr;
    for (const auto& f : funcs) if (has_sem(f)) pick = &f;
    if (!pick)
      for (const auto& f : funcs) if (f.name == "main" || f.name == "vs_main" || f.name == "ps_main" || f.name == "fn") pick = &f;
    if (!pick) throw compile_error("cannot determine the entry function (no parameter or return value carries a semantic)");
    return pick;
  }

  // ---- conversions
  Value convert(const Value& v, const Type& to, const NodePtr& node) {
    if (v.type == to) return v;
    if (to.kind == Kind::Struct || v.type.kind == Kind::Struct || v.type.kind == Kind::Sampler || v.type.kind == Kind::Void)
      err(node, "cannot convert " + to_string(v.type) + " to " + to_string(to));
    std::vector<std::string> src = v.comps;
    if (v.type.n() == 1 && to.n() > 1) {
      src.assign((size_t)to.n(), v.comps[0]);
    } else if (v.type.n() > to.n()) {  // HLSL truncation
      if (v.type.kind == Kind::Matrix || to.kind == Kind::Matrix) {
        if (v.type.kind == Kind::Matrix && to.kind == Kind::Matrix && to.rows <= v.type.rows && to.cols <= v.type.cols) {
          src.clear();
          for (int r = 0; r < to.rows; ++r) for (int c = 0; c < to.cols; ++c) src.push_back(v.comps[(size_t)(r * v.type.cols + c)]);
        } else err(node, "cannot convert " + to_string(v.type) + " to " + to_string(to));
      } else src.resize((size_t)to.n());
    } else if (v.type.n() != to.n()) err(node, "cannot convert " + to_string(v.type) + " to " + to_string(to));
    if (v.type.base != to.base) {
      for (auto& c : src) {
        if (to.base == Base::Bool) c = temp(Base::Bool, "(" + c + " != 0)");
        else c = temp(to.base, std::string("(") + c_base(to.base) + ")(" + c + ")");
      }
    }
    Value out;
    out.type = to; out.comps = src;
    return out;
  }
  // Common type for a binary / ternary operation (scalar broadcast, base promotion, vector truncation)
  Type unify(Value& a, Value& b, const NodePtr& node, bool arith = true) {
    const Type ta = a.type, tb = b.type;
    for (const Type* t : {&ta, &tb}) if (!t->numeric()) err(node, "operand of type " + to_string(*t));
    Base base = rank(ta.base) >= rank(tb.base) ? ta.base : tb.base;
    if (arith && base == Base::Bool) base = Base::Int;
    Type shape;
    if (ta.n() == 1) shape = tb;
    else if (tb.n() == 1) shape = ta;
    else if (ta.kind == Kind::Matrix || tb.kind == Kind::Matrix) {
      if (ta.rows != tb.rows || ta.cols != tb.cols) err(node, "shape mismatch " + to_string(ta) + " vs " + to_string(tb));
      shape = ta;
    } else shape = ta.cols <= tb.cols ? ta : tb;
    const Type to = with_base(shape, base);
    a = convert(a, to, node);
    b = convert(b, to, node);
    return to;
  }
  // Copies lvalue components into temporaries (so a later store cannot change what was read)
  Value materialize(const Value& v) {
    if (!v.lvalue) return v;
    const auto bases = flat_types(v.type);
    Value out;
    out.type = v.type;
    for (size_t k = 0; k < bases.size() && k < v.comps.size(); ++k) out.comps.push_back(temp(bases[k], v.comps[k]));
    return out;
  }

  // ---- scopes
  Value lookup(const std::string& name, const NodePtr& node) const {
    for (auto s = scopes.rbegin(); s != scopes.rend(); ++s) { auto it = s->find(name); if (it != s->end()) return it->second; }
    auto it = uniform_vars.find(name);
    if (it != uniform_vars.end()) return it->second;
    err(node, "undeclared identifier " + quoted(name));
  }
  Value declare(const Type& ty, const std::string& name) {
    ++ntemp;
    const auto bases = flat_types(ty);
    Value v;
    v.type = ty; v.lvalue = true;
    for (size_t k = 0; k < bases.size(); ++k) v.comps.push_back("v" + std::to_string(ntemp) + "_" + name + "_" + std::to_string(k));
    for (size_t k = 0; k < bases.size(); ++k) emit(std::string(c_base(bases[k])) + " " + v.comps[k] + " = 0;");
    scopes.back()[name] = v;
    return v;
  }
  void store(const Value& dst, const Value& src_, const NodePtr& node) {
    const Value src = materialize(convert(src_, dst.type, node));
    for (size_t k = 0; k < dst.comps.size() && k < src.comps.size(); ++k) emit(dst.comps[k] + " = " + src.comps[k] + ";");
  }
  static Value make_value(const Type& t, std::vector<std::string> comps, bool lvalue = false) {
    Value v;
    v.type = t; v.comps = std::move(comps); v.lvalue = lvalue;
    return v;
  }

  // ---- expressions
  Value expr(const NodePtr& n) {
    const std::string& op = n->op;
    if (op == "num") return e_num(n);
    if (op == "bool") return make_value(scalar_of(Base::Bool), {n->str});
    if (op == "var") return lookup(n->str, n);
    if (op == "comma") { expr(n->kids[0]); return expr(n->kids[1]); }
    if (op == "member") return e_member(n);
    if (op == "index") return e_index(n);
    if (op == "un") return e_un(n);
    if (op == "bin") return e_bin(n->str, n->kids[0], n->kids[1], n);
    if (op == "select") return e_select(n);
    if (op == "assign") return e_assign(n);
    if (op == "postinc") return e_postinc(n);
    if (op == "lit") return n->lit;
    if (op == "call") return e_call(n);
    err(n, "unexpected " + op);
  }
  Value e_num(const NodePtr& n) {
    const std::string& t = n->str;
    const bool uns = !t.empty() && (t.back() == 'u' || t.back() == 'U');
    if (lower(t).compare(0, 2, "0x") == 0) {
      const std::string hex = rstrip_set(t, "uU");
      const unsigned long long v = std::strtoull(hex.c_str(), nullptr, 16);
      if (hex.size() > 2 + 16 || v > 0xFFFFFFFFull) err(n, "integer literal " + t + " does not fit 32 bits");
      return make_value(scalar_of(uns ? Base::Uint : Base::Int), {std::to_string(v) + (uns ? "u" : "")});
    }
    if (t.size() > 1 && uns && all_digits(t.substr(0, t.size() - 1))) return make_value(scalar_of(Base::Uint), {t.substr(0, t.size() - 1) + "u"});
    if (all_digits(t) || (t.size() > 1 && (t.back() == 'l' || t.back() == 'L') && all_digits(t.substr(0, t.size() - 1))))
      return make_value(scalar_of(Base::Int), {rstrip_set(t, "lL")});
    std::string body = rstrip_set(t, "fFhHlL");
    if (body.find('.') == std::string::npos && lower(body).find('e') == std::string::npos) body += ".0";
    return make_value(scalar_of(Base::Float), {body + "f"});
  }
  static int swizzle_index(char c) {
    switch (c) { case 'x': case 'r': return 0; case 'y': case 'g': return 1; case 'z': case 'b': return 2; case 'w': case 'a': return 3; default: return -1; }
  }
  Value e_member(const NodePtr& n) {
    const Value base = expr(n->kids[0]);
    const std::string& name = n->str;
    const Type& ty = base.type;
    if (ty.kind == Kind::Struct) {
      size_t off = 0;
      for (const auto& m : members(ty.name)) {
        const size_t k = flat_types(m.type).size();
        if (m.name == name) return make_value(m.type, std::vector<std::string>(base.comps.begin() + off, base.comps.begin() + off + k), base.lvalue);
        off += k;
      }
      err(n, to_string(ty) + " has no member " + quoted(name));
    }
    const bool swz = !name.empty() && name.size() <= 4 && std::all_of(name.begin(), name.end(), [](char c) { return swizzle_index(c) >= 0; });
    if ((ty.kind == Kind::Scalar || ty.kind == Kind::Vector) && swz) {
      std::vector<int> idx;
      for (char c : name) idx.push_back(swizzle_index(c));
      if (*std::max_element(idx.begin(), idx.end()) >= ty.n()) err(n, "swizzle ." + name + " out of range for " + to_string(ty));
      std::vector<std::string> comps;
      for (int k : idx) comps.push_back(base.comps[(size_t)k]);
      const bool distinct = std::set<int>(idx.begin(), idx.end()).size() == idx.size();
      return make_value(vec(ty.base, (int)idx.size()), comps, base.lvalue && distinct);
    }
    if (ty.kind == Kind::Matrix) {  // _mRC (zero-based) | _RC (one-based)
      int r = -1, c = -1;
      if (name.size() == 4 && name[0] == '_' && name[1] == 'm' && name[2] >= '0' && name[2] <= '3' && name[3] >= '0' && name[3] <= '3') { r = name[2] - '0'; c = name[3] - '0'; }
      else if (name.size() == 3 && name[0] == '_' && name[1] >= '1' && name[1] <= '4' && name[2] >= '1' && name[2] <= '4') { r = name[1] - '1'; c = name[2] - '1'; }
      if (r >= 0 && r < ty.rows && c < ty.cols) return make_value(scalar_of(ty.base), {base.comps.at((size_t)(r * ty.cols + c))}, base.lvalue);
    }
    err(n, "cannot take ." + name + " of " + to_string(ty));
  }
  bool in_scopes(const std::string& name) const {
    for (const auto& s : scopes) if (s.count(name)) return true;
    return false;
  }
  Value e_index(const NodePtr& n) {
    const NodePtr& target = n->kids[0];
    const NodePtr& idx = n->kids[1];
    // an element of an array uniform: loads through the address the uniform block holds
    if (target->op == "var" && uniform_arrays.count(target->str) && !in_scopes(target->str)) {
      const std::string& name = target->str;
      const Type ety = uniform_arrays.at(name);
      const Value iv = convert(expr(idx), scalar_of(Base::Int), n);
      const std::string i = temp(Base::Int, iv.comps[0]);
      std::vector<std::string> comps;
      for (int k = 0; k < ety.n(); ++k) comps.push_back(temp(ety.base, "U." + name + "[" + i + " * " + std::to_string(ety.n()) + " + " + std::to_string(k) + "]"));
      return make_value(ety, comps);
    }
    const Value base = expr(target);
    const Type& ty = base.type;
    if (idx->op != "num") {
      // a run-time index into a vector / the rows of a matrix: a chain of selects over the components (read-only)
      const Value iv = expr(idx);
      if (iv.type.kind != Kind::Scalar || (iv.type.base != Base::Int && iv.type.base != Base::Uint)) err(n, "an index must be an integer scalar");
      if (ty.kind != Kind::Vector && ty.kind != Kind::Matrix) err(n, "cannot index " + to_string(ty));
      const std::string i = temp(Base::Int, convert(iv, scalar_of(Base::Int), n).comps[0]);
      const Value src = materialize(base);
      const int count = ty.kind == Kind::Vector ? ty.cols : ty.rows, width = ty.kind == Kind::Vector ? 1 : ty.cols;
      std::vector<std::string> comps;
      for (int c = 0; c < width; ++c) {
        std::string e = src.comps[(size_t)((count - 1) * width + c)];
        for (int r = count - 2; r >= 0; --r) e = "(" + i + " == " + std::to_string(r) + " ? " + src.comps[(size_t)(r * width + c)] + " : " + e + ")";
        comps.push_back(temp(ty.base, e));
      }
      return make_value(vec(ty.base, width), comps);
    }
    const std::string digits = rstrip_set(idx->str, "uUlL");
    const bool hex = digits.size() > 2 && lower(digits).compare(0, 2, "0x") == 0 &&
                     std::all_of(digits.begin() + 2, digits.end(), [](unsigned char c) { return std::isxdigit(c); });
    if (!hex && !all_digits(digits)) err(n, "an index must be an integer literal");
    const long i = std::strtol(digits.c_str(), nullptr, hex ? 16 : 10);
    if (ty.kind == Kind::Vector) {
      if (i >= ty.n()) err(n, "index out of range");
      return make_value(scalar_of(ty.base), {base.comps[(size_t)i]}, base.lvalue);
    }
    if (ty.kind == Kind::Matrix) {
      if (i >= ty.rows) err(n, "index out of range");
      return make_value(vec(ty.base, ty.cols), std::vector<std::string>(base.comps.begin() + i * ty.cols, base.comps.begin() + (i + 1) * ty.cols), base.lvalue);
    }
    err(n, "cannot index " + to_string(ty));
  }
  Value e_un(const NodePtr& n) {
    const std::string& op = n->str;
    Value v = expr(n->kids[0]);
    if (!v.type.numeric()) err(n, "unary " + op + " on " + to_string(v.type));
    if (op == "+") return v;
    std::vector<std::string> out;
    if (op == "!") {
      v = convert(v, with_base(v.type, Base::Bool), n);
      for (const auto& c : v.comps) out.push_back(temp(Base::Bool, "!" + c));
      return make_value(v.type, out);
    }
    if (op == "~") {
      for (const auto& c : v.comps) out.push_back(temp(v.type.base, "~" + c));
      return make_value(v.type, out);
    }
    const Base base = v.type.base == Base::Bool ? Base::Int : v.type.base;
    v = convert(v, with_base(v.type, base), n);
    for (const auto& c : v.comps) out.push_back(temp(base, "-" + c));
    return make_value(v.type, out);
  }
  Value bin_values(const std::string& op, const Value& a, const Value& b, const NodePtr& at) { return e_bin(op, lit_node(a), lit_node(b), at); }
  Value e_bin(const std::string& op, const NodePtr& an, const NodePtr& bn, const NodePtr& n) {
    std::vector<std::string> out;
    Value a = expr(an);
    Value b = expr(bn);
    if (op == "&&" || op == "||") {  // no side effects in operands of the supported subset: evaluate both
      const bool av = a.type.kind == Kind::Vector || a.type.kind == Kind::Matrix, bv = b.type.kind == Kind::Vector || b.type.kind == Kind::Matrix;
      if (av || bv) {  // component-wise on vectors / matrices
        const Type bt = with_base(av ? a.type : b.type, Base::Bool);
        a = convert(a, bt, n);
        b = convert(b, bt, n);
        for (size_t k = 0; k < a.comps.size() && k < b.comps.size(); ++k) out.push_back(temp(Base::Bool, a.comps[k] + " " + op + " " + b.comps[k]));
        return make_value(bt, out);
      }
      a = convert(a, scalar_of(Base::Bool), n);
      b = convert(b, scalar_of(Base::Bool), n);
      return make_value(scalar_of(Base::Bool), {temp(Base::Bool, a.comps[0] + " " + op + " " + b.comps[0])});
    }
    if (op == "==" || op == "!=" || op == "<" || op == ">" || op == "<=" || op == ">=") {
      const Type to = unify(a, b, n, false);
      for (size_t k = 0; k < a.comps.size() && k < b.comps.size(); ++k) out.push_back(temp(Base::Bool, a.comps[k] + " " + op + " " + b.comps[k]));
      return make_value(with_base(to, Base::Bool), out);
    }
    const Type to = unify(a, b, n);
    const bool int_only = op == "%" || op == "&" || op == "|" || op == "^" || op == "<<" || op == ">>";
    if (int_only && to.base == Base::Float) {
      if (op != "%") err(n, "operator " + op + " on floating-point operands");
      for (size_t k = 0; k < a.comps.size() && k < b.comps.size(); ++k) out.push_back(temp(Base::Float, "fmodf(" + a.comps[k] + ", " + b.comps[k] + ")"));
      return make_value(to, out);
    }
    for (size_t k = 0; k < a.comps.size() && k < b.comps.size(); ++k) out.push_back(temp(to.base, a.comps[k] + " " + op + " " + b.comps[k]));
    return make_value(to, out);
  }
  Value e_select(const NodePtr& n) {
    Value c = expr(n->kids[0]);
    Value a = expr(n->kids[1]);
    Value b = expr(n->kids[2]);
    if (a.type.kind == Kind::Struct) err(n, "?: on structs");
    const Type to = unify(a, b, n, false);
    if (!c.type.numeric() || (c.type.n() != 1 && c.type.n() != to.n())) err(n, "?: condition of type " + to_string(c.type) + " with operands of type " + to_string(to));
    const bool one = c.type.n() == 1;
    c = convert(c, make_type(one ? Kind::Scalar : to.kind, Base::Bool, one ? 1 : to.rows, one ? 1 : c.type.cols), n);
    std::vector<std::string> out;
    for (size_t k = 0; k < a.comps.size() && k < b.comps.size(); ++k) {
      if (!one && k >= c.comps.size()) break;
      out.push_back(temp(to.base, (one ? c.comps[0] : c.comps[k]) + " ? " + a.comps[k] + " : " + b.comps[k]));
    }
    return make_value(to, out);
  }
  Value e_assign(const NodePtr& n) {
    Value rhs = expr(n->kids[1]);
    const Value lhs = expr(n->kids[0]);
    if (!lhs.lvalue) err(n, "left side of an assignment is not assignable");
    if (n->str != "=") {
      const Value cur = materialize(lhs);
      rhs = bin_values(n->str.substr(0, n->str.size() - 1), cur, rhs, n);
    }
    store(lhs, rhs, n);
    return lhs;
  }
  Value e_postinc(const NodePtr& n) {
    const Value v = expr(n->kids[0]);
    if (!v.lvalue) err(n, "operand of ++/-- is not assignable");
    const Value old = materialize(v);
    store(v, bin_values(n->str, old, make_value(scalar_of(Base::Int), {"1"}), n), n);
    return old;
  }

  // ---- calls: constructors, intrinsics, user functions
  using Args = std::vector<Value>;
  Value e_call(const NodePtr& n) {
    const std::string& name = n->str;
    Type ty;
    Args args;
    if (parse_type_name(name, ty)) {
      for (const auto& a : n->kids) args.push_back(expr(a));
      return construct(ty, args, n);
    }
    auto f = fn_table.find(name);
    if (f != fn_table.end()) {
      for (const auto& a : n->kids) args.push_back(expr(a));
      return call_user(*f->second, args, n);
    }
    for (const auto& a : n->kids) args.push_back(expr(a));
    const char* unary = unary_math(name);
    if (unary) return map1(args, n, [&](const std::string& c) { return std::string(unary) + "(" + c + ")"; });
    return intrinsic(name, args, n);
  }
  static const char* unary_math(const std::string& name) {
    static const std::map<std::string, const char*> m = {
        {"sqrt", "sqrtf"}, {"exp", "sasl_m_exp"}, {"exp2", "sasl_m_exp2"}, {"log", "sasl_m_log"}, {"log2", "sasl_m_log2"}, {"log10", "sasl_m_log10"},
        {"sin", "sasl_m_sin"}, {"cos", "sasl_m_cos"}, {"tan", "sasl_m_tan"}, {"asin", "sasl_m_asin"}, {"acos", "sasl_m_acos"}, {"atan", "sasl_m_atan"},
        {"sinh", "sasl_m_sinh"}, {"cosh", "sasl_m_cosh"}, {"tanh", "sasl_m_tanh"}, {"floor", "sasl_m_floor"}, {"ceil", "sasl_m_ceil"},
        {"trunc", "sasl_m_trunc"}, {"round", "sasl_m_round"}};
    auto it = m.find(name);
    return it == m.end() ? nullptr : it->second;
  }
  Value construct(const Type& ty, const Args& args, const NodePtr& n) {
    if (args.size() == 1 && args[0].type.numeric() && (args[0].type.n() == 1 || args[0].type.n() >= ty.n())) return convert(args[0], ty, n);
    std::vector<std::string> comps;
    for (const auto& a0 : args) {
      if (!a0.type.numeric()) err(n, "constructor argument of type " + to_string(a0.type));
      const Value a = convert(a0, with_base(a0.type, ty.base), n);
      comps.insert(comps.end(), a.comps.begin(), a.comps.end());
    }
    if ((int)comps.size() != ty.n()) err(n, to_string(ty) + " constructed from " + std::to_string(comps.size()) + " components");
    return make_value(ty, comps);
  }
  std::vector<std::string> ctx_args() const {
    return stage == "ps" ? std::vector<std::string>{"U", "p", "px"} : std::vector<std::string>{"U", "S0"};
  }
  Value call_user(const Func& f, const Args& args, const NodePtr& n) {
    if (args.size() != f.params.size()) err(n, f.name + " expects " + std::to_string(f.params.size()) + " arguments");
    std::vector<std::string> actual;
    for (size_t k = 0; k < args.size(); ++k) {
      const VarDecl& prm = f.params[k];
      if (prm.type.kind == Kind::Sampler) {
        if (args[k].type.kind != Kind::Sampler) err(n, "sampler argument expected");
        actual.insert(actual.end(), args[k].comps.begin(), args[k].comps.end());
      } else {
        const Value a = materialize(prm.type.kind != Kind::Struct ? convert(args[k], prm.type, n) : args[k]);
        actual.insert(actual.end(), a.comps.begin(), a.comps.end());
      }
    }
    std::vector<std::string> rets;
    if (f.ret.kind != Kind::Void) {
      for (Base b : flat_types(f.ret)) {
        rets.push_back("r" + std::to_string(++ntemp));
        emit(std::string(c_base(b)) + " " + rets.back() + " = 0;");
      }
    }
    if (stage == "ps" && divergent && fns_with_derivatives.count(f.name))
      err(n, f.name + " takes screen-space derivatives and is called under divergent control flow");
    std::vector<std::string> all = ctx_args();
    all.insert(all.end(), actual.begin(), actual.end());
    all.insert(all.end(), rets.begin(), rets.end());
    emit("sasl_fn_" + f.name + "(" + join(all, ", ") + ");");
    return make_value(f.ret, rets);
  }
  Value to_base(const Value& v, Base base, const NodePtr& n) {
    if (!v.type.numeric()) err(n, "argument of type " + to_string(v.type));
    return convert(v, with_base(v.type, base), n);
  }
  using Fmt1 = std::function<std::string(const std::string&)>;
  using FmtN = std::function<std::string(const std::vector<std::string>&)>;
  Value map1(const Args& args, const NodePtr& n, const Fmt1& fmt, Base base = Base::Float) {
    if (args.size() != 1) err(n, "expects 1 argument");
    const Value v = to_base(args[0], base, n);
    std::vector<std::string> out;
    for (const auto& c : v.comps) out.push_back(temp(v.type.base, fmt(c)));
    return make_value(v.type, out);
  }
  Value mapn(const Args& args, const NodePtr& n, const FmtN& fmt, size_t count) {
    if (args.size() != count) err(n, "expects " + std::to_string(count) + " arguments");
    std::vector<Value> vs;
    for (const auto& a : args) vs.push_back(to_base(a, Base::Float, n));
    Type shape = vs[0].type;
    for (const auto& v : vs) if (v.type.n() > shape.n()) shape = v.type;  // first of the widest
    bool ragged = false;
    for (const auto& v : vs) if (v.type.n() != 1 && v.type.n() != shape.n()) ragged = true;
    if (ragged) {
      bool have = false;
      for (const auto& v : vs) if (v.type.n() > 1 && (!have || v.type.n() < shape.n())) { shape = v.type; have = true; }  // first of the narrowest
    }
    for (auto& v : vs) v = convert(v, shape, n);
    std::vector<std::string> out;
    size_t len = vs[0].comps.size();
    for (const auto& v : vs) len = std::min(len, v.comps.size());
    for (size_t k = 0; k < len; ++k) {
      std::vector<std::string> cs;
      for (const auto& v : vs) cs.push_back(v.comps[k]);
      out.push_back(temp(Base::Float, fmt(cs)));
    }
    return make_value(shape, out);
  }
  std::string sum_lr(const std::vector<std::string>& terms) {
    std::string acc = terms.at(0);
    for (size_t k = 1; k < terms.size(); ++k) acc = temp(Base::Float, acc + " + " + terms[k]);
    return acc;
  }
  std::string dot_comps(const std::vector<std::string>& a, const std::vector<std::string>& b) {
    std::vector<std::string> prods;
    for (size_t k = 0; k < a.size() && k < b.size(); ++k) prods.push_back(temp(Base::Float, a[k] + " * " + b[k]));
    return sum_lr(prods);
  }
  void need_args(const Args& a, size_t count, const NodePtr& n, const std::string& msg) { if (a.size() != count) err(n, msg); }
  std::string scalar_arg(const Value& v, const NodePtr& n) { return convert(to_base(v, Base::Float, n), scalar_of(Base::Float), n).comps.at(0); }

  // ---- screen-space derivatives and texture sampling (pixel shaders)
  void need_quad(const NodePtr& n, const std::string& what) {
    if (stage != "ps") err(n, what + " is only available in pixel shaders");
    if (divergent) err(n, what + " under divergent control flow is not supported (the four pixels of a quad must reach it together)");
    refl.uses_derivatives = true;
    cur_fn_derivs = true;
  }
  std::string sampler_slot(const Value& v, const NodePtr& n) {
    if (v.type.kind != Kind::Sampler) err(n, "first argument must be a sampler");
    return v.comps.at(0);
  }
  Value tex_result(const std::function<std::string(const std::string&)>& call) {
    ++ntemp;
    std::vector<std::string> r;
    for (int k = 0; k < 4; ++k) r.push_back("x" + std::to_string(ntemp) + "_" + std::to_string(k));
    emit("float " + join(r, ", ") + ";");
    emit(call(join(r, ", ")));
    return make_value(vec(Base::Float, 4), r);
  }
  Value tex2d(const Value& samp, const Value& coord, const NodePtr& n) {  // SASL tex2D == sample_2d_grad with the quad's per-pixel derivatives
    if (stage == "vs") err(n, "vertex shaders sample with tex2Dlod");
    need_quad(n, "tex2D");
    const std::string s = sampler_slot(samp, n);
    const Value uv = convert(to_base(coord, Base::Float, n), vec(Base::Float, 2), n);
    const std::string &u = uv.comps[0], &v = uv.comps[1];
    return tex_result([&](const std::string& r) {
      return "sasl_tex2d_grad(p, px, " + s + ", " + u + ", " + v + ", sasl_ddx(px, " + u + "), sasl_ddx(px, " + v + "), sasl_ddy(px, " + u + "), sasl_ddy(px, " + v + "), 0.0f, " + r + ");";
    });
  }

  Value intrinsic(const std::string& name, const Args& a, const NodePtr& n) {
    const Type F1 = scalar_of(Base::Float);
    auto f1 = [&](const Fmt1& fmt) { return map1(a, n, fmt); };
    auto fn = [&](size_t count, const FmtN& fmt) { return mapn(a, n, fmt, count); };
    if (name == "abs") {
      if (a.size() == 1 && a[0].type.base != Base::Float) return map1(a, n, [](const std::string& c) { return "abs(" + c + ")"; }, Base::Int);
      return f1([](const std::string& c) { return "fabsf(" + c + ")"; });
    }
    if (name == "rsqrt") return f1([](const std::string& c) { return "(1.0f / sqrtf(" + c + "))"; });
    if (name == "frac") return f1([](const std::string& c) { return "(fabsf(" + c + ") - sasl_m_floor(fabsf(" + c + ")))"; });
    if (name == "ldexp") return fn(2, [](const std::vector<std::string>& c) { return "sasl_m_ldexp(" + c[0] + ", " + c[1] + ")"; });
    if (name == "saturate") return f1([](const std::string& c) { return "sasl_clamp(" + c + ", 0.0f, 1.0f)"; });
    if (name == "sign") return f1([](const std::string& c) { return "((" + c + " > 0.0f) ? 1.0f : ((" + c + " < 0.0f) ? -1.0f : 0.0f))"; });
    if (name == "radians") return f1([](const std::string& c) { return "(" + c + " * 0.017453292519943295f)"; });
    if (name == "degrees") return f1([](const std::string& c) { return "(" + c + " * 57.29577951308232f)"; });
    if (name == "min") return fn(2, [](const std::vector<std::string>& c) { return "fminf(" + c[0] + ", " + c[1] + ")"; });
    if (name == "max") return fn(2, [](const std::vector<std::string>& c) { return "fmaxf(" + c[0] + ", " + c[1] + ")"; });
    if (name == "pow") return fn(2, [](const std::vector<std::string>& c) { return "sasl_m_pow(" + c[0] + ", " + c[1] + ")"; });
    if (name == "fmod") return fn(2, [](const std::vector<std::string>& c) { return "fmodf(" + c[0] + ", " + c[1] + ")"; });
    if (name == "atan2") return fn(2, [](const std::vector<std::string>& c) { return "sasl_m_atan2(" + c[0] + ", " + c[1] + ")"; });
    if (name == "step") return fn(2, [](const std::vector<std::string>& c) { return "((" + c[1] + " >= " + c[0] + ") ? 1.0f : 0.0f)"; });
    if (name == "clamp") return fn(3, [](const std::vector<std::string>& c) { return "sasl_clamp(" + c[0] + ", " + c[1] + ", " + c[2] + ")"; });
    if (name == "mad") return fn(3, [](const std::vector<std::string>& c) { return "((" + c[0] + " * " + c[1] + ") + " + c[2] + ")"; });
    if (name == "smoothstep") return fn(3, [](const std::vector<std::string>& c) { return "sasl_smoothstep(" + c[0] + ", " + c[1] + ", " + c[2] + ")"; });
    if (name == "rcp") return f1([](const std::string& c) { return "(1.0f / " + c + ")"; });
    if (name == "lerp") {
      need_args(a, 3, n, "lerp expects 3 arguments");
      const Value d = bin_values("-", a[1], a[0], n);
      const Value scaled = bin_values("*", d, a[2], n);
      return bin_values("+", a[0], scaled, n);
    }
    if (name == "dot") {
      need_args(a, 2, n, "dot expects 2 arguments");
      Value x = to_base(a[0], Base::Float, n), y = to_base(a[1], Base::Float, n);
      unify(x, y, n);
      return make_value(F1, {dot_comps(x.comps, y.comps)});
    }
    if (name == "cross") {
      need_args(a, 2, n, "cross expects 2 arguments");
      const Value x = convert(to_base(a[0], Base::Float, n), vec(Base::Float, 3), n);
      const Value y = convert(to_base(a[1], Base::Float, n), vec(Base::Float, 3), n);
      const std::string &ax = x.comps[0], &ay = x.comps[1], &az = x.comps[2], &bx = y.comps[0], &by = y.comps[1], &bz = y.comps[2];
      const std::string c0 = temp(Base::Float, "(" + ay + " * " + bz + ") - (" + az + " * " + by + ")");
      const std::string c1 = temp(Base::Float, "(" + az + " * " + bx + ") - (" + ax + " * " + bz + ")");
      const std::string c2 = temp(Base::Float, "(" + ax + " * " + by + ") - (" + ay + " * " + bx + ")");
      return make_value(vec(Base::Float, 3), {c0, c1, c2});
    }
    if (name == "dst") {  // distance vector: (1, a.y * b.y, a.z, b.w)
      need_args(a, 2, n, "dst(float4, float4)");
      const Value x = convert(to_base(a[0], Base::Float, n), vec(Base::Float, 4), n);
      const Value y = convert(to_base(a[1], Base::Float, n), vec(Base::Float, 4), n);
      const std::string one = temp(Base::Float, "1.0f");
      const std::string yy = temp(Base::Float, x.comps[1] + " * " + y.comps[1]);
      return make_value(vec(Base::Float, 4), {one, yy, x.comps[2], y.comps[3]});
    }
    if (name == "length" || name == "distance") {
      Value v;
      if (name == "distance") { need_args(a, 2, n, "distance expects 2 arguments"); v = to_base(bin_values("-", a[0], a[1], n), Base::Float, n); }
      else { if (a.empty()) err(n, "length expects 1 argument"); v = to_base(a[0], Base::Float, n); }
      return make_value(F1, {temp(Base::Float, "sqrtf(" + dot_comps(v.comps, v.comps) + ")")});
    }
    if (name == "normalize") {  // eflib normalize3: zero-length vectors are left alone (length := 1)
      if (a.empty()) err(n, "normalize expects 1 argument");
      const Value v = to_base(a[0], Base::Float, n);
      std::string ln = temp(Base::Float, "sqrtf(" + dot_comps(v.comps, v.comps) + ")");
      ln = temp(Base::Float, "sasl_eq_eps(" + ln + ", 0.0f) ? 1.0f : " + ln);
      const std::string inv = temp(Base::Float, "1.0f / " + ln);
      std::vector<std::string> out;
      for (const auto& c : v.comps) out.push_back(temp(Base::Float, c + " * " + inv));
      return make_value(v.type, out);
    }
    if (name == "reflect") {  // eflib reflect3(i, n) = i - 2 * dot(i, n) * n
      need_args(a, 2, n, "reflect expects 2 arguments");
      Value i = to_base(a[0], Base::Float, n), nn = to_base(a[1], Base::Float, n);
      unify(i, nn, n);
      const std::string d = dot_comps(i.comps, nn.comps);
      const std::string s = temp(Base::Float, "2.0f * " + d);
      std::vector<std::string> out;
      for (size_t k = 0; k < i.comps.size() && k < nn.comps.size(); ++k) out.push_back(temp(Base::Float, i.comps[k] + " - (" + s + " * " + nn.comps[k] + ")"));
      return make_value(i.type, out);
    }
    if (name == "mul") {
      need_args(a, 2, n, "mul expects 2 arguments");
      const Value x = to_base(a[0], Base::Float, n), y = to_base(a[1], Base::Float, n);
      const Type &tx = x.type, &ty = y.type;
      if (tx.kind != Kind::Matrix && ty.kind != Kind::Matrix) return bin_values("*", x, y, n);
      std::vector<std::string> out;
      auto prod = [&](size_t xi, size_t yi) { return temp(Base::Float, x.comps[xi] + " * " + y.comps[yi]); };
      if (tx.kind == Kind::Vector && ty.kind == Kind::Matrix) {  // row vector x matrix (eflib transform)
        if (tx.cols != ty.rows) err(n, "mul(" + to_string(tx) + ", " + to_string(ty) + ")");
        for (int j = 0; j < ty.cols; ++j) {
          std::vector<std::string> terms;
          for (int i = 0; i < ty.rows; ++i) terms.push_back(prod((size_t)i, (size_t)(i * ty.cols + j)));
          out.push_back(sum_lr(terms));
        }
        return make_value(vec(Base::Float, ty.cols), out);
      }
      if (tx.kind == Kind::Matrix && ty.kind == Kind::Vector) {
        if (tx.cols != ty.cols) err(n, "mul(" + to_string(tx) + ", " + to_string(ty) + ")");
        for (int i = 0; i < tx.rows; ++i) {
          std::vector<std::string> terms;
          for (int j = 0; j < tx.cols; ++j) terms.push_back(prod((size_t)(i * tx.cols + j), (size_t)j));
          out.push_back(sum_lr(terms));
        }
        return make_value(vec(Base::Float, tx.rows), out);
      }
      if (tx.kind == Kind::Matrix && ty.kind == Kind::Matrix) {
        if (tx.cols != ty.rows) err(n, "mul(" + to_string(tx) + ", " + to_string(ty) + ")");
        for (int i = 0; i < tx.rows; ++i)
          for (int j = 0; j < ty.cols; ++j) {
            std::vector<std::string> terms;
            for (int k = 0; k < tx.cols; ++k) terms.push_back(prod((size_t)(i * tx.cols + k), (size_t)(k * ty.cols + j)));
            out.push_back(sum_lr(terms));
          }
        return make_value(mat(Base::Float, tx.rows, ty.cols), out);
      }
      return bin_values("*", x, y, n);  // scalar * matrix
    }
    if (name == "transpose") {
      if (a.empty() || a[0].type.kind != Kind::Matrix) err(n, "transpose expects a matrix");
      const Value& m = a[0];
      std::vector<std::string> out;
      for (int c = 0; c < m.type.cols; ++c) for (int r = 0; r < m.type.rows; ++r) out.push_back(m.comps[(size_t)(r * m.type.cols + c)]);
      return make_value(mat(m.type.base, m.type.cols, m.type.rows), out);
    }
    if (name == "any" || name == "all") {
      if (a.empty()) err(n, name + " expects 1 argument");
      const Value v = to_base(a[0], Base::Bool, n);
      return make_value(scalar_of(Base::Bool), {temp(Base::Bool, join(v.comps, name == "any" ? " || " : " && "))});
    }
    if (name == "asfloat" || name == "asint" || name == "asuint") {
      if (a.empty()) err(n, name + " expects 1 argument");
      const Base base = name == "asfloat" ? Base::Float : name == "asint" ? Base::Int : Base::Uint;
      std::vector<std::string> out;
      for (const auto& c : a[0].comps) out.push_back(temp(base, "sasl_" + name + "(" + c + ")"));
      return make_value(with_base(a[0].type, base), out);
    }
    if (name == "countbits" || name == "count_bits") {  // both spellings are registered upstream (semantic_analyser.cpp:1981-1982)
      if (a.empty()) err(n, "countbits expects 1 argument");
      const Value v = to_base(a[0], Base::Uint, n);
      std::vector<std::string> out;
      for (const auto& c : v.comps) out.push_back(temp(Base::Uint, "sasl_countbits(" + c + ")"));
      return make_value(v.type, out);
    }
    if (name == "firstbithigh" || name == "firstbitlow" || name == "reversebits") {  // the result keeps the argument's int / uint base
      if (a.size() != 1 || (a[0].type.base != Base::Int && a[0].type.base != Base::Uint)) err(n, "expects one int or uint argument");
      const Value& v = a[0];
      std::vector<std::string> out;
      for (const auto& c : v.comps) out.push_back(temp(v.type.base, std::string("(") + (v.type.base == Base::Int ? "int" : "unsigned") + ")sasl_" + name + "((unsigned)" + c + ")"));
      return make_value(v.type, out);
    }
    if (name == "isinf" || name == "isfinite" || name == "isnan") {  // cgs.cpp:1828-1842
      if (a.size() != 1) err(n, "expects 1 argument");
      const Value v = to_base(a[0], Base::Float, n);
      std::vector<std::string> out;
      for (const auto& c : v.comps) {
        const std::string e = name == "isinf" ? "(fabsf(" + c + ") == sasl_asfloat(0x7F800000u))"
                              : name == "isfinite" ? "(!(fabsf(" + c + ") == sasl_asfloat(0x7F800000u)) && (" + c + " == " + c + "))"
                                                   : "(" + c + " != " + c + ")";
        out.push_back(temp(Base::Bool, e));
      }
      return make_value(with_base(v.type, Base::Bool), out);
    }
    if (name == "refract") {  // cg_impl.cpp:1229-1269, in the reference's order of operations
      need_args(a, 3, n, "refract(i, n, eta)");
      Value i = to_base(a[0], Base::Float, n), nn = to_base(a[1], Base::Float, n);
      unify(i, nn, n);
      const std::string eta = scalar_arg(a[2], n);
      const std::string eta2 = temp(Base::Float, eta + " * " + eta);
      const std::string ndi = dot_comps(nn.comps, i.comps);
      std::vector<std::string> eta_i;
      for (const auto& c : i.comps) eta_i.push_back(temp(Base::Float, eta + " * " + c));
      std::string k = temp(Base::Float, ndi + " * " + ndi);
      k = temp(Base::Float, "1.0f - " + k);
      k = temp(Base::Float, eta2 + " * " + k);
      k = temp(Base::Float, "1.0f - " + k);
      const std::string flag = temp(Base::Bool, k + " < 0.0f");
      k = temp(Base::Float, flag + " ? 0.0f : " + k);
      std::string r = temp(Base::Float, eta + " * " + ndi);
      r = temp(Base::Float, r + " + sqrtf(" + k + ")");
      std::vector<std::string> out;
      for (size_t j = 0; j < eta_i.size() && j < nn.comps.size(); ++j) {
        std::string t = temp(Base::Float, r + " * " + nn.comps[j]);
        t = temp(Base::Float, eta_i[j] + " - " + t);
        out.push_back(temp(Base::Float, flag + " ? 0.0f : " + t));
      }
      return make_value(i.type, out);
    }
    if (name == "faceforward") {  // cg_impl.cpp:1300-1319: dot(i, ng) < 0 ? n : 0 - n
      need_args(a, 3, n, "faceforward(n, i, ng)");
      Value nn = to_base(a[0], Base::Float, n), i = to_base(a[1], Base::Float, n);
      unify(nn, i, n);
      const Value ng = convert(to_base(a[2], Base::Float, n), i.type, n);
      const std::string d = dot_comps(i.comps, ng.comps);
      const std::string flag = temp(Base::Bool, d + " < 0.0f");
      std::vector<std::string> out;
      for (const auto& c : nn.comps) out.push_back(temp(Base::Float, flag + " ? " + c + " : (0.0f - " + c + ")"));
      return make_value(nn.type, out);
    }
    if (name == "lit") {  // cg_impl.cpp:1320-1352: (1, max(n.l, 0), n.l < 0 || n.h < 0 ? 0 : n.h * m, 1)
      need_args(a, 3, n, "lit(n_dot_l, n_dot_h, m)");
      const std::string l = scalar_arg(a[0], n), h = scalar_arg(a[1], n), m = scalar_arg(a[2], n);
      const std::string diffuse = temp(Base::Float, "(" + l + " < 0.0f) ? 0.0f : " + l);
      const std::string spec = temp(Base::Float, "((" + l + " < 0.0f) || (" + h + " < 0.0f)) ? 0.0f : (" + h + " * " + m + ")");
      const std::string one0 = temp(Base::Float, "1.0f");
      const std::string one1 = temp(Base::Float, "1.0f");
      return make_value(vec(Base::Float, 4), {one0, diffuse, spec, one1});
    }
    if (name == "ddx" || name == "ddy") {
      need_quad(n, name);
      return f1([&](const std::string& c) { return "sasl_" + name + "(px, " + c + ")"; });
    }
    if (name == "tex2D") {
      need_args(a, 2, n, "tex2D(sampler, uv)");
      return tex2d(a[0], a[1], n);
    }
    if (name == "tex2Dgrad") {
      if (stage != "ps" || a.size() != 4) err(n, "tex2Dgrad(sampler, uv, ddx, ddy) in a pixel shader");
      const std::string s = sampler_slot(a[0], n);
      const Value uv = convert(to_base(a[1], Base::Float, n), vec(Base::Float, 2), n);
      const Value dx = convert(to_base(a[2], Base::Float, n), vec(Base::Float, 2), n);
      const Value dy = convert(to_base(a[3], Base::Float, n), vec(Base::Float, 2), n);
      return tex_result([&](const std::string& r) {
        return "sasl_tex2d_grad(p, px, " + s + ", " + uv.comps[0] + ", " + uv.comps[1] + ", " + dx.comps[0] + ", " + dx.comps[1] + ", " + dy.comps[0] + ", " + dy.comps[1] + ", 0.0f, " + r + ");";
      });
    }
    if (name == "tex2Dbias") {
      need_args(a, 2, n, "tex2Dbias(sampler, float4(uv, _, bias))");
      need_quad(n, "tex2Dbias");
      const std::string s = sampler_slot(a[0], n);
      const Value c = convert(to_base(a[1], Base::Float, n), vec(Base::Float, 4), n);
      const std::string &u = c.comps[0], &v = c.comps[1];
      return tex_result([&](const std::string& r) {
        return "sasl_tex2d_grad(p, px, " + s + ", " + u + ", " + v + ", sasl_ddx(px, " + u + "), sasl_ddx(px, " + v + "), sasl_ddy(px, " + u + "), sasl_ddy(px, " + v + "), " + c.comps[3] + ", " + r + ");";
      });
    }
    if (name == "tex2Dlod") {
      need_args(a, 2, n, "tex2Dlod(sampler, float4(uv, _, lod))");
      const std::string s = sampler_slot(a[0], n);
      const Value c = convert(to_base(a[1], Base::Float, n), vec(Base::Float, 4), n);
      if (stage == "vs")  // sasl.vs.tex2d.lod = sampler::sample_2d_lod(coord.xy, coord.w) (sampler_api.cpp:50-52)
        return tex_result([&](const std::string& r) { return "sasl_vs_tex2d_lod(S0, " + s + ", " + c.comps[0] + ", " + c.comps[1] + ", " + c.comps[3] + ", " + r + ");"; });
      return tex_result([&](const std::string& r) { return "sasl_tex2d_lod(p, px, " + s + ", " + c.comps[0] + ", " + c.comps[1] + ", " + c.comps[3] + ", " + r + ");"; });
    }
    if (name == "tex2Dproj") {
      need_args(a, 2, n, "tex2Dproj(sampler, float4)");
      const Value c = convert(to_base(a[1], Base::Float, n), vec(Base::Float, 4), n);
      const std::string inv = temp(Base::Float, "1.0f / " + c.comps[3]);
      const std::string u = temp(Base::Float, c.comps[0] + " * " + inv);
      const std::string v = temp(Base::Float, c.comps[1] + " * " + inv);
      return tex2d(a[0], make_value(vec(Base::Float, 2), {u, v}), n);
    }
    err(n, "unknown function " + quoted(name));
  }

  // ---- statements
  void stmt(const NodePtr& n) {
    const std::string& op = n->op;
    if (op == "block") s_block(n->kids);
    else if (op == "decl") s_decl(n);
    else if (op == "expr") expr(n->kids[0]);
    else if (op == "if") s_if(n);
    else if (op == "for") s_for(n);
    else if (op == "dowhile") s_dowhile(n);
    else if (op == "switch") s_switch(n);
    else if (op == "break") s_break(n);
    else if (op == "continue") s_continue(n);
    else if (op == "return") s_return(n);
    else err(n, "unexpected statement " + op);
  }
  void s_block(const std::vector<NodePtr>& stmts) {
    scopes.emplace_back();
    for (const auto& s : stmts) stmt(s);
    scopes.pop_back();
  }
  void s_decl(const NodePtr& n) {
    for (const auto& d : n->decls) {
      if (d.type.kind == Kind::Void || d.type.kind == Kind::Sampler) err(n, "cannot declare a local of type " + to_string(d.type));
      Value init;
      if (d.init) init = materialize(expr(d.init));
      const Value v = declare(d.type, d.name);
      if (d.init) store(v, init, n);
    }
  }
  void s_if(const NodePtr& n) {
    const Value c = convert(expr(n->kids[0]), scalar_of(Base::Bool), n);
    emit("if (" + c.comps[0] + ") {");
    ++indent;
    ++divergent;
    s_block({n->kids[1]});
    --indent;
    if (n->kids[2]) {
      emit("} else {");
      ++indent;
      s_block({n->kids[2]});
      --indent;
    }
    --divergent;
    emit("}");
  }
  // Body inside `do { } while (0)`: `continue` leaves it (the step still runs), `break` raises the loop's flag
  void loop_body(const NodePtr& body) {
    emit("do {");
    ++indent;
    s_block({body});
    --indent;
    emit("} while (0);");
    emit("if (brk" + std::to_string(loop_ids.back()) + ") break;");
  }
  void s_for(const NodePtr& n) {
    const NodePtr &init = n->kids[0], &cond = n->kids[1], &step = n->kids[2], &body = n->kids[3];
    scopes.emplace_back();
    emit("{");
    ++indent;
    if (init) stmt(init);
    loop_ids.push_back(++ntemp);
    emit("bool brk" + std::to_string(ntemp) + " = false;");
    emit("for (;;) {");
    ++indent;
    ++divergent;
    if (cond) {
      const Value c = convert(expr(cond), scalar_of(Base::Bool), n);
      emit("if (!" + c.comps[0] + ") break;");
    }
    loop_body(body);
    if (step) expr(step);
    --divergent;
    --indent;
    emit("}");
    loop_ids.pop_back();
    --indent;
    emit("}");
    scopes.pop_back();
  }
  void s_dowhile(const NodePtr& n) {
    loop_ids.push_back(++ntemp);
    emit("bool brk" + std::to_string(ntemp) + " = false;");
    emit("for (;;) {");
    ++indent;
    ++divergent;
    loop_body(n->kids[0]);
    const Value c = convert(expr(n->kids[1]), scalar_of(Base::Bool), n);
    emit("if (!" + c.comps[0] + ") break;");
    --divergent;
    --indent;
    emit("}");
    loop_ids.pop_back();
  }
  // Case labels are integer literals (optionally negated)
  long long case_value(NodePtr e, const NodePtr& n) {
    bool neg = false;
    while (e->op == "un" && (e->str == "-" || e->str == "+")) { neg ^= e->str == "-"; e = e->kids[0]; }
    const std::string t = e->op == "num" ? e->str : "";
    long long v = 0;
    const std::string hex = rstrip_set(t, "uU");
    if (lower(t).compare(0, 2, "0x") == 0 && hex.size() > 2 && t.size() - hex.size() <= 1 &&
        std::all_of(hex.begin() + 2, hex.end(), [](unsigned char c) { return std::isxdigit(c); }))
      v = (long long)std::strtoull(hex.c_str(), nullptr, 16);
    else if (!t.empty() && (all_digits(t) || (t.size() > 1 && std::string("uUlL").find(t.back()) != std::string::npos && all_digits(t.substr(0, t.size() - 1)))))
      v = std::atoll(t.c_str());
    else err(n, "case labels must be integer literals");
    return neg ? -v : v;
  }
  // switch with C fall-through semantics, lowered to guarded blocks inside one `do { } while (0)`: a group runs when an
  // earlier group fell through into it or the selector equals one of its labels (`default`: none of the switch's labels);
  // `break` leaves the do-while, `continue` (inside a loop) leaves it with the loop's continue flag raised
  void s_switch(const NodePtr& n) {
    const Value sel = convert(expr(n->kids[0]), scalar_of(Base::Int), n);
    const int sid = ++ntemp;
    const std::string S = std::to_string(sid);
    std::vector<std::vector<std::pair<bool, long long>>> labels;  // (is default, value)
    std::vector<long long> all_vals;
    int n_default = 0;
    for (const auto& g : n->groups) {
      labels.emplace_back();
      for (const auto& l : g.labels) {
        if (!l) { labels.back().push_back({true, 0}); ++n_default; }
        else { const long long v = case_value(l, n); labels.back().push_back({false, v}); all_vals.push_back(v); }
      }
    }
    if (std::set<long long>(all_vals.begin(), all_vals.end()).size() != all_vals.size()) err(n, "duplicate case label");
    if (n_default > 1) err(n, "more than one default label");
    emit("{");
    ++indent;
    emit("const int sel" + S + " = " + sel.comps[0] + ";");
    emit("bool fall" + S + " = false;");
    if (!loop_ids.empty()) emit("bool cnt" + S + " = false;");
    emit("do {");
    ++indent;
    ++divergent;
    switch_ids.push_back({sid, loop_ids.size()});
    for (size_t g = 0; g < n->groups.size(); ++g) {
      std::vector<std::string> conds;
      bool has_default = false;
      for (const auto& l : labels[g]) {
        if (l.first) has_default = true;
        else conds.push_back("sel" + S + " == " + std::to_string(l.second));
      }
      if (has_default) {
        std::vector<std::string> none;
        for (long long v : all_vals) none.push_back("sel" + S + " != " + std::to_string(v));
        if (none.empty()) none.push_back("true");
        conds.push_back("(" + join(none, " && ") + ")");
      }
      emit("if (fall" + S + " || " + join(conds, " || ") + ") {");
      ++indent;
      emit("fall" + S + " = true;");
      s_block(n->groups[g].stmts);
      --indent;
      emit("}");
    }
    switch_ids.pop_back();
    --divergent;
    --indent;
    emit("} while (0);");
    if (!loop_ids.empty()) emit("if (cnt" + S + ") break;");  // `continue` inside the switch: leave the loop body's do { } while (0)
    --indent;
    emit("}");
  }
  // The innermost breakable construct is a switch (opened after the innermost loop)
  bool in_switch() const { return !switch_ids.empty() && switch_ids.back().second == loop_ids.size(); }
  void s_break(const NodePtr& n) {
    if (in_switch()) { emit("break;"); return; }  // leaves the switch's do { } while (0)
    if (loop_ids.empty()) err(n, "break outside a loop or switch");
    emit("brk" + std::to_string(loop_ids.back()) + " = true; break;");
  }
  void s_continue(const NodePtr& n) {
    if (loop_ids.empty()) err(n, "continue outside a loop");
    if (in_switch())
      for (const auto& sw : switch_ids)  // every switch opened inside the innermost loop hands the request outwards
        if (sw.second == loop_ids.size()) emit("cnt" + std::to_string(sw.first) + " = true;");
    emit("break;");  // leaves the do { } while (0) around the body; the loop's step still runs
  }
  void s_return(const NodePtr& n) {
    if (n->kids[0]) {
      if (!has_ret) err(n, "void function returns a value");
      store(cur_ret, expr(n->kids[0]), n);
    }
    emit("return;");
  }

  // ---- functions
  template <class F>
  static void walk(const NodePtr& n, const F& visit) {  // pre-order, children in source order
    if (!n) return;
    visit(n);
    for (const auto& d : n->decls) walk(d.init, visit);
    for (const auto& k : n->kids) walk(k, visit);
    for (const auto& g : n->groups) {
      for (const auto& l : g.labels) walk(l, visit);
      for (const auto& s : g.stmts) walk(s, visit);
    }
  }
  std::set<std::string> called_functions(const Func& f) const {
    std::set<std::string> names, out;
    for (const auto& g : funcs) names.insert(g.name);
    walk(f.body, [&](const NodePtr& x) { if (x->op == "call" && names.count(x->str)) out.insert(x->str); });
    return out;
  }
  // Callees before callers (a function may be used before its definition), restricted to what the entry reaches (every
  // function for a library unit); functions on a call cycle land in `recursive`: they get a prototype ahead of all bodies
  std::vector<const Func*> generation_order() {
    std::map<std::string, const Func*> by_name;
    std::map<std::string, std::set<std::string>> callees;
    for (const auto& f : funcs) { by_name[f.name] = &f; callees[f.name] = called_functions(f); }
    std::vector<const Func*> order;
    std::map<std::string, int> state;
    std::vector<std::string> stack;
    std::function<void(const Func*)> dfs = [&](const Func* f) {
      state[f->name] = 1;
      stack.push_back(f->name);
      for (const auto& c : callees[f->name]) {
        if (state[c] == 0) dfs(by_name[c]);
        else if (state[c] == 1) recursive.insert(std::find(stack.begin(), stack.end(), c), stack.end());
      }
      stack.pop_back();
      state[f->name] = 2;
      order.push_back(f);
    };
    if (entry) { if (state[entry->name] == 0) dfs(by_name[entry->name]); }
    else for (const auto& f : funcs) if (state[f.name] == 0) dfs(by_name[f.name]);
    return order;
  }
  std::string fn_head(const Func& f, const std::vector<std::string>& params) const {
    std::vector<std::string> all = stage == "ps" ? std::vector<std::string>{"const SaslUniforms& U", "const slv::RasterParams& p", "const Ctx& px"}
                                                 : std::vector<std::string>{"const SaslUniforms& U", "const SaslSampler& S0"};
    all.insert(all.end(), params.begin(), params.end());
    return std::string(stage == "ps" ? "template <class Ctx>\n" : "") + (recursive.count(f.name) ? "SASL_FN_REC" : "SASL_FN") + " void sasl_fn_" + f.name + "(" + join(all, ", ") + ")";
  }
  // The flattened C++ parameter list of f: parameter declarations, scope of the parameters, names of the results
  void fn_signature(const Func& f, std::vector<std::string>& params, std::map<std::string, Value>& scope, std::vector<std::string>& ret_names) const {
    for (const auto& prm : f.params) {
      if (prm.type.kind == Kind::Sampler) {
        const std::string nm = "a_" + prm.name;
        params.push_back("const int " + nm);
        scope[prm.name] = make_value(make_type(Kind::Sampler), {nm});
        continue;
      }
      std::vector<std::string> names;
      const auto bases = flat_types(prm.type);
      for (size_t k = 0; k < bases.size(); ++k) {
        names.push_back("a_" + prm.name + "_" + std::to_string(k));
        params.push_back(std::string(c_base(bases[k])) + " " + names.back());
      }
      scope[prm.name] = make_value(prm.type, names, true);
    }
    if (f.ret.kind != Kind::Void) {
      const auto bases = flat_types(f.ret);
      for (size_t k = 0; k < bases.size(); ++k) {
        ret_names.push_back("ret_" + std::to_string(k));
        params.push_back(std::string(c_base(bases[k])) + "& ret_" + std::to_string(k));
      }
    }
  }
  // Names of uniform globals (not samplers, not arrays) that `body` stores to, in order of first store; names shadowed by a
  // parameter or by a declaration anywhere in the function are left alone (conservative)
  std::vector<std::string> assigned_globals(const NodePtr& body, const std::map<std::string, Value>& params) const {
    std::set<std::string> declared;
    for (const auto& kv : params) declared.insert(kv.first);
    std::vector<std::string> out;
    auto root = [](NodePtr x) -> std::string {
      while (x && (x->op == "member" || x->op == "index")) x = x->kids[0];
      return x && x->op == "var" ? x->str : std::string();
    };
    auto note = [&](const std::string& r) { if (!r.empty() && std::find(out.begin(), out.end(), r) == out.end()) out.push_back(r); };
    walk(body, [&](const NodePtr& x) {
      for (const auto& d : x->decls) declared.insert(d.name);
      if (x->op == "assign" || x->op == "postinc") note(root(x->kids[0]));
    });
    std::vector<std::string> kept;
    for (const auto& r : out) {
      auto u = uniform_vars.find(r);
      if (declared.count(r) || u == uniform_vars.end() || uniform_arrays.count(r) || !u->second.type.numeric()) continue;
      kept.push_back(r);
    }
    return kept;
  }
  void gen_function(const Func& f) {
    cur_fn_derivs = false;
    if (recursive.count(f.name)) fn_table[f.name] = &f;  // visible to its own body (and to the other members of its cycle)
    std::vector<std::string> params, ret_names;
    std::map<std::string, Value> scope;
    fn_signature(f, params, scope, ret_names);
    has_ret = !ret_names.empty();
    cur_ret = make_value(f.ret, ret_names, true);
    const std::string head = fn_head(f, params) + " {";
    const size_t start = lines.size();
    indent = 1;
    scopes.clear();
    scopes.push_back(scope);
    // a global the function assigns to (sasl/test/repo/input_assigned.svs: `x += 0.5f`): the uniform block is read-only and
    // shared, so the function works on its own copy, initialised from the uniform - the write is local to the invocation
    for (const auto& name : assigned_globals(f.body, scope)) {
      const Value src = uniform_vars.at(name);
      const Value copy = declare(src.type, name);
      store(copy, src, f.body);
    }
    s_block(f.body->kids);
    std::vector<std::string> body(lines.begin() + start, lines.end());
    lines.resize(start);
    lines.push_back(head);
    lines.insert(lines.end(), body.begin(), body.end());
    lines.push_back("}");
    lines.push_back("");
    if (cur_fn_derivs) {
      if (recursive.count(f.name)) err(f.line, f.name + ": screen-space derivatives in a recursive function");
      fns_with_derivatives.insert(f.name);
    }
    fn_table[f.name] = &f;
  }

  // ---- translation unit
  unit run() {
    // globals: uniforms (packed 16-byte aligned, in declaration order) and samplers
    size_t off = 0;
    std::vector<std::string> fields;
    for (const auto& g : globals) {
      if (g.type.kind == Kind::Sampler) {
        uniform_vars[g.name] = make_value(make_type(Kind::Sampler), {std::to_string(refl.samplers.size())});
        refl.samplers.push_back(g.name);
        continue;
      }
      if (g.type.kind == Kind::Struct) err(g.line, "global " + g.name + ": struct uniforms are not supported");
      if (g.array) {
        // an array uniform lives in a buffer of its own (bone palettes do not fit the 256-byte block): the block holds its
        // ADDRESS - device memory for the product (slv_buffer_device_ptr), host memory for host-compiled code
        if (!g.type.numeric() || g.type.base == Base::Bool) err(g.line, "global " + g.name + ": arrays of " + to_string(g.type) + " are not supported");
        if (!g.array_len.empty()) {
          bool found = false;
          for (const auto& x : globals) if (x.name == g.array_len && x.type.kind == Kind::Scalar && (x.type.base == Base::Int || x.type.base == Base::Uint)) found = true;
          if (!found) err(g.line, "global " + g.name + ": the array size " + quoted(g.array_len) + " is not an integer global");
        }
        off = (off + 15) & ~(size_t)15;
        fields.push_back(std::string("  alignas(16) const ") + c_base(g.type.base) + "* " + g.name + ";");
        uniform_arrays[g.name] = g.type;
        refl.uniforms.push_back({g.name, to_string(g.type) + "[]", off, 8});
        off += 8;
        continue;
      }
      const int n = g.type.n();
      off = (off + 15) & ~(size_t)15;
      const char* cb = c_base(g.type.base == Base::Bool ? Base::Int : g.type.base);
      fields.push_back(std::string("  alignas(16) ") + cb + " " + g.name + "[" + std::to_string(n) + "];");
      std::vector<std::string> comps;
      for (int k = 0; k < n; ++k) {
        std::string c = "U." + g.name + "[" + std::to_string(k) + "]";
        if (g.type.base == Base::Bool) c = "(" + c + " != 0)";
        comps.push_back(c);
      }
      uniform_vars[g.name] = make_value(g.type, comps);
      refl.uniforms.push_back({g.name, to_string(g.type), off, (size_t)(4 * n)});
      off += (size_t)(4 * n);
    }
    refl.uniform_bytes = (off + 15) & ~(size_t)15;
    std::vector<std::string> header = {"struct SaslUniforms {"};
    if (fields.empty()) fields.push_back("  int unused_;");
    header.insert(header.end(), fields.begin(), fields.end());
    header.push_back("};");
    header.push_back("");
    const auto order = generation_order();
    for (const Func* f : order)  // prototypes of the functions on call cycles, ahead of every body
      if (recursive.count(f->name)) {
        fn_table[f->name] = f;
        std::vector<std::string> params, rn;
        std::map<std::string, Value> sc;
        fn_signature(*f, params, sc, rn);
        lines.push_back(fn_head(*f, params) + ";");
      }
    if (!recursive.empty()) lines.push_back("");
    for (const Func* f : order) gen_function(*f);
    std::vector<std::string> wrapper;
    if (stage == "vs") wrapper = gen_vs_wrapper();
    else if (stage == "ps") wrapper = gen_ps_wrapper();
    std::vector<std::string> all = header;
    all.insert(all.end(), lines.begin(), lines.end());
    all.insert(all.end(), wrapper.begin(), wrapper.end());
    refl.code = join(all, "\n") + "\n";
    return refl;
  }

  struct IoMember { std::string name; Type type; Semantic sem; };
  // Flattened (name, type, semantic) lists of the entry's inputs and outputs
  void entry_io(std::vector<IoMember>& ins, std::vector<IoMember>& outs) const {
    const Func& f = *entry;
    for (const auto& prm : f.params) {
      if (prm.type.kind == Kind::Struct) for (const auto& m : members(prm.type.name)) ins.push_back({m.name, m.type, norm_semantic(m.has_semantic, m.semantic)});
      else if (prm.type.kind == Kind::Sampler) throw compile_error("the entry function cannot take a sampler");
      else ins.push_back({prm.name, prm.type, norm_semantic(prm.has_semantic, prm.semantic)});
    }
    if (f.ret.kind == Kind::Struct) for (const auto& m : members(f.ret.name)) outs.push_back({m.name, m.type, norm_semantic(m.has_semantic, m.semantic)});
    else if (f.ret.kind != Kind::Void) outs.push_back({"ret", f.ret, norm_semantic(f.has_ret_semantic, f.ret_semantic)});
  }
  static bool is_position(const Semantic& s) { return s.valid && (s.name == "SV_POSITION" || s.name == "POSITION"); }
  static std::string zeros_decl(const std::vector<std::string>& names) {
    std::vector<std::string> parts;
    for (const auto& nm : names) parts.push_back(nm + " = 0");
    return "  float " + join(parts, ", ") + ";";
  }

  std::vector<std::string> gen_vs_wrapper() {
    std::vector<IoMember> ins, outs;
    entry_io(ins, outs);
    std::vector<std::string> L = {"// entry wrapper: input register k <- k-th input semantic; out[0] <- SV_Position, out[1 + k] <- k-th other output",
                                  "SASL_FN void slv_jit_vs(const float4* in, const unsigned char* uniforms, float4* out, const SaslSampler& S0) {",
                                  "  const SaslUniforms& U = *reinterpret_cast<const SaslUniforms*>(uniforms);"};
    std::vector<std::string> args;
    int reg = 0;
    for (const auto& m : ins) {
      if (!m.sem.valid) throw compile_error("vertex-shader input " + m.name + " has no semantic");
      if ((m.type.kind != Kind::Scalar && m.type.kind != Kind::Vector) || m.type.base == Base::Bool)
        throw compile_error("vertex-shader input " + m.name + ": only float / int vectors are supported");
      refl.inputs.push_back({m.sem.name, (uint32_t)m.sem.index, to_string(m.type)});
      // integer inputs: the register holds the element's raw bits (get_vec4 of the *_sint / *_uint formats reinterprets them,
      // stream_assembler.cpp:26-45)
      for (int k = 0; k < m.type.n(); ++k) {
        const std::string lane = "in[" + std::to_string(reg) + "]." + std::string(1, "xyzw"[k]);
        args.push_back(m.type.base == Base::Float ? lane : (m.type.base == Base::Int ? "sasl_asint(" : "sasl_asuint(") + lane + ")");
      }
      ++reg;
    }
    if (reg > 8) throw compile_error("more than 8 vertex-shader inputs");
    std::vector<std::string> rets, stores;
    std::vector<std::vector<std::string>> full;
    std::vector<Semantic> sems;
    for (size_t k = 0; k < outs.size(); ++k) {
      const auto& m = outs[k];
      if (!m.sem.valid) throw compile_error("vertex-shader output " + m.name + " has no semantic");
      if ((m.type.kind != Kind::Scalar && m.type.kind != Kind::Vector) || m.type.base != Base::Float)
        throw compile_error("vertex-shader output " + m.name + ": only float vectors are supported");
      std::vector<std::string> names;
      for (int c = 0; c < m.type.n(); ++c) names.push_back("o" + std::to_string(k) + "_" + std::to_string(c));
      L.push_back(zeros_decl(names));
      rets.insert(rets.end(), names.begin(), names.end());
      names.resize(4, "0.0f");
      full.push_back(names);
      sems.push_back(m.sem);
    }
    // attribute registers follow the reference's semantic array, not the declaration order (reference_semantic_order)
    int attr = 0;
    bool have_pos = false;
    for (int k : reference_semantic_order(sems)) {
      const auto& m = outs[(size_t)k];
      if (is_position(m.sem) && !have_pos) {
        have_pos = true;  // a position narrower than float4 (the reference's semantic test units) is padded with zeros
        stores.push_back("  out[0] = make_float4(" + join(full[(size_t)k], ", ") + ");");
      } else {
        ++attr;
        refl.outputs.push_back({m.sem.name, (uint32_t)m.sem.index, to_string(m.type)});
        stores.push_back("  out[" + std::to_string(attr) + "] = make_float4(" + join(full[(size_t)k], ", ") + ");");
      }
    }
    if (!have_pos) throw compile_error("the vertex shader does not write SV_Position");
    if (attr > 5) throw compile_error("more than 5 vertex-shader output attributes (vs_output_ops, shader.cpp:45-52)");
    refl.n_vs_output_attrs = attr;
    std::vector<std::string> call = {"U", "S0"};
    call.insert(call.end(), args.begin(), args.end());
    call.insert(call.end(), rets.begin(), rets.end());
    L.push_back("  sasl_fn_" + entry->name + "(" + join(call, ", ") + ");");
    L.insert(L.end(), stores.begin(), stores.end());
    L.push_back("}");
    L.push_back("#define SLV_JIT_VS_OUTPUT_ATTRS " + std::to_string(attr));
    L.push_back("#define SLV_JIT_VS_SAMPLERS " + std::to_string(refl.samplers.size()));
    return L;
  }

  std::vector<std::string> gen_ps_wrapper() {
    std::vector<IoMember> flat, outs;
    entry_io(flat, outs);
    std::vector<std::string> L = {"// entry wrapper: k-th input <- interpolated attribute k; colour target 0 <- COLOR / SV_Target", "template <class Ctx>",
                                  "SASL_FN bool slv_jit_ps(const slv::RasterParams& p, const Ctx& px, float4& color) {",
                                  "  const SaslUniforms& U = *reinterpret_cast<const SaslUniforms*>(p.ps_uniforms);"};
    bool all_sem = true;
    std::vector<Semantic> sems;
    for (const auto& m : flat) {
      if ((m.type.kind != Kind::Scalar && m.type.kind != Kind::Vector) || m.type.base != Base::Float)
        throw compile_error("pixel-shader input " + m.name + ": only float vectors are supported");
      if (is_position(m.sem)) throw compile_error("reading SV_Position in a pixel shader is not supported");
      all_sem = all_sem && m.sem.valid;
      sems.push_back(m.sem);
    }
    if (flat.size() > 5) throw compile_error("more than 5 pixel-shader inputs");
    // input -> attribute: position k of the reference's semantic array (with a C++ vertex shader bound the reference hands
    // attribute k to the k-th entry, shader_unit.cpp:129); inputs without a semantic - which the reference rejects - keep the
    // declaration order
    std::vector<int> order;
    if (all_sem) order = reference_semantic_order(sems);
    else for (size_t k = 0; k < flat.size(); ++k) order.push_back((int)k);
    std::vector<int> attr_of(flat.size(), 0);
    for (size_t a = 0; a < order.size(); ++a) {
      attr_of[(size_t)order[a]] = (int)a;
      const auto& m = flat[(size_t)order[a]];
      L.push_back("  const float4 a" + std::to_string(a) + " = px.attr(" + std::to_string(a) + ");");
      refl.inputs.push_back({m.sem.valid ? m.sem.name : "TEXCOORD", (uint32_t)(m.sem.valid ? m.sem.index : (int)a), to_string(m.type)});
    }
    std::vector<std::string> args;
    for (size_t k = 0; k < flat.size(); ++k)
      for (int c = 0; c < flat[k].type.n(); ++c) args.push_back("a" + std::to_string(attr_of[k]) + "." + std::string(1, "xyzw"[c]));
    std::vector<std::string> rets, color;
    for (size_t k = 0; k < outs.size(); ++k) {
      const auto& m = outs[k];
      if ((m.type.kind != Kind::Scalar && m.type.kind != Kind::Vector) || m.type.base != Base::Float)
        throw compile_error("pixel-shader output " + m.name + ": only float vectors are supported");
      std::vector<std::string> names;
      for (int c = 0; c < m.type.n(); ++c) names.push_back("o" + std::to_string(k) + "_" + std::to_string(c));
      L.push_back(zeros_decl(names));
      rets.insert(rets.end(), names.begin(), names.end());
      refl.outputs.push_back({m.sem.valid ? m.sem.name : "COLOR", (uint32_t)(m.sem.valid ? m.sem.index : (int)k), to_string(m.type)});
      if (m.sem.valid && m.sem.name == "DEPTH")
        throw compile_error("pixel-shader depth output is not supported (framebuffer.cpp:348-353 ignores it for cpp shaders too)");
      if (color.empty() && (!m.sem.valid || m.sem.name == "COLOR" || m.sem.name == "SV_TARGET") && (!m.sem.valid || m.sem.index == 0)) {
        color = names;
        color.resize(4, "0.0f");
      }
    }
    if (color.empty()) color.assign(4, "0.0f");
    std::vector<std::string> call = {"U", "p", "px"};
    call.insert(call.end(), args.begin(), args.end());
    call.insert(call.end(), rets.begin(), rets.end());
    L.push_back("  sasl_fn_" + entry->name + "(" + join(call, ", ") + ");");
    L.push_back("  color = make_float4(" + join(color, ", ") + ");");
    L.push_back("  return true;");
    L.push_back("}");
    L.push_back("#define SLV_JIT_PS_SAMPLERS " + std::to_string(refl.samplers.size()));
    return L;
  }
};

}  // namespace detail
}  // namespace sasl
}  // namespace salvia_b200

#include "sasl_frontend_pp.inl"

// Preprocessor and public entry points of the SASL front end (included by sasl_frontend.hpp).
//
// The reference runs Boost.Wave in front of its parser (sasl/src/drivers/compiler_impl.cpp; the units
// sasl/test/repo/{preprocessors,include_main,include_header,include_search_path}.ss).  This is a line-oriented C preprocessor
// for what shaders use: `#define NAME [tokens]` (object-like and function-like macros), `#undef`, `#if` / `#ifdef` / `#ifndef` /
// `#elif` / `#else` / `#endif` with `defined(X)`, integer arithmetic, comparison and logical operators, `#include "file"`
// (directory of the including file, then the user paths) and `#include <file>` (system paths; both forms look into the
// virtual files first - the reference's add_virtual_file), `#error`, `#pragma` / `#line` (ignored).  Skipped and directive
// lines become EMPTY lines, so the line numbers the front end reports stay those of the top-level file.
// Mirrors salviarenderer_b200/sasl/preprocess.py.
#pragma once

#include <sys/stat.h>

namespace salvia_b200 {
namespace sasl {
namespace detail {

struct preprocess_error : std::runtime_error {
  using std::runtime_error::runtime_error;
};

class Preprocessor {
public:
  struct Macro { bool function_like = false; std::vector<std::string> params; std::string body; };
  std::map<std::string, Macro> macros;
  std::vector<std::string> include_dirs, sys_include_dirs;
  std::map<std::string, std::string> virtual_files;
  int max_depth = 32;

  explicit Preprocessor(const options& o) : include_dirs(o.include_dirs), sys_include_dirs(o.sys_include_dirs), virtual_files(o.virtual_files) {
    for (const auto& d : o.defines) { Macro m; m.body = d.second; macros[d.first] = m; }
  }

  static bool is_ident_start(char c) { return std::isalpha((unsigned char)c) || c == '_'; }
  static bool is_ident_char(char c) { return std::isalnum((unsigned char)c) || c == '_'; }
  static bool is_ident(const std::string& s) { return !s.empty() && is_ident_start(s[0]) && std::all_of(s.begin(), s.end(), is_ident_char); }
  static bool is_space(const std::string& s) { return !s.empty() && std::all_of(s.begin(), s.end(), [](unsigned char c) { return std::isspace(c); }); }
  static std::string strip(const std::string& s) {
    size_t a = 0, b = s.size();
    while (a < b && std::isspace((unsigned char)s[a])) ++a;
    while (b > a && std::isspace((unsigned char)s[b - 1])) --b;
    return s.substr(a, b - a);
  }

  // white space | // comment | /* comment */ | "string" | identifier | number-ish | any single character
  static std::vector<std::string> tokens(const std::string& s) {
    std::vector<std::string> out;
    size_t i = 0;
    const size_t n = s.size();
    while (i < n) {
      size_t e = i;
      const char c = s[i];
      if (std::isspace((unsigned char)c)) { while (e < n && std::isspace((unsigned char)s[e])) ++e; }
      else if (c == '/' && i + 1 < n && s[i + 1] == '/') { while (e < n && s[e] != '\n') ++e; }
      else if (c == '/' && i + 1 < n && s[i + 1] == '*' && s.find("*/", i + 2) != std::string::npos) { e = s.find("*/", i + 2) + 2; }
      else if (c == '"' && string_end(s, i) != std::string::npos) { e = string_end(s, i); }
      else if (is_ident_start(c)) { while (e < n && is_ident_char(s[e])) ++e; }
      else if (std::isdigit((unsigned char)c)) { ++e; while (e < n && (is_ident_char(s[e]) || s[e] == '.')) ++e; }
      else e = i + 1;
      out.push_back(s.substr(i, e - i));
      i = e;
    }
    return out;
  }
  // end (one past the closing quote) of the string literal that opens at s[i], npos when it is not closed
  static size_t string_end(const std::string& s, size_t i) {
    for (size_t k = i + 1; k < s.size(); ++k) {
      if (s[k] == '\\') { ++k; continue; }
      if (s[k] == '"') return k + 1;
    }
    return std::string::npos;
  }

  // ---- macro expansion
  std::string expand(const std::string& text, const std::set<std::string>& hide = {}) const {
    std::string out;
    const auto toks = tokens(text);
    size_t i = 0;
    while (i < toks.size()) {
      const std::string& t = toks[i];
      const Macro* m = nullptr;
      if (is_ident(t) && !hide.count(t)) { auto it = macros.find(t); if (it != macros.end()) m = &it->second; }
      if (!m) { out += t; ++i; continue; }
      std::set<std::string> hidden = hide;
      hidden.insert(t);
      if (!m->function_like) { out += expand(m->body, hidden); ++i; continue; }
      // function-like: needs '(' (possibly after white space)
      size_t j = i + 1;
      while (j < toks.size() && is_space(toks[j])) ++j;
      if (j >= toks.size() || toks[j] != "(") { out += t; ++i; continue; }
      int depth = 1;
      std::vector<std::string> args;
      std::string cur;
      bool cur_used = false;
      ++j;
      while (j < toks.size() && depth) {
        const std::string& c = toks[j];
        if (c == "(") ++depth;
        else if (c == ")") { if (--depth == 0) break; }
        if (c == "," && depth == 1) { args.push_back(strip(cur)); cur.clear(); cur_used = false; }
        else { cur += c; cur_used = true; }
        ++j;
      }
      if (depth) throw preprocess_error("unterminated argument list of macro " + t);
      if (cur_used || !args.empty()) args.push_back(strip(cur));
      if (args.size() != m->params.size())
        throw preprocess_error("macro " + t + " takes " + std::to_string(m->params.size()) + " argument(s), " + std::to_string(args.size()) + " given");
      for (auto& a : args) a = expand(a, hide);
      std::string body;
      for (const auto& b : tokens(m->body)) {
        auto p = std::find(m->params.begin(), m->params.end(), b);
        body += p != m->params.end() ? args[(size_t)(p - m->params.begin())] : b;
      }
      out += expand(body, hidden);
      i = j + 1;
    }
    return out;
  }

  // ---- #if expressions: integer constant expressions after `defined` and macro expansion (remaining identifiers are 0)
  struct ExprParser {
    std::vector<std::string> toks;
    size_t i = 0;
    const std::string& where;
    explicit ExprParser(const std::string& text, const std::string& where_) : where(where_) {
      for (const auto& t : tokens(text)) if (!is_space(t)) toks.push_back(t);
      // re-join two-character operators the tokenizer split
      std::vector<std::string> joined;
      for (size_t k = 0; k < toks.size(); ++k) {
        static const char* two[] = {"&&", "||", "==", "!=", "<=", ">=", "<<", ">>"};
        bool done = false;
        if (k + 1 < toks.size())
          for (const char* o : two) if (toks[k] + toks[k + 1] == o) { joined.push_back(o); ++k; done = true; break; }
        if (!done) joined.push_back(toks[k]);
      }
      toks = joined;
    }
    [[noreturn]] void fail() const { throw preprocess_error(where + ": cannot evaluate #if expression"); }
    bool peek(const char* s) const { return i < toks.size() && toks[i] == s; }
    bool accept(const char* s) { if (peek(s)) { ++i; return true; } return false; }
    long long primary() {
      if (accept("(")) { const long long v = ternary(); if (!accept(")")) fail(); return v; }
      if (i >= toks.size()) fail();
      const std::string t = toks[i++];
      if (is_ident(t)) return 0;
      if (!std::isdigit((unsigned char)t[0])) fail();
      const std::string digits = rstrip_set(t, "uUlL");
      char* end = nullptr;
      const long long v = std::strtoll(digits.c_str(), &end, 0);
      if (digits.empty() || *end) fail();
      return v;
    }
    long long unary() {
      if (accept("!")) return !unary();
      if (accept("-")) return -unary();
      if (accept("+")) return unary();
      if (accept("~")) return ~unary();
      return primary();
    }
    long long binary(int level) {
      static const std::vector<std::vector<std::string>> prec = {{"||"}, {"&&"}, {"|"}, {"^"}, {"&"}, {"==", "!="}, {"<", ">", "<=", ">="},
                                                                 {"<<", ">>"}, {"+", "-"}, {"*", "/", "%"}};
      if (level == (int)prec.size()) return unary();
      long long a = binary(level + 1);
      while (i < toks.size() && std::find(prec[level].begin(), prec[level].end(), toks[i]) != prec[level].end()) {
        const std::string op = toks[i++];
        const long long b = binary(level + 1);
        if (op == "||") a = a || b; else if (op == "&&") a = a && b; else if (op == "|") a = a | b; else if (op == "^") a = a ^ b;
        else if (op == "&") a = a & b; else if (op == "==") a = a == b; else if (op == "!=") a = a != b; else if (op == "<") a = a < b;
        else if (op == ">") a = a > b; else if (op == "<=") a = a <= b; else if (op == ">=") a = a >= b; else if (op == "<<") a = a << b;
        else if (op == ">>") a = a >> b; else if (op == "+") a = a + b; else if (op == "-") a = a - b; else if (op == "*") a = a * b;
        else { if (b == 0) fail(); a = op == "/" ? a / b : a % b; }
      }
      return a;
    }
    long long ternary() {
      const long long c = binary(0);
      if (accept("?")) { const long long a = ternary(); if (!accept(":")) fail(); const long long b = ternary(); return c ? a : b; }
      return c;
    }
    long long run() { const long long v = ternary(); if (i != toks.size()) fail(); return v; }
  };
  bool evaluate(const std::string& expr_in, const std::string& where) const {
    // defined(X) / defined X
    std::string expr;
    const auto toks = tokens(expr_in);
    for (size_t k = 0; k < toks.size(); ++k) {
      if (toks[k] != "defined") { expr += toks[k]; continue; }
      size_t j = k + 1;
      while (j < toks.size() && is_space(toks[j])) ++j;
      std::string name;
      size_t end = j;
      if (j < toks.size() && toks[j] == "(") {
        size_t a = j + 1;
        while (a < toks.size() && is_space(toks[a])) ++a;
        size_t b = a + 1;
        while (b < toks.size() && is_space(toks[b])) ++b;
        if (a < toks.size() && is_ident(toks[a]) && b < toks.size() && toks[b] == ")") { name = toks[a]; end = b; }
      } else if (j < toks.size() && j > k + 1 && is_ident(toks[j])) { name = toks[j]; end = j; }
      if (name.empty()) { expr += toks[k]; continue; }
      expr += macros.count(name) ? "1" : "0";
      k = end;
    }
    return ExprParser(expand(expr), where).run() != 0;
  }

  // ---- files
  static bool is_file(const std::string& p) { struct stat st; return ::stat(p.c_str(), &st) == 0 && S_ISREG(st.st_mode); }
  static std::string dir_of(const std::string& p) {
    char* real = ::realpath(p.c_str(), nullptr);
    std::string full = real ? real : p;
    if (real) std::free(real);
    const size_t s = full.rfind('/');
    return s == std::string::npos ? "." : (s == 0 ? "/" : full.substr(0, s));
  }
  bool find(const std::string& name, bool system, const std::string& cur_dir, std::string& path, std::string& text) const {
    auto v = virtual_files.find(name);
    if (v != virtual_files.end()) { path = name; text = v->second; return true; }
    std::vector<std::string> dirs;
    if (!system && !cur_dir.empty()) dirs.push_back(cur_dir);
    if (!system) dirs.insert(dirs.end(), include_dirs.begin(), include_dirs.end());
    dirs.insert(dirs.end(), sys_include_dirs.begin(), sys_include_dirs.end());
    if (system) dirs.insert(dirs.end(), include_dirs.begin(), include_dirs.end());
    for (const auto& d : dirs) {
      const std::string p = d.empty() ? name : (d.back() == '/' ? d + name : d + "/" + name);
      if (!is_file(p)) continue;
      std::ifstream f(p, std::ios::binary);
      std::stringstream ss;
      ss << f.rdbuf();
      path = p; text = ss.str();
      return true;
    }
    return false;
  }

  // Block comments may span lines: replace them by the same number of newlines (line comments are left to the lexer)
  static std::string strip_comments(const std::string& s) {
    std::string out;
    size_t i = 0;
    const size_t n = s.size();
    while (i < n) {
      if (s[i] == '/' && i + 1 < n && s[i + 1] == '*') {
        const size_t close = s.find("*/", i + 2);
        if (close != std::string::npos) {
          const auto nl = std::count(s.begin() + i, s.begin() + close + 2, '\n');
          out += nl ? std::string((size_t)nl, '\n') : std::string(" ");
          i = close + 2;
          continue;
        }
      }
      if (s[i] == '"') { const size_t e = string_end(s, i); if (e != std::string::npos) { out.append(s, i, e - i); i = e; continue; } }
      if (s[i] == '/' && i + 1 < n && s[i + 1] == '/') { size_t e = i; while (e < n && s[e] != '\n') ++e; out.append(s, i, e - i); i = e; continue; }
      out += s[i++];
    }
    return out;
  }
  static std::vector<std::string> split_lines(std::string s) {
    std::string t;
    for (size_t k = 0; k < s.size(); ++k) {
      if (s[k] == '\r') { t += '\n'; if (k + 1 < s.size() && s[k + 1] == '\n') ++k; }
      else t += s[k];
    }
    std::vector<std::string> lines;
    size_t a = 0;
    for (;;) {
      const size_t b = t.find('\n', a);
      if (b == std::string::npos) { lines.push_back(t.substr(a)); break; }
      lines.push_back(t.substr(a, b - a));
      a = b + 1;
    }
    return lines;
  }

  struct Frame { bool active, taken, parent; };
  std::string process(const std::string& src, const std::string& file_name, int depth = 0) {
    if (depth > max_depth) throw preprocess_error("#include nested too deeply");
    const std::string cur_dir = !file_name.empty() && is_file(file_name) ? dir_of(file_name) : std::string();
    const std::string where0 = file_name.empty() ? "<source>" : file_name;
    std::vector<std::string> out;
    std::vector<Frame> stack;
    bool active = true;
    const auto lines = split_lines(strip_comments(src));
    size_t i = 0;
    while (i < lines.size()) {
      std::string line = lines[i];
      size_t n_joined = 0;
      while (!line.empty() && line.back() == '\\' && i + 1 < lines.size()) {  // line continuation
        ++i;
        ++n_joined;
        line = line.substr(0, line.size() - 1) + lines[i];
      }
      const std::string where = where0 + ":" + std::to_string(i + 1 - n_joined);
      // \s*#\s*(\w*)\s*(.*)
      size_t k = 0;
      while (k < line.size() && std::isspace((unsigned char)line[k])) ++k;
      if (k >= line.size() || line[k] != '#') {
        out.push_back(active ? expand(line) : std::string());
        out.insert(out.end(), n_joined, std::string());
        ++i;
        continue;
      }
      ++k;
      while (k < line.size() && std::isspace((unsigned char)line[k])) ++k;
      size_t e = k;
      while (e < line.size() && is_ident_char(line[e])) ++e;
      const std::string cmd = line.substr(k, e - k);
      const std::string rest = strip(line.substr(e));
      std::string emitted;
      if (cmd == "ifdef" || cmd == "ifndef" || cmd == "if") {
        bool cond = false;
        if (active) {
          if (cmd == "if") cond = evaluate(rest, where);
          else {
            size_t m = 0;
            while (m < rest.size() && (m ? is_ident_char(rest[m]) : is_ident_start(rest[m]))) ++m;
            if (!m) throw preprocess_error(where + ": #" + cmd + " needs a name");
            cond = (macros.count(rest.substr(0, m)) != 0) == (cmd == "ifdef");
          }
        }
        stack.push_back({active, cond, active});
        active = active && cond;
      } else if (cmd == "elif" || cmd == "else") {
        if (stack.empty()) throw preprocess_error(where + ": #" + cmd + " without #if");
        const bool taken = stack.back().taken, parent = stack.back().parent;
        const bool cond = parent && !taken && (cmd == "else" ? true : evaluate(rest, where));
        stack.back() = {active, taken || cond, parent};
        active = cond;
      } else if (cmd == "endif") {
        if (stack.empty()) throw preprocess_error(where + ": #endif without #if");
        active = stack.back().parent;
        stack.pop_back();
      } else if (!active) {
      } else if (cmd == "define") {
        // ([A-Za-z_]\w*)(\(([^)]*)\))?\s*(.*)
        size_t m = 0;
        while (m < rest.size() && (m ? is_ident_char(rest[m]) : is_ident_start(rest[m]))) ++m;
        if (!m) throw preprocess_error(where + ": malformed #define");
        Macro mac;
        size_t body_at = m;
        if (m < rest.size() && rest[m] == '(') {
          const size_t close = rest.find(')', m);
          if (close != std::string::npos) {
            mac.function_like = true;
            std::string cur;
            for (size_t q = m + 1; q <= close; ++q) {
              if (q == close || rest[q] == ',') { const std::string prm = strip(cur); if (!prm.empty()) mac.params.push_back(prm); cur.clear(); }
              else cur += rest[q];
            }
            body_at = close + 1;
          }
        }
        mac.body = strip(rest.substr(body_at));
        macros[rest.substr(0, m)] = mac;
      } else if (cmd == "undef") {
        std::istringstream ss(rest);
        std::string name;
        ss >> name;
        macros.erase(name);
      } else if (cmd == "include") {
        const std::string spec = (!rest.empty() && (rest[0] == '"' || rest[0] == '<')) ? rest : expand(rest);
        std::string name;
        bool system = false;
        if (!spec.empty() && spec[0] == '"' && spec.find('"', 1) != std::string::npos && spec.find('"', 1) > 1) name = spec.substr(1, spec.find('"', 1) - 1);
        else if (!spec.empty() && spec[0] == '<' && spec.find('>', 1) != std::string::npos && spec.find('>', 1) > 1) { name = spec.substr(1, spec.find('>', 1) - 1); system = true; }
        else throw preprocess_error(where + ": malformed #include");
        std::string path, text;
        if (!find(name, system, cur_dir, path, text)) throw preprocess_error(where + ": cannot find include file " + quoted(name));
        // an included file contributes its text on ONE output line, so the line numbers of the including file survive
        const auto sub = split_lines(process(text, path, depth + 1));
        for (const auto& s : sub) if (!strip(s).empty()) { if (!emitted.empty()) emitted += " "; emitted += s; }
      } else if (cmd == "error") {
        throw preprocess_error(where + ": #error " + rest);
      } else if (cmd == "pragma" || cmd == "line" || cmd.empty()) {
      } else {
        throw preprocess_error(where + ": unknown directive #" + cmd);
      }
      out.push_back(emitted);
      out.insert(out.end(), n_joined, std::string());
      ++i;
    }
    if (!stack.empty()) throw preprocess_error(where0 + ": unterminated #if");
    return join(out, "\n");
  }
};

}  // namespace detail

// compile(code, profile) of the reference (salvia/include/salvia/core/renderer.h:136-147): stage "vs" | "ps" | "lib" (functions
// only, no entry point: the reference's *.ss test units); `entry` empty = the function that carries semantics.
// Returns false and fills `error` when the source does not compile.
inline bool compile(const std::string& source_in, const std::string& stage, const std::string& entry, const options& opt, unit& out, std::string& error) {
  try {
    if (stage != "vs" && stage != "ps" && stage != "lib") throw compile_error("stage must be 'vs', 'ps' or 'lib'");
    std::string source = source_in;
    if (source.find('#') != std::string::npos || !opt.defines.empty()) {
      try {
        source = detail::Preprocessor(opt).process(source, opt.file_name);
      } catch (const detail::preprocess_error& e) {
        throw compile_error(std::string("preprocessor: ") + e.what());
      }
    }
    detail::Gen g(source, stage, entry);
    out = g.run();
    if (out.samplers.size() > (stage == "ps" ? 2u : 1u))  // RasterParams.sampler0 / sampler1, GeomParams.sampler0
      throw compile_error("at most two samplers per pixel shader and one per vertex shader are supported");
    return true;
  } catch (const std::exception& e) {
    error = e.what();
    return false;
  }
}
inline bool compile(const std::string& source, const std::string& stage, unit& out, std::string& error) { return compile(source, stage, "", options(), out, error); }

// The line-oriented text form of a unit (`SLVSASL 1 ... code NBYTES` + the code) that salviarenderer_b200/sasl/emit.py writes
inline std::string render(const unit& u) {
  std::ostringstream o;
  o << "SLVSASL 1\n" << "stage " << u.stage << "\n" << "n_vs_output_attrs " << u.n_vs_output_attrs << "\n" << "uniform_bytes " << u.uniform_bytes << "\n"
    << "uses_derivatives " << (u.uses_derivatives ? 1 : 0) << "\n";
  for (const auto& x : u.uniforms) o << "uniform " << x.name << " " << x.type << " " << x.offset << " " << x.size << "\n";
  for (size_t k = 0; k < u.samplers.size(); ++k) o << "sampler " << k << " " << u.samplers[k] << "\n";
  for (size_t k = 0; k < u.inputs.size(); ++k) o << "input " << u.inputs[k].semantic << " " << u.inputs[k].index << " " << k << "\n";
  for (size_t k = 0; k < u.outputs.size(); ++k) o << "output " << u.outputs[k].semantic << " " << u.outputs[k].index << " " << k << "\n";
  o << "code " << u.code.size() << "\n" << u.code;
  return o.str();
}

}  // namespace sasl
}  // namespace salvia_b200

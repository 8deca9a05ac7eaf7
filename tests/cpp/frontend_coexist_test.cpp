// An executable that includes the header-only SASL front end AND loads a library that carries its own copy (slv_sasl_translate):
// the configuration of every C++ host that uses shader::compile() next to libsalvia_b200.so.  Both copies are called alternately
// and must agree (tests/test_sasl_frontend_cpp.py).      usage: frontend_coexist_test <library.so>
#include <dlfcn.h>
#include <cstdio>
#include "sasl_frontend.hpp"
#include "salvia_b200.h"
int main(int argc, char** argv) {
  void* lib = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
  if (!lib) { std::printf("dlopen: %s\n", dlerror()); return 1; }
  auto tr = reinterpret_cast<decltype(&slv_sasl_translate)>(dlsym(lib, "slv_sasl_translate"));
  auto fr = reinterpret_cast<decltype(&slv_free)>(dlsym(lib, "slv_free"));
  const char* src = "float4 main(float4 p: TEXCOORD0): COLOR { return p * 2 + sin(p); }";
  for (int k = 0; k < 3; ++k) {
    salvia_b200::sasl::unit u; std::string err;
    if (!salvia_b200::sasl::compile(src, "ps", u, err)) { std::printf("own: %s\n", err.c_str()); return 2; }
    const std::string own = salvia_b200::sasl::render(u);
    char* unit = nullptr; size_t n = 0; char log[512];
    if (tr(1, src, nullptr, &unit, &n, log, sizeof log) != 0) { std::printf("lib: %s\n", log); return 3; }
    if (own != std::string(unit, n)) { std::printf("differ\n"); return 4; }
    fr(unit);
    if (tr(1, "float4 broken(", nullptr, &unit, &n, log, sizeof log) == 0) return 5;
    if (salvia_b200::sasl::compile("float4 broken(", "ps", u, err)) return 6;
  }
  std::printf("ok\n");
  return 0;
}

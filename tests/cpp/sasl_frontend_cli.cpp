// Drives the C++ SASL front end (salviarenderer_b200/host/sasl_frontend.hpp) from the command line, for the comparison with the
// Python front end (tests/test_sasl_frontend_cpp.py): source on stdin -> the unit text (or "error\n<message>") on stdout.
//   sasl_frontend_cli vs|ps|lib|pp [--entry NAME] [-D NAME[=VALUE]]... [-I DIR]... [--sys DIR]... [--file NAME] [--virtual NAME PATH]...
#include <cstdio>
#include <iostream>
#include <iterator>

#include "sasl_frontend.hpp"

int main(int argc, char** argv) {
  namespace sasl = salvia_b200::sasl;
  if (argc < 2) return 64;
  const std::string stage = argv[1];
  std::string entry;
  sasl::options opt;
  for (int i = 2; i < argc; ++i) {
    const std::string a = argv[i];
    if (a == "--entry" && i + 1 < argc) entry = argv[++i];
    else if (a == "-D" && i + 1 < argc) {
      const std::string d = argv[++i];
      const size_t eq = d.find('=');
      opt.defines.push_back({d.substr(0, eq), eq == std::string::npos ? "" : d.substr(eq + 1)});
    } else if (a == "-I" && i + 1 < argc) opt.include_dirs.push_back(argv[++i]);
    else if (a == "--sys" && i + 1 < argc) opt.sys_include_dirs.push_back(argv[++i]);
    else if (a == "--file" && i + 1 < argc) opt.file_name = argv[++i];
    else if (a == "--virtual" && i + 2 < argc) {
      std::ifstream f(argv[i + 2], std::ios::binary);
      opt.virtual_files[argv[i + 1]] = std::string(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
      i += 2;
    } else return 64;
  }
  const std::string source((std::istreambuf_iterator<char>(std::cin)), std::istreambuf_iterator<char>());
  if (stage == "pp") {  // the preprocessor alone
    try {
      std::cout << sasl::detail::Preprocessor(opt).process(source, opt.file_name);
    } catch (const std::exception& e) {
      std::cout << "error\n" << e.what() << "\n";
      return 2;
    }
    return 0;
  }
  sasl::unit u;
  std::string error;
  if (!sasl::compile(source, stage, entry, opt, u, error)) {
    std::cout << "error\n" << error << "\n";
    return 2;
  }
  std::cout << sasl::render(u);
  return 0;
}

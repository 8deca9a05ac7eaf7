// shader::compile() of the C++ host surface (salviarenderer_b200/host/salvia_b200_renderer.hpp) without a device: the SASL front
// end runs in process (or, with SLV_SASL_FRONTEND=python, as a child process) and its unit (reflection + generated device code) is
// parsed into a shader_object.  Prints
// the reflection in a stable text form; the Python test compares it with what the front end reports in process.
//   usage: sasl_compile_test vs|ps < shader.sasl
#include <cstdio>
#include <iostream>
#include <iterator>
#include <string>

#include "salvia_b200_renderer.hpp"

using namespace salvia_b200;

int main(int argc, char** argv) {
  if (argc < 2) return 1;
  std::string src((std::istreambuf_iterator<char>(std::cin)), std::istreambuf_iterator<char>());
  shader::shader_profile prof;
  prof.language = std::string(argv[1]) == "vs" ? shader::lang_vertex_shader : shader::lang_pixel_shader;
  shader::shader_log_ptr log;
  shader::shader_object_ptr so = shader::compile(src, prof, log);
  if (!so) {
    std::printf("error\n%s\n", log ? log->c_str() : "");
    return 2;
  }
  std::printf("n_vs_output_attrs %u\nuniform_bytes %zu\nuses_derivatives %d\n", so->n_vs_output_attrs, so->uniform_bytes, so->uses_derivatives ? 1 : 0);
  for (auto const& u : so->uniforms) std::printf("uniform %s %s %zu %zu\n", u.first.c_str(), u.second.type.c_str(), u.second.offset, u.second.size);
  for (size_t i = 0; i < so->samplers.size(); ++i) std::printf("sampler %zu %s\n", i, so->samplers[i].c_str());
  for (auto const& x : so->inputs) std::printf("input %s %u %u\n", x.semantic.c_str(), x.index, x.slot);
  for (auto const& x : so->outputs) std::printf("output %s %u %u\n", x.semantic.c_str(), x.index, x.slot);
  std::printf("code_bytes %zu\n", so->device_code.size());
  // a shader_profile the surface does not compile
  shader::shader_profile bad;
  bad.language = shader::lang_blending_shader;
  if (shader::compile(src, bad, log)) return 3;
  return 0;
}

// slv_sasl_translate: the SASL front end (salviarenderer_b200/host/sasl_frontend.hpp) behind the C ABI - the first half of the
// reference's compile(code, profile) (salvia/include/salvia/core/renderer.h:136-147); slv_shader_compile is the second half.
// Host code only; compiled by the host compiler and linked into libsalvia_b200.so.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../host/sasl_frontend.hpp"
#include "salvia_b200.h"

extern "C" slv_result slv_sasl_translate(uint32_t stage, const char* source, const char* entry, char** unit, size_t* unit_bytes, char* log,
                                         size_t log_bytes) {
  if (log && log_bytes) log[0] = 0;
  if (!source || !unit || (stage != SLV_STAGE_VS && stage != SLV_STAGE_PS)) return SLV_INVALID_PARAMETER;
  *unit = nullptr;
  if (unit_bytes) *unit_bytes = 0;
  namespace sasl = salvia_b200::sasl;
  sasl::unit u;
  std::string error;
  if (!sasl::compile(source, stage == SLV_STAGE_VS ? "vs" : "ps", entry ? entry : "", sasl::options(), u, error)) {
    if (log && log_bytes) std::snprintf(log, log_bytes, "%s", error.c_str());
    return SLV_FAILED;
  }
  const std::string text = sasl::render(u);
  char* out = static_cast<char*>(std::malloc(text.size() + 1));
  if (!out) return SLV_OUT_OF_MEMORY;
  std::memcpy(out, text.data(), text.size());
  out[text.size()] = 0;
  *unit = out;
  if (unit_bytes) *unit_bytes = text.size();
  return SLV_OK;
}

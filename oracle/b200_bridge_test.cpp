// oracle/b200_bridge_test.cpp — drives scenes through the REFERENCE's own renderer interface (salvia::core::renderer,
// renderer.h:42-131), once into the reference's sync_renderer (create_benchmark_renderer) and once into b200_renderer
// (oracle/b200_renderer.hpp: the renderer_impl subclass of INTEGRATION.md §2, compiled against the reference's headers) bound to
// a C-ABI library given on the command line.  Buffers and counters of the two runs must be equal.
//
// TEST INFRASTRUCTURE.  Built by `make -C oracle bridge` into oracle/_ref/ (needs /root/reference; the binary travels to the GPU
// box like libsalvia_ref.so).     usage: b200_bridge_test <library.so> [width height samples]
#include <salvia/core/renderer.h>
#include <salvia/core/shader.h>
#include <salvia/resource/mapped_resource.h>
#include <salvia/resource/pixel_accessor.h>
#include <salvia/resource/sampler.h>
#include <salvia/resource/texture.h>
#include <salvia/shader/shader_regs.h>

#include <eflib/math/math.h>

#include <cinttypes>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <type_traits>
#include <vector>

#include "b200_renderer.hpp"

using namespace salvia;
using namespace salvia::core;
using namespace salvia::resource;
using namespace salvia::shader;
using eflib::mat44;
using eflib::vec3;
using eflib::vec4;

#define BRIDGE_CLONE()                                           \
  cpp_shader_ptr clone() override {                              \
    typedef std::remove_pointer<decltype(this)>::type this_type; \
    return cpp_shader_ptr(new this_type(*this));                 \
  }

namespace {

// ---- the samples' cpp shaders, each advertising its device twin --------------------------------------------------------------
// samples/ColorizedTriangle/ColorizedTriangle.cpp:29-53 (the SASL vertex shader's cpp form)
struct vs_lights3 : cpp_vertex_shader, device_shader_info {
  mat44 wvp;
  vec4 light_pos[3];
  vs_lights3() {
    declare_constant("wvpMatrix", wvp);
    declare_constant("lightPos0", light_pos[0]);
    declare_constant("lightPos1", light_pos[1]);
    declare_constant("lightPos2", light_pos[2]);
    bind_semantic("POSITION", 0, 0);
    bind_semantic("NORMAL", 0, 1);
  }
  void shader_prog(const vs_input& in, vs_output& out) override {
    vec4 pos = in.attribute(0);
    eflib::transform(out.position(), pos, wvp);
    out.attribute(0) = in.attribute(1);
    out.attribute(1) = light_pos[0] - pos;
    out.attribute(2) = light_pos[1] - pos;
    out.attribute(3) = light_pos[2] - pos;
  }
  uint32_t num_output_attributes() const override { return 4; }
  uint32_t output_attribute_modifiers(uint32_t) const override { return vs_output::am_linear; }
  uint32_t device_program() const override { return SLV_VS_LIGHTS3; }
  size_t pack_uniforms(uint8_t* dst, size_t) const override {
    slv_vs_lights3_uniforms u{};
    std::memcpy(u.wvp, &wvp, 64);
    for (int k = 0; k < 3; ++k) std::memcpy(u.light_pos[k], &light_pos[k], 16);
    std::memcpy(dst, &u, sizeof(u));
    return sizeof(u);
  }
  BRIDGE_CLONE()
};
// samples/ColorizedTriangle/ColorizedTriangle.cpp:55-92
struct ps_lights3 : cpp_pixel_shader, device_shader_info {
  bool shader_prog(const vs_output& in, ps_output& out) override {
    vec3 l0 = in.attribute(1).xyz(), l1 = in.attribute(2).xyz(), l2 = in.attribute(3).xyz();
    vec3 norm = in.attribute(0).xyz();
    float i0 = 1.0f / l0.length(), i1 = 1.0f / l1.length(), i2 = 1.0f / l2.length();
    vec3 n = eflib::normalize3(norm);
    vec3 n0 = l0 * i0, n1 = l1 * i1, n2 = l2 * i2;
    float r0 = eflib::dot_prod3(n, n0), r1 = eflib::dot_prod3(n, n1), r2 = eflib::dot_prod3(n, n2);
    out.color[0] = eflib::clampss(vec4(0.7f, 0.1f, 0.3f, 1.0f) * r0 * i0 * i0 + vec4(0.1f, 0.3f, 0.7f, 1.0f) * r1 * i1 * i1 +
                                      vec4(0.3f, 0.7f, 0.1f, 1.0f) * r2 * i2 * i2,
                                  0.0f, 1.0f);
    out.color[0][3] = 1.0f;
    return true;
  }
  uint32_t device_program() const override { return SLV_PS_LIGHTS3; }
  size_t pack_uniforms(uint8_t*, size_t) const override { return 0; }
  BRIDGE_CLONE()
};
// samples/Sponza/Sponza.cpp:64-97
struct vs_sponza : cpp_vertex_shader, device_shader_info {
  mat44 wvp;
  vec4 light_pos, eye_pos;
  vs_sponza() {
    declare_constant("wvpMatrix", wvp);
    declare_constant("lightPos", light_pos);
    declare_constant("eyePos", eye_pos);
    bind_semantic("POSITION", 0, 0);
    bind_semantic("TEXCOORD", 0, 1);
    bind_semantic("NORMAL", 0, 2);
  }
  void shader_prog(const vs_input& in, vs_output& out) override {
    vec4 pos = in.attribute(0);
    eflib::transform(out.position(), pos, wvp);
    out.attribute(0) = in.attribute(1);
    out.attribute(1) = in.attribute(2);
    out.attribute(2) = light_pos - pos;
    out.attribute(3) = eye_pos - pos;
  }
  uint32_t num_output_attributes() const override { return 4; }
  uint32_t output_attribute_modifiers(uint32_t) const override { return vs_output::am_linear; }
  uint32_t device_program() const override { return SLV_VS_SPONZA; }
  size_t pack_uniforms(uint8_t* dst, size_t) const override {
    slv_vs_sponza_uniforms u{};
    std::memcpy(u.wvp, &wvp, 64);
    std::memcpy(u.light_pos, &light_pos, 16);
    std::memcpy(u.eye_pos, &eye_pos, 16);
    std::memcpy(dst, &u, sizeof(u));
    return sizeof(u);
  }
  BRIDGE_CLONE()
};
// samples/Sponza/Sponza.cpp:99-141
struct ps_sponza : cpp_pixel_shader, device_shader_info {
  sampler_ptr sampler_;
  ps_sponza() { declare_sampler("Sampler", sampler_); }
  bool shader_prog(const vs_output& in, ps_output& out) override {
    vec4 diff = vec4(1.0f, 1.0f, 1.0f, 1.0f);
    if (sampler_) diff = tex2d(*sampler_, 0).get_vec4();
    vec3 norm(eflib::normalize3(in.attribute(1).xyz()));
    vec3 light_dir(eflib::normalize3(in.attribute(2).xyz()));
    float illum = eflib::clamp(eflib::dot_prod3(light_dir, norm), 0.0f, 1.0f);
    out.color[0] = diff * illum;
    out.color[0][3] = 1.0f;
    return true;
  }
  uint32_t device_program() const override { return SLV_PS_SPONZA; }
  size_t pack_uniforms(uint8_t* dst, size_t) const override {
    slv_ps_sponza_uniforms u{sampler_ ? 1u : 0u};
    std::memcpy(dst, &u, sizeof(u));
    return sizeof(u);
  }
  void device_samplers(sampler_ptr (&out)[SLV_MAX_SAMPLERS]) const override { out[0] = sampler_; }
  BRIDGE_CLONE()
};
// ps_sponza with the diffuse texture fetched the way a SASL pixel shader's tex2D fetches it: sampler::sample_2d_grad with the
// quad's per-line / per-column differences (sasl/src/codegen/cg_impl.cpp:902-909, cgs_simd.cpp:275-313) - the cpp twin of the
// SASL pixel shader below.  cpp_pixel_shader keeps the quad private, but execute() (cpp_pixel_shader.cpp:63-75) calls
// shader_prog for pixels 0..3 of the quad in order with in == quad[i], and every worker owns its clone: the call count modulo
// 4 is the pixel's index.
struct ps_sponza_grad : cpp_pixel_shader, device_shader_info {
  sampler_ptr sampler_;
  unsigned calls = 0;
  ps_sponza_grad() { declare_sampler("Sampler", sampler_); }
  bool shader_prog(const vs_output& in, ps_output& out) override {
    unsigned const i = calls++ & 3u;
    vs_output const* quad = &in - i;
    vec4 diff = vec4(1.0f, 1.0f, 1.0f, 1.0f);
    if (sampler_) {
      vec4 dx = quad[i | 1u].attribute(0) - quad[i & ~1u].attribute(0);
      vec4 dy = quad[i | 2u].attribute(0) - quad[i & ~2u].attribute(0);
      diff = sampler_->sample_2d_grad(in.attribute(0).xy(), dx.xy(), dy.xy(), 0.0f).get_vec4();
    }
    vec3 norm(eflib::normalize3(in.attribute(1).xyz()));
    vec3 light_dir(eflib::normalize3(in.attribute(2).xyz()));
    float illum = eflib::clamp(eflib::dot_prod3(light_dir, norm), 0.0f, 1.0f);
    out.color[0] = diff * illum;
    out.color[0][3] = 1.0f;
    return true;
  }
  uint32_t device_program() const override { return SLV_PS_SPONZA_GRAD; }
  size_t pack_uniforms(uint8_t* dst, size_t) const override {
    slv_ps_sponza_grad_uniforms u{sampler_ ? 1u : 0u, 1u};
    std::memcpy(dst, &u, sizeof(u));
    return sizeof(u);
  }
  void device_samplers(sampler_ptr (&out)[SLV_MAX_SAMPLERS]) const override { out[0] = sampler_; }
  BRIDGE_CLONE()
};
// the SASL pair BASELINE configs[3] names (samples/Sponza/Sponza.cpp:39-62 and its pixel shader in SASL)
const char* kSaslVs =
    "float4x4 wvpMatrix; float4 lightPos; float4 eyePos;\n"
    "struct VSIn  { float4 pos: POSITION; float4 tex: TEXCOORD0; float4 norm: NORMAL; };\n"
    "struct VSOut { float4 pos: sv_position; float4 tex: TEXCOORD0; float4 norm: TEXCOORD1; float4 lightDir: TEXCOORD2; float4 eyeDir: TEXCOORD3; };\n"
    "VSOut vs_main(VSIn in) {\n"
    "  VSOut o;\n"
    "  o.norm = in.norm; o.pos = mul(in.pos, wvpMatrix); o.lightDir = lightPos - in.pos; o.eyeDir = eyePos - in.pos; o.tex = in.tex;\n"
    "  return o;\n"
    "}\n";
const char* kSaslPs =
    "sampler texSamp;\n"
    "struct PSIn { float4 tex: TEXCOORD0; float4 norm: TEXCOORD1; float4 lightDir: TEXCOORD2; float4 eyeDir: TEXCOORD3; };\n"
    "float4 ps_main(PSIn in): COLOR {\n"
    "  float4 diff = tex2D(texSamp, in.tex.xy);\n"
    "  float illum = clamp(dot(normalize(in.lightDir.xyz), normalize(in.norm.xyz)), 0.0f, 1.0f);\n"
    "  return float4(diff.xyz * illum, 1.0f);\n"
    "}\n";
struct bs_replace : cpp_blend_shader, device_shader_info {  // ColorizedTriangle.cpp:94-106
  bool shader_prog(size_t sample, pixel_accessor& inout, const ps_output& in) override {
    inout.color(0, sample, color_rgba32f(in.color[0]));
    return true;
  }
  uint32_t device_program() const override { return SLV_BS_REPLACE; }
  size_t pack_uniforms(uint8_t*, size_t) const override { return 0; }
  BRIDGE_CLONE()
};

uint64_t fnv(const void* p, size_t n) {
  const uint8_t* b = static_cast<const uint8_t*>(p);
  uint64_t h = 1469598103934665603ull;
  for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
  return h;
}

struct frame_hashes {
  uint64_t color = 0, depth = 0, resolved = 0;
  size_t drawn = 0;  // samples that are not the clear colour
  pipeline_statistics stats{};
  bool ok = false;
};
#define CHECK(e) do { if ((e) != result::ok) { std::fprintf(stderr, "FAILED: %s (line %d)\n", #e, __LINE__); return out; } } while (0)

template <class T>
result fill_buffer(renderer& r, buffer_ptr const& b, std::vector<T> const& v) {
  mapped_resource m;
  if (r.map(m, b, map_write) != result::ok) return result::failed;
  std::memcpy(m.data, v.data(), v.size() * sizeof(T));
  r.unmap();  // upstream quirk: resource_manager::map_impl never records the map mode, so unmap() always reports `failed`
  return result::ok;  // (resource_manager.cpp:8-54,72-75); the samples ignore it as well
}

// One scene, written against salvia::core::renderer only.  `resolve` is the one call that is not a renderer method upstream
// (surface::resolve, surface.cpp:123-140): the caller passes how its renderer resolves.
// `sasl`: the renderer that compiles SASL (a b200_renderer bound to a library that can), or null: pass 3 then runs the cpp twins
template <class Resolve>
frame_hashes run_scene(renderer& r, size_t W, size_t H, size_t S, Resolve resolve, bool pass3, b200_renderer* sasl = nullptr) {
  frame_hashes out;
  texture_ptr color = r.create_tex2d(W, H, S, pixel_format_color_rgba8), ds = r.create_tex2d(W, H, S, pixel_format_color_rg32f);
  texture_ptr resolved = r.create_tex2d(W, H, 1, pixel_format_color_rgba8);
  surface_ptr cs = color->subresource(0), dss = ds->subresource(0), rs = resolved->subresource(0);
  CHECK(r.set_render_targets(1, &cs, dss));
  viewport vp; vp.x = 0; vp.y = 0; vp.w = (float)W; vp.h = (float)H; vp.minz = 0.0f; vp.maxz = 1.0f;
  CHECK(r.set_viewport(vp));
  raster_desc rd; rd.cm = cull_none;
  CHECK(r.set_rasterizer_state(raster_state_ptr(new raster_state(rd))));

  // a bumpy G x G grid: positions, uv, normals in three streams; u16 indices
  const uint32_t G = 12;
  std::vector<float> gp, guv, gn;
  std::vector<uint16_t> gi;
  for (uint32_t i = 0; i <= G; ++i)
    for (uint32_t j = 0; j <= G; ++j) {
      const float hgt = 0.3f * (float)((i * 5 + j * 3) % 4);
      gp.insert(gp.end(), {-3.0f + 0.5f * (float)i, -1.0f + hgt, -3.0f + 0.5f * (float)j, 1.0f});
      guv.insert(guv.end(), {3.0f * (float)i / (float)G, 3.0f * (float)j / (float)G, 0.0f, 0.0f});
      gn.insert(gn.end(), {0.1f * (float)((i + j) % 3), 1.0f, 0.1f * (float)(j % 2), 0.0f});
    }
  for (uint32_t i = 0; i < G; ++i)
    for (uint32_t j = 0; j < G; ++j) {
      const uint16_t q0 = (uint16_t)(i * (G + 1) + j), q2 = (uint16_t)(q0 + G + 2);
      gi.insert(gi.end(), {q0, (uint16_t)(q0 + 1), q2, q2, (uint16_t)(q2 - 1), q0});
    }
  buffer_ptr b0 = r.create_buffer(gp.size() * 4), b1 = r.create_buffer(guv.size() * 4), b2 = r.create_buffer(gn.size() * 4), bi = r.create_buffer(gi.size() * 2);
  CHECK(fill_buffer(r, b0, gp));
  CHECK(fill_buffer(r, b1, guv));
  CHECK(fill_buffer(r, b2, gn));
  CHECK(fill_buffer(r, bi, gi));
  buffer_ptr bufs[3] = {b0, b1, b2};
  size_t strides[3] = {16, 16, 16}, offsets[3] = {0, 0, 0};
  CHECK(r.set_vertex_buffers(0, 3, bufs, strides, offsets));
  CHECK(r.set_index_buffer(bi, format_r16_uint));
  CHECK(r.set_primitive_topology(primitive_triangle_list));
  input_element_desc descs[] = {{"POSITION", 0, format_r32g32b32a32_float, 0, 0, input_per_vertex, 0},
                                {"TEXCOORD", 0, format_r32g32b32a32_float, 1, 0, input_per_vertex, 0},
                                {"NORMAL", 0, format_r32g32b32a32_float, 2, 0, input_per_vertex, 0}};

  mat44 view, proj, wvp;
  eflib::mat_lookat(view, vec3(0.0f, 2.5f, -5.0f), vec3(0.0f, 0.0f, 0.0f), vec3(0.0f, 1.0f, 0.0f));
  eflib::mat_perspective_fov(proj, 1.0f, (float)W / (float)H, 0.1f, 100.0f);
  eflib::mat_mul(wvp, view, proj);

  async_object_ptr query = r.create_query(async_object_ids::pipeline_statistics);
  CHECK(r.begin(query));
  color_rgba32f clear_c(0.2f, 0.2f, 0.5f, 1.0f);
  CHECK(r.clear_color(cs, clear_c));
  CHECK(r.clear_depth_stencil(dss, clear_depth | clear_stencil, 1.0f, 0));

  // pass 1: the lit mesh (ColorizedTriangle's shaders), indexed
  {
    auto vs = std::make_shared<vs_lights3>();
    CHECK(r.set_vertex_shader(vs));
    CHECK(r.set_pixel_shader(std::make_shared<ps_lights3>()));
    CHECK(r.set_blend_shader(std::make_shared<bs_replace>()));
    CHECK(r.set_input_layout(r.create_input_layout(descs, 3, vs)));
    vec4 l0(2.0f, 3.0f, -2.0f, 1.0f), l1(-3.0f, 1.5f, 1.0f, 1.0f), l2(0.5f, 2.0f, 3.0f, 1.0f);
    // a cpp shader's constants are set on the shader object (Sponza.cpp:244-246); renderer::set_vs_variable feeds SASL cbuffers
    CHECK(vs->set_constant("wvpMatrix", &wvp));
    CHECK(vs->set_constant("lightPos0", &l0));
    CHECK(vs->set_constant("lightPos1", &l1));
    CHECK(vs->set_constant("lightPos2", &l2));
    CHECK(r.draw_index(0, G * G * 2, 0));
  }
  // pass 2: a textured, trilinear-filtered part of the mesh on top (Sponza's shaders), drawn with a start index and less-equal
  {
    const uint32_t TS = 32;
    texture_ptr tex = r.create_tex2d(TS, TS, 1, pixel_format_color_rgba8);
    {
      mapped_resource m;
      CHECK(r.map(m, tex->subresource(0), map_write));
      uint8_t* t = static_cast<uint8_t*>(m.data);
      for (uint32_t y = 0; y < TS; ++y)
        for (uint32_t x = 0; x < TS; ++x) {
          uint8_t* px = t + (y * TS + x) * 4;
          px[0] = (uint8_t)(x * 8); px[1] = (uint8_t)(y * 8); px[2] = (uint8_t)(((x ^ y) & 4) ? 230 : 40); px[3] = 255;
        }
      r.unmap();
    }
    tex->gen_mipmap(filter_linear, true);
    sampler_desc sd;
    sd.min_filter = sd.mag_filter = sd.mip_filter = filter_linear;
    sd.addr_mode_u = sd.addr_mode_v = sd.addr_mode_w = address_wrap;
    sampler_ptr samp = r.create_sampler(sd, tex);
    auto vs = std::make_shared<vs_sponza>();
    auto ps = std::make_shared<ps_sponza>();
    CHECK(r.set_vertex_shader(vs));
    CHECK(r.set_pixel_shader(ps));
    CHECK(r.set_input_layout(r.create_input_layout(descs, 3, vs)));
    vec4 light(2.0f, 4.0f, -1.0f, 1.0f), eye(0.0f, 2.5f, -5.0f, 1.0f);
    CHECK(vs->set_constant("wvpMatrix", &wvp));
    CHECK(vs->set_constant("lightPos", &light));
    CHECK(vs->set_constant("eyePos", &eye));
    CHECK(ps->set_sampler("Sampler", samp));
    depth_stencil_desc dsd;
    dsd.depth_func = compare_function_less_equal;
    CHECK(r.set_depth_stencil_state(depth_stencil_state_ptr(new depth_stencil_state(dsd)), 0));
    CHECK(r.draw_index(G * 2 * 3 * 3, G * 2 * 5, 0));  // rows 3..7 of the grid

    // pass 3: rows 8..11 with the SASL pair and a 16x anisotropic sampler - through compile() / set_vertex_shader_code /
    // set_pixel_shader_code / set_vs_variable_value / set_ps_sampler of the REFERENCE's renderer interface (renderer.h:75-86,
    // 136-147) where the renderer can compile SASL, else (the reference here: no LLVM) with the pair's cpp twins
    if (pass3) {
    sampler_desc sa = sd;
    sa.mip_filter = filter_anisotropic;
    sa.max_anisotropy = 16;
    sampler_ptr samp16 = r.create_sampler(sa, tex);
    if (sasl) {
      std::string log;
      shader_object_ptr vso = sasl->compile(kSaslVs, lang_vertex_shader, &log), pso = sasl->compile(kSaslPs, lang_pixel_shader, &log);
      if (!vso || !pso) { std::fprintf(stderr, "SASL: %s\n", log.c_str()); return out; }
      if (sasl->compile("float4 broken(", lang_pixel_shader, &log) || log.find("line 1") == std::string::npos) return out;
      CHECK(r.set_vertex_shader_code(vso));
      CHECK(r.set_pixel_shader_code(pso));
      CHECK(r.set_input_layout(r.create_input_layout(descs, 3, vso)));
      CHECK(r.set_vs_variable_value("wvpMatrix", &wvp, sizeof(wvp)));
      CHECK(r.set_vs_variable_value("lightPos", &light, sizeof(light)));
      CHECK(r.set_vs_variable_value("eyePos", &eye, sizeof(eye)));
      CHECK(r.set_ps_sampler("texSamp", samp16));
    } else {
      auto ps3 = std::make_shared<ps_sponza_grad>();
      CHECK(r.set_pixel_shader(ps3));
      CHECK(ps3->set_sampler("Sampler", samp16));
    }
    CHECK(r.draw_index(G * 2 * 3 * 8, G * 2 * 4, 0));
    }
  }
  CHECK(r.end(query));
  CHECK(r.flush());
  if (r.get_data(query, &out.stats, false) != async_status::ready) return out;
  if (S > 1) CHECK(resolve(cs, rs));
  CHECK(r.flush());
  mapped_resource m;
  CHECK(r.map(m, cs, map_read));
  out.color = fnv(m.data, W * H * S * 4);
  for (size_t i = 0; i < W * H * S; ++i) out.drawn += static_cast<const uint8_t*>(m.data)[4 * i + 2] != 128 ? 1 : 0;  // clear colour: (51, 51, 128, 255)
  r.unmap();
  CHECK(r.map(m, dss, map_read));
  out.depth = fnv(m.data, W * H * S * 8);
  r.unmap();
  if (S > 1) {
    CHECK(r.map(m, rs, map_read));
    out.resolved = fnv(m.data, W * H * 4);
    r.unmap();
  }
  out.ok = true;
  return out;
}

void print(const char* who, frame_hashes const& h) {
  std::printf("%s color %016" PRIx64 " depth %016" PRIx64 " resolved %016" PRIx64 " | ia_vertices %" PRIu64 " ia_primitives %" PRIu64
              " cinvocations %" PRIu64 " cprimitives %" PRIu64 " ps_invocations %" PRIu64 " | drawn samples %zu\n",
              who, h.color, h.depth, h.resolved, h.stats.ia_vertices, h.stats.ia_primitives, h.stats.cinvocations, h.stats.cprimitives,
              h.stats.ps_invocations, h.drawn);
}

}  // namespace

int main(int argc, char** argv) {
  if (argc < 2) { std::fprintf(stderr, "usage: %s <library.so> [w h samples]\n", argv[0]); return 1; }
  const size_t W = argc > 3 ? std::atoi(argv[2]) : 320, H = argc > 3 ? std::atoi(argv[3]) : 240, S = argc > 4 ? std::atoi(argv[4]) : 4;
  // the reference's own renderer
  renderer_ptr sync = create_benchmark_renderer();  // sync_renderer (renderer.cpp:37-39)
  // the same calls through the renderer_impl subclass into the C ABI
  std::shared_ptr<b200_renderer> b200;
  try {
    b200 = std::make_shared<b200_renderer>(argv[1]);
  } catch (std::exception const& e) {
    std::fprintf(stderr, "%s\n", e.what());
    return 2;
  }
  std::printf("backend %s\n", b200->backend_name().c_str());
  // SASL through the binding where the bound library compiles it on this machine: the CPU checker always (it builds the
  // generated code for the host), the CUDA product when asked (SLV_BRIDGE_SASL=1: NVRTC at run time); else the cpp twins
  // Pass 3 runs on the CPU checkers (and on the CUDA product when SLV_BRIDGE_PASS3 / SLV_BRIDGE_SASL ask for it).
  const bool use_sasl = b200->backend_name() == "oracle" || std::getenv("SLV_BRIDGE_SASL");
  const bool pass3 = use_sasl || b200->backend_name() != "cuda-sm100a" || std::getenv("SLV_BRIDGE_PASS3");
  std::printf("pass 3: %s\n", !pass3 ? "skipped" : use_sasl ? "SASL pair through compile() / set_*_shader_code" : "cpp twins");
  frame_hashes a = run_scene(*sync, W, H, S, [](surface_ptr const& src, surface_ptr const& dst) { src->resolve(*dst); return result::ok; }, pass3);
  frame_hashes b = run_scene(*b200, W, H, S, [&](surface_ptr const& src, surface_ptr const& dst) { return b200->resolve(src, dst); }, pass3, use_sasl ? b200.get() : nullptr);
  print("reference sync_renderer", a);
  print("b200_renderer -> C ABI ", b);
  if (!a.ok || !b.ok || a.drawn < W * H * S / 20) return 3;  // the scene must actually cover part of the target
  const bool same = a.color == b.color && a.depth == b.depth && a.resolved == b.resolved && a.stats.ia_vertices == b.stats.ia_vertices &&
                    a.stats.ia_primitives == b.stats.ia_primitives && a.stats.cinvocations == b.stats.cinvocations &&
                    a.stats.cprimitives == b.stats.cprimitives && a.stats.ps_invocations == b.stats.ps_invocations;
  std::printf(same ? "MATCH\n" : "MISMATCH\n");
  return same ? 0 : 4;
}

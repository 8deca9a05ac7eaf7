// oracle/ref_backend.cpp — TEST INFRASTRUCTURE ONLY.
//
// Drives the UNMODIFIED reference renderer (compiled in place from /root/reference by
// oracle/Makefile, target `ref`) through its own public API, salvia::core::renderer
// (salvia/include/salvia/core/renderer.h:42-131), and exposes it behind the slv_* C ABI of
// include/salvia_b200.h.  This is the ground truth the CUDA path and the CPU restatement
// (oracle/salvia_oracle.cpp) are compared against.  Nothing in the product links this file.
//
// The cpp_vertex_shader / cpp_pixel_shader / cpp_blend_shader subclasses below are the reference-side
// twins of the SLV_VS_* / SLV_PS_* / SLV_BS_* device programs; each cites the sample it restates.

#include "salvia_b200.h"

#include <salvia/core/async_object.h>
#include <salvia/core/framebuffer.h>
#include <salvia/core/raster_state.h>
#include <salvia/core/renderer.h>
#include <salvia/core/shader.h>
#include <salvia/core/sync_renderer.h>
#include <salvia/resource/buffer.h>
#include <salvia/resource/input_layout.h>
#include <salvia/resource/pixel_accessor.h>
#include <salvia/resource/sampler.h>
#include <salvia/resource/surface.h>
#include <salvia/resource/texture.h>
#include <salvia/shader/shader_regs.h>

#include <eflib/math/math.h>

#include <chrono>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

using namespace salvia;
using namespace salvia::core;
using namespace salvia::resource;
using namespace salvia::shader;
using eflib::mat44;
using eflib::vec2;
using eflib::vec3;
using eflib::vec4;

namespace {

#define SLV_CLONE()                                                         \
  cpp_shader_ptr clone() override {                                         \
    typedef std::remove_pointer<decltype(this)>::type this_type;            \
    return cpp_shader_ptr(new this_type(*this));                            \
  }

mat44 load_mat(float const* m) {
  return mat44(m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7], m[8], m[9], m[10], m[11], m[12],
               m[13], m[14], m[15]);
}

// ---------------------------------------------------------------------------------------------
// vertex shaders
struct vs_base : cpp_vertex_shader {
  uint32_t n_attrs = 0;
  uint32_t modifiers[SLV_MAX_VS_OUTPUT_ATTRS] = {};
  void bind_regs(slv_draw_desc const& d) {
    for (uint32_t i = 0; i < d.n_elements; ++i) {
      auto const& e = d.elements[i];
      bind_semantic(e.default_w == 1.0f ? "POSITION" : "TEXCOORD", e.reg, e.reg);
    }
    for (uint32_t i = 0; i < SLV_MAX_VS_OUTPUT_ATTRS; ++i) {
      modifiers[i] = d.vs_attr_modifiers[i] ? d.vs_attr_modifiers[i] : (uint32_t)vs_output::am_linear;
    }
  }
  uint32_t num_output_attributes() const override { return n_attrs; }
  uint32_t output_attribute_modifiers(uint32_t i) const override { return modifiers[i]; }
};

struct vs_mvp_passthrough : vs_base {
  mat44 wvp;
  uint32_t src[5];
  explicit vs_mvp_passthrough(slv_vs_mvp_passthrough_uniforms const& u) : wvp(load_mat(u.wvp)) {
    n_attrs = u.n_attrs;
    memcpy(src, u.src, sizeof(src));
  }
  void shader_prog(const vs_input& in, vs_output& out) override {
    vec4 pos = in.attribute(0);
    eflib::transform(out.position(), pos, wvp);
    for (uint32_t i = 0; i < n_attrs; ++i) {
      out.attribute(i) = in.attribute(src[i]);
    }
  }
  SLV_CLONE()
};

// samples/TextureAndBlending/TextureAndBlending.cpp:121-143 (vs_plane)
struct vs_plane_xz : vs_base {
  mat44 wvp;
  explicit vs_plane_xz(slv_vs_plane_xz_uniforms const& u) : wvp(load_mat(u.wvp)) { n_attrs = 1; }
  void shader_prog(const vs_input& in, vs_output& out) override {
    vec4 pos = in.attribute(0);
    eflib::transform(out.position(), pos, wvp);
    out.attribute(0) = vec4(in.attribute(0).x(), in.attribute(0).z(), 0, 0);
  }
  SLV_CLONE()
};

// cpp twin of the SASL VS in samples/ColorizedTriangle/ColorizedTriangle.cpp:29-53
struct vs_lights3 : vs_base {
  mat44 wvp;
  vec4 light_pos[3];
  explicit vs_lights3(slv_vs_lights3_uniforms const& u) : wvp(load_mat(u.wvp)) {
    n_attrs = 4;
    for (int k = 0; k < 3; ++k) {
      light_pos[k] = vec4(u.light_pos[k][0], u.light_pos[k][1], u.light_pos[k][2], u.light_pos[k][3]);
    }
  }
  void shader_prog(const vs_input& in, vs_output& out) override {
    vec4 pos = in.attribute(0);
    eflib::transform(out.position(), pos, wvp);
    out.attribute(0) = in.attribute(1);
    out.attribute(1) = light_pos[0] - pos;
    out.attribute(2) = light_pos[1] - pos;
    out.attribute(3) = light_pos[2] - pos;
  }
  SLV_CLONE()
};

// cpp twin of the SASL vertex shader of samples/VertexTextureFetch/VertexTextureFetch.cpp:38-61; tex2Dlod is the reference's
// own sasl.vs.tex2d.lod implementation: sampler::sample_2d_lod(coord.xy, coord.w) (salvia/src/resource/sampler_api.cpp:50-52)
struct vs_terrain_vtf : vs_base {
  mat44 wvp;
  float off[2], scale[2];
  sampler_ptr samp;
  vs_terrain_vtf(slv_vs_terrain_vtf_uniforms const& u, sampler_ptr const& s) : wvp(load_mat(u.wvp)), samp(s) {
    n_attrs = 1;
    off[0] = u.offset[0]; off[1] = u.offset[1]; scale[0] = u.scale[0]; scale[1] = u.scale[1];
  }
  void shader_prog(const vs_input& in, vs_output& out) override {
    vec4 pos = in.attribute(0), uv = in.attribute(1);
    eflib::vec2 terrain_uv(off[0] + uv.x() * scale[0], off[1] + uv.y() * scale[1]);
    float displacement = samp->sample_2d_lod(terrain_uv, 0.0f).get_vec4().x();
    vec4 displaced(pos.x() + 0.0f, pos.y() + displacement * 20.0f, pos.z() + 0.0f, 1.0f);
    eflib::transform(out.position(), displaced, wvp);
    out.attribute(0) = vec4(displacement, 0.0f, 0.0f, 0.0f);
  }
  SLV_CLONE()
};

// samples/Sponza/Sponza.cpp:64-97 (sponza_vs)
struct vs_sponza : vs_base {
  mat44 wvp;
  vec4 light_pos, eye_pos;
  explicit vs_sponza(slv_vs_sponza_uniforms const& u)
    : wvp(load_mat(u.wvp))
    , light_pos(u.light_pos[0], u.light_pos[1], u.light_pos[2], u.light_pos[3])
    , eye_pos(u.eye_pos[0], u.eye_pos[1], u.eye_pos[2], u.eye_pos[3]) {
    n_attrs = 4;
  }
  void shader_prog(const vs_input& in, vs_output& out) override {
    vec4 pos = in.attribute(0);
    eflib::transform(out.position(), pos, wvp);
    out.attribute(0) = in.attribute(1);
    out.attribute(1) = in.attribute(2);
    out.attribute(2) = light_pos - pos;
    out.attribute(3) = eye_pos - pos;
  }
  SLV_CLONE()
};

// cpp twin of resources/ssm/Draw.savs (the SASL vertex shader of samples/StandardShadowMap's colour pass)
struct vs_ssm_draw : vs_base {
  mat44 camera_wvp, light_wvp;
  vec4 light_pos, camera_pos;
  explicit vs_ssm_draw(slv_vs_ssm_draw_uniforms const& u)
    : camera_wvp(load_mat(u.camera_wvp))
    , light_wvp(load_mat(u.light_wvp))
    , light_pos(u.light_pos[0], u.light_pos[1], u.light_pos[2], u.light_pos[3])
    , camera_pos(u.camera_pos[0], u.camera_pos[1], u.camera_pos[2], u.camera_pos[3]) {
    n_attrs = 5;
  }
  void shader_prog(const vs_input& in, vs_output& out) override {
    vec4 pos = in.attribute(0);
    out.attribute(1) = in.attribute(1);                          // norm
    eflib::transform(out.position(), pos, camera_wvp);
    eflib::transform(out.attribute(4), pos, light_wvp);          // lightSpacePos
    out.attribute(2) = light_pos - pos;                          // lightDir
    out.attribute(3) = camera_pos - pos;                         // cameraDir
    out.attribute(0) = in.attribute(2);                          // tex
  }
  SLV_CLONE()
};

// ---------------------------------------------------------------------------------------------
// pixel shaders
struct ps_attr0_color : cpp_pixel_shader {
  bool shader_prog(const vs_output& in, ps_output& out) override {
    out.color[0] = in.attribute(0);
    return true;
  }
  SLV_CLONE()
};

// samples/VertexTextureFetch/VertexTextureFetch.cpp:70-113
struct ps_height_color : cpp_pixel_shader {
  bool shader_prog(const vs_output& in, ps_output& out) override {
    float height = in.attribute(0)[0];
    vec4 colors[] = {vec4(0.0f, 0.0f, 0.5f, 1.0f), vec4(0.7f, 0.6f, 0.0f, 1.0f), vec4(0.45f, 0.38f, 0.26f, 1.0f),
                     vec4(0.0f, 0.7f, 0.8f, 1.0f), vec4(0.9f, 0.9f, 1.0f, 1.0f), vec4(0.9f, 0.9f, 1.0f, 1.0f)};
    float boundary_points[] = {0.0f, 0.62f, 0.75f, 0.88f, 1.0f, 1.0f};
    int lower_bound = -1;
    for (int i = 0; i < 5; ++i) {
      if (height < boundary_points[i]) break;
      lower_bound = i;
    }
    if (lower_bound == -1) {
      out.color[0] = colors[0];
    } else {
      float lower_value = boundary_points[lower_bound];
      float interval = boundary_points[lower_bound + 1] - lower_value;
      out.color[0] = eflib::lerp(colors[lower_bound], colors[lower_bound + 1], (height - lower_value) / interval);
    }
    return true;
  }
  SLV_CLONE()
};

// samples/ColorizedTriangle/ColorizedTriangle.cpp:55-92
struct ps_lights3 : cpp_pixel_shader {
  bool shader_prog(const vs_output& in, ps_output& out) override {
    vec3 lightDir0 = in.attribute(1).xyz();
    vec3 lightDir1 = in.attribute(2).xyz();
    vec3 lightDir2 = in.attribute(3).xyz();

    vec3 norm = in.attribute(0).xyz();

    float invLight0Distance = 1.0f / lightDir0.length();
    float invLight1Distance = 1.0f / lightDir1.length();
    float invLight2Distance = 1.0f / lightDir2.length();

    vec3 normalized_norm = eflib::normalize3(norm);
    vec3 normalized_lightDir0 = lightDir0 * invLight0Distance;
    vec3 normalized_lightDir1 = lightDir1 * invLight1Distance;
    vec3 normalized_lightDir2 = lightDir2 * invLight2Distance;

    float refl0 = eflib::dot_prod3(normalized_norm, normalized_lightDir0);
    float refl1 = eflib::dot_prod3(normalized_norm, normalized_lightDir1);
    float refl2 = eflib::dot_prod3(normalized_norm, normalized_lightDir2);

    out.color[0] = eflib::clampss(
        vec4(0.7f, 0.1f, 0.3f, 1.0f) * refl0 * invLight0Distance * invLight0Distance +
            vec4(0.1f, 0.3f, 0.7f, 1.0f) * refl1 * invLight1Distance * invLight1Distance +
            vec4(0.3f, 0.7f, 0.1f, 1.0f) * refl2 * invLight2Distance * invLight2Distance,
        0.0f,
        1.0f);
    out.color[0][3] = 1.0f;
    return true;
  }
  SLV_CLONE()
};

// samples/TextureAndBlending/TextureAndBlending.cpp:96-166 (ps_box / ps_plane)
struct ps_tex_alpha : cpp_pixel_shader {
  sampler_ptr sampler_;
  uint32_t reg;
  float alpha;
  ps_tex_alpha(sampler_ptr const& s, slv_ps_tex_alpha_uniforms const& u)
    : sampler_(s), reg(u.reg), alpha(u.alpha) {}
  bool shader_prog(const vs_output& /*in*/, ps_output& out) override {
    color_rgba32f color = tex2d(*sampler_, reg);
    color.a = alpha;
    out.color[0] = color.get_vec4();
    return true;
  }
  SLV_CLONE()
};

// The SASL tex2D path: tex2D(s,uv) == tex2Dgrad(s, uv, ddx(uv), ddy(uv)) -> sampler::sample_2d_grad
// (sasl/src/codegen/cg_impl.cpp:902-909, salvia/src/resource/sampler_api.cpp:13-18), with the cpp
// quad-derivative convention (cpp_pixel_shader.cpp:13-19).
// sasl_derivatives: ddx per quad row / ddy per quad column (cgs_simd.cpp:275-313).  cpp_pixel_shader keeps the quad private,
// but execute() (cpp_pixel_shader.cpp:63-75) calls shader_prog for pixels 0..3 of `quad` in order with in == quad[i], and every
// worker owns its clone - so the call count modulo 4 is the pixel's index and &in - i the quad.
struct ps_tex_grad_alpha : cpp_pixel_shader {
  sampler_ptr sampler_;
  uint32_t reg;
  float alpha;
  bool sasl_derivatives;
  unsigned calls = 0;
  ps_tex_grad_alpha(sampler_ptr const& s, slv_ps_tex_alpha_uniforms const& u)
    : sampler_(s), reg(u.reg), alpha(u.alpha), sasl_derivatives(u.sasl_derivatives != 0) {}
  bool shader_prog(const vs_output& in, ps_output& out) override {
    unsigned const i = calls++ & 3u;
    color_rgba32f color;
    if (sasl_derivatives) {
      vs_output const* quad = &in - i;
      vec4 dx = quad[i | 1u].attribute(reg) - quad[i & ~1u].attribute(reg);
      vec4 dy = quad[i | 2u].attribute(reg) - quad[i & ~2u].attribute(reg);
      color = sampler_->sample_2d_grad(in.attribute(reg).xy(), dx.xy(), dy.xy(), 0.0f);
    } else {
      color = sampler_->sample_2d_grad(in.attribute(reg).xy(), ddx(reg).xy(), ddy(reg).xy(), 0.0f);
    }
    color.a = alpha;
    out.color[0] = color.get_vec4();
    return true;
  }
  SLV_CLONE()
};

// samples/Sponza/Sponza.cpp:99-141 (sponza_ps)
struct ps_sponza : cpp_pixel_shader {
  sampler_ptr sampler_;
  explicit ps_sponza(sampler_ptr const& s) : sampler_(s) {}
  bool shader_prog(const vs_output& in, ps_output& out) override {
    vec4 diff_color = vec4(1.0f, 1.0f, 1.0f, 1.0f);
    if (sampler_) {
      diff_color = tex2d(*sampler_, 0).get_vec4();
    }
    vec3 norm(eflib::normalize3(in.attribute(1).xyz()));
    vec3 light_dir(eflib::normalize3(in.attribute(2).xyz()));
    float illum_diffuse = eflib::clamp(eflib::dot_prod3(light_dir, norm), 0.0f, 1.0f);
    out.color[0] = diff_color * illum_diffuse;
    out.color[0][3] = 1.0f;
    return true;
  }
  SLV_CLONE()
};

// sponza_ps with the diffuse texture fetched the way a SASL pixel shader's tex2D fetches it (the reference's own
// sampler::sample_2d_grad; sasl/src/codegen/cg_impl.cpp:902-909).  cpp_pixel_shader keeps the quad private, but execute()
// (cpp_pixel_shader.cpp:63-75) calls shader_prog for pixels 0..3 of `quad` in order with in == quad[i], and every worker owns
// its clone - so the call count modulo 4 is the pixel's index and &in - i the quad.  sasl_derivatives: per quad row / column
// (cgs_simd.cpp:275-313) instead of q1 - q0 / q2 - q0.
struct ps_sponza_grad : cpp_pixel_shader {
  sampler_ptr sampler_;
  bool sasl_derivatives;
  unsigned calls = 0;
  ps_sponza_grad(sampler_ptr const& s, bool sasl) : sampler_(s), sasl_derivatives(sasl) {}
  bool shader_prog(const vs_output& in, ps_output& out) override {
    unsigned const i = calls++ & 3u;
    vs_output const* quad = &in - i;
    vec4 diff_color = vec4(1.0f, 1.0f, 1.0f, 1.0f);
    if (sampler_) {
      unsigned const pi = sasl_derivatives ? i : 0u;
      vec4 dx = quad[pi | 1u].attribute(0) - quad[pi & ~1u].attribute(0);
      vec4 dy = quad[pi | 2u].attribute(0) - quad[pi & ~2u].attribute(0);
      diff_color = sampler_->sample_2d_grad(in.attribute(0).xy(), dx.xy(), dy.xy(), 0.0f).get_vec4();
    }
    vec3 norm(eflib::normalize3(in.attribute(1).xyz()));
    vec3 light_dir(eflib::normalize3(in.attribute(2).xyz()));
    float illum_diffuse = eflib::clamp(eflib::dot_prod3(light_dir, norm), 0.0f, 1.0f);
    out.color[0] = diff_color * illum_diffuse;
    out.color[0][3] = 1.0f;
    return true;
  }
  SLV_CLONE()
};

// draw_cpp_ps of samples/StandardShadowMap/StandardShadowMap.cpp:62-142, restated over the reference's own classes
// (two samplers, nine tex2dlod taps of the shadow map, exponential shadow map).  The overloads the sample leaves to the
// platform's <cmath> are pinned: exp / log in float, pow(float, int) in double (include/salvia_b200.h, SLV_PS_SSM_DRAW).
struct ps_ssm_draw : cpp_pixel_shader {
  sampler_ptr texsamp_, dsamp_;
  vec4 ambient, diffuse, specular;
  int shininess;
  ps_ssm_draw(slv_ps_ssm_draw_uniforms const& u, sampler_ptr const& tex, sampler_ptr const& depth)
    : texsamp_(tex), dsamp_(depth)
    , ambient(u.ambient[0], u.ambient[1], u.ambient[2], u.ambient[3])
    , diffuse(u.diffuse[0], u.diffuse[1], u.diffuse[2], u.diffuse[3])
    , specular(u.specular[0], u.specular[1], u.specular[2], u.specular[3])
    , shininess(u.shininess) {}
  bool shader_prog(const vs_output& in, ps_output& out) override {
    const float esm_constant = 25000.0f;
    static const float gaussian_weights[9] = {0.027681f, 0.111014f, 0.027681f, 0.111014f, 0.445213f,
                                              0.111014f, 0.027681f, 0.111014f, 0.027681f};
    float occlusion = 0.0f;
    if (dsamp_) {
      vec3 lis_pos(in.attribute(4).xyz() / in.attribute(4).w());
      vec2 sm_center((lis_pos.x() + 1.0f) * 0.5f, (1.0f - (lis_pos.y() + 1.0f) * 0.5f));
      float sm_offset = 1 / 512.0f;
      float shadow_depth[9];
      for (int i = 0; i < 9; ++i) {
        vec2 d(i % 3 == 0 ? -sm_offset : (i % 3 == 1 ? 0.0f : +sm_offset), i / 3 == 0 ? -sm_offset : (i / 3 == 1 ? 0.0f : +sm_offset));
        vec2 c = sm_center + d;
        vec4 coord_lod(c.x(), c.y(), 0.0f, 0.0f);
        shadow_depth[i] = tex2dlod(*dsamp_, coord_lod).r;
      }
      float occluder = 0.0f;
      for (int i = 1; i < 9; ++i) {
        occluder += gaussian_weights[i] * std::exp(esm_constant * (shadow_depth[i] - shadow_depth[0]));
      }
      occluder += gaussian_weights[0];
      occluder = std::log(occluder);
      occluder += esm_constant * shadow_depth[0];
      occlusion = eflib::clamp(std::exp(occluder - esm_constant * lis_pos[2]), 0.0f, 1.0f);
    }
    color_rgba32f tex_color(1.0f, 1.0f, 1.0f, 1.0f);
    if (texsamp_) {
      tex_color = tex2d(*texsamp_, 0);
    }
    vec3 norm(eflib::normalize3(in.attribute(1).xyz()));
    vec3 light_dir(eflib::normalize3(in.attribute(2).xyz()));
    vec3 eye_dir(eflib::normalize3(in.attribute(3).xyz()));
    float illum_diffuse = eflib::clamp(eflib::dot_prod3(light_dir, norm), 0.0f, 1.0f);
    float illum_specular = eflib::clamp(eflib::dot_prod3(-eflib::reflect3(light_dir, norm), eye_dir), 0.0f, 1.0f);
    float sp = (float)std::pow((double)illum_specular, (double)shininess);
    vec4 illum = ambient + (diffuse * illum_diffuse + specular * sp) * occlusion;
    out.color[0] = tex_color.get_vec4() * illum;
    out.color[0][3] = 1.0f;
    return true;
  }
  SLV_CLONE()
};

struct ps_discard_all : cpp_pixel_shader {
  bool shader_prog(const vs_output& in, ps_output& out) override {
    out.color[0] = in.attribute(0);
    return false;
  }
  SLV_CLONE()
};

// ---------------------------------------------------------------------------------------------
// blend shaders
struct bs_replace : cpp_blend_shader {  // ColorizedTriangle.cpp:94-106
  bool shader_prog(size_t sample, pixel_accessor& inout, const ps_output& in) override {
    inout.color(0, sample, color_rgba32f(in.color[0]));
    return true;
  }
  SLV_CLONE()
};

struct bs_lerp_src_alpha : cpp_blend_shader {  // TextureAndBlending.cpp:168-180
  bool shader_prog(size_t sample, pixel_accessor& inout, const ps_output& in) override {
    color_rgba32f color(in.color[0]);
    inout.color(0, sample, lerp(inout.color(0, sample), color, color.a));
    return true;
  }
  SLV_CLONE()
};

struct bs_replace_and_count : cpp_blend_shader {
  bool shader_prog(size_t sample, pixel_accessor& inout, const ps_output& in) override {
    inout.color(0, sample, color_rgba32f(in.color[0]));
    color_rgba32f c = inout.color(1, sample);
    c.r += 1.0f;
    inout.color(1, sample, c);
    return true;
  }
  SLV_CLONE()
};

// ---------------------------------------------------------------------------------------------
struct resource_entry {
  buffer_ptr buf;
  texture_ptr tex;
  sampler_ptr samp;
};

}  // namespace

struct slv_device_t {
  renderer_ptr r;
  std::vector<resource_entry> res;  // index = handle; [0] unused
  async_object_ptr q_stat, q_internal, q_prof;
  // State objects are kept alive and re-used per distinct descriptor, like an application does: the
  // reference caches on the POINTER identity of the depth-stencil state (framebuffer.cpp:326,361-363), so
  // re-allocating a state object per draw could alias a stale one at a recycled address.
  std::vector<std::pair<slv_depth_stencil_desc, depth_stencil_state_ptr>> ds_states;
  std::vector<std::pair<slv_raster_desc, raster_state_ptr>> rs_states;
  bool q_active = false;
  std::chrono::steady_clock::time_point events[16];

  resource_entry* get(slv_handle h) {
    if (h == 0 || h >= res.size()) return nullptr;
    return &res[h];
  }
  surface_ptr surface_of(slv_handle h) {
    auto e = get(h);
    if (!e || !e->tex) return {};
    return e->tex->subresource(0);
  }
};

extern "C" {

const char* slv_backend_name(void) { return "reference"; }
uint32_t slv_abi_version(void) { return SLV_ABI_VERSION; }

slv_result slv_device_create(int32_t, slv_device* out) {
  if (!out) return SLV_INVALID_PARAMETER;
  auto d = new slv_device_t;
  d->r = create_benchmark_renderer();  // sync renderer (renderer.cpp:37-39)
  d->res.resize(1);
  *out = d;
  return SLV_OK;
}

void slv_device_destroy(slv_device dev) { delete dev; }

slv_result slv_buffer_create(slv_device dev, size_t bytes, slv_handle* out) {
  resource_entry e;
  e.buf = dev->r->create_buffer(bytes);
  if (!e.buf) return SLV_OUT_OF_MEMORY;
  dev->res.push_back(e);
  *out = (slv_handle)(dev->res.size() - 1);
  return SLV_OK;
}

slv_result slv_buffer_upload(slv_device dev, slv_handle h, size_t offset, const void* src, size_t bytes) {
  auto e = dev->get(h);
  if (!e || !e->buf || offset + bytes > e->buf->size()) return SLV_INVALID_PARAMETER;
  memcpy(e->buf->raw_data(offset), src, bytes);
  return SLV_OK;
}

slv_result slv_buffer_readback(slv_device dev, slv_handle h, size_t offset, void* dst, size_t bytes) {
  auto e = dev->get(h);
  if (!e || !e->buf || offset + bytes > e->buf->size()) return SLV_INVALID_PARAMETER;
  dev->r->flush();
  memcpy(dst, e->buf->raw_data(offset), bytes);
  return SLV_OK;
}

slv_result slv_texture_create(slv_device dev, uint32_t w, uint32_t h, uint32_t samples, uint32_t pf,
                              slv_handle* out) {
  if (pf != SLV_PF_RGBA32F && pf != SLV_PF_BGRA8 && pf != SLV_PF_RGBA8 && pf != SLV_PF_RG32F)
    return SLV_INVALID_PARAMETER;
  resource_entry e;
  e.tex = dev->r->create_tex2d(w, h, samples, (pixel_format)pf);
  if (!e.tex) return SLV_OUT_OF_MEMORY;
  dev->res.push_back(e);
  *out = (slv_handle)(dev->res.size() - 1);
  return SLV_OK;
}

slv_result slv_texture_gen_mipmap(slv_device dev, slv_handle h, uint32_t filter) {
  auto e = dev->get(h);
  if (!e || !e->tex) return SLV_INVALID_PARAMETER;
  dev->r->flush();
  e->tex->gen_mipmap((filter_type)filter, true);
  return SLV_OK;
}

slv_result slv_texture_level_count(slv_device dev, slv_handle h, uint32_t* out) {
  auto e = dev->get(h);
  if (!e || !e->tex) return SLV_INVALID_PARAMETER;
  *out = (uint32_t)(e->tex->min_lod() - e->tex->max_lod() + 1);
  return SLV_OK;
}

slv_result slv_texture_level_size(slv_device dev, slv_handle h, uint32_t level, uint32_t* w, uint32_t* hh) {
  auto e = dev->get(h);
  if (!e || !e->tex) return SLV_INVALID_PARAMETER;
  auto s = e->tex->subresource(level);
  if (!s) return SLV_INVALID_PARAMETER;
  *w = (uint32_t)s->width();
  *hh = (uint32_t)s->height();
  return SLV_OK;
}

slv_result slv_texture_upload(slv_device dev, slv_handle h, uint32_t level, const void* src, size_t bytes) {
  auto e = dev->get(h);
  if (!e || !e->tex) return SLV_INVALID_PARAMETER;
  auto s = e->tex->subresource(level);
  if (!s || bytes != s->pitch() * s->height()) return SLV_INVALID_PARAMETER;
  memcpy(s->texel_address(0, 0, 0), src, bytes);
  return SLV_OK;
}

slv_result slv_texture_readback(slv_device dev, slv_handle h, uint32_t level, void* dst, size_t bytes) {
  auto e = dev->get(h);
  if (!e || !e->tex) return SLV_INVALID_PARAMETER;
  auto s = e->tex->subresource(level);
  if (!s || bytes != s->pitch() * s->height()) return SLV_INVALID_PARAMETER;
  dev->r->flush();
  memcpy(dst, s->texel_address(0, 0, 0), bytes);
  return SLV_OK;
}

slv_result slv_sampler_create(slv_device dev, const slv_sampler_desc* d, slv_handle tex, slv_handle* out) {
  auto e = dev->get(tex);
  if (!e || !e->tex) return SLV_INVALID_PARAMETER;
  sampler_desc sd;
  sd.min_filter = (filter_type)d->min_filter;
  sd.mag_filter = (filter_type)d->mag_filter;
  sd.mip_filter = (filter_type)d->mip_filter;
  sd.mip_qual = (mip_quality)d->mip_qual;
  sd.addr_mode_u = (address_mode)d->addr_mode_u;
  sd.addr_mode_v = (address_mode)d->addr_mode_v;
  sd.addr_mode_w = (address_mode)d->addr_mode_w;
  sd.mip_lod_bias = d->mip_lod_bias;
  sd.max_anisotropy = d->max_anisotropy;
  sd.comparison_func = (compare_function)d->comparison_func;
  sd.border_color = color_rgba32f(d->border_color[0], d->border_color[1], d->border_color[2], d->border_color[3]);
  sd.min_lod = d->min_lod;
  sd.max_lod = d->max_lod;
  resource_entry ne;
  ne.samp = dev->r->create_sampler(sd, e->tex);
  dev->res.push_back(ne);
  *out = (slv_handle)(dev->res.size() - 1);
  return SLV_OK;
}

slv_result slv_resource_release(slv_device dev, slv_handle h) {
  auto e = dev->get(h);
  if (!e) return SLV_INVALID_PARAMETER;
  dev->r->flush();
  *e = resource_entry();
  return SLV_OK;
}

static sampler_ptr sampler_of(slv_device dev, slv_handle h) {
  auto e = dev->get(h);
  return e ? e->samp : sampler_ptr();
}

slv_result slv_draw(slv_device dev, const slv_draw_desc* d) {
  renderer* r = dev->r.get();

  // shaders
  std::shared_ptr<vs_base> vs;
  switch (d->vs.program) {
  case SLV_VS_MVP_PASSTHROUGH:
    vs.reset(new vs_mvp_passthrough(*(slv_vs_mvp_passthrough_uniforms const*)d->vs.uniforms));
    break;
  case SLV_VS_PLANE_XZ: vs.reset(new vs_plane_xz(*(slv_vs_plane_xz_uniforms const*)d->vs.uniforms)); break;
  case SLV_VS_LIGHTS3: vs.reset(new vs_lights3(*(slv_vs_lights3_uniforms const*)d->vs.uniforms)); break;
  case SLV_VS_SPONZA: vs.reset(new vs_sponza(*(slv_vs_sponza_uniforms const*)d->vs.uniforms)); break;
  case SLV_VS_SSM_DRAW: vs.reset(new vs_ssm_draw(*(slv_vs_ssm_draw_uniforms const*)d->vs.uniforms)); break;
  case SLV_VS_TERRAIN_VTF: {
    sampler_ptr s = sampler_of(dev, d->vs.samplers[0]);
    if (!s) return SLV_INVALID_PARAMETER;
    vs.reset(new vs_terrain_vtf(*(slv_vs_terrain_vtf_uniforms const*)d->vs.uniforms, s));
  } break;
  default: return SLV_INVALID_PARAMETER;
  }
  vs->bind_regs(*d);

  cpp_pixel_shader_ptr ps;
  switch (d->ps.program) {
  case SLV_PS_ATTR0_COLOR: ps.reset(new ps_attr0_color()); break;
  case SLV_PS_LIGHTS3: ps.reset(new ps_lights3()); break;
  case SLV_PS_TEX_ALPHA:
    ps.reset(new ps_tex_alpha(sampler_of(dev, d->ps.samplers[0]),
                              *(slv_ps_tex_alpha_uniforms const*)d->ps.uniforms));
    break;
  case SLV_PS_TEX_GRAD_ALPHA:
    ps.reset(new ps_tex_grad_alpha(sampler_of(dev, d->ps.samplers[0]),
                                   *(slv_ps_tex_alpha_uniforms const*)d->ps.uniforms));
    break;
  case SLV_PS_SPONZA: {
    auto u = (slv_ps_sponza_uniforms const*)d->ps.uniforms;
    ps.reset(new ps_sponza(u->has_sampler ? sampler_of(dev, d->ps.samplers[0]) : sampler_ptr()));
  } break;
  case SLV_PS_SPONZA_GRAD: {
    auto u = (slv_ps_sponza_grad_uniforms const*)d->ps.uniforms;
    ps.reset(new ps_sponza_grad(u->has_sampler ? sampler_of(dev, d->ps.samplers[0]) : sampler_ptr(), u->sasl_derivatives != 0));
  } break;
  case SLV_PS_DISCARD_ALL: ps.reset(new ps_discard_all()); break;
  case SLV_PS_HEIGHT_COLOR: ps.reset(new ps_height_color()); break;
  case SLV_PS_SSM_DRAW: {
    auto u = (slv_ps_ssm_draw_uniforms const*)d->ps.uniforms;
    sampler_ptr tex = u->has_tex_sampler ? sampler_of(dev, d->ps.samplers[0]) : sampler_ptr();
    sampler_ptr depth = u->has_depth_sampler ? sampler_of(dev, d->ps.samplers[1]) : sampler_ptr();
    if ((u->has_tex_sampler && !tex) || (u->has_depth_sampler && !depth)) return SLV_INVALID_PARAMETER;
    ps.reset(new ps_ssm_draw(*u, tex, depth));
  } break;
  default: return SLV_INVALID_PARAMETER;
  }

  cpp_blend_shader_ptr bs;
  switch (d->bs.program) {
  case SLV_BS_REPLACE: bs.reset(new bs_replace()); break;
  case SLV_BS_LERP_SRC_ALPHA: bs.reset(new bs_lerp_src_alpha()); break;
  case SLV_BS_REPLACE_AND_COUNT: bs.reset(new bs_replace_and_count()); break;
  default: return SLV_INVALID_PARAMETER;
  }

  // input assembler
  std::vector<input_element_desc> descs;
  for (uint32_t i = 0; i < d->n_elements; ++i) {
    auto const& e = d->elements[i];
    descs.emplace_back(e.default_w == 1.0f ? "POSITION" : "TEXCOORD", e.reg, (format)e.format, e.slot,
                       e.aligned_byte_offset, input_per_vertex, 0);
  }
  cpp_vertex_shader_ptr vsp = vs;
  input_layout_ptr layout = r->create_input_layout(descs.data(), descs.size(), vsp);
  for (uint32_t i = 0; i < d->n_streams; ++i) {
    auto e = dev->get(d->streams[i].buffer);
    if (!e || !e->buf) return SLV_INVALID_PARAMETER;
    size_t stride = d->streams[i].stride, offset = d->streams[i].offset;
    r->set_vertex_buffers(i, 1, &e->buf, &stride, &offset);
  }
  if (d->index_buffer) {
    auto e = dev->get(d->index_buffer);
    if (!e || !e->buf) return SLV_INVALID_PARAMETER;
    if (r->set_index_buffer(e->buf, (format)d->index_format) != result::ok) return SLV_FAILED;
  }
  r->set_input_layout(layout);
  if (r->set_primitive_topology((primitive_topology)d->topology) != result::ok) return SLV_FAILED;

  r->set_vertex_shader(vsp);
  r->set_pixel_shader(ps);
  r->set_blend_shader(bs);

  raster_desc rd;
  rd.cm = (cull_mode)d->raster.cull_mode;
  rd.front_ccw = d->raster.front_ccw != 0;
  {
    raster_state_ptr rs;
    for (auto& kv : dev->rs_states)
      if (memcmp(&kv.first, &d->raster, sizeof(slv_raster_desc)) == 0) rs = kv.second;
    if (!rs) {
      rs.reset(new raster_state(rd));
      dev->rs_states.emplace_back(d->raster, rs);
    }
    r->set_rasterizer_state(rs);
  }

  depth_stencil_desc dd;
  dd.depth_enable = d->ds.depth_enable != 0;
  dd.depth_write_mask = d->ds.depth_write_mask != 0;
  dd.depth_func = (compare_function)d->ds.depth_func;
  dd.stencil_enable = d->ds.stencil_enable != 0;
  dd.stencil_read_mask = (uint8_t)d->ds.stencil_read_mask;
  dd.stencil_write_mask = (uint8_t)d->ds.stencil_write_mask;
  auto cvt = [](slv_stencil_op_desc const& s) {
    depth_stencil_op_desc o;
    o.stencil_fail_op = (stencil_op)s.stencil_fail_op;
    o.stencil_depth_fail_op = (stencil_op)s.stencil_depth_fail_op;
    o.stencil_pass_op = (stencil_op)s.stencil_pass_op;
    o.stencil_func = (compare_function)s.stencil_func;
    return o;
  };
  dd.front_face = cvt(d->ds.front_face);
  dd.back_face = cvt(d->ds.back_face);
  {
    depth_stencil_state_ptr dss;
    for (auto& kv : dev->ds_states)
      if (memcmp(&kv.first, &d->ds, sizeof(slv_depth_stencil_desc)) == 0) dss = kv.second;
    if (!dss) {
      dss.reset(new depth_stencil_state(dd));
      dev->ds_states.emplace_back(d->ds, dss);
    }
    r->set_depth_stencil_state(dss, d->stencil_ref);
  }

  std::vector<surface_ptr> colors;
  for (uint32_t i = 0; i < d->n_color_targets; ++i) colors.push_back(dev->surface_of(d->color_targets[i]));
  surface_ptr ds = dev->surface_of(d->ds_target);
  if (r->set_render_targets(colors.size(), colors.data(), ds) != result::ok) return SLV_FAILED;

  viewport vp{d->viewport.x, d->viewport.y, d->viewport.w, d->viewport.h, d->viewport.minz, d->viewport.maxz};
  if (r->set_viewport(vp) != result::ok) return SLV_FAILED;

  result rc = d->index_buffer ? r->draw_index(d->start, d->prim_count, d->base_vertex)
                              : r->draw(d->start, d->prim_count);
  return rc == result::ok ? SLV_OK : SLV_FAILED;
}

slv_result slv_clear_color(slv_device dev, slv_handle h, const float rgba[4]) {
  auto s = dev->surface_of(h);
  if (!s) return SLV_INVALID_PARAMETER;
  return dev->r->clear_color(s, color_rgba32f(rgba[0], rgba[1], rgba[2], rgba[3])) == result::ok ? SLV_OK
                                                                                                  : SLV_FAILED;
}

slv_result slv_clear_depth_stencil(slv_device dev, slv_handle h, uint32_t flags, float depth, uint32_t stencil) {
  auto s = dev->surface_of(h);
  if (!s) return SLV_INVALID_PARAMETER;
  return dev->r->clear_depth_stencil(s, flags, depth, stencil) == result::ok ? SLV_OK : SLV_FAILED;
}

slv_result slv_resolve(slv_device dev, slv_handle src, slv_handle dst) {
  auto s = dev->surface_of(src), t = dev->surface_of(dst);
  if (!s || !t) return SLV_INVALID_PARAMETER;
  dev->r->flush();
  s->resolve(*t);
  return SLV_OK;
}

slv_result slv_flush(slv_device dev) { return dev->r->flush() == result::ok ? SLV_OK : SLV_FAILED; }

slv_result slv_query_begin(slv_device dev) {
  if (dev->q_active) {
    dev->r->end(dev->q_stat);
    dev->r->end(dev->q_internal);
    dev->r->end(dev->q_prof);
  }
  dev->q_stat = dev->r->create_query(async_object_ids::pipeline_statistics);
  dev->q_internal = dev->r->create_query(async_object_ids::internal_statistics);
  dev->r->begin(dev->q_stat);
  dev->q_prof = dev->r->create_query(async_object_ids::pipeline_profiles);
  dev->r->begin(dev->q_internal);
  dev->r->begin(dev->q_prof);
  dev->q_active = true;
  return SLV_OK;
}

static void end_queries(slv_device dev) {
  if (dev->q_active) {
    dev->r->end(dev->q_stat);
    dev->r->end(dev->q_internal);
    dev->r->end(dev->q_prof);
    dev->q_active = false;
  }
}

slv_result slv_profile_get(slv_device dev, slv_pipeline_profiles* out) {
  if (!dev->q_prof) return SLV_FAILED;
  end_queries(dev);
  pipeline_profiles pp;
  if (dev->r->get_data(dev->q_prof, &pp, false) != async_status::ready) return SLV_FAILED;
  out->gather_vtx = pp.gather_vtx;
  out->vtx_proc = pp.vtx_proc;
  out->clipping = pp.clipping;
  out->compact_clip = pp.compact_clip;
  out->vp_trans = pp.vp_trans;
  out->tri_dispatch = pp.tri_dispatch;
  out->ras = pp.ras;
  return SLV_OK;
}

slv_result slv_profile_get_stages(slv_device, double* ms, uint32_t n) {
  if (!ms || n < 6) return SLV_INVALID_PARAMETER;
  for (uint32_t i = 0; i < n; ++i) ms[i] = 0.0;
  return SLV_OK;
}
slv_result slv_set_tile_shard(slv_device, uint32_t rank, uint32_t nranks) {
  return (rank == 0 && nranks == 1) ? SLV_OK : SLV_FAILED;  // the reference is single-process
}

slv_result slv_query_get(slv_device dev, slv_pipeline_statistics* out) {
  if (!dev->q_stat) return SLV_FAILED;
  end_queries(dev);
  pipeline_statistics ps;
  internal_statistics is;
  if (dev->r->get_data(dev->q_stat, &ps, false) != async_status::ready) return SLV_FAILED;
  if (dev->r->get_data(dev->q_internal, &is, false) != async_status::ready) return SLV_FAILED;
  out->ia_vertices = ps.ia_vertices;
  out->ia_primitives = ps.ia_primitives;
  out->vs_invocations = ps.vs_invocations;
  out->gs_invocations = ps.gs_invocations;
  out->gs_primitives = ps.gs_primitives;
  out->cinvocations = ps.cinvocations;
  out->cprimitives = ps.cprimitives;
  out->ps_invocations = ps.ps_invocations;
  out->backend_input_pixels = is.backend_input_pixels;
  return SLV_OK;
}

// measurement / multi-GPU plumbing: the reference has no such facilities; timers use the host clock
slv_result slv_traffic_get(slv_device, slv_traffic_counters*) { return SLV_FAILED; }
slv_result slv_kernel_launch_count(slv_device, uint64_t* out) { *out = 0; return SLV_OK; }
slv_result slv_event_record(slv_device dev, uint32_t slot) {
  if (slot >= 16) return SLV_INVALID_PARAMETER;
  dev->r->flush();
  dev->events[slot] = std::chrono::steady_clock::now();
  return SLV_OK;
}
slv_result slv_event_elapsed_ms(slv_device dev, uint32_t a, uint32_t b, float* ms) {
  if (a >= 16 || b >= 16) return SLV_INVALID_PARAMETER;
  *ms = std::chrono::duration<float, std::milli>(dev->events[b] - dev->events[a]).count();
  return SLV_OK;
}
slv_result slv_profile_enable(slv_device, uint32_t) { return SLV_OK; }
slv_result slv_set_stream(slv_device, void*) { return SLV_OK; }
slv_result slv_texture_readback_async(slv_device dev, slv_handle tex, uint32_t level, void* dst, size_t bytes) {
  return slv_texture_readback(dev, tex, level, dst, bytes);
}
slv_result slv_readback_wait(slv_device) { return SLV_OK; }
slv_result slv_readback_fence(slv_device, slv_handle) { return SLV_OK; }
slv_result slv_buffer_device_ptr(slv_device dev, slv_handle h, void** out, size_t* bytes) {
  auto e = dev->get(h);
  if (!e || !e->buf || !out) return SLV_INVALID_PARAMETER;
  dev->r->flush();
  *out = e->buf->raw_data(0);
  if (bytes) *bytes = e->buf->size();
  return SLV_OK;
}
slv_result slv_external_write_begin(slv_device dev, void*, uint32_t) { dev->r->flush(); return SLV_OK; }
slv_result slv_external_write_end(slv_device, void*) { return SLV_OK; }
slv_result slv_assembly_wait(slv_device, slv_handle, const void*, uint32_t, uint32_t, uint32_t) { return SLV_FAILED; }
slv_result slv_peer_signal_after_consumers(slv_device, slv_handle, void*, uint32_t, uint32_t) { return SLV_FAILED; }
slv_result slv_host_register(slv_device, void*, size_t) { return SLV_OK; }
slv_result slv_host_unregister(slv_device, void*) { return SLV_OK; }
// single process, unsharded: the whole (single-sampled) frame
slv_result slv_texture_export_tiles_async(slv_device dev, slv_handle tex, void* host_frame, size_t bytes) {
  return slv_texture_readback(dev, tex, 0, host_frame, bytes);
}
// peer-memory frame assembly is a property of the CUDA product (NVLink); the CPU checkers do not implement it
slv_result slv_peer_export_texture(slv_device, slv_handle, uint32_t, uint8_t*) { return SLV_FAILED; }
slv_result slv_peer_export_flags(slv_device, uint8_t*) { return SLV_FAILED; }
slv_result slv_texture_level_tracking(slv_device, uint32_t) { return SLV_OK; }
slv_result slv_texture_levels_touched(slv_device, slv_handle, uint32_t* mask) { if (!mask) return SLV_INVALID_PARAMETER; *mask = 0; return SLV_OK; }
slv_result slv_shader_module_load(slv_device, uint32_t, const void*, size_t, uint32_t, slv_handle*) { return SLV_FAILED; }
slv_result slv_shader_compile_cubin(uint32_t, const char*, uint32_t, uint32_t, void**, size_t*, char*, size_t) { return SLV_FAILED; }
slv_result slv_shader_compile(slv_device, uint32_t, const char*, uint32_t, uint32_t, slv_handle*, char*, size_t) { return SLV_FAILED; }
void slv_free(void* p) { free(p); }
slv_result slv_sasl_translate(uint32_t, const char*, const char*, char**, size_t*, char*, size_t) { return SLV_FAILED; }
slv_result slv_peer_open(slv_device, const uint8_t*, void**) { return SLV_FAILED; }
slv_result slv_peer_close(slv_device, void*) { return SLV_FAILED; }
slv_result slv_resolve_target_peer(slv_device, slv_handle, void*) { return SLV_FAILED; }
slv_result slv_peer_signal(slv_device, void*, uint32_t, uint32_t) { return SLV_FAILED; }
slv_result slv_flags_wait(slv_device, const void*, uint32_t, uint32_t, uint32_t) { return SLV_FAILED; }

slv_result slv_texture_device_ptr(slv_device dev, slv_handle h, uint32_t level, void** out, size_t* bytes) {
  auto e = dev->get(h);
  if (!e || !e->tex) return SLV_INVALID_PARAMETER;
  auto s = e->tex->subresource(level);
  if (!s) return SLV_INVALID_PARAMETER;
  *out = s->texel_address(0, 0, 0);
  if (bytes) *bytes = s->pitch() * s->height();
  return SLV_OK;
}
slv_result slv_pack_tiles(slv_device, slv_handle, uint32_t, uint32_t, void*, size_t*) { return SLV_FAILED; }
slv_result slv_unpack_tiles(slv_device, slv_handle, uint32_t, uint32_t, const void*) { return SLV_FAILED; }

slv_result slv_sampler_probe(slv_device dev, slv_handle sh, uint32_t n, const float* coords, const float* ddx,
                             const float* ddy, const float* lod, uint32_t use_lod, float* out) {
  auto s = sampler_of(dev, sh);
  if (!s) return SLV_INVALID_PARAMETER;
  for (uint32_t i = 0; i < n; ++i) {
    color_rgba32f c;
    if (use_lod) {
      c = s->sample_2d_lod(vec2(coords[2 * i], coords[2 * i + 1]), lod[i]);
    } else {
      c = s->sample_2d_grad(vec2(coords[2 * i], coords[2 * i + 1]), vec2(ddx[2 * i], ddx[2 * i + 1]),
                            vec2(ddy[2 * i], ddy[2 * i + 1]), 0.0f);
    }
    out[4 * i + 0] = c.r;
    out[4 * i + 1] = c.g;
    out[4 * i + 2] = c.b;
    out[4 * i + 3] = c.a;
  }
  return SLV_OK;
}

}  // extern "C"

// salvia_b200_renderer.hpp — the reference's Direct3D-style host surface (salvia::core::renderer,
// salvia/include/salvia/core/renderer.h:42-131) over the flat C ABI of include/salvia_b200.h.
//
// Header-only, C++17, no dependency on the reference's headers: same method names, argument meaning, `result` codes and
// shared_ptr ownership, so application code written against salvia::core::renderer ports by changing the namespace.
// The library behind it is bound at run time with dlopen, so the SAME program drives the CUDA product
// (salviarenderer_b200/csrc/libsalvia_b200.so) or, in the tests, one of the CPU checkers that export the same table.
//
// What differs from the reference, and why:
//  * C++ shader objects cannot run on a GPU.  cpp_vertex_shader / cpp_pixel_shader / cpp_blend_shader here are HOST
//    descriptions of a device program: `device_program()` names it (SLV_VS_* / SLV_PS_* / SLV_BS_*, or SLV_PROGRAM_JIT(m) for
//    a SASL shader compiled at run time), constants are declared and set BY NAME with the reference's typed-by-name rules
//    (declare_constant / set_constant, salvia/include/salvia/shader/shader_utility.h:12-74: unknown name or wrong type ->
//    result::failed) and marshalled into the program's POD uniform block by `pack_uniforms`.
//  * map(surface, map_read) copies the texels out (the reference returns a copy too, surface.cpp:99-103);
//    map(..., map_write*) hands out a host staging copy that unmap() uploads.
//  * create_texcube and line / point topologies return result::failed / nullptr (the former is row f-4 of the scope table,
//    the latter are unimplemented upstream as well, rasterizer.cpp:1049-1058).
#pragma once

#include <dlfcn.h>
#include <unistd.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <typeindex>
#include <vector>

#include "salvia_b200.h"
#include "sasl_frontend.hpp"

namespace salvia_b200 {

// ---- enums with the reference's names and values (salvia/include/salvia/common/constants.h) ----
enum class result : uint32_t { ok, failed, out_of_memory, invalid_parameter };
enum class async_status : uint32_t { error, timeout, ready };
enum map_mode { map_mode_none = 0, map_read = 1, map_write = 2, map_read_write = 3, map_write_discard = 4, map_write_no_overwrite = 5 };
enum primitive_topology { primitive_line_list = 0, primitive_line_strip = 1, primitive_triangle_list = 2, primitive_triangle_fan = 3, primitive_triangle_strip = 4 };
enum cull_mode { cull_none = 0, cull_front = 1, cull_back = 2 };
enum address_mode { address_wrap = 0, address_mirror = 1, address_clamp = 2, address_border = 3 };
enum filter_type { filter_point = 0, filter_linear = 1, filter_anisotropic = 2 };
enum mip_quality { mip_lo_quality = 0, mip_mi_quality = 1, mip_hi_quality = 2 };
enum compare_function { compare_function_never = 0, compare_function_less, compare_function_equal, compare_function_less_equal,
                        compare_function_greater, compare_function_not_equal, compare_function_greater_equal, compare_function_always };
enum stencil_op { stencil_op_keep = 1, stencil_op_zero, stencil_op_replace, stencil_op_incr_sat, stencil_op_decr_sat, stencil_op_invert,
                  stencil_op_incr_wrap, stencil_op_decr_wrap };
enum clear_flag { clear_depth = 0x1, clear_stencil = 0x2 };
enum format { format_unknown = 0, format_r32g32b32a32_float = 2, format_r32g32b32_float = 6, format_r32g32_float = 16, format_r32_float = 41,
              format_r32_uint = 42, format_r16_uint = 57 };
enum pixel_format { pixel_format_color_rgba32f = 0, pixel_format_color_bgra8 = 2, pixel_format_color_rgba8 = 3, pixel_format_color_rg32f = 5 };
enum class async_object_ids : uint32_t { none, event, occlusion, pipeline_statistics, occlusion_predicate, internal_statistics, pipeline_profiles };
enum attrib_modifier { am_linear = 1, am_centroid = 2, am_nointerpolation = 4, am_noperspective = 8 };  // vs_output::attrib_modifier_type

struct vec4 { float x = 0, y = 0, z = 0, w = 0; };
struct mat44 { float m[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}}; };  // eflib::mat44: row-major data_[row][col]
struct color_rgba32f { float r = 0, g = 0, b = 0, a = 0; };
struct viewport { float x = 0, y = 0, w = 0, h = 0, minz = 0, maxz = 1; };
using sampler_desc = slv_sampler_desc;                 // field for field (sampler.h:16-45)
using depth_stencil_desc = slv_depth_stencil_desc;     // field for field (framebuffer.h:20-50)
struct raster_desc { cull_mode cm = cull_back; bool front_ccw = false; };  // the fields the rasterizer reads (raster_state.h:17-40)
struct pipeline_statistics { uint64_t ia_vertices, ia_primitives, vs_invocations, gs_invocations, gs_primitives, cinvocations, cprimitives, ps_invocations; };
struct internal_statistics { uint64_t backend_input_pixels; };
struct pipeline_profiles { uint64_t gather_vtx, vtx_proc, clipping, compact_clip, vp_trans, tri_dispatch, ras; };
struct mapped_resource { void* data = nullptr; uint32_t row_pitch = 0, depth_pitch = 0; };
struct input_element_desc {  // input_layout.h:20-52
  std::string semantic_name; uint32_t semantic_index = 0; format data_format = format_unknown; uint32_t input_slot = 0;
  uint32_t aligned_byte_offset = 0;
  input_element_desc() = default;
  input_element_desc(const char* n, uint32_t i, format f, uint32_t slot, uint32_t off) : semantic_name(n), semantic_index(i), data_format(f), input_slot(slot), aligned_byte_offset(off) {}
};

// ---- the C ABI, bound at run time ----
struct abi_table {
  void* lib = nullptr;
#define SLV_HOST_FN(name) decltype(&::name) name = nullptr;
  SLV_HOST_FN(slv_device_create) SLV_HOST_FN(slv_device_destroy) SLV_HOST_FN(slv_backend_name) SLV_HOST_FN(slv_buffer_create)
  SLV_HOST_FN(slv_buffer_upload) SLV_HOST_FN(slv_buffer_readback) SLV_HOST_FN(slv_texture_create) SLV_HOST_FN(slv_texture_gen_mipmap)
  SLV_HOST_FN(slv_texture_level_count) SLV_HOST_FN(slv_texture_level_size) SLV_HOST_FN(slv_texture_upload) SLV_HOST_FN(slv_texture_readback)
  SLV_HOST_FN(slv_sampler_create) SLV_HOST_FN(slv_resource_release) SLV_HOST_FN(slv_draw) SLV_HOST_FN(slv_clear_color)
  SLV_HOST_FN(slv_clear_depth_stencil) SLV_HOST_FN(slv_resolve) SLV_HOST_FN(slv_flush) SLV_HOST_FN(slv_query_begin) SLV_HOST_FN(slv_query_get)
  SLV_HOST_FN(slv_profile_get) SLV_HOST_FN(slv_shader_module_load) SLV_HOST_FN(slv_shader_compile) SLV_HOST_FN(slv_buffer_device_ptr)
#undef SLV_HOST_FN
  explicit abi_table(const std::string& path) {
    lib = dlopen(path.c_str(), RTLD_NOW | RTLD_LOCAL);
    if (!lib) throw std::runtime_error(std::string("cannot load ") + path + ": " + dlerror());
#define SLV_HOST_BIND(name)                                                         \
  name = reinterpret_cast<decltype(name)>(dlsym(lib, #name));                       \
  if (!name) throw std::runtime_error(std::string(#name " is not exported by ") + path);
    SLV_HOST_BIND(slv_device_create) SLV_HOST_BIND(slv_device_destroy) SLV_HOST_BIND(slv_backend_name) SLV_HOST_BIND(slv_buffer_create)
    SLV_HOST_BIND(slv_buffer_upload) SLV_HOST_BIND(slv_buffer_readback) SLV_HOST_BIND(slv_texture_create) SLV_HOST_BIND(slv_texture_gen_mipmap)
    SLV_HOST_BIND(slv_texture_level_count) SLV_HOST_BIND(slv_texture_level_size) SLV_HOST_BIND(slv_texture_upload) SLV_HOST_BIND(slv_texture_readback)
    SLV_HOST_BIND(slv_sampler_create) SLV_HOST_BIND(slv_resource_release) SLV_HOST_BIND(slv_draw) SLV_HOST_BIND(slv_clear_color)
    SLV_HOST_BIND(slv_clear_depth_stencil) SLV_HOST_BIND(slv_resolve) SLV_HOST_BIND(slv_flush) SLV_HOST_BIND(slv_query_begin) SLV_HOST_BIND(slv_query_get)
    SLV_HOST_BIND(slv_profile_get) SLV_HOST_BIND(slv_shader_module_load) SLV_HOST_BIND(slv_shader_compile) SLV_HOST_BIND(slv_buffer_device_ptr)
#undef SLV_HOST_BIND
  }
  ~abi_table() { if (lib) dlclose(lib); }
  abi_table(const abi_table&) = delete;
};
struct device_ctx {  // shared by the renderer and every resource it created (resources release themselves)
  std::shared_ptr<abi_table> abi;
  slv_device dev = nullptr;
  ~device_ctx() { if (dev) abi->slv_device_destroy(dev); }
};
using device_ctx_ptr = std::shared_ptr<device_ctx>;
inline result to_result(slv_result r) { return static_cast<result>(r); }

// ---- resources ----
class buffer {
public:
  buffer(device_ctx_ptr c, slv_handle h, size_t n) : ctx_(std::move(c)), handle_(h), size_(n), staging_(n) {}
  ~buffer() { ctx_->abi->slv_resource_release(ctx_->dev, handle_); }
  size_t size() const { return size_; }
  slv_handle handle() const { return handle_; }
  // buffer::transfer(offset, src, stride, count) (salvia/include/salvia/resource/buffer.h:16)
  result transfer(size_t offset, void const* src, size_t stride, size_t count) {
    if (offset + stride * count > size_) return result::invalid_parameter;
    std::memcpy(staging_.data() + offset, src, stride * count);
    return to_result(ctx_->abi->slv_buffer_upload(ctx_->dev, handle_, offset, staging_.data() + offset, stride * count));
  }
private:
  friend class renderer;
  device_ctx_ptr ctx_; slv_handle handle_; size_t size_; std::vector<uint8_t> staging_;
};
using buffer_ptr = std::shared_ptr<buffer>;

class texture;
class surface {  // mip level of a texture; level 0 is what set_render_targets takes
public:
  size_t width() const { return w_; }
  size_t height() const { return h_; }
  size_t sample_count() const { return samples_; }
  pixel_format get_pixel_format() const { return fmt_; }
  size_t texel_bytes() const { return fmt_ == pixel_format_color_rgba32f ? 16 : (fmt_ == pixel_format_color_rg32f ? 8 : 4); }
  size_t bytes() const { return w_ * h_ * samples_ * texel_bytes(); }
  slv_handle texture_handle() const { return tex_; }
  uint32_t level() const { return level_; }
  // surface::resolve (surface.cpp:123-140)
  result resolve(surface& target);
private:
  friend class texture; friend class renderer;
  device_ctx_ptr ctx_; slv_handle tex_ = 0; uint32_t level_ = 0; size_t w_ = 0, h_ = 0, samples_ = 1; pixel_format fmt_ = pixel_format_color_rgba8;
  std::vector<uint8_t> staging_;
};
using surface_ptr = std::shared_ptr<surface>;

class texture {
public:
  texture(device_ctx_ptr c, slv_handle h, size_t w, size_t hh, size_t s, pixel_format f) : ctx_(std::move(c)), handle_(h), w_(w), h_(hh), samples_(s), fmt_(f) { rebuild_levels(); }
  ~texture() { ctx_->abi->slv_resource_release(ctx_->dev, handle_); }
  slv_handle handle() const { return handle_; }
  size_t width(size_t level = 0) const { return levels_.at(level)->w_; }
  size_t height(size_t level = 0) const { return levels_.at(level)->h_; }
  size_t sample_count() const { return samples_; }
  pixel_format format() const { return fmt_; }
  size_t max_lod() const { return 0; }
  size_t min_lod() const { return levels_.size() - 1; }
  surface_ptr const& subresource(size_t level) const { return levels_.at(level); }
  // texture_2d::gen_mipmap(filter, auto_gen) (texture2d.cpp:25-36)
  void gen_mipmap(filter_type filter, bool /*auto_gen*/) {
    ctx_->abi->slv_texture_gen_mipmap(ctx_->dev, handle_, static_cast<uint32_t>(filter));
    rebuild_levels();
  }
private:
  void rebuild_levels() {
    uint32_t n = 1;
    ctx_->abi->slv_texture_level_count(ctx_->dev, handle_, &n);
    levels_.clear();
    for (uint32_t l = 0; l < n; ++l) {
      uint32_t w = 0, h = 0;
      ctx_->abi->slv_texture_level_size(ctx_->dev, handle_, l, &w, &h);
      auto s = std::make_shared<surface>();
      s->ctx_ = ctx_; s->tex_ = handle_; s->level_ = l; s->w_ = w; s->h_ = h; s->samples_ = samples_; s->fmt_ = fmt_;
      levels_.push_back(s);
    }
  }
  device_ctx_ptr ctx_; slv_handle handle_; size_t w_, h_, samples_; pixel_format fmt_; std::vector<surface_ptr> levels_;
};
using texture_ptr = std::shared_ptr<texture>;

inline result surface::resolve(surface& target) { return to_result(ctx_->abi->slv_resolve(ctx_->dev, tex_, target.tex_)); }

class sampler {  // holds its texture alive, as the reference's sampler does (sampler.h:61)
public:
  sampler(device_ctx_ptr c, slv_handle h, texture_ptr t) : ctx_(std::move(c)), handle_(h), tex_(std::move(t)) {}
  ~sampler() { ctx_->abi->slv_resource_release(ctx_->dev, handle_); }
  slv_handle handle() const { return handle_; }
private:
  device_ctx_ptr ctx_; slv_handle handle_; texture_ptr tex_;
};
using sampler_ptr = std::shared_ptr<sampler>;

class raster_state { public: explicit raster_state(raster_desc const& d) : desc_(d) {} raster_desc const& get_desc() const { return desc_; } private: raster_desc desc_; };
using raster_state_ptr = std::shared_ptr<raster_state>;
class depth_stencil_state { public: explicit depth_stencil_state(depth_stencil_desc const& d) : desc_(d) {} depth_stencil_desc const& get_desc() const { return desc_; } private: depth_stencil_desc desc_; };
using depth_stencil_state_ptr = std::shared_ptr<depth_stencil_state>;
inline depth_stencil_desc default_depth_stencil_desc() {  // depth_stencil_desc() (framebuffer.h:34-50)
  depth_stencil_desc d{};
  d.depth_enable = 1; d.depth_write_mask = 1; d.depth_func = compare_function_less;
  d.stencil_enable = 0; d.stencil_read_mask = 0xFF; d.stencil_write_mask = 0xFF;
  d.front_face = d.back_face = slv_stencil_op_desc{stencil_op_keep, stencil_op_keep, stencil_op_keep, compare_function_always};
  return d;
}

// ---- shaders: host descriptions of device programs, constants typed by name ----
class shader_constants {
public:
  template <class T> result declare_constant(std::string const& name, T& var) {
    table_.erase(name);
    table_.emplace(name, entry{&var, std::type_index(typeid(T)), sizeof(T)});
    return result::ok;
  }
  // set_constant (shader_utility.h:45-74): unknown name or a type other than the declared one fails
  template <class T> result set_constant(std::string const& name, T const* value) {
    auto it = table_.find(name);
    if (it == table_.end() || it->second.type != std::type_index(typeid(T))) return result::failed;
    std::memcpy(it->second.ptr, value, sizeof(T));
    return result::ok;
  }
  result set_constant_raw(std::string const& name, void const* value, size_t size) {  // renderer::set_vs_variable_value path
    auto it = table_.find(name);
    if (it == table_.end() || it->second.size != size) return result::failed;
    std::memcpy(it->second.ptr, value, size);
    return result::ok;
  }
  result declare_sampler(std::string const& name, sampler_ptr& slot) { samplers_[name] = &slot; return result::ok; }
  result set_sampler(std::string const& name, sampler_ptr const& s) {
    auto it = samplers_.find(name);
    if (it == samplers_.end()) return result::failed;
    *it->second = s;
    return result::ok;
  }
private:
  struct entry { void* ptr; std::type_index type; size_t size; };
  std::map<std::string, entry> table_;
  std::map<std::string, sampler_ptr*> samplers_;
};

class cpp_shader : public shader_constants {
public:
  virtual ~cpp_shader() = default;
  virtual uint32_t device_program() const = 0;                                    // SLV_VS_* / SLV_PS_* / SLV_BS_* / SLV_PROGRAM_JIT(m)
  virtual size_t pack_uniforms(uint8_t* /*dst*/, size_t /*cap*/) const { return 0; }  // the program's POD uniform block
  virtual void samplers(slv_handle (&/*out*/)[SLV_MAX_SAMPLERS]) const {}
  void bind(slv_shader_binding& b) const {
    std::memset(&b, 0, sizeof(b));
    b.program = device_program();
    b.uniform_bytes = static_cast<uint32_t>(pack_uniforms(b.uniforms, sizeof(b.uniforms)));
    samplers(b.samplers);
  }
};
class cpp_vertex_shader : public cpp_shader {
public:
  virtual uint32_t num_output_attributes() const = 0;
  virtual uint32_t output_attribute_modifiers(uint32_t /*index*/) const { return am_linear; }
  // bind_semantic(name, semantic index, register) (shader.h:159-165): which input register a layout element feeds
  result bind_semantic(char const* name, size_t semantic_index, size_t reg) { regs_[{name, semantic_index}] = reg; return result::ok; }
  bool find_register(std::string const& name, size_t semantic_index, size_t& reg) const {
    auto it = regs_.find({name, semantic_index});
    if (it == regs_.end()) return false;
    reg = it->second;
    return true;
  }
private:
  std::map<std::pair<std::string, size_t>, size_t> regs_;
};
class cpp_pixel_shader : public cpp_shader {};
class cpp_blend_shader : public cpp_shader {};
using cpp_vertex_shader_ptr = std::shared_ptr<cpp_vertex_shader>;
using cpp_pixel_shader_ptr = std::shared_ptr<cpp_pixel_shader>;
using cpp_blend_shader_ptr = std::shared_ptr<cpp_blend_shader>;

// The device twins of the samples' shaders (program ids and uniform layouts: include/salvia_b200.h).
class vs_mvp_passthrough : public cpp_vertex_shader {  // pos = in[0] * wvp; attribute i = in[src[i]]
public:
  mat44 wvp; std::vector<uint32_t> src;
  explicit vs_mvp_passthrough(std::vector<uint32_t> s) : src(std::move(s)) { declare_constant("wvpMatrix", wvp); bind_semantic("POSITION", 0, 0); }
  uint32_t device_program() const override { return SLV_VS_MVP_PASSTHROUGH; }
  uint32_t num_output_attributes() const override { return static_cast<uint32_t>(src.size()); }
  size_t pack_uniforms(uint8_t* dst, size_t) const override {
    slv_vs_mvp_passthrough_uniforms u{};
    std::memcpy(u.wvp, wvp.m, 64); u.n_attrs = static_cast<uint32_t>(src.size());
    for (size_t i = 0; i < src.size() && i < 5; ++i) u.src[i] = src[i];
    std::memcpy(dst, &u, sizeof(u)); return sizeof(u);
  }
};
class vs_lights3 : public cpp_vertex_shader {  // samples/ColorizedTriangle/ColorizedTriangle.cpp:29-53
public:
  mat44 wvp; vec4 light_pos[3];
  vs_lights3() {
    declare_constant("wvpMatrix", wvp); declare_constant("lightPos0", light_pos[0]); declare_constant("lightPos1", light_pos[1]); declare_constant("lightPos2", light_pos[2]);
    bind_semantic("POSITION", 0, 0); bind_semantic("NORMAL", 0, 1);
  }
  uint32_t device_program() const override { return SLV_VS_LIGHTS3; }
  uint32_t num_output_attributes() const override { return 4; }
  size_t pack_uniforms(uint8_t* dst, size_t) const override {
    slv_vs_lights3_uniforms u{}; std::memcpy(u.wvp, wvp.m, 64); std::memcpy(u.light_pos, light_pos, 48);
    std::memcpy(dst, &u, sizeof(u)); return sizeof(u);
  }
};
class vs_sponza : public cpp_vertex_shader {  // samples/Sponza/Sponza.cpp:64-97
public:
  mat44 wvp; vec4 light_pos, eye_pos;
  vs_sponza() {
    declare_constant("wvpMatrix", wvp); declare_constant("lightPos", light_pos); declare_constant("eyePos", eye_pos);
    bind_semantic("POSITION", 0, 0); bind_semantic("TEXCOORD", 0, 1); bind_semantic("NORMAL", 0, 2);
  }
  uint32_t device_program() const override { return SLV_VS_SPONZA; }
  uint32_t num_output_attributes() const override { return 4; }
  size_t pack_uniforms(uint8_t* dst, size_t) const override {
    slv_vs_sponza_uniforms u{}; std::memcpy(u.wvp, wvp.m, 64); std::memcpy(u.light_pos, &light_pos, 16); std::memcpy(u.eye_pos, &eye_pos, 16);
    std::memcpy(dst, &u, sizeof(u)); return sizeof(u);
  }
};
class vs_terrain_vtf : public cpp_vertex_shader {  // the SASL vertex shader of samples/VertexTextureFetch/VertexTextureFetch.cpp:38-61
public:
  mat44 wvp; float terrain_offset[2] = {0, 0}, terrain_scale[2] = {1, 1}; sampler_ptr sampler_;
  vs_terrain_vtf() {
    declare_constant("wvpMatrix", wvp); declare_constant("terrainOffset", terrain_offset); declare_constant("terrainScale", terrain_scale);
    declare_sampler("terrainSamp", sampler_);
    bind_semantic("POSITION", 0, 0); bind_semantic("TEXCOORD", 0, 1);
  }
  uint32_t device_program() const override { return SLV_VS_TERRAIN_VTF; }
  uint32_t num_output_attributes() const override { return 1; }
  size_t pack_uniforms(uint8_t* dst, size_t) const override {
    slv_vs_terrain_vtf_uniforms u{}; std::memcpy(u.wvp, wvp.m, 64); std::memcpy(u.offset, terrain_offset, 8); std::memcpy(u.scale, terrain_scale, 8);
    std::memcpy(dst, &u, sizeof(u)); return sizeof(u);
  }
  void samplers(slv_handle (&out)[SLV_MAX_SAMPLERS]) const override { out[0] = sampler_ ? sampler_->handle() : 0; }
};
class vs_ssm_draw : public cpp_vertex_shader {  // resources/ssm/Draw.savs (StandardShadowMap.cpp:192, colour pass)
public:
  mat44 camera_wvp, light_wvp; vec4 light_pos, camera_pos;
  vs_ssm_draw() {
    declare_constant("cameraWvp", camera_wvp); declare_constant("lightWvp", light_wvp);
    declare_constant("lightPos", light_pos); declare_constant("cameraPos", camera_pos);
    bind_semantic("POSITION", 0, 0); bind_semantic("NORMAL", 0, 1); bind_semantic("TEXCOORD", 0, 2);
  }
  uint32_t device_program() const override { return SLV_VS_SSM_DRAW; }
  uint32_t num_output_attributes() const override { return 5; }
  size_t pack_uniforms(uint8_t* dst, size_t) const override {
    slv_vs_ssm_draw_uniforms u{}; std::memcpy(u.camera_wvp, camera_wvp.m, 64); std::memcpy(u.light_wvp, light_wvp.m, 64);
    std::memcpy(u.light_pos, &light_pos, 16); std::memcpy(u.camera_pos, &camera_pos, 16);
    std::memcpy(dst, &u, sizeof(u)); return sizeof(u);
  }
};
class ps_lights3 : public cpp_pixel_shader { public: uint32_t device_program() const override { return SLV_PS_LIGHTS3; } };  // ColorizedTriangle.cpp:55-92
class ps_height_color : public cpp_pixel_shader { public: uint32_t device_program() const override { return SLV_PS_HEIGHT_COLOR; } };  // VertexTextureFetch.cpp:70-113
class ps_attr0_color : public cpp_pixel_shader { public: uint32_t device_program() const override { return SLV_PS_ATTR0_COLOR; } };
class ps_sponza : public cpp_pixel_shader {  // samples/Sponza/Sponza.cpp:99-146
public:
  vec4 ambient, diffuse, specular; int shininess = 0; sampler_ptr sampler_;
  ps_sponza() {
    declare_constant("Ambient", ambient); declare_constant("Diffuse", diffuse); declare_constant("Specular", specular);
    declare_constant("Shininess", shininess); declare_sampler("Sampler", sampler_);
  }
  uint32_t device_program() const override { return SLV_PS_SPONZA; }
  size_t pack_uniforms(uint8_t* dst, size_t) const override { slv_ps_sponza_uniforms u{sampler_ ? 1u : 0u}; std::memcpy(dst, &u, sizeof(u)); return sizeof(u); }
  void samplers(slv_handle (&out)[SLV_MAX_SAMPLERS]) const override { out[0] = sampler_ ? sampler_->handle() : 0; }
};
// tex2D(samp, uv) * saturate(dot(normalize(light), normalize(normal))): the cpp twin of the SASL pixel shader BASELINE configs[3]
// runs (sample_2d_grad with per-pixel derivatives; sasl_derivatives selects the per row / per column convention)
class ps_sponza_grad : public cpp_pixel_shader {
public:
  sampler_ptr sampler_; bool sasl_derivatives;
  explicit ps_sponza_grad(bool sasl_deriv = true) : sasl_derivatives(sasl_deriv) { declare_sampler("texSamp", sampler_); }
  uint32_t device_program() const override { return SLV_PS_SPONZA_GRAD; }
  size_t pack_uniforms(uint8_t* dst, size_t) const override { slv_ps_sponza_grad_uniforms u{sampler_ ? 1u : 0u, sasl_derivatives ? 1u : 0u}; std::memcpy(dst, &u, sizeof(u)); return sizeof(u); }
  void samplers(slv_handle (&out)[SLV_MAX_SAMPLERS]) const override { out[0] = sampler_ ? sampler_->handle() : 0; }
};
class ps_ssm_draw : public cpp_pixel_shader {  // draw_cpp_ps, samples/StandardShadowMap/StandardShadowMap.cpp:62-142 (two samplers)
public:
  vec4 ambient, diffuse, specular; int shininess = 0; sampler_ptr texsamp_, dsamp_;
  ps_ssm_draw() {
    declare_constant("Ambient", ambient); declare_constant("Diffuse", diffuse); declare_constant("Specular", specular);
    declare_constant("Shininess", shininess); declare_sampler("DepthSampler", dsamp_); declare_sampler("TexSampler", texsamp_);
  }
  uint32_t device_program() const override { return SLV_PS_SSM_DRAW; }
  size_t pack_uniforms(uint8_t* dst, size_t) const override {
    slv_ps_ssm_draw_uniforms u{};
    std::memcpy(u.ambient, &ambient, 16); std::memcpy(u.diffuse, &diffuse, 16); std::memcpy(u.specular, &specular, 16);
    u.shininess = shininess; u.has_tex_sampler = texsamp_ ? 1u : 0u; u.has_depth_sampler = dsamp_ ? 1u : 0u;
    std::memcpy(dst, &u, sizeof(u)); return sizeof(u);
  }
  void samplers(slv_handle (&out)[SLV_MAX_SAMPLERS]) const override {
    out[0] = texsamp_ ? texsamp_->handle() : 0; out[1] = dsamp_ ? dsamp_->handle() : 0;
  }
};
class ps_tex_alpha : public cpp_pixel_shader {  // samples/TextureAndBlending/TextureAndBlending.cpp:96-166
public:
  uint32_t reg; float alpha; sampler_ptr sampler_; bool grad;
  explicit ps_tex_alpha(uint32_t r, float a, bool sample_grad = false) : reg(r), alpha(a), grad(sample_grad) { declare_constant("Alpha", alpha); declare_sampler("Sampler", sampler_); }
  uint32_t device_program() const override { return grad ? SLV_PS_TEX_GRAD_ALPHA : SLV_PS_TEX_ALPHA; }
  size_t pack_uniforms(uint8_t* dst, size_t) const override { slv_ps_tex_alpha_uniforms u{reg, alpha}; std::memcpy(dst, &u, sizeof(u)); return sizeof(u); }
  void samplers(slv_handle (&out)[SLV_MAX_SAMPLERS]) const override { out[0] = sampler_ ? sampler_->handle() : 0; }
};
class bs_replace : public cpp_blend_shader { public: uint32_t device_program() const override { return SLV_BS_REPLACE; } };          // ColorizedTriangle.cpp:94-106
class bs_lerp_src_alpha : public cpp_blend_shader { public: uint32_t device_program() const override { return SLV_BS_LERP_SRC_ALPHA; } };  // TextureAndBlending.cpp:168-180
// ---- SASL (renderer.h:75-86,136-147; salvia/include/salvia/shader/shader_object.h) -----------------------------------------------
// compile(code, profile) runs the SASL front end IN PROCESS (sasl_frontend.hpp: lexer, parser, semantic analysis, reflection,
// code generation - the counterpart of the reference's sasl library) and returns a shader_object holding the reflection
// (uniform layout, sampler slots, input semantics) and the generated device code.  set_vertex_shader_code /
// set_pixel_shader_code hand that code to slv_shader_compile (NVRTC, in process) once per renderer and bind the module.
// SLV_SASL_FRONTEND=python runs the Python twin of the front end instead (`$SLV_SASL_PYTHON -m salviarenderer_b200.sasl.emit`,
// a child process; the two emit identical units, tests/test_sasl_frontend_cpp.py).
namespace shader {
enum languages { lang_none, lang_general, lang_vertex_shader, lang_pixel_shader, lang_blending_shader };
struct shader_profile { languages language = lang_none; };
class shader_object {
public:
  struct uniform { std::string type; size_t offset = 0, size = 0; };
  struct semantic_slot { std::string semantic; uint32_t index = 0, slot = 0; };
  languages language = lang_none;
  std::string device_code;
  uint32_t n_vs_output_attrs = 0;
  size_t uniform_bytes = 0;
  bool uses_derivatives = false;
  std::map<std::string, uniform> uniforms;
  std::vector<std::string> samplers;       // slot order
  std::vector<semantic_slot> inputs;       // VS: semantic -> input register; PS: semantic -> attribute
  std::vector<semantic_slot> outputs;
};
using shader_object_ptr = std::shared_ptr<shader_object>;
using shader_log_ptr = std::shared_ptr<std::string>;

inline shader_object_ptr compile(std::string const& code, shader_profile const& profile, shader_log_ptr& logs) {
  logs = std::make_shared<std::string>();
  if (profile.language != lang_vertex_shader && profile.language != lang_pixel_shader) { *logs = "only vertex and pixel shaders are compiled"; return nullptr; }
  std::string out;
  const char* which = getenv("SLV_SASL_FRONTEND");
  if (which && std::string(which) == "python") {
    // the Python front end reads the source from a file and writes the unit to stdout
    char src_path[] = "/tmp/slv_sasl_XXXXXX";
    int fd = mkstemp(src_path);
    if (fd < 0) { *logs = "cannot create a temporary file"; return nullptr; }
    FILE* sf = fdopen(fd, "w");
    fwrite(code.data(), 1, code.size(), sf);
    fclose(sf);
    const char* py = getenv("SLV_SASL_PYTHON");
    std::string cmd = std::string(py ? py : "python3") + " -m salviarenderer_b200.sasl.emit " + (profile.language == lang_vertex_shader ? "vs" : "ps") + " < " + src_path + " 2>&1";
    FILE* pf = popen(cmd.c_str(), "r");
    if (pf) {
      char buf[4096];
      size_t n;
      while ((n = fread(buf, 1, sizeof(buf), pf)) > 0) out.append(buf, n);
      pclose(pf);
    }
    unlink(src_path);
  } else {
    sasl::unit u;
    std::string error;
    if (!sasl::compile(code, profile.language == lang_vertex_shader ? "vs" : "ps", u, error)) { *logs = "error\n" + error + "\n"; return nullptr; }
    out = sasl::render(u);
  }
  if (out.compare(0, 10, "SLVSASL 1\n") != 0) { *logs = out.empty() ? "the SASL front end did not run (python3 -m salviarenderer_b200.sasl.emit)" : out; return nullptr; }
  auto obj = std::make_shared<shader_object>();
  obj->language = profile.language;
  size_t pos = 10;
  while (pos < out.size()) {
    size_t eol = out.find('\n', pos);
    if (eol == std::string::npos) eol = out.size();
    std::istringstream ln(out.substr(pos, eol - pos));
    pos = eol + 1;
    std::string key;
    ln >> key;
    if (key == "n_vs_output_attrs") ln >> obj->n_vs_output_attrs;
    else if (key == "uniform_bytes") ln >> obj->uniform_bytes;
    else if (key == "uses_derivatives") { int v = 0; ln >> v; obj->uses_derivatives = v != 0; }
    else if (key == "uniform") { std::string name; shader_object::uniform u; ln >> name >> u.type >> u.offset >> u.size; obj->uniforms[name] = u; }
    else if (key == "sampler") { size_t slot; std::string name; ln >> slot >> name; if (obj->samplers.size() <= slot) obj->samplers.resize(slot + 1); obj->samplers[slot] = name; }
    else if (key == "input" || key == "output") { shader_object::semantic_slot x; ln >> x.semantic >> x.index >> x.slot; (key == "input" ? obj->inputs : obj->outputs).push_back(x); }
    else if (key == "code") { size_t n = 0; ln >> n; obj->device_code = out.substr(pos, n); break; }
  }
  if (obj->device_code.empty()) { *logs = "malformed output of the SASL front end"; return nullptr; }
  return obj;
}
inline shader_object_ptr compile(std::string const& code, shader_profile const& profile) { shader_log_ptr l; return compile(code, profile, l); }
inline shader_object_ptr compile(std::string const& code, languages lang) { shader_profile p; p.language = lang; return compile(code, p); }
inline shader_object_ptr compile_from_file(std::string const& file_name, shader_profile const& profile, shader_log_ptr& logs) {
  std::ifstream f(file_name);
  if (!f) { logs = std::make_shared<std::string>("cannot open " + file_name); return nullptr; }
  std::stringstream ss;
  ss << f.rdbuf();
  return compile(ss.str(), profile, logs);
}
inline shader_object_ptr compile_from_file(std::string const& file_name, shader_profile const& profile) { shader_log_ptr l; return compile_from_file(file_name, profile, l); }
inline shader_object_ptr compile_from_file(std::string const& file_name, languages lang) { shader_profile p; p.language = lang; return compile_from_file(file_name, p); }
}  // namespace shader

// A SASL shader bound to a device: module handle (slv_shader_compile / slv_shader_module_load), globals set by name with the
// (offset, size) table of the compiler's reflection - by value (set_*_variable_value) or by pointer, read at every draw
// (set_vs_variable_pointer, renderer.h:79) - and samplers by name in the reflection's slot order.
class jit_uniform_block {
public:
  explicit jit_uniform_block(size_t uniform_bytes) : block_(uniform_bytes) {}
  void declare_uniform(std::string const& name, size_t offset, size_t size) { layout_[name] = {offset, size}; }
  result set_uniform(std::string const& name, void const* v, size_t size) {
    auto it = layout_.find(name);
    if (it == layout_.end() || it->second.second != size) return result::failed;
    pointers_.erase(name);
    std::memcpy(block_.data() + it->second.first, v, size);
    return result::ok;
  }
  result set_uniform_pointer(std::string const& name, void const* v, size_t size) {
    auto it = layout_.find(name);
    if (it == layout_.end() || it->second.second != size) return result::failed;
    pointers_[name] = v;
    return result::ok;
  }
  // array uniforms (`float4x4 bones[boneCount]`): the block holds the ADDRESS of a device buffer; the data are read through the
  // pointer given to set_*_variable_pointer at every draw (renderer::submit uploads them and patches the address in)
  struct array_binding { size_t offset = 0; void const* src = nullptr; size_t size = 0; std::shared_ptr<buffer> buf; bool uploaded = false; };
  void declare_array(std::string const& name, size_t offset) { arrays_[name].offset = offset; }
  result set_array_pointer(std::string const& name, void const* v, size_t size) {
    auto it = arrays_.find(name);
    if (it == arrays_.end() || !v || !size) return result::failed;
    it->second.src = v; it->second.size = size;
    return result::ok;
  }
  std::map<std::string, array_binding>& arrays() { return arrays_; }
  void patch_address(size_t offset, uint64_t address) { if (offset + 8 <= block_.size()) std::memcpy(block_.data() + offset, &address, 8); }
  size_t pack(uint8_t* dst, size_t cap) const {
    size_t n = block_.size() < cap ? block_.size() : cap;
    std::memcpy(dst, block_.data(), n);
    for (auto const& p : pointers_) {
      auto const& l = layout_.at(p.first);
      if (l.first + l.second <= n) std::memcpy(dst + l.first, p.second, l.second);
    }
    return n;
  }
  void declare_sampler_slot(std::string const& name, size_t slot) { sampler_slots_[name] = slot; if (samplers_.size() <= slot) samplers_.resize(slot + 1); }
  result set_sampler_by_name(std::string const& name, sampler_ptr const& s) {
    auto it = sampler_slots_.find(name);
    if (it == sampler_slots_.end()) return result::failed;
    samplers_[it->second] = s;
    return result::ok;
  }
  void sampler_handles(slv_handle (&out)[SLV_MAX_SAMPLERS]) const {
    for (size_t i = 0; i < samplers_.size() && i < SLV_MAX_SAMPLERS; ++i) out[i] = samplers_[i] ? samplers_[i]->handle() : 0;
  }
private:
  std::vector<uint8_t> block_; std::map<std::string, std::pair<size_t, size_t>> layout_; std::map<std::string, void const*> pointers_;
  std::map<std::string, size_t> sampler_slots_; std::vector<sampler_ptr> samplers_;
  std::map<std::string, array_binding> arrays_;
};
class jit_vertex_shader : public cpp_vertex_shader, public jit_uniform_block {
public:
  jit_vertex_shader(slv_handle module, uint32_t n_attrs, size_t uniform_bytes) : jit_uniform_block(uniform_bytes), module_(module), n_attrs_(n_attrs) {}
  void declare_sampler_name(std::string const& name) { declare_sampler_slot(name, 0); }  // the shader's `sampler` global (slot 0)
  void samplers(slv_handle (&out)[SLV_MAX_SAMPLERS]) const override { sampler_handles(out); }
  uint32_t device_program() const override { return SLV_PROGRAM_JIT(module_); }
  uint32_t num_output_attributes() const override { return n_attrs_; }
  size_t pack_uniforms(uint8_t* dst, size_t cap) const override { return pack(dst, cap); }
private:
  slv_handle module_; uint32_t n_attrs_;
};
class jit_pixel_shader : public cpp_pixel_shader, public jit_uniform_block {
public:
  jit_pixel_shader(slv_handle module, size_t uniform_bytes) : jit_uniform_block(uniform_bytes), module_(module) {}
  void samplers(slv_handle (&out)[SLV_MAX_SAMPLERS]) const override { sampler_handles(out); }
  uint32_t device_program() const override { return SLV_PROGRAM_JIT(module_); }
  size_t pack_uniforms(uint8_t* dst, size_t cap) const override { return pack(dst, cap); }
private:
  slv_handle module_;
};

// ---- input layout: input_element_descs resolved against the vertex shader's register map (stream_assembler.cpp:52-86) ----
class input_layout {
public:
  std::vector<input_element_desc> descs;
  std::vector<slv_input_element> elements;  // register <- (slot, offset, format, default w)
};
using input_layout_ptr = std::shared_ptr<input_layout>;

class async_object {
public:
  explicit async_object(async_object_ids id) : id_(id) {}
  async_object_ids id() const { return id_; }
private:
  friend class renderer;
  async_object_ids id_; bool begun_ = false, ended_ = false; slv_pipeline_statistics stats_{}; slv_pipeline_profiles profs_{};
};
using async_object_ptr = std::shared_ptr<async_object>;

// ---- the renderer ----
class renderer {
public:
  explicit renderer(std::string const& library_path, int device_ordinal = 0) {
    ctx_ = std::make_shared<device_ctx>();
    ctx_->abi = std::make_shared<abi_table>(library_path);
    if (ctx_->abi->slv_device_create(device_ordinal, &ctx_->dev) != SLV_OK) throw std::runtime_error("slv_device_create failed (no usable device; the CUDA product has no CPU fallback)");
    ds_state_ = std::make_shared<depth_stencil_state>(default_depth_stencil_desc());
    rs_state_ = std::make_shared<raster_state>(raster_desc{});
  }
  std::string backend_name() const { return ctx_->abi->slv_backend_name(); }

  // Creators
  buffer_ptr create_buffer(size_t size) {
    slv_handle h = 0;
    if (ctx_->abi->slv_buffer_create(ctx_->dev, size, &h) != SLV_OK) return nullptr;
    return std::make_shared<buffer>(ctx_, h, size);
  }
  texture_ptr create_tex2d(size_t width, size_t height, size_t num_samples, pixel_format fmt) {
    slv_handle h = 0;
    if (ctx_->abi->slv_texture_create(ctx_->dev, (uint32_t)width, (uint32_t)height, (uint32_t)num_samples, (uint32_t)fmt, &h) != SLV_OK) return nullptr;
    return std::make_shared<texture>(ctx_, h, width, height, num_samples, fmt);
  }
  texture_ptr create_texcube(size_t, size_t, size_t, pixel_format) { return nullptr; }  // scope row f-4
  sampler_ptr create_sampler(sampler_desc const& desc, texture_ptr const& tex) {
    slv_handle h = 0;
    if (!tex || ctx_->abi->slv_sampler_create(ctx_->dev, &desc, tex->handle(), &h) != SLV_OK) return nullptr;
    return std::make_shared<sampler>(ctx_, h, tex);
  }
  async_object_ptr create_query(async_object_ids id) {
    if (id != async_object_ids::pipeline_statistics && id != async_object_ids::internal_statistics && id != async_object_ids::pipeline_profiles) return nullptr;
    return std::make_shared<async_object>(id);
  }
  // renderer.h:52-54: the layout is resolved against the SASL shader's input semantics (the reflection's register map)
  input_layout_ptr create_input_layout(input_element_desc const* elem_descs, size_t elems_count, shader::shader_object_ptr const& code) {
    if (!code) return nullptr;
    auto probe = std::make_shared<jit_vertex_shader>(0, code->n_vs_output_attrs, 0);
    for (auto const& in : code->inputs) probe->bind_semantic(in.semantic.c_str(), in.index, in.slot);
    return create_input_layout(elem_descs, elems_count, cpp_vertex_shader_ptr(probe));
  }
  input_layout_ptr create_input_layout(input_element_desc const* elem_descs, size_t elems_count, cpp_vertex_shader_ptr const& vs) {
    auto l = std::make_shared<input_layout>();
    for (size_t i = 0; i < elems_count; ++i) {
      input_element_desc const& e = elem_descs[i];
      l->descs.push_back(e);
      size_t reg = 0;
      if (!vs || !vs->find_register(e.semantic_name, e.semantic_index, reg)) continue;  // elements the shader does not read
      slv_input_element el{};
      el.reg = (uint32_t)reg; el.format = (uint32_t)e.data_format; el.slot = e.input_slot; el.aligned_byte_offset = e.aligned_byte_offset;
      el.default_w = (e.semantic_name == "POSITION" || e.semantic_name == "SV_Position") ? 1.0f : 0.0f;  // shader/constants.h:119
      l->elements.push_back(el);
    }
    return l;
  }

  result map(mapped_resource& mapped, buffer_ptr const& buf, map_mode mm) {
    if (!buf || mapped_buffer_ || mapped_surface_) return result::failed;
    if (mm == map_read || mm == map_read_write)
      if (ctx_->abi->slv_buffer_readback(ctx_->dev, buf->handle_, 0, buf->staging_.data(), buf->size_) != SLV_OK) return result::failed;
    mapped.data = buf->staging_.data(); mapped.row_pitch = (uint32_t)buf->size_; mapped.depth_pitch = (uint32_t)buf->size_;
    mapped_buffer_ = buf; mapped_mode_ = mm;
    return result::ok;
  }
  result map(mapped_resource& mapped, surface_ptr const& surf, map_mode mm) {
    if (!surf || mapped_buffer_ || mapped_surface_) return result::failed;
    surf->staging_.resize(surf->bytes());
    if (mm == map_read || mm == map_read_write)
      if (ctx_->abi->slv_texture_readback(ctx_->dev, surf->tex_, surf->level_, surf->staging_.data(), surf->staging_.size()) != SLV_OK) return result::failed;
    mapped.data = surf->staging_.data();
    mapped.row_pitch = (uint32_t)(surf->w_ * surf->samples_ * surf->texel_bytes());
    mapped.depth_pitch = (uint32_t)surf->staging_.size();
    mapped_surface_ = surf; mapped_mode_ = mm;
    return result::ok;
  }
  result unmap() {
    result r = result::ok;
    bool writes = mapped_mode_ != map_read && mapped_mode_ != map_mode_none;
    if (mapped_buffer_ && writes) r = to_result(ctx_->abi->slv_buffer_upload(ctx_->dev, mapped_buffer_->handle_, 0, mapped_buffer_->staging_.data(), mapped_buffer_->size_));
    if (mapped_surface_ && writes) r = to_result(ctx_->abi->slv_texture_upload(ctx_->dev, mapped_surface_->tex_, mapped_surface_->level_, mapped_surface_->staging_.data(), mapped_surface_->staging_.size()));
    if (!mapped_buffer_ && !mapped_surface_) r = result::failed;
    mapped_buffer_.reset(); mapped_surface_.reset(); mapped_mode_ = map_mode_none;
    return r;
  }

  // State set
  result set_vertex_buffers(size_t starts_slot, size_t buffers_count, buffer_ptr const* buffers, size_t const* strides, size_t const* offsets) {
    if (starts_slot + buffers_count > 8) return result::invalid_parameter;
    for (size_t i = 0; i < buffers_count; ++i) streams_[starts_slot + i] = stream{buffers[i], strides[i], offsets[i]};
    return result::ok;
  }
  result set_index_buffer(buffer_ptr const& hbuf, format index_fmt) {
    if (index_fmt != format_r16_uint && index_fmt != format_r32_uint) return result::failed;  // renderer_impl.cpp:49-54
    index_buffer_ = hbuf; index_format_ = index_fmt;
    return result::ok;
  }
  result set_input_layout(input_layout_ptr const& layout) { layout_ = layout; return result::ok; }
  result set_vertex_shader(cpp_vertex_shader_ptr const& hvs) { vs_ = hvs; vs_code_.reset(); return result::ok; }
  result set_primitive_topology(primitive_topology primtopo) {  // renderer_impl.cpp:71-78
    if (primtopo != primitive_line_list && primtopo != primitive_line_strip && primtopo != primitive_triangle_list && primtopo != primitive_triangle_strip) return result::failed;
    topology_ = primtopo;
    return result::ok;
  }
  result set_vs_variable_value(std::string const& name, void const* pvariable, size_t sz) {
    if (auto j = dynamic_cast<jit_uniform_block*>(vs_.get())) return j->set_uniform(name, pvariable, sz);
    return vs_ ? vs_->set_constant_raw(name, pvariable, sz) : result::failed;
  }
  // the pointed-to value is read at every draw (renderer.h:79; vx_shader_unit::set_variable_pointer)
  result set_vs_variable_pointer(std::string const& name, void const* pvariable, size_t sz) {
    if (auto j = dynamic_cast<jit_uniform_block*>(vs_.get()))
      return j->arrays().count(name) ? j->set_array_pointer(name, pvariable, sz) : j->set_uniform_pointer(name, pvariable, sz);
    return result::failed;
  }
  // SASL shaders (renderer.h:75,86): the generated device code is compiled in process (slv_shader_compile: NVRTC) the first
  // time a shader_object is bound to this renderer; later binds reuse the module.  Fails on a library without run-time
  // compilation (the CPU checkers).
  result set_vertex_shader_code(shader::shader_object_ptr const& so) {
    if (!so || so->language != shader::lang_vertex_shader) return result::invalid_parameter;
    auto it = vs_code_cache_.find(so.get());
    if (it == vs_code_cache_.end()) {
      slv_handle module = 0;
      slv_result rc = ctx_->abi->slv_shader_compile(ctx_->dev, SLV_STAGE_VS, so->device_code.c_str(), so->n_vs_output_attrs, 0, &module, compile_log_, sizeof(compile_log_));
      if (rc != SLV_OK) return to_result(rc);
      auto sh = std::make_shared<jit_vertex_shader>(module, so->n_vs_output_attrs, so->uniform_bytes);
      for (auto const& u : so->uniforms) {
        if (u.second.type.size() > 2 && u.second.type.compare(u.second.type.size() - 2, 2, "[]") == 0) sh->declare_array(u.first, u.second.offset);
        else sh->declare_uniform(u.first, u.second.offset, u.second.size);
      }
      for (size_t i = 0; i < so->samplers.size(); ++i) sh->declare_sampler_slot(so->samplers[i], i);
      for (auto const& in : so->inputs) sh->bind_semantic(in.semantic.c_str(), in.index, in.slot);
      it = vs_code_cache_.emplace(so.get(), std::make_pair(so, sh)).first;
    }
    vs_ = it->second.second; vs_code_ = so;
    return result::ok;
  }
  result set_pixel_shader_code(shader::shader_object_ptr const& so, bool cpp_derivatives = false) {
    if (!so || so->language != shader::lang_pixel_shader) return result::invalid_parameter;
    auto it = ps_code_cache_.find(so.get());
    if (it == ps_code_cache_.end()) {
      slv_handle module = 0;
      slv_result rc = ctx_->abi->slv_shader_compile(ctx_->dev, SLV_STAGE_PS, so->device_code.c_str(), 0, cpp_derivatives ? SLV_COMPILE_DERIV_CPP : 0u, &module, compile_log_, sizeof(compile_log_));
      if (rc != SLV_OK) return to_result(rc);
      auto sh = std::make_shared<jit_pixel_shader>(module, so->uniform_bytes);
      for (auto const& u : so->uniforms) sh->declare_uniform(u.first, u.second.offset, u.second.size);
      for (size_t i = 0; i < so->samplers.size(); ++i) sh->declare_sampler_slot(so->samplers[i], i);
      it = ps_code_cache_.emplace(so.get(), std::make_pair(so, sh)).first;
    }
    ps_ = it->second.second; ps_code_ = so;
    return result::ok;
  }
  shader::shader_object_ptr get_vertex_shader_code() const { return vs_code_; }
  shader::shader_object_ptr get_pixel_shader_code() const { return ps_code_; }
  char const* shader_compile_log() const { return compile_log_; }
  template <typename T> result set_vs_variable(std::string const& name, T const* data) { return set_vs_variable_value(name, data, sizeof(T)); }
  result set_ps_variable(std::string const& name, void const* data, size_t sz) {
    if (auto j = dynamic_cast<jit_uniform_block*>(ps_.get())) return j->set_uniform(name, data, sz);
    return ps_ ? ps_->set_constant_raw(name, data, sz) : result::failed;
  }
  template <typename T> result set_ps_variable(std::string const& name, T const* data) { return set_ps_variable(name, static_cast<void const*>(data), sizeof(T)); }
  result set_ps_sampler(std::string const& name, sampler_ptr const& samp) {
    if (auto j = dynamic_cast<jit_uniform_block*>(ps_.get())) return j->set_sampler_by_name(name, samp);
    return ps_ ? ps_->set_sampler(name, samp) : result::failed;
  }
  result set_vs_sampler(std::string const& name, sampler_ptr const& samp) {  // vertex texture fetch (renderer.h:80)
    if (auto j = dynamic_cast<jit_uniform_block*>(vs_.get())) return j->set_sampler_by_name(name, samp);
    return vs_ ? vs_->set_sampler(name, samp) : result::failed;
  }
  result set_rasterizer_state(raster_state_ptr const& rs) { rs_state_ = rs; return result::ok; }
  result set_blend_shader(cpp_blend_shader_ptr const& hbs) { bs_ = hbs; return result::ok; }
  result set_pixel_shader(cpp_pixel_shader_ptr const& hps) { ps_ = hps; ps_code_.reset(); return result::ok; }
  result set_depth_stencil_state(depth_stencil_state_ptr const& dss, int32_t stencil_ref) { ds_state_ = dss; stencil_ref_ = stencil_ref; return result::ok; }
  result set_render_targets(size_t color_target_count, surface_ptr const* color_targets, surface_ptr const& ds_target) {
    if (color_target_count >= SLV_MAX_RENDER_TARGETS) return result::failed;  // renderer_impl.cpp:159-238
    std::vector<surface_ptr> c(color_targets, color_targets + color_target_count);
    for (auto const& s : c) {
      if (!s) continue;
      if (c[0] && s->samples_ != c[0]->samples_) return result::failed;
      if (ds_target && (ds_target->fmt_ != pixel_format_color_rg32f || ds_target->w_ < s->w_ || ds_target->h_ < s->h_ || ds_target->samples_ != s->samples_)) return result::failed;
    }
    color_targets_ = std::move(c); ds_target_ = ds_target;
    return result::ok;
  }
  result set_viewport(viewport const& vp) {  // renderer_impl.cpp:145-152
    if (vp.x < 0 || vp.y < 0 || vp.w >= SLV_MAX_RENDER_TARGET_SIZE || vp.h >= SLV_MAX_RENDER_TARGET_SIZE) return result::failed;
    vp_ = vp;
    return result::ok;
  }

  // State get
  buffer_ptr get_index_buffer() const { return index_buffer_; }
  format get_index_format() const { return index_format_; }
  primitive_topology get_primitive_topology() const { return topology_; }
  cpp_vertex_shader_ptr get_vertex_shader() const { return vs_; }
  raster_state_ptr get_rasterizer_state() const { return rs_state_; }
  cpp_pixel_shader_ptr get_pixel_shader() const { return ps_; }
  cpp_blend_shader_ptr get_blend_shader() const { return bs_; }
  viewport get_viewport() const { return vp_; }

  // render operations
  result begin(async_object_ptr const& q) {
    if (!q) return result::invalid_parameter;
    q->begun_ = true; q->ended_ = false;
    return to_result(ctx_->abi->slv_query_begin(ctx_->dev));
  }
  result end(async_object_ptr const& q) {
    if (!q || !q->begun_) return result::failed;
    slv_result r = ctx_->abi->slv_query_get(ctx_->dev, &q->stats_);
    if (r == SLV_OK && q->id_ == async_object_ids::pipeline_profiles) r = ctx_->abi->slv_profile_get(ctx_->dev, &q->profs_);
    q->ended_ = r == SLV_OK;
    return to_result(r);
  }
  async_status get_data(async_object_ptr const& q, void* data, bool /*do_not_wait*/) {
    if (!q || !q->ended_ || !data) return async_status::error;
    switch (q->id_) {
    case async_object_ids::pipeline_statistics: std::memcpy(data, &q->stats_, sizeof(pipeline_statistics)); break;
    case async_object_ids::internal_statistics: std::memcpy(data, &q->stats_.backend_input_pixels, sizeof(internal_statistics)); break;
    case async_object_ids::pipeline_profiles: std::memcpy(data, &q->profs_, sizeof(pipeline_profiles)); break;
    default: return async_status::error;
    }
    return async_status::ready;
  }
  result draw(size_t startpos, size_t primcnt) { return submit(startpos, primcnt, 0, false); }
  result draw_index(size_t startpos, size_t primcnt, int basevert) { return submit(startpos, primcnt, basevert, true); }
  result clear_color(surface_ptr const& color_target, color_rgba32f const& c) {
    if (!color_target) return result::invalid_parameter;
    return to_result(ctx_->abi->slv_clear_color(ctx_->dev, color_target->tex_, &c.r));
  }
  result clear_depth_stencil(surface_ptr const& depth_stencil_target, uint32_t f, float d, uint32_t s) {
    if (!depth_stencil_target) return result::invalid_parameter;
    return to_result(ctx_->abi->slv_clear_depth_stencil(ctx_->dev, depth_stencil_target->tex_, f, d, s));
  }
  result flush() { return to_result(ctx_->abi->slv_flush(ctx_->dev)); }

  // SASL: registers a shader compiled by salviarenderer_b200/sasl (the cubin image) with the device
  result load_shader_module(uint32_t stage, void const* image, size_t bytes, uint32_t n_vs_output_attrs, slv_handle& module) {
    return to_result(ctx_->abi->slv_shader_module_load(ctx_->dev, stage, image, bytes, n_vs_output_attrs, &module));
  }

private:
  struct stream { buffer_ptr buf; size_t stride = 0, offset = 0; };
  result submit(size_t startpos, size_t primcnt, int basevert, bool indexed) {
    if (!vs_ || !ps_ || !bs_ || !layout_) return result::failed;
    if (topology_ != primitive_triangle_list && topology_ != primitive_triangle_strip) return result::failed;  // lines: unimplemented upstream too
    if (indexed && !index_buffer_) return result::failed;
    slv_draw_desc d{};
    for (uint32_t slot = 0; slot < 8; ++slot)
      if (streams_[slot].buf) {
        d.n_streams = slot + 1;
        d.streams[slot] = slv_vertex_stream{streams_[slot].buf->handle_, (uint32_t)streams_[slot].stride, (uint32_t)streams_[slot].offset};
      }
    d.n_elements = (uint32_t)layout_->elements.size();
    for (uint32_t i = 0; i < d.n_elements && i < SLV_MAX_VS_INPUT_ATTRS; ++i) d.elements[i] = layout_->elements[i];
    d.index_buffer = indexed ? index_buffer_->handle_ : 0;
    d.index_format = indexed ? (uint32_t)index_format_ : (uint32_t)SLV_INDEX_NONE;
    d.topology = (uint32_t)topology_;
    d.start = (uint32_t)startpos; d.prim_count = (uint32_t)primcnt; d.base_vertex = basevert;
    if (auto j = dynamic_cast<jit_uniform_block*>(vs_.get()))  // array uniforms: push the pointed-to data, patch the buffer's address in
      for (auto& a : j->arrays()) {
        jit_uniform_block::array_binding& ab = a.second;
        if (!ab.src) return result::failed;  // declared by the shader, never set
        if (!ab.buf || ab.buf->size() < ab.size) { ab.buf = create_buffer(ab.size); ab.uploaded = false; }
        if (!ab.buf) return result::failed;
        if (!ab.uploaded || std::memcmp(ab.buf->staging_.data(), ab.src, ab.size) != 0) {  // an upload is a flush point: only when the data changed
          if (ab.buf->transfer(0, ab.src, ab.size, 1) != result::ok) return result::failed;
          ab.uploaded = true;
        }
        void* dptr = nullptr; size_t dbytes = 0;
        if (ctx_->abi->slv_buffer_device_ptr(ctx_->dev, ab.buf->handle_, &dptr, &dbytes) != SLV_OK) return result::failed;
        j->patch_address(ab.offset, reinterpret_cast<uint64_t>(dptr));
      }
    vs_->bind(d.vs); ps_->bind(d.ps); bs_->bind(d.bs);
    for (uint32_t i = 0; i < vs_->num_output_attributes() && i < SLV_MAX_VS_OUTPUT_ATTRS; ++i) d.vs_attr_modifiers[i] = vs_->output_attribute_modifiers(i);
    d.raster.cull_mode = (uint32_t)rs_state_->get_desc().cm; d.raster.front_ccw = rs_state_->get_desc().front_ccw ? 1u : 0u;
    d.ds = ds_state_->get_desc(); d.stencil_ref = stencil_ref_;
    d.viewport = slv_viewport{vp_.x, vp_.y, vp_.w, vp_.h, vp_.minz, vp_.maxz};
    d.n_color_targets = (uint32_t)color_targets_.size();
    for (size_t i = 0; i < color_targets_.size(); ++i) d.color_targets[i] = color_targets_[i] ? color_targets_[i]->tex_ : 0;
    d.ds_target = ds_target_ ? ds_target_->tex_ : 0;
    return to_result(ctx_->abi->slv_draw(ctx_->dev, &d));
  }

  device_ctx_ptr ctx_;
  stream streams_[8];
  buffer_ptr index_buffer_; format index_format_ = format_r16_uint;
  input_layout_ptr layout_; primitive_topology topology_ = primitive_triangle_list;
  cpp_vertex_shader_ptr vs_; cpp_pixel_shader_ptr ps_; cpp_blend_shader_ptr bs_;
  shader::shader_object_ptr vs_code_, ps_code_;
  std::map<shader::shader_object const*, std::pair<shader::shader_object_ptr, std::shared_ptr<jit_vertex_shader>>> vs_code_cache_;
  std::map<shader::shader_object const*, std::pair<shader::shader_object_ptr, std::shared_ptr<jit_pixel_shader>>> ps_code_cache_;
  char compile_log_[8192] = {0};
  raster_state_ptr rs_state_; depth_stencil_state_ptr ds_state_; int32_t stencil_ref_ = 0;
  std::vector<surface_ptr> color_targets_; surface_ptr ds_target_; viewport vp_;
  buffer_ptr mapped_buffer_; surface_ptr mapped_surface_; map_mode mapped_mode_ = map_mode_none;
};
using renderer_ptr = std::shared_ptr<renderer>;

// create_software_renderer() / create_benchmark_renderer() (renderer.h:133-134)
inline renderer_ptr create_b200_renderer(std::string const& library_path, int device_ordinal = 0) { return std::make_shared<renderer>(library_path, device_ordinal); }

}  // namespace salvia_b200

// oracle/b200_renderer.hpp — the reference-side binding of INTEGRATION.md §2, compiled for real.
//
// TEST INFRASTRUCTURE (built by oracle/Makefile target `bridge` against the headers of /root/reference; the product never sees
// it).  `salvia::core::b200_renderer` is the third subclass of the reference's renderer_impl next to sync_renderer and
// async_renderer (salvia/include/salvia/core/renderer_impl.h:23-112): every state setter, resource creator and map / unmap is
// the REFERENCE's own code (renderer_impl.cpp); only commit_state_and_command() (renderer_impl.h:111, renderer_impl.cpp:337-374)
// differs - it marshals the recorded render_state (render_state.h:46-94) into slv_draw_desc and calls the C ABI of
// include/salvia_b200.h in a library loaded at run time (the CUDA product on a GPU box, a CPU checker elsewhere).
//
// Host objects stay the reference's (resource::buffer, texture_2d, surface, sampler); the binding keeps a device handle beside
// each and pushes the host bytes when they changed (after unmap of a write mapping, on first use) - "unmap() pushes it with
// slv_buffer_upload / slv_texture_upload" in INTEGRATION.md.  map(surface, map_read) reads the surface back first.
#pragma once

#include <salvia/core/async_object.h>
#include <salvia/core/framebuffer.h>
#include <salvia/core/raster_state.h>
#include <salvia/core/render_state.h>
#include <salvia/core/renderer_impl.h>
#include <salvia/core/shader.h>
#include <salvia/resource/buffer.h>
#include <salvia/resource/input_layout.h>
#include <salvia/resource/mapped_resource.h>
#include <salvia/resource/sampler.h>
#include <salvia/resource/surface.h>
#include <salvia/resource/texture.h>
#include <salvia/shader/reflection.h>
#include <salvia/shader/shader_object.h>

#include <dlfcn.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "salvia_b200.h"

namespace salvia::core {

// A cpp shader that has a device twin advertises it: the program id, its POD uniform block as documented
// next to the enum in salvia_b200.h, and the samplers it holds, in slot order.
struct device_shader_info {
  virtual ~device_shader_info() = default;
  virtual uint32_t device_program() const = 0;
  virtual size_t pack_uniforms(uint8_t* dst, size_t cap) const = 0;
  virtual void device_samplers(resource::sampler_ptr (&out)[SLV_MAX_SAMPLERS]) const { (void)out; }
};

// compile(code, profile) for this renderer (renderer.h:136-147): what b200_renderer::compile returns.
// A shader_object of the reference's own interface (shader_object.h:20-32) - it is handed to set_vertex_shader_code /
// set_pixel_shader_code / create_input_layout unchanged - that carries the unit slv_sasl_translate produced: the reflection the
// binding marshals with (uniform block layout, sampler slots, input semantics -> registers) and the device code that
// slv_shader_compile turns into a module on first use.  Its shader_reflection answers what renderer_impl itself asks
// (pixel_shader_unit::initialize sizes its CPU-side buffers from total_size: none are needed here).
class b200_shader_object : public shader::shader_object, private shader::shader_reflection {
public:
  struct uniform { std::string type; size_t offset = 0, size = 0; };
  struct semantic_slot { std::string semantic; uint32_t index = 0, slot = 0; };
  shader::languages language = shader::lang_none;
  std::string device_code;
  uint32_t n_vs_output_attrs = 0;
  size_t uniform_bytes = 0;
  std::map<std::string, uniform> uniforms;
  std::vector<std::string> samplers;   // slot order
  std::vector<semantic_slot> inputs;   // VS: semantic -> input register; PS: semantic -> attribute
  mutable slv_handle module = 0;       // compiled for the device on first use (slv_shader_compile)

  shader::shader_reflection const* get_reflection() const override { return this; }
  void* native_function() const override { return nullptr; }

private:
  shader::languages get_language() const override { return language; }
  std::string_view entry_name() const override { return ""; }
  std::vector<shader::sv_layout*> layouts(shader::sv_usage) const override { return {}; }
  size_t layouts_count(shader::sv_usage) const override { return 0; }
  size_t total_size(shader::sv_usage) const override { return 0; }
  shader::sv_layout* input_sv_layout(shader::semantic_value const&) const override { return nullptr; }
  shader::sv_layout* input_sv_layout(std::string_view) const override { return nullptr; }
  shader::sv_layout* output_sv_layout(shader::semantic_value const&) const override { return nullptr; }
  bool has_position_output() const override { return true; }
};

class b200_renderer : public renderer_impl {
public:
  explicit b200_renderer(std::string const& library_path, int ordinal = 0) {
    lib_ = dlopen(library_path.c_str(), RTLD_NOW | RTLD_LOCAL);
    if (!lib_) throw std::runtime_error(std::string("cannot load ") + library_path + ": " + dlerror());
#define B200_BIND(name)                                              \
  name##_ = reinterpret_cast<decltype(&::name)>(dlsym(lib_, #name)); \
  if (!name##_) throw std::runtime_error(#name " is not exported by " + library_path);
    B200_BIND(slv_device_create) B200_BIND(slv_device_destroy) B200_BIND(slv_backend_name) B200_BIND(slv_buffer_create)
    B200_BIND(slv_buffer_upload) B200_BIND(slv_texture_create) B200_BIND(slv_texture_upload) B200_BIND(slv_texture_readback)
    B200_BIND(slv_sampler_create) B200_BIND(slv_draw) B200_BIND(slv_clear_color) B200_BIND(slv_clear_depth_stencil)
    B200_BIND(slv_resolve) B200_BIND(slv_flush) B200_BIND(slv_query_begin) B200_BIND(slv_query_get)
#undef B200_BIND
    if (slv_device_create_(ordinal, &dev_) != SLV_OK) throw std::runtime_error("slv_device_create failed (no usable device)");
  }
  ~b200_renderer() override {
    if (dev_) slv_device_destroy_(dev_);
    if (lib_) dlclose(lib_);
  }
  std::string backend_name() const { return slv_backend_name_(); }

  // salvia::core::compile(code, profile) for this renderer: the library's SASL front end (slv_sasl_translate), the unit parsed
  // into a shader_object.  Null + `log` when the source does not compile.
  shader::shader_object_ptr compile(std::string const& code, shader::languages lang, std::string* log = nullptr) {
    auto translate = reinterpret_cast<decltype(&::slv_sasl_translate)>(dlsym(lib_, "slv_sasl_translate"));
    auto release = reinterpret_cast<decltype(&::slv_free)>(dlsym(lib_, "slv_free"));
    if (!translate || !release || (lang != shader::lang_vertex_shader && lang != shader::lang_pixel_shader)) return nullptr;
    char* unit = nullptr;
    size_t bytes = 0;
    char msg[4096] = "";
    if (translate(lang == shader::lang_vertex_shader ? SLV_STAGE_VS : SLV_STAGE_PS, code.c_str(), nullptr, &unit, &bytes, msg, sizeof(msg)) != SLV_OK) {
      if (log) *log = msg;
      return nullptr;
    }
    const std::string text(unit, bytes);
    release(unit);
    auto obj = std::make_shared<b200_shader_object>();
    obj->language = lang;
    size_t pos = text.find('\n') + 1;  // after "SLVSASL 1"
    while (pos < text.size()) {
      size_t eol = text.find('\n', pos);
      if (eol == std::string::npos) eol = text.size();
      std::istringstream ln(text.substr(pos, eol - pos));
      pos = eol + 1;
      std::string key;
      ln >> key;
      if (key == "n_vs_output_attrs") ln >> obj->n_vs_output_attrs;
      else if (key == "uniform_bytes") ln >> obj->uniform_bytes;
      else if (key == "uniform") { std::string name; b200_shader_object::uniform u; ln >> name >> u.type >> u.offset >> u.size; obj->uniforms[name] = u; }
      else if (key == "sampler") { size_t slot; std::string name; ln >> slot >> name; if (obj->samplers.size() <= slot) obj->samplers.resize(slot + 1); obj->samplers[slot] = name; }
      else if (key == "input") { b200_shader_object::semantic_slot x; ln >> x.semantic >> x.index >> x.slot; obj->inputs.push_back(x); }
      else if (key == "code") { size_t n = 0; ln >> n; obj->device_code = text.substr(pos, n); break; }
    }
    if (obj->device_code.empty()) return nullptr;
    return obj;
  }

  result flush() override { return static_cast<result>(slv_flush_(dev_)); }

  // ---- resources: the reference's host objects, a device handle beside each
  resource::texture_ptr create_tex2d(size_t width, size_t height, size_t num_samples, pixel_format fmt) override {
    resource::texture_ptr t = renderer_impl::create_tex2d(width, height, num_samples, fmt);
    if (t) tex_of_surface_[t->subresource(0).get()] = t;
    return t;
  }
  resource::sampler_ptr create_sampler(resource::sampler_desc const& desc, resource::texture_ptr const& tex) override {
    resource::sampler_ptr s = renderer_impl::create_sampler(desc, tex);
    sampler_info_[s.get()] = {desc, tex, 0};  // the reference's sampler keeps both private (sampler.h:59-61)
    return s;
  }
  result map(resource::mapped_resource& m, resource::buffer_ptr const& buf, map_mode mm) override {
    mapped_buffer_ = buf; mapped_surface_.reset(); mapped_mode_ = mm;
    return renderer_impl::map(m, buf, mm);
  }
  result map(resource::mapped_resource& m, resource::surface_ptr const& surf, map_mode mm) override {
    if (mm == map_read || mm == map_read_write) {  // device -> the host copy the reference is about to hand out (surface.cpp:99-103)
      auto it = surfaces_.find(surf.get());
      if (it != surfaces_.end() && it->second.device_newer) {
        slv_flush_(dev_);
        const size_t bytes = surf->pitch() * surf->height();
        if (slv_texture_readback_(dev_, it->second.tex, it->second.level, surf->texel_address(0, 0, 0), bytes) != SLV_OK) return result::failed;
        it->second.device_newer = false;
      }
    }
    mapped_surface_ = surf; mapped_buffer_.reset(); mapped_mode_ = mm;
    return renderer_impl::map(m, surf, mm);
  }
  result unmap() override {
    result r = renderer_impl::unmap();
    const bool wrote = mapped_mode_ != map_read && mapped_mode_ != map_mode_none;
    if (wrote && mapped_buffer_) dirty_buffers_[mapped_buffer_.get()] = true;
    if (wrote && mapped_surface_) {
      auto it = surfaces_.find(mapped_surface_.get());
      if (it != surfaces_.end()) it->second.host_newer = true;
      else pending_surface_writes_[mapped_surface_.get()] = true;
    }
    mapped_buffer_.reset(); mapped_surface_.reset(); mapped_mode_ = map_mode_none;
    return r;
  }
  // surface::resolve call sites (sample_app.cpp:348-350, swap_chain_impl.cpp:42-46) call this instead
  result resolve(resource::surface_ptr const& src, resource::surface_ptr const& dst) {
    slv_handle hs = surface_handle(src), hd = surface_handle(dst);
    if (!hs || !hd) return result::failed;
    surfaces_[dst.get()].device_newer = true;
    return static_cast<result>(slv_resolve_(dev_, hs, hd));
  }

protected:
  result commit_state_and_command() override {
    render_state const& s = *state_;
    switch (s.cmd) {
    case command_id::clear_color: {
      slv_handle h = surface_handle(s.clear_color_target);
      if (!h) return result::failed;
      surfaces_[s.clear_color_target.get()].device_newer = true;
      const float c[4] = {s.clear_color.r, s.clear_color.g, s.clear_color.b, s.clear_color.a};
      return static_cast<result>(slv_clear_color_(dev_, h, c));
    }
    case command_id::clear_depth_stencil: {
      slv_handle h = surface_handle(s.clear_ds_target);
      if (!h) return result::failed;
      surfaces_[s.clear_ds_target.get()].device_newer = true;
      return static_cast<result>(slv_clear_depth_stencil_(dev_, h, s.clear_f, s.clear_z, s.clear_stencil));
    }
    case command_id::async_begin:  // render_core::async_start (render_core.cpp:113-118): zero the counters, then count
      s.current_async->start_counting();
      return static_cast<result>(slv_query_begin_(dev_));
    case command_id::async_end: {  // render_core::async_stop: the values land, then the pending write is released for get()
      result rq = collect_query(s.current_async);
      s.current_async->stop_counting();
      return rq;
    }
    case command_id::draw:
    case command_id::draw_index: break;
    }
    // a SASL shader object wins over the cpp shader of its stage, as in the reference (rasterizer.cpp:236, framebuffer.cpp:348)
    auto* vso = dynamic_cast<b200_shader_object const*>(s.vx_shader.get());
    auto* pso = dynamic_cast<b200_shader_object const*>(s.px_shader.get());
    auto* vsi = vso ? nullptr : dynamic_cast<device_shader_info*>(s.cpp_vs.get());
    auto* psi = pso ? nullptr : dynamic_cast<device_shader_info*>(s.cpp_ps.get());
    auto* bsi = dynamic_cast<device_shader_info*>(s.cpp_bs.get());
    if (!(vso || vsi) || !(pso || psi) || !bsi || !s.layout) return result::failed;  // a cpp shader without a device twin cannot run on the GPU
    slv_draw_desc d{};
    // input assembler: stream_state + input_layout resolved against the VS register map (stream_assembler.cpp:52-86)
    for (size_t slot = 0; slot < s.str_state.buffer_descs.size() && slot < 8; ++slot) {
      stream_buffer_desc const& b = s.str_state.buffer_descs[slot];
      if (!b.buf) continue;
      d.n_streams = static_cast<uint32_t>(slot + 1);
      d.streams[slot] = slv_vertex_stream{buffer_handle(b.buf), static_cast<uint32_t>(b.stride), static_cast<uint32_t>(b.offset)};
    }
    std::vector<std::pair<shader::semantic_value, uint32_t>> registers;  // semantic -> input register of the bound vertex shader
    if (vso) for (auto const& in : vso->inputs) registers.emplace_back(shader::semantic_value(in.semantic, in.index), in.slot);
    else for (auto const& sv_reg : s.cpp_vs->get_register_map()) registers.emplace_back(sv_reg.first, static_cast<uint32_t>(sv_reg.second));
    for (auto const& sv_reg : registers) {
      resource::input_element_desc const* e = s.layout->find_desc(sv_reg.first);
      if (!e) { d.n_elements = 0; break; }  // stream_assembler.cpp:61-64: one missing element drops them all
      if (d.n_elements >= SLV_MAX_VS_INPUT_ATTRS) return result::failed;
      slv_input_element& el = d.elements[d.n_elements++];
      el.reg = sv_reg.second;
      el.slot = e->input_slot;
      el.aligned_byte_offset = e->aligned_byte_offset;
      el.format = static_cast<uint32_t>(e->data_format);
      el.default_w = shader::semantic_value(e->semantic_name, e->semantic_index).default_w();
    }
    const bool indexed = s.cmd == command_id::draw_index;
    d.index_buffer = indexed ? buffer_handle(s.index_buffer) : 0;
    d.index_format = indexed ? static_cast<uint32_t>(s.index_format) : static_cast<uint32_t>(SLV_INDEX_NONE);
    d.topology = static_cast<uint32_t>(s.prim_topo);
    d.start = s.start_index; d.prim_count = s.prim_count; d.base_vertex = s.base_vertex;
    // shaders: program id + POD uniform block + sampler handles
    if (!(vso ? bind_sasl(d.vs, *vso, s.vx_cbuffer) : bind(d.vs, *vsi)) || !(pso ? bind_sasl(d.ps, *pso, s.px_cbuffer) : bind(d.ps, *psi)) || !bind(d.bs, *bsi))
      return result::failed;
    if (!vso)  // SASL outputs interpolate linearly with perspective (interp_shim.cpp:66-70: im_linear)
      for (uint32_t i = 0; i < s.cpp_vs->num_output_attributes() && i < SLV_MAX_VS_OUTPUT_ATTRS; ++i)
        d.vs_attr_modifiers[i] = s.cpp_vs->output_attribute_modifiers(i);
    // fixed-function state
    d.raster.cull_mode = static_cast<uint32_t>(s.ras_state->get_desc().cm);
    d.raster.front_ccw = s.ras_state->get_desc().front_ccw ? 1u : 0u;
    depth_stencil_desc const& ds = s.ds_state->get_desc();  // field for field, framebuffer.h:20-50
    d.ds.depth_enable = ds.depth_enable; d.ds.depth_write_mask = ds.depth_write_mask; d.ds.depth_func = static_cast<uint32_t>(ds.depth_func);
    d.ds.stencil_enable = ds.stencil_enable; d.ds.stencil_read_mask = ds.stencil_read_mask; d.ds.stencil_write_mask = ds.stencil_write_mask;
    d.ds.front_face = slv_stencil_op_desc{static_cast<uint32_t>(ds.front_face.stencil_fail_op), static_cast<uint32_t>(ds.front_face.stencil_depth_fail_op),
                                          static_cast<uint32_t>(ds.front_face.stencil_pass_op), static_cast<uint32_t>(ds.front_face.stencil_func)};
    d.ds.back_face = slv_stencil_op_desc{static_cast<uint32_t>(ds.back_face.stencil_fail_op), static_cast<uint32_t>(ds.back_face.stencil_depth_fail_op),
                                         static_cast<uint32_t>(ds.back_face.stencil_pass_op), static_cast<uint32_t>(ds.back_face.stencil_func)};
    d.stencil_ref = s.stencil_ref;
    d.viewport = slv_viewport{s.vp.x, s.vp.y, s.vp.w, s.vp.h, s.vp.minz, s.vp.maxz};
    d.n_color_targets = static_cast<uint32_t>(s.color_targets.size());
    for (size_t i = 0; i < s.color_targets.size() && i < 8; ++i) {
      d.color_targets[i] = surface_handle(s.color_targets[i]);
      if (s.color_targets[i]) surfaces_[s.color_targets[i].get()].device_newer = true;
    }
    d.ds_target = surface_handle(s.depth_stencil_target);
    if (s.depth_stencil_target) surfaces_[s.depth_stencil_target.get()].device_newer = true;
    return static_cast<result>(slv_draw_(dev_, &d));
  }

private:
  struct surface_entry { slv_handle tex = 0; uint32_t level = 0; bool host_newer = false, device_newer = false; };
  struct sampler_entry { resource::sampler_desc desc; resource::texture_ptr tex; slv_handle handle; };

  slv_handle buffer_handle(resource::buffer_ptr const& b) {
    if (!b) return 0;
    auto it = buffers_.find(b.get());
    if (it == buffers_.end()) {
      slv_handle h = 0;
      if (slv_buffer_create_(dev_, b->size(), &h) != SLV_OK) return 0;
      it = buffers_.emplace(b.get(), h).first;
      dirty_buffers_[b.get()] = true;
      keep_buffers_.push_back(b);
    }
    auto d = dirty_buffers_.find(b.get());
    if (d != dirty_buffers_.end() && d->second) {
      slv_buffer_upload_(dev_, it->second, 0, b->raw_data(0), b->size());
      d->second = false;
    }
    return it->second;
  }
  // every mip level of a texture is one surface object of the reference; level 0 names the texture
  slv_handle texture_handle(resource::texture_ptr const& t) {
    if (!t) return 0;
    resource::surface_ptr l0 = t->subresource(0);
    auto it = surfaces_.find(l0.get());
    if (it == surfaces_.end() || !it->second.tex) {
      slv_handle h = 0;
      if (slv_texture_create_(dev_, static_cast<uint32_t>(l0->width()), static_cast<uint32_t>(l0->height()), static_cast<uint32_t>(l0->sample_count()),
                              static_cast<uint32_t>(l0->get_pixel_format()), &h) != SLV_OK)
        return 0;
      surface_entry e; e.tex = h; e.level = 0; e.host_newer = pending_surface_writes_.count(l0.get()) != 0;
      surfaces_[l0.get()] = e;
      keep_textures_.push_back(t);
      it = surfaces_.find(l0.get());
    }
    return it->second.tex;
  }
  slv_handle surface_handle(resource::surface_ptr const& s) {
    if (!s) return 0;
    auto known = surfaces_.find(s.get());
    if (known == surfaces_.end() || !known->second.tex) {
      auto t = tex_of_surface_.find(s.get());
      if (t == tex_of_surface_.end()) return 0;
      if (!texture_handle(t->second)) return 0;
      known = surfaces_.find(s.get());
    }
    push_surface(s.get(), known->second);
    return known->second.tex;
  }
  void push_surface(resource::surface* s, surface_entry& e) {
    if (!e.host_newer) return;
    slv_texture_upload_(dev_, e.tex, e.level, s->texel_address(0, 0, 0), s->pitch() * s->height());
    e.host_newer = false;
  }
  // a sampled texture: all the mip levels the host object holds (gen_mipmap ran on the host: texture2d.cpp:25-36) are pushed;
  // the device chain is created by uploading level by level after slv_texture_gen_mipmap sized it
  slv_handle sampler_handle(resource::sampler_ptr const& sp) {
    if (!sp) return 0;
    auto it = sampler_info_.find(sp.get());
    if (it == sampler_info_.end()) return 0;
    sampler_entry& e = it->second;
    slv_handle th = texture_handle(e.tex);
    if (!th) return 0;
    surface_entry& l0 = surfaces_[e.tex->subresource(0).get()];
    const size_t levels = e.tex->min_lod() - e.tex->max_lod() + 1;  // texture.h:45-47: max_lod = 0 is the finest level, min_lod the coarsest
    if (!e.handle || l0.host_newer) {
      push_surface(e.tex->subresource(0).get(), l0);
      if (levels > 1) {
        auto gen = reinterpret_cast<decltype(&::slv_texture_gen_mipmap)>(dlsym(lib_, "slv_texture_gen_mipmap"));
        if (!gen || gen(dev_, th, SLV_FILTER_LINEAR) != SLV_OK) return 0;
        for (size_t l = 1; l < levels; ++l) {  // the host's own levels win (identical for filter_linear: surface.cpp:53-92)
          resource::surface_ptr sl = e.tex->subresource(l);
          slv_texture_upload_(dev_, th, static_cast<uint32_t>(l), sl->texel_address(0, 0, 0), sl->pitch() * sl->height());
        }
      }
    }
    if (!e.handle) {
      slv_sampler_desc sd{};
      static_assert(sizeof(slv_sampler_desc) == sizeof(resource::sampler_desc), "slv_sampler_desc mirrors sampler_desc (sampler.h:16-45)");
      std::memcpy(&sd, &e.desc, sizeof(sd));
      if (slv_sampler_create_(dev_, &sd, th, &e.handle) != SLV_OK) return 0;
    }
    return e.handle;
  }
  bool bind(slv_shader_binding& b, device_shader_info const& info) {
    std::memset(&b, 0, sizeof(b));
    b.program = info.device_program();
    b.uniform_bytes = static_cast<uint32_t>(info.pack_uniforms(b.uniforms, sizeof(b.uniforms)));
    resource::sampler_ptr sm[SLV_MAX_SAMPLERS];
    info.device_samplers(sm);
    for (int i = 0; i < SLV_MAX_SAMPLERS; ++i)
      if (sm[i] && !(b.samplers[i] = sampler_handle(sm[i]))) return false;
    return true;
  }
  // A SASL stage: the module (compiled on first use), the uniform block filled by NAME from the stage's shader_cbuffer
  // (renderer::set_vs_variable_value / set_ps_variable, renderer_impl.cpp:286-325) at the offsets of the unit's reflection, the
  // samplers by name in the reflection's slot order (set_vs_sampler / set_ps_sampler).
  bool bind_sasl(slv_shader_binding& b, b200_shader_object const& so, shader::shader_cbuffer const& cb) {
    std::memset(&b, 0, sizeof(b));
    if (!so.module) {
      auto compile_fn = reinterpret_cast<decltype(&::slv_shader_compile)>(dlsym(lib_, "slv_shader_compile"));
      char log[8192] = "";
      if (!compile_fn || compile_fn(dev_, so.language == shader::lang_vertex_shader ? SLV_STAGE_VS : SLV_STAGE_PS, so.device_code.c_str(), so.n_vs_output_attrs, 0,
                                    &so.module, log, sizeof(log)) != SLV_OK) {
        std::fprintf(stderr, "slv_shader_compile: %s\n", log);
        return false;
      }
    }
    b.program = SLV_PROGRAM_JIT(so.module);
    if (so.uniform_bytes > sizeof(b.uniforms)) return false;
    b.uniform_bytes = static_cast<uint32_t>(so.uniform_bytes);
    for (auto const& var : cb.variables()) {
      auto u = so.uniforms.find(var.first);
      if (u == so.uniforms.end() || u->second.type.find("[]") != std::string::npos) continue;  // not a global of this shader (arrays: not marshalled here)
      void const* src = cb.data_pointer(var.second);
      if (src) std::memcpy(b.uniforms + u->second.offset, src, std::min(var.second.length, u->second.size));
    }
    for (auto const& sm : cb.samplers())
      for (size_t slot = 0; slot < so.samplers.size() && slot < SLV_MAX_SAMPLERS; ++slot)
        if (so.samplers[slot] == sm.first && !(b.samplers[slot] = sampler_handle(sm.second))) return false;
    return true;
  }
  result collect_query(async_object_ptr const& q) {
    slv_pipeline_statistics st{};
    if (slv_query_get_(dev_, &st) != SLV_OK) return result::failed;
    if (q->id() == async_object_ids::pipeline_statistics) {
      using A = async_pipeline_statistics;
      A::accumulate<pipeline_statistic_id::ia_vertices>(q.get(), st.ia_vertices);
      A::accumulate<pipeline_statistic_id::ia_primitives>(q.get(), st.ia_primitives);
      A::accumulate<pipeline_statistic_id::vs_invocations>(q.get(), st.vs_invocations);
      A::accumulate<pipeline_statistic_id::cinvocations>(q.get(), st.cinvocations);
      A::accumulate<pipeline_statistic_id::cprimitives>(q.get(), st.cprimitives);
      A::accumulate<pipeline_statistic_id::ps_invocations>(q.get(), st.ps_invocations);
    } else if (q->id() == async_object_ids::internal_statistics) {
      async_internal_statistics::accumulate<internal_statistics_id::backend_input_pixels>(q.get(), st.backend_input_pixels);
    }
    return result::ok;
  }

  void* lib_ = nullptr;
  slv_device dev_ = nullptr;
#define B200_FN(name) decltype(&::name) name##_ = nullptr;
  B200_FN(slv_device_create) B200_FN(slv_device_destroy) B200_FN(slv_backend_name) B200_FN(slv_buffer_create) B200_FN(slv_buffer_upload)
  B200_FN(slv_texture_create) B200_FN(slv_texture_upload) B200_FN(slv_texture_readback) B200_FN(slv_sampler_create) B200_FN(slv_draw)
  B200_FN(slv_clear_color) B200_FN(slv_clear_depth_stencil) B200_FN(slv_resolve) B200_FN(slv_flush) B200_FN(slv_query_begin) B200_FN(slv_query_get)
#undef B200_FN
  std::unordered_map<resource::buffer*, slv_handle> buffers_;
  std::unordered_map<resource::buffer*, bool> dirty_buffers_;
  std::unordered_map<resource::surface*, surface_entry> surfaces_;
  std::unordered_map<resource::surface*, bool> pending_surface_writes_;
  std::unordered_map<resource::surface*, resource::texture_ptr> tex_of_surface_;
  std::unordered_map<resource::sampler*, sampler_entry> sampler_info_;
  std::vector<resource::buffer_ptr> keep_buffers_;
  std::vector<resource::texture_ptr> keep_textures_;
  resource::buffer_ptr mapped_buffer_;
  resource::surface_ptr mapped_surface_;
  map_mode mapped_mode_ = map_mode_none;
};

}  // namespace salvia::core

// A whole sample application in C++ over the host surface, the way samples/Sponza/Sponza.cpp is written: load an OBJ + MTL
// (salvia_b200_assets.hpp), load each material's map_Kd PNG into a mip-mapped texture, draw one batch per material with the
// Sponza shader twins, save the frame as a PNG and the counters as <name>_Profiling.json.  tests/test_assets_cpp.py runs it on a
// CPU checker library and compares the PNG with the same scene rendered through the Python path.
//   usage: obj_viewer_test <library.so> <scene.obj> <uniforms.bin: wvp[16] light[4] eye[4] floats> <w> <h> <out.png> <report dir>
#include <cinttypes>
#include <cstdio>
#include <cstdlib>

#include "salvia_b200_assets.hpp"
#include "salvia_b200_renderer.hpp"

using namespace salvia_b200;
#define CHECK(e) do { if ((e) != result::ok) { std::fprintf(stderr, "FAILED: %s (line %d)\n", #e, __LINE__); return 2; } } while (0)

int main(int argc, char** argv) {
  if (argc < 8) return 64;
  const size_t W = std::atoi(argv[4]), H = std::atoi(argv[5]);
  renderer_ptr r = create_b200_renderer(argv[1]);

  assets::obj_mesh mesh;
  if (!assets::load_obj(argv[2], false, mesh)) { std::fprintf(stderr, "cannot load %s\n", argv[2]); return 1; }
  float uniforms[24];
  {
    std::ifstream f(argv[3], std::ios::binary);
    if (!f.read(reinterpret_cast<char*>(uniforms), sizeof(uniforms))) return 1;
  }

  // targets, one shared 48-byte-stride vertex buffer (mesh_io_obj.cpp:389-437)
  texture_ptr color = r->create_tex2d(W, H, 1, pixel_format_color_rgba8), ds = r->create_tex2d(W, H, 1, pixel_format_color_rg32f);
  surface_ptr cs = color->subresource(0), dss = ds->subresource(0);
  CHECK(r->set_render_targets(1, &cs, dss));
  CHECK(r->set_viewport(viewport{0, 0, (float)W, (float)H, 0.0f, 1.0f}));
  buffer_ptr vb = r->create_buffer(mesh.vertices.size() * sizeof(float));
  CHECK(vb->transfer(0, mesh.vertices.data(), 48, mesh.vertex_count()));
  auto vs = std::make_shared<vs_sponza>();
  auto ps = std::make_shared<ps_sponza>();
  input_element_desc descs[] = {{"POSITION", 0, format_r32g32b32a32_float, 0, 0}, {"TEXCOORD", 0, format_r32g32b32a32_float, 0, 16},
                                {"NORMAL", 0, format_r32g32b32a32_float, 0, 32}};
  CHECK(r->set_input_layout(r->create_input_layout(descs, 3, vs)));
  buffer_ptr bufs[1] = {vb};
  size_t strides[1] = {48}, offsets[1] = {0};
  CHECK(r->set_vertex_buffers(0, 1, bufs, strides, offsets));
  CHECK(r->set_primitive_topology(primitive_triangle_list));
  CHECK(r->set_vertex_shader(vs));
  CHECK(r->set_pixel_shader(ps));
  CHECK(r->set_blend_shader(std::make_shared<bs_replace>()));
  CHECK(r->set_rasterizer_state(std::make_shared<raster_state>(raster_desc{cull_back, false})));
  mat44 wvp;
  std::memcpy(wvp.m, uniforms, 64);
  vec4 light{uniforms[16], uniforms[17], uniforms[18], uniforms[19]}, eye{uniforms[20], uniforms[21], uniforms[22], uniforms[23]};
  CHECK(r->set_vs_variable("wvpMatrix", &wvp));
  CHECK(r->set_vs_variable("lightPos", &light));
  CHECK(r->set_vs_variable("eyePos", &eye));

  // one texture + trilinear wrap sampler per material that names a map_Kd (Sponza.cpp:265-273)
  std::vector<sampler_ptr> samplers(mesh.materials.size());
  for (size_t m = 0; m < mesh.materials.size(); ++m) {
    if (mesh.materials[m].tex_path.empty()) continue;
    uint32_t tw = 0, th = 0;
    std::vector<uint8_t> texels;
    std::string err;
    if (!assets::load_texture_rgba8(mesh.materials[m].tex_path, tw, th, texels, &err)) { std::fprintf(stderr, "%s\n", err.c_str()); return 1; }
    texture_ptr tex = r->create_tex2d(tw, th, 1, pixel_format_color_rgba8);
    mapped_resource mr;
    CHECK(r->map(mr, tex->subresource(0), map_write));
    std::memcpy(mr.data, texels.data(), texels.size());
    CHECK(r->unmap());
    tex->gen_mipmap(filter_linear, true);
    sampler_desc sd{};
    sd.min_filter = sd.mag_filter = sd.mip_filter = filter_linear;
    sd.mip_qual = mip_mi_quality;
    sd.addr_mode_u = sd.addr_mode_v = sd.addr_mode_w = address_wrap;
    sd.comparison_func = compare_function_always;
    sd.min_lod = -1e20f; sd.max_lod = 1e20f;
    samplers[m] = r->create_sampler(sd, tex);
    if (!samplers[m]) return 3;
  }

  async_object_ptr q = r->create_query(async_object_ids::pipeline_statistics);
  CHECK(r->begin(q));
  CHECK(r->clear_color(cs, color_rgba32f{0.2f, 0.2f, 0.5f, 1.0f}));
  CHECK(r->clear_depth_stencil(dss, clear_depth | clear_stencil, 1.0f, 0));
  for (const auto& g : mesh.material_groups()) {  // one index buffer + one draw per material (mesh_impl::render)
    buffer_ptr ib = r->create_buffer(g.second.size() * 4);
    CHECK(ib->transfer(0, g.second.data(), 4, g.second.size()));
    CHECK(r->set_index_buffer(ib, format_r32_uint));
    CHECK(r->set_ps_sampler("Sampler", samplers[g.first]));
    CHECK(r->draw_index(0, g.second.size() / 3, 0));
  }
  CHECK(r->end(q));
  pipeline_statistics st{};
  if (r->get_data(q, &st, false) != async_status::ready) return 3;
  CHECK(r->flush());

  mapped_resource m;
  CHECK(r->map(m, cs, map_read));
  const bool saved = assets::save_surface_png(argv[6], static_cast<const uint8_t*>(m.data), (uint32_t)W, (uint32_t)H, false);
  CHECK(r->unmap());
  if (!saved) return 4;
  assets::profiling_frame fr = {{"cinvocations", (long long)st.cinvocations}, {"cprimitives", (long long)st.cprimitives}, {"ia_primitives", (long long)st.ia_primitives},
                                {"ia_vertices", (long long)st.ia_vertices}, {"vs_invocations", (long long)st.vs_invocations}, {"ps_invocations", (long long)st.ps_invocations}};
  std::printf("%s\n", assets::save_profiling_json("ObjViewer", r->backend_name(), {fr}, argv[7]).c_str());
  std::printf("stats ia_primitives %" PRIu64 " cprimitives %" PRIu64 " ps_invocations %" PRIu64 "\n", st.ia_primitives, st.cprimitives, st.ps_invocations);
  return 0;
}

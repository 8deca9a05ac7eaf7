// oracle/ref_obj_dump.cpp — pins salviarenderer_b200/assets.py's OBJ + MTL loader to the reference's own loader.
// TEST INFRASTRUCTURE: compiled in place from /root/reference by `make -C oracle obj-dump` (salvia/src/ext/resource/mesh/
// {mesh_io_obj,mesh_impl,material}.cpp + the core objects; the texture loader the materials call needs FreeImage, which is not
// in this image, so load_texture is a stub that records the requested path and returns no texture).
//   usage: ref_obj_dump <file.obj> <flip_tex_v 0|1>
// Prints, per mesh (= per material group, create_mesh_from_obj's order): primitive count, vertex stride, FNV-1a of the vertex
// buffer and of the index buffer, the material's name / ambient / diffuse / specular / shininess / texture name, and the texture paths the loader asked for.
#include <salvia/ext/resource/mesh/material.h>
#include <salvia/ext/resource/mesh/mesh.h>
#include <salvia/ext/resource/mesh/mesh_io_obj.h>
#include <salvia/ext/resource/texture/tex_io.h>

#include <salvia/core/renderer.h>
#include <salvia/resource/buffer.h>

#include <cinttypes>
#include <cstdio>
#include <map>
#include <string>

using namespace salvia;
using namespace salvia::core;
using namespace salvia::resource;
using namespace salvia::ext::resource;

static std::vector<std::string> g_requested;
namespace salvia::ext::resource {
texture_ptr load_texture(renderer*, const std::string& filename, pixel_format) {
  g_requested.push_back(filename);
  return texture_ptr();
}
}  // namespace salvia::ext::resource

static uint64_t fnv(const void* p, size_t n) {
  const uint8_t* b = static_cast<const uint8_t*>(p);
  uint64_t h = 1469598103934665603ull;
  for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
  return h;
}

int main(int argc, char** argv) {
  if (argc < 3) return 1;
  renderer_ptr r = create_benchmark_renderer();
  std::vector<mesh_ptr> meshes = create_mesh_from_obj(r.get(), argv[1], argv[2][0] == '1');
  std::printf("meshes %zu\n", meshes.size());
  for (size_t i = 0; i < meshes.size(); ++i) {
    mesh_ptr const& m = meshes[i];
    buffer_ptr vb = m->get_vertex_buffer(0), ib = m->get_index_buffer();
    auto mtl = std::dynamic_pointer_cast<obj_material>(m->get_attached());
    std::printf("mesh %zu prims %zu vb_bytes %zu vb %016" PRIx64 " ib_bytes %zu ib %016" PRIx64, i, m->get_face_count(), vb ? vb->size() : 0,
                vb ? fnv(vb->raw_data(0), vb->size()) : 0, ib ? ib->size() : 0, ib ? fnv(ib->raw_data(0), ib->size()) : 0);
    if (mtl)
      std::printf(" name %s ambient %.9g %.9g %.9g %.9g diffuse %.9g %.9g %.9g %.9g specular %.9g %.9g %.9g %.9g shininess %d tex_name %s", mtl->name.c_str(),
                  mtl->ambient[0], mtl->ambient[1], mtl->ambient[2], mtl->ambient[3], mtl->diffuse[0], mtl->diffuse[1], mtl->diffuse[2], mtl->diffuse[3],
                  mtl->specular[0], mtl->specular[1], mtl->specular[2], mtl->specular[3], mtl->shininess, mtl->tex_name.c_str());
    std::printf("\n");
  }
  for (auto const& s : g_requested) std::printf("texture %s\n", s.c_str());
  return 0;
}

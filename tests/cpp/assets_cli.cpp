// Drives salviarenderer_b200/host/salvia_b200_assets.hpp for tests/test_assets_cpp.py.
//   assets_cli obj PATH FLIP            fingerprint lines of the meshes (the format of oracle/ref_obj_dump.cpp) + every material
//   assets_cli tex PATH                 "W H FNV" of the rgba8 texels (bottom-up rows)
//   assets_cli png W H BGRA RAW OUT     writes the raw 4-byte texels of RAW as a PNG
//   assets_cli prof COMPILER FRAMES DIR NAME   FRAMES: one line per frame, "key=value key=value ..."; prints the path written
#include <cinttypes>
#include <cstdio>
#include <iostream>

#include "salvia_b200_assets.hpp"

using namespace salvia_b200::assets;

static uint64_t fnv(const void* p, size_t n) {
  uint64_t h = 1469598103934665603ull;
  for (size_t i = 0; i < n; ++i) h = (h ^ static_cast<const uint8_t*>(p)[i]) * 1099511628211ull;
  return h;
}
static std::string vec(const float* v) {
  char b[128];
  std::snprintf(b, sizeof(b), "%.9g %.9g %.9g %.9g", v[0], v[1], v[2], v[3]);
  return b;
}

int main(int argc, char** argv) {
  if (argc < 3) return 64;
  const std::string mode = argv[1];
  if (mode == "obj" && argc >= 4) {
    obj_mesh m;
    if (!load_obj(argv[2], std::string(argv[3]) == "1", m)) return 1;
    const auto groups = m.material_groups();
    std::printf("meshes %zu\n", groups.size());
    for (size_t i = 0; i < groups.size(); ++i) {
      const obj_material& mt = m.materials[groups[i].first];
      const auto& idx = groups[i].second;
      std::printf("mesh %zu prims %zu vb_bytes %zu vb %016" PRIx64 " ib_bytes %zu ib %016" PRIx64 " name %s ambient %s diffuse %s specular %s shininess %d tex_name %s\n",
                  i, idx.size() / 3, m.vertices.size() * 4, fnv(m.vertices.data(), m.vertices.size() * 4), idx.size() * 4, fnv(idx.data(), idx.size() * 4),
                  mt.name.c_str(), vec(mt.ambient).c_str(), vec(mt.diffuse).c_str(), vec(mt.specular).c_str(), mt.shininess, mt.tex_name.c_str());
    }
    for (const auto& mt : m.materials)
      std::printf("material %s ambient %s diffuse %s specular %s alpha %.9g shininess %d is_specular %d tex_name %s tex_path %s\n", mt.name.c_str(),
                  vec(mt.ambient).c_str(), vec(mt.diffuse).c_str(), vec(mt.specular).c_str(), mt.alpha, mt.shininess, mt.is_specular ? 1 : 0, mt.tex_name.c_str(),
                  mt.tex_path.c_str());
    std::printf("indices %016" PRIx64 " attrs %016" PRIx64 "\n", fnv(m.indices.data(), m.indices.size() * 4), fnv(m.attrs.data(), m.attrs.size() * 4));
    return 0;
  }
  if (mode == "tex") {
    uint32_t w = 0, h = 0;
    std::vector<uint8_t> t;
    std::string err;
    if (!load_texture_rgba8(argv[2], w, h, t, &err)) { std::printf("error %s\n", err.c_str()); return 2; }
    std::printf("%u %u %016" PRIx64 "\n", w, h, fnv(t.data(), t.size()));
    return 0;
  }
  if (mode == "png" && argc >= 7) {
    const uint32_t w = std::atoi(argv[2]), h = std::atoi(argv[3]);
    std::ifstream f(argv[5], std::ios::binary);
    std::vector<uint8_t> raw((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    if (raw.size() != size_t(w) * h * 4) return 1;
    return save_surface_png(argv[6], raw.data(), w, h, std::string(argv[4]) == "1") ? 0 : 1;
  }
  if (mode == "prof" && argc >= 6) {
    std::ifstream f(argv[3]);
    std::vector<profiling_frame> frames;
    std::string line;
    while (std::getline(f, line)) {
      profiling_frame fr;
      std::istringstream ss(line);
      std::string kv;
      while (ss >> kv) { const size_t eq = kv.find('='); fr[kv.substr(0, eq)] = std::atoll(kv.c_str() + eq + 1); }
      frames.push_back(fr);
    }
    std::printf("%s\n", save_profiling_json(argv[5], argv[2], frames, argv[4]).c_str());
    return 0;
  }
  return 64;
}

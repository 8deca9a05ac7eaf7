// Asset and report formats either side of the draw path, for C++ hosts: what the reference's
// samples load and what they write, so that real assets (an OBJ + MTL + PNG textures) drop in and outputs can be diffed against
// upstream runs.  The C++ twin of salviarenderer_b200/assets.py - same results, compared by tests/test_assets_cpp.py; the OBJ +
// MTL loader is pinned to the reference's own loader through the same fingerprints (tests/golden/obj_loader.json).
//
// * OBJ + MTL -> the reference's mesh layout (salvia/src/ext/resource/mesh/mesh_io_obj.cpp:41-45, 161-288, 389-452): one shared
//   vertex buffer of 48-byte vertices {pos.xyzw, uv.xyzw, normal.xyzw}, vertices de-duplicated by their (position, texcoord,
//   normal) index triple in first-use order, one u32 index list per material in usemtl order.
// * PNG -> rgba8 texels (salvia/src/ext/resource/texture/tex_io.cpp:28-75): rows bottom-up (FreeImage's scanline order, no
//   flip); images without an alpha channel get alpha 0 (`default_alpha`, freeimage_utilities.h:46-58).
// * surface -> PNG (tex_io.cpp:141-188): surface row 0 is the BOTTOM row of the file.
// * <name>_Profiling.json (salvia/src/utility/common/sample_app.cpp:448-567): every leaf a string; per counter {min, max,
//   total, avg} over the frames.
// Needs zlib (-lz) for the PNG functions; define SLV_ASSETS_NO_PNG to leave them out.
#pragma once

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include <tuple>
#include <vector>

#ifndef SLV_ASSETS_NO_PNG
#include <zlib.h>
#endif

namespace salvia_b200 {
namespace assets {

// obj_material (material.h) with the constructor's defaults (salvia/src/ext/resource/mesh/material.cpp:4-13)
struct obj_material {
  std::string name = "default";
  float ambient[4] = {0.2f, 0.2f, 0.2f, 1.0f};
  float diffuse[4] = {0.5f, 0.5f, 0.5f, 1.0f};
  float specular[4] = {0.7f, 0.7f, 0.7f, 1.0f};
  float alpha = 1.0f;
  int shininess = 2;
  bool is_specular = true;
  std::string tex_name, tex_path;
};

struct obj_mesh {
  std::vector<float> vertices;     // 12 floats per vertex: pos.xyzw, uv.xyzw, normal.xyzw (48-byte stride)
  std::vector<uint32_t> indices;   // 3 per triangle, into `vertices`
  std::vector<uint32_t> attrs;     // material index of each triangle
  std::vector<obj_material> materials;
  size_t vertex_count() const { return vertices.size() / 12; }
  // construct_meshes (mesh_io_obj.cpp:389-437): (material index, u32 index list) for every material that has triangles, in
  // material order; all groups share the one vertex buffer
  std::vector<std::pair<uint32_t, std::vector<uint32_t>>> material_groups() const {
    std::vector<std::pair<uint32_t, std::vector<uint32_t>>> out;
    for (uint32_t m = 0; m < materials.size(); ++m) {
      std::vector<uint32_t> idx;
      for (size_t t = 0; t < attrs.size(); ++t)
        if (attrs[t] == m) idx.insert(idx.end(), indices.begin() + 3 * t, indices.begin() + 3 * t + 3);
      if (!idx.empty()) out.emplace_back(m, std::move(idx));
    }
    return out;
  }
};

namespace detail {
inline std::vector<std::string> split(const std::string& line) {
  std::vector<std::string> out;
  std::istringstream ss(line);
  std::string t;
  while (ss >> t) out.push_back(t);
  return out;
}
// a float32 the way the Python twin reads it: parsed as double, rounded once to float; 0 when it is not a number
inline float to_float(const std::string& tok) {
  char* end = nullptr;
  const double v = std::strtod(tok.c_str(), &end);
  return (tok.empty() || *end) ? 0.0f : static_cast<float>(v);
}
inline std::string dirname(const std::string& p) {
  const size_t s = p.rfind('/');
  return s == std::string::npos ? std::string(".") : (s == 0 ? std::string("/") : p.substr(0, s));
}
}  // namespace detail

// load_material (mesh_io_obj.cpp:62-136): fills the materials `usemtl` created.  Upstream quirk, mirrored (pinned against the
// reference's own loader): a `newmtl` the OBJ never used does NOT deselect the current material, so its statements overwrite
// the previously selected one.
inline bool load_mtl(const std::string& path, std::vector<obj_material>& materials) {
  std::ifstream f(path);
  if (!f) return false;
  obj_material* cur = nullptr;
  const std::string base = detail::dirname(path);
  std::string line;
  while (std::getline(f, line)) {
    const auto tok = detail::split(line);
    if (tok.empty() || tok[0][0] == '#') continue;
    const std::string& cmd = tok[0];
    if (cmd == "newmtl") {
      const std::string name = tok.size() > 1 ? tok[1] : std::string();
      for (auto& m : materials) if (m.name == name) { cur = &m; break; }
      continue;
    }
    if (!cur) continue;
    if ((cmd == "Ka" || cmd == "Kd" || cmd == "Ks") && tok.size() >= 4) {
      float* dst = cmd == "Ka" ? cur->ambient : cmd == "Kd" ? cur->diffuse : cur->specular;
      dst[0] = detail::to_float(tok[1]); dst[1] = detail::to_float(tok[2]); dst[2] = detail::to_float(tok[3]); dst[3] = 0.0f;
    } else if ((cmd == "d" || cmd == "Tr") && tok.size() >= 2) cur->alpha = detail::to_float(tok[1]);
    else if (cmd == "Ns" && tok.size() >= 2) cur->shininess = static_cast<int>(detail::to_float(tok[1]));
    else if (cmd == "illum" && tok.size() >= 2) cur->is_specular = static_cast<int>(detail::to_float(tok[1])) == 2;
    else if (cmd == "map_Kd" && tok.size() >= 2) {
      cur->tex_name = tok[1];
      for (auto& c : cur->tex_name) if (c == '\\') c = '/';
      cur->tex_path = base + "/" + cur->tex_name;
    }
  }
  return true;
}

// load_obj_mesh_c (mesh_io_obj.cpp:140-269).  Triangulated faces only (the reference reads exactly three corners of every `f`
// line and ignores the rest); material 0 is the default material.  Upstream quirk, mirrored: the texcoord / normal indices of
// the de-duplication key are reset per FACE, not per corner, so a corner that omits them is keyed with the indices of the
// face's previous corner (its data are still zero).
inline bool load_obj(const std::string& path, bool flip_tex_v, obj_mesh& out) {
  std::ifstream f(path);
  if (!f) return false;
  out = obj_mesh();
  std::vector<float> positions, uvs, normals;  // 4 floats each
  out.materials.emplace_back();
  uint32_t subset = 0;
  std::string mtl_file, line;
  std::map<std::tuple<long, long, long>, uint32_t> seen;
  auto num = [](const std::vector<std::string>& tok, size_t k) { return k < tok.size() ? detail::to_float(tok[k]) : 0.0f; };
  auto push4 = [](std::vector<float>& v, float a, float b, float c, float d) { v.push_back(a); v.push_back(b); v.push_back(c); v.push_back(d); };
  auto append4 = [](std::vector<float>& dst, const std::vector<float>& src, long one_based, bool present) {
    for (int k = 0; k < 4; ++k) dst.push_back(present ? src.at(static_cast<size_t>((one_based - 1) * 4 + k)) : 0.0f);
  };
  while (std::getline(f, line)) {
    const auto tok = detail::split(line);
    if (tok.empty() || tok[0][0] == '#') continue;
    const std::string& cmd = tok[0];
    if (cmd == "v") push4(positions, num(tok, 1), num(tok, 2), num(tok, 3), 1.0f);
    else if (cmd == "vt") { const float v = num(tok, 2); push4(uvs, num(tok, 1), flip_tex_v ? 1.0f - v : v, 0.0f, 0.0f); }
    else if (cmd == "vn") push4(normals, num(tok, 1), num(tok, 2), num(tok, 3), 0.0f);
    else if (cmd == "f") {
      long ti = 0, ni = 0;  // declared per face, outside the corner loop, upstream: a corner inherits the face's previous indices
      for (size_t c = 1; c < tok.size() && c < 4; ++c) {
        std::vector<std::string> parts;
        std::string cur;
        for (char ch : tok[c]) { if (ch == '/') { parts.push_back(cur); cur.clear(); } else cur += ch; }
        parts.push_back(cur);
        const long pi = std::atol(parts[0].c_str());
        const bool has_t = parts.size() > 1 && !parts[1].empty(), has_n = parts.size() > 2 && !parts[2].empty();
        if (has_t) ti = std::atol(parts[1].c_str());
        if (has_n) ni = std::atol(parts[2].c_str());
        const auto key = std::make_tuple(pi, ti, ni);
        auto it = seen.find(key);
        uint32_t idx;
        if (it == seen.end()) {
          idx = static_cast<uint32_t>(out.vertex_count());
          seen[key] = idx;
          append4(out.vertices, positions, pi, true);
          append4(out.vertices, uvs, ti, has_t);
          append4(out.vertices, normals, ni, has_n);
        } else idx = it->second;
        out.indices.push_back(idx);
      }
      out.attrs.push_back(subset);
    } else if (cmd == "mtllib" && tok.size() > 1) mtl_file = tok[1];
    else if (cmd == "usemtl" && tok.size() > 1) {
      bool found = false;
      for (uint32_t i = 0; i < out.materials.size(); ++i) if (out.materials[i].name == tok[1]) { subset = i; found = true; break; }
      if (!found) {
        subset = static_cast<uint32_t>(out.materials.size());
        out.materials.emplace_back();
        out.materials.back().name = tok[1];
      }
    }
  }
  if (!mtl_file.empty()) load_mtl(detail::dirname(path) + "/" + mtl_file, out.materials);
  return true;
}

#ifndef SLV_ASSETS_NO_PNG
namespace detail {
inline uint32_t be32(const uint8_t* p) { return (uint32_t(p[0]) << 24) | (uint32_t(p[1]) << 16) | (uint32_t(p[2]) << 8) | p[3]; }
inline void put_be32(std::vector<uint8_t>& v, uint32_t x) { v.push_back(uint8_t(x >> 24)); v.push_back(uint8_t(x >> 16)); v.push_back(uint8_t(x >> 8)); v.push_back(uint8_t(x)); }
inline int paeth(int a, int b, int c) {
  const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
  return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}
}  // namespace detail

// load_texture(..., pixel_format_color_rgba8) (tex_io.cpp:28-95) for PNG files: `texels` = h rows of w rgba8 texels, row 0 = the
// file's BOTTOM row; no alpha channel in the file -> alpha 0.  Non-interlaced PNG, 8 bits per channel (grey, grey + alpha, RGB,
// RGBA, palette with optional tRNS) and 1 / 2 / 4-bit grey or palette images.
inline bool load_texture_rgba8(const std::string& path, uint32_t& w, uint32_t& h, std::vector<uint8_t>& texels, std::string* error = nullptr) {
  auto fail = [&](const char* msg) { if (error) *error = path + ": " + msg; return false; };
  std::ifstream f(path, std::ios::binary);
  if (!f) return fail("cannot open");
  const std::vector<uint8_t> file((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
  if (file.size() < 8 || std::memcmp(file.data(), sig, 8) != 0) return fail("not a PNG file");
  uint32_t depth = 0, ctype = 0, interlace = 0;
  std::vector<uint8_t> idat, palette, trns;
  bool have_header = false;
  for (size_t pos = 8; pos + 12 <= file.size();) {
    const uint32_t len = detail::be32(&file[pos]);
    const std::string type(reinterpret_cast<const char*>(&file[pos + 4]), 4);
    if (pos + 12 + len > file.size()) return fail("truncated chunk");
    const uint8_t* data = &file[pos + 8];
    if (type == "IHDR" && len >= 13) {
      w = detail::be32(data); h = detail::be32(data + 4); depth = data[8]; ctype = data[9]; interlace = data[12];
      have_header = true;
    } else if (type == "PLTE") palette.assign(data, data + len);
    else if (type == "tRNS") trns.assign(data, data + len);
    else if (type == "IDAT") idat.insert(idat.end(), data, data + len);
    else if (type == "IEND") break;
    pos += 12 + len;
  }
  if (!have_header || !w || !h) return fail("no IHDR");
  if (interlace) return fail("interlaced PNG files are not supported");
  const uint32_t channels = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
  const bool small_ok = (ctype == 0 || ctype == 3) && (depth == 1 || depth == 2 || depth == 4);
  if (!channels || (depth != 8 && !small_ok)) return fail("unsupported colour type / bit depth");
  const size_t stride = (size_t(w) * channels * depth + 7) / 8, bpp = (channels * depth + 7) / 8;
  std::vector<uint8_t> raw((stride + 1) * h);
  uLongf raw_len = static_cast<uLongf>(raw.size());
  if (uncompress(raw.data(), &raw_len, idat.data(), static_cast<uLong>(idat.size())) != Z_OK || raw_len != raw.size()) return fail("bad image data");
  std::vector<uint8_t> prev(stride, 0), cur(stride);
  texels.assign(size_t(w) * h * 4, 0);
  for (uint32_t y = 0; y < h; ++y) {
    const uint8_t filter = raw[(stride + 1) * y];
    const uint8_t* src = &raw[(stride + 1) * y + 1];
    for (size_t x = 0; x < stride; ++x) {
      const int a = x >= bpp ? cur[x - bpp] : 0, b = prev[x], c = x >= bpp ? prev[x - bpp] : 0;
      int v = src[x];
      switch (filter) {
        case 0: break;
        case 1: v += a; break;
        case 2: v += b; break;
        case 3: v += (a + b) / 2; break;
        case 4: v += detail::paeth(a, b, c); break;
        default: return fail("bad filter type");
      }
      cur[x] = static_cast<uint8_t>(v);
    }
    uint8_t* dst = &texels[size_t(h - 1 - y) * w * 4];  // bottom-up
    for (uint32_t x = 0; x < w; ++x) {
      uint8_t px[4] = {0, 0, 0, 0};
      uint32_t sample = 0;
      if (depth == 8) sample = cur[size_t(x) * channels];
      else { const uint32_t per = 8 / depth, shift = (per - 1 - x % per) * depth; sample = (cur[x / per] >> shift) & ((1u << depth) - 1); }
      switch (ctype) {
        case 0: { const uint8_t g = depth == 8 ? uint8_t(sample) : uint8_t(sample * 255 / ((1u << depth) - 1)); px[0] = px[1] = px[2] = g; break; }
        case 2: px[0] = cur[x * 3]; px[1] = cur[x * 3 + 1]; px[2] = cur[x * 3 + 2]; break;
        case 3:
          if (size_t(sample) * 3 + 2 < palette.size()) { px[0] = palette[sample * 3]; px[1] = palette[sample * 3 + 1]; px[2] = palette[sample * 3 + 2]; }
          if (!trns.empty()) px[3] = sample < trns.size() ? trns[sample] : 255;
          break;
        case 4: px[0] = px[1] = px[2] = cur[x * 2]; px[3] = cur[x * 2 + 1]; break;
        default: px[0] = cur[x * 4]; px[1] = cur[x * 4 + 1]; px[2] = cur[x * 4 + 2]; px[3] = cur[x * 4 + 3]; break;
      }
      std::memcpy(dst + size_t(x) * 4, px, 4);
    }
    prev.swap(cur);
  }
  return true;
}

// save_surface(..., pixel_format_color_bgra8) (tex_io.cpp:141-188): `texels` = h rows of w 4-byte texels in the surface's own
// channel order (bgra = true for a bgra8 surface), row 0 first in memory = bottom row of the PNG.  8-bit RGBA, filter 0.
inline bool save_surface_png(const std::string& path, const uint8_t* texels, uint32_t w, uint32_t h, bool bgra) {
  std::vector<uint8_t> raw;
  raw.reserve((size_t(w) * 4 + 1) * h);
  for (uint32_t y = 0; y < h; ++y) {
    const uint8_t* row = texels + size_t(h - 1 - y) * w * 4;
    raw.push_back(0);
    for (uint32_t x = 0; x < w; ++x) {
      const uint8_t* p = row + size_t(x) * 4;
      raw.push_back(bgra ? p[2] : p[0]); raw.push_back(p[1]); raw.push_back(bgra ? p[0] : p[2]); raw.push_back(p[3]);
    }
  }
  uLongf zlen = compressBound(static_cast<uLong>(raw.size()));
  std::vector<uint8_t> z(zlen);
  if (compress2(z.data(), &zlen, raw.data(), static_cast<uLong>(raw.size()), 6) != Z_OK) return false;
  z.resize(zlen);
  std::vector<uint8_t> out = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
  auto chunk = [&](const char* type, const std::vector<uint8_t>& data) {
    detail::put_be32(out, static_cast<uint32_t>(data.size()));
    const size_t start = out.size();
    out.insert(out.end(), type, type + 4);
    out.insert(out.end(), data.begin(), data.end());
    detail::put_be32(out, static_cast<uint32_t>(crc32(0L, &out[start], static_cast<uInt>(out.size() - start))));
  };
  std::vector<uint8_t> ihdr;
  detail::put_be32(ihdr, w);
  detail::put_be32(ihdr, h);
  const uint8_t tail[5] = {8, 6, 0, 0, 0};
  ihdr.insert(ihdr.end(), tail, tail + 5);
  chunk("IHDR", ihdr);
  chunk("IDAT", z);
  chunk("IEND", {});
  std::ofstream f(path, std::ios::binary);
  f.write(reinterpret_cast<const char*>(out.data()), static_cast<std::streamsize>(out.size()));
  return static_cast<bool>(f);
}
#endif  // SLV_ASSETS_NO_PNG

// sample_app::save_profiling_result (sample_app.cpp:468-567).  One map per frame: counter name -> value (missing = 0).
using profiling_frame = std::map<std::string, long long>;
inline const std::vector<std::string>& pipeline_stat_keys() {
  static const std::vector<std::string> k = {"cinvocations", "cprimitives", "ia_primitives", "ia_vertices", "vs_invocations", "ps_invocations"};
  return k;
}
inline const std::vector<std::string>& pipeline_prof_keys() {
  static const std::vector<std::string> k = {"gather_vtx", "vtx_proc", "clipping", "compact_clip", "vp_trans", "tri_dispatch", "ras"};
  return k;
}
// The JSON text (4-space indent, every leaf a string - as boost::property_tree::write_json emits them), identical to what
// assets.save_profiling_json of the Python twin writes.
inline std::string profiling_json(const std::string& compiler, const std::vector<profiling_frame>& frames) {
  auto esc = [](const std::string& s) {
    std::string o = "\"";
    for (unsigned char c : s) {
      if (c == '"') o += "\\\""; else if (c == '\\') o += "\\\\"; else if (c == '\n') o += "\\n"; else if (c == '\t') o += "\\t";
      else if (c < 0x20) { char b[8]; std::snprintf(b, sizeof(b), "\\u%04x", c); o += b; } else o += static_cast<char>(c);
    }
    return o + "\"";
  };
  auto reduce = [&](const std::string& key, const std::string& pad) {
    long long mn = 0, mx = 0, total = 0;
    for (size_t i = 0; i < frames.size(); ++i) {
      auto it = frames[i].find(key);
      const long long v = it == frames[i].end() ? 0 : it->second;
      if (i == 0 || v < mn) mn = v;
      if (i == 0 || v > mx) mx = v;
      total += v;
    }
    const long long avg = frames.empty() ? 0 : total / static_cast<long long>(frames.size());
    std::ostringstream o;
    o << "{\n" << pad << "    \"min\": \"" << mn << "\",\n" << pad << "    \"max\": \"" << mx << "\",\n" << pad << "    \"total\": \"" << total << "\",\n"
      << pad << "    \"avg\": \"" << avg << "\"\n" << pad << "}";
    return o.str();
  };
  auto group = [&](const std::vector<std::string>& keys) {
    std::string o = "{\n";
    for (size_t i = 0; i < keys.size(); ++i) o += "            " + esc(keys[i]) + ": " + reduce(keys[i], "            ") + (i + 1 < keys.size() ? ",\n" : "\n");
    return o + "        }";
  };
  std::ostringstream o;
  o << "{\n    \"compiler\": " << esc(compiler) << ",\n    \"frames\": \"" << frames.size() << "\",\n    \"async\": {\n"
    << "        \"pipeline_stat\": " << group(pipeline_stat_keys()) << ",\n"
    << "        \"internal_stat\": " << group({"backend_input_pixels"}) << ",\n"
    << "        \"pipeline_prof\": " << group(pipeline_prof_keys()) << "\n    }\n}";
  return o.str();
}
inline std::string save_profiling_json(const std::string& benchmark_name, const std::string& compiler, const std::vector<profiling_frame>& frames,
                                       const std::string& directory = ".") {
  const std::string path = directory + "/" + benchmark_name + "_Profiling.json";
  std::ofstream f(path);
  f << profiling_json(compiler, frames);
  return path;
}

}  // namespace assets
}  // namespace salvia_b200

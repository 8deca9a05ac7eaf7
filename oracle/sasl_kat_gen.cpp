// oracle/sasl_kat_gen.cpp — generates tests/golden/sasl_kat.json: known-answer tests for SASL intrinsics whose expected values
// are computed by the REFERENCE's own math library (eflib, compiled in place from /root/reference) with the inputs and the
// reference expressions of the reference's JIT test suite (sasl/test/jit_test/general.cpp:159-433, test case `intrinsics`:
// the `ref*` values each BOOST_CHECK_CLOSE compares against, tolerance RELATIVE_TORLERANCE_NORMAL = 1e-4 %).
// TEST INFRASTRUCTURE; built and run by `make -C oracle sasl-kat` (needs /root/reference); the JSON it prints is committed.
//
// Each entry: name, the SASL call (over globals a0, a1, ... declared with the listed types), the argument values, the expected
// components.  tests/test_sasl_kat.py evaluates the call through the front end on the CPU (host-compiled generated code) and
// tests/test_gpu_sasl_jit.py through the run-time compiled pixel shader on the GPU, and compares with `expected`.
#include <eflib/math/math.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

using namespace eflib;

// general.cpp:48-57
static vec4 lit_ref(float n_dot_l, float n_dot_h, float m) {
  return vec4(1.0f, std::max(n_dot_l, 0.0f), (n_dot_l < 0.0f) || (n_dot_h < 0.0f) ? 0.0f : (n_dot_h * m), 1.0f);
}
static vec3 faceforward_ref(vec3 n, vec3 i, vec3 ng) { return -n * eflib::sign(eflib::dot_prod3(i, ng)); }

struct Arg { std::string type; std::vector<float> v; };
// JSON has no NaN / infinity: those go out as strings ("nan", "inf", "-inf"); finite values with 9 significant digits (exact)
static std::string num(float v) {
  if (std::isnan(v)) return "\"nan\"";
  if (std::isinf(v)) return v > 0 ? "\"inf\"" : "\"-inf\"";
  char b[64];
  std::snprintf(b, sizeof(b), "%.9g", v);
  return b;
}
static bool first = true;
static void emit(const char* name, const char* call, std::vector<Arg> const& args, std::vector<float> const& expected) {
  std::printf("%s\n  {\"name\": \"%s\", \"call\": \"%s\", \"args\": [", first ? "" : ",", name, call);
  first = false;
  for (size_t i = 0; i < args.size(); ++i) {
    std::printf("%s{\"type\": \"%s\", \"values\": [", i ? ", " : "", args[i].type.c_str());
    for (size_t k = 0; k < args[i].v.size(); ++k) std::printf("%s%s", k ? ", " : "", num(args[i].v[k]).c_str());
    std::printf("]}");
  }
  std::printf("], \"expected\": [");
  for (size_t k = 0; k < expected.size(); ++k) std::printf("%s%s", k ? ", " : "", num(expected[k]).c_str());
  std::printf("]}");
}
static std::vector<float> f3(vec3 const& v) { return {v[0], v[1], v[2]}; }
static std::vector<float> f4(vec4 const& v) { return {v[0], v[1], v[2], v[3]}; }
static Arg a3(vec3 const& v) { return {"float3", f3(v)}; }
static Arg a1(float v) { return {"float", {v}}; }

int main() {
  std::printf("{\"generator\": \"oracle/sasl_kat_gen.cpp\", \"source\": \"eflib of the unmodified reference; inputs and reference expressions of sasl/test/jit_test/general.cpp:159-433\",\n \"relative_tolerance_percent\": 0.0001,\n \"cases\": [");
  {  // general.cpp:230-288
    vec3 lhs(4.0f, 9.3f, -5.9f), rhs(1.0f, -22.0f, 8.28f), ng(8.2f, 1.6f, 0.3f);
    emit("dot_f3", "dot(a0, a1)", {a3(lhs), a3(rhs)}, {dot_prod3(lhs.xyz(), rhs.xyz())});
    emit("normalize_lhs", "normalize(a0)", {a3(lhs)}, f3(normalize3(lhs)));
    emit("normalize_rhs", "normalize(a0)", {a3(rhs)}, f3(normalize3(rhs)));
    emit("reflect_lr", "reflect(a0, a1)", {a3(lhs), a3(rhs)}, f3(reflect3(lhs, rhs)));
    emit("reflect_rl", "reflect(a0, a1)", {a3(rhs), a3(lhs)}, f3(reflect3(rhs, lhs)));
    emit("refract_lr_022", "refract(a0, a1, a2)", {a3(lhs), a3(rhs), a1(0.22f)}, f3(refract3(lhs, rhs, 0.22f)));
    emit("refract_lr_071", "refract(a0, a1, a2)", {a3(lhs), a3(rhs), a1(0.71f)}, f3(refract3(lhs, rhs, 0.71f)));
    emit("refract_rl_022", "refract(a0, a1, a2)", {a3(rhs), a3(lhs), a1(0.22f)}, f3(refract3(rhs, lhs, 0.22f)));
    emit("refract_rl_071", "refract(a0, a1, a2)", {a3(rhs), a3(lhs), a1(0.71f)}, f3(refract3(rhs, lhs, 0.71f)));
    emit("faceforward_0", "faceforward(a0, a1, a2)", {a3(lhs), a3(rhs), a3(ng)}, f3(faceforward_ref(lhs, rhs, ng)));
    emit("faceforward_1", "faceforward(a0, a1, a2)", {a3(rhs), a3(ng), a3(lhs)}, f3(faceforward_ref(rhs, ng, lhs)));
    emit("faceforward_2", "faceforward(a0, a1, a2)", {a3(rhs), a3(lhs), a3(ng)}, f3(faceforward_ref(rhs, lhs, ng)));
    emit("lit_lhs", "lit(a0, a1, a2)", {a1(lhs[0]), a1(lhs[1]), a1(lhs[2])}, f4(lit_ref(lhs[0], lhs[1], lhs[2])));
    emit("lit_rhs", "lit(a0, a1, a2)", {a1(rhs[0]), a1(rhs[1]), a1(rhs[2])}, f4(lit_ref(rhs[0], rhs[1], rhs[2])));
    emit("lit_ng", "lit(a0, a1, a2)", {a1(ng[0]), a1(ng[1]), a1(ng[2])}, f4(lit_ref(ng[0], ng[1], ng[2])));
  }
  {  // general.cpp:290-313: mul(float4x4, float4) against eflib::transform(out, mat, vec)
    mat44 mat(mat44::identity());
    for (int i = 0; i < 16; ++i) ((float*)(&mat))[i] = static_cast<float>(i);
    mat44 tmp;
    mat_mul(mat, mat_rotX(tmp, 0.2f), mat);
    mat_mul(mat, mat_rotY(tmp, -0.3f), mat);
    mat_mul(mat, mat_translate(tmp, 1.7f, -0.9f, 1.1f), mat);
    mat_mul(mat, mat_scale(tmp, 0.5f, 1.2f, 2.0f), mat);
    vec4 rhs2(1.0f, 2.0f, 3.0f, 4.0f), refv;
    eflib::transform(refv, mat, rhs2);
    Arg m{"float4x4", {}};
    for (int i = 0; i < 16; ++i) m.v.push_back(((float*)(&mat))[i]);
    emit("mul_m44v4", "mul(a0, a1)", {m, {"float4", f4(rhs2)}}, f4(refv));
  }
  for (float p : {123.456f, -123.456f, 0.0f}) emit("abs_f", "abs(a0)", {a1(p)}, {fabsf(p)});                 // general.cpp:314-323
  for (float x : {-10.0f, -1.0f, 0.0f, 1.0f, 10.0f}) emit("exp_f", "exp(a0)", {a1(x)}, {expf(x)});          // general.cpp:336-352
  emit("sqrt_f", "sqrt(a0)", {a1(876.625f)}, {sqrtf(876.625f)});                                            // general.cpp:380-387
  emit("sqrt_f2", "sqrt(a0)", {{"float2", {1.7f, 986.27f}}}, {sqrtf(1.7f), sqrtf(986.27f)});
  {  // general.cpp:388-397
    vec3 a(199.7f, -872.5f, 8.63f), b(-98.7f, -37.29f, 77.3f);
    emit("cross_f3", "cross(a0, a1)", {a3(a), a3(b)}, f3(cross_prod3(a, b)));
  }
  {  // general.cpp:398-433
    vec3 v0(227.5f, -0.33f, -76.4f), v1(-113.8f, 17.22f, -9.44f);
    vec2 v2(87.9f, 54.2f);
    float f = 15.5;
    vec4 dst(1.0f, v1[1] * v0[0], v1[2], v0[1]);
    float ref0 = (v2 - v1.xy()).length() + (v2.xyxy() - dst).length();
    vec4 ref1 = vec4(fmodf(v0[0], v1[0]), fmodf(v0[1], v1[1]), fmodf(v0[2], v1[2]), fmodf(f, v1[1]));
    vec3 ref2 = v0 + (v1 - v0) * vec3(v2[0], v2[1], f);
    vec3 ref3 = ((eflib::PI_FLOAT / 180.0f) * v0.xy()).xxy() + (180.0f / eflib::PI_FLOAT) * v0;
    vec2 ref4(v2.length(), v0.xyzy().length());
    Arg av2{"float2", {v2[0], v2[1]}};
    emit("distance_dst", "distance(a0, a2.xy) + distance(dst(a2.xyzx, a1.yxzy), a0.xyxy)", {av2, a3(v0), a3(v1)}, {ref0});
    emit("fmod_f3f", "float4(fmod(a0, a2), fmod(a1, a2.y))", {a3(v0), a1(f), a3(v1)}, f4(ref1));
    emit("lerp_f3", "lerp(a0, a1, a2)", {a3(v0), a3(v1), a3(vec3(v2[0], v2[1], f))}, f3(ref2));
    emit("rad_deg", "radians(a0.xy).xxy + degrees(a0)", {a3(v0)}, f3(ref3));
    vec4 v0q = v0.xyzy();
    emit("length_f2_f4", "float2(length(a0), length(a1))", {av2, {"float4", f4(v0q)}}, {ref4[0], ref4[1]});
  }
  {  // general.cpp:447-566: the math intrinsics on a float3x4 (here: its three float4 rows), against the host functions the
     // reference binds its JIT-ed code to (libm's float functions; eflib's fast_* for floor / ceil / round / log / log2)
    const float arr0[3][4] = {{17.7f, 66.3f, 0.92f, -88.7f}, {8.6f, -0.22f, 17.1f, -64.4f}, {199.8f, 0.1f, -0.1f, 99.73f}};
    const float arr1[3][4] = {{9.62f, 10.33f, -18.2f, 99.7f}, {-0.3f, -76.9f, 93.3f, 0.22f}, {44.1f, 0.027f, 19.9f, -33.5f}};
    struct Fn { const char* name; const char* call; float (*f)(float); };
    const Fn fns[] = {
        {"exp", "exp(a0)", [](float x) { return (float)exp((double)x); }},
        {"exp2", "exp2(a0)", [](float x) { return (float)ldexp(1.0f, (int)x); }},
        {"sin", "sin(a0)", [](float x) { return sinf(x); }}, {"cos", "cos(a0)", [](float x) { return cosf(x); }},
        {"tan", "tan(a0)", [](float x) { return tanf(x); }}, {"sinh", "sinh(a0)", [](float x) { return sinhf(x); }},
        {"cosh", "cosh(a0)", [](float x) { return coshf(x); }}, {"tanh", "tanh(a0)", [](float x) { return tanhf(x); }},
        {"asin", "asin(a0)", [](float x) { return asinf(x); }}, {"acos", "acos(a0)", [](float x) { return acosf(x); }},
        {"atan", "atan(a0)", [](float x) { return atanf(x); }},
        {"ceil", "ceil(a0)", [](float x) { return fast_ceil(x); }}, {"floor", "floor(a0)", [](float x) { return fast_floor(x); }},
        {"round", "round(a0)", [](float x) { return fast_round(x); }}, {"trunc", "trunc(a0)", [](float x) { return (float)trunc(x); }},
        {"log", "log(a0)", [](float x) { return fast_log(x); }}, {"log2", "log2(a0)", [](float x) { return fast_log2(x); }},
        {"log10", "log10(a0)", [](float x) { return log10f(x); }},
        {"rsqrt", "rsqrt(a0)", [](float x) { return 1.0f / sqrtf(x); }}, {"rcp", "rcp(a0)", [](float x) { return 1.0f / x; }},
    };
    for (Fn const& fn : fns)
      for (int r = 0; r < 3; ++r) {
        std::vector<float> in(arr0[r], arr0[r] + 4), out;
        for (float x : in) out.push_back(fn.f(x));
        emit((std::string(fn.name) + "_row" + std::to_string(r)).c_str(), fn.call, {{"float4", in}}, out);
      }
    for (int r = 0; r < 3; ++r) {
      std::vector<float> a(arr0[r], arr0[r] + 4), b(arr1[r], arr1[r] + 4), p, l;
      for (int j = 0; j < 4; ++j) { p.push_back(powf(a[j], b[j])); l.push_back(ldexpf(a[j], (int)b[j])); }
      emit(("pow_row" + std::to_string(r)).c_str(), "pow(a0, a1)", {{"float4", a}, {"float4", b}}, p);
      emit(("ldexp_row" + std::to_string(r)).c_str(), "ldexp(a0, a1)", {{"float4", a}, {"float4", b}}, l);
    }
  }
  {  // general.cpp:1431-1494 (ps_branches): nested if / else if chains over a scalar and a float3; inputs and reference as there
     // (srand(0); rand() / 35.0f), the SASL function restated from the reference logic (kat_branches in tests/sasl_kat.py)
    srand(0);
    for (int i = 0; i < 16; ++i) {
      float in0 = (i * 0.34f) - 1.0f;
      float in1[3];
      for (int j = 0; j < 3; ++j) in1[j] = rand() / 35.0f;
      float r0 = 88.3f, r1 = 75.4f;
      if (in0 > 0.0f) r0 = in0;
      if (in0 > 1.0f) r1 = in1[0]; else r1 = in1[1];
      if (in0 > 2.0f) {
        r1 = in1[2];
        if (in0 > 3.0f) r1 = r1 + 1.0f;
        else if (in0 > 2.5f) r1 = r1 + 2.0f;
      }
      emit(("ps_branches_" + std::to_string(i)).c_str(), "kat_branches(a0, a1)", {a1(in0), {"float3", {in1[0], in1[1], in1[2]}}}, {r0, r1});
    }
  }
  std::printf("\n ],\n \"quad_cases\": [\n");
  // Known answers that need a whole 2x2 quad (PACKAGE_ELEMENT_COUNT = 4 pixels, two per line: pixel = row * 2 + col).
  auto row = [](std::vector<float> const& v) {
    std::string s = "[";
    for (size_t k = 0; k < v.size(); ++k) s += (k ? ", " : "") + num(v[k]);
    return s + "]";
  };
  auto quad = [&](const char* name, const char* what, std::vector<float> const (&in)[4], std::vector<float> const (&out)[4], bool last) {
    std::printf("  {\"name\": \"%s\", \"what\": \"%s\",\n   \"inputs\": [%s, %s, %s, %s],\n   \"expected\": [%s, %s, %s, %s]}%s\n", name, what, row(in[0]).c_str(),
                row(in[1]).c_str(), row(in[2]).c_str(), row(in[3]).c_str(), row(out[0]).c_str(), row(out[1]).c_str(), row(out[2]).c_str(), row(out[3]).c_str(),
                last ? "" : ",");
  };
  {  // general.cpp:1526-1602 (ddx_ddy): inputs srand(0), rand() / 67.0f in the test's order (v0, v1.xy, v2.xyz, v3.xyzw per pixel);
     // ddx = right - left of the pixel's line for both pixels of the line, ddy = lower - upper line for both pixels of the column
     // (get_ddx / get_ddy, general.cpp:1501-1524); out = (ddx(v0) + ddy(v0), ddx(v1).xy + ddy(v1).yx, ddx(v2).xyz + ddy(v2).yzx,
     // ddx(v3).xwzy + ddy(v3).yzxw)
    srand(0);
    std::vector<float> in[4], out[4];
    for (int i = 0; i < 4; ++i) for (int k = 0; k < 10; ++k) in[i].push_back(rand() / 67.0f);
    float ddx[4][10], ddy[4][10];
    for (int k = 0; k < 10; ++k) {
      ddx[0][k] = ddx[1][k] = in[1][k] - in[0][k];
      ddx[2][k] = ddx[3][k] = in[3][k] - in[2][k];
      ddy[0][k] = ddy[2][k] = in[2][k] - in[0][k];
      ddy[1][k] = ddy[3][k] = in[3][k] - in[1][k];
    }
    // component k of the flattened (v0, v1, v2, v3): which ddx / ddy component the reference expression adds
    const int sx[10] = {0, 1, 2, 3, 4, 5, 6, 9, 8, 7};   // .x | .xy | .xyz | .xwzy
    const int sy[10] = {0, 2, 1, 4, 5, 3, 7, 8, 6, 9};   // .x | .yx | .yzx | .yzxw
    for (int i = 0; i < 4; ++i) for (int k = 0; k < 10; ++k) out[i].push_back(ddx[i][sx[k]] + ddy[i][sy[k]]);
    quad("ddx_ddy", "general.cpp:1526-1602", in, out, false);
  }
  {  // general.cpp:1668-1716 (ps_for_loop): x doubled up to ten times, leaving the loop once it exceeds 5000
    srand(0);
    std::vector<float> in[4], out[4];
    for (int i = 0; i < 4; ++i) {
      float x = rand() / 1000.0f;
      in[i].push_back(x);
      for (int j = 0; j < 10; ++j) { x *= 2.0f; if (x > 5000.0f) break; }
      out[i].push_back(x);
    }
    quad("for_loop", "general.cpp:1668-1716", in, out, true);
  }
  std::printf(" ]}\n");
  return 0;
}

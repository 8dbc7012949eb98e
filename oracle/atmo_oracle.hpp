// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product.
//
// Scalar CPU restatement of the atmosphere hot path of Zylann/godot_atmosphere_shader @68766f34,
// written op-for-op after the GDShader sources (cited per function as file:line, paths relative to
// addons/zylann.atmosphere/shaders/). Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load this; the product (godot_atmosphere_shader_b200/)
// never does.
//
// PARITY PIN: the reference ships no tests, golden vectors or CPU implementation, and Godot (the only thing that
// runs GDShader) is absent here. The pin is the reference's OWN SOURCE TEXT: oracle/ref/build_ref.py compiles the
// shader files where they lie under /root/reference as C++ (a purely syntactic rewrite + a header of GLSL types and
// built-ins) into oracle/_ref/libatmo_ref.so, and tests/test_reference_pin.py requires this file (T=float) to equal
// it BIT FOR BIT — LUT, discard masks, fp32 RGBA — for all 7 shipped entry shaders, the BASELINE scale-ups, corner
// cameras and random scenes (and T=double to equal the same sources compiled with `float` = double,
// libatmo_ref64.so); a mutation test shows the comparison has teeth. What stays unpinned is what the
// reference leaves to the engine / GPU (the rounding of GLSL built-ins, texture filtering): defined below, used by
// both sides. Further pins: known-answer tests derived by hand from the shader source (tests/test_oracle_kat.py), the
// fp64 instantiation (same template, T=double), self-generated golden vectors (tests/golden/, generator committed).
//
// Everything is templated on the scalar T: T=float is THE oracle (fp32, the shader's precision;
// build with -ffp-contract=off so no FMA contraction happens), T=double bounds its rounding error.
//
// Engine-side behaviour the reference does not pin is DEFINED here:
//   * texture filtering = exact lerp a+(b-a)*t at LOD 0, texel centres at (i+0.5)/N
//   * LUT: clamp-to-edge (`repeat_disable`, include/atmosphere_funcs_v2.gdshaderinc:7)
//   * 3D shape: repeat on all axes (`repeat_enable`, include/cloud_funcs.gdshaderinc:10)
//   * cubemap: standard GL/Vulkan face tables (consistent with noise_cubemap.gd:110-128), SEAMLESS
//     bilinear via a 1-texel apron copied from the adjacent faces (Vulkan seamless-cube rule; corner
//     apron = rounded mean of the three texels meeting there); no mip selection
//   * pow(x,16) with x<=0 returns 0 (GLSL leaves x<0 undefined; NVIDIA yields NaN -> max(NaN,0)=0)
//   * normalize(v)=v/sqrt(dot(v,v)); mix(a,b,t)=a*(1-t)+b*t; smoothstep per the GLSL spec
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../include/b200atmo.h"

namespace oracle {

// ------------------------------------------------------------------------------------------------
// GLSL value types and built-ins
// ------------------------------------------------------------------------------------------------
template <class T> struct vec2 { T x, y; };
template <class T> struct vec3 { T x, y, z; };
template <class T> struct vec4 { T x, y, z, w; };
// column-major like GLSL: c[col][row]
template <class T> struct mat4 { T c[4][4]; };
template <class T> struct mat2 { T c[2][2]; };

template <class T> inline vec3<T> operator+(vec3<T> a, vec3<T> b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
template <class T> inline vec3<T> operator-(vec3<T> a, vec3<T> b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
template <class T> inline vec3<T> operator-(vec3<T> a) { return {-a.x, -a.y, -a.z}; }
template <class T> inline vec3<T> operator*(vec3<T> a, vec3<T> b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
template <class T> inline vec3<T> operator*(vec3<T> a, T s) { return {a.x * s, a.y * s, a.z * s}; }
template <class T> inline vec3<T> operator*(T s, vec3<T> a) { return {s * a.x, s * a.y, s * a.z}; }
template <class T> inline vec3<T> operator/(vec3<T> a, T s) { return {a.x / s, a.y / s, a.z / s}; }

template <class T> inline T dot(vec3<T> a, vec3<T> b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <class T> inline T length(vec3<T> a) { return std::sqrt(dot(a, a)); }
template <class T> inline T distance(vec3<T> a, vec3<T> b) { return length(a - b); }
template <class T> inline vec3<T> normalize(vec3<T> a) { return a / length(a); }
template <class T> inline T clamp(T x, T lo, T hi) { return std::min(std::max(x, lo), hi); }
template <class T> inline T mix(T a, T b, T t) { return a * (T(1) - t) + b * t; }
template <class T> inline vec3<T> mix(vec3<T> a, vec3<T> b, T t) { return {mix(a.x, b.x, t), mix(a.y, b.y, t), mix(a.z, b.z, t)}; }
template <class T> inline T smoothstep(T e0, T e1, T x) {
    T t = clamp((x - e0) / (e1 - e0), T(0), T(1));
    return t * t * (T(3) - T(2) * t);
}
template <class T> inline vec4<T> mul(const mat4<T>& m, vec4<T> v) {
    vec4<T> r;
    r.x = m.c[0][0] * v.x + m.c[1][0] * v.y + m.c[2][0] * v.z + m.c[3][0] * v.w;
    r.y = m.c[0][1] * v.x + m.c[1][1] * v.y + m.c[2][1] * v.z + m.c[3][1] * v.w;
    r.z = m.c[0][2] * v.x + m.c[1][2] * v.y + m.c[2][2] * v.z + m.c[3][2] * v.w;
    r.w = m.c[0][3] * v.x + m.c[1][3] * v.y + m.c[2][3] * v.z + m.c[3][3] * v.w;
    return r;
}
template <class T> inline mat4<T> mul(const mat4<T>& a, const mat4<T>& b) {
    mat4<T> r;
    for (int col = 0; col < 4; ++col)
        for (int row = 0; row < 4; ++row)
            r.c[col][row] = a.c[0][row] * b.c[col][0] + a.c[1][row] * b.c[col][1] + a.c[2][row] * b.c[col][2] +
                            a.c[3][row] * b.c[col][3];
    return r;
}
template <class T> inline mat4<T> load_mat4(const float* m) {
    mat4<T> r;
    for (int col = 0; col < 4; ++col)
        for (int row = 0; row < 4; ++row) r.c[col][row] = T(m[col * 4 + row]);
    return r;
}
template <class T> inline vec3<T> load_vec3(const float* v) { return {T(v[0]), T(v[1]), T(v[2])}; }

// include/util.gdshaderinc:5-17 — unused by the shipped shaders; used here to restate the MODE_FAR proxy-cube
// coverage of the rasteriser (planet_atmosphere.gd:261-282,302-321). max/min with IEEE fmax/fmin semantics (a NaN
// operand is ignored — what GPU min/max instructions do and what this routine relies on for axis-parallel rays).
template <class T> inline T sel_max(T a, T b) { return std::fmax(a, b); }
template <class T> inline T sel_min(T a, T b) { return std::fmin(a, b); }
template <class T> inline vec2<T> ray_box_intersection(vec3<T> ro, vec3<T> rd, vec3<T> boxSize) {
    vec3<T> m = {T(1) / rd.x, T(1) / rd.y, T(1) / rd.z};
    vec3<T> n = m * ro;
    vec3<T> k = vec3<T>{std::fabs(m.x), std::fabs(m.y), std::fabs(m.z)} * boxSize;
    vec3<T> t1 = -n - k;
    vec3<T> t2 = -n + k;
    T tN = sel_max(sel_max(t1.x, t1.y), t1.z);
    T tF = sel_min(sel_min(t2.x, t2.y), t2.z);
    if (tN > tF || tF < T(0)) return {T(-1), T(-1)};
    return {tN, tF};
}

// include/util.gdshaderinc:49-59
template <class T> inline T pow4(T x) { return x * x * x * x; }
template <class T> inline T pow3(T x) { return x * x * x; }
template <class T> inline T pow2(T x) { return x * x; }

// include/util.gdshaderinc:20-40 — x = first hit, y = second hit; equal (1e6,1e6) if not hit.
template <class T> inline vec2<T> ray_sphere(vec3<T> center, T radius, vec3<T> ray_origin, vec3<T> ray_dir) {
    vec3<T> oc = ray_origin - center;
    T b = dot(oc, ray_dir);
    vec3<T> qc = oc - b * ray_dir;
    T h = radius * radius - dot(qc, qc);
    if (h < T(0)) return {T(1000000.0), T(1000000.0)};
    h = std::sqrt(h);
    return {-b - h, -b + h};
}

// include/util.gdshaderinc:61-69
template <class T> inline vec4<T> blend_colors(vec4<T> self, vec4<T> over) {
    T sa = T(1) - over.w;
    T a = self.w * sa + over.w;
    if (a == T(0)) return {T(0), T(0), T(0), T(0)};
    return {(self.x * self.w * sa + over.x * over.w) / a, (self.y * self.w * sa + over.y * over.w) / a,
            (self.z * self.w * sa + over.z * over.w) / a, a};
}

// ------------------------------------------------------------------------------------------------
// Uniform block converted once per call from the C-ABI POD
// ------------------------------------------------------------------------------------------------
template <class T> struct Uniforms {
    T u_planet_radius, u_atmosphere_height, u_density;
    T u_scattering_strength;
    vec3<T> u_scattering_wavelengths, u_atmosphere_modulate, u_atmosphere_ambient_color;
    T u_sphere_depth_factor;
    T u_cloud_density_scale, u_cloud_bottom, u_cloud_top, u_cloud_blend, u_cloud_shape_invert, u_cloud_coverage_bias,
        u_cloud_shape_factor, u_cloud_shape_scale;
    mat2<T> u_cloud_coverage_rotation;
    mat4<T> u_world_to_model_matrix;
    vec4<T> u_day_color0, u_day_color1, u_night_color0, u_night_color1;
    T u_day_night_transition_scale;
    // textures (borrowed pointers)
    const T* lut = nullptr;              // 256 x 256, row = height ratio, column = 0.5+0.5*cos(theta)
    const uint8_t* shape = nullptr;      // nx*ny*nz, x fastest
    int shape_nx = 0, shape_ny = 0, shape_nz = 0;
    const uint8_t* cube_padded = nullptr;  // 6 x (res+2)^2, see cube_build_padded()
    int cube_res = 0;
};

template <class T> inline Uniforms<T> make_uniforms(const B200AtmoParams& p) {
    Uniforms<T> u;
    u.u_planet_radius = T(p.planet_radius);
    u.u_atmosphere_height = T(p.atmosphere_height);
    u.u_density = T(p.density);
    u.u_scattering_strength = T(p.scattering_strength);
    u.u_scattering_wavelengths = load_vec3<T>(p.scattering_wavelengths);
    u.u_atmosphere_modulate = load_vec3<T>(p.atmosphere_modulate);
    u.u_atmosphere_ambient_color = load_vec3<T>(p.atmosphere_ambient_color);
    u.u_sphere_depth_factor = T(p.sphere_depth_factor);
    u.u_cloud_density_scale = T(p.cloud_density_scale);
    u.u_cloud_bottom = T(p.cloud_bottom);
    u.u_cloud_top = T(p.cloud_top);
    u.u_cloud_blend = T(p.cloud_blend);
    u.u_cloud_shape_invert = T(p.cloud_shape_invert);
    u.u_cloud_coverage_bias = T(p.cloud_coverage_bias);
    u.u_cloud_shape_factor = T(p.cloud_shape_factor);
    u.u_cloud_shape_scale = T(p.cloud_shape_scale);
    u.u_cloud_coverage_rotation.c[0][0] = T(p.cloud_coverage_rotation[0]);
    u.u_cloud_coverage_rotation.c[0][1] = T(p.cloud_coverage_rotation[1]);
    u.u_cloud_coverage_rotation.c[1][0] = T(p.cloud_coverage_rotation[2]);
    u.u_cloud_coverage_rotation.c[1][1] = T(p.cloud_coverage_rotation[3]);
    u.u_world_to_model_matrix = load_mat4<T>(p.world_to_model);
    u.u_day_color0 = {T(p.day_color0[0]), T(p.day_color0[1]), T(p.day_color0[2]), T(p.day_color0[3])};
    u.u_day_color1 = {T(p.day_color1[0]), T(p.day_color1[1]), T(p.day_color1[2]), T(p.day_color1[3])};
    u.u_night_color0 = {T(p.night_color0[0]), T(p.night_color0[1]), T(p.night_color0[2]), T(p.night_color0[3])};
    u.u_night_color1 = {T(p.night_color1[0]), T(p.night_color1[1]), T(p.night_color1[2]), T(p.night_color1[3])};
    u.u_day_night_transition_scale = T(p.day_night_transition_scale);
    return u;
}

// ------------------------------------------------------------------------------------------------
// Texture fetches (oracle-defined conventions, see header comment)
// ------------------------------------------------------------------------------------------------
template <class T> inline T tex_lerp(T a, T b, T t) { return a + (b - a) * t; }

// texture(u_optical_depth_texture, vec2(u, v)).r  — include/atmosphere_funcs_v2.gdshaderinc:28
template <class T> inline T sample_lut(const T* lut, T u, T v) {
    const int N = B200ATMO_LUT_SIZE;
    T x = u * T(N) - T(0.5);
    T y = v * T(N) - T(0.5);
    if (!(x == x) || !(y == y)) return x + y;  // NaN in -> NaN out
    T x0f = std::floor(x), y0f = std::floor(y);
    T fx = x - x0f, fy = y - y0f;
    int x0 = int(x0f), y0 = int(y0f);
    int x1 = x0 + 1, y1 = y0 + 1;
    x0 = std::min(std::max(x0, 0), N - 1);
    x1 = std::min(std::max(x1, 0), N - 1);
    y0 = std::min(std::max(y0, 0), N - 1);
    y1 = std::min(std::max(y1, 0), N - 1);
    T t00 = T(lut[y0 * N + x0]), t10 = T(lut[y0 * N + x1]);
    T t01 = T(lut[y1 * N + x0]), t11 = T(lut[y1 * N + x1]);
    return tex_lerp(tex_lerp(t00, t10, fx), tex_lerp(t01, t11, fx), fy);
}

inline int wrap_index(int i, int n) {
    int m = i % n;
    return m < 0 ? m + n : m;
}

// texture(u_cloud_shape_texture, p).r — include/cloud_funcs.gdshaderinc:49 (repeat, trilinear, LOD 0)
template <class T> inline T sample_shape3d(const Uniforms<T>& u, vec3<T> p) {
    const int nx = u.shape_nx, ny = u.shape_ny, nz = u.shape_nz;
    if (u.shape == nullptr) return T(1);  // unset sampler = white
    T cx = p.x - std::floor(p.x), cy = p.y - std::floor(p.y), cz = p.z - std::floor(p.z);
    T xf = cx * T(nx) - T(0.5), yf = cy * T(ny) - T(0.5), zf = cz * T(nz) - T(0.5);
    if (!(xf == xf) || !(yf == yf) || !(zf == zf)) return xf + yf + zf;
    T x0f = std::floor(xf), y0f = std::floor(yf), z0f = std::floor(zf);
    T fx = xf - x0f, fy = yf - y0f, fz = zf - z0f;
    int x0 = wrap_index(int(x0f), nx), x1 = wrap_index(int(x0f) + 1, nx);
    int y0 = wrap_index(int(y0f), ny), y1 = wrap_index(int(y0f) + 1, ny);
    int z0 = wrap_index(int(z0f), nz), z1 = wrap_index(int(z0f) + 1, nz);
    auto at = [&](int x, int y, int z) { return T(u.shape[(size_t(z) * ny + y) * nx + x]) / T(255); };
    T c00 = tex_lerp(at(x0, y0, z0), at(x1, y0, z0), fx);
    T c10 = tex_lerp(at(x0, y1, z0), at(x1, y1, z0), fx);
    T c01 = tex_lerp(at(x0, y0, z1), at(x1, y0, z1), fx);
    T c11 = tex_lerp(at(x0, y1, z1), at(x1, y1, z1), fx);
    return tex_lerp(tex_lerp(c00, c10, fy), tex_lerp(c01, c11, fy), fz);
}

// Cube face bases: dir = n + sc*S + tc*Tv for face f (+X,-X,+Y,-Y,+Z,-Z), the inverse of the GL/Vulkan
// (sc,tc) tables; consistent with the generator's texel->direction mapping (noise_cubemap.gd:110-128).
static const int kFaceN[6][3] = {{1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}};
static const int kFaceS[6][3] = {{0, 0, -1}, {0, 0, 1}, {1, 0, 0}, {1, 0, 0}, {1, 0, 0}, {-1, 0, 0}};
static const int kFaceT[6][3] = {{0, -1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}, {0, -1, 0}, {0, -1, 0}};

// Integer face select of a point on the cube surface given in "doubled texel units" (face plane at
// +-res, texel centres at odd offsets 2i+1-res). Returns face and texel (i,j).
inline void cube_int_lookup(const int P[3], int res, int* face, int* i, int* j) {
    int f;
    if (std::abs(P[0]) == res) f = P[0] > 0 ? 0 : 1;
    else if (std::abs(P[1]) == res) f = P[1] > 0 ? 2 : 3;
    else f = P[2] > 0 ? 4 : 5;
    int sc = kFaceS[f][0] * P[0] + kFaceS[f][1] * P[1] + kFaceS[f][2] * P[2];
    int tc = kFaceT[f][0] * P[0] + kFaceT[f][1] * P[1] + kFaceT[f][2] * P[2];
    *face = f;
    *i = (sc + res - 1) / 2;
    *j = (tc + res - 1) / 2;
}

// Builds the seamless layout: 6 faces of (res+2)^2 bytes; padded[f][j+1][i+1] = face f texel (i,j);
// edge apron = the texel adjacent across the cube edge; corner apron = round((a+b+c)/3) of the two
// neighbouring apron texels and the face's own corner texel.
inline void cube_build_padded(const uint8_t* faces, int res, uint8_t* out) {
    const int pr = res + 2;
    for (int f = 0; f < 6; ++f) {
        uint8_t* o = out + size_t(f) * pr * pr;
        for (int j = -1; j <= res; ++j) {
            for (int i = -1; i <= res; ++i) {
                bool io = (i < 0 || i >= res), jo = (j < 0 || j >= res);
                if (io && jo) continue;  // corners later
                uint8_t v;
                if (!io && !jo) {
                    v = faces[(size_t(f) * res + j) * res + i];
                } else {
                    int sc = 2 * i + 1 - res, tc = 2 * j + 1 - res;  // one of them is +-(res+1)
                    int nrm = res, s = sc, t = tc;
                    if (io) { nrm = res - 1; s = sc > 0 ? res : -res; }
                    else    { nrm = res - 1; t = tc > 0 ? res : -res; }
                    int P[3];
                    for (int k = 0; k < 3; ++k) P[k] = kFaceN[f][k] * nrm + kFaceS[f][k] * s + kFaceT[f][k] * t;
                    int g, gi, gj;
                    cube_int_lookup(P, res, &g, &gi, &gj);
                    v = faces[(size_t(g) * res + gj) * res + gi];
                }
                o[(j + 1) * pr + (i + 1)] = v;
            }
        }
        const int cs[2] = {0, pr - 1};
        for (int cj = 0; cj < 2; ++cj)
            for (int ci = 0; ci < 2; ++ci) {
                int x = cs[ci], y = cs[cj];
                int xi = x == 0 ? 1 : pr - 2, yi = y == 0 ? 1 : pr - 2;
                int a = o[y * pr + xi], b = o[yi * pr + x], c = o[yi * pr + xi];
                o[y * pr + x] = uint8_t((2 * (a + b + c) + 3) / 6);
            }
    }
}

// texture(u_cloud_coverage_cubemap, dir).r — include/cloud_funcs.gdshaderinc:45
template <class T> inline T sample_cube(const Uniforms<T>& u, vec3<T> d) {
    if (u.cube_padded == nullptr) return T(1);  // unset sampler = white: "cover the whole atmosphere uniformly" (README.md:46)
    const int res = u.cube_res, pr = res + 2;
    T ax = std::fabs(d.x), ay = std::fabs(d.y), az = std::fabs(d.z);
    int f;
    T sc, tc, ma;
    if (ax >= ay && ax >= az) { f = d.x >= T(0) ? 0 : 1; ma = ax; sc = d.x >= T(0) ? -d.z : d.z; tc = -d.y; }
    else if (ay >= az)        { f = d.y >= T(0) ? 2 : 3; ma = ay; sc = d.x; tc = d.y >= T(0) ? d.z : -d.z; }
    else                      { f = d.z >= T(0) ? 4 : 5; ma = az; sc = d.z >= T(0) ? d.x : -d.x; tc = -d.y; }
    T s = T(0.5) * (sc / ma + T(1)), t = T(0.5) * (tc / ma + T(1));
    T xf = s * T(res) - T(0.5), yf = t * T(res) - T(0.5);
    if (!(xf == xf) || !(yf == yf)) return xf + yf;
    T x0f = std::floor(xf), y0f = std::floor(yf);
    T fx = xf - x0f, fy = yf - y0f;
    int x0 = std::min(std::max(int(x0f), -1), res - 1) + 1;  // padded index of the left/top texel
    int y0 = std::min(std::max(int(y0f), -1), res - 1) + 1;
    const uint8_t* o = u.cube_padded + size_t(f) * pr * pr;
    auto at = [&](int x, int y) { return T(o[y * pr + x]) / T(255); };
    return tex_lerp(tex_lerp(at(x0, y0), at(x0 + 1, y0), fx), tex_lerp(at(x0, y0 + 1), at(x0 + 1, y0 + 1), fx), fy);
}

// ------------------------------------------------------------------------------------------------
// include/atmosphere_common.gdshaderinc:12-24
// ------------------------------------------------------------------------------------------------
template <class T> inline T get_atmosphere_density(const Uniforms<T>& u, T height) {
    T sd = height - u.u_planet_radius;
    T h = clamp(sd / u.u_atmosphere_height, T(0), T(1));
    T y = T(1) - h;
    T density = y * y * y * u.u_density;
    return density;
}

// ------------------------------------------------------------------------------------------------
// optical_depth.gdshader:17-31 (get_optical_depth) and :45-69 (fragment) — one LUT texel
// ------------------------------------------------------------------------------------------------
template <class T> inline T get_optical_depth(const Uniforms<T>& u, vec2<T> ray_origin, vec2<T> ray_dir, T ray_len) {
    const int steps = 64;
    T step_len = ray_len / T(steps);
    T optical_depth = T(0);
    for (int i = 0; i < steps; ++i) {
        vec2<T> pos = {ray_origin.x + ray_dir.x * step_len * T(i), ray_origin.y + ray_dir.y * step_len * T(i)};
        T d = std::sqrt(pos.x * pos.x + pos.y * pos.y);
        T density = get_atmosphere_density(u, d);
        optical_depth += density * step_len * u.u_density;
    }
    return optical_depth;
}

template <class T> inline T bake_texel(const Uniforms<T>& u, int i, int j) {
    // canvas UV of the texel centre; the SubViewport is 256x256 (optical_depth_baker.gd:24)
    vec2<T> uv = {(T(i) + T(0.5)) / T(B200ATMO_LUT_SIZE), (T(j) + T(0.5)) / T(B200ATMO_LUT_SIZE)};
    vec2<T> ray_dir;
    ray_dir.y = T(2) * uv.x - T(1);
    ray_dir.x = std::sqrt(T(1) - ray_dir.y * ray_dir.y);
    T height_ratio = uv.y;
    vec2<T> pos = {T(0), u.u_planet_radius + u.u_atmosphere_height * height_ratio};
    vec2<T> rs = ray_sphere<T>({T(0), T(0), T(0)}, u.u_planet_radius + u.u_atmosphere_height, {pos.x, pos.y, T(0)},
                               {ray_dir.x, ray_dir.y, T(0)});
    T distance_through_atmosphere = rs.y - std::max(rs.x, T(0));
    return get_optical_depth(u, pos, ray_dir, distance_through_atmosphere);
    // encode_float_to_viewport (:33-43) + Image.create_from_data(FORMAT_RF) (optical_depth_baker.gd:75-77)
    // is a lossless fp32 round trip through RGBA8; see encode/decode helpers below.
}

// optical_depth.gdshader:33-43 — float bits -> 4 bytes (the /255 and the viewport's *255 cancel exactly)
inline void encode_float_to_viewport(float h, uint8_t out[4]) {
    uint32_t u;
    std::memcpy(&u, &h, 4);
    out[0] = uint8_t((u >> 0u) & 255u);
    out[1] = uint8_t((u >> 8u) & 255u);
    out[2] = uint8_t((u >> 16u) & 255u);
    out[3] = uint8_t((u >> 24u) & 255u);
}
// optical_depth_baker.gd:75-77 — RGBA8 bytes reinterpreted as FORMAT_RF (little endian)
inline float decode_viewport_bytes(const uint8_t in[4]) {
    uint32_t u = uint32_t(in[0]) | (uint32_t(in[1]) << 8) | (uint32_t(in[2]) << 16) | (uint32_t(in[3]) << 24);
    float f;
    std::memcpy(&f, &u, 4);
    return f;
}

// ------------------------------------------------------------------------------------------------
// include/atmosphere_funcs_v2.gdshaderinc
// ------------------------------------------------------------------------------------------------
// :14-29
template <class T> inline T get_baked_optical_depth(const Uniforms<T>& u, vec3<T> pos, vec3<T> dir, vec3<T> planet_center) {
    T height = distance(pos, planet_center) - u.u_planet_radius;
    T height_ratio = clamp(height / u.u_atmosphere_height, T(0), T(1));
    vec3<T> up = normalize(pos - planet_center);
    T uvx = T(0.5) + T(0.5) * dot(up, dir);
    return sample_lut<T>(u.lut, uvx, height_ratio);
}

// :32-101 — returns RGBA
template <class T>
inline vec4<T> compute_atmosphere_v2(const Uniforms<T>& u, int steps, vec3<T> ray_origin, vec3<T> ray_dir,
                                     vec3<T> planet_center, T t_begin, T t_end, T /*linear_depth*/, vec3<T> sun_dir,
                                     T jitter) {
    vec3<T> scattering_coefficients =
        vec3<T>{pow4(T(400) / u.u_scattering_wavelengths.x), pow4(T(400) / u.u_scattering_wavelengths.y),
                pow4(T(400) / u.u_scattering_wavelengths.z)} *
        u.u_scattering_strength;

    T step_len = (t_end - t_begin) / T(steps);
    vec3<T> total_light = {T(0), T(0), T(0)};
    T view_ray_optical_depth = T(0);
    T alpha = T(0);
    vec3<T> pos0 = ray_origin + ray_dir * t_begin;
    vec3<T> pos = pos0;

    for (int i = 0; i < steps; ++i) {
        T sun_ray_optical_depth = get_baked_optical_depth(u, pos, sun_dir, planet_center);

        T height = distance(pos, planet_center);
        T local_density = get_atmosphere_density(u, height) * u.u_density;
        view_ray_optical_depth += local_density * step_len;

        T od = -(sun_ray_optical_depth + view_ray_optical_depth);
        vec3<T> transmittance = {std::exp(od * scattering_coefficients.x), std::exp(od * scattering_coefficients.y),
                                 std::exp(od * scattering_coefficients.z)};

        total_light = total_light + local_density * step_len * transmittance * scattering_coefficients;

        T vtransmittance = std::exp(-local_density * step_len);
        alpha += (T(1) - vtransmittance) * (T(1) - alpha);

        pos = pos + ray_dir * step_len;
    }

    total_light = total_light + u.u_atmosphere_ambient_color;
    total_light = {clamp(total_light.x, T(0), T(1)), clamp(total_light.y, T(0), T(1)), clamp(total_light.z, T(0), T(1))};

    alpha = clamp(alpha + jitter * T(0.02), T(0), T(0.99));

    total_light = total_light * u.u_atmosphere_modulate;

    return {total_light.x, total_light.y, total_light.z, alpha};
}

// ------------------------------------------------------------------------------------------------
// include/atmosphere_funcs_v1.gdshaderinc ("lite")
// ------------------------------------------------------------------------------------------------
// :15-46
template <class T>
inline T get_atmo_factor(const Uniforms<T>& u, int steps, vec3<T> ray_origin, vec3<T> ray_dir, vec3<T> planet_center,
                         T t_begin, T t_end, vec3<T> sun_dir, T* light_factor) {
    T inv_steps = T(1) / T(steps);
    T step_len = (t_end - t_begin) * inv_steps;
    vec3<T> stepv = step_len * ray_dir;
    vec3<T> pos = ray_origin + ray_dir * t_begin;

    T factor = T(1);
    T light_sum = T(0);

    for (int i = 0; i < steps; ++i) {
        T d = distance(pos, planet_center);
        vec3<T> up = (pos - planet_center) / d;
        T density = get_atmosphere_density(u, d);
        T light = clamp(T(1.2) * dot(sun_dir, up) + T(0.5), T(0), T(1));
        light = light * light;
        light_sum += light * inv_steps;
        factor *= (T(1) - density * step_len);
        pos = pos + stepv;
    }

    *light_factor = light_sum;
    return T(1) - factor;
}

// :49-63
template <class T>
inline vec4<T> compute_atmosphere_v1(const Uniforms<T>& u, int steps, vec3<T> ray_origin, vec3<T> ray_dir,
                                     vec3<T> planet_center, T t_begin, T t_end, vec3<T> sun_dir) {
    T light_factor;
    T atmo_factor = get_atmo_factor(u, steps, ray_origin, ray_dir, planet_center, t_begin, t_end, sun_dir, &light_factor);

    auto rgb = [](vec4<T> c) { return vec3<T>{c.x, c.y, c.z}; };
    vec3<T> night_col = mix(rgb(u.u_night_color0), rgb(u.u_night_color1), atmo_factor);
    vec3<T> day_col = mix(rgb(u.u_day_color0), rgb(u.u_day_color1), atmo_factor);

    T day_factor = clamp(light_factor * u.u_day_night_transition_scale, T(0), T(1));
    vec3<T> col = mix(night_col, day_col, day_factor);

    return {col.x, col.y, col.z, clamp(atmo_factor, T(0), T(1))};
}

// ------------------------------------------------------------------------------------------------
// include/cloud_funcs.gdshaderinc
// ------------------------------------------------------------------------------------------------
template <class T> struct CloudSettings {  // :18-23
    T bottom_height, top_height, density_scale, ground_height;
};

// :25-29
template <class T> inline T height_curve(T x) { return T(1) - pow2(T(2) * x - T(1)); }

// :31-68 (CLOUDS_ALWAYS_LOW_QUALITY is hard-defined at planet_atmosphere_main.gdshaderinc:49 => low = true,
// detail = 0.5, `time` has no effect)
template <class T> inline T get_density_full(const Uniforms<T>& u, vec3<T> pos_world, const CloudSettings<T>& settings) {
    T height = length(pos_world) - settings.bottom_height;
    T height_ratio = height / (settings.top_height - settings.bottom_height);

    T hc = std::max(height_curve(height_ratio), T(0));
    T density = T(1);

    // coverage_pos_2d = u_cloud_coverage_rotation * pos_world.xz
    vec2<T> cp2 = {u.u_cloud_coverage_rotation.c[0][0] * pos_world.x + u.u_cloud_coverage_rotation.c[1][0] * pos_world.z,
                   u.u_cloud_coverage_rotation.c[0][1] * pos_world.x + u.u_cloud_coverage_rotation.c[1][1] * pos_world.z};
    vec3<T> coverage_pos = {cp2.x, pos_world.y, cp2.y};
    T coverage = sample_cube(u, coverage_pos);
    coverage = coverage - T(0.25) * height_ratio + u.u_cloud_coverage_bias;

    T shape = mix(T(0.5), sample_shape3d(u, pos_world * u.u_cloud_shape_scale), u.u_cloud_shape_factor);

    T detail = T(0.5);

    if (u.u_cloud_shape_invert == T(1)) shape = T(1) - shape;

    density = (shape - T(0.2) * detail + (mix(T(-1.2), T(1.5), coverage))) * hc;
    density = density * T(50) - T(20);
    density = clamp(density, T(0), T(1));
    return density;
}

// :78-90
template <class T> inline T get_planet_shadow(vec3<T> pos, T /*planet_radius*/, vec3<T> sun_dir) {
    return smoothstep(T(-0.3), T(0.3), dot(normalize(pos), -sun_dir));
}

// pow(dp, 16.0) with the oracle's definition for dp <= 0 (see header)
template <class T> inline T pow16(T x) { return x <= T(0) ? T(0) : std::pow(x, T(16)); }

// :92-102
template <class T>
inline T get_light_cheap(vec3<T> pos_world, vec3<T> ray_dir, vec3<T> sun_dir, T alpha, const CloudSettings<T>& settings) {
    T height = length(pos_world) - settings.bottom_height;
    T height_ratio = height / (settings.top_height - settings.bottom_height);
    T light = height_ratio;
    T dp = dot(ray_dir, sun_dir);
    return light + std::max(pow16(dp), T(0)) * (T(1) - alpha);
}

// :104-151
template <class T>
inline T get_light_raymarched(const Uniforms<T>& u, vec3<T> pos0, vec3<T> sun_dir, T /*jitter*/, T /*alpha0*/,
                              const CloudSettings<T>& settings) {
    const int steps = 6;
    T reach = (settings.top_height - settings.bottom_height) * T(0.15);

    T pos0_height = length(pos0) - settings.bottom_height;
    T pos0_height_ratio = pos0_height / (settings.top_height - settings.bottom_height);

    T inv_steps = T(1) / T(steps);
    T step_len = reach * inv_steps;

    T alpha = T(0);
    for (int i = 0; i < steps; ++i) {
        vec3<T> dir = sun_dir;
        vec3<T> pos = pos0 + T(i) * step_len * dir;
        T density = get_density_full(u, pos, settings);  // both branches (:132-136) are the low-quality path
        density *= step_len * settings.density_scale;
        T transmittance = std::exp(-density);
        alpha += (T(1) - transmittance) * (T(1) - alpha);
        step_len *= T(1.2);
    }

    T light0 = pos0_height_ratio * T(0.2);
    return mix(T(1), light0, alpha);
}

// :153-167
template <class T>
inline T get_light(const Uniforms<T>& u, int light_mode, vec3<T> pos, vec3<T> ray_dir, vec3<T> sun_dir, T jitter, T alpha,
                   const CloudSettings<T>& settings) {
    T light = light_mode == B200ATMO_LIGHT_RAYMARCHED ? get_light_raymarched(u, pos, sun_dir, jitter, alpha, settings)
                                                      : get_light_cheap(pos, ray_dir, sun_dir, alpha, settings);
    T shadow_amount = get_planet_shadow(pos, T(1), sun_dir);
    light = light * mix(T(1), T(0.002), shadow_amount);
    return light;
}

// :175-247 — returns (total_light, alpha)
template <class T>
inline vec2<T> raymarch_cloud(const Uniforms<T>& u, int steps, int light_mode, vec3<T> ray_origin, vec3<T> ray_dir,
                              T t_begin, T t_end, T jitter, vec3<T> sun_dir, const CloudSettings<T>& settings) {
    T march_distance_space =
        T(0.5) * std::sqrt(T(1) - pow2(settings.ground_height / settings.top_height)) * settings.bottom_height;
    T march_distance_ground = T(3) * march_distance_space;
    T march_distance_transition_height_min = settings.bottom_height;
    T march_distance_transition_height_max = settings.top_height * T(1.05);

    T max_d = mix(march_distance_ground, march_distance_space,
                  smoothstep(march_distance_transition_height_min, march_distance_transition_height_max, length(ray_origin)));

    t_end = t_begin + std::min(t_end - t_begin, max_d);

    T inv_steps = T(1) / T(steps);
    T step_len = (t_end - t_begin) * inv_steps;

    T total_transmittance = T(1);
    T total_light = T(0);
    T alpha = T(0);
    vec3<T> pos = ray_origin + jitter * step_len * ray_dir + ray_dir * t_begin;

    for (int i = 0; i < steps; ++i) {
        T light = get_light(u, light_mode, pos, ray_dir, sun_dir, jitter, alpha, settings);
        T density = get_density_full(u, pos, settings);

        density *= settings.density_scale;

        T transmittance = std::exp(-density * step_len);
        total_transmittance *= transmittance;
        total_transmittance = std::max(total_transmittance, T(0.005));

        total_light += light * density * step_len * total_transmittance;

        alpha += (T(1) - transmittance) * (T(1) - alpha);

        pos = pos + ray_dir * step_len;
    }
    return {total_light, alpha};
}

// :249-324
template <class T>
inline void render_clouds(const Uniforms<T>& u, int steps, int light_mode, vec3<T>& out_albedo, T& out_alpha,
                          vec3<T> planet_center_view, vec3<T> ray_origin, vec3<T> ray_dir, T linear_depth,
                          const mat4<T>& inv_view_matrix, vec3<T> sun_dir, T jitter) {
    T clouds_bottom = u.u_planet_radius + u.u_cloud_bottom * u.u_atmosphere_height;
    T clouds_top = u.u_planet_radius + u.u_cloud_top * u.u_atmosphere_height;

    vec2<T> rs_clouds_top = ray_sphere(planet_center_view, clouds_top, ray_origin, ray_dir);

    if (rs_clouds_top.x != rs_clouds_top.y) {
        vec2<T> rs_clouds_bottom = ray_sphere(planet_center_view, clouds_bottom, ray_origin, ray_dir);

        vec2<T> cloud_rs = rs_clouds_top;
        cloud_rs.x = std::max(cloud_rs.x, T(0));
        cloud_rs.y = std::min(cloud_rs.y, linear_depth);

        if (cloud_rs.x < linear_depth && (linear_depth > rs_clouds_bottom.y || rs_clouds_bottom.x > T(0))) {
            mat4<T> view_to_model_matrix = mul(u.u_world_to_model_matrix, inv_view_matrix);
            vec4<T> o4 = mul(view_to_model_matrix, vec4<T>{ray_origin.x, ray_origin.y, ray_origin.z, T(1)});
            vec4<T> d4 = mul(view_to_model_matrix, vec4<T>{ray_dir.x, ray_dir.y, ray_dir.z, T(0)});
            vec4<T> s4 = mul(view_to_model_matrix, vec4<T>{sun_dir.x, sun_dir.y, sun_dir.z, T(0)});
            vec3<T> ray_origin_world = {o4.x, o4.y, o4.z};
            vec3<T> ray_dir_world = {d4.x, d4.y, d4.z};
            vec3<T> sun_dir_world = {s4.x, s4.y, s4.z};

            CloudSettings<T> cs;
            cs.bottom_height = clouds_bottom;
            cs.top_height = clouds_top;
            cs.density_scale = u.u_cloud_density_scale;
            cs.ground_height = u.u_planet_radius;

            vec2<T> cloud_rr = raymarch_cloud(u, steps, light_mode, ray_origin_world, ray_dir_world, cloud_rs.x, cloud_rs.y,
                                              jitter, sun_dir_world, cs);

            vec3<T> cloud_albedo = {cloud_rr.x, cloud_rr.x, cloud_rr.x};
            T cloud_alpha = cloud_rr.y;

            vec4<T> alpha_blended = blend_colors(vec4<T>{out_albedo.x, out_albedo.y, out_albedo.z, out_alpha},
                                                 vec4<T>{cloud_albedo.x, cloud_albedo.y, cloud_albedo.z, cloud_alpha});

            vec4<T> add_blended = {out_albedo.x + cloud_albedo.x * cloud_alpha, out_albedo.y + cloud_albedo.y * cloud_alpha,
                                   out_albedo.z + cloud_albedo.z * cloud_alpha, std::max(out_alpha, cloud_alpha)};

            T b = u.u_cloud_blend;
            out_albedo = {mix(alpha_blended.x, add_blended.x, b), mix(alpha_blended.y, add_blended.y, b),
                          mix(alpha_blended.z, add_blended.z, b)};
            out_alpha = mix(alpha_blended.w, add_blended.w, b);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// include/planet_atmosphere_main.gdshaderinc
// ------------------------------------------------------------------------------------------------
struct Variant {
    int scatter_model;  // B200ATMO_SCATTER_*
    int scatter_steps;  // ATMOSPHERE_RAYMARCH_STEPS
    int cloud_steps;    // CLOUDS_MAX_RAYMARCH_STEPS
    int light_mode;     // B200ATMO_LIGHT_*
};

// :144-196 — everything after the ray has been generated. `linear_depth` is the value of :138.
template <class T>
inline bool fragment_from_ray(const Uniforms<T>& u, const Variant& var, vec3<T> ray_origin, vec3<T> ray_dir, T linear_depth,
                              T jitter, vec3<T> planet_center_view, vec3<T> sun_center_view, const mat4<T>& inv_view_matrix,
                              vec3<T>* out_albedo, T* out_alpha) {
    T atmosphere_radius = u.u_planet_radius + u.u_atmosphere_height;
    vec2<T> rs_atmo = ray_sphere(planet_center_view, atmosphere_radius, ray_origin, ray_dir);

    if (rs_atmo.x != rs_atmo.y) {
        T t_begin = std::max(rs_atmo.x, T(0));
        T t_end = std::max(rs_atmo.y, T(0));

        vec2<T> rs_ground = ray_sphere(planet_center_view, u.u_planet_radius, ray_origin, ray_dir);
        T gd = T(10000000.0);
        if (rs_ground.x != rs_ground.y) gd = rs_ground.x;
        linear_depth = mix(linear_depth, gd, u.u_sphere_depth_factor);

        t_end = std::min(t_end, linear_depth);

        vec3<T> sun_dir = normalize(sun_center_view - planet_center_view);

        vec4<T> atmosphere;
        if (var.scatter_model == B200ATMO_SCATTER_V1)
            atmosphere = compute_atmosphere_v1(u, var.scatter_steps, ray_origin, ray_dir, planet_center_view, t_begin, t_end, sun_dir);
        else
            atmosphere = compute_atmosphere_v2(u, var.scatter_steps, ray_origin, ray_dir, planet_center_view, t_begin, t_end,
                                               linear_depth, sun_dir, jitter);

        *out_albedo = {atmosphere.x, atmosphere.y, atmosphere.z};
        *out_alpha = atmosphere.w;

        if (var.light_mode != B200ATMO_LIGHT_NONE)
            render_clouds(u, var.cloud_steps, var.light_mode, *out_albedo, *out_alpha, planet_center_view, ray_origin, ray_dir,
                          linear_depth, inv_view_matrix, sun_dir, jitter);
        return false;  // out_discard
    }
    *out_albedo = {T(0), T(0), T(0)};
    *out_alpha = T(0);
    return true;
}

// MODE_FAR: is this pixel covered by the node's proxy BoxMesh (edge `clip_box_size`, centred on the node) and does the
// cube's front face pass the depth test?  (engine rasteriser behaviour, restated; 0 = fullscreen quad)
template <class T>
inline bool far_box_covers(const Uniforms<T>& u, const mat4<T>& inv_view_matrix, T clip_box_size, vec3<T> ray_origin,
                           vec3<T> ray_dir, T linear_depth) {
    mat4<T> view_to_model_matrix = mul(u.u_world_to_model_matrix, inv_view_matrix);
    vec4<T> o4 = mul(view_to_model_matrix, vec4<T>{ray_origin.x, ray_origin.y, ray_origin.z, T(1)});
    vec4<T> d4 = mul(view_to_model_matrix, vec4<T>{ray_dir.x, ray_dir.y, ray_dir.z, T(0)});
    T bs = clip_box_size * T(0.5);
    vec2<T> rb = ray_box_intersection<T>({o4.x, o4.y, o4.z}, {d4.x, d4.y, d4.z}, {bs, bs, bs});
    if (rb.x == T(-1) && rb.y == T(-1)) return false;
    return rb.x > T(0) && rb.x < linear_depth;
}

// :69-104 — only the two varyings matter to a raymarcher
template <class T>
inline void atmosphere_vertex_varyings(const mat4<T>& model_matrix, const mat4<T>& view_matrix, vec3<T> u_sun_position,
                                       vec3<T>* planet_center_view, vec3<T>* sun_center_view) {
    vec4<T> world_pos = mul(model_matrix, vec4<T>{T(0), T(0), T(0), T(1)});
    vec4<T> pc = mul(view_matrix, world_pos);
    vec4<T> sc = mul(view_matrix, vec4<T>{u_sun_position.x, u_sun_position.y, u_sun_position.z, T(1)});
    *planet_center_view = {pc.x, pc.y, pc.z};
    *sun_center_view = {sc.x, sc.y, sc.z};
}

// :106-142 — per-pixel ray generation from the depth texture. screen_uv = pixel centre / viewport size.
template <class T>
inline void fragment_make_ray(mat4<T> inv_projection_matrix, mat4<T> inv_view_matrix, bool double_precision, T nonlinear_depth,
                              T screen_u, T screen_v, vec3<T>* ray_origin, vec3<T>* ray_dir, T* linear_depth) {
    if (double_precision) {  // :118-125
        inv_view_matrix.c[3][0] *= T(-1);
        inv_view_matrix.c[3][1] *= T(-1);
        inv_view_matrix.c[3][2] *= T(-1);
    }
    vec3<T> ndc = {screen_u * T(2) - T(1), screen_v * T(2) - T(1), nonlinear_depth};
    vec4<T> view_coords = mul(inv_projection_matrix, vec4<T>{ndc.x, ndc.y, ndc.z, T(1)});
    vec4<T> world_coords = mul(inv_view_matrix, view_coords);
    vec3<T> pos_world = vec3<T>{world_coords.x, world_coords.y, world_coords.z} / world_coords.w;
    vec4<T> cam4 = mul(inv_view_matrix, vec4<T>{T(0), T(0), T(0), T(1)});
    vec3<T> cam_pos_world = {cam4.x, cam4.y, cam4.z};
    *linear_depth = distance(cam_pos_world, pos_world);
    *ray_origin = {T(0), T(0), T(0)};
    *ray_dir = normalize(vec3<T>{view_coords.x, view_coords.y, view_coords.z} - *ray_origin);
}

// ------------------------------------------------------------------------------------------------
// NoiseCubemap._generate_images — noise_cubemap.gd:101-140.
// The noise itself (`noise.get_noise_3dv`, Godot's FastNoiseLite) is engine code outside the reference tree: its
// content is DEFINED by this repo's include/b200atmo.h ("b200 gradient fBm v1"); what follows the reference is the
// texel -> direction mapping (:110-128), `pos * scale`, density = 0.5 + 0.5*n (:130) and the L8 store (:134).
// fp32 only (the generator's output is bytes; parity is bit-exact).
// ------------------------------------------------------------------------------------------------
namespace noisegen {
inline uint32_t hash3(int32_t ix, int32_t iy, int32_t iz, uint32_t seed) {
    uint32_t h = seed;
    h ^= uint32_t(ix) * 0x8da6b343u;
    h ^= uint32_t(iy) * 0xd8163841u;
    h ^= uint32_t(iz) * 0xcb1ab31fu;
    h *= 0x9e3779b1u;
    h ^= h >> 15;
    h *= 0x85ebca6bu;
    h ^= h >> 13;
    return h;
}
inline float grad(uint32_t hash, float x, float y, float z) {  // Perlin's 12 edge directions (+4 repeats)
    const uint32_t h = hash & 15u;
    const float u = h < 8u ? x : y;
    const float v = h < 4u ? y : ((h == 12u || h == 14u) ? x : z);
    return ((h & 1u) ? -u : u) + ((h & 2u) ? -v : v);
}
inline float fade(float t) { return t * t * t * (t * (t * 6.0f - 15.0f) + 10.0f); }
inline float lerp(float a, float b, float t) { return a + (b - a) * t; }
inline float noise3(float x, float y, float z, uint32_t seed) {
    const float x0 = std::floor(x), y0 = std::floor(y), z0 = std::floor(z);
    const int32_t ix = int32_t(x0), iy = int32_t(y0), iz = int32_t(z0);
    const float fx = x - x0, fy = y - y0, fz = z - z0;
    const float u = fade(fx), v = fade(fy), w = fade(fz);
    float c[2][2][2];
    for (int dz = 0; dz < 2; ++dz)
        for (int dy = 0; dy < 2; ++dy)
            for (int dx = 0; dx < 2; ++dx)
                c[dz][dy][dx] = grad(hash3(ix + dx, iy + dy, iz + dz, seed), fx - float(dx), fy - float(dy), fz - float(dz));
    const float a = lerp(lerp(c[0][0][0], c[0][0][1], u), lerp(c[0][1][0], c[0][1][1], u), v);
    const float b = lerp(lerp(c[1][0][0], c[1][0][1], u), lerp(c[1][1][0], c[1][1][1], u), v);
    return lerp(a, b, w);
}
inline float fbm(float x, float y, float z, const B200AtmoNoise& n) {
    float freq = n.frequency, amp = 1.0f, sum = 0.0f, norm = 0.0f;
    for (int o = 0; o < n.octaves; ++o) {
        sum = sum + amp * noise3(x * freq, y * freq, z * freq, uint32_t(n.seed) + uint32_t(o) * 0x632be5abu);
        norm = norm + amp;
        freq = freq * n.lacunarity;
        amp = amp * n.gain;
    }
    return sum / norm;
}
// noise_cubemap.gd:101-140 (without the mipmaps: LOD 0 only)
inline void generate_images(const B200AtmoNoise& noise, int resolution, const float scale[3], uint8_t* out) {
    const float half = 0.5f * float(resolution);  // half_resolution_2d
    for (int side = 0; side < 6; ++side)
        for (int y = 0; y < resolution; ++y)
            for (int x = 0; x < resolution; ++x) {
                const float p2x = (float(x) + 0.5f) / half - 1.0f;
                const float p2y = (float(resolution - y - 1) + 0.5f) / half - 1.0f;
                vec3<float> pos = normalize(vec3<float>{1.0f, p2y, -p2x});  // +X
                switch (side) {
                    case 0: pos = {pos.x, pos.y, pos.z}; break;
                    case 1: pos = {-pos.x, pos.y, -pos.z}; break;
                    case 2: pos = {-pos.z, pos.x, -pos.y}; break;
                    case 3: pos = {-pos.z, -pos.x, pos.y}; break;
                    case 4: pos = {-pos.z, pos.y, pos.x}; break;
                    case 5: pos = {pos.z, pos.y, -pos.x}; break;
                }
                const float density = 0.5f + 0.5f * fbm(pos.x * scale[0], pos.y * scale[1], pos.z * scale[2], noise);
                // Image.set_pixel(x, y, Color(d,d,d)) on FORMAT_L8: uint8(CLAMP(v*255, 0, 255))
                const float v = std::min(std::max(density * 255.0f, 0.0f), 255.0f);
                out[(size_t(side) * resolution + y) * resolution + x] = uint8_t(int(v));
            }
}
// NoiseCubemap._generate_importable_image (noise_cubemap.gd:143-155): 3 x 2 atlas, side = x + 3*y
inline void importable_image(const uint8_t* faces, int res, uint8_t* atlas) {
    const int count_x = 3, count_y = 2;
    for (int ay = 0; ay < count_y; ++ay)
        for (int ax = 0; ax < count_x; ++ax) {
            const int side = ax + ay * count_x;
            for (int y = 0; y < res; ++y)
                std::memcpy(atlas + (size_t(ay * res + y) * count_x + ax) * res, faces + (size_t(side) * res + y) * res, size_t(res));
        }
}
}  // namespace noisegen

}  // namespace oracle

// TEST INFRASTRUCTURE ONLY — GLSL-ES-3.0 types and built-ins as plain C++ (fp32), so that the REFERENCE'S OWN shader
// sources (addons/zylann.atmosphere/shaders/**/*.gdshader*, read from /root/reference at build time and passed through
// the purely syntactic rewrite of oracle/ref/build_ref.py) compile with g++ and can be executed on the CPU. Nothing of
// the reference is restated in this file: it only supplies what the Godot shader compiler / GPU would supply.
//
// Where GLSL leaves a built-in's rounding or an input domain open, the definition is the one documented in
// oracle/atmo_oracle.hpp (same formulas), so the hand-written oracle and the compiled reference can be compared bit for
// bit: dot = left-to-right sum of products, length = sqrt(dot), normalize(v) = v / length(v), mix(a,b,t) = a*(1-t)+b*t,
// clamp = min(max(x,lo),hi), smoothstep per the GLSL spec, mat*vec = column sums left to right, pow(x<=0, y) = 0,
// exp/sqrt/pow = libm fp32. Texture fetches go through the oracle-defined sampling conventions (oracle::sample_*):
// the reference does not define filtering (engine / hardware behaviour), so there is nothing of it to pin there.
// Build with -ffp-contract=off (no FMA contraction), like the oracle.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>

#include <atomic>
#include <thread>
#include <vector>

#include "../atmo_oracle.hpp"

#ifndef GLSL_REAL
#define GLSL_REAL float
#endif

namespace glsl {

typedef GLSL_REAL real;   // float: the shader's precision (THE reference); double: its fp64 twin (rounding-error bound)
typedef unsigned int uint;

struct vec2 {
    union { real x, r; };
    union { real y, g; };
    vec2() : x(0.f), y(0.f) {}
    explicit vec2(real s) : x(s), y(s) {}
    vec2(real x_, real y_) : x(x_), y(y_) {}
};
struct vec3 {
    union { real x, r; };
    union { real y, g; };
    union { real z, b; };
    vec3() : x(0.f), y(0.f), z(0.f) {}
    explicit vec3(real s) : x(s), y(s), z(s) {}
    vec3(real x_, real y_, real z_) : x(x_), y(y_), z(z_) {}
    vec3(vec2 v, real z_) : x(v.x), y(v.y), z(z_) {}
    vec2 xy() const { return vec2(x, y); }
    vec2 xz() const { return vec2(x, z); }
    vec3 xyz() const { return *this; }
    vec3 rgb() const { return *this; }
};
struct vec4 {
    union { real x, r; };
    union { real y, g; };
    union { real z, b; };
    union { real w, a; };
    vec4() : x(0.f), y(0.f), z(0.f), w(0.f) {}
    explicit vec4(real s) : x(s), y(s), z(s), w(s) {}
    vec4(real x_, real y_, real z_, real w_) : x(x_), y(y_), z(z_), w(w_) {}
    vec4(vec3 v, real w_) : x(v.x), y(v.y), z(v.z), w(w_) {}
    vec4(vec2 v, real z_, real w_) : x(v.x), y(v.y), z(z_), w(w_) {}
    vec2 xy() const { return vec2(x, y); }
    vec3 xyz() const { return vec3(x, y, z); }
    vec3 rgb() const { return vec3(x, y, z); }
    real& operator[](int i) { return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w)); }
    real operator[](int i) const { return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w)); }
};
struct ivec2 {
    int x, y;
    explicit ivec2(int s) : x(s), y(s) {}
    ivec2(int x_, int y_) : x(x_), y(y_) {}
    explicit ivec2(vec2 v) : x(int(v.x)), y(int(v.y)) {}   // real -> int conversion truncates toward zero
};
inline ivec2 operator&(ivec2 a, ivec2 b) { return ivec2(a.x & b.x, a.y & b.y); }

// ---- componentwise arithmetic ------------------------------------------------------------------------------------
inline vec2 operator+(vec2 a, vec2 b) { return vec2(a.x + b.x, a.y + b.y); }
inline vec2 operator-(vec2 a, vec2 b) { return vec2(a.x - b.x, a.y - b.y); }
inline vec2 operator*(vec2 a, vec2 b) { return vec2(a.x * b.x, a.y * b.y); }
inline vec2 operator/(vec2 a, vec2 b) { return vec2(a.x / b.x, a.y / b.y); }
inline vec2 operator-(vec2 a) { return vec2(-a.x, -a.y); }
inline vec2 operator*(vec2 a, real s) { return vec2(a.x * s, a.y * s); }
inline vec2 operator*(real s, vec2 a) { return vec2(s * a.x, s * a.y); }
inline vec2 operator/(vec2 a, real s) { return vec2(a.x / s, a.y / s); }
inline vec2 operator+(vec2 a, real s) { return vec2(a.x + s, a.y + s); }
inline vec2 operator-(vec2 a, real s) { return vec2(a.x - s, a.y - s); }
inline vec2& operator+=(vec2& a, vec2 b) { a = a + b; return a; }
inline vec2& operator*=(vec2& a, real s) { a = a * s; return a; }

inline vec3 operator+(vec3 a, vec3 b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline vec3 operator-(vec3 a, vec3 b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline vec3 operator*(vec3 a, vec3 b) { return vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline vec3 operator/(vec3 a, vec3 b) { return vec3(a.x / b.x, a.y / b.y, a.z / b.z); }
inline vec3 operator-(vec3 a) { return vec3(-a.x, -a.y, -a.z); }
inline vec3 operator*(vec3 a, real s) { return vec3(a.x * s, a.y * s, a.z * s); }
inline vec3 operator*(real s, vec3 a) { return vec3(s * a.x, s * a.y, s * a.z); }
inline vec3 operator/(vec3 a, real s) { return vec3(a.x / s, a.y / s, a.z / s); }
inline vec3 operator/(real s, vec3 a) { return vec3(s / a.x, s / a.y, s / a.z); }
inline vec3 operator+(vec3 a, real s) { return vec3(a.x + s, a.y + s, a.z + s); }
inline vec3 operator-(vec3 a, real s) { return vec3(a.x - s, a.y - s, a.z - s); }
inline vec3& operator+=(vec3& a, vec3 b) { a = a + b; return a; }
inline vec3& operator-=(vec3& a, vec3 b) { a = a - b; return a; }
inline vec3& operator*=(vec3& a, vec3 b) { a = a * b; return a; }
inline vec3& operator*=(vec3& a, real s) { a = a * s; return a; }

inline vec4 operator+(vec4 a, vec4 b) { return vec4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
inline vec4 operator-(vec4 a, vec4 b) { return vec4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
inline vec4 operator*(vec4 a, vec4 b) { return vec4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
inline vec4 operator/(vec4 a, vec4 b) { return vec4(a.x / b.x, a.y / b.y, a.z / b.z, a.w / b.w); }
inline vec4 operator*(vec4 a, real s) { return vec4(a.x * s, a.y * s, a.z * s, a.w * s); }
inline vec4 operator*(real s, vec4 a) { return vec4(s * a.x, s * a.y, s * a.z, s * a.w); }
inline vec4 operator/(vec4 a, real s) { return vec4(a.x / s, a.y / s, a.z / s, a.w / s); }

// ---- matrices (column-major, m[col][row]) ------------------------------------------------------------------------
struct mat2 {
    vec2 c[2];
    mat2() { c[0] = vec2(1.f, 0.f); c[1] = vec2(0.f, 1.f); }
    vec2& operator[](int i) { return c[i]; }
    const vec2& operator[](int i) const { return c[i]; }
};
struct mat4 {
    vec4 c[4];
    mat4() { c[0] = vec4(1.f, 0.f, 0.f, 0.f); c[1] = vec4(0.f, 1.f, 0.f, 0.f); c[2] = vec4(0.f, 0.f, 1.f, 0.f); c[3] = vec4(0.f, 0.f, 0.f, 1.f); }
    vec4& operator[](int i) { return c[i]; }
    const vec4& operator[](int i) const { return c[i]; }
};
inline vec2 operator*(const mat2& m, vec2 v) { return vec2(m[0].x * v.x + m[1].x * v.y, m[0].y * v.x + m[1].y * v.y); }
inline vec4 operator*(const mat4& m, vec4 v) {
    return vec4(m[0].x * v.x + m[1].x * v.y + m[2].x * v.z + m[3].x * v.w, m[0].y * v.x + m[1].y * v.y + m[2].y * v.z + m[3].y * v.w,
                m[0].z * v.x + m[1].z * v.y + m[2].z * v.z + m[3].z * v.w, m[0].w * v.x + m[1].w * v.y + m[2].w * v.z + m[3].w * v.w);
}
inline mat4 operator*(const mat4& a, const mat4& b) {
    mat4 r;
    for (int col = 0; col < 4; ++col)
        for (int row = 0; row < 4; ++row)
            r[col][row] = a[0][row] * b[col][0] + a[1][row] * b[col][1] + a[2][row] * b[col][2] + a[3][row] * b[col][3];
    return r;
}

// ---- built-in functions --------------------------------------------------------------------------------------------
inline real max(real a, real b) { return std::max(a, b); }
inline real min(real a, real b) { return std::min(a, b); }
inline real abs(real a) { return std::fabs(a); }
inline vec3 abs(vec3 a) { return vec3(std::fabs(a.x), std::fabs(a.y), std::fabs(a.z)); }
inline real sqrt(real a) { return std::sqrt(a); }
inline real exp(real a) { return std::exp(a); }
inline vec3 exp(vec3 a) { return vec3(std::exp(a.x), std::exp(a.y), std::exp(a.z)); }
inline real pow(real x, real y) { return x <= 0.0f ? 0.0f : std::pow(x, y); }   // GLSL: undefined for x < 0 (see header)
inline real clamp(real x, real lo, real hi) { return std::min(std::max(x, lo), hi); }
inline vec3 clamp(vec3 v, vec3 lo, vec3 hi) { return vec3(clamp(v.x, lo.x, hi.x), clamp(v.y, lo.y, hi.y), clamp(v.z, lo.z, hi.z)); }
inline real mix(real a, real b, real t) { return a * (1.0f - t) + b * t; }
inline vec3 mix(vec3 a, vec3 b, real t) { return vec3(mix(a.x, b.x, t), mix(a.y, b.y, t), mix(a.z, b.z, t)); }
inline vec4 mix(vec4 a, vec4 b, real t) { return vec4(mix(a.x, b.x, t), mix(a.y, b.y, t), mix(a.z, b.z, t), mix(a.w, b.w, t)); }
inline real smoothstep(real e0, real e1, real x) {
    real t = clamp((x - e0) / (e1 - e0), 0.0f, 1.0f);
    return t * t * (3.0f - 2.0f * t);
}
inline real dot(vec2 a, vec2 b) { return a.x * b.x + a.y * b.y; }
inline real dot(vec3 a, vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline real length(vec2 a) { return std::sqrt(dot(a, a)); }
inline real length(vec3 a) { return std::sqrt(dot(a, a)); }
inline real distance(vec3 a, vec3 b) { return length(a - b); }
inline vec3 normalize(vec3 a) { return a / length(a); }
// fp32 bit pattern of the argument; the fp64 twin also keeps the unrounded value (the RGBA8 viewport encoding of
// optical_depth.gdshader is an fp32 container, the twin's bake reads the value from here instead)
inline real& last_float_bits_arg() {
    static thread_local real v = real(0);
    return v;
}
inline uint floatBitsToUint(real f) {
    last_float_bits_arg() = f;
    const float g = float(f);
    uint u;
    std::memcpy(&u, &g, 4);
    return u;
}

// ---- samplers: what the engine binds; fetches follow the oracle-defined conventions ----------------------------------
struct sampler2D {
    enum Kind { UNSET, LUT_F32, DEPTH_F32, L8 } kind = UNSET;
    const real* lut = nullptr;    // LUT: 256 x 256 (linear filter, clamp)
    const float* f32 = nullptr;   // DEPTH: w x h fp32 as handed over by the engine (nearest)
    const uint8_t* u8 = nullptr;  // L8: w x h (texelFetch only)
    int w = 0, h = 0;
};
struct sampler3D {
    const oracle::Uniforms<real>* tex = nullptr;   // holds the shape texture pointers (oracle::sample_shape3d)
};
struct samplerCube {
    const oracle::Uniforms<real>* tex = nullptr;   // holds the padded cube (oracle::cube_build_padded layout)
};
inline vec4 texture(const sampler2D& s, vec2 uv) {
    if (s.kind == sampler2D::LUT_F32) {
        const real v = oracle::sample_lut<real>(s.lut, uv.x, uv.y);
        return vec4(v, 0.f, 0.f, 1.f);
    }
    if (s.kind == sampler2D::DEPTH_F32) {   // hint_depth_texture: nearest texel under the fragment
        const int ix = std::min(std::max(int(std::floor(uv.x * real(s.w))), 0), s.w - 1);
        const int iy = std::min(std::max(int(std::floor(uv.y * real(s.h))), 0), s.h - 1);
        return vec4(real(s.f32[size_t(iy) * s.w + ix]), 0.f, 0.f, 1.f);
    }
    return vec4(1.f);   // unset sampler = white
}
inline vec4 texelFetch(const sampler2D& s, ivec2 p, int /*lod*/) {
    if (s.kind != sampler2D::L8) return vec4(0.f, 0.f, 0.f, 1.f);   // no blue-noise texture bound: jitter 0 (oracle convention)
    const real v = real(s.u8[size_t(p.y) * s.w + p.x]) / 255.0f;
    return vec4(v, v, v, 1.f);
}
inline vec4 texture(const sampler3D& s, vec3 p) {
    const real v = s.tex ? oracle::sample_shape3d<real>(*s.tex, oracle::vec3<real>{p.x, p.y, p.z}) : 1.0f;
    return vec4(v, v, v, 1.f);
}
inline vec4 texture(const samplerCube& s, vec3 d) {
    const real v = s.tex ? oracle::sample_cube<real>(*s.tex, oracle::vec3<real>{d.x, d.y, d.z}) : 1.0f;
    return vec4(v, v, v, 1.f);
}

}  // namespace glsl

// Brought into each generated namespace by using-DECLARATIONS (they hide ::sqrt, ::exp, ::pow ... of <cmath> for
// unqualified calls, which a using-directive would not).
#define GLSL_USING                                                                                                     \
    using glsl::real; using glsl::uint; using glsl::vec2; using glsl::vec3; using glsl::vec4; using glsl::ivec2; using glsl::mat2; using glsl::mat4;     \
    using glsl::sampler2D; using glsl::sampler3D; using glsl::samplerCube; using glsl::max; using glsl::min; using glsl::abs;          \
    using glsl::sqrt; using glsl::exp; using glsl::pow; using glsl::clamp; using glsl::mix; using glsl::smoothstep; using glsl::dot;   \
    using glsl::length; using glsl::distance; using glsl::normalize; using glsl::floatBitsToUint; using glsl::texture;                \
    using glsl::texelFetch;

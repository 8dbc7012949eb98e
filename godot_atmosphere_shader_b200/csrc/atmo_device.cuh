// Device functions of the atmosphere hot path (sm_100a). Included by atmo_kernels.cu.
//
// NUMERIC POLICY. The translation unit is compiled with -fmad=false, so every expression written
// with plain * + - / sqrtf is IEEE fp32 with no FMA contraction, i.e. bit-identical to the scalar
// shader arithmetic. That "exact" style is used wherever rounding is amplified downstream:
//   * per-ray set-up and every hit / miss / visibility decision (ray_sphere, t_begin/t_end, cloud test)
//   * ray positions (pos0, pos += dir*step) — height = |pos-C| - R cancels 1-2 digits
//   * the cloud height curve (its output is multiplied by 50 and thresholded)
//   * texture coordinates (a 1e-7 coordinate error times a texel-to-texel slope times 135 is visible)
//   * the LUT bake
// FMA (fmaf) and MUFU approximations (ex2/rsqrt/rcp.approx) are requested explicitly, only in the
// accumulation arithmetic of the hot loops, where errors do not amplify.
//
// The file also compiles as plain C++ (tests/hostsim) so host-side logic tests can run without a GPU;
// that build is test infrastructure and is never loaded by the product.
//
// Reference being implemented (addons/zylann.atmosphere/shaders/...):
//   include/planet_atmosphere_main.gdshaderinc:106-197  atmosphere_fragment      -> shade_ray / make_ray
//   include/atmosphere_funcs_v2.gdshaderinc:14-101       compute_atmosphere_v2    -> scatter_v2
//   include/atmosphere_funcs_v1.gdshaderinc:15-63        compute_atmosphere       -> scatter_v1
//   include/cloud_funcs.gdshaderinc:25-324               render_clouds & friends  -> render_clouds
#pragma once

#include "atmo_internal.h"

// B200ATMO_LITERAL: audit build — every round-2 shortcut that is claimed to be bit-identical is replaced by the literal form
// (two-instruction expressions instead of the exact FMA folds, the plain shell test instead of hc_min, the shader's seventh
// density evaluation in the light march, the literal density bound test and shape mix). The CPU test suite compiles the
// host build both ways and requires identical bits; profiles/build_variants.sh can do the same for the GPU library.
#ifdef B200ATMO_LITERAL
#define B200ATMO_EXACT_FOLDS 0
#define B200ATMO_NO_HCMIN 1
#define B200ATMO_LIGHT_RESAMPLE_FIRST 1
#endif
// tuning knobs (defaults chosen by profiles/tune_scatter.sh on B200)
#ifndef B200ATMO_SCATTER_UNROLL
#define B200ATMO_SCATTER_UNROLL 8
#endif
// unroll factors of the cloud loops, chosen on B200 (profiles/r02/tune_clouds.txt): the 6-step light march by 3 (-2.8 % on
// cfg4), the cloud march by 4 when the light is cheap (-3.6 % on cfg3) and not at all around the raymarched light
#ifndef B200ATMO_LIGHT_UNROLL
#define B200ATMO_LIGHT_UNROLL 3
#endif
#ifndef B200ATMO_CLOUD_UNROLL_CHEAP
#define B200ATMO_CLOUD_UNROLL_CHEAP 4
#endif
#ifndef B200ATMO_CLOUD_UNROLL_RM
#define B200ATMO_CLOUD_UNROLL_RM 1
#endif
#define B200_PRAGMA(x) _Pragma(#x)
#ifdef __CUDACC__
#define B200_DEV __device__ __forceinline__
#define B200_DEV_RARE __device__ __noinline__
#define B200_UNROLL(n) B200_PRAGMA(unroll n)
__device__ __forceinline__ float4 ldg4(const float4* p) { return __ldg(p); }
#else
#define B200_UNROLL(n)
// ---- host simulation shims (tests/hostsim only) ----
#include <cmath>
#include <cstring>
#define B200_DEV static inline
#define B200_DEV_RARE static inline
struct float4 {
    float x, y, z, w;
};
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline float __ldg(const float* p) { return *p; }
static inline float4 ldg4(const float4* p) { return *p; }
static inline float __saturatef(float x) { return x != x ? 0.0f : (x < 0.0f ? 0.0f : (x > 1.0f ? 1.0f : x)); }
static inline int __float_as_int(float f) {
    int i;
    std::memcpy(&i, &f, 4);
    return i;
}
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
#endif

namespace b200atmo {

// ------------------------------------------------------------------------------------------------
// small vector helpers (exact arithmetic: no fmaf here)
// ------------------------------------------------------------------------------------------------
struct f3 {
    float x, y, z;
};
struct f2 {
    float x, y;
};
B200_DEV f3 mk3(float x, float y, float z) { return f3{x, y, z}; }
B200_DEV f3 ld3(const float* p) { return f3{p[0], p[1], p[2]}; }
B200_DEV f3 operator+(f3 a, f3 b) { return f3{a.x + b.x, a.y + b.y, a.z + b.z}; }
B200_DEV f3 operator-(f3 a, f3 b) { return f3{a.x - b.x, a.y - b.y, a.z - b.z}; }
B200_DEV f3 operator*(f3 a, float s) { return f3{a.x * s, a.y * s, a.z * s}; }
B200_DEV f3 operator*(float s, f3 a) { return f3{s * a.x, s * a.y, s * a.z}; }
B200_DEV float dot3(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
B200_DEV float mixf(float a, float b, float t) { return a * (1.0f - t) + b * t; }  // GLSL mix
B200_DEV float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
B200_DEV float smoothstepf(float e0, float e1, float x) {
    float t = clampf((x - e0) / (e1 - e0), 0.0f, 1.0f);
    return t * t * (3.0f - 2.0f * t);
}
// mat4 (column-major) * vec4 in the shader's summation order
B200_DEV void mat4_mul(const float* m, float x, float y, float z, float w, float out[4]) {
    for (int r = 0; r < 4; ++r) out[r] = m[0 + r] * x + m[4 + r] * y + m[8 + r] * z + m[12 + r] * w;
}

// ------------------------------------------------------------------------------------------------
// approximate hardware ops (MUFU), used only where stated
// ------------------------------------------------------------------------------------------------
#ifdef __CUDACC__
B200_DEV float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
B200_DEV float rsqrt_approx(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
B200_DEV float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
#else
B200_DEV float ex2_approx(float x) { return exp2f(x); }
B200_DEV float rsqrt_approx(float x) { return 1.0f / sqrtf(x); }
B200_DEV float rcp_approx(float x) { return 1.0f / x; }
#endif

// sqrt(x) and 1/sqrt(x) from ONE MUFU op: rsqrt.approx + one Newton step on the root with an exact
// (fma) residual. The pre-rounding error is ~2^-44, so the result equals IEEE sqrtf except when the
// true root lies within that distance of a rounding midpoint (~1e-6 of inputs, then 1 ulp off).
B200_DEV float sqrt_refined(float x, float& inv) {
    inv = rsqrt_approx(x);
    const float r = x * inv;
    const float e = fmaf(-r, r, x);
    return fmaf(e, 0.5f * inv, r);
}
// a/b given rb ~ 1/b (correctly rounded or 1-ulp approximate): q0 = a*rb, one exact-residual
// correction. Equals the IEEE quotient except for the same ~1e-6 near-midpoint cases.
B200_DEV float div_refined(float a, float b, float rb) {
    const float q = a * rb;
    const float r = fmaf(-q, b, a);
    return fmaf(r, rb, q);
}
// floor(x) for |x| < 2^22 without the conversion pipe: round-to-nearest of x-0.5 through the
// 1.5*2^23 magic constant. Returns the integer part and writes the fraction. At exact integers it
// may return (x-1, frac=1): the same point of a continuous interpolant.
constexpr float kMagic = 12582912.0f;  // 1.5 * 2^23
constexpr int kMagicBits = 0x4B400000;
B200_DEV int floor_frac(float x, float& frac) {
    const float xm = (x - 0.5f) + kMagic;
    const float x0 = xm - kMagic;
    frac = x - x0;
    return __float_as_int(xm) - kMagicBits;
}
B200_DEV float lerp_fma(float a, float b, float t) { return fmaf(b - a, t, a); }  // a + (b-a)*t

// ------------------------------------------------------------------------------------------------
// include/util.gdshaderinc:20-40 (exact)
// ------------------------------------------------------------------------------------------------
B200_DEV f2 ray_sphere(f3 center, float radius, f3 ray_origin, f3 ray_dir) {
    const f3 oc = ray_origin - center;
    const float b = dot3(oc, ray_dir);
    const f3 qc = oc - b * ray_dir;
    float h = radius * radius - dot3(qc, qc);
    if (h < 0.0f) return f2{1000000.0f, 1000000.0f};
    h = sqrtf(h);
    return f2{-b - h, -b + h};
}

// ------------------------------------------------------------------------------------------------
// texture fetches in the hot loops (textures are fp32 copies with a one-texel apron, see atmo_kernels.cu)
// ------------------------------------------------------------------------------------------------
// ---- cell layouts of the two noise textures ------------------------------------------------------
// Built once at upload (atmo_kernels.cu) from the padded fp32 copies. A cell holds the texels of one
// interpolation footprint with the x-differences precomputed, so a fetch is 1 (cube) / 2 (3D) 16-byte loads and the
// lerps a+(b-a)*t become fma(b-a, t, a) with the SAME fp32 value of (b-a): results are bit-identical to lerping the
// raw texels.
//   cube cell (f, yi, xi)      = (t00, t10-t00, t01, t11-t01)                        yi, xi in [0, res]
//   shape cell (zi, yi, xi)[0] = (t000, t100-t000, t010, t110-t010)   [1] = same at z+1   indices in [0, n]
B200_DEV float4 make_cube_cell(const float* __restrict__ cube_pad, int res, int f, int yi, int xi) {
    const int pr = res + 2;
    const float* p = cube_pad + (size_t(f) * pr + yi) * pr + xi;
    return make_float4(p[0], p[1] - p[0], p[pr], p[pr + 1] - p[pr]);
}
B200_DEV float4 make_shape_cell(const float* __restrict__ shp, int nx, int ny, int zi, int yi, int xi) {
    const int px = nx + 2, pxy = px * (ny + 2);
    const float* p = shp + size_t(zi) * pxy + yi * px + xi;   // zi may already include the +1 of the second half
    return make_float4(p[0], p[1] - p[0], p[px], p[px + 1] - p[px]);
}

// Exact single-instruction forms of two-operation expressions in which one operation cannot round:
//   0.5*(q + 1) == fma(q, 0.5, 0.5): scaling by a power of two commutes with rounding, so rn(q+1)/2 == rn((q+1)/2)
//   2*h - 1     == fma(h, 2, -1),   c - 0.25*h == fma(h, -0.25, c): the product is exact, one rounding either way
// (no subnormals in range). Bit-identical to the two-instruction forms, one issue slot less each.
#ifndef B200ATMO_EXACT_FOLDS
#define B200ATMO_EXACT_FOLDS 1
#endif
B200_DEV float half_of_sum1(float q) {
#if B200ATMO_EXACT_FOLDS
    return fmaf(q, 0.5f, 0.5f);
#else
    return 0.5f * (q + 1.0f);
#endif
}
B200_DEV float twice_minus1(float h) {
#if B200ATMO_EXACT_FOLDS
    return fmaf(h, 2.0f, -1.0f);
#else
    return 2.0f * h - 1.0f;
#endif
}
B200_DEV float minus_quarter(float c, float h) {
#if B200ATMO_EXACT_FOLDS
    return fmaf(h, -0.25f, c);
#else
    return c - 0.25f * h;
#endif
}
// coord*n - 0.5 (texel coordinate of a normalised coordinate). When n is a power of two the product is exact, so the fused
// form rounds once to the same value: bit-identical, one issue slot less. POW2 = every dimension of the coverage cube and
// of the shape volume is a power of two (the usual case: Godot's NoiseTexture3D is 64^3, the NoiseCubemap 256^2); the
// host selects the kernel instantiation (bit 2 of the LIGHT template argument, kLightPow2).
constexpr int kLightPow2 = 4;
template <bool POW2> B200_DEV float texel_coord(float s, float n) {
    if (POW2) return fmaf(s, n, -0.5f);
    return s * n - 0.5f;
}
// x - floor(x) for |x| < 2^22. B200ATMO_MAGIC_WRAP: floor by the magic-constant add instead of FRND.FLOOR (a conversion-
// pipe instruction). At exact integers the magic form may return 1.0 where floorf gives 0.0: the same point of the
// periodic interpolant (cell n instead of cell 0 of the repeat-padded volume, same texels, same weights).
B200_DEV float wrap01(float x) {
#ifdef B200ATMO_MAGIC_WRAP
    const float f = ((x - 0.5f) + kMagic) - kMagic;
    return x - f;
#else
    return x - floorf(x);
#endif
}

// texture(u_cloud_coverage_cubemap, d).r — cloud_funcs:45 (seamless bilinear, LOD 0)
template <bool POW2> B200_DEV float sample_cube(const float4* __restrict__ cells, int res, float x, float y, float z) {
    const float ax = fabsf(x), ay = fabsf(y), az = fabsf(z);
    int f;
    float sc, tc, ma;
    if (ax >= ay && ax >= az) { f = x >= 0.0f ? 0 : 1; ma = ax; sc = x >= 0.0f ? -z : z; tc = -y; }
    else if (ay >= az)        { f = y >= 0.0f ? 2 : 3; ma = ay; sc = x; tc = y >= 0.0f ? z : -z; }
    else                      { f = z >= 0.0f ? 4 : 5; ma = az; sc = z >= 0.0f ? x : -x; tc = -y; }
    const float inv_ma = rcp_approx(ma);
    // s = 0.5*(sc/ma + 1); xf = s*res - 0.5   (exact op order; the two divisions share one MUFU.RCP)
    const float s = half_of_sum1(div_refined(sc, ma, inv_ma));
    const float t = half_of_sum1(div_refined(tc, ma, inv_ma));
    float fx, fy;
    const float resf = float(res);
    const int xi = floor_frac(texel_coord<POW2>(s, resf), fx) + 1;   // in [0, res] by construction (|sc|,|tc| <= ma)
    const int yi = floor_frac(texel_coord<POW2>(t, resf), fy) + 1;
    const unsigned rc = unsigned(res + 1);
    unsigned idx = (unsigned(f) * rc + unsigned(yi)) * rc + unsigned(xi);
    idx = min(idx, 6u * rc * rc - 1u);                            // NaN guard only
    const float4 q = ldg4(cells + idx);
    const float c0 = fmaf(q.y, fx, q.x), c1 = fmaf(q.w, fx, q.z);
    return lerp_fma(c0, c1, fy);
}

// texture(u_cloud_shape_texture, c).r — cloud_funcs:49 (repeat, trilinear, LOD 0)
template <bool POW2> B200_DEV float sample_shape(const float4* __restrict__ cells, int nx, int ny, int nz, float cx, float cy, float cz) {
    cx = wrap01(cx);  // repeat: wrap to [0,1]
    cy = wrap01(cy);
    cz = wrap01(cz);
    float fx, fy, fz;
    const int xi = floor_frac(texel_coord<POW2>(cx, float(nx)), fx) + 1;   // in [0, n] by construction
    const int yi = floor_frac(texel_coord<POW2>(cy, float(ny)), fy) + 1;
    const int zi = floor_frac(texel_coord<POW2>(cz, float(nz)), fz) + 1;
    const unsigned cxn = unsigned(nx + 1), cyn = unsigned(ny + 1);
    unsigned idx = (unsigned(zi) * cyn + unsigned(yi)) * cxn + unsigned(xi);
    idx = min(idx, cxn * cyn * unsigned(nz + 1) - 1u);            // NaN guard only
    const float4 a = ldg4(cells + 2u * idx), b = ldg4(cells + 2u * idx + 1u);
    const float c00 = fmaf(a.y, fx, a.x), c10 = fmaf(a.w, fx, a.z);
    const float c01 = fmaf(b.y, fx, b.x), c11 = fmaf(b.w, fx, b.z);
    return lerp_fma(lerp_fma(c00, c10, fy), lerp_fma(c01, c11, fy), fz);
}

B200_DEV float4 make_lut_cell(const float* __restrict__ lut_pad, int xi, int yi) {
    const float* p = lut_pad + yi * kLutPad + xi;
    const float t00 = p[0], t10 = p[1], t01 = p[kLutPad], t11 = p[kLutPad + 1];
    const float dx = t10 - t00, dy = t01 - t00, dxy = (t11 - t01) - dx;
    // expansion about the cell centre, g,h in [-0.5, 0.5]: t = tc + dxc*g + h*(dyc + dxy*g)
    return make_float4(t00 + 0.5f * dx + 0.5f * dy + 0.25f * dxy, dx + 0.5f * dxy, dy + 0.5f * dxy, dxy);
}

// Register-resident loop constants: an opaque move stops ptxas from re-materialising them (MOV / LDC)
// inside the march loop, which is issue-bound.
#ifdef __CUDACC__
B200_DEV float pin_reg(float v) {
    asm volatile("mov.f32 %0, %0;" : "+f"(v));
    return v;
}
B200_DEV unsigned pin_reg(unsigned v) {
    asm volatile("mov.b32 %0, %0;" : "+r"(v));
    return v;
}
B200_DEV const float4* pin_reg(const float4* v) {
    asm volatile("mov.b64 %0, %0;" : "+l"(v));
    return v;
}
#else
B200_DEV float pin_reg(float v) { return v; }
B200_DEV unsigned pin_reg(unsigned v) { return v; }
B200_DEV const float4* pin_reg(const float4* v) { return v; }
#endif

// Rare path of scatter_v2: the literal alpha recurrence of funcs_v2:78-79 with a correctly rounded exp.
// render_clouds' blend_colors (util:61-69) is DISCONTINUOUS at total alpha == 0 (it returns vec4(0)), so for rays that
// barely touch the atmosphere with jitter == 0 it matters whether alpha is exactly 0: the shader's per-step
// (1 - exp(-x)) terms all vanish for x <= 2^-25, while 1 - exp(-sum x) may not. Only rays with |view_od| < 2e-5 come
// here (a thin rim of the limb). All ld_step share one sign, so each |x| <= |view_od| < 2e-5 and exp(-x) rounds like
// 1 - x + x^2/2 (next term 1e-15): no fp64, no libm.
B200_DEV_RARE float scatter_alpha_recurrence(const DevConsts& c, f3 o, f3 d, float t_begin, float step_len, int steps) {
    const f3 C = ld3(c.C);
    f3 pos = o + d * t_begin;
    const f3 dstep = d * step_len;
    const float ld_scale = c.rho2 * step_len;
    const float neg_inv_H = -c.inv_H;
    float alpha = 0.0f;
    for (int i = 0; i < steps; ++i) {
        const f3 rel = pos - C;
        const float d2 = fmaf(rel.z, rel.z, fmaf(rel.y, rel.y, rel.x * rel.x));
        float inv;
        const float dist = sqrt_refined(d2, inv);
        const float y = __saturatef(fmaf(dist - c.R, neg_inv_H, 1.0f));
        const float ld_step = (y * y) * (y * ld_scale);
        const float vt = 1.0f + fmaf(ld_step, 0.5f * ld_step, -ld_step);   // exp(-x) for tiny x, one final rounding  :78
        alpha = alpha + (1.0f - vt) * (1.0f - alpha);    // :79
        pos = pos + dstep;
    }
    return alpha;
}

// ------------------------------------------------------------------------------------------------
// include/atmosphere_funcs_v2.gdshaderinc:32-101 — N-step in-scatter march against the baked LUT
//
// LUT fetch (funcs_v2:14-29) from the CELL layout (atmo_kernels.cu: lut_cells_kernel): one float4 per pair of adjacent
// padded texel rows/columns holding the bilinear patch expanded about the cell centre, so texture(LUT, uv) is ONE 16-byte
// load and three FMAs. Padded texel coordinates: xp = u*256 + 0.5 = 128*mu + 128.5, yp = 256*(1-y) + 0.5 with
// mu = dot(up, sun_dir), y = 1 - height_ratio. The cell index n = floor(xp) = rn(xp - 0.5) comes out of ONE add against
// the 1.5*2^23 constant (integer in the low mantissa bits) and the centred fraction g = (xp - 0.5) - n out of two more.
// The LUT is smooth, so (unlike the noise textures) its coordinates need not reproduce the shader's rounding.
// ------------------------------------------------------------------------------------------------
B200_DEV float4 scatter_v2(const DevConsts& c, f3 o, f3 d, float t_begin, float t_end, float jitter) {
    const int steps = c.scatter_steps;
    const f3 C = ld3(c.C);
    const f3 sun128 = mk3(c.sun_dir[0] * 128.0f, c.sun_dir[1] * 128.0f, c.sun_dir[2] * 128.0f);  // exact scaling
    const float step_len = (t_end - t_begin) / float(steps);
    f3 pos = o + d * t_begin;   // pos0, :57-58 (exact)
    const f3 dstep = d * step_len;
    // local_density*step_len = y^3 * (rho^2 * step_len): the constant factor is applied once after the loop, the loop
    // accumulates S = sum y^3 (view optical depth / ld_scale) and sum y^3 * T_c
    const float ld_scale = c.rho2 * step_len;
    const float k0 = c.neg_coef_log2e[0], k1 = c.neg_coef_log2e[1], k2 = c.neg_coef_log2e[2];
    const float neg_inv_H = pin_reg(-c.inv_H);
    const float n256 = pin_reg(-256.0f);
    constexpr unsigned off_bias = 0u - unsigned(kMagicBits) * unsigned(kLutCells + 1);  // both magic biases, mod 2^32
    const float4* cells = pin_reg(c.lut_cells);
    float L0 = 0.0f, L1 = 0.0f, L2 = 0.0f, S = 0.0f;

B200_UNROLL(B200ATMO_SCATTER_UNROLL)
    for (int i = 0; i < steps; ++i) {
        const f3 rel = pos - C;                                             // exact: carries the shader's position rounding
        const float d2 = fmaf(rel.z, rel.z, fmaf(rel.y, rel.y, rel.x * rel.x));
        const float sd = fmaf(rel.z, sun128.z, fmaf(rel.y, sun128.y, rel.x * sun128.x));
        float inv;
        const float dist = sqrt_refined(d2, inv);                           // distance(pos, planet_center)
        // y = 1 - clamp((dist-R)/H, 0, 1) in one rounding (atmosphere_common:13-15); height_ratio (:17-18) = 1 - y
        const float y = __saturatef(fmaf(dist - c.R, neg_inv_H, 1.0f));
        // 128*mu = sd*inv = 128 * dot(normalize(pos-C), sun_dir) (:19-20) is never materialised: one fma gives the cell
        // column n_x = rn(128 mu + 128) in the low mantissa bits, a second one the centred fraction g = (128 mu + 128) - n_x
        const float xm = fmaf(sd, inv, kMagic + 128.0f);
        const float ym = fmaf(y, n256, kMagic + 256.0f);                    // cell row n_y = rn(256 - 256 y)
        const float g = fmaf(sd, inv, (kMagic + 128.0f) - xm);              // centred fractions in [-0.5, 0.5]
        const float h = fmaf(y, n256, (kMagic + 256.0f) - ym);              // (256 - 256 y) - n_y
        // row*257 + col; the unsigned min turns the garbage of a NaN coordinate (pos == planet centre) into an
        // in-range read instead of a fault
        unsigned off = unsigned(__float_as_int(ym)) * unsigned(kLutCells) + unsigned(__float_as_int(xm));
        off = min(off + off_bias, unsigned(kLutCells * kLutCells - 1));
        const float4 q = ldg4(cells + off);                                 // :28
        const float sun_od = fmaf(fmaf(q.w, g, q.z), h, fmaf(q.y, g, q.x));
        const float y3 = (y * y) * y;                                       // local_density*step_len / ld_scale, :64-65
        S += y3;                                                            // :66
        const float od = fmaf(S, ld_scale, sun_od);                         // sun_ray + view_ray optical depth
        const float T0 = ex2_approx(od * k0), T1 = ex2_approx(od * k1), T2 = ex2_approx(od * k2);  // :71-73
        L0 = fmaf(y3, T0, L0);                                              // :75 (ld_scale and the coefficient are applied after the loop)
        L1 = fmaf(y3, T1, L1);
        L2 = fmaf(y3, T2, L2);
        pos = pos + dstep;                                                  // :81 (exact)
    }
    const float view_od = S * ld_scale;
    // alpha: the recurrence a += (1-vt)(1-a), vt = exp(-ld*step) telescopes to 1 - exp(-view_od)  (:78-79)
    float alpha = 1.0f - ex2_approx(view_od * -1.4426950408889634f);
    if (fabsf(view_od) < 2e-5f) alpha = scatter_alpha_recurrence(c, o, d, t_begin, step_len, steps);  // exact-zero cases
    const float r = clampf(fmaf(L0 * ld_scale, c.coef[0], c.ambient[0]), 0.0f, 1.0f) * c.modulate[0];  // :91, :98
    const float g = clampf(fmaf(L1 * ld_scale, c.coef[1], c.ambient[1]), 0.0f, 1.0f) * c.modulate[1];
    const float b = clampf(fmaf(L2 * ld_scale, c.coef[2], c.ambient[2]), 0.0f, 1.0f) * c.modulate[2];
    alpha = clampf(fmaf(jitter, 0.02f, alpha), 0.0f, 0.99f);                                 // :96
    return make_float4(r, g, b, alpha);
}

// ------------------------------------------------------------------------------------------------
// include/atmosphere_funcs_v1.gdshaderinc:15-63 — "lite" model
// ------------------------------------------------------------------------------------------------
B200_DEV float4 scatter_v1(const DevConsts& c, f3 o, f3 d, float t_begin, float t_end) {
    const int steps = c.scatter_steps;
    const f3 C = ld3(c.C), sun = ld3(c.sun_dir);
    const float inv_steps = 1.0f / float(steps);
    const float step_len = (t_end - t_begin) * inv_steps;
    const f3 stepv = step_len * d;
    f3 pos = o + d * t_begin;
    float factor = 1.0f, light_sum = 0.0f;
B200_UNROLL(2)
    for (int i = 0; i < steps; ++i) {
        const f3 rel = pos - C;
        const float d2 = dot3(rel, rel);
        float inv;
        const float dist = sqrt_refined(d2, inv);
        const float h = __saturatef(div_refined(dist - c.R, c.H, c.inv_H));
        const float y = 1.0f - h;
        const float density = y * y * y * c.rho;                       // density applied once in v1 (funcs_v1:33)
        const float sdot = fmaf(rel.z, sun.z, fmaf(rel.y, sun.y, rel.x * sun.x)) * inv;   // dot(sun_dir, up) :31,35
        float light = __saturatef(fmaf(1.2f, sdot, 0.5f));
        light = light * light;                                          // :36
        light_sum = fmaf(light, inv_steps, light_sum);                  // :38
        factor *= fmaf(-density, step_len, 1.0f);                       // :39
        pos = pos + stepv;
    }
    const float atmo_factor = 1.0f - factor;
    const float day_factor = __saturatef(light_sum * c.day_night_scale);
    float rgb[3];
    for (int k = 0; k < 3; ++k) {
        const float night = mixf(c.night0[k], c.night1[k], atmo_factor);
        const float day = mixf(c.day0[k], c.day1[k], atmo_factor);
        rgb[k] = mixf(night, day, day_factor);
    }
    return make_float4(rgb[0], rgb[1], rgb[2], __saturatef(atmo_factor));
}

// ------------------------------------------------------------------------------------------------
// include/cloud_funcs.gdshaderinc
// ------------------------------------------------------------------------------------------------
#ifdef B200ATMO_STEP_STATS
static long long b200atmo_step_stats_skipped = 0;
#endif
// height_ratio (:36-37, :95-96, :111-112): (|p| - bottom) / (top - bottom)
B200_DEV float cloud_height_ratio(const DevConsts& c, float len) {
    return div_refined(len - c.hot.bottom_h, c.hot.thickness, c.hot.inv_thickness);
}

// get_density_full (:31-68) with |p| and height_ratio already computed; clamped density in [0,1].
// Exact skip: outside the shell height_curve clamps to 0, so (..)*0*50-20 clamps to exactly 0.
template <bool POW2> B200_DEV float cloud_density(const DevConsts& c, f3 p, float hr) {
    const float a = twice_minus1(hr);                                          // height_curve :25-29, exact
    const float hc = 1.0f - a * a;
    // outside the shell (hc <= 0) and in the rim of the shell where not even the largest coverage texel with the largest
    // shape value reaches a positive density (hc <= c.hc_min, an exact bound: atmo_consts.h cloud_hc_min)
#ifdef B200ATMO_NO_HCMIN   // tuning knob: the plain shell test
    if (!(hc > 0.0f)) return 0.0f;
#else
    if (!(hc > c.hot.hc_min)) return 0.0f;
#endif
    const float cpx = c.hot.rot[0] * p.x + c.hot.rot[2] * p.z;                 // u_cloud_coverage_rotation * p.xz :43, exact
    const float cpz = c.hot.rot[1] * p.x + c.hot.rot[3] * p.z;
    float coverage = sample_cube<POW2>(c.hot.cube_cells, c.hot.cube_res, cpx, p.y, cpz);   // :45
    coverage = minus_quarter(coverage, hr) + c.hot.coverage_bias;              // :46
    const float cov_term = mixf(-1.2f, 1.5f, coverage);
    // Exact early-out before the 3D fetch: the expression below is monotone in `shape` (every op is monotone under
    // round-to-nearest, hc > 0), so if it is <= 0 for the largest possible shape value it is <= 0 for the real one
    // and the clamped density is exactly 0. ~3/4 of the in-shell samples of the demo scene end here.
    // (c.dens_y_min = the largest y with y*50 - 20 <= 0 in fp32, found on the host: the same decision as evaluating
    // (..)*hc*50 - 20 > 0 with two instructions less)
#ifdef B200ATMO_LITERAL
    if (!((c.hot.shape_hi_m01 + cov_term) * hc * 50.0f - 20.0f > 0.0f)) return 0.0f;
#else
    if (!((c.hot.shape_hi_m01 + cov_term) * hc > c.hot.dens_y_min)) return 0.0f;
#endif
    const float tex = sample_shape<POW2>(c.hot.shape_cells, c.hot.nx, c.hot.ny, c.hot.nz, p.x * c.hot.shape_scale,
                                   p.y * c.hot.shape_scale, p.z * c.hot.shape_scale);
#ifdef B200ATMO_LITERAL
    float shape = mixf(0.5f, tex, c.hot.shape_factor);                         // :48-50
#else
    float shape = c.hot.shape_mix0 + tex * c.hot.shape_factor;                 // mix(0.5, tex, factor) :48-50; 0.5*(1-factor) from the host
#endif
    if (c.hot.shape_invert) shape = 1.0f - shape;                              // :57-59
    // detail = 0.5 (CLOUDS_ALWAYS_LOW_QUALITY, main:49) => 0.2*detail = 0.1 (same fp32 product)
    float density = (shape - 0.2f * 0.5f + cov_term) * hc;                     // :61
    density = density * 50.0f - 20.0f;                                         // :62
    return __saturatef(density);                                               // :64
}

// get_light_raymarched (:104-151). `dens0` = the clamped density at pos0, which the caller has just evaluated (> 0: the light
// is only needed then). The shader's first light sample (i = 0) sits at pos0 + 0*step*sun_dir = pos0 — exactly, 0*x is +-0 and
// p + (+-0) == p — and get_density is the function the march itself just called there (:131-136: the quality switch is dead,
// CLOUDS_ALWAYS_LOW_QUALITY), so its value is dens0 bit for bit: six of the shader's seven density evaluations per lit cloud step
// remain. B200ATMO_LIGHT_RESAMPLE_FIRST restores the literal seventh (tuning / audit knob).
template <bool POW2> B200_DEV float light_raymarched(const DevConsts& c, f3 pos0, f3 sun, float hr0, float dens0) {
    float step_len = c.hot.light_reach * (1.0f / 6.0f);   // reach * inv_steps
#ifdef B200ATMO_LIGHT_RESAMPLE_FIRST
    float transm = 1.0f;                               // 1 - alpha
    constexpr int kFirst = 0;
#else
    // i = 0: a = 0 + (1 - tr)(1 - 0)  =>  1 - a = tr                                                       :138-142
    float transm = ex2_approx(dens0 * (step_len * c.hot.density_scale) * -1.4426950408889634f);
    step_len *= 1.2f;                                                                                       // :143
    constexpr int kFirst = 1;
#endif
B200_UNROLL(B200ATMO_LIGHT_UNROLL)
    for (int i = kFirst; i < 6; ++i) {
        const float t = float(i) * step_len;
        const f3 p = mk3(pos0.x + t * sun.x, pos0.y + t * sun.y, pos0.z + t * sun.z);  // :129, exact
        float inv;
        const float len = sqrt_refined(dot3(p, p), inv);
        const float dens = cloud_density<POW2>(c, p, cloud_height_ratio(c, len));
        if (dens > 0.0f) transm *= ex2_approx(dens * (step_len * c.hot.density_scale) * -1.4426950408889634f);  // :138-142
        step_len *= 1.2f;                                                                                   // :143
    }
    const float alpha = 1.0f - transm;
    return mixf(1.0f, hr0 * 0.2f, alpha);              // :146-150
}

// raymarch_cloud (:175-247) in model space; returns (total_light, alpha)
// LIGHT = light mode (bits 0-1) | kLightPow2 (texture sizes are powers of two)
template <int LIGHT> B200_DEV f2 raymarch_cloud(const DevConsts& c, f3 o, f3 d, float t_begin, float t_end, float jitter, f3 sun) {
    constexpr int MODE = LIGHT & 3;
    constexpr bool POW2 = (LIGHT & kLightPow2) != 0;
    constexpr int kUnroll = MODE == B200ATMO_LIGHT_RAYMARCHED ? B200ATMO_CLOUD_UNROLL_RM : B200ATMO_CLOUD_UNROLL_CHEAP;
    (void)kUnroll;   // only the CUDA build has the pragma
    const int steps = c.cloud_steps;
    // march-length cap (:186-204), exact
    const float max_d = mixf(c.march_ground, c.march_space, smoothstepf(c.march_hmin, c.march_hmax, sqrtf(dot3(o, o))));
    t_end = t_begin + fminf(t_end - t_begin, max_d);
    const float inv_steps = 1.0f / float(steps);
    const float step_len = (t_end - t_begin) * inv_steps;
    f3 pos = o + jitter * step_len * d + d * t_begin;  // :213, exact
    const f3 dstep = d * step_len;

    // loop invariant: max(pow(dot(ray_dir, sun_dir), 16), 0) (:98-101; a non-positive base gives 0)
    float sunpeek = 0.0f;
    if (MODE == B200ATMO_LIGHT_CHEAP) {
        const float dp = dot3(d, sun);
        if (dp > 0.0f) {
            const float p2 = dp * dp, p4 = p2 * p2, p8 = p4 * p4;
            sunpeek = p8 * p8;
        }
    }
    const float k = step_len * c.hot.density_scale;
    float T_clamped = 1.0f;  // total_transmittance (:222-223)
    float T_alpha = 1.0f;    // 1 - alpha (:228 telescopes to a product of transmittances)
    float total_light = 0.0f;
    // Exact skip of the steps that cannot meet cloud: the run that passes UNDER the shell (a ray that reaches the ground, or
    // crosses the shell twice, spends a third of its steps there) and the steps in the empty rim at the TOP of the shell
    // where the march starts and ends. The positions are pos + i*dstep up to the rounding of the additions, so
    // |pos_i|^2 < under_r2 and |pos_i|^2 > over_r2 are quadratic inequalities in i, solved once per ray. The two radii
    // (atmo_consts.h: cloud_skip_r2) sit far enough outside the band that can hold cloud to absorb that rounding and the fp32
    // error of this solve (the guard on |pos|^2 keeps the latter below 5e-6 * r^2), and one whole step is given away at
    // every end on top of that: in a skipped step the shader's density is exactly 0, so nothing but `pos += dstep` happens.
    // The march becomes  skip | evaluate [b0, e0) | skip | evaluate [b1, e1) | (the rest only moves pos: dropped).  On the GPU
    // the four bounds are widened to the union over the lanes that march together, so the trip counts are warp-uniform and
    // the evaluated loop carries no per-step test (8x4 pixel tiles: 1/3 of the warp-steps of the cfg4 frame have no lane
    // in the shell).
    int b0 = 0, e0 = steps, b1 = steps, e1 = steps;
#if !defined(B200ATMO_LITERAL) && !defined(B200ATMO_NO_UNDER_SKIP)
    {
        const float qa = dot3(dstep, dstep), qb = dot3(pos, dstep), q0 = dot3(pos, pos);
        if (qa > 1e-30f && q0 < 4.0f * c.hot.under_r2) {
            const float ia = 1.0f / qa, nsteps = float(steps);
            int in_b = 0, in_e = steps;          // steps that may be below over_r: [in_b, in_e)
            const float disc_o = qb * qb - qa * (q0 - c.hot.over_r2);
            if (!(disc_o > 0.0f)) in_e = 0;      // the whole line stays above the shell (NaN: nothing is finite, nothing to march)
            else {
                const float sq = sqrtf(disc_o);
                const float lo = (-qb - sq) * ia, hi = (-qb + sq) * ia;   // |pos + x*dstep|^2 < over_r2 for x in (lo, hi)
                if (fabsf(lo) < 1e9f && fabsf(hi) < 1e9f) {
                    in_b = int(fminf(fmaxf(floorf(lo) - 1.0f, 0.0f), nsteps));
                    in_e = int(fminf(fmaxf(ceilf(hi) + 2.0f, 0.0f), nsteps));
                }
            }
            int un_b = steps, un_e = steps;      // steps surely below under_r: [un_b, un_e)
            const float disc_u = qb * qb - qa * (q0 - c.hot.under_r2);
            if (disc_u > 0.0f) {
                const float sq = sqrtf(disc_u);
                const float lo = (-qb - sq) * ia, hi = (-qb + sq) * ia;   // |pos + x*dstep|^2 < under_r2 for x in (lo, hi)
                if (fabsf(lo) < 1e9f && fabsf(hi) < 1e9f) {
                    const float first = fmaxf(ceilf(lo) + 1.0f, 0.0f), end = fminf(floorf(hi), nsteps);
                    if (end > first) { un_b = int(first); un_e = int(end); }
                }
            }
            if (!(disc_o == disc_o) || !(disc_u == disc_u)) { in_b = 0; in_e = steps; un_b = un_e = steps; }   // NaN anywhere: march literally
            b0 = in_b; e0 = un_b < in_e ? un_b : in_e;
            b1 = un_e > in_b ? un_e : in_b; e1 = in_e;
            if (e0 < b0) e0 = b0;
            if (b1 < e0) b1 = e0;
            if (e1 < b1) e1 = b1;
#ifdef B200ATMO_STEP_STATS   // host builds only: how many steps the skip covers (so a comparison can show it was exercised)
            b200atmo_step_stats_skipped += steps - (e0 - b0) - (e1 - b1);
#endif
        }
    }
#ifdef __CUDA_ARCH__
    {   // widen to the lanes that are here together (any converged subset is a valid group)
        const unsigned grp = __activemask();
        const bool none0 = e0 <= b0, none1 = e1 <= b1;       // an empty segment must not widen the union
        b0 = __reduce_min_sync(grp, none0 ? 0x7fffffff : b0); e0 = __reduce_max_sync(grp, none0 ? 0 : e0);
        b1 = __reduce_min_sync(grp, none1 ? 0x7fffffff : b1); e1 = __reduce_max_sync(grp, none1 ? 0 : e1);
        if (e0 <= b0) b0 = e0 = 0;                           // no lane has a first segment
        if (e1 <= b1) b1 = e1 = e0;                          // ... a second one
        if (b1 < e0) { e0 = e1 > e0 ? e1 : e0; b1 = e1 = e0; }   // the lanes' first segments reach into a second one: one segment
    }
#endif
#endif
    int i = 0;
B200_UNROLL(1)
    for (int seg = 0; seg < 2; ++seg) {
    const int seg_b = seg == 0 ? b0 : b1, seg_e = seg == 0 ? e0 : e1;
    for (; i < seg_b; ++i) pos = pos + dstep;   // :236 — outside the cloud band: density 0, transmittance 1, nothing accumulates
B200_UNROLL(kUnroll)
    for (; i < seg_e; ++i) {
        float inv;
        const float len = sqrt_refined(dot3(pos, pos), inv);
        const float hr = cloud_height_ratio(c, len);
        const float dens01 = cloud_density<POW2>(c, pos, hr);
        if (dens01 > 0.0f) {  // density == 0 => transmittance 1, no light added, alpha unchanged: exact skip
            float light;      // get_light (:153-167)
            if (MODE == B200ATMO_LIGHT_RAYMARCHED) light = light_raymarched<POW2>(c, pos, sun, hr, dens01);
            else light = fmaf(sunpeek, T_alpha, hr);                                         // :95-101
            const float sdot = -(fmaf(pos.z, sun.z, fmaf(pos.y, sun.y, pos.x * sun.x)) * inv);  // dot(normalize(pos), -sun)
            const float st = __saturatef((sdot + 0.3f) * (1.0f / 0.6f));                     // smoothstep(-0.3, 0.3, .) :87
            const float shadow = st * st * fmaf(-2.0f, st, 3.0f);
            light *= fmaf(shadow, -0.998f, 1.0f);                                             // mix(1, 0.002, shadow) :164
            const float dens_step = dens01 * k;                                               // density*scale*step_len
            const float tr = ex2_approx(dens_step * -1.4426950408889634f);                    // :221
            T_clamped = fmaxf(T_clamped * tr, 0.005f);                                        // :222-223
            total_light = fmaf(light * dens_step, T_clamped, total_light);                    // :226
            T_alpha *= tr;                                                                    // :228
        }
        pos = pos + dstep;  // :236, exact
    }
    }
    return f2{total_light, 1.0f - T_alpha};
}

// render_clouds (:249-324) in three parts: set-up + visibility test (exact), the march, the blend.
struct CloudRay {
    bool active;      // the visibility test of :263-278 passed
    f3 o, d;          // model-space ray (:286-287), not renormalised
    float t0, t1;     // cloud_rs (:269-271)
};
B200_DEV CloudRay clouds_setup(const DevConsts& c, f3 o, f3 d, float linear_depth) {
    CloudRay r;
    r.active = false;
    r.o = r.d = mk3(0.f, 0.f, 0.f);
    r.t0 = r.t1 = 0.0f;
    const f3 C = ld3(c.C);
    const f2 rs_top = ray_sphere(C, c.cloud_top_h, o, d);
    if (rs_top.x == rs_top.y) return r;
    const f2 rs_bottom = ray_sphere(C, c.cloud_bottom_h, o, d);
    r.t0 = fmaxf(rs_top.x, 0.0f);
    r.t1 = fminf(rs_top.y, linear_depth);
    if (!(r.t0 < linear_depth && (linear_depth > rs_bottom.y || rs_bottom.x > 0.0f))) return r;  // :273-278
    float om[4], dm[4];
    mat4_mul(c.v2m, o.x, o.y, o.z, 1.0f, om);  // :286-287
    mat4_mul(c.v2m, d.x, d.y, d.z, 0.0f, dm);
    r.o = mk3(om[0], om[1], om[2]);
    r.d = mk3(dm[0], dm[1], dm[2]);
    r.active = true;
    return r;
}
B200_DEV void clouds_blend(const DevConsts& c, float4& px, f2 rr) {
    const float cl = rr.x, ca = rr.y;
    // blend_colors(self = atmosphere, over = cloud), util:61-69
    const float sa = 1.0f - ca;
    const float a = px.w * sa + ca;
    float br = 0.f, bg = 0.f, bb = 0.f, ba = 0.f;
    if (a != 0.0f) {
        br = (px.x * px.w * sa + cl * ca) / a;
        bg = (px.y * px.w * sa + cl * ca) / a;
        bb = (px.z * px.w * sa + cl * ca) / a;
        ba = a;
    }
    const float ar = px.x + cl * ca, ag = px.y + cl * ca, ab = px.z + cl * ca, aa = fmaxf(px.w, ca);  // :308-310
    px.x = mixf(br, ar, c.cloud_blend);  // :318
    px.y = mixf(bg, ag, c.cloud_blend);
    px.z = mixf(bb, ab, c.cloud_blend);
    px.w = mixf(ba, aa, c.cloud_blend);
}
template <int LIGHT> B200_DEV void render_clouds(const DevConsts& c, float4& px, f3 o, f3 d, float linear_depth, float jitter) {
    const CloudRay r = clouds_setup(c, o, d, linear_depth);
    if (!r.active) return;
    clouds_blend(c, px, raymarch_cloud<LIGHT>(c, r.o, r.d, r.t0, r.t1, jitter, ld3(c.hot.sun)));
}

#ifdef __CUDACC__
// ------------------------------------------------------------------------------------------------
// Warp-wide compaction of the raymarched cloud light (B200ATMO_LIGHT_QUEUE). The 6-step light march (:104-151) costs six
// density evaluations but only runs for the samples whose density is > 0; in raymarch_cloud<RAYMARCHED> the lanes of a
// warp without such a sample idle meanwhile. Here every lane marches its own ray, but a sample that needs the light
// march is PUSHED into a per-warp shared-memory FIFO (position, height ratio, and the three factors the light value is
// multiplied with); whenever 32 items are queued all 32 lanes pop one each and run the light march at full occupancy,
// then every lane folds the results of ITS items into its accumulator in queue (= step) order. Per ray the arithmetic
// and its order are exactly those of raymarch_cloud<RAYMARCHED>: results are bit-identical.
//   q : 64 ring slots x 2 float4 per warp: {pos.xyz, height_ratio}, {shadow factor, density*step, T_clamped, density (in) / light (out)}
// Whether it pays depends on how coherent the lanes of a warp are (profiles/r02/warp_model.txt): measured in
// profiles/r02/tune_clouds.txt.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long rotr64(unsigned long long v, unsigned s) { return s ? (v >> s) | (v << (64u - s)) : v; }
__device__ __forceinline__ unsigned long long rotl64(unsigned long long v, unsigned s) { return s ? (v << s) | (v >> (64u - s)) : v; }

template <bool POW2> __device__ __noinline__ f2 raymarch_cloud_light_queue(const DevConsts& c, const CloudRay r, float jitter, f3 sun, float4* q) {
    constexpr unsigned kFull = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31u;
    if (!__any_sync(kFull, r.active)) return f2{0.0f, 0.0f};
    const int steps = c.cloud_steps;
    const f3 o = r.o, d = r.d;
    const float t_begin = r.t0;
    // march-length cap (:186-204), exact
    const float max_d = mixf(c.march_ground, c.march_space, smoothstepf(c.march_hmin, c.march_hmax, sqrtf(dot3(o, o))));
    const float t_end = t_begin + fminf(r.t1 - t_begin, max_d);
    const float inv_steps = 1.0f / float(steps);
    const float step_len = (t_end - t_begin) * inv_steps;
    f3 pos = o + jitter * step_len * d + d * t_begin;  // :213, exact
    const f3 dstep = d * step_len;
    const float k = step_len * c.hot.density_scale;
    float T_clamped = 1.0f, T_alpha = 1.0f, total_light = 0.0f;
    unsigned head = 0, count = 0;       // warp-uniform ring state
    unsigned long long mine = 0;        // ring slots that hold items of this lane

    auto drain = [&](unsigned n) {      // pop n <= 32 items, light-march them on n lanes, fold the results back in FIFO order
        __syncwarp();
        if (lane < n) {
            const unsigned slot = (head + lane) & 63u;
            const float4 a = q[2 * slot];
            const float L = light_raymarched<POW2>(c, mk3(a.x, a.y, a.z), sun, a.w, reinterpret_cast<const float*>(q + 2 * slot + 1)[3]);
            reinterpret_cast<float*>(q + 2 * slot + 1)[3] = L;
        }
        __syncwarp();
        const unsigned long long window = n >= 64u ? ~0ull : ((1ull << n) - 1ull);
        unsigned long long m = rotr64(mine, head) & window;
        while (m) {
            const unsigned kbit = unsigned(__ffsll((long long)m)) - 1u;
            m &= m - 1ull;
            const float4 e = q[2 * ((head + kbit) & 63u) + 1];
            const float light = e.w * e.x;                                    // get_light * mix(1, 0.002, shadow) :164
            total_light = fmaf(light * e.y, e.z, total_light);                // :226
        }
        mine &= ~rotl64(window, head);
        head = (head + n) & 63u;
        count -= n;
        __syncwarp();
    };

B200_UNROLL(1)
    for (int i = 0; i < steps; ++i) {
        bool hit = false;
        float hr = 0.0f, sf = 0.0f, ds = 0.0f, d0 = 0.0f;
        if (r.active) {
            float inv;
            const float len = sqrt_refined(dot3(pos, pos), inv);
            hr = cloud_height_ratio(c, len);
            const float dens01 = cloud_density<POW2>(c, pos, hr);
            if (dens01 > 0.0f) {
                hit = true;
                d0 = dens01;
                const float sdot = -(fmaf(pos.z, sun.z, fmaf(pos.y, sun.y, pos.x * sun.x)) * inv);
                const float st = __saturatef((sdot + 0.3f) * (1.0f / 0.6f));
                const float shadow = st * st * fmaf(-2.0f, st, 3.0f);
                sf = fmaf(shadow, -0.998f, 1.0f);
                ds = dens01 * k;
                const float tr = ex2_approx(ds * -1.4426950408889634f);       // :221
                T_clamped = fmaxf(T_clamped * tr, 0.005f);                    // :222-223
                T_alpha *= tr;                                                // :228
            }
        }
        const unsigned b = __ballot_sync(kFull, hit);
        if (b) {
            if (hit) {
                const unsigned slot = (head + count + unsigned(__popc(b & ((1u << lane) - 1u)))) & 63u;
                q[2 * slot] = make_float4(pos.x, pos.y, pos.z, hr);
                q[2 * slot + 1] = make_float4(sf, ds, T_clamped, d0);   // .w: density in, light out
                mine |= 1ull << slot;
            }
            count += unsigned(__popc(b));
            if (count >= 32u) drain(32u);
        }
        pos = pos + dstep;  // :236, exact
    }
    if (count) drain(count);
    return f2{total_light, 1.0f - T_alpha};
}
#endif

// ------------------------------------------------------------------------------------------------
// NoiseCubemap content: "b200 gradient fBm v1" (see include/b200atmo.h). Exact arithmetic throughout (no fmaf), so the
// bytes are identical on every IEEE machine.
// ------------------------------------------------------------------------------------------------
B200_DEV unsigned noise_hash(int ix, int iy, int iz, unsigned seed) {
    unsigned h = seed;
    h ^= unsigned(ix) * 0x8da6b343u;
    h ^= unsigned(iy) * 0xd8163841u;
    h ^= unsigned(iz) * 0xcb1ab31fu;
    h *= 0x9e3779b1u;
    h ^= h >> 15;
    h *= 0x85ebca6bu;
    h ^= h >> 13;
    return h;
}
B200_DEV float noise_grad(unsigned hash, float x, float y, float z) {
    const unsigned h = hash & 15u;
    const float u = h < 8u ? x : y;
    const float v = h < 4u ? y : ((h == 12u || h == 14u) ? x : z);
    return ((h & 1u) ? -u : u) + ((h & 2u) ? -v : v);
}
B200_DEV float noise_fade(float t) { return t * t * t * (t * (t * 6.0f - 15.0f) + 10.0f); }
B200_DEV float noise_lerp(float a, float b, float t) { return a + (b - a) * t; }
B200_DEV float noise3(float x, float y, float z, unsigned seed) {
    const float fx0 = floorf(x), fy0 = floorf(y), fz0 = floorf(z);
    const int ix = int(fx0), iy = int(fy0), iz = int(fz0);
    const float fx = x - fx0, fy = y - fy0, fz = z - fz0;
    const float u = noise_fade(fx), v = noise_fade(fy), w = noise_fade(fz);
    const float n000 = noise_grad(noise_hash(ix, iy, iz, seed), fx, fy, fz);
    const float n100 = noise_grad(noise_hash(ix + 1, iy, iz, seed), fx - 1.0f, fy, fz);
    const float n010 = noise_grad(noise_hash(ix, iy + 1, iz, seed), fx, fy - 1.0f, fz);
    const float n110 = noise_grad(noise_hash(ix + 1, iy + 1, iz, seed), fx - 1.0f, fy - 1.0f, fz);
    const float n001 = noise_grad(noise_hash(ix, iy, iz + 1, seed), fx, fy, fz - 1.0f);
    const float n101 = noise_grad(noise_hash(ix + 1, iy, iz + 1, seed), fx - 1.0f, fy, fz - 1.0f);
    const float n011 = noise_grad(noise_hash(ix, iy + 1, iz + 1, seed), fx, fy - 1.0f, fz - 1.0f);
    const float n111 = noise_grad(noise_hash(ix + 1, iy + 1, iz + 1, seed), fx - 1.0f, fy - 1.0f, fz - 1.0f);
    const float a = noise_lerp(noise_lerp(n000, n100, u), noise_lerp(n010, n110, u), v);
    const float b = noise_lerp(noise_lerp(n001, n101, u), noise_lerp(n011, n111, u), v);
    return noise_lerp(a, b, w);
}
// fractal sum normalised to about [-1, 1]
B200_DEV float noise_fbm(float x, float y, float z, const B200AtmoNoise& n) {
    float freq = n.frequency, amp = 1.0f, sum = 0.0f, norm = 0.0f;
    for (int o = 0; o < n.octaves; ++o) {
        sum = sum + amp * noise3(x * freq, y * freq, z * freq, unsigned(n.seed) + unsigned(o) * 0x632be5abu);
        norm = norm + amp;
        freq = freq * n.lacunarity;
        amp = amp * n.gain;
    }
    return sum / norm;
}
// One texel of NoiseCubemap._generate_images (noise_cubemap.gd:108-134)
B200_DEV unsigned char noise_cube_texel(int side, int x, int y, int res, const float scale[3], const B200AtmoNoise& n) {
    const float half = 0.5f * float(res);                                         // half_resolution_2d, :102
    const float px = (float(x) + 0.5f) / half - 1.0f;                             // :110-111
    const float py = (float(res - y - 1) + 0.5f) / half - 1.0f;
    f3 pos = mk3(1.0f, py, -px);                                                  // :113, then .normalized()
    const float l = sqrtf(dot3(pos, pos));
    pos = mk3(pos.x / l, pos.y / l, pos.z / l);
    f3 q;
    switch (side) {                                                               // :116-128
        case 0: q = mk3(pos.x, pos.y, pos.z); break;
        case 1: q = mk3(-pos.x, pos.y, -pos.z); break;
        case 2: q = mk3(-pos.z, pos.x, -pos.y); break;
        case 3: q = mk3(-pos.z, -pos.x, pos.y); break;
        case 4: q = mk3(-pos.z, pos.y, pos.x); break;
        default: q = mk3(pos.z, pos.y, -pos.x); break;
    }
    const float density = 0.5f + 0.5f * noise_fbm(q.x * scale[0], q.y * scale[1], q.z * scale[2], n);   // :130
    // Image.set_pixel on FORMAT_L8: uint8(CLAMP(v * 255.0, 0, 255)) (engine behaviour)
    const float v = fminf(fmaxf(density * 255.0f, 0.0f), 255.0f);
    return (unsigned char)(int(v));
}

// ------------------------------------------------------------------------------------------------
// include/planet_atmosphere_main.gdshaderinc:144-196 — one fragment, ray already generated
// ------------------------------------------------------------------------------------------------
template <int MODEL, int LIGHT> B200_DEV bool shade_ray(const DevConsts& c, f3 o, f3 d, float linear_depth, float jitter, float4& out) {
    const f3 C = ld3(c.C);
    const f2 rs_atmo = ray_sphere(C, c.atmo_radius, o, d);
    if (rs_atmo.x == rs_atmo.y) {  // miss (or tangent): discard, :150,191-196
        out = make_float4(0.f, 0.f, 0.f, 0.f);
        return true;
    }
    const float t_begin = fmaxf(rs_atmo.x, 0.0f);
    float t_end = fmaxf(rs_atmo.y, 0.0f);
    const f2 rs_ground = ray_sphere(C, c.R, o, d);
    float gd = 10000000.0f;
    if (rs_ground.x != rs_ground.y) gd = rs_ground.x;
    linear_depth = mixf(linear_depth, gd, c.sphere_depth_factor);  // :160
    t_end = fminf(t_end, linear_depth);                            // :162
    if (MODEL == B200ATMO_SCATTER_V1) out = scatter_v1(c, o, d, t_begin, t_end);
    else out = scatter_v2(c, o, d, t_begin, t_end, jitter);
    if ((LIGHT & 3) != B200ATMO_LIGHT_NONE) render_clouds<LIGHT>(c, out, o, d, linear_depth, jitter);
    return false;
}

#ifdef __CUDACC__
// shade_ray for the raymarched-light variant with the per-warp light queue: ALL 32 lanes of the warp must call it
// (`valid` = this lane has a ray); the scatter march and the cloud set-up run per lane, the cloud march cooperatively.
template <int MODEL, bool POW2> __device__ __forceinline__ bool shade_ray_light_queue(const DevConsts& c, bool valid, f3 o, f3 d, float linear_depth,
                                                                         float jitter, float4& out, float4* q) {
    bool disc = true;
    CloudRay cr;
    cr.active = false;
    cr.o = cr.d = mk3(0.f, 0.f, 0.f);
    cr.t0 = cr.t1 = 0.0f;
    out = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid) {
        const f3 C = ld3(c.C);
        const f2 rs_atmo = ray_sphere(C, c.atmo_radius, o, d);
        if (rs_atmo.x != rs_atmo.y) {
            disc = false;
            const float t_begin = fmaxf(rs_atmo.x, 0.0f);
            float t_end = fmaxf(rs_atmo.y, 0.0f);
            const f2 rs_ground = ray_sphere(C, c.R, o, d);
            float gd = 10000000.0f;
            if (rs_ground.x != rs_ground.y) gd = rs_ground.x;
            linear_depth = mixf(linear_depth, gd, c.sphere_depth_factor);  // :160
            t_end = fminf(t_end, linear_depth);                            // :162
            if (MODEL == B200ATMO_SCATTER_V1) out = scatter_v1(c, o, d, t_begin, t_end);
            else out = scatter_v2(c, o, d, t_begin, t_end, jitter);
            cr = clouds_setup(c, o, d, linear_depth);
        }
    }
    __syncwarp();
    const f2 rr = raymarch_cloud_light_queue<POW2>(c, cr, jitter, ld3(c.hot.sun), q);
    if (cr.active) clouds_blend(c, out, rr);
    return disc;
}
#endif

// MODE_FAR coverage (planet_atmosphere.gd:261-282,302-321): the node draws a BoxMesh of edge `atmo_clip_distance`
// centred on itself, so the fragment shader only runs where that cube's front faces are rasterised and pass the
// depth test. Restated with the reference's own (unused) slab test, include/util.gdshaderinc:5-17, in model space:
// a point o + t*d keeps its parameter t under the affine view->model map, so tN is comparable with linear_depth.
// min/max have IEEE fmin/fmax semantics (NaN operands ignored: axis-parallel rays produce 0*inf), as in the oracle.
B200_DEV float sel_max(float a, float b) { return fmaxf(a, b); }
B200_DEV float sel_min(float a, float b) { return fminf(a, b); }
B200_DEV bool far_box_covers(const DevConsts& c, f3 o, f3 d, float linear_depth) {
    float om[4], dm[4];
    mat4_mul(c.v2m, o.x, o.y, o.z, 1.0f, om);
    mat4_mul(c.v2m, d.x, d.y, d.z, 0.0f, dm);
    const float bs = c.clip_box_half;
    const float mx = 1.0f / dm[0], my = 1.0f / dm[1], mz = 1.0f / dm[2];       // util:6
    const float nx = mx * om[0], ny = my * om[1], nz = mz * om[2];            // util:7
    const float kx = fabsf(mx) * bs, ky = fabsf(my) * bs, kz = fabsf(mz) * bs; // util:8
    const float tN = sel_max(sel_max(-nx - kx, -ny - ky), -nz - kz);           // util:9-11
    const float tF = sel_min(sel_min(-nx + kx, -ny + ky), -nz + kz);           // util:10-12
    if (tN > tF || tF < 0.0f) return false;                                    // util:13-15
    return tN > 0.0f && tN < linear_depth;  // front faces in front of the camera, nearer than the opaque depth
}

// main:128-142 — ray generation from the depth texture (exact arithmetic, shader op order)
// The x- and y-terms of INV_PROJECTION_MATRIX * vec4(ndc, 1) depend on the column / the row only: ray_tables_kernel
// evaluates them once per frame size with the same operations in the same order, so the sums below are bit-identical to
// mat4_mul(inv_proj, ndc.x, ndc.y, depth, 1) = ((m0*x + m4*y) + m8*z) + m12*1 while the pixel saves two IEEE divisions,
// the ndc arithmetic and 8 of the 16 products.
B200_DEV float4 ray_col_entry(const DevConsts& c, int x) {
    const float su = (float(x) + 0.5f) / float(c.fw);          // SCREEN_UV.x of the fragment centre
    const float nx = su * 2.0f - 1.0f;                          // :130
    return make_float4(c.inv_proj[0] * nx, c.inv_proj[1] * nx, c.inv_proj[2] * nx, c.inv_proj[3] * nx);
}
B200_DEV float4 ray_row_entry(const DevConsts& c, int y) {
    const float sv = (float(y) + 0.5f) / float(c.fh);
    const float ny = sv * 2.0f - 1.0f;
    return make_float4(c.inv_proj[4] * ny, c.inv_proj[5] * ny, c.inv_proj[6] * ny, c.inv_proj[7] * ny);
}

B200_DEV void make_ray(const DevConsts& c, int x, int y, float nonlinear_depth, f3& o, f3& d, float& linear_depth, float& jitter,
                       const float4* ray_col = nullptr, const float4* ray_row = nullptr) {
    const float4 cx = ray_col ? ldg4(ray_col + x) : ray_col_entry(c, x);
    const float4 cy = ray_row ? ldg4(ray_row + y) : ray_row_entry(c, y);
    float vc[4], wc[4];
    vc[0] = cx.x + cy.x + c.inv_proj[8] * nonlinear_depth + c.inv_proj[12] * 1.0f;   // :130-131
    vc[1] = cx.y + cy.y + c.inv_proj[9] * nonlinear_depth + c.inv_proj[13] * 1.0f;
    vc[2] = cx.z + cy.z + c.inv_proj[10] * nonlinear_depth + c.inv_proj[14] * 1.0f;
    vc[3] = cx.w + cy.w + c.inv_proj[11] * nonlinear_depth + c.inv_proj[15] * 1.0f;
    mat4_mul(c.inv_view_ray, vc[0], vc[1], vc[2], vc[3], wc);                             // :134
    const f3 pos_world = mk3(wc[0] / wc[3], wc[1] / wc[3], wc[2] / wc[3]);                // :135
    const f3 dc = ld3(c.cam_pos_world) - pos_world;
    linear_depth = sqrtf(dot3(dc, dc));                                                   // :138
    o = mk3(0.f, 0.f, 0.f);                                                               // :141
    const f3 v = mk3(vc[0], vc[1], vc[2]) - o;
    const float l = sqrtf(dot3(v, v));
    d = mk3(v.x / l, v.y / l, v.z / l);                                                   // :142
    jitter = 0.0f;
    // :168-169: texelFetch(u_blue_noise_texture, ivec2(px) & ivec2(0xff), 0) — a FIXED 256x256 window whatever the texture
    // size (uploads smaller than 256 are rejected: the shader would fetch out of range)
    if (c.blue_noise) jitter = float(c.blue_noise[(y & 0xff) * c.bn_w + (x & 0xff)]) / 255.0f;
}

}  // namespace b200atmo

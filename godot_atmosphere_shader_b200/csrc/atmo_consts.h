// Host-side derivation of the per-frame kernel constants (DevConsts) from the uniform block.
// Plain fp32 C++ in the shader's op order; the including TU must be compiled without FMA
// contraction (-ffp-contract=off) so the values are what the shader computes per fragment.
#pragma once

#include <cmath>
#include <cstring>

#include "atmo_internal.h"

namespace b200atmo {

struct Variant {
    int scatter_model = B200ATMO_SCATTER_V2;
    int scatter_steps = 8;   // ATMOSPHERE_RAYMARCH_STEPS of planet_atmosphere_no_clouds.gdshader:4
    int cloud_steps = 0;
    int light_mode = B200ATMO_LIGHT_NONE;
};

struct DeviceTextures {
    const float* lut_pad = nullptr;
    const float4* lut_cells = nullptr;
    const float4* cube_cells = nullptr;
    int cube_res = 0;
    const float4* shape_cells = nullptr;
    int nx = 0, ny = 0, nz = 0;
    float cube_max = 1.0f;   // largest texel of the coverage cube as the sampler returns it (u8/255); 1 = unknown
    const uint8_t* blue_noise = nullptr;
    int bn_w = 0, bn_h = 0;
};

namespace hostmath {
inline void mat4_mul_vec(const float* m, float x, float y, float z, float w, float out[4]) {
    for (int r = 0; r < 4; ++r) out[r] = m[0 + r] * x + m[4 + r] * y + m[8 + r] * z + m[12 + r] * w;
}
// (a*b)[col][row] = sum_k a[k][row]*b[col][k], k ascending — GLSL mat4*mat4
inline void mat4_mul_mat(const float* a, const float* b, float* out) {
    for (int col = 0; col < 4; ++col)
        for (int row = 0; row < 4; ++row)
            out[col * 4 + row] = a[0 + row] * b[col * 4 + 0] + a[4 + row] * b[col * 4 + 1] + a[8 + row] * b[col * 4 + 2] +
                                 a[12 + row] * b[col * 4 + 3];
}
inline float pow4(float x) { return x * x * x * x; }
}  // namespace hostmath

// Largest height-curve value hc for which the cloud density (cloud_funcs:39-64) is exactly 0 whatever the coverage and
// shape texels are. The shader's expression is monotone non-decreasing in the coverage sample, in the shape term and
// (for a positive sum) in hc under round-to-nearest, so evaluating it IN THE SHADER'S OWN fp32 OP ORDER with the upper
// bounds of both gives an upper bound of the real value; where that bound is <= 0 the clamped density is 0. Bilinear
// filtering (fma lerps of texels <= cube_max) can exceed cube_max by a few ulps at most: 8 ulps of slack are added.
// In the shell 0 < height_ratio < 1, so `coverage - 0.25*hr` never exceeds the coverage sample. Found by bisection over
// the float bits of hc in (0, 1]; 0 when no useful bound exists.
inline float cloud_hc_min(float cube_max, float coverage_bias, float shape_hi_m01) {
    const float tex_hi = cube_max * (1.0f + 8.0f * 1.1920929e-7f);
    const float cov_hi = tex_hi + coverage_bias;
    const float covterm_hi = -1.2f * (1.0f - cov_hi) + 1.5f * cov_hi;             // GLSL mix(-1.2, 1.5, cov) as the kernels evaluate it
    const float total_hi = shape_hi_m01 + covterm_hi;
    if (!(total_hi > 0.0f)) return 1.0f;                                         // hc <= 1 always: no sample can be dense
    auto positive = [&](float hc) { return total_hi * hc * 50.0f - 20.0f > 0.0f; };
    if (positive(1.1754944e-38f)) return 0.0f;
    if (!positive(1.0f)) return 1.0f;
    uint32_t lo, hi;                                                             // invariant: !positive(lo), positive(hi)
    float flo = 1.1754944e-38f, fhi = 1.0f;
    std::memcpy(&lo, &flo, 4);
    std::memcpy(&hi, &fhi, 4);
    while (hi - lo > 1u) {
        const uint32_t mid = lo + (hi - lo) / 2u;
        float fm;
        std::memcpy(&fm, &mid, 4);
        if (positive(fm)) hi = mid;
        else lo = mid;
    }
    std::memcpy(&flo, &lo, 4);
    return flo;
}

// Largest fp32 y with (y * 50) - 20 <= 0 evaluated like the shader (cloud_funcs:62; both operations round): bisection over
// the float bits of y in [0, 1] (the expression is monotone in y).
inline float cloud_dens_y_min() {
    auto positive = [](float y) { return y * 50.0f - 20.0f > 0.0f; };
    uint32_t lo = 0u, hi;
    float fhi = 1.0f, f;
    std::memcpy(&hi, &fhi, 4);
    while (hi - lo > 1u) {
        const uint32_t mid = lo + (hi - lo) / 2u;
        std::memcpy(&f, &mid, 4);
        if (positive(f)) hi = mid;
        else lo = mid;
    }
    std::memcpy(&f, &lo, 4);
    return f;
}

// Squares of two radii r_u < r_o such that a march position whose exact distance from the planet centre is below r_u or above
// r_o has density exactly 0 with room to spare: outside the shell, or in its bottom / top rim where height_curve <= hc_min
// (cloud_hc_min). The room (1e-6 * (steps + 64) relative) covers the rounding the shader's `pos += dir * step` accumulates
// over `steps` additions (<= steps * 1.1e-7 relative) and that of the per-ray quadratics that use these bounds
// (raymarch_cloud) about ten times over. under = 0 switches the skip off (degenerate shells, absurd step counts).
inline void cloud_skip_r2(float bottom_h, float thickness, float hc_min, int steps, float& under_r2, float& over_r2) {
    under_r2 = 0.0f;
    over_r2 = 3.0e38f;
    if (!(bottom_h > 0.0f) || !(thickness > 0.0f) || steps > 100000) return;
    const double s = std::sqrt(std::fmax(0.0, 1.0 - double(std::fmin(std::fmax(hc_min, 0.0f), 1.0f))));
    const double room = 1e-6 * (double(steps) + 64.0);
    const double r_u = (double(bottom_h) + double(thickness) * (1.0 - s) * 0.5) * (1.0 - room);   // height_curve(hr) <= hc_min for hr <= (1 - s)/2
    const double r_o = (double(bottom_h) + double(thickness) * (1.0 + s) * 0.5) * (1.0 + room);   // ... and for hr >= (1 + s)/2
    if (!(r_u > 0.0) || !(r_o * r_o < 1.0e38)) return;
    float u = float(r_u * r_u), o = float(r_o * r_o);
    while (double(u) > r_u * r_u) u = std::nextafterf(u, 0.0f);
    while (double(o) < r_o * r_o) o = std::nextafterf(o, 3.0e38f);
    under_r2 = u;
    over_r2 = o;
}

// The packed copy the cloud loops read (CloudHot, atmo_internal.h); call after any of its sources changed.
inline void consts_pack_cloud_hot(DevConsts& c) {
    CloudHot& h = c.hot;
    for (int k = 0; k < 3; ++k) h.sun[k] = c.sun_dir_model[k];
    h.bottom_h = c.cloud_bottom_h;
    h.inv_thickness = c.inv_cloud_thickness;
    h.thickness = c.cloud_thickness;
    h.hc_min = c.hc_min;
    h.coverage_bias = c.coverage_bias;
    for (int k = 0; k < 4; ++k) h.rot[k] = c.rot[k];
    h.shape_hi_m01 = c.shape_hi_m01;
    h.dens_y_min = c.dens_y_min;
    h.shape_scale = c.shape_scale;
    h.shape_factor = c.shape_factor;
    h.shape_mix0 = c.shape_mix0;
    h.density_scale = c.density_scale;
    h.shape_invert = c.shape_invert;
    h.light_reach = c.light_reach;
    h.cube_cells = c.cube_cells;
    h.shape_cells = c.shape_cells;
    h.cube_res = c.cube_res;
    h.nx = c.shape_nx;
    h.ny = c.shape_ny;
    h.nz = c.shape_nz;
    cloud_skip_r2(c.cloud_bottom_h, c.cloud_thickness, c.hc_min, c.cloud_steps, h.under_r2, h.over_r2);
    for (float& x : h.pad_) x = 0.0f;
}

// Uniform-only part (everything that does not depend on the frame).
inline void consts_from_params(DevConsts& c, const B200AtmoParams& p, const Variant& v, const DeviceTextures& t) {
    std::memset(&c, 0, sizeof(c));
    c.R = p.planet_radius;
    c.H = p.atmosphere_height;
    c.rho = p.density;
    c.atmo_radius = p.planet_radius + p.atmosphere_height;  // main:144
    c.sphere_depth_factor = p.sphere_depth_factor;
    for (int k = 0; k < 3; ++k) {
        c.coef[k] = hostmath::pow4(400.0f / p.scattering_wavelengths[k]) * p.scattering_strength;  // funcs_v2:47-51
        c.neg_coef_log2e[k] = -c.coef[k] * 1.4426950408889634f;
        c.ambient[k] = p.atmosphere_ambient_color[k];
        c.modulate[k] = p.atmosphere_modulate[k];
        c.day0[k] = p.day_color0[k];
        c.day1[k] = p.day_color1[k];
        c.night0[k] = p.night_color0[k];
        c.night1[k] = p.night_color1[k];
    }
    c.inv_H = 1.0f / c.H;
    c.rho2 = c.rho * c.rho;
    c.day_night_scale = p.day_night_transition_scale;
    c.lut_pad = t.lut_pad;
    c.lut_cells = t.lut_cells;
    // clouds
    c.cloud_bottom_h = p.planet_radius + p.cloud_bottom * p.atmosphere_height;  // cloud_funcs:260
    c.cloud_top_h = p.planet_radius + p.cloud_top * p.atmosphere_height;        // cloud_funcs:261
    c.cloud_thickness = c.cloud_top_h - c.cloud_bottom_h;
    c.inv_cloud_thickness = 1.0f / c.cloud_thickness;
    c.density_scale = p.cloud_density_scale;
    c.cloud_blend = p.cloud_blend;
    c.coverage_bias = p.cloud_coverage_bias;
    c.shape_factor = p.cloud_shape_factor;
    c.shape_scale = p.cloud_shape_scale;
    c.shape_invert = (p.cloud_shape_invert == 1.0f) ? 1 : 0;                     // cloud_funcs:57
    for (int k = 0; k < 4; ++k) c.rot[k] = p.cloud_coverage_rotation[k];
    {
        const float ratio = c.R / c.cloud_top_h;                                 // ground_height / top_height
        c.march_space = 0.5f * std::sqrt(1.0f - ratio * ratio) * c.cloud_bottom_h;  // cloud_funcs:186-189
        c.march_ground = 3.0f * c.march_space;                                   // :190
        c.march_hmin = c.cloud_bottom_h;                                         // :191
        c.march_hmax = c.cloud_top_h * 1.05f;                                    // :192
        c.light_reach = (c.cloud_top_h - c.cloud_bottom_h) * 0.15f;              // :108
    }
    {
        // Largest value `shape` (cloud_funcs:48-59) can take for any texel in [0,1], in the shader's own arithmetic:
        // mix(0.5, tex, factor) is monotone in tex, the optional invert flips it.
        const float f = c.shape_factor;
        const float s0 = 0.5f * (1.0f - f) + 0.0f * f, s1 = 0.5f * (1.0f - f) + 1.0f * f;
        const float lo = s0 < s1 ? s0 : s1, hi = s0 < s1 ? s1 : s0;
        const float shape_hi = c.shape_invert ? 1.0f - lo : hi;
        c.shape_hi_m01 = shape_hi - 0.2f * 0.5f;
    }
    c.hc_min = cloud_hc_min(t.cube_max, c.coverage_bias, c.shape_hi_m01);
    c.shape_mix0 = 0.5f * (1.0f - c.shape_factor);       // GLSL mix(0.5, tex, f) = 0.5*(1-f) + tex*f, the first product
    c.dens_y_min = cloud_dens_y_min();
    c.cube_cells = t.cube_cells;
    c.cube_res = t.cube_res;
    c.shape_cells = t.shape_cells;
    c.shape_nx = t.nx;
    c.shape_ny = t.ny;
    c.shape_nz = t.nz;
    c.blue_noise = t.blue_noise;
    c.bn_w = t.bn_w;
    c.bn_h = t.bn_h;
    c.scatter_steps = v.scatter_steps;
    c.cloud_steps = v.cloud_steps > 0 ? v.cloud_steps : 1;
    consts_pack_cloud_hot(c);
}

// Frame-dependent part: the varyings (planet/sun centre in view space) and INV_VIEW_MATRIX.
inline void consts_set_frame(DevConsts& c, const B200AtmoParams& p, const float planet_center_view[3],
                             const float sun_center_view[3], const float inv_view[16]) {
    float v[3];
    for (int k = 0; k < 3; ++k) {
        c.C[k] = planet_center_view[k];
        v[k] = sun_center_view[k] - planet_center_view[k];
    }
    const float len = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    for (int k = 0; k < 3; ++k) c.sun_dir[k] = v[k] / len;                       // main:164
    hostmath::mat4_mul_mat(p.world_to_model, inv_view, c.v2m);                   // cloud_funcs:285
    float s4[4];
    hostmath::mat4_mul_vec(c.v2m, c.sun_dir[0], c.sun_dir[1], c.sun_dir[2], 0.0f, s4);  // cloud_funcs:288
    for (int k = 0; k < 3; ++k) c.sun_dir_model[k] = s4[k];
    consts_pack_cloud_hot(c);
}

// Frame API: derive the varyings like atmosphere_vertex (main:101-103) and the ray-generation matrices.
inline void consts_set_camera(DevConsts& c, const B200AtmoParams& p, const B200AtmoCamera& cam, int w, int h, int row_begin,
                              int row_end) {
    float world_pos[4], pc[4], sc[4];
    hostmath::mat4_mul_vec(cam.model, 0.0f, 0.0f, 0.0f, 1.0f, world_pos);
    hostmath::mat4_mul_vec(cam.view, world_pos[0], world_pos[1], world_pos[2], world_pos[3], pc);
    hostmath::mat4_mul_vec(cam.view, p.sun_position[0], p.sun_position[1], p.sun_position[2], 1.0f, sc);
    float inv_view[16];
    std::memcpy(inv_view, cam.inv_view, sizeof(inv_view));
    if (cam.double_precision) {  // main:118-125
        inv_view[12] *= -1.0f;
        inv_view[13] *= -1.0f;
        inv_view[14] *= -1.0f;
    }
    consts_set_frame(c, p, pc, sc, inv_view);
    std::memcpy(c.inv_proj, cam.inv_projection, sizeof(c.inv_proj));
    std::memcpy(c.inv_view_ray, inv_view, sizeof(c.inv_view_ray));
    float cp[4];
    hostmath::mat4_mul_vec(inv_view, 0.0f, 0.0f, 0.0f, 1.0f, cp);                // main:136
    for (int k = 0; k < 3; ++k) c.cam_pos_world[k] = cp[k];
    c.fw = w;
    c.fh = h;
    c.row_begin = row_begin;
    c.row_end = row_end;
    c.clip_box_half = cam.clip_box_size > 0.0f ? cam.clip_box_size * 0.5f : 0.0f;
    c.row_pitch = 8;
}

}  // namespace b200atmo

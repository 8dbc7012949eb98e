// TEST INFRASTRUCTURE ONLY — never loaded by the product.
// Compiles the product's device functions (csrc/atmo_device.cuh) as plain host C++ so the logic of the
// kernels can be exercised without a GPU (`-m "not gpu"` tests). MUFU approximations are replaced by
// libm (ex2 -> exp2f, rsqrt -> 1/sqrtf, rcp -> 1/x), so this checks logic and rounding policy, not the
// hardware approximations; the GPU parity tests do that.
#include "../../godot_atmosphere_shader_b200/csrc/atmo_consts.h"
#include "../../godot_atmosphere_shader_b200/csrc/atmo_device.cuh"

using namespace b200atmo;

extern "C" {

typedef struct HostsimTextures {
    const float* lut_pad;
    const float* cube_pad;
    int32_t cube_res;
    const float* shape_pad;
    int32_t nx, ny, nz;
    const uint8_t* blue_noise;
    int32_t bn_w, bn_h;
} HostsimTextures;

}  // extern "C"

#include <vector>
static std::vector<float4> g_cells, g_cube, g_shape;  // host restatement of lut_cells_kernel (single-threaded test helper)
static DeviceTextures tex_of(const HostsimTextures* t) {
    DeviceTextures d;
    d.lut_pad = t->lut_pad;
    g_cells.resize(size_t(kLutCells) * kLutCells);
    for (int yi = 0; yi < kLutCells; ++yi)
        for (int xi = 0; xi < kLutCells; ++xi) g_cells[size_t(yi) * kLutCells + xi] = make_lut_cell(t->lut_pad, xi, yi);
    d.lut_cells = g_cells.data();
    {   // host restatement of cube_cells_kernel / shape_cells_kernel
        const int res = t->cube_res, rc = res + 1;
        g_cube.resize(size_t(6) * rc * rc);
        for (int f = 0; f < 6; ++f)
            for (int yi = 0; yi < rc; ++yi)
                for (int xi = 0; xi < rc; ++xi) g_cube[(size_t(f) * rc + yi) * rc + xi] = make_cube_cell(t->cube_pad, res, f, yi, xi);
        const int cx = t->nx + 1, cy = t->ny + 1, cz = t->nz + 1;
        g_shape.resize(size_t(cx) * cy * cz * 2);
        for (int zi = 0; zi < cz; ++zi)
            for (int yi = 0; yi < cy; ++yi)
                for (int xi = 0; xi < cx; ++xi) {
                    const size_t idx = (size_t(zi) * cy + yi) * cx + xi;
                    g_shape[2 * idx] = make_shape_cell(t->shape_pad, t->nx, t->ny, zi, yi, xi);
                    g_shape[2 * idx + 1] = make_shape_cell(t->shape_pad, t->nx, t->ny, zi + 1, yi, xi);
                }
    }
    d.cube_cells = g_cube.data();
    d.cube_res = t->cube_res;
    d.cube_max = 0.0f;   // as upload_cube (atmo_capi.cu): the largest texel bounds the cloud density
    for (size_t i = 0; i < size_t(6) * (t->cube_res + 2) * (t->cube_res + 2); ++i) d.cube_max = t->cube_pad[i] > d.cube_max ? t->cube_pad[i] : d.cube_max;
    d.shape_cells = g_shape.data();
    d.nx = t->nx; d.ny = t->ny; d.nz = t->nz;
    d.blue_noise = t->blue_noise;
    d.bn_w = t->bn_w; d.bn_h = t->bn_h;
    return d;
}

template <int M, int L>
static void run_rays(const DevConsts& c, const float* od, const float* dj, size_t n, float* rgba, uint8_t* disc) {
    for (size_t i = 0; i < n; ++i) {
        float4 out;
        bool d = shade_ray<M, L>(c, mk3(od[4 * i], od[4 * i + 1], od[4 * i + 2]), mk3(dj[4 * i], dj[4 * i + 1], dj[4 * i + 2]),
                                 od[4 * i + 3], dj[4 * i + 3], out);
        rgba[4 * i] = out.x; rgba[4 * i + 1] = out.y; rgba[4 * i + 2] = out.z; rgba[4 * i + 3] = out.w;
        if (disc) disc[i] = d ? 1 : 0;
    }
}

extern "C" {

void hostsim_render_rays(const B200AtmoParams* p, const int32_t variant[4], const B200AtmoFrame* fr, const HostsimTextures* t,
                         const float* od, const float* dj, size_t n, float* rgba, uint8_t* disc) {
    Variant v;
    v.scatter_model = variant[0]; v.scatter_steps = variant[1]; v.cloud_steps = variant[2]; v.light_mode = variant[3];
    DevConsts c;
    consts_from_params(c, *p, v, tex_of(t));
    consts_set_frame(c, *p, fr->planet_center_view, fr->sun_center_view, fr->inv_view);
    // like launch_rays_t (atmo_kernels.cu): power-of-two texture sizes select the instantiation with the fused texel coordinates
    auto pow2 = [](int x) { return x > 0 && (x & (x - 1)) == 0; };
    const bool all_pow2 = pow2(t->cube_res) && pow2(t->nx) && pow2(t->ny) && pow2(t->nz);
    const int key = v.scatter_model * 8 + (v.light_mode ? (v.light_mode | (all_pow2 ? kLightPow2 : 0)) : 0);
    switch (key) {
        case 0: run_rays<0, 0>(c, od, dj, n, rgba, disc); break;
        case 1: run_rays<0, 1>(c, od, dj, n, rgba, disc); break;
        case 2: run_rays<0, 2>(c, od, dj, n, rgba, disc); break;
        case 5: run_rays<0, 5>(c, od, dj, n, rgba, disc); break;
        case 6: run_rays<0, 6>(c, od, dj, n, rgba, disc); break;
        case 8: run_rays<1, 0>(c, od, dj, n, rgba, disc); break;
        case 9: run_rays<1, 1>(c, od, dj, n, rgba, disc); break;
        case 10: run_rays<1, 2>(c, od, dj, n, rgba, disc); break;
        case 13: run_rays<1, 5>(c, od, dj, n, rgba, disc); break;
        default: run_rays<1, 6>(c, od, dj, n, rgba, disc); break;
    }
}

// frame front-end + frame constants (make_rays_kernel / b200atmo_make_rays on the host)
void hostsim_make_rays(const B200AtmoParams* p, const B200AtmoCamera* cam, const HostsimTextures* t, const float* depth, int w,
                       int h, float* od, float* dj, B200AtmoFrame* frame_out) {
    Variant v;
    DevConsts c;
    consts_from_params(c, *p, v, tex_of(t));
    consts_set_camera(c, *p, *cam, w, h, 0, h);
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            size_t i = size_t(y) * w + x;
            f3 o, d;
            float ld, jit;
            make_ray(c, x, y, depth[i], o, d, ld, jit);
            od[4 * i] = o.x; od[4 * i + 1] = o.y; od[4 * i + 2] = o.z; od[4 * i + 3] = ld;
            dj[4 * i] = d.x; dj[4 * i + 1] = d.y; dj[4 * i + 2] = d.z; dj[4 * i + 3] = jit;
        }
    if (frame_out) {
        for (int k = 0; k < 3; ++k) frame_out->planet_center_view[k] = c.C[k];
        std::memcpy(frame_out->inv_view, c.inv_view_ray, sizeof(frame_out->inv_view));
        // sun centre is not kept in DevConsts; recompute like consts_set_camera
        float sc[4];
        hostmath::mat4_mul_vec(cam->view, p->sun_position[0], p->sun_position[1], p->sun_position[2], 1.0f, sc);
        for (int k = 0; k < 3; ++k) frame_out->sun_center_view[k] = sc[k];
    }
}

void hostsim_noise_cubemap(const B200AtmoNoise* noise, int res, const float scale[3], uint8_t* out) {
    for (int side = 0; side < 6; ++side)
        for (int y = 0; y < res; ++y)
            for (int x = 0; x < res; ++x) out[(size_t(side) * res + y) * res + x] = noise_cube_texel(side, x, y, res, scale, *noise);
}

float hostsim_sqrt_refined(float x) { float inv; return sqrt_refined(x, inv); }
float hostsim_div_refined(float a, float b) { return div_refined(a, b, 1.0f / b); }
int hostsim_floor_frac(float x, float* frac) { return floor_frac(x, *frac); }

// steps covered by the under-shell skip of raymarch_cloud since the last call (builds with -DB200ATMO_STEP_STATS; else -1)
long long hostsim_take_skipped_steps() {
#ifdef B200ATMO_STEP_STATS
    const long long v = b200atmo_step_stats_skipped;
    b200atmo_step_stats_skipped = 0;
    return v;
#else
    return -1;
#endif
}

}  // extern "C"

// ORACLE — TEST INFRASTRUCTURE ONLY (see atmo_oracle.hpp). C exports for ctypes.
// Build: make -C oracle   (g++ -O2 -ffp-contract=off, no fast-math, no intrinsics)
#include "atmo_oracle.hpp"

#include <atomic>
#include <thread>

using namespace oracle;

extern "C" {

typedef struct OracleTextures {
    const float* lut;          // 256*256 fp32 (f32 entry points)
    const double* lut64;       // 256*256 fp64 (f64 entry points; may be NULL -> converted from lut)
    const uint8_t* shape;      // nx*ny*nz or NULL
    int32_t nx, ny, nz;
    const uint8_t* cube_padded;  // 6*(res+2)^2 from oracle_cube_build_padded, or NULL
    int32_t cube_res;
    const uint8_t* blue_noise;   // bn_w*bn_h (frame API only)
    int32_t bn_w, bn_h;
} OracleTextures;

typedef struct OracleVariant {
    int32_t scatter_model, scatter_steps, cloud_steps, light_mode;
} OracleVariant;

int oracle_hardware_threads(void) {
    unsigned n = std::thread::hardware_concurrency();
    return n ? int(n) : 1;
}

}  // extern "C"

namespace {

template <class F> void parallel_for(size_t n, int threads, F&& body) {
    if (threads <= 0) threads = oracle_hardware_threads();
    if (threads == 1 || n < 1024) {
        body(size_t(0), n);
        return;
    }
    // dynamic chunks: rays that miss are much cheaper than rays that march
    const size_t chunk = 4096;
    std::atomic<size_t> next{0};
    std::vector<std::thread> pool;
    pool.reserve(size_t(threads));
    for (int t = 0; t < threads; ++t)
        pool.emplace_back([&] {
            for (;;) {
                size_t b = next.fetch_add(chunk);
                if (b >= n) break;
                body(b, std::min(n, b + chunk));
            }
        });
    for (auto& th : pool) th.join();
}

template <class T> Uniforms<T> bind(const B200AtmoParams* p, const OracleTextures* tex, const T* lut) {
    Uniforms<T> u = make_uniforms<T>(*p);
    u.lut = lut;
    if (tex) {
        u.shape = tex->shape;
        u.shape_nx = tex->nx;
        u.shape_ny = tex->ny;
        u.shape_nz = tex->nz;
        u.cube_padded = tex->cube_padded;
        u.cube_res = tex->cube_res;
    }
    return u;
}

template <class T> void bake_lut(const B200AtmoParams* p, T* out) {
    Uniforms<T> u = make_uniforms<T>(*p);
    for (int j = 0; j < B200ATMO_LUT_SIZE; ++j)
        for (int i = 0; i < B200ATMO_LUT_SIZE; ++i) out[j * B200ATMO_LUT_SIZE + i] = bake_texel(u, i, j);
}

template <class T>
void render_rays(const B200AtmoParams* p, const OracleVariant* v, const B200AtmoFrame* fr, const OracleTextures* tex,
                 const T* lut, const float* origin_depth, const float* dir_jitter, size_t n, T* rgba, uint8_t* discard,
                 int threads) {
    Uniforms<T> u = bind<T>(p, tex, lut);
    Variant var{v->scatter_model, v->scatter_steps, v->cloud_steps, v->light_mode};
    vec3<T> pc = load_vec3<T>(fr->planet_center_view), sc = load_vec3<T>(fr->sun_center_view);
    mat4<T> inv_view = load_mat4<T>(fr->inv_view);
    parallel_for(n, threads, [&](size_t b, size_t e) {
        for (size_t i = b; i < e; ++i) {
            vec3<T> o = {T(origin_depth[4 * i]), T(origin_depth[4 * i + 1]), T(origin_depth[4 * i + 2])};
            T depth = T(origin_depth[4 * i + 3]);
            vec3<T> d = {T(dir_jitter[4 * i]), T(dir_jitter[4 * i + 1]), T(dir_jitter[4 * i + 2])};
            T jitter = T(dir_jitter[4 * i + 3]);
            vec3<T> albedo;
            T alpha;
            bool disc = fragment_from_ray(u, var, o, d, depth, jitter, pc, sc, inv_view, &albedo, &alpha);
            rgba[4 * i] = albedo.x;
            rgba[4 * i + 1] = albedo.y;
            rgba[4 * i + 2] = albedo.z;
            rgba[4 * i + 3] = alpha;
            if (discard) discard[i] = disc ? 1 : 0;
        }
    });
}

template <class T>
void render_frame(const B200AtmoParams* p, const OracleVariant* v, const B200AtmoCamera* cam, const OracleTextures* tex,
                  const T* lut, const float* depth, int w, int h, int row_begin, int row_end, T* rgba, uint8_t* discard,
                  int threads) {
    Uniforms<T> u = bind<T>(p, tex, lut);
    Variant var{v->scatter_model, v->scatter_steps, v->cloud_steps, v->light_mode};
    mat4<T> inv_proj = load_mat4<T>(cam->inv_projection), inv_view = load_mat4<T>(cam->inv_view);
    mat4<T> view = load_mat4<T>(cam->view), model = load_mat4<T>(cam->model);
    vec3<T> pc, sc;
    atmosphere_vertex_varyings(model, view, load_vec3<T>(p->sun_position), &pc, &sc);
    const bool dp = cam->double_precision != 0;
    mat4<T> inv_view_frag = inv_view;  // the by-value parameter of atmosphere_fragment, after :118-125
    if (dp) {
        inv_view_frag.c[3][0] *= T(-1);
        inv_view_frag.c[3][1] *= T(-1);
        inv_view_frag.c[3][2] *= T(-1);
    }
    const size_t n = size_t(row_end - row_begin) * size_t(w);
    parallel_for(n, threads, [&](size_t b, size_t e) {
        for (size_t k = b; k < e; ++k) {
            int y = row_begin + int(k / size_t(w)), x = int(k % size_t(w));
            size_t i = size_t(y) * w + x;
            // SCREEN_UV of the fragment centre
            T su = (T(x) + T(0.5)) / T(w), sv = (T(y) + T(0.5)) / T(h);
            vec3<T> o, d;
            T linear_depth;
            fragment_make_ray(inv_proj, inv_view, dp, T(depth[i]), su, sv, &o, &d, &linear_depth);
            // main:168-169: texelFetch(tex, ivec2(viewport_size*screen_uv) & ivec2(0xff), 0): a fixed 256x256 window (bn_w, bn_h >= 256)
            T jitter = T(0);
            if (tex && tex->blue_noise)
                jitter = T(tex->blue_noise[size_t(y & 0xff) * tex->bn_w + (x & 0xff)]) / T(255);
            vec3<T> albedo = {T(0), T(0), T(0)};
            T alpha = T(0);
            bool disc = true;
            if (!(cam->clip_box_size > 0.0f) || far_box_covers(u, inv_view_frag, T(cam->clip_box_size), o, d, linear_depth))
                disc = fragment_from_ray(u, var, o, d, linear_depth, jitter, pc, sc, inv_view_frag, &albedo, &alpha);
            rgba[4 * i] = albedo.x;
            rgba[4 * i + 1] = albedo.y;
            rgba[4 * i + 2] = albedo.z;
            rgba[4 * i + 3] = alpha;
            if (discard) discard[i] = disc ? 1 : 0;
        }
    });
}

std::vector<double> widen_lut(const OracleTextures* tex) {
    std::vector<double> l(size_t(B200ATMO_LUT_SIZE) * B200ATMO_LUT_SIZE);
    for (size_t i = 0; i < l.size(); ++i) l[i] = double(tex->lut[i]);
    return l;
}

}  // namespace

extern "C" {

// ---- LUT bake (optical_depth.gdshader) -------------------------------------------------------------
void oracle_bake_lut_f32(const B200AtmoParams* p, float* out) { bake_lut<float>(p, out); }
void oracle_bake_lut_f64(const B200AtmoParams* p, double* out) { bake_lut<double>(p, out); }
// Full reference round trip: shader output -> RGBA8 viewport bytes -> FORMAT_RF reinterpretation
void oracle_bake_lut_via_rgba8(const B200AtmoParams* p, float* out) {
    std::vector<float> tmp(size_t(B200ATMO_LUT_SIZE) * B200ATMO_LUT_SIZE);
    bake_lut<float>(p, tmp.data());
    for (size_t i = 0; i < tmp.size(); ++i) {
        uint8_t px[4];
        encode_float_to_viewport(tmp[i], px);
        out[i] = decode_viewport_bytes(px);
    }
}

void oracle_cube_build_padded(const uint8_t* faces, int res, uint8_t* out) { cube_build_padded(faces, res, out); }

// ---- ray batch / frame -------------------------------------------------------------------------------
void oracle_render_rays_f32(const B200AtmoParams* p, const OracleVariant* v, const B200AtmoFrame* fr,
                            const OracleTextures* tex, const float* origin_depth, const float* dir_jitter, size_t n,
                            float* rgba, uint8_t* discard, int threads) {
    render_rays<float>(p, v, fr, tex, tex->lut, origin_depth, dir_jitter, n, rgba, discard, threads);
}
void oracle_render_rays_f64(const B200AtmoParams* p, const OracleVariant* v, const B200AtmoFrame* fr,
                            const OracleTextures* tex, const float* origin_depth, const float* dir_jitter, size_t n,
                            double* rgba, uint8_t* discard, int threads) {
    std::vector<double> tmp;
    const double* lut = tex->lut64;
    if (!lut) { tmp = widen_lut(tex); lut = tmp.data(); }
    render_rays<double>(p, v, fr, tex, lut, origin_depth, dir_jitter, n, rgba, discard, threads);
}
void oracle_render_frame_f32(const B200AtmoParams* p, const OracleVariant* v, const B200AtmoCamera* cam,
                             const OracleTextures* tex, const float* depth, int w, int h, int row_begin, int row_end,
                             float* rgba, uint8_t* discard, int threads) {
    render_frame<float>(p, v, cam, tex, tex->lut, depth, w, h, row_begin, row_end, rgba, discard, threads);
}
void oracle_render_frame_f64(const B200AtmoParams* p, const OracleVariant* v, const B200AtmoCamera* cam,
                             const OracleTextures* tex, const float* depth, int w, int h, int row_begin, int row_end,
                             double* rgba, uint8_t* discard, int threads) {
    std::vector<double> tmp;
    const double* lut = tex->lut64;
    if (!lut) { tmp = widen_lut(tex); lut = tmp.data(); }
    render_frame<double>(p, v, cam, tex, lut, depth, w, h, row_begin, row_end, rgba, discard, threads);
}
// Frame front-end only (main:101-103, 128-142): rays as the batch API wants them + the frame constants.
void oracle_make_rays_f32(const B200AtmoParams* p, const B200AtmoCamera* cam, const OracleTextures* tex, const float* depth,
                          int w, int h, float* origin_depth, float* dir_jitter, B200AtmoFrame* frame_out) {
    mat4<float> inv_proj = load_mat4<float>(cam->inv_projection), inv_view = load_mat4<float>(cam->inv_view);
    vec3<float> pc, sc;
    atmosphere_vertex_varyings(load_mat4<float>(cam->model), load_mat4<float>(cam->view), load_vec3<float>(p->sun_position), &pc, &sc);
    const bool dp = cam->double_precision != 0;
    if (frame_out) {
        frame_out->planet_center_view[0] = pc.x; frame_out->planet_center_view[1] = pc.y; frame_out->planet_center_view[2] = pc.z;
        frame_out->sun_center_view[0] = sc.x; frame_out->sun_center_view[1] = sc.y; frame_out->sun_center_view[2] = sc.z;
        for (int i = 0; i < 16; ++i) frame_out->inv_view[i] = cam->inv_view[i];
        if (dp) for (int i = 12; i < 15; ++i) frame_out->inv_view[i] *= -1.0f;
    }
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            size_t i = size_t(y) * w + x;
            float su = (float(x) + 0.5f) / float(w), sv = (float(y) + 0.5f) / float(h);
            vec3<float> o, d;
            float ld;
            fragment_make_ray(inv_proj, inv_view, dp, depth[i], su, sv, &o, &d, &ld);
            float jitter = 0.f;
            if (tex && tex->blue_noise)
                jitter = float(tex->blue_noise[size_t(y & 0xff) * tex->bn_w + (x & 0xff)]) / 255.0f;
            origin_depth[4 * i] = o.x; origin_depth[4 * i + 1] = o.y; origin_depth[4 * i + 2] = o.z; origin_depth[4 * i + 3] = ld;
            dir_jitter[4 * i] = d.x; dir_jitter[4 * i + 1] = d.y; dir_jitter[4 * i + 2] = d.z; dir_jitter[4 * i + 3] = jitter;
        }
}

// ---- known-answer-test hooks (one per shader function) -----------------------------------------------
void oracle_ray_sphere_f32(const float c[3], float r, const float o[3], const float d[3], float out[2]) {
    vec2<float> rs = ray_sphere<float>({c[0], c[1], c[2]}, r, {o[0], o[1], o[2]}, {d[0], d[1], d[2]});
    out[0] = rs.x; out[1] = rs.y;
}
void oracle_ray_sphere_f64(const double c[3], double r, const double o[3], const double d[3], double out[2]) {
    vec2<double> rs = ray_sphere<double>({c[0], c[1], c[2]}, r, {o[0], o[1], o[2]}, {d[0], d[1], d[2]});
    out[0] = rs.x; out[1] = rs.y;
}
void oracle_ray_box_f32(const float ro[3], const float rd[3], const float box[3], float out[2]) {
    vec2<float> r = ray_box_intersection<float>({ro[0], ro[1], ro[2]}, {rd[0], rd[1], rd[2]}, {box[0], box[1], box[2]});
    out[0] = r.x; out[1] = r.y;
}
float oracle_atmosphere_density_f32(const B200AtmoParams* p, float height) {
    return get_atmosphere_density(make_uniforms<float>(*p), height);
}
void oracle_blend_colors_f32(const float s[4], const float o[4], float out[4]) {
    vec4<float> r = blend_colors<float>({s[0], s[1], s[2], s[3]}, {o[0], o[1], o[2], o[3]});
    out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}
float oracle_sample_lut_f32(const float* lut, float u, float v) { return sample_lut<float>(lut, u, v); }
float oracle_sample_shape_f32(const B200AtmoParams* p, const OracleTextures* tex, const float pos[3]) {
    Uniforms<float> u = bind<float>(p, tex, tex->lut);
    return sample_shape3d(u, vec3<float>{pos[0], pos[1], pos[2]});
}
float oracle_sample_cube_f32(const B200AtmoParams* p, const OracleTextures* tex, const float dir[3]) {
    Uniforms<float> u = bind<float>(p, tex, tex->lut);
    return sample_cube(u, vec3<float>{dir[0], dir[1], dir[2]});
}
static CloudSettings<float> cloud_settings_of(const Uniforms<float>& u) {
    CloudSettings<float> cs;
    cs.bottom_height = u.u_planet_radius + u.u_cloud_bottom * u.u_atmosphere_height;
    cs.top_height = u.u_planet_radius + u.u_cloud_top * u.u_atmosphere_height;
    cs.density_scale = u.u_cloud_density_scale;
    cs.ground_height = u.u_planet_radius;
    return cs;
}
float oracle_cloud_density_f32(const B200AtmoParams* p, const OracleTextures* tex, const float pos[3]) {
    Uniforms<float> u = bind<float>(p, tex, tex->lut);
    return get_density_full(u, vec3<float>{pos[0], pos[1], pos[2]}, cloud_settings_of(u));
}
float oracle_cloud_light_f32(const B200AtmoParams* p, const OracleTextures* tex, int light_mode, const float pos[3],
                             const float ray_dir[3], const float sun_dir[3], float jitter, float alpha) {
    Uniforms<float> u = bind<float>(p, tex, tex->lut);
    return get_light(u, light_mode, vec3<float>{pos[0], pos[1], pos[2]}, vec3<float>{ray_dir[0], ray_dir[1], ray_dir[2]},
                     vec3<float>{sun_dir[0], sun_dir[1], sun_dir[2]}, jitter, alpha, cloud_settings_of(u));
}
// raymarch_cloud in model space: out = (total_light, alpha)
void oracle_raymarch_cloud_f32(const B200AtmoParams* p, const OracleTextures* tex, int steps, int light_mode,
                               const float o[3], const float d[3], float t_begin, float t_end, float jitter,
                               const float sun[3], float out[2]) {
    Uniforms<float> u = bind<float>(p, tex, tex->lut);
    vec2<float> r = raymarch_cloud(u, steps, light_mode, vec3<float>{o[0], o[1], o[2]}, vec3<float>{d[0], d[1], d[2]}, t_begin,
                                   t_end, jitter, vec3<float>{sun[0], sun[1], sun[2]}, cloud_settings_of(u));
    out[0] = r.x; out[1] = r.y;
}
// compute_atmosphere_v2 alone: out = rgba
void oracle_compute_atmosphere_v2_f32(const B200AtmoParams* p, const float* lut, int steps, const float o[3], const float d[3],
                                      const float pc[3], float t_begin, float t_end, const float sun_dir[3], float jitter,
                                      float out[4]) {
    Uniforms<float> u = make_uniforms<float>(*p);
    u.lut = lut;
    vec4<float> r = compute_atmosphere_v2(u, steps, vec3<float>{o[0], o[1], o[2]}, vec3<float>{d[0], d[1], d[2]},
                                          vec3<float>{pc[0], pc[1], pc[2]}, t_begin, t_end, 0.0f,
                                          vec3<float>{sun_dir[0], sun_dir[1], sun_dir[2]}, jitter);
    out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}
void oracle_noise_cubemap(const B200AtmoNoise* noise, int res, const float scale[3], uint8_t* out) {
    noisegen::generate_images(*noise, res, scale, out);
}
float oracle_noise3(float x, float y, float z, uint32_t seed) { return noisegen::noise3(x, y, z, seed); }
void oracle_cubemap_atlas(const uint8_t* faces, int res, uint8_t* atlas) { noisegen::importable_image(faces, res, atlas); }
void oracle_encode_float(float h, uint8_t out[4]) { encode_float_to_viewport(h, out); }
float oracle_decode_float(const uint8_t in[4]) { return decode_viewport_bytes(in); }

}  // extern "C"

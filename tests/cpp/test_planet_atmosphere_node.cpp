// Tests of the C++ PlanetAtmosphere / OpticalDepthBaker core (godot_atmosphere_shader_b200/csrc/node/).
//
//   test_node logic                      host logic against a RECORDING stand-in of the C-ABI table: no GPU, no compute
//   test_node render <in.bin> <out.bin>  drives the real library on cuda:0 through the node (render_host) with the scene
//                                        tests/test_cpp_node.py wrote; pytest compares the output with the Python mirror
//                                        (bit-identical) and the oracle
// The cases follow tests/test_planet_atmosphere_node.py line by line, citing planet_atmosphere.gd / optical_depth_baker.gd.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include "../../godot_atmosphere_shader_b200/csrc/node/planet_atmosphere_node.hpp"

using namespace b200atmo;

static int g_checks = 0, g_failed = 0;
#define CHECK(cond)                                                                  \
    do {                                                                             \
        ++g_checks;                                                                  \
        if (!(cond)) {                                                               \
            ++g_failed;                                                              \
            std::fprintf(stderr, "%s:%d: CHECK failed: %s\n", __FILE__, __LINE__, #cond); \
        }                                                                            \
    } while (0)
static bool approx(double a, double b, double tol = 1e-6) { return std::fabs(a - b) <= tol * (1.0 + std::fabs(b)); }

// ---- recording stand-in for the C-ABI (host-logic tests only) ---------------------------------------------------
struct Call {
    std::string name;
    int a[4];
};
static std::vector<Call> g_calls;
static B200AtmoParams g_last_params;
static int count(const char* name) {
    int n = 0;
    for (const Call& c : g_calls) n += c.name == name;
    return n;
}
static bool has_variant(int model, int steps, int csteps, int light) {
    for (const Call& c : g_calls)
        if (c.name == "set_variant" && c.a[0] == model && c.a[1] == steps && c.a[2] == csteps && c.a[3] == light) return true;
    return false;
}
static int f_create(int, b200atmo_ctx** out) {
    g_calls.push_back({"create", {}});
    *out = reinterpret_cast<b200atmo_ctx*>(0x1);
    return 0;
}
static void f_destroy(b200atmo_ctx*) { g_calls.push_back({"destroy", {}}); }
static const char* f_last_error(const b200atmo_ctx*) { return "fake"; }
static int f_set_params(b200atmo_ctx*, const B200AtmoParams* p) {
    g_calls.push_back({"set_params", {}});
    g_last_params = *p;
    return 0;
}
static int f_set_variant(b200atmo_ctx*, int m, int s, int c, int l) {
    g_calls.push_back({"set_variant", {m, s, c, l}});
    return 0;
}
static int f_bn(b200atmo_ctx*, const uint8_t*, int w, int h) {
    g_calls.push_back({"upload_blue_noise", {w, h, 0, 0}});
    return 0;
}
static int f_shape(b200atmo_ctx*, const uint8_t*, int x, int y, int z) {
    g_calls.push_back({"upload_shape3d", {x, y, z, 0}});
    return 0;
}
static int f_cube(b200atmo_ctx*, const uint8_t*, int r) {
    g_calls.push_back({"upload_coverage_cube", {r, 0, 0, 0}});
    return 0;
}
static int f_bake(b200atmo_ctx*, void*) {
    g_calls.push_back({"bake_optical_depth", {}});
    return 0;
}
static int f_rf(b200atmo_ctx*, const B200AtmoCamera*, const float*, int, int, int, int, float*, uint8_t*, void*) {
    g_calls.push_back({"render_frame", {}});
    return 0;
}
static int f_rfc(b200atmo_ctx*, const B200AtmoCamera*, const float*, int, int, int, int, float*, void*) {
    g_calls.push_back({"render_frame_composite", {}});
    return 0;
}
static int f_rfh(b200atmo_ctx*, const B200AtmoCamera*, const float*, int, int, float*, uint8_t*) {
    g_calls.push_back({"render_frame_host", {}});
    return 0;
}
static int f_cfh(b200atmo_ctx*, const B200AtmoCamera*, const float*, int, int, void*, int fmt) {
    g_calls.push_back({"composite_frame_host", {fmt, 0, 0, 0}});
    return 0;
}
// default_params is a pure host function of the library (no device work): the real one is used
static const Api kFake = {f_create, f_destroy, f_last_error, b200atmo_default_params, f_set_params, f_set_variant, f_bn,
                          f_shape,  f_cube,    f_bake,       f_rf,                    f_rfc,        f_rfh, f_cfh};

static std::vector<std::string> g_log;
static Logger recording_logger() {
    return [](LogLevel l, const std::string& m) { g_log.push_back((l == LogLevel::WARNING ? "W:" : l == LogLevel::ERROR ? "E:" : "P:") + m); };
}
static bool has_prop(const std::vector<PropertyInfo>& props, const char* uniform) {
    for (const PropertyInfo& p : props)
        if (p.name == std::string("shader_params/") + uniform) return true;
    return false;
}

static void test_defaults_and_init() {
    g_calls.clear();
    PlanetAtmosphere n(0, kFake, recording_logger());
    CHECK(n.ok());
    CHECK(n.get_planet_radius() == 1.0f && n.get_atmosphere_height() == 0.1f);             // planet_atmosphere.gd:20,28
    const B200AtmoParams& p = n.get_material_params();
    CHECK(p.sun_position[0] == 5000.0f && p.sun_position[1] == 0.0f && p.sun_position[2] == 0.0f);   // :106
    CHECK(p.clip_mode == 0.0f && n.get_mode() == PlanetAtmosphere::MODE_FAR);              // :108, :58
    CHECK(n.clouds_rotation_speed == 1.0f && !n.force_fullscreen);                          // :52, :54
    CHECK(approx(n.get_extra_cull_margin(), 1.1));                                          // :241-242
    CHECK(has_variant(B200ATMO_SCATTER_V2, 8, 0, B200ATMO_LIGHT_NONE));                     // default shader, :13-14
    CHECK(n.get_shader_parameter("u_sun_position") == Variant::vector3(5000.f, 0.f, 0.f));
    CHECK(n.get_far_mesh_size() == 1.0f);                                                   // :99
    CHECK(PlanetAtmosphere::MODE_NEAR == 0 && PlanetAtmosphere::MODE_FAR == 1 && PlanetAtmosphere::SWITCH_MARGIN_RATIO == 1.1f);
}

static void test_setters_clamp_and_trigger_rebake() {
    g_calls.clear();
    PlanetAtmosphere n(0, kFake, recording_logger());
    CHECK(n.get_optical_depth_baker() == nullptr);
    CHECK(n.set_custom_shader("planet_atmosphere_no_clouds.gdshader"));   // declares u_optical_depth_texture -> baking on
    const OpticalDepthBaker* baker = n.get_optical_depth_baker();
    CHECK(baker && baker->state() == OpticalDepthBaker::STATE_REQUEST_BAKE && baker->is_processing());
    n._process(0.016, nullptr);   // frame 1: _setup_bake (optical_depth_baker.gd:66-72)
    CHECK(baker->state() == OpticalDepthBaker::STATE_PENDING_RENDER && count("bake_optical_depth") == 1);
    CHECK(!n.is_optical_depth_ready());
    n._process(0.016, nullptr);   // frame 2: `baked` (:74-85)
    CHECK(baker->state() == OpticalDepthBaker::STATE_IDLE && n.is_optical_depth_ready() && !baker->is_processing());
    n.set_planet_radius(-5.0f);                                            // maxf(new_radius, 0.0), :233
    CHECK(n.get_planet_radius() == 0.0f && baker->state() == OpticalDepthBaker::STATE_REQUEST_BAKE);
    n._process(0, nullptr), n._process(0, nullptr);
    n.set_atmosphere_height(0.3f);
    CHECK(n.get_atmosphere_height() == 0.3f && approx(n.get_extra_cull_margin(), 0.3));
    CHECK(baker->state() == OpticalDepthBaker::STATE_REQUEST_BAKE);
    n._process(0, nullptr), n._process(0, nullptr);
    const int bakes = count("bake_optical_depth");
    n.set_atmosphere_height(0.3f);                                         // unchanged: early return, :246-247
    CHECK(baker->state() == OpticalDepthBaker::STATE_IDLE);
    CHECK(n._set("shader_params/u_scattering_strength", Variant(3.0)));    // not in _shader_params_affecting_optical_depth
    CHECK(baker->state() == OpticalDepthBaker::STATE_IDLE);
    CHECK(n._set("shader_params/u_density", Variant(0.7)));                // :217-218
    CHECK(baker->state() == OpticalDepthBaker::STATE_REQUEST_BAKE && approx(n.get_material_params().density, 0.7));
    n._process(0, nullptr), n._process(0, nullptr);
    CHECK(count("bake_optical_depth") == bakes + 1);
    CHECK(approx(g_last_params.density, 0.7) && g_last_params.atmosphere_height == 0.3f);   // uniforms copied to the bake (:55-59)
    CHECK(!n._set("planet_radius", Variant(2.0)));                         // not a shader_params/ key: not handled
}

static void test_shader_params_surface() {
    g_calls.clear();
    PlanetAtmosphere n(0, kFake, recording_logger());
    std::vector<PropertyInfo> props = n._get_property_list();
    CHECK(has_prop(props, "u_density") && has_prop(props, "u_scattering_wavelengths") && has_prop(props, "u_sphere_depth_factor"));
    for (const char* hidden : {"u_planet_radius", "u_atmosphere_height", "u_clip_mode", "u_sun_position", "u_world_to_model_matrix",
                               "u_blue_noise_texture", "u_cloud_coverage_rotation", "u_optical_depth_texture"})   // :68-77
        CHECK(!has_prop(props, hidden));
    CHECK(!has_prop(props, "u_cloud_density_scale"));                      // no clouds in the default shader
    n.set_custom_shader("planet_atmosphere_clouds_high");
    props = n._get_property_list();
    CHECK(has_prop(props, "u_cloud_density_scale") && has_prop(props, "u_cloud_coverage_cubemap") && has_prop(props, "u_cloud_shape_texture"));
    CHECK(!has_prop(props, "u_cloud_coverage_rotation") && !has_prop(props, "u_world_to_model_matrix"));
    n.set_custom_shader("planet_atmosphere_v1_clouds");
    props = n._get_property_list();
    CHECK(has_prop(props, "u_day_color0") && !has_prop(props, "u_scattering_strength"));
    for (const PropertyInfo& p : props)
        if (p.name == "shader_params/u_day_color0") CHECK(p.is_color && p.type == Variant::COLOR);
    Variant v;
    CHECK(n._get("shader_params/u_cloud_top", &v) && v.type() == Variant::FLOAT && v.as_float() == 0.5f);   // shader default, :206-207
    n._set("shader_params/u_cloud_top", Variant(0.6));
    CHECK(n._get("shader_params/u_cloud_top", &v) && approx(v.as_float(), 0.6) && approx(n.get_material_params().cloud_top, 0.6));
    CHECK(!n._get("planet_radius", &v));
    // source_color uniforms are converted sRGB -> linear before upload; get returns what was set
    n.set_custom_shader("planet_atmosphere_no_clouds");
    n.set_shader_parameter("u_atmosphere_modulate", Variant::color(1.0f, 0.5f, 0.0f));
    CHECK(n.get_shader_parameter("u_atmosphere_modulate") == Variant::color(1.0f, 0.5f, 0.0f));
    const float* m = n.get_material_params().atmosphere_modulate;
    CHECK(m[0] == 1.0f && approx(m[1], 0.21404114) && m[2] == 0.0f);
    // textures go to the upload entry points
    auto shape = std::make_shared<Texture>();
    shape->kind = Texture::TEXTURE_3D, shape->width = shape->height = shape->depth = 4, shape->texels.assign(64, 128);
    auto cube = std::make_shared<Texture>();
    cube->kind = Texture::CUBEMAP, cube->width = cube->height = 2, cube->texels.assign(24, 200);
    auto bn = std::make_shared<Texture>();
    bn->width = bn->height = 256, bn->texels.assign(65536, 7);
    n.set_custom_shader("planet_atmosphere_clouds");
    n._set("shader_params/u_cloud_shape_texture", Variant::texture(shape));
    n._set("shader_params/u_cloud_coverage_cubemap", Variant::texture(cube));
    n.set_shader_parameter("u_blue_noise_texture", Variant::texture(bn));
    CHECK(count("upload_shape3d") == 1 && count("upload_coverage_cube") == 1 && count("upload_blue_noise") == 1);
    // unknown parameters are kept by the material, silently
    n.set_shader_parameter("u_not_a_uniform", Variant(1.0));
    CHECK(n.get_shader_parameter("u_not_a_uniform") == Variant(1.0) && n.get_shader_parameter("u_never_set").is_nil());
}

static void test_deprecated_accessors_warn() {
    g_calls.clear(), g_log.clear();
    PlanetAtmosphere n(0, kFake, recording_logger());
    n.set_shader_param("u_density", Variant(0.4));                        // :164-166
    CHECK(approx(n.get_shader_param("u_density").as_float(), 0.4));       // :170-172
    CHECK(g_log.size() == 2 && g_log[0] == "W:set_shader_param is deprecated, use set_shader_parameter" &&
          g_log[1] == "W:get_shader_param is deprecated, use get_shader_parameter");
}

static void test_variant_table_matches_entry_shaders() {
    ShaderVariant v;
    CHECK(find_shader_variant("planet_atmosphere_no_clouds", &v) && v == (ShaderVariant{B200ATMO_SCATTER_V2, 8, 0, B200ATMO_LIGHT_NONE}));
    CHECK(find_shader_variant("planet_atmosphere_clouds", &v) && v == (ShaderVariant{B200ATMO_SCATTER_V2, 8, 32, B200ATMO_LIGHT_CHEAP}));
    CHECK(find_shader_variant("planet_atmosphere_clouds_high", &v) && v == (ShaderVariant{B200ATMO_SCATTER_V2, 8, 64, B200ATMO_LIGHT_CHEAP}));
    CHECK(find_shader_variant("planet_atmosphere_clouds_high_rm", &v) && v == (ShaderVariant{B200ATMO_SCATTER_V2, 8, 64, B200ATMO_LIGHT_RAYMARCHED}));
    CHECK(find_shader_variant("planet_atmosphere_v1_no_clouds", &v) && v == (ShaderVariant{B200ATMO_SCATTER_V1, 16, 0, B200ATMO_LIGHT_NONE}));
    CHECK(find_shader_variant("planet_atmosphere_v1_clouds", &v) && v == (ShaderVariant{B200ATMO_SCATTER_V1, 16, 32, B200ATMO_LIGHT_CHEAP}));
    CHECK(find_shader_variant("planet_atmosphere_v1_clouds_high", &v) && v == (ShaderVariant{B200ATMO_SCATTER_V1, 16, 64, B200ATMO_LIGHT_CHEAP}));
    CHECK(shipped_shader_names().size() == 7);
    g_calls.clear(), g_log.clear();
    PlanetAtmosphere n(0, kFake, recording_logger());
    CHECK(n.set_custom_shader("res://addons/zylann.atmosphere/shaders/planet_atmosphere_clouds_high_m.gdshader"));   // README.md:35 spelling
    CHECK(has_variant(B200ATMO_SCATTER_V2, 8, 64, B200ATMO_LIGHT_RAYMARCHED));
    CHECK(n.set_custom_shader_variant({B200ATMO_SCATTER_V2, 32, 128, B200ATMO_LIGHT_RAYMARCHED}));   // BASELINE scale-up step counts
    CHECK(has_variant(B200ATMO_SCATTER_V2, 32, 128, B200ATMO_LIGHT_RAYMARCHED));
    CHECK(!n.set_custom_shader("no_such_shader") && !g_log.empty() && g_log.back().rfind("E:", 0) == 0);
    CHECK(n.set_custom_shader(""));                                        // null -> DefaultShader (:122-123)
    CHECK(n.get_variant() == (ShaderVariant{B200ATMO_SCATTER_V2, 8, 0, B200ATMO_LIGHT_NONE}));
    // v1 shaders have no u_optical_depth_texture: a fresh node on a v1 shader never bakes (:132-139)
    PlanetAtmosphere lite(0, kFake, recording_logger());
    lite.set_custom_shader("planet_atmosphere_v1_no_clouds");
    lite.set_planet_radius(3.0f);
    CHECK(lite.get_optical_depth_baker() == nullptr);
}

static void test_process_mode_switch_sun_and_rotation() {
    g_calls.clear();
    PlanetAtmosphere n(0, kFake, recording_logger());
    n.set_planet_radius(100.0f), n.set_atmosphere_height(8.0f);
    const double clip = 1.75 * (100.0 + 8.0 + double(0.1f)) * double(1.1f);
    uint64_t now_ms = 0;
    n.set_ticks_msec_source([&] { return now_ms; });
    PlanetAtmosphere::Camera cam = {{0.f, 0.f, float(clip * 1.01)}, 0.1f};
    n._process(0.0, &cam);
    CHECK(n.get_mode() == PlanetAtmosphere::MODE_FAR && n.get_material_params().clip_mode == 0.0f && approx(n.get_far_mesh_size(), clip));
    cam.position[2] = float(clip * 0.99);
    n._process(0.0, &cam);
    CHECK(n.get_mode() == PlanetAtmosphere::MODE_NEAR && n.get_material_params().clip_mode == 1.0f);   // :268-275
    cam.position[2] = 1e6f;
    n._process(0.0, &cam);
    CHECK(n.get_mode() == PlanetAtmosphere::MODE_FAR);
    n.force_fullscreen = true;                                             // :309
    n._process(0.0, &cam);
    CHECK(n.get_mode() == PlanetAtmosphere::MODE_NEAR);
    n.force_fullscreen = false;
    // no camera: editor falls back to 10 * (R + H + near) away on +X (:296-299) -> far; at run time cam_pos = 0 -> near
    n.editor_hint = true;
    n._process(0.0, nullptr);
    CHECK(n.get_mode() == PlanetAtmosphere::MODE_FAR);
    n.editor_hint = false;
    n._process(0.0, nullptr);
    CHECK(n.get_mode() == PlanetAtmosphere::MODE_NEAR);
    auto warn = n._get_configuration_warnings();
    CHECK(warn.size() == 1 && warn[0] == "The path to the sun is not assigned.");            // :222-223
    n.set_sun_path("../Sun");
    n.set_sun_resolver([](const std::string& p) {
        PlanetAtmosphere::SunLookup s;
        s.exists = p == "../Sun" || p == "../NotSpatial";
        s.is_node3d = p == "../Sun";
        s.origin[0] = 1.0f, s.origin[1] = 2.0f, s.origin[2] = 478.677f;
        return s;
    });
    CHECK(n._get_configuration_warnings().empty());
    n.set_sun_path("../NotSpatial");
    warn = n._get_configuration_warnings();
    CHECK(warn.size() == 1 && warn[0] == "The assigned sun node is not a Node3D.");          // :225-226
    n._process(0.0, &cam);
    CHECK(n.get_material_params().sun_position[0] == 5000.0f);                               // `if sun is Node3D` (:330)
    n.set_sun_path("../Sun");
    // a rotated + translated node: global_transform.inverse() = (B^T, -B^T o)
    const float T[16] = {0.f, 1.f, 0.f, 0.f, -1.f, 0.f, 0.f, 0.f, 0.f, 0.f, 1.f, 0.f, 10.f, 0.f, 0.f, 1.f};   // 90 deg about Z, origin (10,0,0)
    n.set_global_transform(T);
    n.clouds_rotation_speed = 90.0f;
    now_ms = 1000;                                                         // t = 1 s -> 90 degrees
    PlanetAtmosphere::Camera cam0 = {{0.f, 0.f, 0.f}, 0.1f};
    n._process(0.0, &cam0);
    const B200AtmoParams& p = n.get_material_params();
    CHECK(p.sun_position[0] == 1.0f && p.sun_position[1] == 2.0f && p.sun_position[2] == 478.677f);
    const float* w = p.world_to_model;   // column-major
    CHECK(w[0] == 0.f && w[1] == -1.f && w[4] == 1.f && w[5] == 0.f && w[10] == 1.f && w[15] == 1.f);
    CHECK(approx(w[12], 0.0) && approx(w[13], 10.0) && approx(w[14], 0.0));                  // -B^T * (10,0,0) = (0,10,0)
    const float* r = p.cloud_coverage_rotation;
    CHECK(approx(r[0], 0.0, 1e-6) && approx(r[1], 1.0) && approx(r[2], -1.0) && approx(r[3], 0.0, 1e-6));   // Transform2D().rotated(pi/2)
    // the draw: MODE_NEAR passes clip_box_size 0 (fullscreen quad), MODE_FAR the BoxMesh edge; set_params precedes the draw
    const float I[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    CHECK(n.get_mode() == PlanetAtmosphere::MODE_NEAR && n.make_camera(I, I, I).clip_box_size == 0.0f);
    n._process(0.0, &cam);
    const B200AtmoCamera fc = n.make_camera(I, I, I, true);
    CHECK(n.get_mode() == PlanetAtmosphere::MODE_FAR && approx(fc.clip_box_size, clip) && fc.double_precision == 1 && fc.model[12] == 10.f);
    const int sp = count("set_params");
    CHECK(n.render(fc, nullptr, 4, 4, nullptr, nullptr, nullptr) == B200ATMO_OK);
    CHECK(count("set_params") == sp + 1 && g_calls.back().name == "render_frame");
    CHECK(n.render_composite(fc, nullptr, 4, 4, nullptr, nullptr) == B200ATMO_OK && g_calls.back().name == "render_frame_composite");
    CHECK(n.render_host(fc, nullptr, 4, 4, nullptr, nullptr) == B200ATMO_OK && g_calls.back().name == "render_frame_host");
    CHECK(n.composite_host(fc, nullptr, 4, 4, nullptr, B200ATMO_COLOR_RGBA16F) == B200ATMO_OK &&
          g_calls.back().name == "composite_frame_host" && g_calls.back().a[0] == B200ATMO_COLOR_RGBA16F);
}

static void test_create_failure_is_loud() {
    Api broken = kFake;
    broken.create = [](int, b200atmo_ctx** out) { *out = nullptr; return int(B200ATMO_E_CUDA); };
    g_log.clear();
    PlanetAtmosphere n(0, broken, recording_logger());
    CHECK(!n.ok() && !n.init_error().empty() && !g_log.empty() && g_log[0].rfind("E:", 0) == 0);
    const float I[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    CHECK(n.render(n.make_camera(I, I, I), nullptr, 4, 4, nullptr, nullptr, nullptr) == B200ATMO_E_STATE);   // no CPU fallback
}

static int run_logic() {
    test_defaults_and_init();
    test_setters_clamp_and_trigger_rebake();
    test_shader_params_surface();
    test_deprecated_accessors_warn();
    test_variant_table_matches_entry_shaders();
    test_process_mode_switch_sun_and_rotation();
    test_create_failure_is_loud();
    std::printf("%d checks, %d failed\n", g_checks, g_failed);
    return g_failed ? 1 : 0;
}

// ---- GPU: one frame through the node with the real library ------------------------------------------------------
// in.bin : int32 w, h, shape_n, cube_res; B200AtmoParams demo; float inv_projection[16], inv_view[16], view[16];
//          float cam_pos[3]; float depth[w*h]; uint8 shape[n^3]; uint8 cube[6*res^2]; uint8 blue_noise[256*256]
// out.bin: float rgba[w*h*4]; uint8 discard[w*h]; B200AtmoParams as uploaded
template <class T> static bool rd(std::ifstream& f, T* p, size_t n) { return bool(f.read(reinterpret_cast<char*>(p), std::streamsize(sizeof(T) * n))); }

static int run_render(const char* in_path, const char* out_path) {
    std::ifstream f(in_path, std::ios::binary);
    int32_t hdr[4];
    B200AtmoParams demo;
    float inv_p[16], inv_v[16], view[16], cam_pos[3];
    if (!f || !rd(f, hdr, 4) || !rd(f, &demo, 1) || !rd(f, inv_p, 16) || !rd(f, inv_v, 16) || !rd(f, view, 16) || !rd(f, cam_pos, 3)) return 2;
    const int w = hdr[0], h = hdr[1], sn = hdr[2], cr = hdr[3];
    std::vector<float> depth(size_t(w) * h);
    auto shape = std::make_shared<Texture>(), cube = std::make_shared<Texture>(), bn = std::make_shared<Texture>();
    shape->kind = Texture::TEXTURE_3D, shape->width = shape->height = shape->depth = sn, shape->texels.resize(size_t(sn) * sn * sn);
    cube->kind = Texture::CUBEMAP, cube->width = cube->height = cr, cube->texels.resize(size_t(6) * cr * cr);
    bn->width = bn->height = 256, bn->texels.resize(65536);
    if (!rd(f, depth.data(), depth.size()) || !rd(f, shape->texels.data(), shape->texels.size()) ||
        !rd(f, cube->texels.data(), cube->texels.size()) || !rd(f, bn->texels.data(), bn->texels.size()))
        return 2;

    PlanetAtmosphere n(0);   // linked_api(): the real libb200atmo.so; fails loudly without a GPU
    if (!n.ok()) {
        std::fprintf(stderr, "create failed: %s\n", n.init_error().c_str());
        return 3;
    }
    n.set_planet_radius(demo.planet_radius), n.set_atmosphere_height(demo.atmosphere_height);
    n._ready();
    n.set_shader_parameter("u_blue_noise_texture", Variant::texture(bn));
    n.set_custom_shader("planet_atmosphere_clouds");
    n._set("shader_params/u_density", Variant(demo.density));
    n._set("shader_params/u_scattering_strength", Variant(demo.scattering_strength));
    n._set("shader_params/u_cloud_density_scale", Variant(demo.cloud_density_scale));
    n._set("shader_params/u_cloud_top", Variant(demo.cloud_top));
    n._set("shader_params/u_cloud_shape_invert", Variant(demo.cloud_shape_invert));
    n._set("shader_params/u_cloud_shape_factor", Variant(demo.cloud_shape_factor));
    n._set("shader_params/u_cloud_shape_scale", Variant(demo.cloud_shape_scale));
    n._set("shader_params/u_cloud_shape_texture", Variant::texture(shape));
    n._set("shader_params/u_cloud_coverage_cubemap", Variant::texture(cube));
    n.set_sun_path("Sun");
    n.set_sun_resolver([&](const std::string&) {
        PlanetAtmosphere::SunLookup s;
        s.exists = s.is_node3d = true;
        for (int k = 0; k < 3; ++k) s.origin[k] = demo.sun_position[k];
        return s;
    });
    n.set_ticks_msec_source([] { return uint64_t(0); });
    PlanetAtmosphere::Camera cam = {{cam_pos[0], cam_pos[1], cam_pos[2]}, 0.1f};
    for (int frame = 0; frame < 2; ++frame)
        if (n._process(0.016, &cam) != B200ATMO_OK) return 4;
    if (!n.is_optical_depth_ready()) return 5;
    const B200AtmoCamera c = n.make_camera(inv_p, inv_v, view);
    std::vector<float> rgba(size_t(w) * h * 4);
    std::vector<uint8_t> disc(size_t(w) * h);
    if (n.render_host(c, depth.data(), w, h, rgba.data(), disc.data()) != B200ATMO_OK) {
        std::fprintf(stderr, "render failed: %s\n", n.last_error().c_str());
        return 6;
    }
    std::ofstream o(out_path, std::ios::binary);
    o.write(reinterpret_cast<const char*>(rgba.data()), std::streamsize(rgba.size() * 4));
    o.write(reinterpret_cast<const char*>(disc.data()), std::streamsize(disc.size()));
    o.write(reinterpret_cast<const char*>(&n.get_material_params()), sizeof(B200AtmoParams));
    std::printf("rendered %dx%d through the C++ node, mode %d\n", w, h, n.get_mode());
    return o ? 0 : 7;
}

int main(int argc, char** argv) {
    if (argc >= 2 && std::strcmp(argv[1], "logic") == 0) return run_logic();
    if (argc >= 4 && std::strcmp(argv[1], "render") == 0) return run_render(argv[2], argv[3]);
    std::fprintf(stderr, "usage: %s logic | render <in.bin> <out.bin>\n", argv[0]);
    return 64;
}

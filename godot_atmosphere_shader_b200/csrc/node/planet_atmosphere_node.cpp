// See planet_atmosphere_node.hpp. Host logic only; the arithmetic of the hot path is behind the C-ABI.
#include "planet_atmosphere_node.hpp"

#include <chrono>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstring>

namespace b200atmo {

const Api& linked_api() {
    static const Api api = {
        b200atmo_create,        b200atmo_destroy,        b200atmo_last_error,          b200atmo_default_params,
        b200atmo_set_params,    b200atmo_set_variant,    b200atmo_upload_blue_noise,   b200atmo_upload_shape3d,
        b200atmo_upload_coverage_cube, b200atmo_bake_optical_depth, b200atmo_render_frame, b200atmo_render_frame_composite,
        b200atmo_render_frame_host, b200atmo_composite_frame_host,
    };
    return api;
}

// ---------------------------------------------------------------------------------------------------------------
// Variant
// ---------------------------------------------------------------------------------------------------------------
Variant Variant::vector3(float x, float y, float z) {
    Variant v;
    v.type_ = VECTOR3;
    v.v_[0] = x, v.v_[1] = y, v.v_[2] = z;
    return v;
}
Variant Variant::color(float r, float g, float b, float a) {
    Variant v;
    v.type_ = COLOR;
    v.v_[0] = r, v.v_[1] = g, v.v_[2] = b, v.v_[3] = a;
    return v;
}
Variant Variant::transform2d(float c0x, float c0y, float c1x, float c1y) {
    Variant v;
    v.type_ = TRANSFORM2D;
    v.v_[0] = c0x, v.v_[1] = c0y, v.v_[2] = c1x, v.v_[3] = c1y;
    return v;
}
Variant Variant::transform3d(const float m[16]) {
    Variant v;
    v.type_ = TRANSFORM3D;
    for (int i = 0; i < 16; ++i) v.v_[i] = m[i];
    return v;
}
Variant Variant::texture(std::shared_ptr<const Texture> t) {
    Variant v;
    v.type_ = TEXTURE;
    v.tex_ = std::move(t);
    return v;
}
bool Variant::operator==(const Variant& o) const { return type_ == o.type_ && v_ == o.v_ && tex_ == o.tex_; }

float srgb_to_linear(float c) {
    // Color::srgb_to_linear: c < 0.04045 ? c / 12.92 : pow((c + 0.055) / 1.055, 2.4)
    const double x = c;
    return float(x < 0.04045 ? x / 12.92 : std::pow((x + 0.055) / 1.055, 2.4));
}

// ---------------------------------------------------------------------------------------------------------------
// shader registry: #defines of the entry shaders + the uniforms each one declares, in declaration order
// ---------------------------------------------------------------------------------------------------------------
namespace {

struct NamedVariant {
    const char* name;
    ShaderVariant v;
};
const NamedVariant kShaders[] = {
    // shaders/planet_atmosphere_*.gdshader:4-7 — ATMOSPHERE_RAYMARCH_STEPS, CLOUDS_MAX_RAYMARCH_STEPS, lighting
    {"planet_atmosphere_no_clouds", {B200ATMO_SCATTER_V2, 8, 0, B200ATMO_LIGHT_NONE}},
    {"planet_atmosphere_clouds", {B200ATMO_SCATTER_V2, 8, 32, B200ATMO_LIGHT_CHEAP}},
    {"planet_atmosphere_clouds_high", {B200ATMO_SCATTER_V2, 8, 64, B200ATMO_LIGHT_CHEAP}},
    {"planet_atmosphere_clouds_high_rm", {B200ATMO_SCATTER_V2, 8, 64, B200ATMO_LIGHT_RAYMARCHED}},
    // shaders/planet_atmosphere_v1_*.gdshader:4-7 — ATMOSPHERE_LITE, 16 steps
    {"planet_atmosphere_v1_no_clouds", {B200ATMO_SCATTER_V1, 16, 0, B200ATMO_LIGHT_NONE}},
    {"planet_atmosphere_v1_clouds", {B200ATMO_SCATTER_V1, 16, 32, B200ATMO_LIGHT_CHEAP}},
    {"planet_atmosphere_v1_clouds_high", {B200ATMO_SCATTER_V1, 16, 64, B200ATMO_LIGHT_CHEAP}},
};
const char* const kDefaultShader = "planet_atmosphere_no_clouds";   // planet_atmosphere.gd:13-14

enum class Kind { FLOAT, BOOL, VEC3, COLOR3, COLOR4, MAT2, MAT4, TEX2D, TEX3D, CUBE };
enum Group { COMMON = 1, V1 = 2, V2 = 4, CLOUDS = 8, MAIN = 16 };

struct Uniform {
    const char* name;
    Kind kind;
    int group;
    std::ptrdiff_t offset;   // into B200AtmoParams, -1 = not a POD field (samplers)
    bool api;                // planet_atmosphere.gd:68-77 (_api_shader_params): assigned internally, hidden
};
#define P_OFF(f) std::ptrdiff_t(offsetof(B200AtmoParams, f))
// Declaration order as the preprocessor sees it: planet_common, atmosphere_common, funcs_v1 | funcs_v2, cloud_funcs, main.
const Uniform kUniforms[] = {
    {"u_planet_radius", Kind::FLOAT, COMMON, P_OFF(planet_radius), true},                       // planet_common:4
    {"u_atmosphere_height", Kind::FLOAT, COMMON, P_OFF(atmosphere_height), true},               // planet_common:5
    {"u_sun_position", Kind::VEC3, COMMON, P_OFF(sun_position), true},                          // planet_common:6
    {"u_density", Kind::FLOAT, COMMON, P_OFF(density), false},                                  // atmosphere_common:10
    {"u_day_color0", Kind::COLOR4, V1, P_OFF(day_color0), false},                               // funcs_v1:8-12
    {"u_day_color1", Kind::COLOR4, V1, P_OFF(day_color1), false},
    {"u_night_color0", Kind::COLOR4, V1, P_OFF(night_color0), false},
    {"u_night_color1", Kind::COLOR4, V1, P_OFF(night_color1), false},
    {"u_day_night_transition_scale", Kind::FLOAT, V1, P_OFF(day_night_transition_scale), false},
    {"u_optical_depth_texture", Kind::TEX2D, V2, -1, true},                                     // funcs_v2:7-11
    {"u_scattering_strength", Kind::FLOAT, V2, P_OFF(scattering_strength), false},
    {"u_scattering_wavelengths", Kind::VEC3, V2, P_OFF(scattering_wavelengths), false},
    {"u_atmosphere_modulate", Kind::COLOR3, V2, P_OFF(atmosphere_modulate), false},
    {"u_atmosphere_ambient_color", Kind::COLOR3, V2, P_OFF(atmosphere_ambient_color), false},
    {"u_cloud_density_scale", Kind::FLOAT, CLOUDS, P_OFF(cloud_density_scale), false},           // cloud_funcs:5-16
    {"u_cloud_bottom", Kind::FLOAT, CLOUDS, P_OFF(cloud_bottom), false},
    {"u_cloud_top", Kind::FLOAT, CLOUDS, P_OFF(cloud_top), false},
    {"u_cloud_blend", Kind::FLOAT, CLOUDS, P_OFF(cloud_blend), false},
    {"u_world_to_model_matrix", Kind::MAT4, CLOUDS, P_OFF(world_to_model), true},
    {"u_cloud_shape_texture", Kind::TEX3D, CLOUDS, -1, false},
    {"u_cloud_shape_invert", Kind::FLOAT, CLOUDS, P_OFF(cloud_shape_invert), false},
    {"u_cloud_coverage_bias", Kind::FLOAT, CLOUDS, P_OFF(cloud_coverage_bias), false},
    {"u_cloud_shape_factor", Kind::FLOAT, CLOUDS, P_OFF(cloud_shape_factor), false},
    {"u_cloud_shape_scale", Kind::FLOAT, CLOUDS, P_OFF(cloud_shape_scale), false},
    {"u_cloud_coverage_cubemap", Kind::CUBE, CLOUDS, -1, false},
    {"u_cloud_coverage_rotation", Kind::MAT2, CLOUDS, P_OFF(cloud_coverage_rotation), true},
    {"u_clip_mode", Kind::BOOL, MAIN, P_OFF(clip_mode), true},                                  // main:55
    {"u_sphere_depth_factor", Kind::FLOAT, MAIN, P_OFF(sphere_depth_factor), false},            // main:60
    {"u_blue_noise_texture", Kind::TEX2D, MAIN, -1, true},                                      // main:63
};
#undef P_OFF

int groups_of(const ShaderVariant& v) {
    int g = COMMON | MAIN | (v.scatter_model == B200ATMO_SCATTER_V1 ? V1 : V2);
    if (v.light_mode != B200ATMO_LIGHT_NONE) g |= CLOUDS;
    return g;
}
const Uniform* find_uniform(const std::string& name) {
    for (const Uniform& u : kUniforms)
        if (name == u.name) return &u;
    return nullptr;
}
bool shader_has(const ShaderVariant& v, const Uniform& u) { return (groups_of(v) & u.group) != 0; }

Variant::Type variant_type(Kind k) {
    switch (k) {
        case Kind::FLOAT: case Kind::BOOL: return Variant::FLOAT;
        case Kind::VEC3: return Variant::VECTOR3;
        case Kind::COLOR3: case Kind::COLOR4: return Variant::COLOR;
        case Kind::MAT2: return Variant::TRANSFORM2D;
        case Kind::MAT4: return Variant::TRANSFORM3D;
        default: return Variant::TEXTURE;
    }
}

const char kShaderParamsPrefix[] = "shader_params/";
const size_t kPrefixLen = sizeof(kShaderParamsPrefix) - 1;

const std::array<float, 16> kIdentity = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};

}  // namespace

const std::vector<std::string>& shipped_shader_names() {
    static const std::vector<std::string> names = [] {
        std::vector<std::string> n;
        for (const NamedVariant& s : kShaders) n.push_back(s.name);
        return n;
    }();
    return names;
}

bool find_shader_variant(const std::string& shader, ShaderVariant* out, std::string* canonical_name) {
    std::string stem = shader;
    const size_t slash = stem.find_last_of('/');
    if (slash != std::string::npos) stem = stem.substr(slash + 1);
    const std::string ext = ".gdshader";
    if (stem.size() > ext.size() && stem.compare(stem.size() - ext.size(), ext.size(), ext) == 0)
        stem.resize(stem.size() - ext.size());
    if (stem == "planet_atmosphere_clouds_high_m") stem = "planet_atmosphere_clouds_high_rm";   // README.md:35
    for (const NamedVariant& s : kShaders) {
        if (stem == s.name) {
            if (out) *out = s.v;
            if (canonical_name) *canonical_name = s.name;
            return true;
        }
    }
    return false;
}

// ---------------------------------------------------------------------------------------------------------------
// OpticalDepthBaker
// ---------------------------------------------------------------------------------------------------------------
void OpticalDepthBaker::request_bake(const B200AtmoParams& atmosphere_material) {
    // Not baking right now: _process must run twice, in order (optical_depth_baker.gd:41-46)
    state_ = STATE_REQUEST_BAKE;
    material_ = atmosphere_material;
    processing_ = true;
}

int OpticalDepthBaker::_process(double) {
    if (state_ == STATE_REQUEST_BAKE) {
        int rc = api_.set_params(ctx_, &material_);                       // _setup_bake (:49-64)
        if (rc == B200ATMO_OK) rc = api_.bake_optical_depth(ctx_, nullptr);
        state_ = STATE_PENDING_RENDER;
        return rc;
    }
    if (state_ == STATE_PENDING_RENDER) {                                 // :74-85
        for (auto& fn : baked_) fn();
        state_ = STATE_IDLE;
        processing_ = false;
    }
    return B200ATMO_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// PlanetAtmosphere
// ---------------------------------------------------------------------------------------------------------------
PlanetAtmosphere::PlanetAtmosphere(int cuda_device, const Api& api, Logger logger)
    : api_(api), logger_(std::move(logger)), global_transform_(kIdentity) {
    const auto t0 = std::chrono::steady_clock::now();
    ticks_msec_ = [t0] {
        return uint64_t(std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now() - t0).count());
    };
    api_.default_params(&params_);
    find_shader_variant(kDefaultShader, &variant_, &shader_name_);
    if (api_.create(cuda_device, &ctx_) != B200ATMO_OK) {
        const char* e = api_.last_error(nullptr);
        init_error_ = e ? e : "b200atmo_create failed";
        ctx_ = nullptr;
        log(LogLevel::ERROR, init_error_);
        return;
    }
    update_cull_margin();
    // defaults for the builtin shader (:105-108); the blue-noise texture arrives through
    // set_shader_parameter("u_blue_noise_texture", ...) — the addon's blue_noise.png is engine-side content
    params_.sun_position[0] = 5000.0f, params_.sun_position[1] = 0.0f, params_.sun_position[2] = 0.0f;
    params_.clip_mode = 0.0f;
    raw_params_["u_sun_position"] = Variant::vector3(5000.0f, 0.0f, 0.0f);
    raw_params_["u_clip_mode"] = Variant(0.0);
    apply_variant();
}

PlanetAtmosphere::~PlanetAtmosphere() {
    if (ctx_) api_.destroy(ctx_);
}

void PlanetAtmosphere::log(LogLevel l, const std::string& m) const {
    if (logger_) {
        logger_(l, m);
        return;
    }
    std::fprintf(stderr, "%s%s\n", l == LogLevel::ERROR ? "ERROR: " : (l == LogLevel::WARNING ? "WARNING: " : ""), m.c_str());
}

std::string PlanetAtmosphere::last_error() const {
    const char* e = api_.last_error(ctx_);
    return e ? e : "";
}

void PlanetAtmosphere::apply_variant() {
    if (!ctx_) return;
    if (api_.set_variant(ctx_, variant_.scatter_model, variant_.scatter_steps, variant_.cloud_steps, variant_.light_mode) != B200ATMO_OK)
        log(LogLevel::ERROR, last_error());
}

void PlanetAtmosphere::_ready() {
    // assigned here because the scene loader sets them after _init (:113-115)
    params_.planet_radius = planet_radius_;
    params_.atmosphere_height = atmosphere_height_;
    raw_params_["u_planet_radius"] = Variant(planet_radius_);
    raw_params_["u_atmosphere_height"] = Variant(atmosphere_height_);
}

bool PlanetAtmosphere::set_custom_shader(const std::string& shader) {
    ShaderVariant v;
    std::string name;
    if (shader.empty()) {
        find_shader_variant(kDefaultShader, &v, &name);          // mat.shader = DefaultShader (:122-123)
    } else if (!find_shader_variant(shader, &v, &name)) {
        log(LogLevel::ERROR, "unknown atmosphere shader '" + shader + "'");
        return false;
    }
    custom_shader_ = shader;
    shader_name_ = name;
    return set_custom_shader_variant(v);
}

bool PlanetAtmosphere::set_custom_shader_variant(const ShaderVariant& v) {
    variant_ = v;
    apply_variant();
    // the LUT is baked for shaders that declare u_optical_depth_texture = every v2 variant; the flag is never
    // cleared again (:132-139)
    if (v.scatter_model == B200ATMO_SCATTER_V2) uses_baked_optical_depth_ = true;
    if (uses_baked_optical_depth_) request_bake_optical_depth();
    return true;   // notify_property_list_changed() is the wrapper's business
}

void PlanetAtmosphere::request_bake_optical_depth() {
    if (!ctx_) return;
    if (!baker_) {
        baker_.reset(new OpticalDepthBaker(api_, ctx_));
        baker_->connect_baked([this] { on_optical_depth_baked(); });
    }
    optical_depth_ready_ = false;
    baker_->request_bake(params_);
}

void PlanetAtmosphere::on_optical_depth_baked() { optical_depth_ready_ = true; }

void PlanetAtmosphere::set_shader_param(const std::string& name, const Variant& value) {
    log(LogLevel::WARNING, "set_shader_param is deprecated, use set_shader_parameter");
    set_shader_parameter(name, value);
}

Variant PlanetAtmosphere::get_shader_param(const std::string& name) {
    log(LogLevel::WARNING, "get_shader_param is deprecated, use get_shader_parameter");
    return get_shader_parameter(name);
}

void PlanetAtmosphere::set_shader_parameter(const std::string& name, const Variant& value) {
    raw_params_[name] = value;   // ShaderMaterial keeps any parameter, known to the shader or not
    const Uniform* u = find_uniform(name);
    if (!u) return;
    const std::array<float, 16>& v = value.values();
    float* dst = u->offset >= 0 ? reinterpret_cast<float*>(reinterpret_cast<char*>(&params_) + u->offset) : nullptr;
    switch (u->kind) {
        case Kind::FLOAT:
        case Kind::BOOL:
            if (value.type() == Variant::FLOAT) dst[0] = v[0];
            break;
        case Kind::VEC3:
            if (value.type() == Variant::VECTOR3 || value.type() == Variant::COLOR)
                for (int k = 0; k < 3; ++k) dst[k] = v[k];
            break;
        case Kind::COLOR3:   // source_color: sRGB -> linear before upload
            if (value.type() == Variant::COLOR || value.type() == Variant::VECTOR3)
                for (int k = 0; k < 3; ++k) dst[k] = srgb_to_linear(v[k]);
            break;
        case Kind::COLOR4:
            if (value.type() == Variant::COLOR) {
                for (int k = 0; k < 3; ++k) dst[k] = srgb_to_linear(v[k]);
                dst[3] = v[3];
            }
            break;
        case Kind::MAT2:
            if (value.type() == Variant::TRANSFORM2D)
                for (int k = 0; k < 4; ++k) dst[k] = v[k];
            break;
        case Kind::MAT4:
            if (value.type() == Variant::TRANSFORM3D)
                for (int k = 0; k < 16; ++k) dst[k] = v[k];
            break;
        case Kind::TEX2D:
        case Kind::TEX3D:
        case Kind::CUBE: {
            if (name == "u_optical_depth_texture") break;   // owned by the context (baker output)
            if (!ctx_ || value.type() != Variant::TEXTURE || !value.tex()) break;
            const Texture& t = *value.tex();
            int rc = B200ATMO_OK;
            if (u->kind == Kind::TEX2D) rc = api_.upload_blue_noise(ctx_, t.texels.data(), t.width, t.height);
            else if (u->kind == Kind::TEX3D) rc = api_.upload_shape3d(ctx_, t.texels.data(), t.width, t.height, t.depth);
            else rc = api_.upload_coverage_cube(ctx_, t.texels.data(), t.width);
            if (rc != B200ATMO_OK) log(LogLevel::ERROR, last_error());
            break;
        }
    }
}

Variant PlanetAtmosphere::get_shader_parameter(const std::string& name) const {
    const auto it = raw_params_.find(name);
    return it == raw_params_.end() ? Variant() : it->second;
}

std::vector<PropertyInfo> PlanetAtmosphere::_get_property_list() const {
    std::vector<PropertyInfo> props;
    for (const Uniform& u : kUniforms) {
        if (!shader_has(variant_, u) || u.api) continue;   // :190-191
        props.push_back(PropertyInfo{std::string(kShaderParamsPrefix) + u.name, variant_type(u.kind),
                                     u.kind == Kind::COLOR3 || u.kind == Kind::COLOR4});
    }
    return props;
}

static Variant shader_default(const Uniform& u, const Api& api) {
    // RenderingServer.shader_get_parameter_default: the value written in the shader source. `source_color`
    // defaults are stored as authored; the colour defaults of this addon are only used through the POD block.
    if (u.offset < 0) return Variant();
    B200AtmoParams d;
    api.default_params(&d);
    const float* src = reinterpret_cast<const float*>(reinterpret_cast<const char*>(&d) + u.offset);
    switch (u.kind) {
        case Kind::FLOAT: case Kind::BOOL: return Variant(src[0]);
        case Kind::VEC3: return Variant::vector3(src[0], src[1], src[2]);
        case Kind::COLOR3: return Variant::color(src[0], src[1], src[2], 1.0f);
        case Kind::COLOR4: return Variant::color(src[0], src[1], src[2], src[3]);
        case Kind::MAT2: return Variant::transform2d(src[0], src[1], src[2], src[3]);
        case Kind::MAT4: return Variant::transform3d(src);
        default: return Variant();
    }
}

bool PlanetAtmosphere::_get(const std::string& key, Variant* out) const {
    if (key.compare(0, kPrefixLen, kShaderParamsPrefix) != 0) return false;
    const std::string param_name = key.substr(kPrefixLen);
    Variant value = get_shader_parameter(param_name);
    if (value.is_nil()) {                                   // :206-207
        const Uniform* u = find_uniform(param_name);
        if (u && shader_has(variant_, *u)) value = shader_default(*u, api_);
    }
    if (out) *out = value;
    return true;
}

bool PlanetAtmosphere::_set(const std::string& key, const Variant& value) {
    if (key.compare(0, kPrefixLen, kShaderParamsPrefix) != 0) return false;
    const std::string param_name = key.substr(kPrefixLen);
    set_shader_parameter(param_name, value);
    // _shader_params_affecting_optical_depth (:79-81, :217-218)
    if (uses_baked_optical_depth_ && param_name == "u_density") request_bake_optical_depth();
    return true;
}

std::vector<std::string> PlanetAtmosphere::_get_configuration_warnings() const {
    if (sun_path_.empty()) return {"The path to the sun is not assigned."};
    SunLookup s;
    if (sun_resolver_) s = sun_resolver_(sun_path_);
    if (!(s.exists && s.is_node3d)) return {"The assigned sun node is not a Node3D."};
    return {};
}

void PlanetAtmosphere::set_planet_radius(float new_radius) {
    if (planet_radius_ == new_radius) return;
    planet_radius_ = std::fmax(new_radius, 0.0f);
    params_.planet_radius = planet_radius_;
    raw_params_["u_planet_radius"] = Variant(planet_radius_);
    update_cull_margin();
    if (uses_baked_optical_depth_) request_bake_optical_depth();
}

void PlanetAtmosphere::update_cull_margin() { extra_cull_margin_ = planet_radius_ + atmosphere_height_; }

void PlanetAtmosphere::set_atmosphere_height(float new_height) {
    if (atmosphere_height_ == new_height) return;
    atmosphere_height_ = std::fmax(new_height, 0.0f);
    params_.atmosphere_height = atmosphere_height_;
    raw_params_["u_atmosphere_height"] = Variant(atmosphere_height_);
    update_cull_margin();
    if (uses_baked_optical_depth_) request_bake_optical_depth();
}

void PlanetAtmosphere::set_sun_path(const std::string& new_sun_path) { sun_path_ = new_sun_path; }

void PlanetAtmosphere::set_global_transform(const float m[16]) {
    for (int i = 0; i < 16; ++i) global_transform_[i] = m[i];
}

void PlanetAtmosphere::set_mode(int mode) {
    if (mode == mode_) return;
    mode_ = mode;
    if (mode_ == MODE_NEAR) {
        if (stdout_verbose) log(LogLevel::PRINT, "Switching PlanetAtmosphere to near mode");
        // fullscreen quad at the near plane (:268-275)
        params_.clip_mode = 1.0f;
    } else {
        if (stdout_verbose) log(LogLevel::PRINT, "Switching PlanetAtmosphere to far mode");
        params_.clip_mode = 0.0f;
    }
    raw_params_["u_clip_mode"] = Variant(params_.clip_mode);
}

int PlanetAtmosphere::_process(double delta, const Camera* cam) {
    double cam_pos[3] = {0.0, 0.0, 0.0};
    double cam_near = 0.1;
    const double origin[3] = {global_transform_[12], global_transform_[13], global_transform_[14]};
    if (cam) {
        for (int k = 0; k < 3; ++k) cam_pos[k] = cam->position[k];
        cam_near = cam->near;
    } else if (editor_hint) {   // :296-299
        cam_pos[0] = origin[0] + 10.0 * (double(planet_radius_) + double(atmosphere_height_) + cam_near);
        cam_pos[1] = origin[1];
        cam_pos[2] = origin[2];
    }
    // 1.75 ~ sqrt(3): the far mesh is a cube, its largest distance from the centre counts (:301-304)
    const double atmo_clip_distance = 1.75 * (double(planet_radius_) + double(atmosphere_height_) + cam_near) * double(SWITCH_MARGIN_RATIO);
    const double dx = origin[0] - cam_pos[0], dy = origin[1] - cam_pos[1], dz = origin[2] - cam_pos[2];
    const double d = std::sqrt(dx * dx + dy * dy + dz * dz);
    const bool is_near = d < atmo_clip_distance;
    set_mode((is_near || force_fullscreen) ? MODE_NEAR : MODE_FAR);
    if (mode_ == MODE_FAR && double(prev_atmo_clip_distance_) != atmo_clip_distance) {
        // the mesh instance is never scaled: a new BoxMesh of that edge replaces the old one (:314-321)
        prev_atmo_clip_distance_ = float(atmo_clip_distance);
        far_mesh_size_ = float(atmo_clip_distance);
    }
    if (sun_resolver_ && !sun_path_.empty()) {   // :328-331
        const SunLookup s = sun_resolver_(sun_path_);
        if (s.exists && s.is_node3d) {
            for (int k = 0; k < 3; ++k) params_.sun_position[k] = s.origin[k];
            raw_params_["u_sun_position"] = Variant::vector3(s.origin[0], s.origin[1], s.origin[2]);
        }
    }
    // global_transform.inverse() (:335-336): Transform3D::inverse() = transposed basis and -B^T * origin (it assumes
    // an orthonormal basis, as the node is never scaled, :315)
    const std::array<float, 16>& g = global_transform_;
    float w2m[16] = {g[0], g[4], g[8], 0.0f, g[1], g[5], g[9], 0.0f, g[2], g[6], g[10], 0.0f, 0.0f, 0.0f, 0.0f, 1.0f};
    for (int r = 0; r < 3; ++r)
        w2m[12 + r] = float(-(double(w2m[0 + r]) * g[12] + double(w2m[4 + r]) * g[13] + double(w2m[8 + r]) * g[14]));
    std::memcpy(params_.world_to_model, w2m, sizeof(w2m));
    raw_params_["u_world_to_model_matrix"] = Variant::transform3d(w2m);
    // Transform2D().rotated(time * deg_to_rad(speed)) (:339-341): columns (cos, sin), (-sin, cos)
    const double time = double(ticks_msec_()) / 1000.0;
    const double a = time * (double(clouds_rotation_speed) * (3.14159265358979323846 / 180.0));
    const float c = float(std::cos(a)), s = float(std::sin(a));
    params_.cloud_coverage_rotation[0] = c, params_.cloud_coverage_rotation[1] = s;
    params_.cloud_coverage_rotation[2] = -s, params_.cloud_coverage_rotation[3] = c;
    raw_params_["u_cloud_coverage_rotation"] = Variant::transform2d(c, s, -s, c);
    // the baker is a child node: its _process runs in the same frame
    if (baker_ && baker_->is_processing()) {
        const int rc = baker_->_process(delta);
        if (rc != B200ATMO_OK) {
            log(LogLevel::ERROR, last_error());
            return rc;
        }
    }
    return B200ATMO_OK;
}

B200AtmoCamera PlanetAtmosphere::make_camera(const float inv_projection[16], const float inv_view[16], const float view[16],
                                             bool double_precision) const {
    B200AtmoCamera cam;
    std::memcpy(cam.inv_projection, inv_projection, sizeof(cam.inv_projection));
    std::memcpy(cam.inv_view, inv_view, sizeof(cam.inv_view));
    std::memcpy(cam.view, view, sizeof(cam.view));
    std::memcpy(cam.model, global_transform_.data(), sizeof(cam.model));
    cam.double_precision = double_precision ? 1 : 0;
    cam.clip_box_size = mode_ == MODE_FAR ? far_mesh_size_ : 0.0f;
    return cam;
}

int PlanetAtmosphere::push_params() {
    if (!ctx_) return B200ATMO_E_STATE;
    return api_.set_params(ctx_, &params_);
}

int PlanetAtmosphere::render(const B200AtmoCamera& cam, const float* d_depth, int w, int h, float* d_rgba, uint8_t* d_discard,
                             void* stream) {
    int rc = push_params();
    if (rc == B200ATMO_OK) rc = api_.render_frame(ctx_, &cam, d_depth, w, h, 0, h, d_rgba, d_discard, stream);
    return rc;
}

int PlanetAtmosphere::render_composite(const B200AtmoCamera& cam, const float* d_depth, int w, int h, float* d_color_inout,
                                       void* stream) {
    int rc = push_params();
    if (rc == B200ATMO_OK) rc = api_.render_frame_composite(ctx_, &cam, d_depth, w, h, 0, h, d_color_inout, stream);
    return rc;
}

int PlanetAtmosphere::render_host(const B200AtmoCamera& cam, const float* h_depth, int w, int h, float* h_rgba, uint8_t* h_discard) {
    int rc = push_params();
    if (rc == B200ATMO_OK) rc = api_.render_frame_host(ctx_, &cam, h_depth, w, h, h_rgba, h_discard);
    return rc;
}

int PlanetAtmosphere::composite_host(const B200AtmoCamera& cam, const float* h_depth, int w, int h, void* h_color_inout,
                                     int color_format) {
    int rc = push_params();
    if (rc == B200ATMO_OK) rc = api_.composite_frame_host(ctx_, &cam, h_depth, w, h, h_color_inout, color_format);
    return rc;
}

}  // namespace b200atmo

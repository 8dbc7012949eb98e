// Engine-independent C++ core of the drop-in `PlanetAtmosphere` node and its `OpticalDepthBaker`, above the C-ABI
// of include/b200atmo.h. A GDExtension class (gdextension/planet_atmosphere_b200.cpp) forwards its bound methods to
// this class one to one; everything Godot did implicitly (scene tree, camera, Time, push_warning) is injected.
//
// Reference surface being mirrored (addons/zylann.atmosphere/...):
//   planet_atmosphere.gd:9-11     MODE_NEAR / MODE_FAR / SWITCH_MARGIN_RATIO
//   planet_atmosphere.gd:20-54    exported properties planet_radius, atmosphere_height, sun_path, custom_shader,
//                                 clouds_rotation_speed, force_fullscreen
//   planet_atmosphere.gd:84-115   _init defaults, _ready
//   planet_atmosphere.gd:118-156  set_custom_shader, _request_bake_optical_depth, _on_optical_depth_baked
//   planet_atmosphere.gd:164-218  set/get_shader_param(eter), _get_property_list, _get, _set ("shader_params/*")
//   planet_atmosphere.gd:221-282  configuration warnings, setters, _set_mode
//   planet_atmosphere.gd:285-341  _process (mode switch, sun, u_world_to_model_matrix, u_cloud_coverage_rotation)
//   optical_depth_baker.gd:3-85   STATE_IDLE / STATE_REQUEST_BAKE / STATE_PENDING_RENDER, request_bake, _process, `baked`
//
// No arithmetic of the hot path lives here: the class fills B200AtmoParams / B200AtmoCamera and calls the library.
#pragma once

#include <array>
#include <cstdint>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "b200atmo.h"

namespace b200atmo {

// The entry points of include/b200atmo.h this class uses, as a table: linked_api() binds the library this file is
// linked against; a GDExtension that dlopen()s libb200atmo.so fills it with dlsym(); host-logic tests record calls.
struct Api {
    int (*create)(int, b200atmo_ctx**);
    void (*destroy)(b200atmo_ctx*);
    const char* (*last_error)(const b200atmo_ctx*);
    void (*default_params)(B200AtmoParams*);
    int (*set_params)(b200atmo_ctx*, const B200AtmoParams*);
    int (*set_variant)(b200atmo_ctx*, int, int, int, int);
    int (*upload_blue_noise)(b200atmo_ctx*, const uint8_t*, int, int);
    int (*upload_shape3d)(b200atmo_ctx*, const uint8_t*, int, int, int);
    int (*upload_coverage_cube)(b200atmo_ctx*, const uint8_t*, int);
    int (*bake_optical_depth)(b200atmo_ctx*, void*);
    int (*render_frame)(b200atmo_ctx*, const B200AtmoCamera*, const float*, int, int, int, int, float*, uint8_t*, void*);
    int (*render_frame_composite)(b200atmo_ctx*, const B200AtmoCamera*, const float*, int, int, int, int, float*, void*);
    int (*render_frame_host)(b200atmo_ctx*, const B200AtmoCamera*, const float*, int, int, float*, uint8_t*);
    int (*composite_frame_host)(b200atmo_ctx*, const B200AtmoCamera*, const float*, int, int, void*, int);
};
const Api& linked_api();

// 8-bit single-channel texture contents (Image FORMAT_L8): 2D w*h, 3D w*h*d, cubemap 6 faces of w*w (+X,-X,+Y,-Y,+Z,-Z).
struct Texture {
    enum Kind { TEXTURE_2D, TEXTURE_3D, CUBEMAP } kind = TEXTURE_2D;
    int width = 0, height = 0, depth = 0;
    std::vector<uint8_t> texels;
};

// The few Variant types a shader parameter of this addon can hold.
class Variant {
public:
    enum Type { NIL, FLOAT, VECTOR3, COLOR, TRANSFORM2D, TRANSFORM3D, TEXTURE };
    Variant() = default;
    Variant(double f) : type_(FLOAT) { v_[0] = float(f); }
    static Variant vector3(float x, float y, float z);
    static Variant color(float r, float g, float b, float a = 1.0f);      // sRGB, as edited in the inspector
    static Variant transform2d(float c0x, float c0y, float c1x, float c1y);  // the 2x2 basis, column-major
    static Variant transform3d(const float colmajor16[16]);
    static Variant texture(std::shared_ptr<const Texture> t);
    Type type() const { return type_; }
    bool is_nil() const { return type_ == NIL; }
    float as_float() const { return v_[0]; }
    const std::array<float, 16>& values() const { return v_; }
    const std::shared_ptr<const Texture>& tex() const { return tex_; }
    bool operator==(const Variant& o) const;
    bool operator!=(const Variant& o) const { return !(*this == o); }

private:
    Type type_ = NIL;
    std::array<float, 16> v_{};
    std::shared_ptr<const Texture> tex_;
};

// Color.srgb_to_linear(): Godot converts `source_color` uniforms before upload (engine behaviour).
float srgb_to_linear(float c);

// custom_shader: the entry shaders are nothing but #defines (shaders/planet_atmosphere_*.gdshader:4-7).
struct ShaderVariant {
    int scatter_model, scatter_steps, cloud_steps, light_mode;
    bool operator==(const ShaderVariant& o) const {
        return scatter_model == o.scatter_model && scatter_steps == o.scatter_steps && cloud_steps == o.cloud_steps &&
               light_mode == o.light_mode;
    }
};
// Looks up a shipped shader by resource path or stem ("res://.../planet_atmosphere_clouds_high.gdshader",
// "planet_atmosphere_clouds_high"; README.md:35's "_clouds_high_m" spelling maps to the _rm file). false if unknown.
bool find_shader_variant(const std::string& shader, ShaderVariant* out, std::string* canonical_name = nullptr);
const std::vector<std::string>& shipped_shader_names();

struct PropertyInfo {
    std::string name;      // "shader_params/<uniform>"
    Variant::Type type;
    bool is_color;         // `source_color` hint
};

enum class LogLevel { PRINT, WARNING, ERROR };
using Logger = std::function<void(LogLevel, const std::string&)>;

class PlanetAtmosphere;

// optical_depth_baker.gd minus the SubViewport: _process #1 launches the bake kernel (_setup_bake copied the uniforms
// by name, :55-59 — here: set_params), _process #2 emits `baked`; the texture stays on the device, owned by the context.
class OpticalDepthBaker {
public:
    static constexpr int STATE_IDLE = 0, STATE_REQUEST_BAKE = 1, STATE_PENDING_RENDER = 2;  // :3-5
    OpticalDepthBaker(const Api& api, b200atmo_ctx* ctx) : api_(api), ctx_(ctx) {}
    void connect_baked(std::function<void()> fn) { baked_.push_back(std::move(fn)); }   // signal baked(texture), :10
    void request_bake(const B200AtmoParams& atmosphere_material);                        // :37-46
    int _process(double delta);                                                          // :66-85; returns a B200ATMO_* code
    int state() const { return state_; }
    bool is_processing() const { return processing_; }                                   // set_process()

private:
    const Api& api_;
    b200atmo_ctx* ctx_;
    int state_ = STATE_IDLE;
    bool processing_ = false;
    B200AtmoParams material_{};
    std::vector<std::function<void()>> baked_;
};

class PlanetAtmosphere {
public:
    static constexpr int MODE_NEAR = 0, MODE_FAR = 1;        // planet_atmosphere.gd:9-10
    static constexpr float SWITCH_MARGIN_RATIO = 1.1f;       // :11

    // _init (:84-108). Creates the device context (no CPU fallback: ok() is false and every call fails without a GPU).
    explicit PlanetAtmosphere(int cuda_device = 0, const Api& api = linked_api(), Logger logger = Logger());
    ~PlanetAtmosphere();
    PlanetAtmosphere(const PlanetAtmosphere&) = delete;
    PlanetAtmosphere& operator=(const PlanetAtmosphere&) = delete;
    bool ok() const { return ctx_ != nullptr; }
    const std::string& init_error() const { return init_error_; }

    // ---- exported properties (:20-54) ----
    float get_planet_radius() const { return planet_radius_; }
    void set_planet_radius(float new_radius);                 // :230-238
    float get_atmosphere_height() const { return atmosphere_height_; }
    void set_atmosphere_height(float new_height);             // :245-253
    const std::string& get_sun_path() const { return sun_path_; }
    void set_sun_path(const std::string& new_sun_path);       // :256-258
    const std::string& get_custom_shader() const { return custom_shader_; }
    // :118-141. "" = null (default shader). Returns false (and logs an error) for an unknown shader.
    bool set_custom_shader(const std::string& shader);
    // Custom step counts (BASELINE scale-ups: 32 in-scatter steps, 128 cloud steps ...), same effect as a forked shader.
    bool set_custom_shader_variant(const ShaderVariant& v);
    float clouds_rotation_speed = 1.0f;                       // :52, degrees per second
    bool force_fullscreen = false;                            // :54

    void _ready();                                            // :111-115

    // ---- shader parameters (:164-218) ----
    void set_shader_param(const std::string& name, const Variant& value);    // deprecated, push_warning (:164-166)
    Variant get_shader_param(const std::string& name);                       // deprecated (:170-172)
    void set_shader_parameter(const std::string& name, const Variant& value);  // :175-176
    Variant get_shader_parameter(const std::string& name) const;             // :179-180 (NIL if never set)
    std::vector<PropertyInfo> _get_property_list() const;                    // :185-198
    bool _get(const std::string& key, Variant* out) const;                   // :201-208 (false = not handled)
    bool _set(const std::string& key, const Variant& value);                 // :211-218

    std::vector<std::string> _get_configuration_warnings() const;            // :221-227

    // ---- what the scene tree provided ----
    struct SunLookup {
        bool exists = false, is_node3d = false;
        float origin[3] = {0.f, 0.f, 0.f};                    // sun.global_transform.origin
    };
    void set_sun_resolver(std::function<SunLookup(const std::string&)> fn) { sun_resolver_ = std::move(fn); }
    void set_global_transform(const float colmajor16[16]);    // node's global_transform (affine)
    void set_ticks_msec_source(std::function<uint64_t()> fn) { ticks_msec_ = std::move(fn); }   // Time.get_ticks_msec
    bool editor_hint = false;                                 // Engine.is_editor_hint()
    bool stdout_verbose = false;                              // OS.is_stdout_verbose()

    struct Camera {                                           // get_viewport().get_camera_3d()
        float position[3];                                    // cam.global_transform.origin
        float near;                                           // cam.near
    };
    // :285-341; cam == nullptr when the viewport has no camera. Returns a B200ATMO_* code (bake errors).
    int _process(double delta, const Camera* cam);

    int get_mode() const { return mode_; }
    float get_far_mesh_size() const { return far_mesh_size_; }             // BoxMesh edge (:314-321)
    float get_extra_cull_margin() const { return extra_cull_margin_; }     // :241-242
    bool is_optical_depth_ready() const { return optical_depth_ready_; }   // `baked` was emitted after the last request
    const OpticalDepthBaker* get_optical_depth_baker() const { return baker_.get(); }
    const B200AtmoParams& get_material_params() const { return params_; }  // the uniform block as uploaded
    ShaderVariant get_variant() const { return variant_; }

    // ---- the draw (what the engine does with the material after _process) ----
    // Spatial-shader built-ins -> B200AtmoCamera; MODE_FAR passes the proxy-cube edge so only its pixels are shaded.
    B200AtmoCamera make_camera(const float inv_projection[16], const float inv_view[16], const float view[16],
                               bool double_precision = false) const;
    int render(const B200AtmoCamera& cam, const float* d_depth, int w, int h, float* d_rgba, uint8_t* d_discard, void* stream);
    int render_composite(const B200AtmoCamera& cam, const float* d_depth, int w, int h, float* d_color_inout, void* stream);
    int render_host(const B200AtmoCamera& cam, const float* h_depth, int w, int h, float* h_rgba, uint8_t* h_discard);
    // the whole transparent pass on host buffers: blends into h_color_inout (B200ATMO_COLOR_RGBA32F / _RGBA16F) in place
    int composite_host(const B200AtmoCamera& cam, const float* h_depth, int w, int h, void* h_color_inout, int color_format);
    std::string last_error() const;

private:
    void update_cull_margin();
    void request_bake_optical_depth();                        // :144-150
    void on_optical_depth_baked();                            // :153-156
    void set_mode(int mode);                                  // :261-282
    void apply_variant();
    void log(LogLevel l, const std::string& m) const;
    int push_params();

    const Api& api_;
    Logger logger_;
    b200atmo_ctx* ctx_ = nullptr;
    std::string init_error_;
    B200AtmoParams params_{};
    float planet_radius_ = 1.0f;        // :19
    float atmosphere_height_ = 0.1f;    // :27
    std::string sun_path_;
    std::string custom_shader_;
    std::string shader_name_;           // canonical stem of the bound shader
    ShaderVariant variant_{};
    int mode_ = MODE_FAR;               // :58
    float prev_atmo_clip_distance_ = 0.0f;
    float far_mesh_size_ = 1.0f;        // :99
    float extra_cull_margin_ = 0.0f;
    bool uses_baked_optical_depth_ = false;
    bool optical_depth_ready_ = false;
    std::unique_ptr<OpticalDepthBaker> baker_;
    std::map<std::string, Variant> raw_params_;   // ShaderMaterial's parameter store: what get_shader_parameter returns
    std::array<float, 16> global_transform_;
    std::function<SunLookup(const std::string&)> sun_resolver_;
    std::function<uint64_t()> ticks_msec_;
};

}  // namespace b200atmo

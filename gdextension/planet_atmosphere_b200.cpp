// See planet_atmosphere_b200.h. SOURCE ONLY (godot-cpp absent in the build image): every decision with behaviour is
// in b200atmo::PlanetAtmosphere, which is compiled and tested in this repo; this file converts types and forwards.
#include "planet_atmosphere_b200.h"

#include <godot_cpp/classes/camera3d.hpp>
#include <godot_cpp/classes/cubemap.hpp>
#include <godot_cpp/classes/engine.hpp>
#include <godot_cpp/classes/image.hpp>
#include <godot_cpp/classes/os.hpp>
#include <godot_cpp/classes/render_scene_buffers_rd.hpp>
#include <godot_cpp/classes/render_scene_data.hpp>
#include <godot_cpp/classes/rendering_device.hpp>
#include <godot_cpp/classes/rendering_server.hpp>
#include <godot_cpp/classes/resource_loader.hpp>
#include <godot_cpp/classes/texture3d.hpp>
#include <godot_cpp/classes/time.hpp>
#include <godot_cpp/classes/viewport.hpp>
#include <godot_cpp/core/class_db.hpp>
#include <godot_cpp/variant/utility_functions.hpp>

#include <cstring>

using namespace godot;

namespace {

void store_colmajor(float out[16], const Transform3D& t) {
    const Basis& b = t.basis;
    for (int c = 0; c < 3; ++c) {
        const Vector3 col = b.get_column(c);
        out[c * 4 + 0] = col.x, out[c * 4 + 1] = col.y, out[c * 4 + 2] = col.z, out[c * 4 + 3] = 0.0f;
    }
    out[12] = t.origin.x, out[13] = t.origin.y, out[14] = t.origin.z, out[15] = 1.0f;
}
void store_colmajor(float out[16], const Projection& p) {
    for (int c = 0; c < 4; ++c)
        for (int r = 0; r < 4; ++r) out[c * 4 + r] = p.columns[c][r];
}

// L8 texel bytes of an Image (converted if needed): what the kernels sample as UNORM8
std::vector<uint8_t> l8_bytes(Ref<Image> img) {
    if (img->is_compressed()) img->decompress();
    if (img->get_format() != Image::FORMAT_L8) img->convert(Image::FORMAT_L8);
    const PackedByteArray d = img->get_data();
    return std::vector<uint8_t>(d.ptr(), d.ptr() + int64_t(img->get_width()) * img->get_height());
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
// Variant conversion (only the types a shader parameter of this addon can hold)
// ---------------------------------------------------------------------------------------------------------------
b200atmo::Variant PlanetAtmosphereB200::to_core(const godot::Variant& v) {
    switch (v.get_type()) {
        case godot::Variant::BOOL: return b200atmo::Variant(bool(v) ? 1.0 : 0.0);
        case godot::Variant::INT:
        case godot::Variant::FLOAT: return b200atmo::Variant(double(v));
        case godot::Variant::VECTOR3: {
            const Vector3 a = v;
            return b200atmo::Variant::vector3(a.x, a.y, a.z);
        }
        case godot::Variant::COLOR: {
            const Color c = v;
            return b200atmo::Variant::color(c.r, c.g, c.b, c.a);
        }
        case godot::Variant::TRANSFORM2D: {
            const Transform2D t = v;
            return b200atmo::Variant::transform2d(t.columns[0].x, t.columns[0].y, t.columns[1].x, t.columns[1].y);
        }
        case godot::Variant::TRANSFORM3D: {
            float m[16];
            store_colmajor(m, Transform3D(v));
            return b200atmo::Variant::transform3d(m);
        }
        case godot::Variant::OBJECT: {
            auto tex = std::make_shared<b200atmo::Texture>();
            Object* o = v;
            if (auto* t3 = Object::cast_to<Texture3D>(o)) {            // u_cloud_shape_texture (NoiseTexture3D ...)
                tex->kind = b200atmo::Texture::TEXTURE_3D;
                tex->width = t3->get_width(), tex->height = t3->get_height(), tex->depth = t3->get_depth();
                const TypedArray<Image> slices = t3->get_data();
                for (int z = 0; z < slices.size(); ++z) {
                    const std::vector<uint8_t> s = l8_bytes(slices[z]);
                    tex->texels.insert(tex->texels.end(), s.begin(), s.end());
                }
            } else if (auto* cm = Object::cast_to<Cubemap>(o)) {         // u_cloud_coverage_cubemap (NoiseCubemap), +X,-X,+Y,-Y,+Z,-Z
                tex->kind = b200atmo::Texture::CUBEMAP;
                tex->width = tex->height = cm->get_width();
                for (int f = 0; f < 6; ++f) {
                    const std::vector<uint8_t> s = l8_bytes(cm->get_layer_data(f));
                    tex->texels.insert(tex->texels.end(), s.begin(), s.end());
                }
            } else if (auto* t2 = Object::cast_to<Texture2D>(o)) {       // u_blue_noise_texture
                tex->width = t2->get_width(), tex->height = t2->get_height();
                tex->texels = l8_bytes(t2->get_image());
            } else {
                return b200atmo::Variant();
            }
            return b200atmo::Variant::texture(tex);
        }
        default: return b200atmo::Variant();
    }
}

godot::Variant PlanetAtmosphereB200::from_core(const b200atmo::Variant& v) {
    const auto& a = v.values();
    switch (v.type()) {
        case b200atmo::Variant::FLOAT: return double(a[0]);
        case b200atmo::Variant::VECTOR3: return Vector3(a[0], a[1], a[2]);
        case b200atmo::Variant::COLOR: return Color(a[0], a[1], a[2], a[3]);
        case b200atmo::Variant::TRANSFORM2D: return Transform2D(Vector2(a[0], a[1]), Vector2(a[2], a[3]), Vector2());
        case b200atmo::Variant::TRANSFORM3D:
            return Transform3D(Basis(Vector3(a[0], a[1], a[2]), Vector3(a[4], a[5], a[6]), Vector3(a[8], a[9], a[10])), Vector3(a[12], a[13], a[14]));
        default: return godot::Variant();   // textures are returned from the wrapper's own store
    }
}

// ---------------------------------------------------------------------------------------------------------------
// node
// ---------------------------------------------------------------------------------------------------------------
PlanetAtmosphereB200::PlanetAtmosphereB200() {
    auto logger = [](b200atmo::LogLevel l, const std::string& m) {
        if (l == b200atmo::LogLevel::ERROR) UtilityFunctions::push_error(String(m.c_str()));
        else if (l == b200atmo::LogLevel::WARNING) UtilityFunctions::push_warning(String(m.c_str()));
        else UtilityFunctions::print(String(m.c_str()));
    };
    core_.reset(new b200atmo::PlanetAtmosphere(0, b200atmo::linked_api(), logger));
    core_->editor_hint = Engine::get_singleton()->is_editor_hint();
    core_->stdout_verbose = OS::get_singleton()->is_stdout_verbose();
    core_->set_ticks_msec_source([] { return uint64_t(Time::get_singleton()->get_ticks_msec()); });
    core_->set_sun_resolver([this](const std::string&) {
        b200atmo::PlanetAtmosphere::SunLookup s;
        if (has_node(sun_path_)) {                                     // planet_atmosphere.gd:328-331
            Node* n = get_node_or_null(sun_path_);
            s.exists = n != nullptr;
            if (auto* n3 = Object::cast_to<Node3D>(n)) {
                s.is_node3d = true;
                const Vector3 o = n3->get_global_transform().origin;
                s.origin[0] = o.x, s.origin[1] = o.y, s.origin[2] = o.z;
            }
        }
        return s;
    });
    // BlueNoiseTexture = preload("./blue_noise.png") (:15, :107)
    Ref<Texture2D> bn = ResourceLoader::get_singleton()->load("res://addons/zylann.atmosphere/blue_noise.png");
    if (bn.is_valid()) core_->set_shader_parameter("u_blue_noise_texture", to_core(bn));
    effect_.instantiate();
    effect_->set_owner_node(this);
    set_process(true);
}

void PlanetAtmosphereB200::_ready() { core_->_ready(); }

void PlanetAtmosphereB200::_process(double delta) {
    float g[16];
    store_colmajor(g, get_global_transform());
    core_->set_global_transform(g);
    b200atmo::PlanetAtmosphere::Camera cam;
    const Camera3D* c = get_viewport() ? get_viewport()->get_camera_3d() : nullptr;
    if (c) {
        const Vector3 o = c->get_global_transform().origin;
        cam.position[0] = o.x, cam.position[1] = o.y, cam.position[2] = o.z;
        cam.near = c->get_near();
    }
    core_->_process(delta, c ? &cam : nullptr);
}

void PlanetAtmosphereB200::set_planet_radius(double r) { core_->set_planet_radius(float(r)); }
double PlanetAtmosphereB200::get_planet_radius() const { return core_->get_planet_radius(); }
void PlanetAtmosphereB200::set_atmosphere_height(double h) { core_->set_atmosphere_height(float(h)); }
double PlanetAtmosphereB200::get_atmosphere_height() const { return core_->get_atmosphere_height(); }
void PlanetAtmosphereB200::set_sun_path(const NodePath& p) {
    sun_path_ = p;
    core_->set_sun_path(String(p).utf8().get_data());
    update_configuration_warnings();                                   // :258
}
NodePath PlanetAtmosphereB200::get_sun_path() const { return sun_path_; }
void PlanetAtmosphereB200::set_custom_shader(const Ref<Shader>& shader) {
    custom_shader_ = shader;
    core_->set_custom_shader(shader.is_valid() ? std::string(shader->get_path().utf8().get_data()) : std::string());
    notify_property_list_changed();                                    // :141
}
Ref<Shader> PlanetAtmosphereB200::get_custom_shader() const { return custom_shader_; }
void PlanetAtmosphereB200::set_clouds_rotation_speed(double s) { core_->clouds_rotation_speed = float(s); }
double PlanetAtmosphereB200::get_clouds_rotation_speed() const { return core_->clouds_rotation_speed; }
void PlanetAtmosphereB200::set_force_fullscreen(bool f) { core_->force_fullscreen = f; }
bool PlanetAtmosphereB200::get_force_fullscreen() const { return core_->force_fullscreen; }

void PlanetAtmosphereB200::set_shader_param(const String& name, const godot::Variant& value) {
    core_->set_shader_param(name.utf8().get_data(), to_core(value));
}
godot::Variant PlanetAtmosphereB200::get_shader_param(const String& name) { return from_core(core_->get_shader_param(name.utf8().get_data())); }
void PlanetAtmosphereB200::set_shader_parameter(const StringName& name, const godot::Variant& value) {
    core_->set_shader_parameter(String(name).utf8().get_data(), to_core(value));
}
godot::Variant PlanetAtmosphereB200::get_shader_parameter(const StringName& name) const {
    return from_core(core_->get_shader_parameter(String(name).utf8().get_data()));
}

void PlanetAtmosphereB200::_get_property_list(List<PropertyInfo>* p_list) const {
    for (const b200atmo::PropertyInfo& p : core_->_get_property_list()) {
        godot::Variant::Type t = godot::Variant::FLOAT;
        PropertyHint hint = PROPERTY_HINT_NONE;
        String hint_string;
        switch (p.type) {
            case b200atmo::Variant::VECTOR3: t = godot::Variant::VECTOR3; break;
            case b200atmo::Variant::COLOR: t = godot::Variant::COLOR; break;
            case b200atmo::Variant::TEXTURE: t = godot::Variant::OBJECT, hint = PROPERTY_HINT_RESOURCE_TYPE, hint_string = "Texture"; break;
            default: break;
        }
        p_list->push_back(PropertyInfo(t, String(p.name.c_str()), hint, hint_string));
    }
}
bool PlanetAtmosphereB200::_get(const StringName& p_name, godot::Variant& r_ret) const {
    b200atmo::Variant v;
    if (!core_->_get(String(p_name).utf8().get_data(), &v)) return false;
    r_ret = from_core(v);
    return true;
}
bool PlanetAtmosphereB200::_set(const StringName& p_name, const godot::Variant& p_value) {
    return core_->_set(String(p_name).utf8().get_data(), to_core(p_value));
}
PackedStringArray PlanetAtmosphereB200::_get_configuration_warnings() const {
    PackedStringArray out;
    for (const std::string& w : core_->_get_configuration_warnings()) out.push_back(String(w.c_str()));
    return out;
}

void PlanetAtmosphereB200::_bind_methods() {
    BIND_CONSTANT(MODE_NEAR);
    BIND_CONSTANT(MODE_FAR);
#define B200_PROP(type, name)                                                                                      \
    ClassDB::bind_method(D_METHOD("set_" #name, "value"), &PlanetAtmosphereB200::set_##name);                       \
    ClassDB::bind_method(D_METHOD("get_" #name), &PlanetAtmosphereB200::get_##name);                                \
    ADD_PROPERTY(PropertyInfo(type, #name), "set_" #name, "get_" #name)
    B200_PROP(godot::Variant::FLOAT, planet_radius);
    B200_PROP(godot::Variant::FLOAT, atmosphere_height);
    B200_PROP(godot::Variant::NODE_PATH, sun_path);
    ClassDB::bind_method(D_METHOD("set_custom_shader", "shader"), &PlanetAtmosphereB200::set_custom_shader);
    ClassDB::bind_method(D_METHOD("get_custom_shader"), &PlanetAtmosphereB200::get_custom_shader);
    ADD_PROPERTY(PropertyInfo(godot::Variant::OBJECT, "custom_shader", PROPERTY_HINT_RESOURCE_TYPE, "Shader"), "set_custom_shader", "get_custom_shader");
    B200_PROP(godot::Variant::FLOAT, clouds_rotation_speed);
    B200_PROP(godot::Variant::BOOL, force_fullscreen);
#undef B200_PROP
    // the draw happens in a CompositorEffect: add it to the Compositor of the camera / WorldEnvironment that should show
    // the atmosphere (the GDScript node added a MeshInstance3D child instead, planet_atmosphere.gd:84-103)
    ClassDB::bind_method(D_METHOD("get_compositor_effect"), &PlanetAtmosphereB200::get_compositor_effect);
    ClassDB::bind_method(D_METHOD("set_shader_param", "param_name", "value"), &PlanetAtmosphereB200::set_shader_param);
    ClassDB::bind_method(D_METHOD("get_shader_param", "param_name"), &PlanetAtmosphereB200::get_shader_param);
    ClassDB::bind_method(D_METHOD("set_shader_parameter", "param_name", "value"), &PlanetAtmosphereB200::set_shader_parameter);
    ClassDB::bind_method(D_METHOD("get_shader_parameter", "param_name"), &PlanetAtmosphereB200::get_shader_parameter);
}

// ---------------------------------------------------------------------------------------------------------------
// the draw: depth read-back -> b200atmo_render_frame_host -> blend_mix -> colour write-back
// ---------------------------------------------------------------------------------------------------------------
B200AtmosphereEffect::B200AtmosphereEffect() {
    // The reference draws its mesh in the TRANSPARENT pass, i.e. after the sky: POST_OPAQUE would run before the sky pass and
    // the sky would then overwrite every far-plane pixel (the limb against space). PRE_TRANSPARENT = after opaque + sky.
    set_effect_callback_type(EFFECT_CALLBACK_TYPE_PRE_TRANSPARENT);
    set_access_resolved_depth(true);
}

void B200AtmosphereEffect::_render_callback(int32_t, RenderData* p_render_data) {
    if (!owner_ || !owner_->core() || !owner_->core()->ok()) return;
    Ref<RenderSceneBuffersRD> buffers = Object::cast_to<RenderSceneBuffersRD>(p_render_data->get_render_scene_buffers().ptr());
    RenderSceneData* scene = p_render_data->get_render_scene_data();
    RenderingDevice* rd = RenderingServer::get_singleton()->get_rendering_device();
    if (buffers.is_null() || !scene || !rd) return;
    const Vector2i size = buffers->get_internal_size();
    const int w = size.x, h = size.y;
    // the built-ins the fragment stage consumed (planet_atmosphere_no_clouds.gdshader:13-26)
    // INV_PROJECTION_MATRIX as shaders see it is NOT cam_projection.inverse(): RenderSceneDataRD::update_ubo multiplies the
    // camera projection by the depth-correction matrix first (y flipped for Vulkan, reverse-Z, depth remapped to 0..1):
    //   projection = correction * cam_projection;  inv_projection = projection.inverse()
    // make_ray (csrc/atmo_device.cuh) builds ndc from the top-left SCREEN_UV and the raw reverse-Z depth, exactly like
    // planet_atmosphere_main.gdshaderinc:128-132, so it needs that corrected matrix (tests/test_godot_conventions.py).
    float inv_p[16], inv_v[16], view[16];
    const Projection correction = Projection::create_depth_correction(true);   // flip_y = true; 4.3: reverse-Z + 0..1 remap
    store_colmajor(inv_p, (correction * scene->get_cam_projection()).inverse());
    store_colmajor(inv_v, scene->get_cam_transform());
    store_colmajor(view, scene->get_cam_transform().affine_inverse());
    const B200AtmoCamera cam = owner_->core()->make_camera(inv_p, inv_v, view);
    // KNOWN COST: texture_get_data / texture_update are synchronous read-backs on the render thread. They stand in for the
    // zero-copy path (export the depth / colour images with VK_KHR_external_memory and import them with
    // cudaImportExternalMemory, then call b200atmo_render_frame_composite_fmt on the device pointers), which needs
    // RenderingDevice to expose the VkDeviceMemory handles. INTEGRATION.md §2.
    const PackedByteArray depth_bytes = rd->texture_get_data(buffers->get_depth_layer(0), 0);   // reverse-Z
    if (depth_bytes.size() != int64_t(w) * h * 4) {
        // D32_SFLOAT has the memory layout of R32_SFLOAT (4 bytes per texel); D24_UNORM_S8 / D32_SFLOAT_S8 do not
        UtilityFunctions::push_error(String("PlanetAtmosphereB200: the scene depth buffer is not a 32-bit float format (") +
                                     String::num_int64(depth_bytes.size()) + String(" bytes for ") + String::num_int64(int64_t(w) * h) +
                                     String(" pixels); add a depth -> R32F copy pass or select D32_SFLOAT"));
        return;
    }
    depth_.resize(size_t(w) * h);
    std::memcpy(depth_.data(), depth_bytes.ptr(), depth_bytes.size());
    // render_mode unshaded + blend_mix (planet_atmosphere_*.gdshader:2) straight into the RGBA16F 3D colour target:
    // depth + colour up, fp32 render + blend on the GPU, colour down (b200atmo_composite_frame_host)
    PackedByteArray color = rd->texture_get_data(buffers->get_color_layer(0), 0);
    if (color.size() != int64_t(w) * h * 8) {
        UtilityFunctions::push_error(String("PlanetAtmosphereB200: the 3D colour target is not RGBA16F"));
        return;
    }
    if (owner_->core()->composite_host(cam, depth_.data(), w, h, color.ptrw(), B200ATMO_COLOR_RGBA16F) != B200ATMO_OK) {
        UtilityFunctions::push_error(String(owner_->core()->last_error().c_str()));
        return;
    }
    rd->texture_update(buffers->get_color_layer(0), 0, color);
}

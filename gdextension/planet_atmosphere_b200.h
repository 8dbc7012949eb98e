// Godot 4.3 binding of the B200 atmosphere path: a Node3D with the exported surface of
// addons/zylann.atmosphere/planet_atmosphere.gd, forwarding to b200atmo::PlanetAtmosphere
// (godot_atmosphere_shader_b200/csrc/node/planet_atmosphere_node.hpp).
// SOURCE ONLY — godot-cpp is not in this repo's build image; see gdextension/README.md.
#pragma once

#include <godot_cpp/classes/compositor_effect.hpp>
#include <godot_cpp/classes/node3d.hpp>
#include <godot_cpp/classes/render_data.hpp>
#include <godot_cpp/classes/shader.hpp>
#include <godot_cpp/classes/texture2d.hpp>
#include <godot_cpp/variant/node_path.hpp>

#include <memory>
#include <vector>

#include "planet_atmosphere_node.hpp"

class PlanetAtmosphereB200;

// Runs after the opaque pass (depth available), where the reference's transparent draw would happen.
class B200AtmosphereEffect : public godot::CompositorEffect {
    GDCLASS(B200AtmosphereEffect, godot::CompositorEffect)
public:
    B200AtmosphereEffect();
    void _render_callback(int32_t p_effect_callback_type, godot::RenderData* p_render_data) override;
    void set_owner_node(PlanetAtmosphereB200* n) { owner_ = n; }

protected:
    static void _bind_methods() {}

private:
    PlanetAtmosphereB200* owner_ = nullptr;
    std::vector<float> depth_;
};

class PlanetAtmosphereB200 : public godot::Node3D {
    GDCLASS(PlanetAtmosphereB200, godot::Node3D)
public:
    enum Mode { MODE_NEAR = 0, MODE_FAR = 1 };                      // planet_atmosphere.gd:9-10

    PlanetAtmosphereB200();                                         // _init, :84-108
    void _ready() override;                                         // :111-115
    void _process(double delta) override;                           // :285-341

    // exported properties (:20-54)
    void set_planet_radius(double r);
    double get_planet_radius() const;
    void set_atmosphere_height(double h);
    double get_atmosphere_height() const;
    void set_sun_path(const godot::NodePath& p);
    godot::NodePath get_sun_path() const;
    void set_custom_shader(const godot::Ref<godot::Shader>& shader);   // the resource path selects the variant
    godot::Ref<godot::Shader> get_custom_shader() const;
    void set_clouds_rotation_speed(double s);
    double get_clouds_rotation_speed() const;
    void set_force_fullscreen(bool f);
    bool get_force_fullscreen() const;

    // :164-180
    void set_shader_param(const godot::String& name, const godot::Variant& value);      // deprecated
    godot::Variant get_shader_param(const godot::String& name);                         // deprecated
    void set_shader_parameter(const godot::StringName& name, const godot::Variant& value);
    godot::Variant get_shader_parameter(const godot::StringName& name) const;

    // :185-227
    void _get_property_list(godot::List<godot::PropertyInfo>* p_list) const;
    bool _get(const godot::StringName& p_name, godot::Variant& r_ret) const;
    bool _set(const godot::StringName& p_name, const godot::Variant& p_value);
    godot::PackedStringArray _get_configuration_warnings() const override;

    b200atmo::PlanetAtmosphere* core() { return core_.get(); }
    godot::Ref<B200AtmosphereEffect> get_compositor_effect() const { return effect_; }

protected:
    static void _bind_methods();

private:
    static b200atmo::Variant to_core(const godot::Variant& v);
    static godot::Variant from_core(const b200atmo::Variant& v);

    std::unique_ptr<b200atmo::PlanetAtmosphere> core_;
    godot::NodePath sun_path_;
    godot::Ref<godot::Shader> custom_shader_;
    godot::Ref<B200AtmosphereEffect> effect_;
};

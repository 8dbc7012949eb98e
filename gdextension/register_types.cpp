// GDExtension entry point (godot-cpp 4.3). Source only: godot-cpp is not in the build image.
#include <gdextension_interface.h>
#include <godot_cpp/core/class_db.hpp>
#include <godot_cpp/godot.hpp>

#include "planet_atmosphere_b200.h"

static void initialize_b200atmo(godot::ModuleInitializationLevel level) {
    if (level != godot::MODULE_INITIALIZATION_LEVEL_SCENE) return;
    godot::ClassDB::register_class<B200AtmosphereEffect>();
    godot::ClassDB::register_class<PlanetAtmosphereB200>();
}
static void uninitialize_b200atmo(godot::ModuleInitializationLevel) {}

extern "C" GDExtensionBool GDE_EXPORT b200atmo_gdextension_init(GDExtensionInterfaceGetProcAddress get_proc, GDExtensionClassLibraryPtr lib,
                                                                 GDExtensionInitialization* init) {
    godot::GDExtensionBinding::InitObject obj(get_proc, lib, init);
    obj.register_initializer(initialize_b200atmo);
    obj.register_terminator(uninitialize_b200atmo);
    obj.set_minimum_library_initialization_level(godot::MODULE_INITIALIZATION_LEVEL_SCENE);
    return obj.init();
}

// TEST INFRASTRUCTURE ONLY — one entry point over the compiled entry shaders of oracle/_ref (see build_ref.py):
// picks the shader the way PlanetAtmosphere.custom_shader does (scatter model + clouds + lighting) and lists the
// #defines each shipped entry shader carries.
#include <cstdint>
#include <cstring>

#include "../../include/b200atmo.h"

struct RefTextures;
struct RefVariant {
    int32_t scatter_model, scatter_steps, cloud_steps, light_mode;
};
#define REF_DECL(name)                                                                                                        \
    extern "C" void ref_##name##_defines(int out[4]);                                                                         \
    extern "C" void ref_##name##_render_frame(const B200AtmoParams*, const RefVariant*, const B200AtmoCamera*, const RefTextures*, \
                                                  const float*, int, int, int, int, int, void*, uint8_t*, int);
REF_DECL(planet_atmosphere_no_clouds)
REF_DECL(planet_atmosphere_clouds)
REF_DECL(planet_atmosphere_clouds_high)
REF_DECL(planet_atmosphere_clouds_high_rm)
REF_DECL(planet_atmosphere_v1_no_clouds)
REF_DECL(planet_atmosphere_v1_clouds)
REF_DECL(planet_atmosphere_v1_clouds_high)

namespace {
typedef void (*DefinesFn)(int[4]);
typedef void (*RenderFn)(const B200AtmoParams*, const RefVariant*, const B200AtmoCamera*, const RefTextures*, const float*, int, int,
                         int, int, int, void*, uint8_t*, int);
struct Entry {
    const char* name;
    DefinesFn defines;
    RenderFn render;
    int lite, atmo_steps, cloud_steps, rm;   // captured at load time, before any render changes the step variables
};
#define REF_ENTRY_ROW(n) {#n, ref_##n##_defines, ref_##n##_render_frame, 0, 0, 0, 0}
Entry g_entries[] = {
    REF_ENTRY_ROW(planet_atmosphere_no_clouds),    REF_ENTRY_ROW(planet_atmosphere_clouds),
    REF_ENTRY_ROW(planet_atmosphere_clouds_high),  REF_ENTRY_ROW(planet_atmosphere_clouds_high_rm),
    REF_ENTRY_ROW(planet_atmosphere_v1_no_clouds), REF_ENTRY_ROW(planet_atmosphere_v1_clouds),
    REF_ENTRY_ROW(planet_atmosphere_v1_clouds_high),
};
const int kEntries = int(sizeof(g_entries) / sizeof(g_entries[0]));
struct Init {
    Init() {
        for (Entry& e : g_entries) {
            int d[4];
            e.defines(d);
            e.lite = d[0], e.atmo_steps = d[1], e.cloud_steps = d[2], e.rm = d[3];
        }
    }
} g_init;
}  // namespace

extern "C" {

int ref_entry_count(void) { return kEntries; }
const char* ref_entry_name(int i) { return (i >= 0 && i < kEntries) ? g_entries[i].name : nullptr; }
// {ATMOSPHERE_LITE, ATMOSPHERE_RAYMARCH_STEPS, CLOUDS_MAX_RAYMARCH_STEPS (0 = clouds disabled), CLOUDS_RAYMARCHED_LIGHTING}
int ref_entry_defines(int i, int out[4]) {
    if (i < 0 || i >= kEntries) return -1;
    out[0] = g_entries[i].lite, out[1] = g_entries[i].atmo_steps, out[2] = g_entries[i].cloud_steps, out[3] = g_entries[i].rm;
    return 0;
}

// rgba: float[h*w*4] in libatmo_ref.so, double[h*w*4] in libatmo_ref64.so (ref_bake_real_size() tells which).
// Renders with entry shader `name`, or (name == NULL) with the first shipped shader whose feature #defines match the
// variant (step counts are runtime values in the compiled shaders). Returns 0, or -1 if there is no such shader.
int ref_render_frame(const char* name, const B200AtmoParams* p, const RefVariant* v, const B200AtmoCamera* cam,
                         const RefTextures* tex, const float* depth, int w, int h, int row_begin, int row_end, int row_stride,
                         void* rgba, uint8_t* discard, int threads) {
    for (const Entry& e : g_entries) {
        if (name) {
            if (std::strcmp(name, e.name) != 0) continue;
        } else {
            const bool lite = v->scatter_model == B200ATMO_SCATTER_V1;
            const bool clouds = v->light_mode != B200ATMO_LIGHT_NONE, rm = v->light_mode == B200ATMO_LIGHT_RAYMARCHED;
            if (lite != (e.lite != 0) || clouds != (e.cloud_steps != 0) || (clouds && rm != (e.rm != 0))) continue;
        }
        e.render(p, v, cam, tex, depth, w, h, row_begin, row_end, row_stride, rgba, discard, threads);
        return 0;
    }
    return -1;
}

}  // extern "C"

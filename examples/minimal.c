/* Minimal C client of libb200atmo.so: renders one 640x360 frame of the default atmosphere through host buffers.
 *   gcc -Iinclude examples/minimal.c -Lgodot_atmosphere_shader_b200 -lb200atmo -Wl,-rpath,$PWD/godot_atmosphere_shader_b200 -lm -o minimal
 * (needs a CUDA device at run time; there is no CPU fallback) */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "b200atmo.h"

static void identity(float m[16]) { memset(m, 0, 16 * sizeof(float)); m[0] = m[5] = m[10] = m[15] = 1.0f; }

int main(void) {
    const int w = 640, h = 360;
    b200atmo_ctx* ctx = NULL;
    if (b200atmo_create(0, &ctx) != B200ATMO_OK) { fprintf(stderr, "%s\n", b200atmo_last_error(NULL)); return 1; }
    B200AtmoParams p;
    b200atmo_default_params(&p);                 /* shader-source defaults */
    p.planet_radius = 1.0f; p.atmosphere_height = 0.2f; p.density = 10.0f; p.scattering_strength = 0.5f;  /* planet_atmosphere.tscn */
    p.sun_position[0] = 5000.0f;
    b200atmo_set_params(ctx, &p);
    b200atmo_set_variant(ctx, B200ATMO_SCATTER_V2, 8, 0, B200ATMO_LIGHT_NONE);   /* planet_atmosphere_no_clouds.gdshader */

    /* camera 3 units from the planet on +Z looking down -Z; Godot-style reverse-Z projection with flipped y */
    B200AtmoCamera cam;
    memset(&cam, 0, sizeof cam);
    identity(cam.inv_view); identity(cam.view); identity(cam.model); identity(cam.inv_projection);
    cam.inv_view[14] = 3.0f; cam.view[14] = -3.0f;
    const float fovy = 60.0f * 3.14159265f / 180.0f, aspect = (float)w / h, n = 0.05f, f = 100.0f;
    const float cot = 1.0f / tanf(0.5f * fovy);
    /* inverse of P = [cot/a 0 0 0; 0 -cot 0 0; 0 0 n/(f-n) fn/(f-n); 0 0 -1 0] (column-major) */
    memset(cam.inv_projection, 0, sizeof cam.inv_projection);
    cam.inv_projection[0] = aspect / cot; cam.inv_projection[5] = -1.0f / cot;
    cam.inv_projection[11] = (f - n) / (f * n); cam.inv_projection[14] = -1.0f; cam.inv_projection[15] = 1.0f / f;

    float* depth = calloc((size_t)w * h, sizeof(float));          /* 0 = far plane everywhere (nothing opaque) */
    float* rgba = malloc((size_t)w * h * 4 * sizeof(float));
    int rc = b200atmo_render_frame_host(ctx, &cam, depth, w, h, rgba, NULL);
    if (rc != B200ATMO_OK) { fprintf(stderr, "%s\n", b200atmo_last_error(ctx)); return 1; }
    const float* c = rgba + 4 * ((size_t)(h / 2) * w + w / 2);
    printf("centre pixel rgba = %.5f %.5f %.5f %.5f ; kernels launched: %llu\n", c[0], c[1], c[2], c[3],
           (unsigned long long)b200atmo_launch_count(ctx));
    free(depth); free(rgba);
    b200atmo_destroy(ctx);
    return 0;
}

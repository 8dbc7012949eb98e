/*
 * b200atmo.h — C-ABI of the B200-native batched atmosphere raymarcher.
 *
 * Drop-in boundary for ONE hot path of Zylann/godot_atmosphere_shader: the per-pixel integration of
 *   addons/zylann.atmosphere/shaders/include/planet_atmosphere_main.gdshaderinc:106-197   (atmosphere_fragment)
 *   addons/zylann.atmosphere/shaders/include/atmosphere_funcs_v2.gdshaderinc:14-101       (LUT fetch + in-scatter march)
 *   addons/zylann.atmosphere/shaders/include/atmosphere_funcs_v1.gdshaderinc:15-63        (v1 "lite" model)
 *   addons/zylann.atmosphere/shaders/include/cloud_funcs.gdshaderinc:25-324               (cloud march + lighting)
 *   addons/zylann.atmosphere/shaders/optical_depth.gdshader:17-69                         (LUT bake)
 *
 * The reference has no FFI: its "operator interface" is Godot's ShaderMaterial uniform set plus the
 * spatial-shader built-ins, driven by the PlanetAtmosphere GDScript node
 * (addons/zylann.atmosphere/planet_atmosphere.gd). Each entry point below names the reference
 * interface it replaces. INTEGRATION.md shows the GDExtension-side binding.
 *
 * Conventions
 *   - plain C, no torch / C++ types in signatures; `stream` is a cudaStream_t passed as void* (NULL = default stream)
 *   - all matrices are COLUMN-MAJOR float[16] (GLSL mat4: m[col*4+row]), vectors are float[3]
 *   - colours are LINEAR (Godot converts `source_color` uniforms sRGB->linear before upload; do the same)
 *   - `d_*` pointers are device memory owned by the caller, `h_*` pointers are host memory owned by the caller
 *   - the context owns the optical-depth LUT and all textures; one context per device; not thread-safe
 *     (the reference only ever runs on Godot's main thread, planet_atmosphere.gd:285)
 *   - every call returns 0 on success or a negative B200ATMO_E_* code; b200atmo_last_error() gives the text
 *   - there is NO CPU fallback: without a CUDA device b200atmo_create() fails with B200ATMO_E_CUDA
 */
#ifndef B200ATMO_H
#define B200ATMO_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200ATMO_VERSION 2

enum {
    B200ATMO_OK = 0,
    B200ATMO_E_INVALID = -1, /* bad argument (null pointer, negative size, unsupported variant ...) */
    B200ATMO_E_CUDA = -2,    /* CUDA runtime error (text in last_error) */
    B200ATMO_E_NOMEM = -3,   /* host or device allocation failed */
    B200ATMO_E_STATE = -4    /* call sequence error */
};

/* Scattering model = which include the entry shader pulls in (planet_atmosphere_main.gdshaderinc:27-31). */
enum {
    B200ATMO_SCATTER_V2 = 0,  /* atmosphere_funcs_v2.gdshaderinc  (planet_atmosphere_*.gdshader)    */
    B200ATMO_SCATTER_V1 = 1   /* atmosphere_funcs_v1.gdshaderinc  (planet_atmosphere_v1_*.gdshader) */
};

/* Cloud lighting = CLOUDS_ENABLED / CLOUDS_RAYMARCHED_LIGHTING (planet_atmosphere_main.gdshaderinc:33,50). */
enum {
    B200ATMO_LIGHT_NONE = 0,       /* clouds disabled               (*_no_clouds.gdshader)         */
    B200ATMO_LIGHT_CHEAP = 1,      /* get_light_cheap               (*_clouds[_high].gdshader)      */
    B200ATMO_LIGHT_RAYMARCHED = 2  /* get_light_raymarched, 6 steps (*_clouds_high_rm.gdshader)     */
};

/*
 * The shader uniform set. Field name = uniform name without the `u_` prefix.
 * Defaults are the shader-source defaults; b200atmo_default_params() fills them.
 */
typedef struct B200AtmoParams {
    /* planet_common.gdshaderinc:4-6 */
    float planet_radius;            /* u_planet_radius = 1.0 */
    float atmosphere_height;        /* u_atmosphere_height = 0.1 */
    float sun_position[3];          /* u_sun_position (world space); only the frame API reads it */
    /* atmosphere_common.gdshaderinc:10 */
    float density;                  /* u_density = 0.2 (affects the LUT) */
    /* atmosphere_funcs_v2.gdshaderinc:8-11 */
    float scattering_strength;      /* 20.0 */
    float scattering_wavelengths[3];/* (700, 530, 440) */
    float atmosphere_modulate[3];   /* (1,1,1), linear */
    float atmosphere_ambient_color[3]; /* (0,0,0.002), linear */
    /* planet_atmosphere_main.gdshaderinc:55,60 */
    float clip_mode;                /* u_clip_mode: vertex stage only; carried for surface parity, unused by kernels */
    float sphere_depth_factor;      /* 0.0 */
    /* cloud_funcs.gdshaderinc:5-16 */
    float cloud_density_scale;      /* 50.0 */
    float cloud_bottom;             /* 0.2 */
    float cloud_top;                /* 0.5 */
    float cloud_blend;              /* 0.5 */
    float cloud_shape_invert;       /* 0.0 (tested == 1.0) */
    float cloud_coverage_bias;      /* 0.0 */
    float cloud_shape_factor;       /* 0.8 */
    float cloud_shape_scale;        /* 1.0 */
    float cloud_coverage_rotation[4]; /* mat2, column-major: (c0.x, c0.y, c1.x, c1.y); identity by default */
    float world_to_model[16];       /* u_world_to_model_matrix; identity by default */
    /* atmosphere_funcs_v1.gdshaderinc:7-11 (only read when scatter model is V1) */
    float day_color0[4];            /* (0.5,0.8,1,1) linear */
    float day_color1[4];
    float night_color0[4];          /* (0.2,0.4,0.8,1) linear */
    float night_color1[4];
    float day_night_transition_scale; /* 2.0 */
} B200AtmoParams;

/*
 * Per-frame constants for the RAY-BATCH API: the two varyings written by atmosphere_vertex
 * (planet_atmosphere_main.gdshaderinc:101-103) plus INV_VIEW_MATRIX (needed by render_clouds,
 * cloud_funcs.gdshaderinc:285). Rays are expressed in the same (view) space as these.
 */
typedef struct B200AtmoFrame {
    float planet_center_view[3];    /* v_planet_center_viewspace */
    float sun_center_view[3];       /* v_sun_center_viewspace */
    float inv_view[16];             /* INV_VIEW_MATRIX */
} B200AtmoFrame;

/*
 * Camera block for the FRAME API = the spatial-shader built-ins consumed by vertex()/fragment()
 * (planet_atmosphere_no_clouds.gdshader:13-26).
 */
typedef struct B200AtmoCamera {
    float inv_projection[16];       /* INV_PROJECTION_MATRIX (Vulkan 0..1 depth, as Godot passes it) */
    float inv_view[16];             /* INV_VIEW_MATRIX */
    float view[16];                 /* VIEW_MATRIX */
    float model[16];                /* MODEL_MATRIX of the PlanetAtmosphere node */
    int32_t double_precision;       /* != 0: DOUBLE_PRECISION workaround, negate inv_view origin (main:118-125) */
    float clip_box_size;            /* MODE_FAR proxy mesh: edge of the BoxMesh centred on the node (planet_atmosphere.gd:302-321);
                                       only pixels whose view ray enters that box in front of the opaque depth are shaded, the
                                       rest are discarded like un-rasterised pixels. 0 = MODE_NEAR / fullscreen quad (u_clip_mode) */
} B200AtmoCamera;

typedef struct b200atmo_ctx b200atmo_ctx;

/* ---- lifetime / errors ---------------------------------------------------------------------- */
int b200atmo_version(void);
size_t b200atmo_sizeof_params(void);   /* ABI self-check for bindings */
size_t b200atmo_sizeof_frame(void);
size_t b200atmo_sizeof_camera(void);
size_t b200atmo_sizeof_peer_targets(void);
/* Replaces: ShaderMaterial.new() + material.shader = ... (planet_atmosphere.gd:84-108). */
int b200atmo_create(int cuda_device, b200atmo_ctx** out);
void b200atmo_destroy(b200atmo_ctx* ctx);
/* Replaces: push_error/push_warning strings (planet_atmosphere.gd:165,171). ctx may be NULL (create errors). */
const char* b200atmo_last_error(const b200atmo_ctx* ctx);

/* ---- uniforms -------------------------------------------------------------------------------- */
void b200atmo_default_params(B200AtmoParams* out);
/* Replaces: material.set_shader_parameter(...) for every scalar/vector/matrix uniform
 * (planet_atmosphere.gd:175-180, 211-218, 230-253, 328-341). A change of planet_radius,
 * atmosphere_height or density marks the LUT stale exactly like _request_bake_optical_depth
 * (planet_atmosphere.gd:144-150, 217-218, 237-238, 252-253). */
int b200atmo_set_params(b200atmo_ctx* ctx, const B200AtmoParams* p);
int b200atmo_get_params(const b200atmo_ctx* ctx, B200AtmoParams* out);
/* Replaces: custom_shader selection, i.e. the compile-time #defines of the entry shaders
 * (shaders/planet_atmosphere_*.gdshader:4-7). 1 <= scatter_steps <= 65536; 1 <= cloud_steps <= 65536 unless light_mode is NONE. */
int b200atmo_set_variant(b200atmo_ctx* ctx, int scatter_model, int scatter_steps, int cloud_steps, int light_mode);

/* ---- textures (host -> device; the context keeps its own device copy) -------------------------- */
/* u_blue_noise_texture (planet_atmosphere_main.gdshaderinc:63,168-169): w x h, 8-bit. The shader fetches
 * texelFetch(tex, ivec2(pixel) & ivec2(0xff), 0): always the top-left 256 x 256 texels, whatever the texture size, so
 * 256 <= w, h <= 4096 is required (a smaller texture would be fetched out of range); the reference ships 256 x 256. */
int b200atmo_upload_blue_noise(b200atmo_ctx* ctx, const uint8_t* h_texels, int w, int h);
/* u_cloud_shape_texture (cloud_funcs.gdshaderinc:10,48-50): nx*ny*nz 8-bit, x fastest; trilinear, repeat. */
int b200atmo_upload_shape3d(b200atmo_ctx* ctx, const uint8_t* h_texels, int nx, int ny, int nz);
/* u_cloud_coverage_cubemap (cloud_funcs.gdshaderinc:15,43-45): 6 faces of res*res 8-bit, face order
 * +X,-X,+Y,-Y,+Z,-Z (noise_cubemap.gd:116-128), row 0 = top. Seamless bilinear, LOD 0. */
int b200atmo_upload_coverage_cube(b200atmo_ctx* ctx, const uint8_t* h_faces6, int res);

/* ---- NoiseCubemap generator (replaces NoiseCubemap._generate_images, noise_cubemap.gd:101-140) ------ */
/*
 * The reference evaluates a Godot `Noise` (FastNoiseLite, engine code outside the addon) at 6*res*res texel
 * directions in a GDScript loop ("This is really slow", noise_cubemap.gd:100). The engine's noise is not part of
 * the reference tree, so the CONTENT is defined here ("b200 gradient fBm v1": hashed 3D gradient noise, quintic
 * fade, `octaves` octaves); the texel -> direction mapping, the `* scale`, density = 0.5 + 0.5*noise and the L8
 * quantisation follow noise_cubemap.gd:110-134 exactly. Face order +X,-X,+Y,-Y,+Z,-Z, row 0 = top.
 */
typedef struct B200AtmoNoise {
    int32_t seed;        /* FastNoiseLite.seed */
    float frequency;     /* FastNoiseLite.frequency (default 0.01) */
    int32_t octaves;     /* fractal_octaves (>= 1) */
    float lacunarity;    /* fractal_lacunarity (2.0) */
    float gain;          /* fractal_gain (0.5) */
} B200AtmoNoise;
/* Generates on the device. h_faces6_out (6*res*res bytes) may be NULL; set_as_coverage != 0 also installs the result
 * as u_cloud_coverage_cubemap (device to device). res in [1, 4096] (noise_cubemap.gd:30). */
int b200atmo_generate_noise_cubemap(b200atmo_ctx* ctx, const B200AtmoNoise* noise, int res, const float scale[3],
                                    uint8_t* h_faces6_out, int set_as_coverage);

/* ---- optical-depth LUT (replaces OpticalDepthBaker, optical_depth_baker.gd:37-85) -------------- */
#define B200ATMO_LUT_SIZE 256
/* NB a re-bake (explicit, or implied by the first render call after planet_radius / atmosphere_height / density changed)
 * rewrites the LUT in place and therefore synchronises the device once — like a texture upload, and unlike every other
 * render call, it must not happen inside a CUDA graph capture. The reference takes two frames for the same event. */
int b200atmo_bake_optical_depth(b200atmo_ctx* ctx, void* stream);
int b200atmo_download_lut(b200atmo_ctx* ctx, float* h_lut256x256);    /* bakes first if stale; synchronises */
/* Debug/parity: device texture layouts as seen by the kernels. */
int b200atmo_download_cube_padded(b200atmo_ctx* ctx, uint8_t* h_out, size_t cap, int* out_res);

/* ---- RAY-BATCH API (device buffers, asynchronous on `stream`) ---------------------------------- */
/*
 * One ray = one fragment invocation. SoA of float4:
 *   d_origin_depth[i] = (ray_origin.xyz [view space], linear_depth)   (main:138,141; depth BEFORE the sphere mix :160)
 *   d_dir_jitter[i]   = (ray_dir.xyz [normalised],    jitter)         (main:142,169)
 *   d_rgba[i]         = (ALBEDO.rgb, ALPHA); (0,0,0,0) when discarded
 *   d_discard[i]      = 1 if `discard` (main:191-196) else 0; may be NULL
 * Replaces: one draw of the atmosphere mesh (fragment stage), planet_atmosphere_*.gdshader fragment().
 */
int b200atmo_render_rays(b200atmo_ctx* ctx, const B200AtmoFrame* frame,
                         const float* d_origin_depth, const float* d_dir_jitter, size_t n_rays,
                         float* d_rgba, uint8_t* d_discard, void* stream);
/* The same call for a ray batch that is a width x height pixel grid in row-major order (n_rays = width*height): warps are
 * mapped to 8x4 pixel tiles instead of 32 consecutive rays, so the lanes of a warp enter / leave the cloud shell together
 * (lane-occupancy model profiles/r02/warp_model.txt: 8-15 % fewer issued instructions on the cloud variants). Results
 * are bit-identical to b200atmo_render_rays. */
int b200atmo_render_rays_2d(b200atmo_ctx* ctx, const B200AtmoFrame* frame,
                            const float* d_origin_depth, const float* d_dir_jitter, int width, int height,
                            float* d_rgba, uint8_t* d_discard, void* stream);
/* Same call with HOST buffers: H2D of the rays, render, D2H of the result; synchronous. */
int b200atmo_render_rays_host(b200atmo_ctx* ctx, const B200AtmoFrame* frame,
                              const float* h_origin_depth, const float* h_dir_jitter, size_t n_rays,
                              float* h_rgba, uint8_t* h_discard);

/* ---- FRAME API --------------------------------------------------------------------------------- */
/*
 * Runs atmosphere_fragment for every pixel of a w x h target (rows [row_begin, row_end) only — the
 * screen-tile shard used for multi-GPU): depth fetch, ray generation (main:128-142), the varyings of
 * atmosphere_vertex (main:101-103) from camera->view/model and params.sun_position, blue-noise fetch.
 *   d_depth : w*h floats, the depth texture (non-linear, as sampled by `texture(depth_texture, uv).x`), row 0 = top
 *   d_rgba  : w*h float4, only rows [row_begin,row_end) are written
 *   d_discard: w*h bytes or NULL
 */
int b200atmo_render_frame(b200atmo_ctx* ctx, const B200AtmoCamera* cam, const float* d_depth,
                          int w, int h, int row_begin, int row_end,
                          float* d_rgba, uint8_t* d_discard, void* stream);
/* Scheduling note (no effect on any pixel): with B200ATMO_LIGHT_RAYMARCHED the frame calls and b200atmo_render_rays_2d remember,
 * per stream and launch geometry, how long every thread block of the previous launch ran and dispatch the next launch's blocks
 * longest-first (one extra 10-us kernel per launch). Environment B200ATMO_BLOCK_ORDER=0 switches it off. */
/* Same as b200atmo_render_frame, but the result is alpha-blended straight into the frame's colour buffer, which is what
 * the reference's `render_mode unshaded` + default blend_mix does in the ROP (planet_atmosphere_*.gdshader:2):
 *   color.rgb = ALBEDO * ALPHA + color.rgb * (1 - ALPHA)   for every non-discarded pixel;  color.a is left untouched.
 * d_color_inout: w*h float4 (linear HDR colour), rows [row_begin,row_end) are updated in place. */
int b200atmo_render_frame_composite(b200atmo_ctx* ctx, const B200AtmoCamera* cam, const float* d_depth, int w, int h,
                                    int row_begin, int row_end, float* d_color_inout, void* stream);
/* Same, for either colour-buffer format. Godot's Forward+ renderer keeps the 3D colour target in RGBA16F: the blend is
 * computed in fp32 from the fp32 ALBEDO/ALPHA and the stored value is rounded to nearest-even, as the ROP does; the
 * destination alpha bits are left untouched. d_color_inout: w*h float4 (RGBA32F) or w*h half4 (RGBA16F). */
enum {
    B200ATMO_COLOR_RGBA32F = 0,
    B200ATMO_COLOR_RGBA16F = 1
};
int b200atmo_render_frame_composite_fmt(b200atmo_ctx* ctx, const B200AtmoCamera* cam, const float* d_depth, int w, int h,
                                        int row_begin, int row_end, void* d_color_inout, int color_format, void* stream);
/* HOST-buffer variant, the whole transparent pass of the reference in one call: H2D depth + colour, render + blend,
 * D2H colour (in place in h_color_inout); synchronous. RGBA16F moves 4 + 8 B/pixel up and 8 B/pixel down (the
 * un-blended b200atmo_render_frame_host: 4 up, 16 down). */
int b200atmo_composite_frame_host(b200atmo_ctx* ctx, const B200AtmoCamera* cam, const float* h_depth, int w, int h,
                                  void* h_color_inout, int color_format);
/* Frame front-end only (main:101-103,128-142): depth buffer -> the SoA ray buffers of the ray-batch API and
 * the frame constants that go with them (frame_out may be NULL). */
int b200atmo_make_rays(b200atmo_ctx* ctx, const B200AtmoCamera* cam, const float* d_depth, int w, int h,
                       float* d_origin_depth, float* d_dir_jitter, B200AtmoFrame* frame_out, void* stream);
/* HOST-buffer variant (the e2e path): H2D depth, render, D2H rgba (+discard); synchronous. */
int b200atmo_render_frame_host(b200atmo_ctx* ctx, const B200AtmoCamera* cam, const float* h_depth,
                               int w, int h, float* h_rgba, uint8_t* h_discard);

/* Output-format variants of the two host-buffer frame calls. B200ATMO_COLOR_RGBA16F writes h_rgba as w*h half4 (8 B/pixel
 * instead of 16): every channel is the fp32 result rounded to nearest-even, (0,0,0,0) when discarded — i.e. exactly what
 * storing ALBEDO/ALPHA into Godot's RGBA16F colour target does. The D2H of the result is the PCIe-bound leg of a host frame,
 * so this halves the end-to-end time; the fp32 format stays the parity path. */
int b200atmo_render_frame_host_fmt(b200atmo_ctx* ctx, const B200AtmoCamera* cam, const float* h_depth,
                                   int w, int h, void* h_rgba, int rgba_format, uint8_t* h_discard);
/* Device-buffer frame call with an output format (rows [row_begin,row_end) of d_rgba: float4 or half4 per pixel). */
int b200atmo_render_frame_fmt(b200atmo_ctx* ctx, const B200AtmoCamera* cam, const float* d_depth, int w, int h, int row_begin,
                              int row_end, void* d_rgba, int rgba_format, uint8_t* d_discard, void* stream);

/* ---- multi-GPU: fused render + delivery over NVLink / NVSwitch peer memory ------------------------------------- */
/*
 * The screen-tile shard (one process per GPU; rank g renders its own tile, a row band [g*H/G, (g+1)*H/G) or the 8-row tiles
 * g, g+G, ... of one frame) ends with the consuming rank(s) holding every pixel. Instead of rendering locally and then
 * gathering, these calls make the render kernel store each finished pixel directly into the consuming ranks' copies of a
 * symmetric buffer (same layout on all GPUs, mapped into this process by CUDA IPC / symmetric memory, e.g.
 * torch.distributed._symmetric_memory): one peer-to-peer store per listed rank, or — with `d_rgba_multicast`, the NVLS
 * multicast mapping of that buffer — one store that the NVSwitch replicates. The calls are asynchronous on `stream`; the
 * pixels of the other ranks are complete on a consumer after an inter-rank barrier that follows the kernels (e.g. the
 * symmetric-memory handle's barrier), or when the kernel returns if it carries the hand-shake itself (B200AtmoPeerSync).
 * No discard mask. Measured delivery rates: DESIGN.md §7.
 */
#define B200ATMO_MAX_PEERS 8
/*
 * Hand-shake of the fused render + delivery, carried out BY THE RENDER KERNEL (no barrier kernel, no extra launch). Flags are
 * uint32_t epochs in symmetric memory (one array per rank, mapped everywhere); epochs grow by one per frame (wrap-around safe).
 *   producer : the kernel's LAST block to finish (per-block fence + atomic block count) publishes `epoch` into element
 *              `done_slot` of every listed consumer flag array            -> "my pixels of frame `epoch` have landed"
 *              every block first waits until d_credit_flags[credit_first_slot .. +n_credit) have reached credit_epoch
 *                                                                         -> "the buffer I am about to overwrite was consumed"
 *   consumer : the kernel's FIRST block publishes consumed_epoch into element consumed_slot of the listed arrays when the
 *              kernel starts, i.e. after everything queued before it      -> "I have finished reading frame consumed_epoch"
 *              the last block then waits until d_wait_flags[wait_first_slot .. +n_wait) have reached `epoch`: the kernel
 *              only completes when every producer's pixels are here, so work queued behind it may read them.
 * Any part may be left empty (n_* = 0 / NULL). Waits give up after ~2 s (b200atmo_peers_wait_timeouts).
 */
typedef struct B200AtmoPeerSync {
    void* d_done_flags[B200ATMO_MAX_PEERS];
    int32_t n_done_flags;
    int32_t done_slot;
    uint32_t epoch;
    uint32_t credit_epoch;
    const void* d_credit_flags;
    int32_t credit_first_slot;
    int32_t n_credit;
    void* d_consumed_flags[B200ATMO_MAX_PEERS];
    int32_t n_consumed_flags;
    int32_t consumed_slot;
    uint32_t consumed_epoch;
    int32_t n_wait;
    const void* d_wait_flags;
    int32_t wait_first_slot;
    int32_t reserved;
} B200AtmoPeerSync;
typedef struct B200AtmoPeerTargets {
    void* d_rgba_peers[B200ATMO_MAX_PEERS]; /* the symmetric buffer as mapped here, one pointer per rank (own rank included) */
    int32_t n_peers;                        /* 1..B200ATMO_MAX_PEERS */
    void* d_rgba_multicast;                 /* NVLS multicast mapping of the same buffer, or NULL */
    uint64_t elem_offset;                   /* float4 elements added to the pixel / ray index (this rank's slot) */
    int32_t first_peer;                     /* P2P path: index the store loop starts at (wraps around). Pass (rank + 1) % n_peers so
                                               that at any moment the ranks address DIFFERENT destinations (no incast on one NVLink port) */
    int32_t use_tma;                        /* b200atmo_render_rays_peers, P2P path: != 0 stages each block's 128 results in shared
                                               memory and sends them with one TMA bulk store (cp.async.bulk) per peer instead of one
                                               STG.128 per thread and peer. Needs 16-byte aligned buffers and elem_offset
                                               (half4 tiles: an even elem_offset). */
    int32_t rgba_format;                    /* B200ATMO_COLOR_RGBA32F (float4 per pixel) or B200ATMO_COLOR_RGBA16F (half4 per pixel,
                                               each channel the fp32 result rounded to nearest-even): the tile format on the wire.
                                               Half the NVLink bytes; elem_offset counts pixels of that format. */
    int32_t reserved;
    B200AtmoPeerSync sync;                  /* hand-shake fused into the kernel (all zero = none: follow the call with a barrier) */
} B200AtmoPeerTargets;
/* Delivery patterns are chosen by the pointer list alone: all ranks' mappings = all-gather (every GPU ends with every
 * tile); ONLY the consuming rank's mapping (n_peers = 1) = deliver-to-root (1/world of the fabric traffic of the
 * all-gather; the root's NVLink ingress is then the only loaded port). */
int b200atmo_render_frame_peers(b200atmo_ctx* ctx, const B200AtmoCamera* cam, const float* d_depth, int w, int h,
                                int row_begin, int row_end, const B200AtmoPeerTargets* targets, void* stream);
/* Interleaved screen-tile shard of ONE frame ("prefer interleaved tile-rows ... with a fixed, deterministic
 * mapping"): the frame is cut into 8-row tiles and this call renders tiles first_tile, first_tile + tile_pitch, ... in one
 * launch. Rank g of G passes (g, G). Unlike contiguous bands, every rank gets the same mix of sky, limb, ground and cloud
 * rows, so the ranks finish together. Pixels are bit-identical to the unsharded frame. */
int b200atmo_render_frame_peers_interleaved(b200atmo_ctx* ctx, const B200AtmoCamera* cam, const float* d_depth, int w, int h,
                                            int first_tile, int tile_pitch, const B200AtmoPeerTargets* targets, void* stream);
int b200atmo_render_rays_peers(b200atmo_ctx* ctx, const B200AtmoFrame* frame, const float* d_origin_depth,
                               const float* d_dir_jitter, size_t n_rays, const B200AtmoPeerTargets* targets, void* stream);
/* The same hand-shake as stand-alone calls (for protocols driven from the host side or mixed with other work):
 *   b200atmo_peers_wait   : queues a one-warp kernel on `stream` that returns when d_flags[first_slot + k] has reached `epoch`
 *                           for every k < n_slots (acquire at system scope; wrap-around safe). A consumer calls it after its
 *                           own render: work queued behind it sees every producer's pixels. Producers use it for flow control
 *                           (wait for the consumer's "consumed" flag before overwriting a buffer).
 *   b200atmo_peers_signal : queues a kernel that publishes `epoch` into element `slot` of each listed flag array (release at
 *                           system scope), ordered after everything already queued on `stream` — e.g. "I have consumed frame e".
 * A wait gives up after ~2 s (a peer died) instead of hanging the GPU; b200atmo_peers_wait_timeouts() counts those.
 * A wait SPINS ON THE GPU: only wait for flags that another GPU publishes, or that work queued EARLIER on this GPU publishes.
 * A flag published by a launch queued later on this same GPU may never arrive, because streams can share a hardware queue. */
int b200atmo_peers_wait(b200atmo_ctx* ctx, const void* d_flags, int first_slot, int n_slots, uint32_t epoch, void* stream);
int b200atmo_peers_signal(b200atmo_ctx* ctx, void* const* d_flags_peers, int n_peers, int slot, uint32_t epoch, void* stream);
int b200atmo_peers_wait_timeouts(b200atmo_ctx* ctx);   /* synchronises the device; >= 0 = number of timed-out waits so far */

/*
 * Pipelined form of b200atmo_render_frame_host for a stream of frames (one per _process tick, or the tiles of an
 * offscreen target; B200ATMO_PIPELINE_SLOTS slots — with 3 in use the copy engine that downloads frame k never waits for the
 * host to submit frame k+2): enqueues H2D(depth) -> render -> D2H(rgba [, discard]) on pipeline slot `slot` and returns without
 * waiting; b200atmo_frame_wait(ctx, slot) blocks until that frame's host buffers are complete. Each slot owns a stream
 * and device staging, so while frame k downloads (the PCIe-bound leg: 16 B/pixel), frame k+1 on the other slot uploads
 * and renders. Uniforms, variant and camera are captured at submit time. The host buffers must stay valid (and should be
 * pinned) until the wait returns; a slot must be waited on before it is submitted again (B200ATMO_E_STATE otherwise).
 * Texture uploads, re-bakes and b200atmo_destroy wait for frames in flight themselves.
 * Results are bit-identical to b200atmo_render_frame_host.
 */
#define B200ATMO_PIPELINE_SLOTS 4
int b200atmo_render_frame_host_submit(b200atmo_ctx* ctx, const B200AtmoCamera* cam, const float* h_depth,
                                      int w, int h, float* h_rgba, uint8_t* h_discard, int slot);
int b200atmo_render_frame_host_submit_fmt(b200atmo_ctx* ctx, const B200AtmoCamera* cam, const float* h_depth,
                                          int w, int h, void* h_rgba, int rgba_format, uint8_t* h_discard, int slot);
int b200atmo_frame_wait(b200atmo_ctx* ctx, int slot);

/* Number of kernels this context has launched since creation (bench.py's gpu_launches claim). */
uint64_t b200atmo_launch_count(const b200atmo_ctx* ctx);
/* Number of times the frame front end rebuilt its per-column / per-row ray tables (a cache keyed by stream, frame size and
 * projection; alternating viewports must not rebuild, and no frame call ever synchronises the device for them). */
uint64_t b200atmo_table_build_count(const b200atmo_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* B200ATMO_H */

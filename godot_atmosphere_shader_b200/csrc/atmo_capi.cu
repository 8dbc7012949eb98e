// C-ABI of include/b200atmo.h: context, uniforms, texture uploads, LUT bake, ray-batch and frame
// entry points. Host code only (the kernels live in atmo_kernels.cu). There is no CPU fallback:
// every compute entry point launches CUDA kernels or fails with B200ATMO_E_CUDA.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "atmo_consts.h"
#include "atmo_internal.h"

using namespace b200atmo;

struct b200atmo_ctx {
    int device = 0;
    B200AtmoParams params;
    Variant variant;
    bool lut_stale = true;
    float* d_lut = nullptr;       // [256][256]
    float* d_lut_pad = nullptr;   // [258][258]
    float4* d_lut_cells = nullptr; // [257][257] bilinear cells
    uint8_t* d_cube_pad_u8 = nullptr;
    float4* d_cube_cells = nullptr;
    int cube_res = 0;
    float cube_max = 1.0f;         // largest coverage texel / 255 (bounds the cloud density, atmo_consts.h cloud_hc_min)
    float4* d_shape_cells = nullptr;
    int nx = 0, ny = 0, nz = 0;
    uint8_t* d_blue = nullptr;
    int bn_w = 0, bn_h = 0;
    // staging for the host-buffer entry points (grown on demand, reused across calls)
    void* d_stage_in0 = nullptr;
    void* d_stage_in1 = nullptr;
    void* d_stage_out = nullptr;
    uint8_t* d_stage_disc = nullptr;
    size_t cap_in0 = 0, cap_in1 = 0, cap_out = 0, cap_disc = 0;
    cudaStream_t streams[2] = {nullptr, nullptr};
    // pipelined host-buffer frames (b200atmo_render_frame_host_submit / b200atmo_frame_wait): per slot its own stream
    // and device staging, so the D2H of frame k overlaps the H2D + kernel of frame k+1
    struct Slot {
        cudaStream_t stream = nullptr;
        void* d_depth = nullptr;
        void* d_rgba = nullptr;
        void* d_disc = nullptr;
        size_t cap_depth = 0, cap_rgba = 0, cap_disc = 0;
        bool in_flight = false;
    } slots[B200ATMO_PIPELINE_SLOTS];
    // frame front end: per-column / per-row tables of INV_PROJECTION_MATRIX * ndc, cached PER STREAM (kTableStreams
    // streams x kTableEntries (w, h, projection) keys, LRU). An entry is only ever built and read by work on its own
    // stream, so builds are stream-ordered and no call ever synchronises: alternating viewports hit the cache, a
    // projection that changes every frame (TAA jitter) costs one 2-us kernel per frame. Callers that use more streams
    // than kTableStreams get the inline path (the kernel computes the same values per pixel: bit-identical).
    static constexpr int kTableStreams = 16, kTableEntries = 4;
    struct TableEntry {
        float4* d = nullptr;
        size_t cap = 0;
        int w = 0, h = 0;
        float inv_proj[16] = {};
        uint64_t tick = 0;
    };
    // heaviest-first block dispatch of the raymarched-cloud launches (atmo_kernels.cu: block_order_kernel): per stream and launch
    // geometry the per-block cycle counts of the previous launch and the dispatch order sorted from them. Same ownership rule as
    // the ray tables: an entry is only touched by work queued on its stream.
    static constexpr int kOrderEntries = 4;
    struct OrderEntry {
        unsigned* d_cost = nullptr;     // [n] + d_order [n] in one allocation
        unsigned* d_order = nullptr;
        size_t cap = 0;
        int kind = 0, w = 0, h = 0, row_begin = 0, row_end = 0, row_pitch = 0, cloud_steps = 0;
        unsigned n = 0;
        bool valid = false;             // d_order holds a permutation sorted from a previous launch
        uint64_t tick = 0;
    };
    struct TableStream {
        cudaStream_t stream = nullptr;
        bool used = false;
        TableEntry e[kTableEntries];
        OrderEntry o[kOrderEntries];
    } tables[kTableStreams];
    int block_order_enabled = 1;        // B200ATMO_BLOCK_ORDER=0 switches it off (A/B measurements); 2 = also for the cheap light
                                        // (needs a library built with -DB200ATMO_ORDER_CHEAP=1)
    uint64_t table_tick = 0;
    uint64_t table_builds = 0;
    // fused completion signal of the peers kernels: a ring of block counters (zero between launches; each launch takes the
    // next one, so launches on different streams never share one) and the count of timed-out waits
    static constexpr int kBlockCounters = 64;
    unsigned* d_block_counters = nullptr;   // [kBlockCounters] counters + [1] timeouts
    int next_block_counter = 0;
    uint64_t launches = 0;
    std::string last_error;
};

namespace {

thread_local std::string g_create_error;

int fail(b200atmo_ctx* ctx, int code, const std::string& msg) {
    if (ctx) ctx->last_error = msg;
    else g_create_error = msg;
    return code;
}

#define CU_TRY(ctx, expr)                                                                                     \
    do {                                                                                                      \
        cudaError_t _e = (expr);                                                                              \
        if (_e != cudaSuccess)                                                                                \
            return fail(ctx, _e == cudaErrorMemoryAllocation ? B200ATMO_E_NOMEM : B200ATMO_E_CUDA,            \
                        std::string(#expr) + ": " + cudaGetErrorString(_e));                                  \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

int ensure(b200atmo_ctx* ctx, void** p, size_t* cap, size_t need) {
    if (*cap >= need) return B200ATMO_OK;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *cap = 0;
    CU_TRY(ctx, cudaMalloc(p, need));
    *cap = need;
    return B200ATMO_OK;
}

DeviceTextures textures_of(const b200atmo_ctx* ctx) {
    DeviceTextures t;
    t.lut_pad = ctx->d_lut_pad;
    t.lut_cells = ctx->d_lut_cells;
    t.cube_cells = ctx->d_cube_cells;
    t.cube_res = ctx->cube_res;
    t.cube_max = ctx->cube_max;
    t.shape_cells = ctx->d_shape_cells;
    t.nx = ctx->nx;
    t.ny = ctx->ny;
    t.nz = ctx->nz;
    t.blue_noise = ctx->d_blue;
    t.bn_w = ctx->bn_w;
    t.bn_h = ctx->bn_h;
    return t;
}

// Frames submitted with b200atmo_render_frame_host_submit read the LUT / textures asynchronously: anything that rewrites
// or frees those buffers waits for them first (rare paths: re-bake, texture upload, destroy).
int drain_slots(b200atmo_ctx* ctx) {
    for (auto& sl : ctx->slots) {
        if (!sl.in_flight) continue;
        CU_TRY(ctx, cudaStreamSynchronize(sl.stream));
        sl.in_flight = false;
    }
    return B200ATMO_OK;
}

int bake_if_stale(b200atmo_ctx* ctx, cudaStream_t s) {
    if (!ctx->lut_stale) return B200ATMO_OK;
    int drained = drain_slots(ctx);
    if (drained != B200ATMO_OK) return drained;
    CU_TRY(ctx, cudaDeviceSynchronize());   // rare path: frames on any stream may still sample the old LUT
    CU_TRY(ctx, launch_bake_lut(ctx->params.planet_radius, ctx->params.atmosphere_height, ctx->params.density, ctx->d_lut,
                                ctx->d_lut_pad, ctx->d_lut_cells, s));
    ctx->launches += 2;
    // re-bakes are rare (R, H or u_density changed); finishing here keeps later launches on OTHER streams safe
    CU_TRY(ctx, cudaStreamSynchronize(s));
    ctx->lut_stale = false;
    return B200ATMO_OK;
}

// device allocation that frees itself unless released (no leaks on the error paths of the upload functions)
struct DevBuf {
    void* p = nullptr;
    ~DevBuf() {
        if (p) cudaFree(p);
    }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes); }
    template <class T> T* as() const { return static_cast<T*>(p); }
    template <class T> T* release() {
        T* q = static_cast<T*>(p);
        p = nullptr;
        return q;
    }
};

// installs a cube whose raw faces are either on the host (h_faces6) or already on the device (d_src); the raw faces and
// the padded fp32 copy only live during the build, the context keeps the integer seamless layout (for
// b200atmo_download_cube_padded) and the cells the kernels sample
int upload_cube(b200atmo_ctx* ctx, const uint8_t* h_faces6, int res, cudaStream_t s, const uint8_t* d_src = nullptr) {
    const size_t raw = size_t(6) * res * res, pad = size_t(6) * (res + 2) * (res + 2);
    int drained = drain_slots(ctx);
    if (drained != B200ATMO_OK) return drained;
    DevBuf d_raw, d_pad8, d_pad, d_cells;
    CU_TRY(ctx, d_raw.alloc(raw));
    CU_TRY(ctx, d_pad8.alloc(pad));
    CU_TRY(ctx, d_pad.alloc(pad * sizeof(float)));
    CU_TRY(ctx, d_cells.alloc(size_t(6) * (res + 1) * (res + 1) * sizeof(float4)));
    CU_TRY(ctx, cudaMemcpyAsync(d_raw.p, d_src ? d_src : h_faces6, raw, d_src ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s));
    // largest texel (the seamless apron only repeats face texels or averages three of them): faces generated on the device
    // are read back once for it (uploads are rare)
    unsigned max_texel = 0;
    {
        std::vector<uint8_t> tmp;
        const uint8_t* src = h_faces6;
        if (d_src) {
            tmp.resize(raw);
            CU_TRY(ctx, cudaMemcpyAsync(tmp.data(), d_src, raw, cudaMemcpyDeviceToHost, s));
            CU_TRY(ctx, cudaStreamSynchronize(s));
            src = tmp.data();
        }
        for (size_t i = 0; i < raw; ++i) max_texel = src[i] > max_texel ? src[i] : max_texel;
    }
    CU_TRY(ctx, launch_cube_pad(d_raw.as<uint8_t>(), res, d_pad8.as<uint8_t>(), d_pad.as<float>(), d_cells.as<float4>(), s));
    ctx->launches += 2;
    CU_TRY(ctx, cudaStreamSynchronize(s));
    cudaFree(ctx->d_cube_pad_u8);
    cudaFree(ctx->d_cube_cells);
    ctx->d_cube_pad_u8 = d_pad8.release<uint8_t>();
    ctx->d_cube_cells = d_cells.release<float4>();
    ctx->cube_res = res;
    ctx->cube_max = float(max_texel) / 255.0f;
    return B200ATMO_OK;
}

int upload_shape(b200atmo_ctx* ctx, const uint8_t* h, int nx, int ny, int nz, cudaStream_t s) {
    const size_t raw = size_t(nx) * ny * nz, pad = size_t(nx + 2) * (ny + 2) * (nz + 2);
    int drained = drain_slots(ctx);
    if (drained != B200ATMO_OK) return drained;
    DevBuf d_raw, d_pad, d_cells;
    CU_TRY(ctx, d_raw.alloc(raw));
    CU_TRY(ctx, d_pad.alloc(pad * sizeof(float)));
    CU_TRY(ctx, d_cells.alloc(size_t(nx + 1) * (ny + 1) * (nz + 1) * 2 * sizeof(float4)));
    CU_TRY(ctx, cudaMemcpyAsync(d_raw.p, h, raw, cudaMemcpyHostToDevice, s));
    CU_TRY(ctx, launch_shape_pad(d_raw.as<uint8_t>(), nx, ny, nz, d_pad.as<float>(), d_cells.as<float4>(), s));
    ctx->launches += 2;
    CU_TRY(ctx, cudaStreamSynchronize(s));
    cudaFree(ctx->d_shape_cells);
    ctx->d_shape_cells = d_cells.release<float4>();
    ctx->nx = nx;
    ctx->ny = ny;
    ctx->nz = nz;
    return B200ATMO_OK;
}

}  // namespace

extern "C" {

enum { kOrderFrame = 1, kOrderRays2D = 2, kOrderPeers = 3 };
static b200atmo_ctx::OrderEntry* block_order_begin(b200atmo_ctx* ctx, cudaStream_t s, int kind, const DevConsts& c, unsigned n, RayIO& io);
static int block_order_finish(b200atmo_ctx* ctx, b200atmo_ctx::OrderEntry* e, cudaStream_t s);

int b200atmo_version(void) { return B200ATMO_VERSION; }
size_t b200atmo_sizeof_params(void) { return sizeof(B200AtmoParams); }
size_t b200atmo_sizeof_frame(void) { return sizeof(B200AtmoFrame); }
size_t b200atmo_sizeof_camera(void) { return sizeof(B200AtmoCamera); }
size_t b200atmo_sizeof_peer_targets(void) { return sizeof(B200AtmoPeerTargets); }

void b200atmo_default_params(B200AtmoParams* p) {
    if (!p) return;
    std::memset(p, 0, sizeof(*p));
    p->planet_radius = 1.0f;
    p->atmosphere_height = 0.1f;
    p->density = 0.2f;
    p->scattering_strength = 20.0f;
    p->scattering_wavelengths[0] = 700.0f;
    p->scattering_wavelengths[1] = 530.0f;
    p->scattering_wavelengths[2] = 440.0f;
    for (int k = 0; k < 3; ++k) p->atmosphere_modulate[k] = 1.0f;
    p->atmosphere_ambient_color[2] = 0.002f;
    p->cloud_density_scale = 50.0f;
    p->cloud_bottom = 0.2f;
    p->cloud_top = 0.5f;
    p->cloud_blend = 0.5f;
    p->cloud_shape_factor = 0.8f;
    p->cloud_shape_scale = 1.0f;
    p->cloud_coverage_rotation[0] = 1.0f;
    p->cloud_coverage_rotation[3] = 1.0f;
    for (int k = 0; k < 4; ++k) p->world_to_model[k * 5] = 1.0f;
    const float day[4] = {0.5f, 0.8f, 1.0f, 1.0f}, night[4] = {0.2f, 0.4f, 0.8f, 1.0f};
    for (int k = 0; k < 4; ++k) {
        p->day_color0[k] = p->day_color1[k] = day[k];
        p->night_color0[k] = p->night_color1[k] = night[k];
    }
    p->day_night_transition_scale = 2.0f;
}

int b200atmo_create(int cuda_device, b200atmo_ctx** out) {
    if (!out) return fail(nullptr, B200ATMO_E_INVALID, "b200atmo_create: out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0)
        return fail(nullptr, B200ATMO_E_CUDA,
                    std::string("b200atmo_create: no CUDA device (") + cudaGetErrorString(e) + "); there is no CPU fallback");
    if (cuda_device < 0 || cuda_device >= count) return fail(nullptr, B200ATMO_E_INVALID, "b200atmo_create: bad device index");
    b200atmo_ctx* ctx = new (std::nothrow) b200atmo_ctx();
    if (!ctx) return fail(nullptr, B200ATMO_E_NOMEM, "b200atmo_create: out of host memory");
    ctx->device = cuda_device;
    if (const char* e = std::getenv("B200ATMO_BLOCK_ORDER")) ctx->block_order_enabled = std::atoi(e);
    b200atmo_default_params(&ctx->params);
    DeviceGuard g(cuda_device);
    auto bail = [&](int code) {
        g_create_error = ctx->last_error;
        b200atmo_destroy(ctx);
        return code;
    };
#define CREATE_TRY(expr)                                                                           \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            ctx->last_error = std::string(#expr) + ": " + cudaGetErrorString(_e);                  \
            return bail(B200ATMO_E_CUDA);                                                          \
        }                                                                                          \
    } while (0)
    CREATE_TRY(cudaMalloc(&ctx->d_lut, sizeof(float) * kLut * kLut));
    CREATE_TRY(cudaMalloc(&ctx->d_lut_pad, sizeof(float) * kLutPad * kLutPad));
    CREATE_TRY(cudaMalloc(&ctx->d_lut_cells, sizeof(float4) * kLutCells * kLutCells));
    CREATE_TRY(cudaStreamCreateWithFlags(&ctx->streams[0], cudaStreamNonBlocking));
    CREATE_TRY(cudaStreamCreateWithFlags(&ctx->streams[1], cudaStreamNonBlocking));
    for (auto& sl : ctx->slots) CREATE_TRY(cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking));
    CREATE_TRY(cudaMalloc(&ctx->d_block_counters, sizeof(unsigned) * (b200atmo_ctx::kBlockCounters + 1)));
    CREATE_TRY(cudaMemset(ctx->d_block_counters, 0, sizeof(unsigned) * (b200atmo_ctx::kBlockCounters + 1)));
#undef CREATE_TRY
    // unset samplers read as white (README.md:46: "by default they cover the whole atmosphere uniformly")
    const uint8_t white[6] = {255, 255, 255, 255, 255, 255};
    int rc = upload_cube(ctx, white, 1, ctx->streams[0]);
    if (rc == B200ATMO_OK) rc = upload_shape(ctx, white, 1, 1, 1, ctx->streams[0]);
    if (rc != B200ATMO_OK) return bail(rc);
    *out = ctx;
    return B200ATMO_OK;
}

void b200atmo_destroy(b200atmo_ctx* ctx) {
    if (!ctx) return;
    DeviceGuard g(ctx->device);
    for (auto& sl : ctx->slots) {
        if (sl.stream) {
            cudaStreamSynchronize(sl.stream);
            cudaStreamDestroy(sl.stream);
        }
        cudaFree(sl.d_depth);
        cudaFree(sl.d_rgba);
        cudaFree(sl.d_disc);
    }
    for (auto& ts : ctx->tables) {
        for (auto& e : ts.e) cudaFree(e.d);
        for (auto& o : ts.o) cudaFree(o.d_cost);
    }
    cudaFree(ctx->d_block_counters);
    cudaFree(ctx->d_lut);
    cudaFree(ctx->d_lut_pad);
    cudaFree(ctx->d_lut_cells);
    cudaFree(ctx->d_cube_pad_u8);
    cudaFree(ctx->d_cube_cells);
    cudaFree(ctx->d_shape_cells);
    cudaFree(ctx->d_blue);
    cudaFree(ctx->d_stage_in0);
    cudaFree(ctx->d_stage_in1);
    cudaFree(ctx->d_stage_out);
    cudaFree(ctx->d_stage_disc);
    for (auto& s : ctx->streams)
        if (s) cudaStreamDestroy(s);
    delete ctx;
}

const char* b200atmo_last_error(const b200atmo_ctx* ctx) { return ctx ? ctx->last_error.c_str() : g_create_error.c_str(); }

int b200atmo_set_params(b200atmo_ctx* ctx, const B200AtmoParams* p) {
    if (!ctx || !p) return fail(ctx, B200ATMO_E_INVALID, "b200atmo_set_params: NULL argument");
    // planet_atmosphere.gd:79-81,217-218,237-238,252-253: R, H and u_density invalidate the baked LUT
    if (p->planet_radius != ctx->params.planet_radius || p->atmosphere_height != ctx->params.atmosphere_height ||
        p->density != ctx->params.density)
        ctx->lut_stale = true;
    ctx->params = *p;
    return B200ATMO_OK;
}

int b200atmo_get_params(const b200atmo_ctx* ctx, B200AtmoParams* out) {
    if (!ctx || !out) return B200ATMO_E_INVALID;
    *out = ctx->params;
    return B200ATMO_OK;
}

int b200atmo_set_variant(b200atmo_ctx* ctx, int scatter_model, int scatter_steps, int cloud_steps, int light_mode) {
    if (!ctx) return B200ATMO_E_INVALID;
    if (scatter_model != B200ATMO_SCATTER_V2 && scatter_model != B200ATMO_SCATTER_V1)
        return fail(ctx, B200ATMO_E_INVALID, "b200atmo_set_variant: unknown scatter model");
    if (scatter_steps < 1 || scatter_steps > 65536 || cloud_steps > 65536)
        return fail(ctx, B200ATMO_E_INVALID, "b200atmo_set_variant: step counts must be in [1, 65536]");
    if (light_mode < B200ATMO_LIGHT_NONE || light_mode > B200ATMO_LIGHT_RAYMARCHED)
        return fail(ctx, B200ATMO_E_INVALID, "b200atmo_set_variant: unknown light mode");
    if (light_mode != B200ATMO_LIGHT_NONE && cloud_steps < 1)
        return fail(ctx, B200ATMO_E_INVALID, "b200atmo_set_variant: cloud_steps must be >= 1 when clouds are enabled");
    ctx->variant.scatter_model = scatter_model;
    ctx->variant.scatter_steps = scatter_steps;
    ctx->variant.cloud_steps = cloud_steps;
    ctx->variant.light_mode = light_mode;
    return B200ATMO_OK;
}

int b200atmo_upload_blue_noise(b200atmo_ctx* ctx, const uint8_t* h_texels, int w, int h) {
    if (!ctx || !h_texels) return fail(ctx, B200ATMO_E_INVALID, "b200atmo_upload_blue_noise: NULL argument");
    // main:168-169 fetches texel (x & 0xff, y & 0xff): the top-left 256 x 256 window of whatever is bound
    if (w < 256 || h < 256 || w > 4096 || h > 4096)
        return fail(ctx, B200ATMO_E_INVALID, "b200atmo_upload_blue_noise: sizes must be in [256, 4096] (the shader reads texel (x & 0xff, y & 0xff))");
    DeviceGuard g(ctx->device);
    DevBuf d;
    CU_TRY(ctx, d.alloc(size_t(w) * h));
    CU_TRY(ctx, cudaMemcpy(d.p, h_texels, size_t(w) * h, cudaMemcpyHostToDevice));
    cudaFree(ctx->d_blue);
    ctx->d_blue = d.release<uint8_t>();
    ctx->bn_w = w;
    ctx->bn_h = h;
    return B200ATMO_OK;
}

int b200atmo_upload_shape3d(b200atmo_ctx* ctx, const uint8_t* h_texels, int nx, int ny, int nz) {
    if (!ctx || !h_texels) return fail(ctx, B200ATMO_E_INVALID, "b200atmo_upload_shape3d: NULL argument");
    if (nx < 1 || ny < 1 || nz < 1 || nx > 512 || ny > 512 || nz > 512)
        return fail(ctx, B200ATMO_E_INVALID, "b200atmo_upload_shape3d: each dimension must be in [1, 512]");
    DeviceGuard g(ctx->device);
    return upload_shape(ctx, h_texels, nx, ny, nz, ctx->streams[0]);
}

int b200atmo_upload_coverage_cube(b200atmo_ctx* ctx, const uint8_t* h_faces6, int res) {
    if (!ctx || !h_faces6) return fail(ctx, B200ATMO_E_INVALID, "b200atmo_upload_coverage_cube: NULL argument");
    if (res < 1 || res > 4096)  // noise_cubemap.gd:30 clamps the resolution to [1, 4096]
        return fail(ctx, B200ATMO_E_INVALID, "b200atmo_upload_coverage_cube: res must be in [1, 4096]");
    DeviceGuard g(ctx->device);
    return upload_cube(ctx, h_faces6, res, ctx->streams[0]);
}

int b200atmo_generate_noise_cubemap(b200atmo_ctx* ctx, const B200AtmoNoise* noise, int res, const float scale[3],
                                    uint8_t* h_faces6_out, int set_as_coverage) {
    if (!ctx || !noise || !scale) return fail(ctx, B200ATMO_E_INVALID, "b200atmo_generate_noise_cubemap: NULL argument");
    if (res < 1 || res > 4096) return fail(ctx, B200ATMO_E_INVALID, "b200atmo_generate_noise_cubemap: res must be in [1, 4096]");
    if (noise->octaves < 1 || noise->octaves > 32) return fail(ctx, B200ATMO_E_INVALID, "b200atmo_generate_noise_cubemap: octaves must be in [1, 32]");
    DeviceGuard g(ctx->device);
    cudaStream_t s = ctx->streams[0];
    const size_t raw = size_t(6) * res * res;
    DevBuf d_faces;
    CU_TRY(ctx, d_faces.alloc(raw));
    CU_TRY(ctx, launch_noise_cube(*noise, scale, res, d_faces.as<uint8_t>(), s));
    ctx->launches++;
    if (h_faces6_out) CU_TRY(ctx, cudaMemcpyAsync(h_faces6_out, d_faces.p, raw, cudaMemcpyDeviceToHost, s));
    CU_TRY(ctx, cudaStreamSynchronize(s));
    return set_as_coverage ? upload_cube(ctx, nullptr, res, s, d_faces.as<uint8_t>()) : B200ATMO_OK;
}

int b200atmo_bake_optical_depth(b200atmo_ctx* ctx, void* stream) {
    if (!ctx) return B200ATMO_E_INVALID;
    DeviceGuard g(ctx->device);
    ctx->lut_stale = true;
    return bake_if_stale(ctx, static_cast<cudaStream_t>(stream));
}

int b200atmo_download_lut(b200atmo_ctx* ctx, float* h_lut) {
    if (!ctx || !h_lut) return fail(ctx, B200ATMO_E_INVALID, "b200atmo_download_lut: NULL argument");
    DeviceGuard g(ctx->device);
    int rc = bake_if_stale(ctx, ctx->streams[0]);
    if (rc != B200ATMO_OK) return rc;
    CU_TRY(ctx, cudaMemcpyAsync(h_lut, ctx->d_lut, sizeof(float) * kLut * kLut, cudaMemcpyDeviceToHost, ctx->streams[0]));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->streams[0]));
    return B200ATMO_OK;
}

int b200atmo_download_cube_padded(b200atmo_ctx* ctx, uint8_t* h_out, size_t cap, int* out_res) {
    if (!ctx || !h_out) return fail(ctx, B200ATMO_E_INVALID, "b200atmo_download_cube_padded: NULL argument");
    const size_t need = size_t(6) * (ctx->cube_res + 2) * (ctx->cube_res + 2);
    if (cap < need) return fail(ctx, B200ATMO_E_INVALID, "b200atmo_download_cube_padded: buffer too small");
    DeviceGuard g(ctx->device);
    CU_TRY(ctx, cudaMemcpy(h_out, ctx->d_cube_pad_u8, need, cudaMemcpyDeviceToHost));
    if (out_res) *out_res = ctx->cube_res;
    return B200ATMO_OK;
}

static int render_rays_impl(b200atmo_ctx* ctx, const B200AtmoFrame* frame, const float* d_origin_depth, const float* d_dir_jitter,
                            size_t n_rays, int width, int height, float* d_rgba, uint8_t* d_discard, void* stream) {
    if (!ctx || !frame) return fail(ctx, B200ATMO_E_INVALID, "b200atmo_render_rays: NULL ctx/frame");
    if (n_rays == 0) return B200ATMO_OK;
    if (!d_origin_depth || !d_dir_jitter || !d_rgba) return fail(ctx, B200ATMO_E_INVALID, "b200atmo_render_rays: NULL buffer");
    if (n_rays > (size_t(1) << 38)) return fail(ctx, B200ATMO_E_INVALID, "b200atmo_render_rays: n_rays too large");
    DeviceGuard g(ctx->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int rc = bake_if_stale(ctx, s);
    if (rc != B200ATMO_OK) return rc;
    DevConsts c;
    consts_from_params(c, ctx->params, ctx->variant, textures_of(ctx));
    consts_set_frame(c, ctx->params, frame->planet_center_view, frame->sun_center_view, frame->inv_view);
    c.fw = width;    // > 0: the batch is a width x height pixel grid, warps cover 8x4 tiles (launch_rays_t)
    c.fh = height;
    RayIO io{};
    io.origin_depth = d_origin_depth;
    io.dir_jitter = d_dir_jitter;
    io.rgba = d_rgba;
    io.discard = d_discard;
    io.n = n_rays;
    b200atmo_ctx::OrderEntry* oe = width > 0 ? block_order_begin(ctx, s, kOrderRays2D, c, rays2d_grid_blocks(width, height), io) : nullptr;
    CU_TRY(ctx, launch_render_rays(c, io, ctx->variant.scatter_model, ctx->variant.light_mode, s));
    ctx->launches++;
    return block_order_finish(ctx, oe, s);
}

int b200atmo_render_rays(b200atmo_ctx* ctx, const B200AtmoFrame* frame, const float* d_origin_depth, const float* d_dir_jitter,
                         size_t n_rays, float* d_rgba, uint8_t* d_discard, void* stream) {
    return render_rays_impl(ctx, frame, d_origin_depth, d_dir_jitter, n_rays, 0, 0, d_rgba, d_discard, stream);
}

int b200atmo_render_rays_2d(b200atmo_ctx* ctx, const B200AtmoFrame* frame, const float* d_origin_depth, const float* d_dir_jitter,
                            int width, int height, float* d_rgba, uint8_t* d_discard, void* stream) {
    if (width < 0 || height < 0) return fail(ctx, B200ATMO_E_INVALID, "b200atmo_render_rays_2d: negative size");
    return render_rays_impl(ctx, frame, d_origin_depth, d_dir_jitter, size_t(width) * size_t(height), width, height, d_rgba, d_discard,
                            stream);
}

int b200atmo_render_rays_host(b200atmo_ctx* ctx, const B200AtmoFrame* frame, const float* h_origin_depth,
                              const float* h_dir_jitter, size_t n_rays, float* h_rgba, uint8_t* h_discard) {
    if (!ctx || !frame) return fail(ctx, B200ATMO_E_INVALID, "b200atmo_render_rays_host: NULL ctx/frame");
    if (n_rays == 0) return B200ATMO_OK;
    if (!h_origin_depth || !h_dir_jitter || !h_rgba) return fail(ctx, B200ATMO_E_INVALID, "b200atmo_render_rays_host: NULL buffer");
    DeviceGuard g(ctx->device);
    const size_t bytes = n_rays * 4 * sizeof(float);
    int rc;
    if ((rc = ensure(ctx, &ctx->d_stage_in0, &ctx->cap_in0, bytes)) != B200ATMO_OK) return rc;
    if ((rc = ensure(ctx, &ctx->d_stage_in1, &ctx->cap_in1, bytes)) != B200ATMO_OK) return rc;
    if ((rc = ensure(ctx, &ctx->d_stage_out, &ctx->cap_out, bytes)) != B200ATMO_OK) return rc;
    if (h_discard && (rc = ensure(ctx, reinterpret_cast<void**>(&ctx->d_stage_disc), &ctx->cap_disc, n_rays)) != B200ATMO_OK) return rc;
    // chunked, double-buffered over two streams so H2D, compute and D2H overlap (PCIe is full duplex)
    const size_t chunk = (n_rays + 7) / 8 < 65536 ? n_rays : (n_rays + 7) / 8;
    int k = 0;
    for (size_t b = 0; b < n_rays; b += chunk, ++k) {
        const size_t m = (n_rays - b < chunk) ? n_rays - b : chunk;
        cudaStream_t s = ctx->streams[k & 1];
        float* in0 = static_cast<float*>(ctx->d_stage_in0) + 4 * b;
        float* in1 = static_cast<float*>(ctx->d_stage_in1) + 4 * b;
        float* out = static_cast<float*>(ctx->d_stage_out) + 4 * b;
        CU_TRY(ctx, cudaMemcpyAsync(in0, h_origin_depth + 4 * b, m * 16, cudaMemcpyHostToDevice, s));
        CU_TRY(ctx, cudaMemcpyAsync(in1, h_dir_jitter + 4 * b, m * 16, cudaMemcpyHostToDevice, s));
        rc = b200atmo_render_rays(ctx, frame, in0, in1, m, out, h_discard ? ctx->d_stage_disc + b : nullptr, s);
        if (rc != B200ATMO_OK) return rc;
        CU_TRY(ctx, cudaMemcpyAsync(h_rgba + 4 * b, out, m * 16, cudaMemcpyDeviceToHost, s));
        if (h_discard) CU_TRY(ctx, cudaMemcpyAsync(h_discard + b, ctx->d_stage_disc + b, m, cudaMemcpyDeviceToHost, s));
    }
    CU_TRY(ctx, cudaStreamSynchronize(ctx->streams[0]));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->streams[1]));
    return B200ATMO_OK;
}

static int frame_consts(b200atmo_ctx* ctx, const B200AtmoCamera* cam, int w, int h, int row_begin, int row_end, DevConsts& c) {
    if (w < 1 || h < 1 || row_begin < 0 || row_end > h || row_begin > row_end)
        return fail(ctx, B200ATMO_E_INVALID, "frame: bad size / row range");
    consts_from_params(c, ctx->params, ctx->variant, textures_of(ctx));
    consts_set_camera(c, ctx->params, *cam, w, h, row_begin, row_end);
    return B200ATMO_OK;
}

// Per-column / per-row tables of the frame front end (make_ray), from the per-stream cache (see b200atmo_ctx::tables).
// Never synchronises: a miss allocates (first use of an entry, or a larger frame) and launches ray_tables_kernel on `s`,
// ahead of the frame kernel that reads it on the same stream.
static int frame_tables(b200atmo_ctx* ctx, const DevConsts& c, RayIO& io, cudaStream_t s) {
    b200atmo_ctx::TableStream* ts = nullptr;
    for (auto& t : ctx->tables)
        if (t.used && t.stream == s) { ts = &t; break; }
    if (!ts)
        for (auto& t : ctx->tables)
            if (!t.used) { ts = &t; t.used = true; t.stream = s; break; }
    if (!ts) {   // more caller streams than cache rows: inline path (same arithmetic per pixel)
        io.ray_col = nullptr;
        io.ray_row = nullptr;
        return B200ATMO_OK;
    }
    b200atmo_ctx::TableEntry* hit = nullptr;
    b200atmo_ctx::TableEntry* lru = &ts->e[0];
    for (auto& e : ts->e) {
        if (e.d && e.w == c.fw && e.h == c.fh && std::memcmp(e.inv_proj, c.inv_proj, sizeof(e.inv_proj)) == 0) { hit = &e; break; }
        if (e.tick < lru->tick) lru = &e;
    }
    if (!hit) {
        const size_t need = size_t(c.fw + c.fh) * sizeof(float4);
        if (lru->cap < need) {
            // stream-ordered free + allocate: kernels already queued on `s` may still read the old buffer
            if (lru->d) CU_TRY(ctx, cudaFreeAsync(lru->d, s));
            lru->d = nullptr;
            lru->cap = 0;
            CU_TRY(ctx, cudaMallocAsync(reinterpret_cast<void**>(&lru->d), need, s));
            lru->cap = need;
        }
        lru->w = 0;
        CU_TRY(ctx, launch_ray_tables(c, lru->d, lru->d + c.fw, s));
        ctx->launches++;
        ctx->table_builds++;
        lru->w = c.fw;
        lru->h = c.fh;
        std::memcpy(lru->inv_proj, c.inv_proj, sizeof(lru->inv_proj));
        hit = lru;
    }
    hit->tick = ++ctx->table_tick;
    io.ray_col = hit->d;
    io.ray_row = hit->d + c.fw;
    return B200ATMO_OK;
}

// Heaviest-first block dispatch for a raymarched-cloud launch of `n` blocks with this geometry on stream `s`: hands the kernel
// the cost slots to fill and, if an earlier launch with the same geometry filled them, the order sorted from those. Returns
// the entry (call block_order_finish after the launch) or null (not a raymarched launch, switched off, no cache row).
static b200atmo_ctx::OrderEntry* block_order_begin(b200atmo_ctx* ctx, cudaStream_t s, int kind, const DevConsts& c, unsigned n, RayIO& io) {
    io.block_order = nullptr;
    io.block_cost = nullptr;
    const bool wanted = ctx->variant.light_mode == B200ATMO_LIGHT_RAYMARCHED ||
                        (ctx->block_order_enabled == 2 && ctx->variant.light_mode == B200ATMO_LIGHT_CHEAP);
    if (!ctx->block_order_enabled || !wanted || n < 2u) return nullptr;
    b200atmo_ctx::TableStream* ts = nullptr;
    for (auto& t : ctx->tables)
        if (t.used && t.stream == s) { ts = &t; break; }
    if (!ts)
        for (auto& t : ctx->tables)
            if (!t.used) { ts = &t; t.used = true; t.stream = s; break; }
    if (!ts) return nullptr;
    b200atmo_ctx::OrderEntry* hit = nullptr;
    b200atmo_ctx::OrderEntry* lru = &ts->o[0];
    for (auto& o : ts->o) {
        if (o.d_cost && o.kind == kind && o.n == n && o.w == c.fw && o.h == c.fh && o.row_begin == c.row_begin && o.row_end == c.row_end &&
            o.row_pitch == c.row_pitch && o.cloud_steps == c.cloud_steps) { hit = &o; break; }
        if (o.tick < lru->tick) lru = &o;
    }
    if (!hit) {
        const size_t need = size_t(n) * 2 * sizeof(unsigned);
        if (lru->cap < need) {
            if (lru->d_cost && cudaFreeAsync(lru->d_cost, s) != cudaSuccess) return nullptr;
            lru->d_cost = nullptr;
            lru->cap = 0;
            if (cudaMallocAsync(reinterpret_cast<void**>(&lru->d_cost), need, s) != cudaSuccess) { cudaGetLastError(); return nullptr; }
            lru->cap = need;
        }
        lru->d_order = lru->d_cost + n;
        if (cudaMemsetAsync(lru->d_cost, 0, size_t(n) * sizeof(unsigned), s) != cudaSuccess) return nullptr;
        lru->kind = kind; lru->n = n; lru->w = c.fw; lru->h = c.fh; lru->row_begin = c.row_begin; lru->row_end = c.row_end;
        lru->row_pitch = c.row_pitch; lru->cloud_steps = c.cloud_steps;
        lru->valid = false;
        hit = lru;
    }
    hit->tick = ++ctx->table_tick;
    io.block_cost = hit->d_cost;
    io.block_order = hit->valid ? hit->d_order : nullptr;
    return hit;
}
static int block_order_finish(b200atmo_ctx* ctx, b200atmo_ctx::OrderEntry* e, cudaStream_t s) {
    if (!e) return B200ATMO_OK;
    CU_TRY(ctx, launch_block_order(e->d_cost, e->d_order, e->n, s));
    e->valid = true;
    ctx->launches++;
    return B200ATMO_OK;
}

static bool valid_format(int f) { return f == B200ATMO_COLOR_RGBA32F || f == B200ATMO_COLOR_RGBA16F; }
static size_t format_bytes(int f) { return f == B200ATMO_COLOR_RGBA16F ? 8 : 16; }

// one frame-kernel launch for rows [c.row_begin, c.row_end) in either result format
static cudaError_t launch_frame_any(const DevConsts& c, const RayIO& io, int rgba_format, const Variant& v, cudaStream_t s) {
    if (rgba_format == B200ATMO_COLOR_RGBA16F && !io.color_inout) {
        RayIO16 io16;
        static_cast<RayIO&>(io16) = io;
        return launch_render_frame16(c, io16, v.scatter_model, v.light_mode, s);
    }
    return launch_render_frame(c, io, v.scatter_model, v.light_mode, s);
}

int b200atmo_render_frame_fmt(b200atmo_ctx* ctx, const B200AtmoCamera* cam, const float* d_depth, int w, int h, int row_begin,
                              int row_end, void* d_rgba, int rgba_format, uint8_t* d_discard, void* stream) {
    if (!ctx || !cam || !d_depth || !d_rgba) return fail(ctx, B200ATMO_E_INVALID, "b200atmo_render_frame: NULL argument");
    if (!valid_format(rgba_format)) return fail(ctx, B200ATMO_E_INVALID, "b200atmo_render_frame: unknown result format");
    DeviceGuard g(ctx->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    DevConsts c;
    int rc = frame_consts(ctx, cam, w, h, row_begin, row_end, c);
    if (rc != B200ATMO_OK) return rc;
    if (row_begin == row_end) return B200ATMO_OK;
    if ((rc = bake_if_stale(ctx, s)) != B200ATMO_OK) return rc;   // LUT buffers are allocated once: pointers in c stay valid
    RayIO io{};
    io.depth = d_depth;
    io.rgba = d_rgba;
    io.discard = d_discard;
    io.n = size_t(w) * h;
    if ((rc = frame_tables(ctx, c, io, s)) != B200ATMO_OK) return rc;
    b200atmo_ctx::OrderEntry* oe = block_order_begin(ctx, s, kOrderFrame, c, frame_grid_blocks(c), io);
    CU_TRY(ctx, launch_frame_any(c, io, rgba_format, ctx->variant, s));
    ctx->launches++;
    return block_order_finish(ctx, oe, s);
}

int b200atmo_render_frame(b200atmo_ctx* ctx, const B200AtmoCamera* cam, const float* d_depth, int w, int h, int row_begin,
                          int row_end, float* d_rgba, uint8_t* d_discard, void* stream) {
    return b200atmo_render_frame_fmt(ctx, cam, d_depth, w, h, row_begin, row_end, d_rgba, B200ATMO_COLOR_RGBA32F, d_discard, stream);
}

int b200atmo_render_frame_composite_fmt(b200atmo_ctx* ctx, const B200AtmoCamera* cam, const float* d_depth, int w, int h,
                                        int row_begin, int row_end, void* d_color_inout, int color_format, void* stream) {
    if (!ctx || !cam || !d_depth || !d_color_inout)
        return fail(ctx, B200ATMO_E_INVALID, "b200atmo_render_frame_composite: NULL argument");
    if (color_format != B200ATMO_COLOR_RGBA32F && color_format != B200ATMO_COLOR_RGBA16F)
        return fail(ctx, B200ATMO_E_INVALID, "b200atmo_render_frame_composite: unknown colour format");
    DeviceGuard g(ctx->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int rc = bake_if_stale(ctx, s);
    if (rc != B200ATMO_OK) return rc;
    DevConsts c;
    if ((rc = frame_consts(ctx, cam, w, h, row_begin, row_end, c)) != B200ATMO_OK) return rc;
    if (row_begin == row_end) return B200ATMO_OK;
    RayIO io{};
    io.depth = d_depth;
    io.color_inout = d_color_inout;
    io.color_format = color_format;
    io.n = size_t(w) * h;
    if ((rc = frame_tables(ctx, c, io, s)) != B200ATMO_OK) return rc;
    CU_TRY(ctx, launch_render_frame(c, io, ctx->variant.scatter_model, ctx->variant.light_mode, s));
    ctx->launches++;
    return B200ATMO_OK;
}

int b200atmo_render_frame_composite(b200atmo_ctx* ctx, const B200AtmoCamera* cam, const float* d_depth, int w, int h,
                                    int row_begin, int row_end, float* d_color_inout, void* stream) {
    return b200atmo_render_frame_composite_fmt(ctx, cam, d_depth, w, h, row_begin, row_end, d_color_inout, B200ATMO_COLOR_RGBA32F,
                                               stream);
}

int b200atmo_composite_frame_host(b200atmo_ctx* ctx, const B200AtmoCamera* cam, const float* h_depth, int w, int h,
                                  void* h_color_inout, int color_format) {
    if (!ctx || !cam || !h_depth || !h_color_inout) return fail(ctx, B200ATMO_E_INVALID, "b200atmo_composite_frame_host: NULL argument");
    if (w < 1 || h < 1) return fail(ctx, B200ATMO_E_INVALID, "b200atmo_composite_frame_host: bad size");
    if (color_format != B200ATMO_COLOR_RGBA32F && color_format != B200ATMO_COLOR_RGBA16F)
        return fail(ctx, B200ATMO_E_INVALID, "b200atmo_composite_frame_host: unknown colour format");
    DeviceGuard g(ctx->device);
    const size_t npx = size_t(w) * h, px_bytes = color_format == B200ATMO_COLOR_RGBA16F ? 8 : 16;
    int rc;
    if ((rc = ensure(ctx, &ctx->d_stage_in0, &ctx->cap_in0, npx * sizeof(float))) != B200ATMO_OK) return rc;
    if ((rc = ensure(ctx, &ctx->d_stage_out, &ctx->cap_out, npx * px_bytes)) != B200ATMO_OK) return rc;
    if ((rc = bake_if_stale(ctx, ctx->streams[0])) != B200ATMO_OK) return rc;
    float* d_depth = static_cast<float*>(ctx->d_stage_in0);
    char* d_color = static_cast<char*>(ctx->d_stage_out);
    char* h_color = static_cast<char*>(h_color_inout);
    DevConsts c;
    if ((rc = frame_consts(ctx, cam, w, h, 0, h, c)) != B200ATMO_OK) return rc;
    RayIO io{};
    io.depth = d_depth;
    io.color_inout = d_color;
    io.color_format = color_format;
    io.n = npx;
    // equal row bands alternating over two streams: the uploads of band k+1 (12 or 20 B/pixel) run while band k downloads
    // (8 or 16 B/pixel); PCIe is full duplex, so the slower direction bounds the frame
    const int bands = h >= 512 ? 8 : (h >= 256 ? 4 : 1);   // the upload is the longer leg here: finer bands shorten the tail
    for (int k = 0; k < bands; ++k) {
        const int r0 = int(int64_t(h) * k / bands), r1 = int(int64_t(h) * (k + 1) / bands);
        if (r1 <= r0) continue;
        cudaStream_t s = ctx->streams[k & 1];
        const size_t off = size_t(r0) * w, cnt = size_t(r1 - r0) * w;
        CU_TRY(ctx, cudaMemcpyAsync(d_depth + off, h_depth + off, cnt * sizeof(float), cudaMemcpyHostToDevice, s));
        CU_TRY(ctx, cudaMemcpyAsync(d_color + off * px_bytes, h_color + off * px_bytes, cnt * px_bytes, cudaMemcpyHostToDevice, s));
        c.row_begin = r0;
        c.row_end = r1;
        if ((rc = frame_tables(ctx, c, io, s)) != B200ATMO_OK) return rc;   // per-stream cache: each band stream has its own copy
        CU_TRY(ctx, launch_render_frame(c, io, ctx->variant.scatter_model, ctx->variant.light_mode, s));
        ctx->launches++;
        CU_TRY(ctx, cudaMemcpyAsync(h_color + off * px_bytes, d_color + off * px_bytes, cnt * px_bytes, cudaMemcpyDeviceToHost, s));
    }
    CU_TRY(ctx, cudaStreamSynchronize(ctx->streams[0]));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->streams[1]));
    return B200ATMO_OK;
}

int b200atmo_make_rays(b200atmo_ctx* ctx, const B200AtmoCamera* cam, const float* d_depth, int w, int h, float* d_origin_depth,
                       float* d_dir_jitter, B200AtmoFrame* frame_out, void* stream) {
    if (!ctx || !cam || !d_depth || !d_origin_depth || !d_dir_jitter)
        return fail(ctx, B200ATMO_E_INVALID, "b200atmo_make_rays: NULL argument");
    DeviceGuard g(ctx->device);
    DevConsts c;
    int rc = frame_consts(ctx, cam, w, h, 0, h, c);
    if (rc != B200ATMO_OK) return rc;
    RayIO io{};
    io.depth = d_depth;
    io.out_origin_depth = d_origin_depth;
    io.out_dir_jitter = d_dir_jitter;
    io.n = size_t(w) * h;
    if ((rc = frame_tables(ctx, c, io, static_cast<cudaStream_t>(stream))) != B200ATMO_OK) return rc;
    CU_TRY(ctx, launch_make_rays(c, io, static_cast<cudaStream_t>(stream)));
    ctx->launches++;
    if (frame_out) {
        // the varyings of atmosphere_vertex (main:101-103) + the (fixed-up) INV_VIEW_MATRIX
        float world_pos[4], pc[4], sc[4];
        hostmath::mat4_mul_vec(cam->model, 0.0f, 0.0f, 0.0f, 1.0f, world_pos);
        hostmath::mat4_mul_vec(cam->view, world_pos[0], world_pos[1], world_pos[2], world_pos[3], pc);
        hostmath::mat4_mul_vec(cam->view, ctx->params.sun_position[0], ctx->params.sun_position[1], ctx->params.sun_position[2],
                               1.0f, sc);
        for (int k = 0; k < 3; ++k) {
            frame_out->planet_center_view[k] = pc[k];
            frame_out->sun_center_view[k] = sc[k];
        }
        std::memcpy(frame_out->inv_view, c.inv_view_ray, sizeof(frame_out->inv_view));
    }
    return B200ATMO_OK;
}

int b200atmo_render_frame_host_fmt(b200atmo_ctx* ctx, const B200AtmoCamera* cam, const float* h_depth, int w, int h, void* h_rgba,
                                   int rgba_format, uint8_t* h_discard) {
    if (!ctx || !cam || !h_depth || !h_rgba) return fail(ctx, B200ATMO_E_INVALID, "b200atmo_render_frame_host: NULL argument");
    if (w < 1 || h < 1) return fail(ctx, B200ATMO_E_INVALID, "b200atmo_render_frame_host: bad size");
    if (!valid_format(rgba_format)) return fail(ctx, B200ATMO_E_INVALID, "b200atmo_render_frame_host: unknown result format");
    DeviceGuard g(ctx->device);
    const size_t npx = size_t(w) * h, pxb = format_bytes(rgba_format);
    int rc;
    if ((rc = ensure(ctx, &ctx->d_stage_in0, &ctx->cap_in0, npx * sizeof(float))) != B200ATMO_OK) return rc;
    if ((rc = ensure(ctx, &ctx->d_stage_out, &ctx->cap_out, npx * pxb)) != B200ATMO_OK) return rc;
    if (h_discard && (rc = ensure(ctx, reinterpret_cast<void**>(&ctx->d_stage_disc), &ctx->cap_disc, npx)) != B200ATMO_OK) return rc;
    if ((rc = bake_if_stale(ctx, ctx->streams[0])) != B200ATMO_OK) return rc;
    float* d_depth = static_cast<float*>(ctx->d_stage_in0);
    char* d_rgba = static_cast<char*>(ctx->d_stage_out);
    char* h_out = static_cast<char*>(h_rgba);
    DevConsts c;
    if ((rc = frame_consts(ctx, cam, w, h, 0, h, c)) != B200ATMO_OK) return rc;
    RayIO io{};
    io.depth = d_depth;
    io.rgba = d_rgba;
    io.discard = h_discard ? ctx->d_stage_disc : nullptr;
    io.n = npx;
    // Row bands, alternating over two streams: H2D(depth band) -> kernel(band) -> D2H(rgba band). The D2H (16 or 8 B/px)
    // is the PCIe-bound leg, so the first band is small (the D2H engine starts early) and bands grow by
    // ~1.5x: each band's upload + kernel hides behind the previous band's download. Few bands: the host issues
    // ~3 API calls per band at ~5 us each.
    int bands = h >= 256 ? 4 : 1;   // measured on B200: 3-5 bands are equivalent (0.72-0.75 ms at 1080p, PCIe floor 0.61 ms)
    if (const char* e = std::getenv("B200ATMO_E2E_BANDS")) {  // experiment knob
        const int v = std::atoi(e);
        if (v >= 1 && v <= h) bands = v;
    }
    const double r = 1.5;
    double total = 0.0, acc = 0.0, pw = 1.0;
    for (int k = 0; k < bands; ++k, pw *= r) total += pw;
    pw = 1.0;
    int r0 = 0;
    for (int k = 0; k < bands; ++k, pw *= r) {
        acc += pw;
        const int r1 = (k == bands - 1) ? h : int(double(h) * acc / total);
        if (r1 <= r0) continue;
        cudaStream_t s = ctx->streams[k & 1];
        const size_t off = size_t(r0) * w, cnt = size_t(r1 - r0) * w;
        CU_TRY(ctx, cudaMemcpyAsync(d_depth + off, h_depth + off, cnt * sizeof(float), cudaMemcpyHostToDevice, s));
        c.row_begin = r0;
        c.row_end = r1;
        if ((rc = frame_tables(ctx, c, io, s)) != B200ATMO_OK) return rc;   // per-stream cache: each band stream has its own copy
        b200atmo_ctx::OrderEntry* oe = block_order_begin(ctx, s, kOrderFrame, c, frame_grid_blocks(c), io);
        CU_TRY(ctx, launch_frame_any(c, io, rgba_format, ctx->variant, s));
        ctx->launches++;
        if ((rc = block_order_finish(ctx, oe, s)) != B200ATMO_OK) return rc;
        CU_TRY(ctx, cudaMemcpyAsync(h_out + off * pxb, d_rgba + off * pxb, cnt * pxb, cudaMemcpyDeviceToHost, s));
        if (h_discard) CU_TRY(ctx, cudaMemcpyAsync(h_discard + off, ctx->d_stage_disc + off, cnt, cudaMemcpyDeviceToHost, s));
        r0 = r1;
    }
    CU_TRY(ctx, cudaStreamSynchronize(ctx->streams[0]));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->streams[1]));
    return B200ATMO_OK;
}

int b200atmo_render_frame_host(b200atmo_ctx* ctx, const B200AtmoCamera* cam, const float* h_depth, int w, int h, float* h_rgba,
                               uint8_t* h_discard) {
    return b200atmo_render_frame_host_fmt(ctx, cam, h_depth, w, h, h_rgba, B200ATMO_COLOR_RGBA32F, h_discard);
}

static int peers_to_io(b200atmo_ctx* ctx, const B200AtmoPeerTargets* t, RayIOPeers& io, const char* who) {
    if (!t) return fail(ctx, B200ATMO_E_INVALID, std::string(who) + ": NULL targets");
    if (t->n_peers < 1 || t->n_peers > B200ATMO_MAX_PEERS) return fail(ctx, B200ATMO_E_INVALID, std::string(who) + ": n_peers out of range");
    for (int r = 0; r < t->n_peers; ++r) {
        if (!t->d_rgba_peers[r]) return fail(ctx, B200ATMO_E_INVALID, std::string(who) + ": NULL peer buffer");
        io.rgba_peers[r] = t->d_rgba_peers[r];
    }
    io.n_peers = t->n_peers;
    io.first_peer = (t->first_peer >= 0 && t->first_peer < t->n_peers) ? t->first_peer : 0;
    io.use_tma = t->use_tma != 0;
    io.rgba_multicast = t->d_rgba_multicast;
    io.peer_offset = size_t(t->elem_offset);
    if (!valid_format(t->rgba_format)) return fail(ctx, B200ATMO_E_INVALID, std::string(who) + ": unknown tile format");
    io.rgba_half = t->rgba_format == B200ATMO_COLOR_RGBA16F ? 1 : 0;
    if (io.use_tma && io.rgba_half && (t->elem_offset & 1u))   // cp.async.bulk needs 16-byte aligned destinations
        return fail(ctx, B200ATMO_E_INVALID, std::string(who) + ": TMA stores of half4 tiles need an even elem_offset");
    const B200AtmoPeerSync& y = t->sync;
    if (y.n_done_flags < 0 || y.n_done_flags > B200ATMO_MAX_PEERS || y.n_consumed_flags < 0 || y.n_consumed_flags > B200ATMO_MAX_PEERS ||
        y.n_credit < 0 || y.n_credit > 32 || y.n_wait < 0 || y.n_wait > 32 || y.done_slot < 0 || y.consumed_slot < 0 ||
        y.credit_first_slot < 0 || y.wait_first_slot < 0)
        return fail(ctx, B200ATMO_E_INVALID, std::string(who) + ": bad hand-shake block");
    if (y.n_done_flags || y.n_consumed_flags || y.n_credit || y.n_wait) {
        if (io.use_tma) return fail(ctx, B200ATMO_E_INVALID, std::string(who) + ": the fused hand-shake is not available with use_tma");
        for (int k = 0; k < y.n_done_flags; ++k)
            if (!y.d_done_flags[k]) return fail(ctx, B200ATMO_E_INVALID, std::string(who) + ": NULL completion-flag array");
        for (int k = 0; k < y.n_consumed_flags; ++k)
            if (!y.d_consumed_flags[k]) return fail(ctx, B200ATMO_E_INVALID, std::string(who) + ": NULL consumed-flag array");
        if ((y.n_credit && !y.d_credit_flags) || (y.n_wait && !y.d_wait_flags))
            return fail(ctx, B200ATMO_E_INVALID, std::string(who) + ": NULL flag array to wait on");
        io.sync = y;
        io.timeouts = ctx->d_block_counters + b200atmo_ctx::kBlockCounters;
        io.block_counter = ctx->d_block_counters + ctx->next_block_counter;
        ctx->next_block_counter = (ctx->next_block_counter + 1) % b200atmo_ctx::kBlockCounters;
    }
    return B200ATMO_OK;
}

static int render_frame_peers_impl(b200atmo_ctx* ctx, const B200AtmoCamera* cam, const float* d_depth, int w, int h, int row_begin,
                                   int row_end, int row_pitch, const B200AtmoPeerTargets* targets, void* stream) {
    if (!ctx || !cam || !d_depth) return fail(ctx, B200ATMO_E_INVALID, "b200atmo_render_frame_peers: NULL argument");
    DeviceGuard g(ctx->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    RayIOPeers io{};
    int rc = peers_to_io(ctx, targets, io, "b200atmo_render_frame_peers");
    if (rc != B200ATMO_OK) return rc;
    DevConsts c;
    if ((rc = frame_consts(ctx, cam, w, h, row_begin, row_end, c)) != B200ATMO_OK) return rc;
    if (row_begin >= row_end) return B200ATMO_OK;
    c.row_pitch = row_pitch;
    if ((rc = bake_if_stale(ctx, s)) != B200ATMO_OK) return rc;
    io.depth = d_depth;
    io.n = size_t(w) * h;
    if ((rc = frame_tables(ctx, c, io, s)) != B200ATMO_OK) return rc;
    b200atmo_ctx::OrderEntry* oe = block_order_begin(ctx, s, kOrderPeers, c, frame_grid_blocks(c), io);
    CU_TRY(ctx, launch_render_frame_peers(c, io, ctx->variant.scatter_model, ctx->variant.light_mode, s));
    ctx->launches++;
    return block_order_finish(ctx, oe, s);
}

int b200atmo_render_frame_peers(b200atmo_ctx* ctx, const B200AtmoCamera* cam, const float* d_depth, int w, int h, int row_begin,
                                int row_end, const B200AtmoPeerTargets* targets, void* stream) {
    return render_frame_peers_impl(ctx, cam, d_depth, w, h, row_begin, row_end, 8, targets, stream);
}

int b200atmo_render_frame_peers_interleaved(b200atmo_ctx* ctx, const B200AtmoCamera* cam, const float* d_depth, int w, int h,
                                            int first_tile, int tile_pitch, const B200AtmoPeerTargets* targets, void* stream) {
    if (first_tile < 0 || tile_pitch < 1 || first_tile >= tile_pitch)
        return fail(ctx, B200ATMO_E_INVALID, "b200atmo_render_frame_peers_interleaved: need 0 <= first_tile < tile_pitch");
    if (h < 1) return fail(ctx, B200ATMO_E_INVALID, "frame: bad size / row range");
    const int row_begin = first_tile * 8 < h ? first_tile * 8 : h;   // more ranks than row tiles: nothing to do for this one
    return render_frame_peers_impl(ctx, cam, d_depth, w, h, row_begin, h, 8 * tile_pitch, targets, stream);
}

int b200atmo_render_rays_peers(b200atmo_ctx* ctx, const B200AtmoFrame* frame, const float* d_origin_depth, const float* d_dir_jitter,
                               size_t n_rays, const B200AtmoPeerTargets* targets, void* stream) {
    if (!ctx || !frame) return fail(ctx, B200ATMO_E_INVALID, "b200atmo_render_rays_peers: NULL ctx/frame");
    RayIOPeers io{};
    int rc = peers_to_io(ctx, targets, io, "b200atmo_render_rays_peers");
    if (rc != B200ATMO_OK) return rc;
    if (n_rays == 0) return B200ATMO_OK;
    if (!d_origin_depth || !d_dir_jitter) return fail(ctx, B200ATMO_E_INVALID, "b200atmo_render_rays_peers: NULL buffer");
    if (n_rays > (size_t(1) << 38)) return fail(ctx, B200ATMO_E_INVALID, "b200atmo_render_rays_peers: n_rays too large");
    DeviceGuard g(ctx->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if ((rc = bake_if_stale(ctx, s)) != B200ATMO_OK) return rc;
    DevConsts c;
    consts_from_params(c, ctx->params, ctx->variant, textures_of(ctx));
    consts_set_frame(c, ctx->params, frame->planet_center_view, frame->sun_center_view, frame->inv_view);
    io.origin_depth = d_origin_depth;
    io.dir_jitter = d_dir_jitter;
    io.n = n_rays;
    CU_TRY(ctx, launch_render_rays_peers(c, io, ctx->variant.scatter_model, ctx->variant.light_mode, s));
    ctx->launches++;
    return B200ATMO_OK;
}

int b200atmo_render_frame_host_submit_fmt(b200atmo_ctx* ctx, const B200AtmoCamera* cam, const float* h_depth, int w, int h,
                                          void* h_rgba, int rgba_format, uint8_t* h_discard, int slot) {
    if (!ctx || !cam || !h_depth || !h_rgba) return fail(ctx, B200ATMO_E_INVALID, "b200atmo_render_frame_host_submit: NULL argument");
    if (w < 1 || h < 1) return fail(ctx, B200ATMO_E_INVALID, "b200atmo_render_frame_host_submit: bad size");
    if (!valid_format(rgba_format)) return fail(ctx, B200ATMO_E_INVALID, "b200atmo_render_frame_host_submit: unknown result format");
    if (slot < 0 || slot >= B200ATMO_PIPELINE_SLOTS) return fail(ctx, B200ATMO_E_INVALID, "b200atmo_render_frame_host_submit: bad slot");
    b200atmo_ctx::Slot& sl = ctx->slots[slot];
    if (sl.in_flight) return fail(ctx, B200ATMO_E_STATE, "b200atmo_render_frame_host_submit: slot still in flight (call b200atmo_frame_wait)");
    DeviceGuard g(ctx->device);
    const size_t npx = size_t(w) * h, pxb = format_bytes(rgba_format);
    int rc;
    if ((rc = ensure(ctx, &sl.d_depth, &sl.cap_depth, npx * sizeof(float))) != B200ATMO_OK) return rc;
    if ((rc = ensure(ctx, &sl.d_rgba, &sl.cap_rgba, npx * pxb)) != B200ATMO_OK) return rc;
    if (h_discard && (rc = ensure(ctx, &sl.d_disc, &sl.cap_disc, npx)) != B200ATMO_OK) return rc;
    if ((rc = bake_if_stale(ctx, sl.stream)) != B200ATMO_OK) return rc;
    DevConsts c;   // uniforms, variant and camera are captured here (kernel parameter space)
    if ((rc = frame_consts(ctx, cam, w, h, 0, h, c)) != B200ATMO_OK) return rc;
    RayIO io{};
    io.depth = static_cast<float*>(sl.d_depth);
    io.rgba = sl.d_rgba;
    io.discard = h_discard ? static_cast<uint8_t*>(sl.d_disc) : nullptr;
    io.n = npx;
    if ((rc = frame_tables(ctx, c, io, sl.stream)) != B200ATMO_OK) return rc;
    sl.in_flight = true;   // from the first enqueue on the host buffers are in use, also if a later enqueue fails
    CU_TRY(ctx, cudaMemcpyAsync(sl.d_depth, h_depth, npx * sizeof(float), cudaMemcpyHostToDevice, sl.stream));
    b200atmo_ctx::OrderEntry* oe = block_order_begin(ctx, sl.stream, kOrderFrame, c, frame_grid_blocks(c), io);
    CU_TRY(ctx, launch_frame_any(c, io, rgba_format, ctx->variant, sl.stream));
    ctx->launches++;
    if ((rc = block_order_finish(ctx, oe, sl.stream)) != B200ATMO_OK) return rc;
    CU_TRY(ctx, cudaMemcpyAsync(h_rgba, sl.d_rgba, npx * pxb, cudaMemcpyDeviceToHost, sl.stream));
    if (h_discard) CU_TRY(ctx, cudaMemcpyAsync(h_discard, sl.d_disc, npx, cudaMemcpyDeviceToHost, sl.stream));
    return B200ATMO_OK;
}

int b200atmo_render_frame_host_submit(b200atmo_ctx* ctx, const B200AtmoCamera* cam, const float* h_depth, int w, int h,
                                      float* h_rgba, uint8_t* h_discard, int slot) {
    return b200atmo_render_frame_host_submit_fmt(ctx, cam, h_depth, w, h, h_rgba, B200ATMO_COLOR_RGBA32F, h_discard, slot);
}

int b200atmo_frame_wait(b200atmo_ctx* ctx, int slot) {
    if (!ctx) return fail(ctx, B200ATMO_E_INVALID, "b200atmo_frame_wait: NULL context");
    if (slot < 0 || slot >= B200ATMO_PIPELINE_SLOTS) return fail(ctx, B200ATMO_E_INVALID, "b200atmo_frame_wait: bad slot");
    b200atmo_ctx::Slot& sl = ctx->slots[slot];
    if (!sl.in_flight) return B200ATMO_OK;
    DeviceGuard g(ctx->device);
    sl.in_flight = false;
    CU_TRY(ctx, cudaStreamSynchronize(sl.stream));
    return B200ATMO_OK;
}

int b200atmo_peers_wait(b200atmo_ctx* ctx, const void* d_flags, int first_slot, int n_slots, uint32_t epoch, void* stream) {
    if (!ctx || !d_flags) return fail(ctx, B200ATMO_E_INVALID, "b200atmo_peers_wait: NULL argument");
    if (first_slot < 0 || n_slots < 0 || n_slots > 32) return fail(ctx, B200ATMO_E_INVALID, "b200atmo_peers_wait: bad slot range");
    if (n_slots == 0) return B200ATMO_OK;
    DeviceGuard g(ctx->device);
    CU_TRY(ctx, launch_peers_wait(static_cast<const unsigned*>(d_flags) + first_slot, n_slots, epoch,
                                  ctx->d_block_counters + b200atmo_ctx::kBlockCounters, static_cast<cudaStream_t>(stream)));
    ctx->launches++;
    return B200ATMO_OK;
}

int b200atmo_peers_signal(b200atmo_ctx* ctx, void* const* d_flags_peers, int n_peers, int slot, uint32_t epoch, void* stream) {
    if (!ctx || !d_flags_peers) return fail(ctx, B200ATMO_E_INVALID, "b200atmo_peers_signal: NULL argument");
    if (n_peers < 0 || n_peers > B200ATMO_MAX_PEERS || slot < 0) return fail(ctx, B200ATMO_E_INVALID, "b200atmo_peers_signal: bad peer list");
    if (n_peers == 0) return B200ATMO_OK;
    PeerFlagList l{};
    for (int k = 0; k < n_peers; ++k) {
        if (!d_flags_peers[k]) return fail(ctx, B200ATMO_E_INVALID, "b200atmo_peers_signal: NULL flag array");
        l.p[k] = static_cast<unsigned*>(d_flags_peers[k]);
    }
    DeviceGuard g(ctx->device);
    CU_TRY(ctx, launch_peers_signal(l, n_peers, unsigned(slot), epoch, static_cast<cudaStream_t>(stream)));
    ctx->launches++;
    return B200ATMO_OK;
}

int b200atmo_peers_wait_timeouts(b200atmo_ctx* ctx) {
    if (!ctx) return B200ATMO_E_INVALID;
    DeviceGuard g(ctx->device);
    unsigned v = 0;
    CU_TRY(ctx, cudaMemcpy(&v, ctx->d_block_counters + b200atmo_ctx::kBlockCounters, sizeof(v), cudaMemcpyDeviceToHost));
    return int(v & 0x7fffffffu);
}

uint64_t b200atmo_launch_count(const b200atmo_ctx* ctx) { return ctx ? ctx->launches : 0; }
uint64_t b200atmo_table_build_count(const b200atmo_ctx* ctx) { return ctx ? ctx->table_builds : 0; }

}  // extern "C"

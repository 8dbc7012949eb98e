// Internal interface between the C-ABI (atmo_capi.cu) and the kernels (atmo_kernels.cu).
// Not installed; the public surface is include/b200atmo.h.
#pragma once

#ifndef __CUDACC__
struct float4;
#endif
#include <stddef.h>
#include <stdint.h>

#include "../../include/b200atmo.h"

namespace b200atmo {

constexpr int kLut = B200ATMO_LUT_SIZE;      // 256
constexpr int kLutPad = kLut + 2;            // clamp-to-edge apron of one texel on every side
constexpr int kLutCells = kLut + 1;          // bilinear cells between padded texels: 257 x 257

// The constants of the cloud density evaluation, packed into 16-byte aligned quads in the order the evaluation reads them: the
// kernels fetch kernel parameters with LDC / LDCU inside the (register-starved) march loops, and ptxas merges adjacent aligned
// scalars into one 64- or 128-bit load: 17 constant loads per density evaluation become 7 (of ~170 issued instructions).
// Copies of fields of DevConsts (atmo_consts.h: consts_pack_cloud_hot), appended at its END so no other offset moves.
struct alignas(16) CloudHot {
    float sun[3];            // sun_dir_model
    float bottom_h;          // cloud_bottom_h
    float inv_thickness, thickness, hc_min, coverage_bias;
    float rot[4];
    float shape_hi_m01, dens_y_min, shape_scale, shape_factor;
    float shape_mix0, density_scale;
    int shape_invert;
    float light_reach;
    const float4* cube_cells;
    const float4* shape_cells;
    int cube_res, nx, ny, nz;
    float under_r2, over_r2; // squares of two radii safely below / above the band of the shell that can hold cloud (atmo_consts.h: cloud_skip_r2)
    float pad_[2];
};

// Everything a render kernel needs, passed by value as a __grid_constant__ kernel parameter.
// Host code (atmo_consts.h) fills it with plain fp32 arithmetic in the shader's op order, no FMA
// contraction, so per-frame constants are bit-identical to what the shader computes per fragment.
struct DevConsts {
    // --- planet / atmosphere (planet_common, atmosphere_common) ---
    float R, H, rho, atmo_radius;
    float sphere_depth_factor;
    float C[3];            // v_planet_center_viewspace
    float sun_dir[3];      // normalize(v_sun_center_viewspace - v_planet_center_viewspace), main:164
    // --- scattering v2 ---
    float coef[3];         // scattering_coefficients, funcs_v2:47-51
    float neg_coef_log2e[3];
    float ambient[3], modulate[3];
    float inv_H;           // 1/H, correctly rounded
    float rho2;            // rho*rho (density is applied twice, funcs_v2:65)
    const float* lut_pad;  // [kLutPad][kLutPad] fp32
    const float4* lut_cells;  // [kLutCells][kLutCells] bilinear patches expanded about the cell centre (tc, dxc, dyc, dxy)
    // --- scattering v1 ---
    float day0[3], day1[3], night0[3], night1[3];
    float day_night_scale;
    // --- clouds ---
    float cloud_bottom_h, cloud_top_h;   // R + u_cloud_bottom*H, R + u_cloud_top*H  (cloud_funcs:260-261)
    float cloud_thickness;               // top - bottom
    float inv_cloud_thickness;           // 1/(top-bottom), correctly rounded
    float density_scale, cloud_blend, coverage_bias, shape_factor, shape_scale;
    int shape_invert;                    // u_cloud_shape_invert == 1.0
    float rot[4];                        // mat2 column-major
    float v2m[16];                       // view_to_model = u_world_to_model_matrix * inv_view (cloud_funcs:285), column-major
    float sun_dir_model[3];              // (view_to_model * vec4(sun_dir, 0)).xyz, cloud_funcs:288
    float march_space, march_ground;     // cloud_funcs:186-190
    float march_hmin, march_hmax;        // cloud_funcs:191-192
    float light_reach;                   // (top-bottom)*0.15, cloud_funcs:108
    float shape_hi_m01;                  // upper bound of (shape - 0.2*detail) over all texel values, see cloud_density
    float shape_mix0;                    // 0.5*(1 - u_cloud_shape_factor): the constant term of mix(0.5, tex, factor), cloud_funcs:48-50
    float dens_y_min;                    // largest y for which y*50 - 20 <= 0 in fp32 (cloud_funcs:62), see cloud_density
    float hc_min;                        // height-curve values <= hc_min cannot give a positive density for ANY coverage / shape texel
                                         // (exact bound from the largest coverage texel, atmo_consts.h: cloud_hc_min); 0 = no bound
    const float4* cube_cells;            // [6][res+1][res+1] bilinear footprints of the seamless padded faces (u8/255 as fp32)
    int cube_res;
    const float4* shape_cells;           // [nz+1][ny+1][nx+1][2] trilinear footprints of the repeat-padded volume
    int shape_nx, shape_ny, shape_nz;
    // --- variant ---
    int scatter_steps, cloud_steps;
    // --- frame front-end (unused by the ray-batch kernels) ---
    float inv_proj[16], inv_view_ray[16];  // inv_view_ray: INV_VIEW_MATRIX after the DOUBLE_PRECISION fix-up (main:118-125)
    float cam_pos_world[3];                // (inv_view * (0,0,0,1)).xyz, main:136
    const uint8_t* blue_noise;
    int bn_w, bn_h;
    int fw, fh, row_begin, row_end;
    float clip_box_half;                   // MODE_FAR proxy cube half edge (0 = fullscreen)
    int row_pitch;                         // frame kernel: rows between the 8-row tiles of consecutive blockIdx.y (8 = contiguous band;
                                           // 8*world = the interleaved multi-GPU shard, b200atmo_render_frame_peers_interleaved)
    CloudHot hot;                          // see above; keep LAST
};

struct RayIO {
    const void* origin_depth;  // float4[n]   ray batch in
    const void* dir_jitter;    // float4[n]
    const float* depth;        // float[w*h]  frame in
    void* rgba;                // float4[...] out
    void* color_inout;         // frame colour buffer to blend into (frame kernel; then rgba may be null): float4[w*h] or half4[w*h]
    int color_format;          // B200ATMO_COLOR_RGBA32F / B200ATMO_COLOR_RGBA16F
    uint8_t* discard;          // nullable
    void* out_origin_depth;    // make_rays only
    void* out_dir_jitter;
    size_t n;
    // frame front end: per-column / per-row partial products of INV_PROJECTION_MATRIX * ndc, hoisted out of the pixel
    // (ray_tables_kernel); null = compute inline. Appended AFTER n: the offsets of the fields above must not move (the
    // ptxas schedule of the march loop depends on them).
    const float4* ray_col;     // [fw]  m[0..3] * ndc.x(x)
    const float4* ray_row;     // [fh]  m[4..7] * ndc.y(y)
    // raymarched-cloud launches on a 2D block grid: heaviest-first dispatch order learned from the previous launch with the
    // same geometry (atmo_kernels.cu: block_order_kernel). Null = blockIdx order / nothing recorded.
    const unsigned* block_order;   // [gridDim.x * gridDim.y] permutation: the k-th dispatched block renders logical block block_order[k]
    unsigned* block_cost;          // [same] cycles each logical block took (max over its warps) in THIS launch
};
// RGBA16F result (b200atmo_render_frame*_fmt): the same fields, a different TYPE, so the store is chosen at compile time and
// the fp32 kernels keep their code byte for byte. rgba = half4[...].
struct RayIO16 : RayIO {};
// fused render + all-gather (b200atmo_render_*_peers): the result goes to the same symmetric buffer on every GPU instead
// of `rgba`: one multimem store when the NVLS multicast mapping is given, else one P2P store per peer. A separate
// parameter type, so the single-GPU kernels keep exactly the code (and parameter layout) they have without this path.
struct RayIOPeers : RayIO {
    void* rgba_peers[B200ATMO_MAX_PEERS];
    void* rgba_multicast;
    int n_peers;
    int first_peer;            // store loop starts here and wraps (ranks stagger their destinations)
    int use_tma;               // ray kernel: stage the block's results in smem, one cp.async.bulk per peer
    size_t peer_offset;        // pixels (float4 or half4 elements) added to the pixel / ray index in the peer buffers
    int rgba_half;             // tile format on the wire: 0 = float4, 1 = half4 (RTN-even of the fp32 result)
    // hand-shake fused into the kernel (B200AtmoPeerSync)
    unsigned* block_counter;   // local scratch, 0 before the launch, reset by the last block; null = no hand-shake
    unsigned* timeouts;        // local counter of waits that gave up
    B200AtmoPeerSync sync;
};

#ifdef __CUDACC__
// kernels (atmo_kernels.cu)
cudaError_t launch_bake_lut(float R, float H, float rho, float* d_lut, float* d_lut_pad, float4* d_lut_cells, cudaStream_t s);
cudaError_t launch_cube_pad(const uint8_t* d_faces, int res, uint8_t* d_padded, float* d_padded_f32, float4* d_cells, cudaStream_t s);
cudaError_t launch_shape_pad(const uint8_t* d_src, int nx, int ny, int nz, float* d_dst, float4* d_cells, cudaStream_t s);
cudaError_t launch_noise_cube(const B200AtmoNoise& noise, const float scale[3], int res, uint8_t* d_faces, cudaStream_t s);
cudaError_t launch_render_rays(const DevConsts& c, const RayIO& io, int scatter_model, int light_mode, cudaStream_t s);
cudaError_t launch_render_rays_peers(const DevConsts& c, const RayIOPeers& io, int scatter_model, int light_mode, cudaStream_t s);
cudaError_t launch_render_frame_peers(const DevConsts& c, const RayIOPeers& io, int scatter_model, int light_mode, cudaStream_t s);
cudaError_t launch_render_frame(const DevConsts& c, const RayIO& io, int scatter_model, int light_mode, cudaStream_t s);
cudaError_t launch_render_frame16(const DevConsts& c, const RayIO16& io, int scatter_model, int light_mode, cudaStream_t s);
cudaError_t launch_make_rays(const DevConsts& c, const RayIO& io, cudaStream_t s);
cudaError_t launch_ray_tables(const DevConsts& c, float4* d_col, float4* d_row, cudaStream_t s);
cudaError_t launch_block_order(unsigned* d_cost, unsigned* d_order, unsigned n, cudaStream_t s);
unsigned frame_grid_blocks(const DevConsts& c);     // blocks of a frame-kernel launch for rows [c.row_begin, c.row_end) at c.row_pitch
unsigned rays2d_grid_blocks(int w, int h);          // blocks of a tile-mapped cloud ray-batch launch
cudaError_t launch_peers_wait(const unsigned* d_flags, int n, unsigned epoch, unsigned* d_timeouts, cudaStream_t s);
struct PeerFlagList {
    unsigned* p[B200ATMO_MAX_PEERS];
};
cudaError_t launch_peers_signal(const PeerFlagList& flags, int n, unsigned slot, unsigned epoch, cudaStream_t s);
#endif

}  // namespace b200atmo

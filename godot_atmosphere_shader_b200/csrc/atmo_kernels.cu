// sm_100a kernels of the atmosphere hot path + their launchers. Compile with -fmad=false (see the
// numeric policy in atmo_device.cuh).
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "atmo_device.cuh"

namespace b200atmo {

// ------------------------------------------------------------------------------------------------
// optical_depth.gdshader:17-69 — LUT bake, exact arithmetic => bit-identical to the shader's floats.
// One thread per texel; also writes the clamp-to-edge padded copy the render kernels sample.
// Replaces the SubViewport render + RGBA8 bit-pack + FORMAT_RF reinterpretation
// (optical_depth.gdshader:33-43, optical_depth_baker.gd:74-85): that round trip is lossless.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bake_lut_kernel(float R, float H, float rho, float* __restrict__ lut,
                                                        float* __restrict__ lut_pad) {
    const int i = blockIdx.x * 16 + (threadIdx.x & 15);
    const int j = blockIdx.y * 16 + (threadIdx.x >> 4);
    const float uvx = (float(i) + 0.5f) / float(kLut);  // canvas UV of the texel centre
    const float uvy = (float(j) + 0.5f) / float(kLut);
    const float dir_y = 2.0f * uvx - 1.0f;               // :48-51
    const float dir_x = sqrtf(1.0f - dir_y * dir_y);
    const float pos_y = R + H * uvy;                      // :53-55
    const f2 rs = ray_sphere(mk3(0.f, 0.f, 0.f), R + H, mk3(0.0f, pos_y, 0.0f), mk3(dir_x, dir_y, 0.0f));
    const float ray_len = rs.y - fmaxf(rs.x, 0.0f);       // :63
    // get_optical_depth (:17-31), 64-step left Riemann sum
    const float step_len = ray_len / 64.0f;
    float od = 0.0f;
    for (int s = 0; s < 64; ++s) {
        const float px = 0.0f + dir_x * step_len * float(s);
        const float py = pos_y + dir_y * step_len * float(s);
        const float d = sqrtf(px * px + py * py);
        const float sd = d - R;                           // get_atmosphere_density (atmosphere_common:12-24)
        const float h = clampf(sd / H, 0.0f, 1.0f);
        const float y = 1.0f - h;
        const float density = y * y * y * rho;
        od += density * step_len * rho;
    }
    lut[j * kLut + i] = od;
    // padded copy: pad[j+1][i+1]; aprons replicate the edge (repeat_disable = clamp to edge)
    const int xs = (i == 0) ? 0 : i + 1, xe = (i == kLut - 1) ? kLutPad - 1 : i + 1;
    const int ys = (j == 0) ? 0 : j + 1, ye = (j == kLut - 1) ? kLutPad - 1 : j + 1;
    for (int y2 = ys; y2 <= ye; ++y2)
        for (int x2 = xs; x2 <= xe; ++x2) lut_pad[y2 * kLutPad + x2] = od;
}

// Bilinear cells for the scatter loop: one float4 per pair of adjacent padded texel rows/columns holding the
// patch expanded about the cell centre (make_lut_cell). 257*257*16 B = 1.06 MB, L2-resident.
__global__ void __launch_bounds__(256) lut_cells_kernel(const float* __restrict__ lut_pad, float4* __restrict__ cells) {
    const int idx = blockIdx.x * 256 + threadIdx.x;
    if (idx >= kLutCells * kLutCells) return;
    const int yi = idx / kLutCells, xi = idx % kLutCells;
    cells[idx] = make_lut_cell(lut_pad, xi, yi);
}

cudaError_t launch_bake_lut(float R, float H, float rho, float* d_lut, float* d_lut_pad, float4* d_lut_cells, cudaStream_t s) {
    bake_lut_kernel<<<dim3(kLut / 16, kLut / 16), 256, 0, s>>>(R, H, rho, d_lut, d_lut_pad);
    lut_cells_kernel<<<(kLutCells * kLutCells + 255) / 256, 256, 0, s>>>(d_lut_pad, d_lut_cells);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Seamless cube layout (integer, bit-exact with the oracle's cube_build_padded): per face a
// (res+2)^2 image whose one-texel apron holds the texel adjacent across the cube edge; the corner
// apron is the rounded mean of the three texels meeting at that cube corner. Bilinear filtering
// inside the padded face is then continuous across faces (the Vulkan seamless-cube rule).
// Face order +X,-X,+Y,-Y,+Z,-Z and (sc,tc) tables agree with noise_cubemap.gd:110-128.
// ------------------------------------------------------------------------------------------------
__constant__ int kFaceN[6][3] = {{1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}};
__constant__ int kFaceS[6][3] = {{0, 0, -1}, {0, 0, 1}, {1, 0, 0}, {1, 0, 0}, {1, 0, 0}, {-1, 0, 0}};
__constant__ int kFaceT[6][3] = {{0, -1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}, {0, -1, 0}, {0, -1, 0}};

// texel (i,j) of face f where at most one of i,j lies one step outside [0,res): fold over the cube edge
__device__ int cube_edge_texel(const uint8_t* faces, int res, int f, int i, int j) {
    const bool io = (i < 0 || i >= res), jo = (j < 0 || j >= res);
    if (!io && !jo) return faces[(size_t(f) * res + j) * res + i];
    const int sc = 2 * i + 1 - res, tc = 2 * j + 1 - res;  // doubled texel units, face plane at +-res
    int s = sc, t = tc;
    const int nrm = res - 1;
    if (io) s = sc > 0 ? res : -res;
    else t = tc > 0 ? res : -res;
    int P[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) P[k] = kFaceN[f][k] * nrm + kFaceS[f][k] * s + kFaceT[f][k] * t;
    int g;
    if (abs(P[0]) == res) g = P[0] > 0 ? 0 : 1;
    else if (abs(P[1]) == res) g = P[1] > 0 ? 2 : 3;
    else g = P[2] > 0 ? 4 : 5;
    const int gs = kFaceS[g][0] * P[0] + kFaceS[g][1] * P[1] + kFaceS[g][2] * P[2];
    const int gt = kFaceT[g][0] * P[0] + kFaceT[g][1] * P[1] + kFaceT[g][2] * P[2];
    const int gi = (gs + res - 1) / 2, gj = (gt + res - 1) / 2;
    return faces[(size_t(g) * res + gj) * res + gi];
}

__global__ void cube_pad_kernel(const uint8_t* __restrict__ faces, int res, uint8_t* __restrict__ padded,
                                float* __restrict__ padded_f32) {
    const int pr = res + 2;
    const size_t total = size_t(6) * pr * pr;
    for (size_t idx = blockIdx.x * size_t(blockDim.x) + threadIdx.x; idx < total; idx += size_t(gridDim.x) * blockDim.x) {
        const int f = int(idx / (size_t(pr) * pr));
        const int rem = int(idx % (size_t(pr) * pr));
        const int j = rem / pr - 1, i = rem % pr - 1;
        const bool io = (i < 0 || i >= res), jo = (j < 0 || j >= res);
        int v;
        if (io && jo) {
            const int ii = i < 0 ? 0 : res - 1, jj = j < 0 ? 0 : res - 1;
            const int a = cube_edge_texel(faces, res, f, ii, j);   // apron next to the corner, same row
            const int b = cube_edge_texel(faces, res, f, i, jj);   // apron next to the corner, same column
            const int c = cube_edge_texel(faces, res, f, ii, jj);  // the face's own corner texel
            v = (2 * (a + b + c) + 3) / 6;
        } else {
            v = cube_edge_texel(faces, res, f, i, j);
        }
        padded[idx] = uint8_t(v);
        padded_f32[idx] = float(v) / 255.0f;  // IEEE division: the value a UNORM8 fetch returns
    }
}

__global__ void cube_cells_kernel(const float* __restrict__ padded_f32, int res, float4* __restrict__ cells) {
    const int rc = res + 1;
    const size_t total = size_t(6) * rc * rc;
    for (size_t idx = blockIdx.x * size_t(blockDim.x) + threadIdx.x; idx < total; idx += size_t(gridDim.x) * blockDim.x) {
        const int f = int(idx / (size_t(rc) * rc)), rem = int(idx % (size_t(rc) * rc));
        cells[idx] = make_cube_cell(padded_f32, res, f, rem / rc, rem % rc);
    }
}

cudaError_t launch_cube_pad(const uint8_t* d_faces, int res, uint8_t* d_padded, float* d_padded_f32, float4* d_cells,
                            cudaStream_t s) {
    const size_t total = size_t(6) * (res + 2) * (res + 2);
    const int blocks = int((total + 255) / 256);
    cube_pad_kernel<<<blocks > 4096 ? 4096 : blocks, 256, 0, s>>>(d_faces, res, d_padded, d_padded_f32);
    cube_cells_kernel<<<blocks > 4096 ? 4096 : blocks, 256, 0, s>>>(d_padded_f32, res, d_cells);
    return cudaGetLastError();
}

// 3D shape: repeat-padded fp32 copy [nz+2][ny+2][nx+2], padded[z+1][y+1][x+1] = texel(x mod nx, ..)/255
__global__ void shape_pad_kernel(const uint8_t* __restrict__ src, int nx, int ny, int nz, float* __restrict__ dst) {
    const int px = nx + 2, py = ny + 2, pz = nz + 2;
    const size_t total = size_t(px) * py * pz;
    for (size_t idx = blockIdx.x * size_t(blockDim.x) + threadIdx.x; idx < total; idx += size_t(gridDim.x) * blockDim.x) {
        const int x = int(idx % px), y = int((idx / px) % py), z = int(idx / (size_t(px) * py));
        const int sx = (x - 1 + nx) % nx, sy = (y - 1 + ny) % ny, sz = (z - 1 + nz) % nz;
        dst[idx] = float(src[(size_t(sz) * ny + sy) * nx + sx]) / 255.0f;
    }
}

__global__ void shape_cells_kernel(const float* __restrict__ padded, int nx, int ny, int nz, float4* __restrict__ cells) {
    const int cx = nx + 1, cy = ny + 1;
    const size_t total = size_t(cx) * cy * (nz + 1);
    for (size_t idx = blockIdx.x * size_t(blockDim.x) + threadIdx.x; idx < total; idx += size_t(gridDim.x) * blockDim.x) {
        const int xi = int(idx % cx), yi = int((idx / cx) % cy), zi = int(idx / (size_t(cx) * cy));
        cells[2 * idx] = make_shape_cell(padded, nx, ny, zi, yi, xi);
        cells[2 * idx + 1] = make_shape_cell(padded, nx, ny, zi + 1, yi, xi);
    }
}

cudaError_t launch_shape_pad(const uint8_t* d_src, int nx, int ny, int nz, float* d_dst, float4* d_cells, cudaStream_t s) {
    const size_t total = size_t(nx + 2) * (ny + 2) * (nz + 2);
    const int blocks = int((total + 255) / 256);
    shape_pad_kernel<<<blocks > 8192 ? 8192 : blocks, 256, 0, s>>>(d_src, nx, ny, nz, d_dst);
    shape_cells_kernel<<<blocks > 8192 ? 8192 : blocks, 256, 0, s>>>(d_dst, nx, ny, nz, d_cells);
    return cudaGetLastError();
}

// NoiseCubemap._generate_images (noise_cubemap.gd:101-140) on the device: one thread per texel.
struct NoiseCubeArgs {
    B200AtmoNoise noise;
    float scale[3];
    int res;
};
__global__ void __launch_bounds__(256) noise_cube_kernel(const __grid_constant__ NoiseCubeArgs a, uint8_t* __restrict__ faces) {
    const size_t total = size_t(6) * a.res * a.res;
    for (size_t idx = blockIdx.x * size_t(blockDim.x) + threadIdx.x; idx < total; idx += size_t(gridDim.x) * blockDim.x) {
        const int side = int(idx / (size_t(a.res) * a.res)), rem = int(idx % (size_t(a.res) * a.res));
        faces[idx] = noise_cube_texel(side, rem % a.res, rem / a.res, a.res, a.scale, a.noise);
    }
}

cudaError_t launch_noise_cube(const B200AtmoNoise& noise, const float scale[3], int res, uint8_t* d_faces, cudaStream_t s) {
    NoiseCubeArgs a;
    a.noise = noise;
    for (int k = 0; k < 3; ++k) a.scale[k] = scale[k];
    a.res = res;
    const size_t total = size_t(6) * res * res;
    const size_t blocks = (total + 255) / 256;
    noise_cube_kernel<<<unsigned(blocks > 148 * 64 ? 148 * 64 : blocks), 256, 0, s>>>(a, d_faces);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// render kernels
// ------------------------------------------------------------------------------------------------
#ifndef B200ATMO_BLOCK
#define B200ATMO_BLOCK 128
#endif
#ifdef B200ATMO_MIN_BLOCKS
#define B200ATMO_BOUNDS __launch_bounds__(B200ATMO_BLOCK, B200ATMO_MIN_BLOCKS)
#else
#define B200ATMO_BOUNDS __launch_bounds__(B200ATMO_BLOCK)
#endif
constexpr int kBlock = B200ATMO_BLOCK;
// warp-wide compaction of the raymarched cloud light (atmo_device.cuh: raymarch_cloud_light_queue); 0 = per-thread march
#ifndef B200ATMO_LIGHT_QUEUE
#define B200ATMO_LIGHT_QUEUE 0
#endif
constexpr bool kLightQueue = B200ATMO_LIGHT_QUEUE != 0;
// block size of the ray kernels that march clouds (ragged work: a block's warp slots are held until its slowest warp
// is done, so smaller blocks keep more warps resident); tuning knob, profiles/r02/tune_clouds.txt
#ifndef B200ATMO_CLOUD_BLOCK
#define B200ATMO_CLOUD_BLOCK B200ATMO_BLOCK
#endif
__host__ __device__ constexpr int ray_block(int light) { return light ? B200ATMO_CLOUD_BLOCK : kBlock; }

// Result store. Multi-GPU shards write straight into every rank's copy of a symmetric buffer over NVLink: with the NVLS
// multicast mapping one 16-byte store is replicated by the NVSwitch to all GPUs (the render IS the all-gather, no
// second pass over the tile); without it, one peer-to-peer store per rank.
// Overloaded on the parameter type (RayIOPeers only for the *_peers kernels): the single-GPU kernels compile to
// exactly the code they have without this path.
// RGBA16F: every channel rounded to nearest-even, like a store into an RGBA16F colour target
__device__ __forceinline__ uint2 pack_half4(float4 v) {
    const __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
    return make_uint2(*reinterpret_cast<const unsigned*>(&a), *reinterpret_cast<const unsigned*>(&b));
}
__device__ __forceinline__ void store_rgba(const RayIO& io, size_t i, float4 v) { __stcs(static_cast<float4*>(io.rgba) + i, v); }
__device__ __forceinline__ void store_rgba(const RayIO16& io, size_t i, float4 v) {
    const uint2 h = pack_half4(v);
    __stcs(reinterpret_cast<float2*>(io.rgba) + i, make_float2(__uint_as_float(h.x), __uint_as_float(h.y)));
}
__device__ __forceinline__ void store_rgba(const RayIOPeers& io, size_t i, float4 v) {
    if (io.rgba_half) {   // half4 tiles: 8 bytes per pixel on the wire
        const uint2 h = pack_half4(v);
        if (io.rgba_multicast) {
            uint2* p = static_cast<uint2*>(io.rgba_multicast) + io.peer_offset + i;
            asm volatile("multimem.st.relaxed.sys.global.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(__uint_as_float(h.x)), "f"(__uint_as_float(h.y))
                         : "memory");
        } else {
            int r = io.first_peer;
            for (int k = 0; k < io.n_peers; ++k) {
                static_cast<uint2*>(io.rgba_peers[r])[io.peer_offset + i] = h;
                r = (r + 1 == io.n_peers) ? 0 : r + 1;
            }
        }
        return;
    }
    if (io.rgba_multicast) {
        float4* p = static_cast<float4*>(io.rgba_multicast) + io.peer_offset + i;
        asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                     : "memory");
    } else {
        int r = io.first_peer;
        for (int k = 0; k < io.n_peers; ++k) {
            static_cast<float4*>(io.rgba_peers[r])[io.peer_offset + i] = v;
            r = (r + 1 == io.n_peers) ? 0 : r + 1;
        }
    }
}

// Completion signal of the fused render + delivery kernels: when the LAST block of the grid has stored its pixels, it
// publishes io.done_epoch into the consumers' flag arrays. Per block: barrier (all stores of the block issued), system
// fence, one atomic on a local counter; the block that completes the count knows every other block's fence came first
// (atomics on one location are totally ordered), fences again and releases the flags. All threads of every block must call.
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// spin until flags[0..n) have all reached `epoch` (signed distance: wrap-around safe); gives up after ~2 s
__device__ __forceinline__ void wait_flags(const unsigned* flags, int n, unsigned epoch, unsigned* timeouts) {
    const long long t0 = clock64();
    for (int k = 0; k < n; ++k) {
        while (int(ld_acquire_sys(flags + k) - epoch) < 0) {
            if (clock64() - t0 > 4000000000ll) {   // a peer died: do not hang the GPU
                if (timeouts) atomicAdd(timeouts, 1u);
                return;
            }
            __nanosleep(64);
        }
    }
}
template <class IO> __device__ __forceinline__ void peer_begin(const IO&) {}
template <class IO> __device__ __forceinline__ void peer_done(const IO&) {}
// kernel start: the first block tells the producers which frame this rank has consumed; every block waits for its credit
__device__ __forceinline__ void peer_begin(const RayIOPeers& io) {
    if (!io.block_counter) return;
    const B200AtmoPeerSync& y = io.sync;
    if (threadIdx.x == 0) {
        if (blockIdx.x == 0 && blockIdx.y == 0)
            for (int k = 0; k < y.n_consumed_flags; ++k)
                st_release_sys(static_cast<unsigned*>(y.d_consumed_flags[k]) + y.consumed_slot, y.consumed_epoch);
        if (y.n_credit > 0) wait_flags(static_cast<const unsigned*>(y.d_credit_flags) + y.credit_first_slot, y.n_credit, y.credit_epoch, io.timeouts);
    }
    if (y.n_credit > 0) __syncthreads();
}
// kernel end: per block a barrier (all its stores are issued), a device-scope fence and one atomic on a local counter. The
// block that completes the count has observed every other block's fence + atomic (device scope), so its own SYSTEM-scope
// fence orders all of the grid's peer stores before the flags it then releases (causality is transitive across scopes).
// On a consumer that block finally waits for the producers' flags: the kernel ends when every tile has landed.
__device__ __forceinline__ void peer_done(const RayIOPeers& io) {
    if (!io.block_counter) return;
    const B200AtmoPeerSync& y = io.sync;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned total = gridDim.x * gridDim.y;
        if (atomicAdd(io.block_counter, 1u) == total - 1u) {
            *io.block_counter = 0u;   // ready for the next launch that is handed this counter
            __threadfence_system();
            for (int k = 0; k < y.n_done_flags; ++k) st_release_sys(static_cast<unsigned*>(y.d_done_flags[k]) + y.done_slot, y.epoch);
            if (y.n_wait > 0) wait_flags(static_cast<const unsigned*>(y.d_wait_flags) + y.wait_first_slot, y.n_wait, y.epoch, io.timeouts);
            __threadfence_system();
        }
    }
}
// ------------------------------------------------------------------------------------------------------------------
// Heaviest-first dispatch of the blocks of a raymarched-cloud launch. A block whose rays cross dense cloud lives ~5x longer
// than the average block (one ray is a serial chain of up to 128 x 6 density evaluations), blocks are dispatched in blockIdx
// order, and whatever heavy blocks start late finish alone: a list-scheduling model of the 3840x2160 frame on 148 x 9 block slots
// (profiles/microbench/tail_model.py) puts that tail at +7 % on one GPU and +35 % when the frame is split over 8 — the measured
// strong-scaling loss — and at +0 % / +3 % when the blocks are dispatched longest first. The costs come from the previous
// launch with the same geometry on the same stream (frames are temporally coherent): every warp adds its elapsed cycles to its
// logical block's slot (atomicMax), block_order_kernel then sorts the blocks by cost into `block_order` (counting sort, 256
// bins, one 1024-thread block) and clears the costs. The k-th dispatched block renders logical block block_order[k]: a
// permutation of the blocks — pixels are untouched. Only the raymarched-light kernels use it (their blocks are long enough).
// ------------------------------------------------------------------------------------------------------------------
#ifndef B200ATMO_ORDER_CHEAP
#define B200ATMO_ORDER_CHEAP 0      // tuning knob: also order the cheap-light cloud kernels (with B200ATMO_BLOCK_ORDER=2)
#endif
__host__ __device__ constexpr bool uses_block_order(int light) {
    return (light & 3) == B200ATMO_LIGHT_RAYMARCHED || (B200ATMO_ORDER_CHEAP && (light & 3) == B200ATMO_LIGHT_CHEAP);
}
__device__ __forceinline__ unsigned logical_block(const RayIO& io) {
    unsigned lb = blockIdx.y * gridDim.x + blockIdx.x;
    if (io.block_order) lb = __ldg(io.block_order + lb);
    return lb;
}
__device__ __forceinline__ void record_block_cost(const RayIO& io, unsigned lb, long long t0) {
    if (!io.block_cost) return;
    const unsigned lanes = __activemask();
    if ((threadIdx.x & 31u) == unsigned(__ffs(int(lanes)) - 1)) {
        const long long dt = clock64() - t0;
        atomicMax(io.block_cost + lb, dt > 0xffffffffll ? 0xffffffffu : unsigned(dt));
    }
}
__global__ void __launch_bounds__(1024) block_order_kernel(unsigned* __restrict__ cost, unsigned* __restrict__ order, unsigned n) {
    __shared__ unsigned s_max, s_hist[256], s_pos[256];
    const unsigned tid = threadIdx.x;
    if (tid == 0) s_max = 0u;
    if (tid < 256u) s_hist[tid] = 0u;
    __syncthreads();
    unsigned m = 0u;
    for (unsigned i = tid; i < n; i += 1024u) m = max(m, cost[i]);
    atomicMax(&s_max, m);
    __syncthreads();
    const unsigned long long mx = s_max ? s_max : 1u;
    for (unsigned i = tid; i < n; i += 1024u) atomicAdd(&s_hist[255u - unsigned(cost[i] * 255ull / mx)], 1u);   // bin 0 = the heaviest
    __syncthreads();
    if (tid == 0) {
        unsigned acc = 0u;
        for (int b = 0; b < 256; ++b) {
            s_pos[b] = acc;
            acc += s_hist[b];
        }
    }
    __syncthreads();
    for (unsigned i = tid; i < n; i += 1024u) order[atomicAdd(&s_pos[255u - unsigned(cost[i] * 255ull / mx)], 1u)] = i;
    __syncthreads();
    for (unsigned i = tid; i < n; i += 1024u) cost[i] = 0u;
}
cudaError_t launch_block_order(unsigned* d_cost, unsigned* d_order, unsigned n, cudaStream_t s) {
    block_order_kernel<<<1, 1024, 0, s>>>(d_cost, d_order, n);
    return cudaGetLastError();
}

template <class IO> struct IsPeers { static constexpr bool value = false; };
template <> struct IsPeers<RayIOPeers> { static constexpr bool value = true; };

// Stand-alone wait: one thread spins until flags[0..n) have reached `epoch`.
__global__ void __launch_bounds__(32) peers_wait_kernel(const unsigned* __restrict__ flags, int n, unsigned epoch, unsigned* timeouts) {
    if (threadIdx.x == 0) wait_flags(flags, n, epoch, timeouts);
    __threadfence_system();
}
__global__ void __launch_bounds__(32) peers_signal_kernel(const PeerFlagList flags, int n, unsigned slot, unsigned epoch) {
    if (int(threadIdx.x) < n) {
        __threadfence_system();
        st_release_sys(flags.p[threadIdx.x] + slot, epoch);
    }
}
cudaError_t launch_peers_wait(const unsigned* d_flags, int n, unsigned epoch, unsigned* d_timeouts, cudaStream_t s) {
    peers_wait_kernel<<<1, 32, 0, s>>>(d_flags, n, epoch, d_timeouts);
    return cudaGetLastError();
}
cudaError_t launch_peers_signal(const PeerFlagList& flags, int n, unsigned slot, unsigned epoch, cudaStream_t s) {
    peers_signal_kernel<<<1, 32, 0, s>>>(flags, n, slot, epoch);
    return cudaGetLastError();
}

// Ray batch: thread i <-> ray i. Two coalesced LDG.128 in (streaming), one STG.128 out.
// TILED (b200atmo_render_rays_2d): the batch is a c.fw x c.fh pixel grid; a warp covers an 8x4 pixel tile like the frame
// kernel (four 128-byte segments per load instead of one 512-byte run), so its lanes enter and leave the cloud shell
// together. A separate instantiation: the linear kernel keeps its code byte for byte.
// launch bounds of the ray kernels: scatter-only kernels take B200ATMO_MIN_BLOCKS, cloud kernels B200ATMO_CLOUD_MIN_BLOCKS
// (tuning knobs; without them ptxas' own register heuristic is used — an explicit minimum of 1 would make it spend 112
// registers on the cloud kernels)
template <int L> struct RayBounds {
#if defined(B200ATMO_CLOUD_MIN_BLOCKS)
    static constexpr int kCloudMin = B200ATMO_CLOUD_MIN_BLOCKS;
#else
    static constexpr int kCloudMin = 0;
#endif
#if defined(B200ATMO_MIN_BLOCKS)
    static constexpr int kScatterMin = B200ATMO_MIN_BLOCKS;
#else
    static constexpr int kScatterMin = 0;
#endif
    static constexpr int kMin = L ? kCloudMin : kScatterMin;
};
#define B200ATMO_RAY_BOUNDS(L) __launch_bounds__(ray_block(L), RayBounds<L>::kMin)
template <int MODEL, int LIGHT, class IO, bool TILED>
__global__ void B200ATMO_RAY_BOUNDS(LIGHT) render_rays_kernel(const __grid_constant__ DevConsts c, const IO io) {
    constexpr int BS = ray_block(LIGHT), WX = BS >= 64 ? 2 : 1, WY = BS / 32 / WX;   // block tile = (8*WX) x (4*WY) pixels
    size_t i;
    bool valid;
    constexpr bool ORDERED = TILED && uses_block_order(LIGHT);   // heaviest-first block dispatch (block_order_kernel)
    unsigned lb = 0;
    long long t_start = 0;
    if (TILED) {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        unsigned bx = blockIdx.x, by = blockIdx.y;
        if (ORDERED) {
            lb = logical_block(io);
            bx = lb % gridDim.x;
            by = lb / gridDim.x;
            t_start = clock64();
        }
        const int x = bx * (8 * WX) + (warp % WX) * 8 + (lane & 7);
        const int y = by * (4 * WY) + (warp / WX) * 4 + (lane >> 3);
        valid = x < c.fw && y < c.fh;
        i = size_t(y) * c.fw + x;
    } else {
        i = blockIdx.x * size_t(BS) + threadIdx.x;
        valid = i < io.n;
    }
    if constexpr ((LIGHT & 3) == B200ATMO_LIGHT_RAYMARCHED && kLightQueue) {
        // every lane of the warp stays (lanes without a ray are passive): the cloud march is warp-cooperative
        __shared__ float4 s_queue[BS / 32][128];
        peer_begin(io);
        float4 od = make_float4(0.f, 0.f, 0.f, 0.f), dj = od, out;
        if (valid) {
            od = __ldcs(static_cast<const float4*>(io.origin_depth) + i);
            dj = __ldcs(static_cast<const float4*>(io.dir_jitter) + i);
        }
        const bool disc = shade_ray_light_queue<MODEL, (LIGHT & kLightPow2) != 0>(c, valid, mk3(od.x, od.y, od.z), mk3(dj.x, dj.y, dj.z), od.w, dj.w, out,
                                                        s_queue[threadIdx.x >> 5]);
        if (valid) {
            store_rgba(io, i, out);
            if (io.discard) io.discard[i] = disc ? 1 : 0;
        }
        peer_done(io);
        return;
    }
    if (!IsPeers<IO>::value && !valid) return;   // the peers kernels keep every thread for the fused hand-shake
    peer_begin(io);
    if (valid) {
        const float4 od = __ldcs(static_cast<const float4*>(io.origin_depth) + i);
        const float4 dj = __ldcs(static_cast<const float4*>(io.dir_jitter) + i);
        float4 out;
        const bool disc = shade_ray<MODEL, LIGHT>(c, mk3(od.x, od.y, od.z), mk3(dj.x, dj.y, dj.z), od.w, dj.w, out);
        store_rgba(io, i, out);
        if (io.discard) io.discard[i] = disc ? 1 : 0;
        if (ORDERED) record_block_cost(io, lb, t_start);
    }
    peer_done(io);
}

// Peer variant with TMA bulk stores (RayIOPeers::use_tma): the block's 128 results are staged in shared memory and ONE
// elected thread sends the 2 KB tile to every rank with cp.async.bulk (shared::cta -> global, the global address being the
// peer mapping), instead of 128 threads x n_peers STG.128. No thread leaves before the barrier.
template <int MODEL, int LIGHT, int UNUSED>
__global__ void B200ATMO_BOUNDS render_rays_tma_peers_kernel(const __grid_constant__ DevConsts c, const RayIOPeers io) {
    __shared__ __align__(128) float4 s_out[kBlock];   // half4 tiles use the first half of it
    const size_t base = blockIdx.x * size_t(kBlock);
    const size_t i = base + threadIdx.x;
    float4 out = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < io.n) {
        const float4 od = __ldcs(static_cast<const float4*>(io.origin_depth) + i);
        const float4 dj = __ldcs(static_cast<const float4*>(io.dir_jitter) + i);
        shade_ray<MODEL, LIGHT>(c, mk3(od.x, od.y, od.z), mk3(dj.x, dj.y, dj.z), od.w, dj.w, out);
    }
    const unsigned px_bytes = io.rgba_half ? 8u : 16u;
    if (io.rgba_half) reinterpret_cast<uint2*>(s_out)[threadIdx.x] = pack_half4(out);
    else s_out[threadIdx.x] = out;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the async (TMA) proxy
    __syncthreads();
    if (threadIdx.x == 0) {
        const size_t left = io.n - base;
        const unsigned all = unsigned(left < size_t(kBlock) ? left : size_t(kBlock)) * px_bytes;
        const unsigned bytes = all & ~15u;   // cp.async.bulk moves multiples of 16 bytes: an odd half4 tail pixel goes by a plain store
        const unsigned src = unsigned(__cvta_generic_to_shared(s_out));
        int r = io.first_peer;
        for (int k = 0; k < io.n_peers; ++k) {
            char* dst = static_cast<char*>(io.rgba_peers[r]) + (io.peer_offset + base) * px_bytes;
            if (all != bytes) *reinterpret_cast<uint2*>(dst + bytes) = reinterpret_cast<const uint2*>(s_out)[bytes / 8u];
            if (bytes == 0u) { r = (r + 1 == io.n_peers) ? 0 : r + 1; continue; }
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
            r = (r + 1 == io.n_peers) ? 0 : r + 1;
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // the tile has left shared memory and been written
    }
}

// Frame: a warp covers an 8x4 pixel tile (coherent LUT / texture footprints, full 128 B store
// segments per tile row), a block of 4 warps a 16x8 tile.
template <int MODEL, int LIGHT, class IO>
__device__ __forceinline__ void frame_pixel(const DevConsts& c, const IO& io, int x, int y) {
    const size_t i = size_t(y) * c.fw + x;
    f3 o, d;
    float linear_depth, jitter;
    make_ray(c, x, y, __ldcs(io.depth + i), o, d, linear_depth, jitter, io.ray_col, io.ray_row);
    float4 out = make_float4(0.f, 0.f, 0.f, 0.f);
    bool disc = true;
    if (!(c.clip_box_half > 0.0f) || far_box_covers(c, o, d, linear_depth))  // MODE_FAR: outside the proxy cube = not rasterised
        disc = shade_ray<MODEL, LIGHT>(c, o, d, linear_depth, jitter, out);
    if (io.color_inout) {  // the ROP's blend_mix of an unshaded spatial shader, exact arithmetic; discard = no write
        if (!disc) {
            const float ia = 1.0f - out.w;
            if (io.color_format == B200ATMO_COLOR_RGBA16F) {
                // the Forward+ 3D colour target: blended in fp32, stored round-to-nearest-even like the ROP; alpha bits untouched
                uint2* dst = static_cast<uint2*>(io.color_inout) + i;
                const uint2 raw = *dst;
                const float2 rg = __half22float2(*reinterpret_cast<const __half2*>(&raw.x));
                const float b = __half2float(__ushort_as_half((unsigned short)(raw.y & 0xffffu)));
                const __half2 o_rg = __floats2half2_rn(out.x * out.w + rg.x * ia, out.y * out.w + rg.y * ia);
                const unsigned o_b = __half_as_ushort(__float2half_rn(out.z * out.w + b * ia));
                *dst = make_uint2(*reinterpret_cast<const unsigned*>(&o_rg), (raw.y & 0xffff0000u) | o_b);
            } else {
                float4* dst = static_cast<float4*>(io.color_inout) + i;
                const float4 bg = *dst;
                *dst = make_float4(out.x * out.w + bg.x * ia, out.y * out.w + bg.y * ia, out.z * out.w + bg.z * ia, bg.w);
            }
        }
    } else {
        store_rgba(io, i, out);
    }
    if (io.discard) io.discard[i] = disc ? 1 : 0;
}
template <int MODEL, int LIGHT, class IO>
__global__ void __launch_bounds__(kBlock) render_frame_kernel(const __grid_constant__ DevConsts c, const IO io) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr bool ORDERED = uses_block_order(LIGHT);   // heaviest-first block dispatch (block_order_kernel)
    unsigned lb = 0, bx = blockIdx.x, by = blockIdx.y;
    long long t_start = 0;
    if (ORDERED) {
        lb = logical_block(io);
        bx = lb % gridDim.x;
        by = lb / gridDim.x;
        t_start = clock64();
    }
    int x, y;
    if (IsPeers<IO>::value && (LIGHT & 3) == B200ATMO_LIGHT_NONE) {
        // peer stores of a scatter-only frame are NVLink-bound, not issue-bound: a warp covers 16x2 pixels, so a tile row is one
        // 256-byte (float4) / 128-byte (half4) run instead of 128 / 64 bytes — the ncu NVLink counters show 19 % / 41 % protocol
        // bytes on top of the pixels for the 8x4 shape (profiles/r02/nvlink_counters.txt). Same 16x8 block tile, same pixels.
        x = bx * 16 + (lane & 15);
        y = c.row_begin + by * c.row_pitch + warp * 2 + (lane >> 4);
    } else {
        x = bx * 16 + (warp & 1) * 8 + (lane & 7);
        y = c.row_begin + by * c.row_pitch + (warp >> 1) * 4 + (lane >> 3);
    }
    const bool valid = x < c.fw && y < c.row_end;
    if (!IsPeers<IO>::value && !valid) return;   // the peers kernels keep every thread for the fused hand-shake
    peer_begin(io);
    if (valid) {
        frame_pixel<MODEL, LIGHT>(c, io, x, y);
        if (ORDERED) record_block_cost(io, lb, t_start);
    }
    peer_done(io);
}

// Frame front-end only: depth buffer -> SoA rays for the batch API.
__global__ void __launch_bounds__(kBlock) make_rays_kernel(const __grid_constant__ DevConsts c, const RayIO io) {
    const size_t i = blockIdx.x * size_t(kBlock) + threadIdx.x;
    if (i >= io.n) return;
    const int x = int(i % size_t(c.fw)), y = int(i / size_t(c.fw));
    f3 o, d;
    float linear_depth, jitter;
    make_ray(c, x, y, io.depth[i], o, d, linear_depth, jitter, io.ray_col, io.ray_row);
    static_cast<float4*>(io.out_origin_depth)[i] = make_float4(o.x, o.y, o.z, linear_depth);
    static_cast<float4*>(io.out_dir_jitter)[i] = make_float4(d.x, d.y, d.z, jitter);
}

// Per-column / per-row tables of the frame front end (see make_ray): fw + fh threads, once per (size, projection).
__global__ void __launch_bounds__(256) ray_tables_kernel(const __grid_constant__ DevConsts c, float4* __restrict__ col,
                                                          float4* __restrict__ row) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < c.fw) col[i] = ray_col_entry(c, i);
    else if (i < c.fw + c.fh) row[i - c.fw] = ray_row_entry(c, i - c.fw);
}
cudaError_t launch_ray_tables(const DevConsts& c, float4* d_col, float4* d_row, cudaStream_t s) {
    ray_tables_kernel<<<(c.fw + c.fh + 255) / 256, 256, 0, s>>>(c, d_col, d_row);
    return cudaGetLastError();
}

// KERNEL<MODEL, LIGHT, TAIL...>: TAIL = the remaining template arguments (IO type). LIGHT = light mode, plus kLightPow2
// when every texture dimension is a power of two (`light_mode` arrives here already combined: see light_template_arg)
#define B200ATMO_DISPATCH(KERNEL, GRID, TAIL, ...)                                                            \
    do {                                                                                                     \
        switch ((scatter_model == B200ATMO_SCATTER_V1 ? 8 : 0) + light_mode) {                                \
            case 0: KERNEL<0, 0, TAIL><<<GRID, kBlock, 0, s>>>(__VA_ARGS__); break;                           \
            case 1: KERNEL<0, 1, TAIL><<<GRID, kBlock, 0, s>>>(__VA_ARGS__); break;                           \
            case 2: KERNEL<0, 2, TAIL><<<GRID, kBlock, 0, s>>>(__VA_ARGS__); break;                           \
            case 5: KERNEL<0, 5, TAIL><<<GRID, kBlock, 0, s>>>(__VA_ARGS__); break;                           \
            case 6: KERNEL<0, 6, TAIL><<<GRID, kBlock, 0, s>>>(__VA_ARGS__); break;                           \
            case 8: KERNEL<1, 0, TAIL><<<GRID, kBlock, 0, s>>>(__VA_ARGS__); break;                           \
            case 9: KERNEL<1, 1, TAIL><<<GRID, kBlock, 0, s>>>(__VA_ARGS__); break;                           \
            case 10: KERNEL<1, 2, TAIL><<<GRID, kBlock, 0, s>>>(__VA_ARGS__); break;                          \
            case 13: KERNEL<1, 5, TAIL><<<GRID, kBlock, 0, s>>>(__VA_ARGS__); break;                          \
            default: KERNEL<1, 6, TAIL><<<GRID, kBlock, 0, s>>>(__VA_ARGS__); break;                          \
        }                                                                                                    \
    } while (0)
// light mode of the C-ABI + "all texture dimensions are powers of two" -> the LIGHT template argument
static int light_template_arg(const DevConsts& c, int light_mode) {
    auto pow2 = [](int v) { return v > 0 && (v & (v - 1)) == 0; };
    if (light_mode == B200ATMO_LIGHT_NONE) return 0;
    const bool all = pow2(c.cube_res) && pow2(c.shape_nx) && pow2(c.shape_ny) && pow2(c.shape_nz);
    return light_mode | (all ? kLightPow2 : 0);
}
#define B200ATMO_COMMA ,

template <int M, int L, class IO> static void launch_rays_one(const DevConsts& c, const IO& io, cudaStream_t s) {
    constexpr int BS = ray_block(L), WX = BS >= 64 ? 2 : 1, WY = BS / 32 / WX;
    if (c.fw > 0) {   // 2D batch: one (8*WX) x (4*WY) pixel tile per block
        const dim3 grid((c.fw + 8 * WX - 1) / (8 * WX), (c.fh + 4 * WY - 1) / (4 * WY));
        render_rays_kernel<M, L, IO, true><<<grid, BS, 0, s>>>(c, io);
    } else {
        render_rays_kernel<M, L, IO, false><<<unsigned((io.n + BS - 1) / BS), BS, 0, s>>>(c, io);
    }
}
template <class IO> static cudaError_t launch_rays_t(const DevConsts& c, const IO& io, int scatter_model, int light_mode, cudaStream_t s) {
    if (io.n == 0) return cudaSuccess;
    switch ((scatter_model == B200ATMO_SCATTER_V1 ? 8 : 0) + light_template_arg(c, light_mode)) {
        case 0: launch_rays_one<0, 0>(c, io, s); break;
        case 1: launch_rays_one<0, 1>(c, io, s); break;
        case 2: launch_rays_one<0, 2>(c, io, s); break;
        case 5: launch_rays_one<0, 5>(c, io, s); break;
        case 6: launch_rays_one<0, 6>(c, io, s); break;
        case 8: launch_rays_one<1, 0>(c, io, s); break;
        case 9: launch_rays_one<1, 1>(c, io, s); break;
        case 10: launch_rays_one<1, 2>(c, io, s); break;
        case 13: launch_rays_one<1, 5>(c, io, s); break;
        default: launch_rays_one<1, 6>(c, io, s); break;
    }
    return cudaGetLastError();
}
template <class IO> static cudaError_t launch_frame_t(const DevConsts& c, const IO& io, int scatter_model, int light_mode, cudaStream_t s) {
    const int rows = c.row_end - c.row_begin;
    if (rows <= 0 || c.fw <= 0) return cudaSuccess;
    const dim3 grid((c.fw + 15) / 16, (rows + c.row_pitch - 1) / c.row_pitch);   // row_pitch 8: one block row per 8 rows
    light_mode = light_template_arg(c, light_mode);
    B200ATMO_DISPATCH(render_frame_kernel, grid, IO, c, io);
    return cudaGetLastError();
}
unsigned frame_grid_blocks(const DevConsts& c) {
    const int rows = c.row_end - c.row_begin;
    if (rows <= 0 || c.fw <= 0 || c.row_pitch <= 0) return 0u;
    return unsigned((c.fw + 15) / 16) * unsigned((rows + c.row_pitch - 1) / c.row_pitch);
}
unsigned rays2d_grid_blocks(int w, int h) {
    constexpr int BS = ray_block(B200ATMO_LIGHT_RAYMARCHED), WX = BS >= 64 ? 2 : 1, WY = BS / 32 / WX;
    if (w <= 0 || h <= 0) return 0u;
    return unsigned((w + 8 * WX - 1) / (8 * WX)) * unsigned((h + 4 * WY - 1) / (4 * WY));
}
cudaError_t launch_render_rays(const DevConsts& c, const RayIO& io, int scatter_model, int light_mode, cudaStream_t s) {
    return launch_rays_t(c, io, scatter_model, light_mode, s);
}
cudaError_t launch_render_rays_peers(const DevConsts& c, const RayIOPeers& io, int scatter_model, int light_mode, cudaStream_t s) {
    if (!io.use_tma || io.rgba_multicast || c.fw > 0) return launch_rays_t(c, io, scatter_model, light_mode, s);   // the TMA kernel stages LINEAR runs of 128 results
    if (io.n == 0) return cudaSuccess;
    const unsigned grid = unsigned((io.n + kBlock - 1) / kBlock);
    light_mode = light_template_arg(c, light_mode);
    B200ATMO_DISPATCH(render_rays_tma_peers_kernel, grid, 0, c, io);
    return cudaGetLastError();
}
cudaError_t launch_render_frame(const DevConsts& c, const RayIO& io, int scatter_model, int light_mode, cudaStream_t s) {
    return launch_frame_t(c, io, scatter_model, light_mode, s);
}
cudaError_t launch_render_frame16(const DevConsts& c, const RayIO16& io, int scatter_model, int light_mode, cudaStream_t s) {
    return launch_frame_t(c, io, scatter_model, light_mode, s);
}
cudaError_t launch_render_frame_peers(const DevConsts& c, const RayIOPeers& io, int scatter_model, int light_mode, cudaStream_t s) {
    return launch_frame_t(c, io, scatter_model, light_mode, s);
}

cudaError_t launch_make_rays(const DevConsts& c, const RayIO& io, cudaStream_t s) {
    if (io.n == 0) return cudaSuccess;
    const unsigned grid = unsigned((io.n + kBlock - 1) / kBlock);
    make_rays_kernel<<<grid, kBlock, 0, s>>>(c, io);
    return cudaGetLastError();
}

}  // namespace b200atmo

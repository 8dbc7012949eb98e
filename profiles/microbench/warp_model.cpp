// PROFILE TOOLING (CPU) — lane-occupancy model of the cloud march, never loaded by the product.
// Runs the product's own device functions (csrc/atmo_device.cuh, host build) for the rays of sampled warps and records,
// per cloud step and lane, how far the step gets (out of shell / coverage early-out / shape fetch / density > 0) and the
// same for the 6 light-march samples of every hit. From those masks it prices three execution strategies in issued
// warp-instructions:
//   per_thread   : today's kernel — a warp pays for a section if ANY lane needs it
//   light_queue  : hits are pushed into a per-warp FIFO and the 6-step light march runs on batches of 32 queued items
//   ideal        : every section costs (lanes needing it) / 32 — the floor of any regrouping
// The section costs are parameters (SASS instruction counts, profiles/r02/sass_cloud_sections.txt).
#include "../../godot_atmosphere_shader_b200/csrc/atmo_consts.h"
#include "../../godot_atmosphere_shader_b200/csrc/atmo_device.cuh"

#include <vector>

using namespace b200atmo;

namespace {

// cloud_density with the exit point reported: 0 out of shell, 1 coverage early-out, 2 shape fetched (dens may be 0), 3 dens > 0
int density_stage(const DevConsts& c, f3 p, float hr, float& dens) {
    dens = 0.0f;
    const float a = 2.0f * hr - 1.0f;
    const float hc = 1.0f - a * a;
    if (!(hc > c.hc_min)) return 0;
    const float cpx = c.rot[0] * p.x + c.rot[2] * p.z;
    const float cpz = c.rot[1] * p.x + c.rot[3] * p.z;
    float coverage = sample_cube<false>(c.hot.cube_cells, c.hot.cube_res, cpx, p.y, cpz);
    coverage = coverage - 0.25f * hr + c.coverage_bias;
    const float cov_term = mixf(-1.2f, 1.5f, coverage);
    if (!((c.shape_hi_m01 + cov_term) * hc * 50.0f - 20.0f > 0.0f)) return 1;
    dens = cloud_density<false>(c, p, hr);
    return dens > 0.0f ? 3 : 2;
}

struct Costs {
    double base, cube, shape, hit_tail, light_step_base, light_tail, push, pop;
};

struct Acc {
    double per_thread = 0, light_queue = 0, ideal = 0;
    double lanes_steps = 0, lanes_shell = 0, lanes_shape = 0, lanes_hit = 0, warp_steps = 0, warp_any_hit = 0;
    double light_items = 0, light_batches = 0, warp_steps_no_shell = 0;
};

struct Item {
    unsigned char st[6];
};

double light_cost_mask(const Costs& k, const std::vector<Item>& items, size_t b, size_t e) {
    double cost = 0;
    for (int j = 0; j < 6; ++j) {
        bool any1 = false, any2 = false;
        for (size_t i = b; i < e; ++i) {
            any1 |= items[i].st[j] >= 1;
            any2 |= items[i].st[j] >= 2;
        }
        cost += k.light_step_base + (any1 ? k.cube : 0) + (any2 ? k.shape : 0);
    }
    return cost + k.light_tail;
}

}  // namespace

extern "C" {

// rays: od/dj float4 SoA of n_warps*32 rays (warp w = rays [32w, 32w+32)); out[16] receives the accumulators
void warp_model_run(const B200AtmoParams* p, const B200AtmoFrame* fr, const float* lut_pad, const float* cube_pad, int cube_res,
                    const float* shape_pad, int nx, int ny, int nz, int cloud_steps, const float* od, const float* dj,
                    size_t n_warps, const double* costs8, double* out, double* per_warp /* nullable: per-thread-strategy cost of each warp */) {
    std::vector<float4> cube_cells, shape_cells;
    const int rc = cube_res + 1;
    cube_cells.resize(size_t(6) * rc * rc);
    for (int f = 0; f < 6; ++f)
        for (int yi = 0; yi < rc; ++yi)
            for (int xi = 0; xi < rc; ++xi) cube_cells[(size_t(f) * rc + yi) * rc + xi] = make_cube_cell(cube_pad, cube_res, f, yi, xi);
    const int cx = nx + 1, cy = ny + 1, cz = nz + 1;
    shape_cells.resize(size_t(cx) * cy * cz * 2);
    for (int zi = 0; zi < cz; ++zi)
        for (int yi = 0; yi < cy; ++yi)
            for (int xi = 0; xi < cx; ++xi) {
                const size_t idx = (size_t(zi) * cy + yi) * cx + xi;
                shape_cells[2 * idx] = make_shape_cell(shape_pad, nx, ny, zi, yi, xi);
                shape_cells[2 * idx + 1] = make_shape_cell(shape_pad, nx, ny, zi + 1, yi, xi);
            }
    DeviceTextures t;
    t.lut_pad = lut_pad;
    t.cube_cells = cube_cells.data();
    t.cube_res = cube_res;
    t.shape_cells = shape_cells.data();
    t.nx = nx; t.ny = ny; t.nz = nz;
    Variant v;
    v.scatter_steps = 8;
    v.cloud_steps = cloud_steps;
    v.light_mode = B200ATMO_LIGHT_RAYMARCHED;
    DevConsts c;
    consts_from_params(c, *p, v, t);
    consts_set_frame(c, *p, fr->planet_center_view, fr->sun_center_view, fr->inv_view);
    const Costs k{costs8[0], costs8[1], costs8[2], costs8[3], costs8[4], costs8[5], costs8[6], costs8[7]};
    Acc A;
    const f3 C = ld3(c.C), sun = ld3(c.sun_dir_model);
    for (size_t w = 0; w < n_warps; ++w) {
        const double cost_before = A.per_thread;
        if (per_warp) per_warp[w] = 0.0;
        // per-lane march set-up (render_clouds + raymarch_cloud prologue)
        bool act[32];
        f3 pos[32], dstep[32];
        bool any_act = false;
        for (int l = 0; l < 32; ++l) {
            const size_t i = w * 32 + l;
            const f3 o = mk3(od[4 * i], od[4 * i + 1], od[4 * i + 2]), d = mk3(dj[4 * i], dj[4 * i + 1], dj[4 * i + 2]);
            float linear_depth = od[4 * i + 3];
            const float jitter = dj[4 * i + 3];
            act[l] = false;
            const f2 rs_atmo = ray_sphere(C, c.atmo_radius, o, d);
            if (rs_atmo.x == rs_atmo.y) continue;
            const f2 rs_ground = ray_sphere(C, c.R, o, d);
            float gd = 10000000.0f;
            if (rs_ground.x != rs_ground.y) gd = rs_ground.x;
            linear_depth = mixf(linear_depth, gd, c.sphere_depth_factor);
            const f2 rs_top = ray_sphere(C, c.cloud_top_h, o, d);
            if (rs_top.x == rs_top.y) continue;
            const f2 rs_bottom = ray_sphere(C, c.cloud_bottom_h, o, d);
            const float t0 = fmaxf(rs_top.x, 0.0f);
            float t1 = fminf(rs_top.y, linear_depth);
            if (!(t0 < linear_depth && (linear_depth > rs_bottom.y || rs_bottom.x > 0.0f))) continue;
            float om[4], dm[4];
            mat4_mul(c.v2m, o.x, o.y, o.z, 1.0f, om);
            mat4_mul(c.v2m, d.x, d.y, d.z, 0.0f, dm);
            const f3 o2 = mk3(om[0], om[1], om[2]), d2 = mk3(dm[0], dm[1], dm[2]);
            const float max_d = mixf(c.march_ground, c.march_space, smoothstepf(c.march_hmin, c.march_hmax, sqrtf(dot3(o2, o2))));
            t1 = t0 + fminf(t1 - t0, max_d);
            const float step_len = (t1 - t0) * (1.0f / float(cloud_steps));
            pos[l] = o2 + jitter * step_len * d2 + d2 * t0;
            dstep[l] = d2 * step_len;
            act[l] = true;
            any_act = true;
        }
        if (!any_act) continue;
        std::vector<Item> queue;   // light-march work items in push order (lane order within a step)
        for (int s = 0; s < cloud_steps; ++s) {
            int n_act = 0, n1 = 0, n2 = 0, n3 = 0;
            std::vector<Item> step_items;
            for (int l = 0; l < 32; ++l) {
                if (!act[l]) continue;
                ++n_act;
                float inv, dens;
                const float len = sqrt_refined(dot3(pos[l], pos[l]), inv);
                const float hr = cloud_height_ratio(c, len);
                const int st = density_stage(c, pos[l], hr, dens);
                n1 += st >= 1; n2 += st >= 2; n3 += st >= 3;
                if (st == 3) {
                    Item it;
                    float sl = c.light_reach * (1.0f / 6.0f);
                    for (int j = 0; j < 6; ++j) {
                        const float tt = float(j) * sl;
                        const f3 q = mk3(pos[l].x + tt * sun.x, pos[l].y + tt * sun.y, pos[l].z + tt * sun.z);
                        float inv2, dd;
                        const float len2 = sqrt_refined(dot3(q, q), inv2);
                        it.st[j] = (unsigned char)density_stage(c, q, cloud_height_ratio(c, len2), dd);
                        sl *= 1.2f;
                    }
                    step_items.push_back(it);
                }
                pos[l] = pos[l] + dstep[l];
            }
            A.warp_steps += 1;
            if (n1 == 0) A.warp_steps_no_shell += 1;
            A.lanes_steps += n_act; A.lanes_shell += n1; A.lanes_shape += n2; A.lanes_hit += n3;
            const double front = k.base + (n1 ? k.cube : 0) + (n2 ? k.shape : 0);
            A.per_thread += front;
            A.light_queue += front;
            A.ideal += (k.base * 32 /* the loop itself runs on every lane of an active warp */ + k.cube * n1 + k.shape * n2) / 32.0;
            if (n3) {
                A.warp_any_hit += 1;
                A.per_thread += k.hit_tail + light_cost_mask(k, step_items, 0, step_items.size());
                A.light_queue += k.hit_tail + k.push;
                for (auto& it : step_items) {
                    queue.push_back(it);
                    double ci = k.hit_tail + k.light_tail;
                    for (int j = 0; j < 6; ++j) ci += k.light_step_base + (it.st[j] >= 1 ? k.cube : 0) + (it.st[j] >= 2 ? k.shape : 0);
                    A.ideal += ci / 32.0;
                }
            }
        }
        if (per_warp) per_warp[w] = A.per_thread - cost_before;
        A.light_items += double(queue.size());
        for (size_t b = 0; b < queue.size(); b += 32) {
            const size_t e = b + 32 < queue.size() ? b + 32 : queue.size();
            A.light_queue += k.pop + light_cost_mask(k, queue, b, e);
            A.light_batches += 1;
        }
    }
    out[0] = A.per_thread; out[1] = A.light_queue; out[2] = A.ideal; out[3] = A.lanes_steps; out[4] = A.lanes_shell;
    out[5] = A.lanes_shape; out[6] = A.lanes_hit; out[7] = A.warp_steps; out[8] = A.warp_any_hit; out[9] = A.light_items;
    out[10] = A.light_batches;
    out[11] = A.warp_steps_no_shell;
}

}  // extern "C"

#!/usr/bin/env python
"""bench.py — ray-steps/s of the atmosphere hot path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W]              our CUDA path
  python bench.py --impl reference [--gpus N] [--steps K] ...      the CPU arm (scalar C++ oracle, all host threads)

A "step" = one pass of the hot path over one frame of synthetic rays. Workload at N=1 = BASELINE.json
configs[1]: 1920x1080, 32 in-scatter steps, LUT mode, no clouds, scene "demo",
Camera B (every ray hits the atmosphere, so ray-steps = W*H*32). N>1: weak scaling, every rank renders
its own 1080p tile of an N-tile offscreen target (no data-path collective; the optional NCCL all-gather
of the RGBA tiles is measured separately under "gather").

`value`     : ray-steps/s, rays resident in HBM, kernel timed with CUDA events on the launch stream,
              L2 flushed between timed iterations.
`e2e`       : same metric through the host-buffer C-ABI (pinned host buffers), every step's depth H2D and RGBA D2H
              inside the timed region: b200atmo_render_frame_host_submit / b200atmo_frame_wait over two pipeline
              slots; the one-frame-at-a-time call b200atmo_render_frame_host is reported under e2e.synchronous.
`roofline`  : algorithmic HBM bytes (48 B/ray: 2 x float4 in, 1 x float4 out) / kernel time vs the measured
              copy bandwidth. NB this path is FP32-issue/MUFU bound at N=32; see DESIGN.md.
`cpu_baseline`: the oracle (kind "port": the reference has no CPU implementation) on this box's host cores.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "ray_steps_per_sec"
UNIT = "ray-steps/s"
ALGO_BYTES_PER_RAY = 48  # 32 B in (2 x float4) + 16 B out (RGBA f32)
FRAME_BYTES_PER_PIXEL_IN, FRAME_BYTES_PER_PIXEL_OUT = 4, 16


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--scatter-steps", type=int, default=32)
    ap.add_argument("--cloud-steps", type=int, default=0)
    ap.add_argument("--light", type=int, default=0)
    ap.add_argument("--camera", choices=["A", "B"], default="B")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=60)
    return ap.parse_args()


def workload_config(a, n_gpus):
    return {
        "workload": f"{a.width}x{a.height} frame, {a.scatter_steps} in-scatter steps (LUT mode), "
                    + (f"{a.cloud_steps} cloud steps light={a.light}, " if a.light else "no clouds, ")
                    + f"scene demo (R=100,H=8,u_density=0.5), camera {a.camera}"
                    + (" (all rays hit)" if a.camera == "B" else ""),
        "width": a.width, "height": a.height, "scatter_steps": a.scatter_steps, "cloud_steps": a.cloud_steps,
        "light_mode": a.light, "camera": a.camera, "rays_per_gpu": a.width * a.height,
        "parallelism": f"screen-tile shard x{n_gpus}" if n_gpus > 1 else "single GPU",
        "l2": "flushed between timed iterations (256 MiB write)",
    }


def build_scene(a):
    from godot_atmosphere_shader_b200 import scenes
    p = scenes.demo_params()
    cam = scenes.camera_b(a.width, a.height, p) if a.camera == "B" else scenes.camera_a(a.width, a.height)
    depth = scenes.synth_depth(cam, p, a.width, a.height)
    tex = dict(bn=scenes.blue_noise_tile())
    if a.light:
        tex["shape"] = scenes.shape_texture(64, seed=1)
        tex["cube"] = scenes.coverage_cubemap(256, seed=1)
    return p, cam, depth, tex


# ------------------------------------------------------------------------------------------------
# clocks (NVML sampled DURING the timed region)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.nv is not None:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()

    def stop(self):
        self._stop.set()
        if self._t is not None:
            self._t.join()
        return {"sm_mhz": (int(np.median(self.samples)) if self.samples else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the scalar C++ oracle on all host threads
# ------------------------------------------------------------------------------------------------
def usable_cpus():
    """Threads the CPU arm should use: min(hardware threads, scheduler affinity, cgroup CPU quota)."""
    n = os.cpu_count() or 1
    info = {"hardware_threads": n}
    try:
        aff = len(os.sched_getaffinity(0))
        info["affinity"] = aff
        n = min(n, aff)
    except Exception:
        pass
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            txt = open(path).read().split()
            if path.endswith("cpu.max"):
                if txt[0] != "max":
                    q = max(1, int(float(txt[0]) / float(txt[1]) + 0.5))
                    info["cgroup_quota_cpus"] = q
                    n = min(n, q)
            else:
                quota = int(txt[0])
                if quota > 0:
                    period = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
                    q = max(1, int(quota / period + 0.5))
                    info["cgroup_quota_cpus"] = q
                    n = min(n, q)
            break
        except Exception:
            continue
    return n, info


def cpu_frame_runner(a, budget_s_per_step):
    """Returns (run, info): run() renders one bounded sample on the host cores and returns (seconds, ray_steps).

    kind "reference": oracle/_ref — the reference's own GDShader sources compiled as C++ (oracle/ref/build_ref.py; built
    where /root/reference exists, the .so travels with the repo), its fragment() run per pixel of every `stride`-th row.
    kind "port" (only if that library is missing): the hand-written scalar C++ oracle on the same rows."""
    from oracle import pyoracle as O
    from oracle import pyref as R
    p, cam, depth, tex = build_scene(a)
    otex = O.Textures(lut=O.bake_lut(p), shape=tex.get("shape"), cube_faces=tex.get("cube"), blue_noise=tex["bn"])
    var = O.variant(a.scatter_steps, a.cloud_steps, a.light)
    threads, cpu_info = usable_cpus()
    w, h = a.width, a.height
    use_ref = R.available()
    if use_ref:
        def render(stride):
            t0 = time.perf_counter()
            _, disc = R.render_frame(p, var, cam, otex, depth, w, h, threads=threads, row_stride=stride)
            dt = time.perf_counter() - t0
            return dt, int((disc[::stride] == 0).sum()) * a.scatter_steps
    else:
        od, dj, fr = O.make_rays(p, cam, otex, depth, w, h)

        def render(stride):
            rows = np.arange(0, h, stride)
            idx = (rows[:, None] * w + np.arange(w)[None, :]).reshape(-1)
            s_od, s_dj = np.ascontiguousarray(od[idx]), np.ascontiguousarray(dj[idx])
            t0 = time.perf_counter()
            _, disc = O.render_rays(p, var, fr, otex, s_od, s_dj, threads=threads)
            dt = time.perf_counter() - t0
            return dt, int((disc == 0).sum()) * a.scatter_steps

    # calibrate on 1/16 of the rows, then pick a row stride that keeps one step under the budget
    render(16)
    dt, _ = render(16)
    full_est = dt * 16
    stride = 1
    while full_est / stride > budget_s_per_step and stride < h:
        stride *= 2
    n_rows = len(range(0, h, stride))
    what = ("the reference's GDShader sources compiled as C++ (oracle/_ref), fragment() per pixel" if use_ref
            else "scalar C++ oracle port -O2 no-FMA (oracle/_ref not built)")
    info = {"cores": threads, "host": cpu_info, "kind": "reference" if use_ref else "port",
            "sample": (f"full {w}x{h} frame" if stride == 1 else f"every {stride}th row of the {w}x{h} frame ({n_rows} rows)")
                      + f", {a.scatter_steps} steps, {what}, {threads} std::thread workers"}
    return (lambda: render(stride)), info


def run_reference(a, rank, world):
    if rank != 0:
        return
    total = a.steps + a.warmup
    budget = max(0.02, min(2.0, 150.0 / max(total, 1)))
    run, info = cpu_frame_runner(a, budget)
    for _ in range(a.warmup):
        run()
    t_sum, steps_sum = 0.0, 0
    for _ in range(a.steps):
        dt, rs = run()
        t_sum += dt
        steps_sum += rs
    value = steps_sum / t_sum
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * t_sum / a.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(a, a.gpus),
        "mpixels_per_s": value / a.scatter_steps / 1e6,
        "cpu_baseline": dict(info, value=value, unit=UNIT),
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "the reference ships no CPU implementation (GDShader only); this arm runs its shader sources compiled as C++ on the host cores (kind reference), or the oracle port if that library is missing",
    }
    emit(line)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def load_ncu_facts(a):
    """Per-launch facts of the dominant kernel from the committed `ncu --set full` capture of this workload
    (profiles/roofline_traffic.json): DRAM bytes (read+write) and executed warp instructions. {} if none."""
    path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    try:
        d = json.load(open(path))
        key = f"{a.width}x{a.height}x{a.scatter_steps}_c{a.cloud_steps}_l{a.light}_cam{a.camera}"
        return d.get(key, {})
    except Exception:
        return {}


def bind_to_gpu_numa_node(local_rank):
    """Run this rank's host thread on the CPUs of its GPU's NUMA node (so its pinned buffers are allocated there and the
    PCIe copies do not cross the socket interconnect). Best effort: returns the node or None, never raises."""
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(local_rank)).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:   # NVML prints an 8-digit PCI domain, sysfs a 4-digit one
            bus = bus[4:]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        usable = cpus & os.sched_getaffinity(0)
        if usable:
            os.sched_setaffinity(0, usable)
            return node
    except Exception:
        pass
    return None


def run_ours(a, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    from godot_atmosphere_shader_b200 import context

    assert torch.cuda.is_available(), "bench.py needs a CUDA device; there is no CPU fallback for the product path"
    numa_node = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    w, h, N = a.width, a.height, a.scatter_steps
    n_rays = w * h
    p, cam, depth, tex = build_scene(a)

    ctx = context.AtmosphereContext(local_rank)
    ctx.set_params(p)
    ctx.set_variant(N, a.cloud_steps, a.light)
    ctx.upload_blue_noise(tex["bn"])
    if a.light:
        ctx.upload_shape3d(tex["shape"])
        ctx.upload_coverage_cube(tex["cube"])
    stream = torch.cuda.current_stream().cuda_stream

    d_depth = torch.from_numpy(depth).to(dev)
    d_od = torch.empty((n_rays, 4), dtype=torch.float32, device=dev)
    d_dj = torch.empty((n_rays, 4), dtype=torch.float32, device=dev)
    fr = ctx.make_rays(cam, d_depth, w, h, d_od, d_dj, stream=stream)
    d_rgba = torch.empty((n_rays, 4), dtype=torch.float32, device=dev)
    d_disc = torch.empty((n_rays,), dtype=torch.uint8, device=dev)
    ctx.render_rays(fr, d_od, d_dj, n_rays, d_rgba, d_disc, stream=stream)
    torch.cuda.synchronize()
    hit_rays = int((d_disc == 0).sum().item())
    ray_steps = hit_rays * N
    checksum = float(d_rgba.double().sum().item())

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_loop(fn, steps, warmup):
        for _ in range(warmup):
            flush.zero_()
            fn()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        t0 = time.perf_counter()
        for s, e in ev:
            flush.zero_()  # L2 flush, outside the event pair
            s.record()
            fn()
            e.record()
        barrier()
        wall = time.perf_counter() - t0
        ms = sum(s.elapsed_time(e) for s, e in ev)
        return ms, wall

    # A timed region that saw a hardware / thermal slowdown is rejected and measured again, once (all ranks decide together).
    remeasured = False
    for attempt in range(2):
        sampler = ClockSampler(local_rank)
        l0 = ctx.launch_count
        sampler.start()
        ms_total, wall = timed_loop(lambda: ctx.render_rays(fr, d_od, d_dj, n_rays, d_rgba, None, stream=stream), a.steps, a.warmup)
        clocks = sampler.stop()
        launches = ctx.launch_count - l0 - a.warmup
        bad = bool(set(clocks.get("reasons", [])) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"})
        flag = torch.tensor([1 if bad else 0], dtype=torch.int32, device=dev)
        if world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        if not int(flag.item()) or attempt == 1:
            break
        remeasured = True
        time.sleep(2.0)
    clocks["remeasured_after_slowdown"] = remeasured

    # max over ranks
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_per_step = ms_total / a.steps
    value = world * ray_steps / (ms_per_step * 1e-3)

    # the frame API on device buffers (depth in, rays generated on the fly: 20 B/pixel instead of 48 B/ray)
    d_rgba2 = torch.empty_like(d_rgba)
    fms, _ = timed_loop(lambda: ctx.render_frame(cam, d_depth, w, h, d_rgba2, None, stream=stream), max(20, a.steps // 4), 3)
    tf = torch.tensor([fms / max(20, a.steps // 4)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tf, op=dist.ReduceOp.MAX)
    frame_ms = float(tf.item())
    frame_ok = bool(torch.equal(d_rgba2, d_rgba))

    # optional: compute + NCCL all-gather of the RGBA tiles (BASELINE config[4]); not part of `value`
    gather = None
    if world > 1:
        gathered = torch.empty((world * n_rays, 4), dtype=torch.float32, device=dev)

        def step_gather():
            ctx.render_rays(fr, d_od, d_dj, n_rays, d_rgba, None, stream=stream)
            dist.all_gather_into_tensor(gathered, d_rgba)

        gms, _ = timed_loop(step_gather, max(10, a.steps // 4), 3)
        g = torch.tensor([gms / max(10, a.steps // 4)], dtype=torch.float64, device=dev)
        dist.all_reduce(g, op=dist.ReduceOp.MAX)
        gather = {"ms_per_step": float(g.item()), "value": world * ray_steps / (float(g.item()) * 1e-3), "unit": UNIT,
                  "bytes_gathered_per_gpu": world * n_rays * 16, "collective": "ncclAllGather (torch.distributed)"}
        # fused: the render kernel stores every finished RGBA value into all ranks' symmetric buffers over NVLink
        # (NVLS multicast when the fabric offers it) -- the render IS the all-gather, no collective pass
        from godot_atmosphere_shader_b200.sharding import SymmetricTiles, render_rays_and_gather_fused
        fused = {}
        for label, use_mc, use_tma in (("multicast", True, False), ("p2p", False, False), ("p2p_tma", False, True)):
            try:
                tiles = SymmetricTiles(world, n_rays, dev, use_multicast=use_mc, use_tma=use_tma)
                if use_mc and not tiles.multicast_ptr:
                    fused[label] = {"unavailable": "no NVLS multicast mapping on this fabric"}
                    continue
                fms2, _ = timed_loop(lambda: render_rays_and_gather_fused(ctx, fr, d_od, d_dj, n_rays, tiles, stream=stream),
                                     max(10, a.steps // 4), 3)
                fz = torch.tensor([fms2 / max(10, a.steps // 4)], dtype=torch.float64, device=dev)
                dist.all_reduce(fz, op=dist.ReduceOp.MAX)
                ok = bool(torch.equal(tiles.tensor.view(world * n_rays, 4), gathered))
                fused[label] = {"ms_per_step": float(fz.item()), "value": world * ray_steps / (float(fz.item()) * 1e-3),
                                "tiles_match_nccl": ok}
                del tiles
            except Exception as exc:  # symmetric memory unavailable (no P2P / driver support): report, do not fail the bench
                fused[label] = {"unavailable": repr(exc)[:200]}
        gather["fused"] = dict(fused, api="b200atmo_render_rays_peers + symmetric-memory barrier (timed incl. the barrier)")
        # chunked + overlapped delivery through the frame API (sharding.render_tile_and_gather_overlapped)
        from godot_atmosphere_shader_b200.sharding import render_tile_and_gather_overlapped
        mine = torch.empty((h, w, 4), dtype=torch.float32, device=dev)
        best = None
        for chunks in (2, 4, 8):
            if h % chunks:
                continue
            slabs = torch.empty((chunks, world, h // chunks, w, 4), dtype=torch.float32, device=dev)
            oms, _ = timed_loop(lambda: render_tile_and_gather_overlapped(ctx, cam, d_depth, w, h, mine, slabs, rank, world,
                                                                          chunks, stream=stream), max(10, a.steps // 4), 3)
            o = torch.tensor([oms / max(10, a.steps // 4)], dtype=torch.float64, device=dev)
            dist.all_reduce(o, op=dist.ReduceOp.MAX)
            ok = bool(torch.equal(slabs[:, rank].reshape(-1, 4), d_rgba)) and bool(torch.equal(slabs[:, (rank + 1) % world], slabs[:, rank]))
            if best is None or float(o.item()) < best[1]:
                best = (chunks, float(o.item()), ok)
            del slabs
        gather["overlapped"] = {"ms_per_step": best[1], "value": world * ray_steps / (best[1] * 1e-3), "chunks": best[0],
                                "tiles_match": best[2], "layout": "chunk-major [chunks, world, rows, w, 4]"}

    # e2e through the host-buffer C-ABI (pinned host memory). Every step uploads that step's depth buffer and reads that
    # step's RGBA back. Two forms: the synchronous call (one frame at a time), and the pipelined submit/wait pair over
    # two slots (frame k downloads while frame k+1 uploads and renders) — the form a stream of frames uses.
    h_depth = torch.from_numpy(depth).pin_memory()
    h_rgba = torch.empty((n_rays, 4), dtype=torch.float32).pin_memory()
    for _ in range(3):
        ctx.render_frame_host(cam, h_depth, w, h, h_rgba, None)
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.e2e_steps):
        ctx.render_frame_host(cam, h_depth, w, h, h_rgba, None)
    torch.cuda.synchronize()
    e2e_sync_s = (time.perf_counter() - t0) / a.e2e_steps
    e2e_ok = bool(np.array_equal(h_rgba.numpy(), d_rgba.cpu().numpy()))
    h_depths = [h_depth, torch.from_numpy(depth.copy()).pin_memory()]
    h_rgbas = [h_rgba, torch.empty((n_rays, 4), dtype=torch.float32).pin_memory()]
    h_rgbas[0].zero_()

    def pipelined(steps):
        for k in range(steps):
            ctx.frame_wait(k & 1)
            ctx.render_frame_host_submit(cam, h_depths[k & 1], w, h, h_rgbas[k & 1], None, slot=k & 1)
        ctx.frame_wait(0)
        ctx.frame_wait(1)

    pipelined(4)
    barrier()
    t0 = time.perf_counter()
    pipelined(a.e2e_steps)
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / a.e2e_steps
    e2e_ok = e2e_ok and all(bool(np.array_equal(b.numpy(), d_rgba.cpu().numpy())) for b in h_rgbas[:min(2, a.e2e_steps)])
    # the whole transparent pass against Godot's RGBA16F colour target: depth + colour up, render + blend_mix, colour down
    from godot_atmosphere_shader_b200 import abi as _abi
    h_color = torch.zeros((n_rays, 4), dtype=torch.float16).pin_memory()
    for _ in range(3):
        ctx.composite_frame_host(cam, h_depth, w, h, h_color, color_format=_abi.COLOR_RGBA16F)
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.e2e_steps):
        ctx.composite_frame_host(cam, h_depth, w, h, h_color, color_format=_abi.COLOR_RGBA16F)
    torch.cuda.synchronize()
    comp_s = (time.perf_counter() - t0) / a.e2e_steps
    te = torch.tensor([e2e_s, e2e_sync_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s, e2e_sync_s = float(te[0].item()), float(te[1].item())

    if rank != 0:
        return
    peak, peak_src = load_peaks()
    facts = load_ncu_facts(a)
    achieved = ALGO_BYTES_PER_RAY * n_rays / (ms_per_step * 1e-3) / 1e9
    # second roofline: the resource that actually binds this kernel is the warp-instruction issue rate
    # (1 instr/clk/SMSP). Peak = 148 SMs x 4 SMSPs x SM clock under load.
    issue = None
    if facts.get("warp_instructions") and clocks.get("sm_mhz"):
        sms = torch.cuda.get_device_properties(local_rank).multi_processor_count
        ipeak = sms * 4 * clocks["sm_mhz"] * 1e6
        iach = facts["warp_instructions"] / (ms_per_step * 1e-3)
        issue = {"bound": "issue", "achieved": iach / 1e9, "peak": ipeak / 1e9, "unit": "G warp-instr/s", "frac": iach / ipeak,
                 "warp_instructions_per_launch": facts["warp_instructions"], "source": facts.get("source")}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(a, world),
        "mpixels_per_s": world * n_rays / (ms_per_step * 1e-3) / 1e6,
        "hit_fraction": hit_rays / n_rays, "ray_steps_per_step": world * ray_steps,
        "clocks": clocks,
        "e2e": {"value": world * ray_steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": n_rays * FRAME_BYTES_PER_PIXEL_IN,
                "d2h_bytes_per_step": n_rays * FRAME_BYTES_PER_PIXEL_OUT, "ms_per_step": e2e_s * 1e3,
                "api": "b200atmo_render_frame_host_submit + b200atmo_frame_wait (pinned host depth in, RGBA out; 2 pipeline "
                       "slots: frame k's D2H overlaps frame k+1's H2D + kernel)",
                "timer": "host perf_counter around the whole loop incl. the final waits, max over ranks", "matches_device_path": e2e_ok,
                "host_numa_node_rank0": numa_node,
                "synchronous": {"value": world * ray_steps / e2e_sync_s, "ms_per_step": e2e_sync_s * 1e3,
                                "api": "b200atmo_render_frame_host (one frame at a time, 4 row bands over 2 streams)"}},
        "e2e_composite_rgba16f": {"ms_per_step": comp_s * 1e3, "value": ray_steps / comp_s, "unit": UNIT,
                                  "h2d_bytes_per_step": n_rays * 12, "d2h_bytes_per_step": n_rays * 8,
                                  "api": "b200atmo_composite_frame_host (rank 0; fp32 render + blend_mix into an RGBA16F frame)"},
        "gpu_launches": launches,
        "frame_api": {"ms_per_step": frame_ms, "value": world * ray_steps / (frame_ms * 1e-3), "unit": UNIT,
                      "api": "b200atmo_render_frame (device depth in, 4 B + 16 B per pixel)", "bit_identical_to_ray_api": frame_ok},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": facts.get("dram_bytes"), "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": ALGO_BYTES_PER_RAY * n_rays,
                     "kernel": "render_rays_kernel<V2, no clouds>" if not a.light else "render_rays_kernel<V2, clouds>",
                     "note": "FP32-issue/MUFU bound at N=32 by construction (1.5 B/ray-step); see DESIGN.md"},
        "timed_wall_s": wall, "checksum": checksum,
    }
    if issue:
        line["roofline_issue"] = issue
    if gather:
        line["gather"] = gather
    if world == 1 and not a.no_cpu_baseline:
        run, info = cpu_frame_runner(a, budget_s_per_step=4.0)
        run()
        ts = [run() for _ in range(3)]
        best = min(ts, key=lambda x: x[0])
        line["cpu_baseline"] = dict(info, value=best[1] / best[0], unit=UNIT, ms_per_step=best[0] * 1e3)
    emit(line)
    ctx.close()


_REAL_STDOUT = None


def quiet_stdout():
    """Everything any library prints to fd 1 (NCCL's "NCCL version ..." banner goes to stdout) is routed to stderr; the
    single JSON line is written to the real stdout by emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    fd = _REAL_STDOUT if _REAL_STDOUT is not None else 1
    while data:
        n = os.write(fd, data)
        data = data[n:]


def main():
    a = parse_args()
    quiet_stdout()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if a.impl == "reference":
        run_reference(a, rank, world)
        return
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(a, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()

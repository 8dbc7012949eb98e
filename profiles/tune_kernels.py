#!/usr/bin/env python
"""Kernel-variant timing harness (GPU). Each build variant of libb200atmo.so (profiles/build_variants.sh -> tune_libs/) is
loaded in its own process through B200ATMO_LIB; prints one line per variant: kernel ms per workload and a hash of every
output buffer, so variants that claim to be bit-identical can be checked against the base build.
usage: B200ATMO_LIB=tune_libs/lib_x.so python profiles/tune_kernels.py [--quick]"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    quick = "--quick" in sys.argv
    only = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--only=")]
    torch.cuda.set_device(0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def timed(fn, steps, warmup=3):
        for _ in range(warmup):
            flush.zero_()
            fn()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        torch.cuda.synchronize()
        for s, e in ev:
            flush.zero_()
            s.record()
            fn()
            e.record()
        torch.cuda.synchronize()
        ts = sorted(s.elapsed_time(e) for s, e in ev)
        return sum(ts) / len(ts), ts[len(ts) // 2]

    W = bench.Workload
    work = [("cfg2", W(1920, 1080, 32, 0, 0, "B"), 100), ("cfg2_n8", W(1920, 1080, 8, 0, 0, "B"), 100),
            ("cfg3A", W(1920, 1080, 8, 64, 1, "A"), 40), ("cfg3C", W(1920, 1080, 8, 64, 1, "C"), 30),
            ("cfg4A", W(3840, 2160, 8, 128, 2, "A"), 8), ("cfg4C", W(3840, 2160, 8, 128, 2, "C"), 4),
            ("rm1080A", W(1920, 1080, 8, 64, 2, "A"), 20)]
    if quick:
        work = [w for w in work if w[0] in ("cfg2", "cfg3A", "cfg4A")]
    if only:
        work = [w for w in work if w[0] in only]
    out = {"lib": os.environ.get("B200ATMO_LIB", "default")}
    for name, wl, steps in work:
        R = bench.Runner(torch, wl, 0)
        lin, lin_med = timed(lambda: R.render_rays(grid=False), steps)
        h1 = hashlib.sha1(R.d_rgba.cpu().numpy().tobytes()).hexdigest()[:12]
        til, til_med = timed(lambda: R.render_rays(grid=True), steps)
        h2 = hashlib.sha1(R.d_rgba.cpu().numpy().tobytes()).hexdigest()[:12]
        d_out = torch.empty_like(R.d_rgba)
        frm, frm_med = timed(lambda: R.ctx.render_frame(R.cam, R.d_depth, wl.width, wl.height, d_out, None), steps)
        h3 = hashlib.sha1(d_out.cpu().numpy().tobytes()).hexdigest()[:12]
        out[name] = {"linear_ms": round(lin, 5), "tiled_ms": round(til, 5), "frame_ms": round(frm, 5),
                     "linear_med": round(lin_med, 5), "tiled_med": round(til_med, 5), "hash": h1, "all_equal": h1 == h2 == h3}
        R.close()
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()

#!/usr/bin/env bash
# Runs on the B200 box (under gpurun): GPU tests, bench lines and ncu captures. Outputs -> gpurun_out/.
# Numbers printed by a run under ncu are never bench values; bench lines come from the plain runs.
set -u
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15) > gpurun_out/pytest_gpu.log
# cfg2 (headline): 1920x1080 x 32 steps, no clouds
timeout 600 python bench.py > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err
# cfg2 with camera A (realistic orbit view, ~85% hit)
timeout 300 python bench.py --camera A --no-cpu-baseline > gpurun_out/bench_cfg2_camA.json 2>> gpurun_out/bench_cfg2.err
# reference step count N=8
timeout 300 python bench.py --scatter-steps 8 --no-cpu-baseline > gpurun_out/bench_1080p_n8.json 2>> gpurun_out/bench_cfg2.err
# cfg3: scatter 8 + clouds_high (64 steps, cheap light), camera A and B
timeout 300 python bench.py --scatter-steps 8 --cloud-steps 64 --light 1 --camera A --steps 50 --no-cpu-baseline > gpurun_out/bench_cfg3_camA.json 2> gpurun_out/bench_cfg3.err
timeout 300 python bench.py --scatter-steps 8 --cloud-steps 64 --light 1 --camera B --steps 50 --no-cpu-baseline > gpurun_out/bench_cfg3_camB.json 2>> gpurun_out/bench_cfg3.err
# cfg4: 3840x2160, clouds_high_rm 128 x 6
timeout 600 python bench.py --width 3840 --height 2160 --scatter-steps 8 --cloud-steps 128 --light 2 --camera A --steps 10 --warmup 3 --e2e-steps 3 --no-cpu-baseline > gpurun_out/bench_cfg4_camA.json 2> gpurun_out/bench_cfg4.err
# ncu: launch list of the default bench command, then full captures of the dominant kernels
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_cfg2.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/ncu_l.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_rays -s 3 -c 1 -f -o gpurun_out/prof_1920x1080x32_c0_l0_camB python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_a.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_rays -s 3 -c 1 -f -o gpurun_out/prof_1920x1080x8_c64_l1_camA python bench.py --scatter-steps 8 --cloud-steps 64 --light 1 --camera A --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_b.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_rays -s 2 -c 1 -f -o gpurun_out/prof_1920x1080x8_c128_l2_camA python bench.py --width 1920 --height 1080 --scatter-steps 8 --cloud-steps 128 --light 2 --camera A --steps 2 --warmup 2 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_c.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_rays -s 3 -c 1 -f -o gpurun_out/prof_1920x1080x8_c0_l0_camB python bench.py --scatter-steps 8 --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_d.log 2>&1
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference_cfg2.json 2> gpurun_out/bench_reference.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 --scatter-steps 8 --cloud-steps 64 --light 1 --camera A > gpurun_out/bench_reference_cfg3_camA.json 2>> gpurun_out/bench_reference.err
cat gpurun_out/pytest_gpu.log
for f in gpurun_out/bench_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read())
    print({k:d[k] for k in ("value","ms_per_step","mpixels_per_s","hit_fraction")}, "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "roof", round(d["roofline"]["frac"],4), d["clocks"], d.get("cpu_baseline",{}).get("value"))
except Exception as e: print("ERR", e)
PY
done
for f in gpurun_out/bench_*.err; do tail -n 2 "$f"; done

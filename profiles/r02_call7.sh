#!/usr/bin/env bash
# GPU session 7 of round 2 (one B200): full GPU test suite, compute-sanitizer on the new paths, ncu captures of the final
# kernels, launch list of the default bench command, bench lines (default + reference arm).
set -u
O=gpurun_out/r02
mkdir -p $O
python -m pytest tests -m gpu -q 2>&1 | tail -60 > $O/pytest_gpu_final.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_configs.py -m gpu -q -x -k "rgba16f or tile_mapped or interleaved or handshake or blue_noise" 2>&1 | tail -15 > $O/sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_configs.py -m gpu -q -x -k "rgba16f or handshake" 2>&1 | tail -15 > $O/sanitizer_racecheck.log
NCU="ncu --set full --clock-control none --import-source on -k regex:render_rays -s 1 -c 1 -f"
timeout 600 $NCU -o $O/prof_cfg3A_tiled python profiles/prof_one.py cfg3A tiled > $O/ncu_cfg3A.log 2>&1
timeout 600 $NCU -o $O/prof_cfg3C_tiled python profiles/prof_one.py cfg3C tiled > $O/ncu_cfg3C.log 2>&1
timeout 900 $NCU -o $O/prof_cfg4A_tiled python profiles/prof_one.py cfg4A tiled > $O/ncu_cfg4A.log 2>&1
timeout 900 $NCU -o $O/prof_cfg4C_tiled python profiles/prof_one.py cfg4C tiled > $O/ncu_cfg4C.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_frame_kernel -c 1 -f -o $O/prof_cfg2_frame_rgba16f python - > $O/ncu_frame16.log 2>&1 <<'PY'
import sys, torch
sys.path.insert(0, ".")
import bench
from godot_atmosphere_shader_b200 import abi
R = bench.Runner(torch, bench.Workload(1920, 1080, 32, 0, 0, "B"), 0)
out = torch.empty((1080, 1920, 4), dtype=torch.float16, device="cuda")
R.ctx.render_frame(R.cam, R.d_depth, 1920, 1080, out, None, rgba_format=abi.COLOR_RGBA16F)
torch.cuda.synchronize()
PY
# keep the merge-back under 64 MiB: summaries are extracted here, only the cfg4A capture travels as a .ncu-rep
for f in $O/prof_*.ncu-rep; do python profiles/ncu_summary.py $f > ${f%.ncu-rep}.summary.txt 2>&1; done
python profiles/extract_facts.py $O r02 > $O/roofline_traffic_r02.json 2> $O/extract_facts.err
for f in $O/prof_*.ncu-rep; do case $f in *cfg4A_tiled*) ;; *) rm -f $f ;; esac; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 2 > $O/ncu_launches.log 2>&1
python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "rc=$?" >> $O/bench_n1.err
python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_reference_n1.json 2> $O/bench_reference_n1.err
tail -4 $O/pytest_gpu_final.log; tail -3 $O/sanitizer_memcheck.log; tail -3 $O/sanitizer_racecheck.log; ls $O/*.ncu-rep | wc -l; tail -2 $O/bench_n1.err

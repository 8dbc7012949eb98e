#!/usr/bin/env bash
# GPU session 12 of round 2 (one B200): final kernels — full tests, A/B against the previous build, ncu captures (cfg3A/cfg3C/cfg4A/cfg4C tiled,
# cfg2 linear), launch list, bench lines.
set -u
O=gpurun_out/r02
mkdir -p $O
python -m pytest tests -m gpu -q 2>&1 | tail -8 > $O/pytest_gpu_final5.log
: > $O/tune_clouds6.jsonl
for lib in tune_libs/lib_base4.so godot_atmosphere_shader_b200/libb200atmo.so tune_libs/lib_base4.so godot_atmosphere_shader_b200/libb200atmo.so; do
    B200ATMO_LIB=$lib timeout 300 python profiles/tune_kernels.py --only=cfg2 --only=cfg3A --only=cfg4A --only=cfg4C --only=rm1080A --only=cfg3C >> $O/tune_clouds6.jsonl 2>> $O/tune_clouds6.err
done
NCU="ncu --set full --clock-control none --import-source on -k regex:render_rays -s 1 -c 1 -f"
timeout 600 $NCU -o $O/prof_cfg2_linear python profiles/prof_one.py cfg2 linear > $O/ncu_cfg2.log 2>&1
timeout 600 $NCU -o $O/prof_cfg3A_tiled python profiles/prof_one.py cfg3A tiled > $O/ncu_cfg3A.log 2>&1
timeout 600 $NCU -o $O/prof_cfg3C_tiled python profiles/prof_one.py cfg3C tiled > $O/ncu_cfg3C.log 2>&1
timeout 900 $NCU -o $O/prof_cfg4A_tiled python profiles/prof_one.py cfg4A tiled > $O/ncu_cfg4A.log 2>&1
timeout 900 $NCU -o $O/prof_cfg4C_tiled python profiles/prof_one.py cfg4C tiled > $O/ncu_cfg4C.log 2>&1
for f in $O/prof_*.ncu-rep; do python profiles/ncu_summary.py $f > ${f%.ncu-rep}.summary.txt 2>&1; done
for f in $O/prof_*.ncu-rep; do case $f in *cfg4A_tiled*) ;; *) rm -f $f ;; esac; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 2 > $O/ncu_launches.log 2>&1
python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "rc=$?" >> $O/bench_n1.err
tail -3 $O/pytest_gpu_final5.log; cut -c1-400 $O/tune_clouds6.jsonl; tail -1 $O/bench_n1.err

#!/usr/bin/env bash
# GPU session 3 of round 2 (one B200): second kernel-variant sweep, ncu captures of the shipped kernels, bench line.
set -u
O=gpurun_out/r02
mkdir -p $O
: > $O/tune_clouds2.jsonl
for lib in tune_libs/lib_*.so; do
    B200ATMO_LIB=$lib timeout 300 python profiles/tune_kernels.py --only=cfg3A --only=cfg4A --only=cfg4C --only=rm1080A --only=cfg3C >> $O/tune_clouds2.jsonl 2>> $O/tune_clouds2.err
done
NCU="ncu --set full --clock-control none --import-source on -k regex:render_rays -s 1 -c 1 -f"
timeout 600 $NCU -o $O/prof_cfg2_linear python profiles/prof_one.py cfg2 linear > $O/ncu_cfg2.log 2>&1
timeout 600 $NCU -o $O/prof_cfg3A_tiled python profiles/prof_one.py cfg3A tiled > $O/ncu_cfg3A.log 2>&1
timeout 900 $NCU -o $O/prof_cfg4A_tiled python profiles/prof_one.py cfg4A tiled > $O/ncu_cfg4A.log 2>&1
timeout 900 $NCU -o $O/prof_cfg4A_linear python profiles/prof_one.py cfg4A linear > $O/ncu_cfg4A_lin.log 2>&1
timeout 900 $NCU -o $O/prof_cfg4C_tiled python profiles/prof_one.py cfg4C tiled > $O/ncu_cfg4C.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 2 > $O/ncu_launches.log 2>&1
python bench.py > $O/bench_n1_b.json 2> $O/bench_n1_b.err; echo "rc=$?" >> $O/bench_n1_b.err
wc -l $O/tune_clouds2.jsonl; ls -la $O/*.ncu-rep; tail -2 $O/bench_n1_b.err

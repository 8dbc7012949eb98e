#!/usr/bin/env bash
# GPU session 14 of round 2 (one B200): heaviest-first block dispatch — tests, A/B timing (B200ATMO_BLOCK_ORDER=0/1), bench.
set -u
O=gpurun_out/r02
mkdir -p $O
python -m pytest tests -m gpu -q 2>&1 | tail -25 > $O/pytest_gpu_final6.log
: > $O/tune_block_order.jsonl
for v in 0 1 0 1; do
    B200ATMO_BLOCK_ORDER=$v B200ATMO_LIB=godot_atmosphere_shader_b200/libb200atmo.so timeout 300 python profiles/tune_kernels.py --only=cfg4A --only=cfg4C --only=rm1080A >> $O/tune_block_order.jsonl 2>> $O/tune_block_order.err
done
python bench.py > $O/bench_n1_d.json 2> $O/bench_n1_d.err; echo "rc=$?" >> $O/bench_n1_d.err
tail -4 $O/pytest_gpu_final6.log; cut -c1-330 $O/tune_block_order.jsonl; tail -1 $O/bench_n1_d.err

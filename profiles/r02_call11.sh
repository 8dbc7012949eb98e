#!/usr/bin/env bash
# GPU session 11 of round 2 (one B200): the 6-evaluation light march (first sample reused) against the literal 7, unroll variants; full tests; bench.
set -u
O=gpurun_out/r02
mkdir -p $O
: > $O/tune_clouds5.jsonl
for lib in tune_libs/lib_*.so; do
    B200ATMO_LIB=$lib timeout 300 python profiles/tune_kernels.py --only=cfg3A --only=cfg4A --only=cfg4C --only=rm1080A --only=cfg3C >> $O/tune_clouds5.jsonl 2>> $O/tune_clouds5.err
done
python -m pytest tests -m gpu -q 2>&1 | tail -8 > $O/pytest_gpu_final4.log
python bench.py > $O/bench_n1_c.json 2> $O/bench_n1_c.err; echo "rc=$?" >> $O/bench_n1_c.err
tail -3 $O/pytest_gpu_final4.log; tail -1 $O/bench_n1_c.err

#!/usr/bin/env bash
# GPU session 2 of round 2 (one B200): full GPU test suite, kernel-variant sweep, issue-mix microbenchmark, bench line.
set -u
O=gpurun_out/r02
mkdir -p $O
python -m pytest tests -m gpu -q 2>&1 | tail -25 > $O/pytest_gpu.log
: > $O/tune_clouds.jsonl
for lib in tune_libs/lib_*.so; do
    B200ATMO_LIB=$lib timeout 300 python profiles/tune_kernels.py >> $O/tune_clouds.jsonl 2>> $O/tune_clouds.err
done
(cd profiles/microbench && timeout 300 python issue_mix.py > ../../$O/issue_mix.txt 2>&1)
python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "rc=$?" >> $O/bench_n1.err
tail -4 $O/pytest_gpu.log; wc -l $O/tune_clouds.jsonl; tail -3 $O/issue_mix.txt; tail -2 $O/bench_n1.err

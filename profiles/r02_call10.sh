#!/usr/bin/env bash
# GPU session 10 of round 2 (one B200): centre-out dispatch order of the cloud kernels, A/B against top-to-bottom; full GPU tests.
set -u
O=gpurun_out/r02
mkdir -p $O
: > $O/tune_centerout.jsonl
for lib in tune_libs/lib_topdown.so tune_libs/lib_centerout.so tune_libs/lib_topdown.so tune_libs/lib_centerout.so; do
    B200ATMO_LIB=$lib timeout 300 python profiles/tune_kernels.py --only=cfg3A --only=cfg4A --only=cfg4C --only=rm1080A --only=cfg3C >> $O/tune_centerout.jsonl 2>> $O/tune_centerout.err
done
python -m pytest tests -m gpu -q 2>&1 | tail -8 > $O/pytest_gpu_final3.log
cat $O/tune_centerout.jsonl | cut -c1-700; tail -3 $O/pytest_gpu_final3.log

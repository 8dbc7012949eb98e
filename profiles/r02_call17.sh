#!/usr/bin/env bash
# GPU session 17 of round 2 (one B200): the final build — full GPU suite, smoke(), default bench line.
set -u
O=gpurun_out/r02
mkdir -p $O
python -m pytest tests -m gpu -q 2>&1 | tail -8 > $O/pytest_gpu_final7.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
python bench.py > $O/bench_n1_final.json 2> $O/bench_n1_final.err; echo "rc=$?" >> $O/bench_n1_final.err
tail -n 3 $O/pytest_gpu_final7.log; tail -n 1 $O/smoke.log; tail -n 1 $O/bench_n1_final.err

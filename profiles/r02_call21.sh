#!/usr/bin/env bash
# GPU session 21 of round 2: the default bench line of the final build
O=gpurun_out/r02
mkdir -p $O
timeout 200 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "rc=$?" >> $O/bench_n1.err
tail -c 3000 $O/bench_n1.json; tail -2 $O/bench_n1.err

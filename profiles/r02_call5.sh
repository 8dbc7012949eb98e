#!/usr/bin/env bash
# GPU session 5 of round 2 (TWO B200s): fused multi-GPU test with the completion flags, bench --gpus 2.
set -u
O=gpurun_out/r02
mkdir -p $O
timeout 900 python -m pytest tests/test_multigpu_fused.py -m gpu -q -x 2>&1 | tail -30 > $O/pytest_multigpu_n2b.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 > $O/bench_n2b.json 2> $O/bench_n2b.err; echo "rc=$?" >> $O/bench_n2b.err
tail -5 $O/pytest_multigpu_n2b.log; tail -3 $O/bench_n2b.err; wc -c $O/bench_n2b.json

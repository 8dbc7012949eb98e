#!/usr/bin/env bash
# GPU session 16 of round 2 (EIGHT B200s): strong scaling with the heaviest-first block dispatch — bench --gpus 8 / 4 / 2
# (headline delivery mode + strong-scaling runs; --quick-delivery skips the other delivery modes measured in r02_call8.sh).
set -u
O=gpurun_out/r02
mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 60 --warmup 5 --quick-delivery --e2e-steps 10 > $O/bench_n8_order.json 2> $O/bench_n8_order.err; echo "rc=$?" >> $O/bench_n8_order.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 4 --steps 60 --warmup 5 --quick-delivery --e2e-steps 10 > $O/bench_n4_order.json 2> $O/bench_n4_order.err; echo "rc=$?" >> $O/bench_n4_order.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29524 bench.py --gpus 2 --steps 60 --warmup 5 --quick-delivery --e2e-steps 10 > $O/bench_n2_order.json 2> $O/bench_n2_order.err; echo "rc=$?" >> $O/bench_n2_order.err
tail -1 $O/bench_n8_order.err $O/bench_n4_order.err $O/bench_n2_order.err; wc -c $O/bench_n*_order.json

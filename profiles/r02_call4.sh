#!/usr/bin/env bash
# GPU session 4 of round 2 (TWO B200s): the fused multi-GPU test, bench --gpus 2, and single-GPU kernel variants on GPU 0.
set -u
O=gpurun_out/r02
mkdir -p $O
nvidia-smi topo -m > $O/topo_n2.txt 2>&1
timeout 900 python -m pytest tests/test_multigpu_fused.py -m gpu -q -x 2>&1 | tail -30 > $O/pytest_multigpu_n2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 > $O/bench_n2.json 2> $O/bench_n2.err; echo "rc=$?" >> $O/bench_n2.err
: > $O/tune_clouds3.jsonl
for lib in tune_libs/lib_*.so; do
    CUDA_VISIBLE_DEVICES=0 B200ATMO_LIB=$lib timeout 300 python profiles/tune_kernels.py --only=cfg3A --only=cfg4A --only=cfg4C --only=cfg3C >> $O/tune_clouds3.jsonl 2>> $O/tune_clouds3.err
done
tail -5 $O/pytest_multigpu_n2.log; tail -3 $O/bench_n2.err; wc -c $O/bench_n2.json

#!/usr/bin/env bash
# GPU session 6 of round 2 (TWO B200s): tests with both hand-shakes, bench --gpus 2 (barrier default), NVLink byte counters via ncu.
set -u
O=gpurun_out/r02
mkdir -p $O
timeout 900 python -m pytest tests/test_multigpu_fused.py tests/test_gpu_configs.py -m gpu -q -x -k "multigpu or fused or handshake" 2>&1 | tail -30 > $O/pytest_multigpu_n2c.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 > $O/bench_n2.json 2> $O/bench_n2.err; echo "rc=$?" >> $O/bench_n2.err
timeout 600 ncu --metrics nvltx__bytes.sum,nvltx__bytes_data_user.sum,nvltx__bytes_data_protocol.sum,nvlrx__bytes.sum,nvlrx__bytes_data_user.sum,gpu__time_duration.sum,dram__bytes_write.sum --clock-control none -k regex:render_frame_kernel --devices 0 --csv --log-file $O/nvlink_ncu.csv python profiles/nvlink_probe.py > $O/nvlink_probe.log 2>&1
tail -5 $O/pytest_multigpu_n2c.log; tail -3 $O/bench_n2.err; tail -5 $O/nvlink_probe.log; tail -5 $O/nvlink_ncu.csv | cut -c1-300

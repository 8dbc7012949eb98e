#!/usr/bin/env bash
# GPU session 9 of round 2 (TWO B200s): all GPU tests on GPU 0, the fused multi-GPU test, bench --gpus 2, NVLink counters of the
# peers kernel with the 16x2 warp tiles, smoke().
set -u
O=gpurun_out/r02
mkdir -p $O
CUDA_VISIBLE_DEVICES=0 python -m pytest tests -m gpu -q 2>&1 | tail -30 > $O/pytest_gpu_final2.log
timeout 900 python -m pytest tests/test_multigpu_fused.py -m gpu -q -x 2>&1 | tail -30 > $O/pytest_multigpu_n2d.log
CUDA_VISIBLE_DEVICES=0 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 > $O/bench_n2_final.json 2> $O/bench_n2_final.err; echo "rc=$?" >> $O/bench_n2_final.err
timeout 600 ncu --metrics nvltx__bytes.sum,nvltx__bytes_data_user.sum,nvltx__bytes_data_protocol.sum,nvlrx__bytes.sum,gpu__time_duration.sum --clock-control none -k regex:render_frame_kernel --devices 0 --csv --log-file $O/nvlink_ncu_16x2.csv python profiles/nvlink_probe.py > $O/nvlink_probe_16x2.log 2>&1
tail -4 $O/pytest_gpu_final2.log; tail -3 $O/pytest_multigpu_n2d.log; tail -2 $O/smoke.log; tail -2 $O/bench_n2_final.err; grep -E "nvltx__bytes" $O/nvlink_ncu_16x2.csv | cut -d, -f1,5,13,15

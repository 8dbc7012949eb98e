#!/usr/bin/env bash
# GPU session 8 of round 2 (EIGHT B200s): fused multi-GPU test (world 4), bench --gpus 8 (1080p tiles: BASELINE configs[4] shape),
# bench --gpus 8 with 3840x2160 tiles (configs[4] literally), NVLink byte counters of the peers kernel via ncu (2 GPUs, 1 process).
set -u
O=gpurun_out/r02
mkdir -p $O
nvidia-smi topo -m > $O/topo_n8.txt 2>&1
timeout 600 python -m pytest tests/test_multigpu_fused.py -m gpu -q -x 2>&1 | tail -30 > $O/pytest_multigpu_n8box.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 100 --warmup 5 > $O/bench_n8.json 2> $O/bench_n8.err; echo "rc=$?" >> $O/bench_n8.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 50 --warmup 5 --width 3840 --height 2160 --no-strong --e2e-steps 20 > $O/bench_n8_4k_tiles.json 2> $O/bench_n8_4k_tiles.err; echo "rc=$?" >> $O/bench_n8_4k_tiles.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 4 --steps 100 --warmup 5 > $O/bench_n4.json 2> $O/bench_n4.err; echo "rc=$?" >> $O/bench_n4.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29524 bench.py --gpus 2 --steps 100 --warmup 5 > $O/bench_n2.json 2> $O/bench_n2.err; echo "rc=$?" >> $O/bench_n2.err
CUDA_VISIBLE_DEVICES=0,1 timeout 600 ncu --metrics nvltx__bytes.sum,nvltx__bytes_data_user.sum,nvltx__bytes_data_protocol.sum,nvlrx__bytes.sum,nvlrx__bytes_data_user.sum,gpu__time_duration.sum,dram__bytes_write.sum --clock-control none -k regex:render_frame_kernel --devices 0 --csv --log-file $O/nvlink_ncu_final.csv python profiles/nvlink_probe.py > $O/nvlink_probe_final.log 2>&1
CUDA_VISIBLE_DEVICES=0 timeout 300 python -m pytest tests/test_gpu_configs.py -m gpu -q -x -k handshake 2>&1 | tail -5 > $O/pytest_handshake.log
tail -3 $O/pytest_multigpu_n8box.log; tail -2 $O/bench_n8.err; tail -2 $O/bench_n8_4k_tiles.err; tail -2 $O/bench_n4.err; tail -4 $O/nvlink_probe_final.log; tail -3 $O/pytest_handshake.log; wc -c $O/*.json

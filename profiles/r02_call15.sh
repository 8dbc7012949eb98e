#!/usr/bin/env bash
# GPU session 15 of round 2 (one B200): the block-order test on the final build; A/B of ordering the cheap-light kernels too.
set -u
O=gpurun_out/r02
mkdir -p $O
python -m pytest tests/test_gpu_configs.py -m gpu -q -k "heaviest or handshake or interleaved or rgba16f" 2>&1 | tail -25 > $O/pytest_block_order.log
: > $O/tune_order_cheap.jsonl
for v in 0 2 0 2; do
    B200ATMO_BLOCK_ORDER=$v B200ATMO_LIB=tune_libs/lib_ordercheap.so timeout 300 python profiles/tune_kernels.py --only=cfg3A --only=cfg3C >> $O/tune_order_cheap.jsonl 2>> $O/tune_order_cheap.err
done
tail -4 $O/pytest_block_order.log; cut -c1-330 $O/tune_order_cheap.jsonl

#!/usr/bin/env bash
# GPU session 13 of round 2 (one B200): compute-sanitizer on the round-2 paths of the final build, smoke(), default bench line.
set -u
O=gpurun_out/r02
mkdir -p $O
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_configs.py -m gpu -q -x -k "rgba16f or tile_mapped or interleaved or handshake or blue_noise or table_cache" 2>&1 | tail -12 > $O/sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_configs.py tests/test_gpu_parity.py -m gpu -q -x -k "rgba16f or handshake or peer_store" 2>&1 | tail -12 > $O/sanitizer_racecheck.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "frame_parity and (rm128 or clouds_high_rm or odd_counts or scatter32_clouds64)" 2>&1 | tail -8 > $O/sanitizer_memcheck_clouds.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
python bench.py > $O/bench_n1_final.json 2> $O/bench_n1_final.err; echo "rc=$?" >> $O/bench_n1_final.err
tail -4 $O/sanitizer_memcheck.log; tail -4 $O/sanitizer_racecheck.log; tail -4 $O/sanitizer_memcheck_clouds.log; tail -1 $O/smoke.log; tail -1 $O/bench_n1_final.err

#!/usr/bin/env bash
# GPU session 22 of round 2: compute-sanitizer synccheck of the warp reductions in the cloud march (ragged frames, non-finite rays)
O=gpurun_out/r02
mkdir -p $O
timeout 75 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "tiny_and_ragged or non_finite or camera_below" 2>&1 | tail -12 > $O/sanitizer_synccheck_clouds.log
echo "rc=$?"; cat $O/sanitizer_synccheck_clouds.log

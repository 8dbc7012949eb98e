#!/usr/bin/env bash
# session 19: under-shell skip of the cloud march (current build) against the previous build (tune_libs/lib_prev.so), A/B/A/B
mkdir -p gpurun_out
for rep in 1 2; do
  for lib in tune_libs/lib_prev.so godot_atmosphere_shader_b200/libb200atmo.so; do
    B200ATMO_LIB=$PWD/$lib timeout 200 python profiles/tune_kernels.py >> gpurun_out/tune_under_skip.txt 2>&1
  done
done
cat gpurun_out/tune_under_skip.txt

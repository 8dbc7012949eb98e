#!/usr/bin/env bash
# session 18: texture-footprint what-if of the cloud kernels (shipped library)
mkdir -p gpurun_out
timeout 300 python profiles/whatif_textures.py > gpurun_out/whatif_textures.txt 2>&1
echo "rc=$?"
cat gpurun_out/whatif_textures.txt

#!/usr/bin/env bash
# Builds libb200atmo.so (sm_100a) in-tree: godot_atmosphere_shader_b200/libb200atmo.so
#   -fmad=false + -ffp-contract=off : see the numeric policy in atmo_device.cuh
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="${B200ATMO_OUT:-${HERE}/../libb200atmo.so}"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false
       -Xcompiler -fPIC,-ffp-contract=off,-fno-fast-math,-Wall -Xptxas -v)
"${NVCC}" "${FLAGS[@]}" -shared -o "${OUT}" "${HERE}/atmo_kernels.cu" "${HERE}/atmo_capi.cu" "$@"
echo "built ${OUT}"

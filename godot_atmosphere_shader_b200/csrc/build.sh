#!/usr/bin/env bash
# Builds in-tree:
#   godot_atmosphere_shader_b200/libb200atmo.so       the C-ABI library (sm_100a kernels), nvcc
#   godot_atmosphere_shader_b200/libb200atmo_node.so  the engine-independent C++ core of the PlanetAtmosphere node, g++
#   -fmad=false + -ffp-contract=off : see the numeric policy in atmo_device.cuh
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="${B200ATMO_OUT:-${HERE}/../libb200atmo.so}"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
CXX="${CXX:-g++}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false
       -Xcompiler -fPIC,-ffp-contract=off,-fno-fast-math,-Wall -Xptxas -v)
"${NVCC}" "${FLAGS[@]}" -shared -o "${OUT}" "${HERE}/atmo_kernels.cu" "${HERE}/atmo_capi.cu" "$@"
echo "built ${OUT}"
if [ -z "${B200ATMO_OUT:-}" ]; then
    NODE_OUT="${HERE}/../libb200atmo_node.so"
    "${CXX}" -O2 -std=c++17 -fPIC -Wall -Wextra -shared -I"${HERE}/../../include" -o "${NODE_OUT}" \
        "${HERE}/node/planet_atmosphere_node.cpp" -L"${HERE}/.." -lb200atmo -Wl,-rpath,'$ORIGIN'
    echo "built ${NODE_OUT}"
fi

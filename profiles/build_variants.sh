#!/usr/bin/env bash
# Builds kernel-tuning variants of libb200atmo.so into tune_libs/ (git-ignored; travels to the GPU box with gpurun).
# usage: profiles/build_variants.sh name1:"-DFLAG1 -DFLAG2" name2:"..." ...
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
mkdir -p "${HERE}/../tune_libs"
for spec in "$@"; do
    name="${spec%%:*}"; flags="${spec#*:}"
    ( B200ATMO_OUT="${HERE}/../tune_libs/lib_${name}.so" bash "${HERE}/../godot_atmosphere_shader_b200/csrc/build.sh" ${flags} > "${HERE}/../tune_libs/build_${name}.log" 2>&1 \
      && echo "built ${name} (${flags})" || { echo "FAILED ${name}"; tail -5 "${HERE}/../tune_libs/build_${name}.log"; } ) &
done
wait

#!/usr/bin/env bash
# Builds tuning variants of libb200atmo.so (here, no GPU needed) ...   profiles/tune_scatter.sh build
# ... and benches them on the GPU box ...                               profiles/tune_scatter.sh run
set -u
cd "$(dirname "$0")/.."
VARIANTS=("default:-DB200ATMO_SCATTER_UNROLL=4"
          "u2_default:-DB200ATMO_SCATTER_UNROLL=2"
          "u8_default:-DB200ATMO_SCATTER_UNROLL=8"
          "u4_b256_default:-DB200ATMO_SCATTER_UNROLL=4 -DB200ATMO_BLOCK=256"
          "u4_b128_m12:-DB200ATMO_SCATTER_UNROLL=4 -DB200ATMO_BLOCK=128 -DB200ATMO_MIN_BLOCKS=12"
          "u4_b128_m1:-DB200ATMO_SCATTER_UNROLL=4 -DB200ATMO_BLOCK=128 -DB200ATMO_MIN_BLOCKS=1"
          "u2_b128_m1:-DB200ATMO_SCATTER_UNROLL=2 -DB200ATMO_BLOCK=128 -DB200ATMO_MIN_BLOCKS=1"
          "u8_b128_m1:-DB200ATMO_SCATTER_UNROLL=8 -DB200ATMO_BLOCK=128 -DB200ATMO_MIN_BLOCKS=1"
          "u4_b128_m16:-DB200ATMO_SCATTER_UNROLL=4 -DB200ATMO_BLOCK=128 -DB200ATMO_MIN_BLOCKS=16"
          "u4_b256_m1:-DB200ATMO_SCATTER_UNROLL=4 -DB200ATMO_BLOCK=256 -DB200ATMO_MIN_BLOCKS=1"
          "u4_b64_m1:-DB200ATMO_SCATTER_UNROLL=4 -DB200ATMO_BLOCK=64 -DB200ATMO_MIN_BLOCKS=1"
          "u4_b128_m8:-DB200ATMO_SCATTER_UNROLL=4 -DB200ATMO_BLOCK=128 -DB200ATMO_MIN_BLOCKS=8"
          "u1_b128_m1:-DB200ATMO_SCATTER_UNROLL=1 -DB200ATMO_BLOCK=128 -DB200ATMO_MIN_BLOCKS=1")
mkdir -p tune_libs gpurun_out
if [ "${1:-build}" = "build" ]; then
  for v in "${VARIANTS[@]}"; do
    name="${v%%:*}"; flags="${v#*:}"
    B200ATMO_OUT="$PWD/tune_libs/lib_${name}.so" godot_atmosphere_shader_b200/csrc/build.sh $flags 2>&1 | grep -E "render_rays_kernelILi0ELi0" -A3 | grep -E "registers" | sed "s/^/${name}: /"
  done
else
  for f in tune_libs/lib_*.so; do
    B200ATMO_LIB="$PWD/$f" timeout 120 python bench.py --no-cpu-baseline --e2e-steps 2 --steps 300 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$f', round(d['ms_per_step']*1e3,2),'us', '%.3e'%d['value'])"
  done | tee gpurun_out/tune_scatter.txt
fi

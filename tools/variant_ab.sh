#!/usr/bin/env bash
# Parity + same-box A/B of a build variant of the step kernels (robot_lab_b200/build.py VARIANTS) against the default
# library. Build the variant HERE first (nvcc cross-compiles, the .so travels with the snapshot):
#
#   python -m robot_lab_b200.build --variant shared_norms
#   gpurun --timeout 400 -- 'bash tools/variant_ab.sh shared_norms'
#
# Output: gpurun_out/variant_<name>_tests.log (the whole -m gpu suite run against the variant library) and
# gpurun_out/variant_<name>_ab.log (tools/launch_breakdown.py --short, default / variant alternating, two rounds).
set -uo pipefail
name="${1:?variant name}"
lib="$PWD/robot_lab_b200/_lib/libmdpstep_${name}.so"
[ -f "$lib" ] || { echo "missing $lib: build it first (python -m robot_lab_b200.build --variant $name)"; exit 2; }
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
RL_MDPSTEP_LIB="$lib" timeout 240 python -m pytest tests -m gpu -x -q -p no:cacheprovider > "gpurun_out/variant_${name}_tests.log" 2>&1
echo "rc=$?" >> "gpurun_out/variant_${name}_tests.log"
tail -n 4 "gpurun_out/variant_${name}_tests.log"
ab="gpurun_out/variant_${name}_ab.log"
: > "$ab"
for round in 1 2; do
  for which in default "$name"; do
    echo "== $which" >> "$ab"
    if [ "$which" = default ]; then
      timeout 90 python tools/launch_breakdown.py "${2:-4096}" 16 go2_rough 32 --short >> "$ab" 2>&1
    else
      RL_MDPSTEP_LIB="$lib" timeout 90 python tools/launch_breakdown.py "${2:-4096}" 16 go2_rough 32 --short >> "$ab" 2>&1
    fi
  done
done
cat "$ab"

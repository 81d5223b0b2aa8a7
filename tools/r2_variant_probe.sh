#!/usr/bin/env bash
# Round-2 first GPU call: same-box timing of the default library against the build variants written at the end of
# round 1 (shared = norms + ctx prepass, persistent = tile loop) at 4096 / 16384 / 65536 envs.
# Output: gpurun_out/r2_variant_probe.log
set -uo pipefail
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
out=gpurun_out/r2_variant_probe.log
: > "$out"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader >> "$out" 2>&1
for n in 4096 16384 65536; do
  for which in default shared persistent default; do
    echo "== $which N=$n" >> "$out"
    if [ "$which" = default ]; then
      timeout 120 python tools/launch_breakdown.py "$n" 16 go2_rough 32 --short >> "$out" 2>&1
    else
      RL_MDPSTEP_LIB="$PWD/robot_lab_b200/_lib/libmdpstep_${which}.so" timeout 120 python tools/launch_breakdown.py "$n" 16 go2_rough 32 --short >> "$out" 2>&1
    fi
  done
done
cat "$out"

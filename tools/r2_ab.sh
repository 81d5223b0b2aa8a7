#!/usr/bin/env bash
# Same-box A/B of two builds of the library: RL_MDPSTEP_LIB=<other .so> against the default, interleaved.
# usage: bash tools/r2_ab.sh <other.so> [out]
set -uo pipefail
other="$1"; out="${2:-gpurun_out/r2_ab.log}"
mkdir -p gpurun_out; : > "$out"
export PYTHONUNBUFFERED=1
for rep in 1 2; do
  for n in 4096 65536; do
    for lib in default "$other"; do
      echo "== rep=$rep lib=$lib N=$n" >> "$out"
      if [ "$lib" = default ]; then
        timeout 120 python tools/launch_breakdown.py "$n" 16 go2_rough 32 --short 2>&1 | grep -E "pre-reset|post-reset|env step" >> "$out"
      else
        RL_MDPSTEP_LIB="$lib" timeout 120 python tools/launch_breakdown.py "$n" 16 go2_rough 32 --short 2>&1 | grep -E "pre-reset|post-reset|env step" >> "$out"
      fi
    done
  done
done
cat "$out"

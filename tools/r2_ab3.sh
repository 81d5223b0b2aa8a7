#!/usr/bin/env bash
# Same-box A/B of builds of the library (default first), interleaved: usage: bash tools/r2_ab3.sh out.log lib1.so lib2.so ...
set -uo pipefail
out="$1"; shift
mkdir -p gpurun_out; : > "$out"
export PYTHONUNBUFFERED=1
for rep in 1 2; do
  for spec in "4096 1x1x16" "16384 1x1x8" "65536 1x1x8" "65536 1x1x16"; do
    set -- $spec "$@"; n=$1; cfg=$2; shift 2
    for lib in default "$@"; do
      echo "== rep=$rep lib=$lib N=$n cfg=$cfg" >> "$out"
      if [ "$lib" = default ]; then
        RL_MDPSTEP_V2_CFG=$cfg timeout 120 python tools/launch_breakdown.py "$n" 16 go2_rough 32 --short 2>&1 | grep -E "pre-reset|post-reset|env step" >> "$out"
      else
        RL_MDPSTEP_V2_CFG=$cfg RL_MDPSTEP_LIB="$lib" timeout 120 python tools/launch_breakdown.py "$n" 16 go2_rough 32 --short 2>&1 | grep -E "pre-reset|post-reset|env step" >> "$out"
      fi
    done
  done
done
cat "$out"

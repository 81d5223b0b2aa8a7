#!/usr/bin/env bash
# Third GPU session of round 2: whole -m gpu suite; A/B of unrolled vs rolled term loops, PDL on/off, at 4096 / 65536.
set -uo pipefail
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
t=gpurun_out/r2_gpu_tests3.log
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > "$t" 2>&1
echo "rc=$?" >> "$t"
tail -n 25 "$t"
out=gpurun_out/r2_v2_ab3.log
: > "$out"
rolled="$PWD/robot_lab_b200/_lib/libmdpstep_rolled.so"
for n in 4096 65536; do
  for cfg in 1x1x16 1x1x8 2x2x16; do
    for lib in default rolled; do
      echo "== lib=$lib cfg=$cfg N=$n" >> "$out"
      if [ "$lib" = default ]; then
        RL_MDPSTEP_V2_CFG=$cfg timeout 120 python tools/launch_breakdown.py "$n" 16 go2_rough 32 --short --pdl >> "$out" 2>&1
      else
        RL_MDPSTEP_LIB="$rolled" RL_MDPSTEP_V2_CFG=$cfg timeout 120 python tools/launch_breakdown.py "$n" 16 go2_rough 32 --short >> "$out" 2>&1
      fi
    done
  done
done
cat "$out"

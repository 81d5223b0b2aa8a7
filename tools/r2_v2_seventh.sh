#!/usr/bin/env bash
set -uo pipefail
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
t=gpurun_out/r2_gpu_tests7.log
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > "$t" 2>&1
echo "rc=$?" >> "$t"
tail -n 6 "$t"
out=gpurun_out/r2_v2_cfg_timing7.log
: > "$out"
for n in 4096 65536; do
  for cfg in off 1x1x16 1x1x8 2x2x16 4x4x16; do
    echo "== cfg=$cfg N=$n" >> "$out"
    if [ "$cfg" = off ]; then
      RL_MDPSTEP_V2=0 timeout 120 python tools/launch_breakdown.py "$n" 16 go2_rough 32 --short >> "$out" 2>&1
    else
      RL_MDPSTEP_V2_CFG=$cfg timeout 120 python tools/launch_breakdown.py "$n" 16 go2_rough 32 --short >> "$out" 2>&1
    fi
  done
done
cat "$out"
out=gpurun_out/r2_v2_timeline7.log
: > "$out"
L="$PWD/robot_lab_b200/_lib/libmdpstep_stamps.so"
RL_MDPSTEP_LIB="$L" timeout 200 python tools/v2_timeline.py 4096 >> "$out" 2>&1
RL_MDPSTEP_V2_CFG=1x1x8 RL_MDPSTEP_LIB="$L" timeout 200 python tools/v2_timeline.py 65536 >> "$out" 2>&1
cat "$out"

#!/usr/bin/env bash
set -uo pipefail
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
t=gpurun_out/r2_gpu_tests5.log
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > "$t" 2>&1
echo "rc=$?" >> "$t"
tail -n 15 "$t"
out=gpurun_out/r2_v2_cfg_timing5.log
: > "$out"
for n in 4096 16384 65536; do
  for cfg in 1x1x16 1x1x8 2x2x16 4x4x16; do
    echo "== cfg=$cfg N=$n" >> "$out"
    RL_MDPSTEP_V2_CFG=$cfg timeout 120 python tools/launch_breakdown.py "$n" 16 go2_rough 32 --short >> "$out" 2>&1
  done
done
cat "$out"
out=gpurun_out/r2_v2_timeline5.log
: > "$out"
RL_MDPSTEP_LIB="$PWD/robot_lab_b200/_lib/libmdpstep_stamps.so" timeout 200 python tools/v2_timeline.py 4096 >> "$out" 2>&1
RL_MDPSTEP_V2_CFG=4x4x16 RL_MDPSTEP_LIB="$PWD/robot_lab_b200/_lib/libmdpstep_stamps.so" timeout 200 python tools/v2_timeline.py 4096 >> "$out" 2>&1
RL_MDPSTEP_LIB="$PWD/robot_lab_b200/_lib/libmdpstep_stamps.so" timeout 200 python tools/v2_timeline.py 65536 >> "$out" 2>&1
cat "$out"

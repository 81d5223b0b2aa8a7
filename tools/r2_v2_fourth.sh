#!/usr/bin/env bash
set -uo pipefail
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
t=gpurun_out/r2_gpu_tests4.log
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > "$t" 2>&1
echo "rc=$?" >> "$t"
tail -n 15 "$t"
out=gpurun_out/r2_v2_timeline.log
: > "$out"
for n in 4096 65536; do
  RL_MDPSTEP_LIB="$PWD/robot_lab_b200/_lib/libmdpstep_stamps.so" timeout 200 python tools/v2_timeline.py $n >> "$out" 2>&1
done
RL_MDPSTEP_V2_CFG=4x4x16 RL_MDPSTEP_LIB="$PWD/robot_lab_b200/_lib/libmdpstep_stamps.so" timeout 200 python tools/v2_timeline.py 4096 >> "$out" 2>&1
cat "$out"
bash tools/r2_sanitizer.sh > /dev/null 2>&1
cat gpurun_out/r2_sanitizer.txt

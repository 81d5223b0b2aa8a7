#!/usr/bin/env bash
set -uo pipefail
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_v2_parity.py tests/test_gpu_step_parity.py -x -q -p no:cacheprovider 2>&1 | tail -3
out=gpurun_out/r2_timeline13.log
: > "$out"
for n in 4096 65536; do
  RL_MDPSTEP_LIB=robot_lab_b200/_lib/libmdpstep_stamps.so timeout 200 python tools/v2_timeline.py $n >> "$out" 2>&1
done
cat "$out"
out=gpurun_out/r2_v2_cfg_timing13.log
: > "$out"
for n in 4096 16384 65536; do
  echo "== N=$n" >> "$out"
  timeout 120 python tools/launch_breakdown.py "$n" 16 go2_rough 32 --short >> "$out" 2>&1
done
grep -E "==|pre-reset|post-reset|env step" "$out"

#!/usr/bin/env bash
set -uo pipefail
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
t=gpurun_out/r2_gpu_tests18.log
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > "$t" 2>&1
echo "rc=$?" >> "$t"
tail -n 8 "$t"
out=gpurun_out/r2_timeline19.log
: > "$out"
RL_MDPSTEP_LIB=robot_lab_b200/_lib/libmdpstep_stamps.so timeout 200 python tools/v2_timeline.py 4096 >> "$out" 2>&1
RL_MDPSTEP_LIB=robot_lab_b200/_lib/libmdpstep_stamps.so timeout 200 python tools/v2_timeline.py 65536 >> "$out" 2>&1
grep -v "^     \|graph of\|^   pre  \|^   post  \|process_action" "$out"
out=gpurun_out/r2_v2_cfg_timing19.log
: > "$out"
for n in 4096 16384 65536; do
  echo "== N=$n" >> "$out"
  timeout 120 python tools/launch_breakdown.py "$n" 16 go2_rough 32 --short >> "$out" 2>&1
done
grep -E "==|pre-reset|post-reset|env step" "$out"

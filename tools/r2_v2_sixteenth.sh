#!/usr/bin/env bash
set -uo pipefail
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
out=gpurun_out/r2_timeline20.log
: > "$out"
RL_MDPSTEP_LIB=robot_lab_b200/_lib/libmdpstep_stamps.so timeout 200 python tools/v2_timeline.py 4096 --warm >> "$out" 2>&1
RL_MDPSTEP_LIB=robot_lab_b200/_lib/libmdpstep_stamps.so timeout 200 python tools/v2_timeline.py 4096 >> "$out" 2>&1
for k in a1_flat g1_rough; do
RL_MDPSTEP_LIB=robot_lab_b200/_lib/libmdpstep_stamps.so timeout 200 python tools/v2_timeline.py 4096 $k --warm 2>&1 | grep -E "^go2|^a1|^g1|---|per-warp|per-task|slowest" >> "$out"
done
grep -v "^     \|graph of\|^   pre  \|^   post  \|process_action" "$out"

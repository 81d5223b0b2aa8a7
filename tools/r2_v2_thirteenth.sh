#!/usr/bin/env bash
set -uo pipefail
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_v2_parity.py -x -q -p no:cacheprovider 2>&1 | tail -3
out=gpurun_out/r2_timeline14.log
: > "$out"
RL_MDPSTEP_LIB=robot_lab_b200/_lib/libmdpstep_stamps.so timeout 200 python tools/v2_timeline.py 4096 >> "$out" 2>&1
RL_MDPSTEP_LIB=robot_lab_b200/_lib/libmdpstep_stamps.so timeout 200 python tools/v2_timeline.py 4096 --warm >> "$out" 2>&1
RL_MDPSTEP_LIB=robot_lab_b200/_lib/libmdpstep_stamps.so timeout 200 python tools/v2_timeline.py 65536 >> "$out" 2>&1
grep -v "^     \|graph of\|^   pre  \|^   post  \|process_action" "$out"

#!/usr/bin/env bash
# Source-level ncu summary of the final step kernels at 4096 envs (cold launches of tools/ncu_targets.py).
set -uo pipefail
d=gpurun_out/final
mkdir -p "$d"
export PYTHONUNBUFFERED=1
timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -k 'regex:v2_pre_kernel|v2_post_kernel' \
    -f -o /tmp/full_4096_src python tools/ncu_targets.py 4096 > "$d/full_4096_src.log" 2>&1
{ echo "ncu --set full --import-source on, tools/ncu_targets.py 4096 (final round-2 build), aggregated by tools/ncu_source_summary.py"; echo;
  python tools/ncu_source_summary.py /tmp/full_4096_src.ncu-rep regex:v2_pre 22; echo;
  python tools/ncu_source_summary.py /tmp/full_4096_src.ncu-rep regex:v2_post 22; } > "$d/ncu_source_4096.txt" 2>&1
head -30 "$d/ncu_source_4096.txt" | cut -c1-160

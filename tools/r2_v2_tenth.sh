#!/usr/bin/env bash
# Same-box sweep of the one-tile configurations at large env counts, one `ncu --set full` capture of the two step kernels
# at 65536 envs (default configuration), then the sanitizer pass.
set -uo pipefail
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
out=gpurun_out/r2_v2_cfg_timing10.log
: > "$out"
for n in 16384 65536; do
  for cfg in 1x1x16 1x1x8 1x1x4; do
    echo "== cfg=$cfg N=$n" >> "$out"
    RL_MDPSTEP_V2_CFG=$cfg timeout 120 python tools/launch_breakdown.py "$n" 16 go2_rough 32 --short >> "$out" 2>&1
  done
done
cat "$out"
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k 'regex:v2_pre_kernel|v2_post_kernel' \
    -f -o gpurun_out/r2_full_65536 python tools/ncu_targets.py 65536 > gpurun_out/r2_full_65536.log 2>&1
echo "ncu rc=$?"
ncu -i gpurun_out/r2_full_65536.ncu-rep --page raw --csv > gpurun_out/r2_full_65536_raw.csv 2>/dev/null
ls -la gpurun_out/r2_full_65536*
bash tools/r2_sanitizer.sh > /dev/null 2>&1
grep -E "^===|rc=|SUMMARY" gpurun_out/r2_sanitizer.txt | paste - - - | cut -c1-220

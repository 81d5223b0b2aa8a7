#!/usr/bin/env bash
# Second GPU session of round 2: the whole -m gpu suite on the new default (cluster kernels on), same-box timing of
# every compiled configuration at three env counts, one ncu --set full capture of the default configuration.
set -uo pipefail
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
t=gpurun_out/r2_gpu_tests.log
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > "$t" 2>&1
echo "rc=$?" >> "$t"
tail -n 30 "$t"
out=gpurun_out/r2_v2_cfg_timing.log
: > "$out"
for n in 4096 16384 65536; do
  for cfg in off 1x1x16 1x1x8 1x1x4 2x2x16 4x4x16; do
    echo "== cfg=$cfg N=$n" >> "$out"
    if [ "$cfg" = off ]; then
      RL_MDPSTEP_V2=0 timeout 120 python tools/launch_breakdown.py "$n" 16 go2_rough 32 --short >> "$out" 2>&1
    else
      RL_MDPSTEP_V2_CFG=$cfg timeout 120 python tools/launch_breakdown.py "$n" 16 go2_rough 32 --short >> "$out" 2>&1
    fi
  done
done
cat "$out"
for n in 4096 65536; do
  timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off \
      -k 'regex:v2_pre|v2_post|process_action' -o gpurun_out/r2_v2_full_$n -f python tools/ncu_targets.py $n > gpurun_out/r2_v2_ncu_$n.log 2>&1
  tail -n 3 gpurun_out/r2_v2_ncu_$n.log
done
ls -la gpurun_out/*.ncu-rep

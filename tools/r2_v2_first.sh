#!/usr/bin/env bash
# First GPU contact of the cluster kernels: parity against the general kernel, then same-box timing of every
# configuration. Output: gpurun_out/r2_v2_first_tests.log, gpurun_out/r2_v2_first_timing.log
set -uo pipefail
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
t=gpurun_out/r2_v2_first_tests.log
timeout 420 python -m pytest tests/test_gpu_v2_parity.py -x -q -p no:cacheprovider > "$t" 2>&1
echo "rc=$?" >> "$t"
tail -n 40 "$t"
out=gpurun_out/r2_v2_first_timing.log
: > "$out"
for n in 4096 65536; do
  for cfg in off 4x4 2x2 1x1; do
    echo "== cfg=$cfg N=$n" >> "$out"
    if [ "$cfg" = off ]; then
      RL_MDPSTEP_V2=0 timeout 120 python tools/launch_breakdown.py "$n" 16 go2_rough 32 --short >> "$out" 2>&1
    else
      RL_MDPSTEP_V2_CFG=$cfg timeout 120 python tools/launch_breakdown.py "$n" 16 go2_rough 32 --short >> "$out" 2>&1
    fi
  done
done
cat "$out"

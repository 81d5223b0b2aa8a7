#!/usr/bin/env bash
# compute-sanitizer (memcheck, racecheck, synccheck) over the hand-rolled inter-warp / inter-CTA protocols of the step
# kernels: the cluster kernels (mbarrier + TMA, DSMEM + cluster barriers, early release tickets), the general kernel's
# build-time specialised and generic instantiations (named-barrier tails, early tickets), a ragged env count.
# Output: gpurun_out/r2_sanitizer.txt (summary lines of every run)
set -uo pipefail
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
out=gpurun_out/r2_sanitizer.txt
: > "$out"
run() {   # tool, label, env assignments..., -- pytest selection
  local tool="$1" label="$2"; shift 2
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  echo "=== $tool :: $label :: ${envs[*]:-} :: $*" >> "$out"
  env "${envs[@]}" timeout 600 compute-sanitizer --tool "$tool" --error-exitcode 86 --print-limit 20 \
      python -m pytest "$@" -x -q -p no:cacheprovider > gpurun_out/san_tmp.log 2>&1
  local rc=$?
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|error" gpurun_out/san_tmp.log | tail -n 6 >> "$out"
  grep -E "^========= (Error|Race|Barrier|Invalid|Warning)" gpurun_out/san_tmp.log | sort | uniq -c | head -n 12 >> "$out"
  if [ $rc -ne 0 ]; then   # the full text of the first hazards (both accesses)
    grep -A 8 -E "^========= (Error|Warning)" gpurun_out/san_tmp.log | head -n 40 >> "$out"
  fi
  echo "rc=$rc" >> "$out"
}
for tool in memcheck racecheck synccheck; do
  run $tool "new kernels, one tile per CTA" RL_MDPSTEP_V2_CFG=1x1x16 -- "tests/test_gpu_step_parity.py::test_two_launch_step_in_reference_order"
  run $tool "new kernels, one tile per CTA, 8 warps" RL_MDPSTEP_V2_CFG=1x1x8 -- "tests/test_gpu_step_parity.py::test_two_launch_step_in_reference_order"
  run $tool "new kernels, 4-CTA cluster (DSMEM, cluster barriers)" RL_MDPSTEP_V2_CFG=4x4x16 -- "tests/test_gpu_step_parity.py::test_two_launch_step_in_reference_order"
  run $tool "general kernel, baked" RL_MDPSTEP_V2=0 -- "tests/test_gpu_step_parity.py::test_two_launch_step_in_reference_order"
  run $tool "general kernel, generic" RL_MDPSTEP_V2=0 RL_MDPSTEP_GENERIC=1 -- "tests/test_gpu_step_parity.py::test_two_launch_step_in_reference_order"
  run $tool "ragged env counts (general kernel)" RL_X=1 -- "tests/test_gpu_step_parity.py::test_ragged_env_counts"
done
cat "$out"

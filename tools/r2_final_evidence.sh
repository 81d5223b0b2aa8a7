#!/usr/bin/env bash
# Final evidence of the round on one GPU: GPU tests, the bench lines (default flags, the driver's flags, the other BASELINE
# configs, the reference arm), the ncu launch list of bench.py, one `ncu --set full` capture of every kernel at 4096 and
# 65536 envs (cold, tools/ncu_targets.py), the sanitizer pass. Everything lands in gpurun_out/final/.
set -uo pipefail
d=gpurun_out/final
mkdir -p "$d"
export PYTHONUNBUFFERED=1
if [ "${SKIP_TESTS:-0}" != 1 ]; then
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > "$d/gpu_tests.log" 2>&1
echo "rc=$?" >> "$d/gpu_tests.log"
tail -n 4 "$d/gpu_tests.log"
fi
python bench.py > "$d/bench_n1.json" 2> "$d/bench_n1.err"; echo "bench rc=$?"
python bench.py --gpus 1 --steps 20 --warmup 5 > "$d/bench_n1_driver_flags.json" 2> "$d/bench_n1_driver_flags.err"; echo "bench(driver flags) rc=$?"
bash tools/ncu_launch_list.sh "$d/launches.csv" > "$d/launches_summary.txt" 2>&1; tail -n 2 "$d/launches_summary.txt"
for n in 4096 65536; do
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
      -k 'regex:v2_pre_kernel|v2_post_kernel|process_action_kernel' -f -o "$d/full_$n" python tools/ncu_targets.py $n > "$d/full_$n.log" 2>&1
  ncu -i "$d/full_$n.ncu-rep" --page raw --csv > "$d/full_${n}_raw.csv" 2>/dev/null
  rm -f "$d/full_$n.ncu-rep"   # 20 MB each: gpurun_out/ travels back only below 64 MiB
done
ls -la "$d"/full_* | head
for t in RobotLab-Isaac-Velocity-Flat-Unitree-Go2-v0 RobotLab-Isaac-Velocity-Rough-Unitree-G1-v0 RobotLab-Isaac-Velocity-Rough-Unitree-G1-37dof-v0; do
  python bench.py --task "$t" --no-cpu-baseline --large-n 0 --no-neighbours > "$d/bench_n1_$t.json" 2> "$d/bench_n1_$t.err"; echo "bench $t rc=$?"
done
[ "${SKIP_REF:-0}" = 1 ] || python bench.py --impl reference --steps 20 --warmup 3 > "$d/bench_reference_arm.json" 2> "$d/bench_reference_arm.err"; echo "reference arm rc=$?"
bash tools/r2_sanitizer.sh > /dev/null 2>&1
cp gpurun_out/r2_sanitizer.txt "$d/sanitizer.txt"
grep -E "^===|rc=" "$d/sanitizer.txt" | paste - - | cut -c1-200
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/final/bench_*.json")):
    try:
        d = json.load(open(f))
    except Exception as e:
        print(f, "unreadable", e); continue
    r = d.get("roofline") or {}
    print(f.split("/")[-1], d.get("value"), d.get("ms_per_step"), d.get("steps"), "roofline", r.get("kernel_us"), r.get("frac"),
          (r.get("other_kernel") or {}).get("kernel_us"), "e2e", (d.get("e2e") or {}).get("value"))
    l = d.get("roofline_large_n")
    if l: print("   large", l["kernel"], l["kernel_us"], l["frac"], l["other_kernel"]["kernel_us"], l["other_kernel"]["frac"], l["env_step_us"])
PY

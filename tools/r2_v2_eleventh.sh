#!/usr/bin/env bash
set -uo pipefail
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
t=gpurun_out/r2_gpu_tests11.log
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > "$t" 2>&1
echo "rc=$?" >> "$t"
tail -n 6 "$t"
out=gpurun_out/r2_v2_cfg_timing11.log
: > "$out"
for n in 4096 16384 65536; do
  for cfg in 1x1x16 1x1x8; do
    echo "== cfg=$cfg N=$n" >> "$out"
    RL_MDPSTEP_V2_CFG=$cfg timeout 120 python tools/launch_breakdown.py "$n" 16 go2_rough 32 --short >> "$out" 2>&1
  done
done
grep -E "==|pre-reset|post-reset|env step" "$out"
python bench.py > gpurun_out/r2_bench_n1d.json 2> gpurun_out/r2_bench_n1d.err
tail -2 gpurun_out/r2_bench_n1d.err
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_n1d_driver_args.json 2> gpurun_out/r2_bench_n1d_driver_args.err
python - <<'PY'
import json
for f in ("gpurun_out/r2_bench_n1d.json", "gpurun_out/r2_bench_n1d_driver_args.json"):
    d = json.load(open(f))
    print(f, {k: d[k] for k in ("value", "ms_per_step", "steps", "warmup", "gpu_launches")})
    r = d["roofline"]; print("  roofline", r["kernel_us"], r["frac"], r["other_kernel"]["kernel_us"], r["other_kernel"]["frac"])
    l = d.get("roofline_large_n")
    if l: print("  large", l["kernel"], l["kernel_us"], l["frac"], l["other_kernel"]["kernel_us"], l["other_kernel"]["frac"], l["env_step_us"])
    e = d["e2e"]; print("  e2e", e["value"], e.get("frac_of_pcie"), e.get("api"))
PY
source tools/r2_sanitizer.sh > /dev/null 2>&1 || true

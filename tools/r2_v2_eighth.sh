#!/usr/bin/env bash
set -uo pipefail
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
t=gpurun_out/r2_gpu_tests8.log
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > "$t" 2>&1
echo "rc=$?" >> "$t"
tail -n 6 "$t"
out=gpurun_out/r2_v2_cfg_timing8.log
: > "$out"
for n in 4096 16384 65536; do
  for cfg in 1x1x16 1x1x8; do
    echo "== cfg=$cfg N=$n" >> "$out"
    RL_MDPSTEP_V2_CFG=$cfg timeout 120 python tools/launch_breakdown.py "$n" 16 go2_rough 32 --short >> "$out" 2>&1
  done
done
cat "$out"
python bench.py > gpurun_out/r2_bench_n1b.json 2> gpurun_out/r2_bench_n1b.err
tail -2 gpurun_out/r2_bench_n1b.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench_n1b.json"))
print({k: d[k] for k in ("value", "ms_per_step")})
r = d["roofline"]; print("roofline", r["kernel_us"], r["frac"], r["other_kernel"]["kernel_us"], r["other_kernel"]["frac"])
l = d["roofline_large_n"]; print("large", l["kernel"], l["kernel_us"], l["frac"], l["other_kernel"]["kernel_us"], l["other_kernel"]["frac"], l["env_step_us"])
e = d["e2e"]; print("e2e", e["value"], e["frac_of_pcie"], e["api"])
print("neigh", {k: round(v["kernel_us"], 2) for k, v in d["neighbours"].items()})
PY

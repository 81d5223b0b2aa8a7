#!/usr/bin/env bash
set -uo pipefail
d=gpurun_out/final
mkdir -p "$d"
export PYTHONUNBUFFERED=1
python bench.py --gpus 1 --steps 20 --warmup 5 > "$d/bench_n1_driver_flags.json" 2> "$d/bench_n1_driver_flags.err"; echo "rc=$?"; wc -l "$d/bench_n1_driver_flags.json"
python bench.py > "$d/bench_n1.json" 2> "$d/bench_n1.err"; echo "rc=$?"; wc -l "$d/bench_n1.json"
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > "$d/gpu_tests.log" 2>&1; echo "tests rc=$?"; tail -n 2 "$d/gpu_tests.log"
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python - <<'PY'
import json
for f in ("gpurun_out/final/bench_n1_driver_flags.json", "gpurun_out/final/bench_n1.json"):
    d = json.load(open(f)); r = d["roofline"]; l = d["roofline_large_n"]
    print(f.split("/")[-1], d["value"], d["ms_per_step"], d["steps"], r["kernel_us"], r["frac"], r["other_kernel"]["kernel_us"], r["traffic"], l["kernel_us"], l["frac"], d["e2e"]["value"], d["e2e"]["frac_of_pcie"])
PY

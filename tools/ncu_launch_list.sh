#!/usr/bin/env bash
# ncu launch list of bench.py (B200_PROFILING.md "launch list" pass): per-launch gpu__time_duration of OUR kernels in
# the order bench.py launches them. Cold-cache, serialised times: compare the kernels' SHARES of the step with the
# warm CUDA-event numbers of the bench line, not the absolutes.
#
#   gpurun -- 'bash tools/ncu_launch_list.sh gpurun_out/launches.csv'
#
# The -k filter matters: without it the launch cap is spent on torch's fill / copy kernels while bench.py builds its
# 24 state sets (that is how the refresh of profiles/r1_launches.csv was lost at the end of round 1).
set -euo pipefail
out="${1:-gpurun_out/launches.csv}"
mkdir -p "$(dirname "$out")"
ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:process_action_kernel|mdp_step_kernel|v2_pre_kernel|v2_post_kernel' -c 400 --csv \
    --log-file "$out" \
    python bench.py --steps 24 --warmup 3 --no-cpu-baseline --no-e2e --no-neighbours --skip-handoff --large-n 0 --pdl off > "${out%.csv}.bench.log" 2>&1
python - "$out" <<'PY'
import csv, io, statistics, sys
lines = open(sys.argv[1]).read().split("\n")
h = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[h:]))))
pa = [int(r["Metric Value"]) for r in rows if "process_action" in r["Kernel Name"]]
st = [int(r["Metric Value"]) for r in rows if "mdp_step_kernel" in r["Kernel Name"]]
pre, post = st[0::2], st[1::2]          # the general kernel's step launches alternate: DONES|REWARDS|COMPACT, RESET|COMMAND|OBS
pre += [int(r["Metric Value"]) for r in rows if "v2_pre_kernel" in r["Kernel Name"]]
post += [int(r["Metric Value"]) for r in rows if "v2_post_kernel" in r["Kernel Name"]]
m = lambda x: statistics.mean(x) / 1e3 if x else float("nan")
tot = m(pa) + m(pre) + m(post)
print(f"{len(rows)} launches: process_action {m(pa):.2f} us, pre-reset {m(pre):.2f} us, post-reset {m(post):.2f} us; "
      f"shares {m(pa) / tot:.2f} / {m(pre) / tot:.2f} / {m(post) / tot:.2f}")
PY

#!/usr/bin/env bash
set -uo pipefail
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
t=gpurun_out/r2_gpu_tests20.log
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > "$t" 2>&1
echo "rc=$?" >> "$t"
tail -n 5 "$t"
bash tools/r2_ab.sh robot_lab_b200/_lib/libmdpstep_prev.so gpurun_out/r2_ab_final.log > /dev/null 2>&1
python - <<'PY'
import re
cur=None; rows={}
for l in open('gpurun_out/r2_ab_final.log'):
    m=re.match(r"== rep=(\d) lib=(\S+) N=(\d+)",l)
    if m: cur=(m.group(3),'prev' if 'prev' in m.group(2) else 'now',m.group(1)); rows[cur]={}; continue
    m=re.search(r"\((pre|post)-reset\)\s+([\d.]+) us",l)
    if m: rows[cur][m.group(1)]=float(m.group(2))
    m=re.search(r"env step.*?([\d.]+) us",l)
    if m: rows[cur]['step']=float(m.group(1))
for k,v in sorted(rows.items()): print(k,v)
PY

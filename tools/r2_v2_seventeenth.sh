#!/usr/bin/env bash
set -uo pipefail
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
out=gpurun_out/r2_timeline21.log
: > "$out"
for lib in stamps stamps_few0; do
  echo "##### $lib" >> "$out"
  RL_MDPSTEP_LIB=robot_lab_b200/_lib/libmdpstep_$lib.so timeout 200 python tools/v2_timeline.py 4096 --warm 2>&1 | grep -E "^go2|---|per-warp|per-task|slowest|tasks \+ barrier|prologue|entry ->|first CTA" >> "$out"
done
cut -c1-700 "$out"
bash tools/r2_ab.sh robot_lab_b200/_lib/libmdpstep_few0.so gpurun_out/r2_ab_few.log > /dev/null 2>&1
python - <<'PY'
import re
cur=None; rows={}
for l in open('gpurun_out/r2_ab_few.log'):
    m=re.match(r"== rep=(\d) lib=(\S+) N=(\d+)",l)
    if m: cur=(m.group(3),'few0' if 'few0' in m.group(2) else 'now',m.group(1)); rows[cur]={}; continue
    m=re.search(r"\((pre|post)-reset\)\s+([\d.]+) us",l)
    if m: rows[cur][m.group(1)]=float(m.group(2))
    m=re.search(r"env step.*?([\d.]+) us",l)
    if m: rows[cur]['step']=float(m.group(1))
for k,v in sorted(rows.items()): print(k,v)
PY

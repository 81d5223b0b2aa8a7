#!/usr/bin/env bash
# bench.py on N GPUs of one box, launched the way the driver launches it. usage: bash tools/r2_final_multi.sh N
set -uo pipefail
n="$1"
d=gpurun_out/final
mkdir -p "$d"
export PYTHONUNBUFFERED=1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus "$n" --steps 2400 --warmup 240 > "$d/bench_n$n.json" 2> "$d/bench_n$n.err"
echo "rc=$?"
tail -n 3 "$d/bench_n$n.err"
python - "$d/bench_n$n.json" <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
print({k: d.get(k) for k in ("value", "ms_per_step", "n_gpus", "shard_parity")})
h = d.get("handoff") or {}
print({k: h.get(k) for k in ("compute_only_ms", "unstreamed_ms", "streamed_ms", "streamed_graph_ms", "steps_per_gather", "floor_ms", "value_incl_handoff", "efficiency_vs_compute_only", "streamed_note")})
print("value_incl_handoff", d.get("value_incl_handoff"))
PY

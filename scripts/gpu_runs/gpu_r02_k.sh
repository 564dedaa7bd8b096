#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 --timeout 600 2>&1 | tail -5
B="python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e --no-aten-gpu --no-configs"
show='import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(sys.argv[1], round(d["ms_per_step"],4), {k:round(v["ms"],4) for k,v in d["stages"].items() if not k.startswith("_")}, "graph", (d.get("cuda_graph_replay") or {}).get("ms_per_step"))'
for w in acdc2d_loss la3d acdc2d_trainstep; do
  $B --workload $w 2>/dev/null | python -c "$show" "$w default"
  ARCO_PREFILL_MIN_MB=64 $B --workload $w 2>/dev/null | python -c "$show" "$w prefill>=64MB"
  ARCO_PREFILL_GRAD=0 $B --workload $w 2>/dev/null | python -c "$show" "$w prefill off"
done

#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-aten-gpu --no-configs"
show='import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(sys.argv[1], round(d["ms_per_step"],4), {k:round(v["ms"],4) for k,v in d["stages"].items() if not k.startswith("_")}, "graph", (d.get("cuda_graph_replay") or {}).get("ms_per_step"))'
for w in acdc2d_trainstep cityscapes; do
  for tail in 1 0; do
    ARCO_PROTO_TAIL=$tail $B --workload $w 2>/dev/null | python -c "$show" "$w tail=$tail"
  done
done
for per in 4 32; do
  for w in cityscapes la3d acdc2d_trainstep acdc2d_loss; do
  ARCO_CLASSIFY_CTAS=$per $B --workload $w 2>/dev/null | python -c "$show" "$w classify_ctas=$per"
  done
done

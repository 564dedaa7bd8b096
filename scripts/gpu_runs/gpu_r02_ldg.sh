#!/bin/bash
B="python bench.py --warmup 3 --steps 20 --no-cpu --no-e2e --no-aten-gpu --no-configs"
for w in acdc2d_loss cityscapes la3d acdc2d_trainstep; do
  timeout 200 $B --workload $w 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1])
print('$w', round(d['ms_per_step'],4), 'graph', round((d.get('cuda_graph_replay') or {}).get('ms_per_step',0),4), {k:round(v['ms'],4) for k,v in d['stages'].items() if not k.startswith('_')})"
done
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2

#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_dist.py -x -q --timeout 240 2>&1 | tail -2
for n in 1 2; do
timeout 300 python bench.py --gpus $n --steps 30 --warmup 5 --no-configs --no-cpu --no-e2e --no-aten-gpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1])
print(d['n_gpus'], round(d['ms_per_step'],4), round(d['value'],1), d.get('multi_gpu_check'))"
done

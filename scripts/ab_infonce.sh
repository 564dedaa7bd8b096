#!/bin/bash
# A/B of the InfoNCE stage budget (bytes of staged bank rows per warp): prints the per-stage times per setting.
for w in acdc2d_trainstep cityscapes acdc2d_loss; do
  for sb in 4608 9216 18432; do
    echo "== $w stage=$sb"
    ARCO_INFONCE_STAGE=$sb python bench.py --workload $w --steps 20 --warmup 5 --no-cpu --no-e2e | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms_per_step'],4), {k:round(v['ms'],4) for k,v in d['stages'].items() if k=='infonce'})"
  done
done

#!/bin/bash
# usage: ab_env.sh VAR "v1 v2 ..." [workload]   -- per-stage times of the bench under each value of an env knob
VAR=$1; VALS=$2; W=${3:-acdc2d_trainstep}
for v in $VALS; do
  echo "== $W $VAR=$v"
  env $VAR=$v python bench.py --workload $W --steps 20 --warmup 5 --no-cpu --no-e2e | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms_per_step'],4), {k:round(v['ms'],4) for k,v in d['stages'].items() if k[0]!='_'})"
done

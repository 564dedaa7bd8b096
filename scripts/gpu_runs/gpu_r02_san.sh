#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --warmup 3 --no-cpu --no-e2e --no-aten-gpu --no-configs"
for w in acdc2d_loss la3d; do
  ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 40 --csv --log-file gpurun_out/r02_launches_$w.csv $B --workload $w --steps 3 > /dev/null 2>&1
done
SAN_TIMEOUT=500 bash scripts/sanitize.sh
grep -c "Potential" gpurun_out/sanitizer/racecheck.log

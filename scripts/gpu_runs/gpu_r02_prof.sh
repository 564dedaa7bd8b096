#!/bin/bash
# ncu launch lists + full captures of the small-shape kernels
mkdir -p gpurun_out
B="python bench.py --warmup 3 --no-cpu --no-e2e --no-aten-gpu --no-configs"
for w in acdc2d_loss la3d; do
  ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 40 --csv --log-file gpurun_out/r02_launches_$w.csv $B --workload $w --steps 3 > /dev/null 2>&1
  ncu --set full --clock-control none --import-source on -k regex:"infonce|proto_|classify" -s 8 -c 6 -o gpurun_out/r02_prof_$w $B --workload $w --steps 2 > gpurun_out/r02_prof_$w.log 2>&1
done
ls -la gpurun_out/*.ncu-rep | tail -3

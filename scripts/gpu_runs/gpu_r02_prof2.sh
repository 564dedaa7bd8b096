#!/bin/bash
# round 2 evidence: launch lists (shares) + full ncu captures of the dominant kernels
mkdir -p gpurun_out
B="python bench.py --warmup 3 --no-cpu --no-e2e --no-aten-gpu --no-configs"
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 48 --csv --log-file gpurun_out/r02_launches_trainstep.csv $B --steps 3 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 48 --csv --log-file gpurun_out/r02_launches_cityscapes.csv $B --workload cityscapes --steps 3 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"proto_tc|infonce|classify|grad_scatter|fill_zero" -s 16 -c 8 -o gpurun_out/r02_prof_trainstep $B --steps 2 > gpurun_out/r02_prof_trainstep.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"proto_tc32|infonce_kernel|classify_kernel" -s 6 -c 4 -o gpurun_out/r02_prof_cityscapes $B --workload cityscapes --steps 2 > gpurun_out/r02_prof_cityscapes.log 2>&1
ls -la gpurun_out/r02_prof_*.ncu-rep

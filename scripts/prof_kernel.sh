#!/bin/bash
# usage: prof_kernel.sh <kernel-regex> <out-name> [workload]  -- one ncu --set full capture of a kernel inside the bench
ncu --set full --clock-control none --import-source on -k regex:"$1" -s ${SKIP:-8} -c 1 -f -o gpurun_out/$2 \
  python bench.py --workload ${3:-acdc2d_trainstep} --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_$2.log 2>&1
tail -2 gpurun_out/ncu_$2.log

#!/bin/bash
# ncu --set full of the InfoNCE kernel on the trainstep workload; ARCO_INFONCE_MMA values as arguments (default "1")
for v in ${@:-1}; do
  ARCO_INFONCE_MMA=$v ncu --set full --clock-control none --import-source on -k regex:"infonce" -s 8 -c 1 -f -o gpurun_out/prof_infonce_mma$v \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_infonce_$v.log 2>&1
  tail -2 gpurun_out/ncu_infonce_$v.log
done

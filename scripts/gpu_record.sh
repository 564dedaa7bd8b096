#!/bin/bash
# The lines and ncu captures that get summarised under profiles/ (run on the GPU box).
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_trainstep.json 2> gpurun_out/bench_trainstep.err; tail -c 300 gpurun_out/bench_trainstep.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -c 300 gpurun_out/bench_reference.err
for w in acdc2d_loss la3d cityscapes; do
  python bench.py --workload $w --steps 20 --warmup 5 --no-cpu --no-e2e > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; tail -c 300 gpurun_out/bench_$w.err
done
python bench.py --workload cityscapes --blocky --steps 20 --warmup 5 --no-cpu --no-e2e > gpurun_out/bench_cityscapes_blocky.json 2>/dev/null
python bench.py --blocky --steps 20 --warmup 5 --no-cpu --no-e2e > gpurun_out/bench_trainstep_blocky.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 60 --csv --log-file gpurun_out/launches_trainstep.csv python bench.py --steps 4 --warmup 5 --no-cpu --no-e2e > gpurun_out/ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"_kernel" -s 60 -c 10 -o gpurun_out/prof_trainstep_final python bench.py --steps 2 --warmup 5 --no-cpu --no-e2e > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"proto_|infonce" -s 8 -c 3 -o gpurun_out/prof_city_final python bench.py --workload cityscapes --steps 2 --warmup 3 --no-cpu --no-e2e >> gpurun_out/ncu_full.log 2>&1
python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --aten-gpu > gpurun_out/bench_trainstep_aten.json 2>/dev/null
python bench.py --workload la3d --bank cold --steps 20 --warmup 5 --no-cpu --no-e2e > gpurun_out/bench_la3d_coldbank.json 2>/dev/null
python scripts/bench_prepare.py > gpurun_out/prepare.jsonl 2>/dev/null

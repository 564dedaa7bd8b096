#!/bin/bash
# One GPU-box pass: parity tests, then the bench lines of every workload (outputs under gpurun_out/).
timeout 900 python -m pytest tests -m gpu -q --maxfail=8 --timeout 90 2>&1 | tail -6
python bench.py --steps 20 --warmup 5 ${BENCH_FLAGS:---no-cpu --no-e2e} > gpurun_out/bench_trainstep.json 2> gpurun_out/bench_trainstep.err; tail -c 400 gpurun_out/bench_trainstep.err
for w in acdc2d_loss la3d cityscapes; do
  python bench.py --workload $w --steps 20 --warmup 5 --no-cpu --no-e2e > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; tail -c 300 gpurun_out/bench_$w.err
done

#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 --timeout 600 2>&1 | tail -25 > gpurun_out/r02j_pytest.txt; tail -8 gpurun_out/r02j_pytest.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02j_bench.json 2> gpurun_out/r02j_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r02j_bench.err
python scripts/host_overhead.py acdc2d_loss 2>&1 | head -40 > gpurun_out/r02j_host.txt; head -3 gpurun_out/r02j_host.txt

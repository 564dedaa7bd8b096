#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 --timeout 600 2>&1 | tail -6 > gpurun_out/r02p_pytest.txt; tail -3 gpurun_out/r02p_pytest.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02p_bench.json 2> gpurun_out/r02p_bench.err; echo "bench rc=$?"; tail -c 400 gpurun_out/r02p_bench.err
SAN_TIMEOUT=600 SAN_PRINT=6000 bash scripts/sanitize.sh
head -30 gpurun_out/sanitizer/racecheck_by_line.txt

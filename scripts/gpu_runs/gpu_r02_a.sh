#!/bin/bash
# round 2, first GPU pass: full parity suite (incl. the production-parameter cases), the default bench line, sanitizers
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 --timeout 600 2>&1 | tail -25 > gpurun_out/r02a_pytest.txt; tail -8 gpurun_out/r02a_pytest.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r02a_bench.err; head -c 600 gpurun_out/r02a_bench.json
SAN_TIMEOUT=420 bash scripts/sanitize.sh

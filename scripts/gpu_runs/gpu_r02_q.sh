#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 200 /usr/local/cuda/bin/compute-sanitizer --tool initcheck --print-limit 2 python scripts/probe/initcheck_cublas.py 2>&1 | grep -E "Uninitialized|ERROR SUMMARY|at void|^ok " | head -6
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 --timeout 600 2>&1 | tail -3 > gpurun_out/r02q_pytest.txt; tail -2 gpurun_out/r02q_pytest.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02q_bench.json 2> gpurun_out/r02q_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r02q_bench.err

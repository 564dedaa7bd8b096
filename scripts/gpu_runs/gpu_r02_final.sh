#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-400
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 --timeout 600 2>&1 | tail -6 > gpurun_out/r02_final_pytest.txt; tail -2 gpurun_out/r02_final_pytest.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_final_bench.json 2> gpurun_out/r02_final_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r02_final_bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_final_bench_reference.json 2>/dev/null; echo "ref rc=$?"
timeout 300 python scripts/bench_step_terms.py > gpurun_out/r02_step_terms.jsonl 2>/dev/null
timeout 300 python scripts/bench_producers.py > gpurun_out/r02_producers.jsonl 2>/dev/null

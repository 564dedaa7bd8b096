#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_revisit.py -q --timeout 600 2>&1 | tail -15
python scripts/bench_step_terms.py revisit 2>&1 | tail -4 | tee gpurun_out/r02_step_terms.jsonl

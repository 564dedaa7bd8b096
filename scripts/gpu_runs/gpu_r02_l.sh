#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_revisit.py tests/test_gpu_stepterms.py -q --timeout 600 2>&1 | tail -15
cd scripts && python bench_step_terms.py 2>&1 | tail -5 | tee ../gpurun_out/r02_step_terms.jsonl

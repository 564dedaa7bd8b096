#!/bin/bash
mkdir -p gpurun_out/sanitizer
ARCO_PROTO_TAIL=0 timeout 300 /usr/local/cuda/bin/compute-sanitizer --tool memcheck --print-limit 5 python scripts/sanitize_run.py tc_bf16 2>&1 | grep -v "^=========     Host Frame" | head -40
SAN_TIMEOUT=400 bash scripts/sanitize.sh

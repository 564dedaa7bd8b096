#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_revisit.py -x -q 2>&1 | tail -2
timeout 120 python scripts/probe/rv_probe.py
timeout 300 ncu --set full --import-source on --clock-control none -k regex:revisit_dots -s 4 -c 1 -o gpurun_out/r02_prof_revisit -f python scripts/probe/rv_probe.py > gpurun_out/r02_prof_revisit.log 2>&1
ls -la gpurun_out/r02_prof_revisit.ncu-rep

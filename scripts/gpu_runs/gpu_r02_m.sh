#!/bin/bash
mkdir -p gpurun_out
cd scripts && timeout 900 python fullstep.py 4 2>&1 | tail -12 | tee ../gpurun_out/r02_fullstep.jsonl

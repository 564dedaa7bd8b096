#!/bin/bash
# compute-sanitizer pass over every kernel family (run on the GPU box; logs under gpurun_out/sanitizer/).
# memcheck: out-of-bounds / misaligned global, shared and local accesses; racecheck: shared-memory hazards between the
# warp roles of the hand-rolled mbarrier pipelines; synccheck: illegal barrier use; initcheck: reads of uninitialised
# global memory (the workspace is a fresh torch.empty every step).
mkdir -p gpurun_out/sanitizer
CS=${CS:-/usr/local/cuda/bin/compute-sanitizer}
for tool in memcheck racecheck synccheck initcheck; do
  extra=""
  [ "$tool" = "racecheck" ] && extra="--racecheck-report all"
  timeout ${SAN_TIMEOUT:-900} $CS --tool $tool $extra --print-limit ${SAN_PRINT:-400} --error-exitcode 77 \
      --log-file gpurun_out/sanitizer/$tool.log python scripts/sanitize_run.py "$@" > gpurun_out/sanitizer/$tool.out 2>&1
  echo "$tool rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitizer/$tool.log | tail -1)"
done
# hazards by kernel and source line (racecheck does not model mbarrier arrive/wait ordering; see profiles/r02_sanitizer.md)
grep -E "(Write|Read) Thread" gpurun_out/sanitizer/racecheck.log | sed -E 's/Thread \([0-9,]+\)//; s/\+0x[0-9a-f]+//; s/\(CUtensorMap[^)]*\)//' | sort | uniq -c | sort -rn > gpurun_out/sanitizer/racecheck_by_line.txt

#!/bin/bash
# Round-2 recording on one B200: tests, the driver-format bench line + reference arm, ncu launch lists (kernel shares) and full
# captures of every dominant kernel.  Outputs under gpurun_out/ (scratch); profiles/summarize.py + kernel_shares.py turn them
# into the tracked summaries.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 --timeout 600 2>&1 | tail -6 > gpurun_out/r02_final_pytest.txt; tail -3 gpurun_out/r02_final_pytest.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_final_bench.json 2> gpurun_out/r02_final_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r02_final_bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_final_bench_reference.json 2>/dev/null; echo "ref rc=$?"
B="python bench.py --warmup 3 --no-cpu --no-e2e --no-aten-gpu --no-configs"
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 48 --csv --log-file gpurun_out/r02_launches_trainstep.csv $B --steps 3 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 48 --csv --log-file gpurun_out/r02_launches_cityscapes.csv $B --workload cityscapes --steps 3 > /dev/null 2>&1
for w in acdc2d_loss la3d; do
  ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 40 --csv --log-file gpurun_out/r02_launches_$w.csv $B --workload $w --steps 3 > /dev/null 2>&1
  ncu --set full --clock-control none --import-source on -k regex:"infonce|proto_|classify" -s 8 -c 6 -f -o gpurun_out/r02_prof_$w $B --workload $w --steps 2 > gpurun_out/r02_prof_$w.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:"proto_tc|infonce|classify|grad_scatter|fill_zero" -s 16 -c 8 -f -o gpurun_out/r02_prof_trainstep $B --steps 2 > gpurun_out/r02_prof_trainstep.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"proto_tc32|infonce_kernel|classify_kernel" -s 6 -c 4 -f -o gpurun_out/r02_prof_cityscapes $B --workload cityscapes --steps 2 > gpurun_out/r02_prof_cityscapes.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"revisit_dots" -s 4 -c 2 -f -o gpurun_out/r02_prof_revisit python scripts/probe/rv_probe.py > gpurun_out/r02_prof_revisit.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"keys_transform|proto_transform|anchor_gather" -s 6 -c 3 -f -o gpurun_out/r02_prof_producers python scripts/bench_producers.py > gpurun_out/r02_prof_producers.log 2>&1
timeout 300 python scripts/bench_step_terms.py > gpurun_out/r02_step_terms.jsonl 2>/dev/null
timeout 300 python scripts/bench_producers.py > gpurun_out/r02_producers.jsonl 2>/dev/null
for f in gpurun_out/r02_prof_*.ncu-rep; do ncu -i $f --page raw --csv > ${f%.ncu-rep}.raw.csv 2>/dev/null; done
ls -la gpurun_out/r02_prof_*.ncu-rep
# only the headline capture travels back as a report (source-level inspection); the others as their raw-page CSV
for f in gpurun_out/r02_prof_*.ncu-rep; do case $f in *trainstep*) ;; *) rm -f $f;; esac; done
du -sh gpurun_out

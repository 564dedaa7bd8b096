#!/bin/bash
# Round-2 FINAL recording on one B200 (after the step-skipping prototype pass and the arco_forward replay cache): tests, the
# driver-format bench line + reference arm, ncu launch lists of the four workloads and full captures of the kernels that
# changed (classify, both tensor-core prototype kernels) plus InfoNCE.  Outputs under gpurun_out/ (scratch);
# profiles/summarize.py + kernel_shares.py turn them into the tracked summaries.
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-400
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 --timeout 600 2>&1 | tail -6 > gpurun_out/r02_final_pytest.txt; tail -3 gpurun_out/r02_final_pytest.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_final_bench.json 2> gpurun_out/r02_final_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r02_final_bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_final_bench_reference.json 2>/dev/null; echo "ref rc=$?"
B="python bench.py --warmup 3 --no-cpu --no-e2e --no-aten-gpu --no-configs"
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 48 --csv --log-file gpurun_out/r02_launches_trainstep.csv $B --steps 3 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 48 --csv --log-file gpurun_out/r02_launches_cityscapes.csv $B --workload cityscapes --steps 3 > /dev/null 2>&1
for w in acdc2d_loss la3d; do
  ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 40 --csv --log-file gpurun_out/r02_launches_$w.csv $B --workload $w --steps 3 > /dev/null 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:"proto_tc|infonce|classify" -s 12 -c 6 -f -o gpurun_out/r02_prof_trainstep $B --steps 2 > gpurun_out/r02_prof_trainstep.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"proto_tc32|classify_kernel" -s 4 -c 4 -f -o gpurun_out/r02_prof_cityscapes_coherent $B --workload cityscapes --coherent --steps 2 > gpurun_out/r02_prof_cityscapes_coherent.log 2>&1
for f in gpurun_out/r02_prof_trainstep.ncu-rep gpurun_out/r02_prof_cityscapes_coherent.ncu-rep; do ncu -i $f --page raw --csv > ${f%.ncu-rep}.raw.csv 2>/dev/null; done
rm -f gpurun_out/r02_prof_cityscapes_coherent.ncu-rep
ls -la gpurun_out/r02_prof_*.ncu-rep
du -sh gpurun_out

B="python bench.py --steps 30 --warmup 5 --no-cpu --no-aten-gpu --no-fullstep --no-e2e --no-configs"
for w in la3d acdc2d_loss; do
  for env in "X=1" "ARCO_FWD_GRAPH=0" "ARCO_PREFILL_GRAD=0" "ARCO_FWD_GRAPH=2"; do
    echo "== $w $env" >> gpurun_out/r02_ab_small.txt
    env $env $B --workload $w 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d.get('cuda_graph_replay',{}).get('ms_per_step'))" >> gpurun_out/r02_ab_small.txt
  done
done
for env in "X=1" "ARCO_FWD_GRAPH=2"; do
  echo "== acdc2d_trainstep $env" >> gpurun_out/r02_ab_small.txt
  env $env $B 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d.get('cuda_graph_replay',{}).get('ms_per_step'))" >> gpurun_out/r02_ab_small.txt
done
cat gpurun_out/r02_ab_small.txt

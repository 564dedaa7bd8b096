# A/B of the small (launch-bound) shapes under environment switches; appends "ms eager / ms torch-graph replay" lines.
B="python bench.py --steps 30 --warmup 5 --no-cpu --no-aten-gpu --no-fullstep --no-e2e --no-configs"
OUT=${OUT:-gpurun_out/r02_ab_small.txt}
for rep in 1 2; do
for w in acdc2d_loss la3d; do
  for env in "X=1" "ARCO_PREFILL_MIN_MB=64" "ARCO_PREFILL_GRAD=0"; do
    echo "== $w $env" >> $OUT
    env $env $B --workload $w 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d.get('cuda_graph_replay',{}).get('ms_per_step'))" >> $OUT
  done
done
done
cat $OUT

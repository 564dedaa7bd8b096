# A/B: sampler stream priority (forward.cu) on the default bench line incl. e2e + ATen baseline (the order the driver runs it in)
for env in "ARCO_SAMPLER_PRIORITY=0" "ARCO_SAMPLER_PRIORITY=1" "ARCO_SAMPLER_PRIORITY=0" "ARCO_SAMPLER_PRIORITY=1"; do
  env $env python bench.py --steps 20 --warmup 5 --no-cpu --no-fullstep 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$env', 'head %.4f' % d['ms_per_step'], ' '.join('%s %.4f/%.4f' % (k, b['ms_per_step'], b['ms_per_step_median']) for k,b in d['configs'].items() if 'trainstep' not in k))
" >> gpurun_out/r02_ab_prio.txt
done
cat gpurun_out/r02_ab_prio.txt

import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k: d.get(k) for k in ('value', 'ms_per_step', 'step_frac_hbm', 'gpu_launches', 'n_gpus')})
if d.get('e2e'): print('e2e', d['e2e']['value'])
if d.get('aten_gpu_baseline'): print('aten ms', d['aten_gpu_baseline']['ms_per_step'], 'speedup', d['aten_gpu_baseline'].get('speedup_of_this_op'))
if d.get('cpu_baseline'): print('cpu', d['cpu_baseline']['value'])
if d.get('multi_gpu_check'): print('mgpu', d['multi_gpu_check'])
def st(s): return {k: (round(v['ms'], 4), round(v['frac_hbm'], 3)) for k, v in s.items() if not k.startswith('_')}
if d.get('stages'): print('stages', st(d['stages']))
if d.get('cuda_graph_replay'): print('graph', d['cuda_graph_replay'])
for k, v in (d.get('configs') or {}).items():
    print(k, round(v['ms_per_step'], 4), round(v['value'], 1), 'step_frac', round(v.get('step_frac_hbm', 0) or 0, 3), st(v['stages']) if v.get('stages') else '', v.get('multi_gpu_check', ''), 'GRAPH', (v.get('cuda_graph_replay') or {}).get('ms_per_step', (v.get('cuda_graph_replay') or {}).get('error')))

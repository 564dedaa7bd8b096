for m in 1 2 1 2; do
export ARCO_PROTO_TC=$m
timeout 120 python -m pytest tests/test_gpu_large.py -m gpu -q -k trainstep 2>&1 | tail -1
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('TC mode $m', 'proto ms', j['stages']['proto_enqueue']['ms'], 'step', j['ms_per_step'])"
done

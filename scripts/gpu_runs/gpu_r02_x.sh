#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_dist.py -x -q --timeout 240 2>&1 | tail -5; cat gpurun_out/dist_res.txt; echo
timeout 600 python -m pytest tests -m gpu -q --maxfail=5 --timeout 300 2>&1 | tail -4
timeout 300 python bench.py --steps 20 --warmup 5 --no-configs --no-cpu --no-e2e --no-aten-gpu > gpurun_out/r02x_bench1.json 2> gpurun_out/r02x_bench1.err; echo "bench1 rc=$?"
timeout 300 python bench.py --gpus 2 --steps 20 --warmup 5 --no-configs --no-cpu --no-e2e > gpurun_out/r02x_bench2.json 2> gpurun_out/r02x_bench2.err; echo "bench2 rc=$?"; tail -c 300 gpurun_out/r02x_bench2.err
python - <<'PY'
import json
for f in ("gpurun_out/r02x_bench1.json","gpurun_out/r02x_bench2.json"):
    try:
        d=json.loads(open(f).read().strip().split("\n")[-1])
        print(f, d["n_gpus"], d["ms_per_step"], d["value"], {k:round(v["ms"],4) for k,v in (d.get("stages") or {}).items() if not k.startswith("_")}, d.get("multi_gpu_check"))
    except Exception as e:
        print(f, "ERR", e)
PY

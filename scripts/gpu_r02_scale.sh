#!/bin/bash
mkdir -p gpurun_out
for n in 1 2 4 8; do
  timeout 300 python bench.py --gpus $n --steps 30 --warmup 5 --no-configs --no-cpu --no-e2e --no-aten-gpu > gpurun_out/r02s_bench$n.json 2> gpurun_out/r02s_bench$n.err; echo "bench$n rc=$?"
done
python - <<'PY'
import json
base=None
for n in (1,2,4,8):
    f="gpurun_out/r02s_bench%d.json"%n
    try:
        d=json.loads(open(f).read().strip().split("\n")[-1])
        if base is None: base=d["value"]
        print(n, round(d["ms_per_step"],4), round(d["value"],1), "eff", round(d["value"]/(n*base),4), {k:round(v["ms"],4) for k,v in (d.get("stages") or {}).items() if not k.startswith("_")}, d.get("multi_gpu_check"))
    except Exception as e:
        print(f, "ERR", e)
PY

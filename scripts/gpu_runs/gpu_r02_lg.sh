#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_logits.py -x -q 2>&1 | tail -15
timeout 200 /usr/local/cuda/bin/compute-sanitizer --tool initcheck --print-limit 3 python scripts/probe/initcheck_cublas.py 2>&1 | grep -E "Uninitialized|ERROR SUMMARY|at void|ok " | head -8
python - <<'PY'
import torch, sys
sys.path.insert(0, ".")
import arco_b200
from arco_b200.synth import bench_inputs, bench_bank
dev = torch.device("cuda", 0)
spec, x = bench_inputs("acdc2d_trainstep", dev)
g = torch.Generator(device=dev).manual_seed(3)
n_l, n_u, Cn, sp = spec.n_lab, spec.n_unlab, spec.classes, tuple(spec.spatial)
pl, pu, ps = (torch.randn(n, Cn, *sp, device=dev, generator=g) for n in (n_l, n_u, n_u))
lab = x["labels"]
ll, lu = lab[:n_l].contiguous(), lab[n_l:].contiguous()
rep = x["rep"].requires_grad_(True)
bank, ptr, caps = bench_bank(spec)
kw = dict(delta_n=0.97, func="smc", num_queries=256, num_negatives=512, temp=0.5)
def two():
    rep.grad = None
    p = arco_b200.prepare_contrast_inputs(ps, pl, pu, ll, lu, 20.0)
    _, loss = arco_b200.compute_contra_memobank_loss(rep, p["label_l"], p["label_u"], p["prob_l_teacher"], p["prob_u_teacher"], p["low_mask_all"], p["high_mask_all"], bank, ptr, caps, x["rep_teacher"], **kw)
    loss.backward()
def one():
    rep.grad = None
    _, loss = arco_b200.compute_contra_memobank_loss_from_logits(rep, ll, lu, pl, pu, ps, 20.0, bank, ptr, caps, x["rep_teacher"], **kw)
    loss.backward()
def timed(f, n=20):
    for _ in range(3): f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): f()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / n
print("prepare + loss (two calls): %.4f ms; logits-in: %.4f ms" % (timed(two), timed(one)))
PY

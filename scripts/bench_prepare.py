"""time the device-side mask / threshold preparation against the trainers' own op mix
(torch softmax / entropy on the GPU, boolean index -> .cpu().numpy() -> np.percentile twice, CPU one-hot scatter ->
.cuda().long(); restated in oracle/prepare_oracle.py, run here with CUDA tensors) on the BASELINE shapes.
One JSON line per workload."""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
import arco_b200
from arco_b200.synth import WORKLOADS
from oracle import prepare_oracle

dev = torch.device("cuda", 0)
for name in ("acdc2d_trainstep", "la3d", "cityscapes"):
    w = WORKLOADS[name]
    nl, nu, C, sp = w["n_lab"], w["n_unlab"], w["classes"], tuple(w["spatial"])
    g = torch.Generator(device=dev).manual_seed(1)
    mk = lambda b: torch.randn((b, C) + sp, device=dev, generator=g)
    pred_u, pl_t, pu_t = mk(nu), mk(nl), mk(nu)
    lab_l = torch.randint(0, C, (nl,) + sp, device=dev, generator=g)
    lab_u = torch.randint(0, C, (nu,) + sp, device=dev, generator=g)
    lab_u[torch.rand(lab_u.shape, device=dev, generator=g) < 0.05] = -1
    alpha = 14.0

    def ours():
        return arco_b200.prepare_contrast_inputs(pred_u, pl_t, pu_t, lab_l, lab_u, alpha)

    def ref():
        r = prepare_oracle.prepare(pred_u, pl_t, pu_t, lab_l, lab_u, alpha, C)
        r["label_l"] = r["label_l"].cuda().long()          # what the call site does (train_arco_2d.py:394)
        r["label_u"] = r["label_u"].cuda().long()
        return r

    a, b = ours(), ref()
    torch.cuda.synchronize()
    same_low = int((a["low_mask_all"] != b["low_mask_all"].to(dev)).sum())
    same_high = int((a["high_mask_all"] != b["high_mask_all"].to(dev)).sum())
    for _ in range(3):
        ours()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ours()
    e1.record()
    e1.synchronize()
    ms_ours = e0.elapsed_time(e1) / 20
    ref()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        ref()
    torch.cuda.synchronize()
    ms_ref = (time.perf_counter() - t0) / 3 * 1e3
    P = (nl + nu) * int(torch.tensor(sp).prod())
    byt = (nl + 2 * nu) * C * (P // (nl + nu)) * 4 + (nl + nu) * C * (P // (nl + nu)) * 4 + P * 8 + 2 * P * 4
    print(json.dumps(dict(workload=name, pixels=P, ms_device_prepare=ms_ours, ms_reference_ops_on_gpu=ms_ref,
                          speedup=ms_ref / ms_ours, algorithmic_bytes=byt, gbs=byt / ms_ours / 1e6,
                          mask_pixels_differing_from_reference_ops=[same_low, same_high],
                          note="reference ops = torch GPU softmax/entropy + 2x (boolean index, D2H, np.percentile) + CPU one-hot scatter + H2D")),
          flush=True)

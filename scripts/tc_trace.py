"""Debug: per-role clock stamps of proto_tc_kernel CTA 0 (build with -DARCO_TC_TRACE into libarco_b200_trace.so).

    python -c "from arco_b200.build import build; build(extra=['-DARCO_TC_TRACE'], out='arco_b200/lib/libarco_b200_trace.so')"
    ARCO_B200_LIB=arco_b200/lib/libarco_b200_trace.so python scripts/tc_trace.py
"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import arco_b200
from arco_b200 import _cabi
from arco_b200.synth import bench_bank, bench_inputs

dev = torch.device("cuda", 0)
spec, x = bench_inputs(sys.argv[1] if len(sys.argv) > 1 else "acdc2d_trainstep", dev)
bank, ptr, caps = bench_bank(spec)
for _ in range(3):
    rep = x["rep"].clone().requires_grad_(True)
    arco_b200.compute_contra_memobank_loss(rep, x["label_l"], x["label_u"], x["prob_l"], x["prob_u"], x["low_mask"],
                                           x["high_mask"], bank, ptr, caps, x["rep_teacher"], num_queries=256,
                                           num_negatives=512)
torch.cuda.synchronize()
buf = np.zeros(8 * 256 + 4 * 160, np.int64)
rc = _cabi.lib.arco_debug_tc_trace(buf.ctypes.data_as(C.c_void_p))
assert rc == 0, rc
out = buf[: 8 * 256].reshape(8, 256)
cta = buf[8 * 256:].reshape(4, 160)[:, :148].astype(np.float64)
k0 = cta[0].min()
cta = (cta - k0) / 1e3
for i, nm in enumerate(["entry", "first full", "all MMAs done", "exit"]):
    print(f"CTA {nm:14s} (us after the first CTA entered): min {cta[i].min():8.2f}  median {np.median(cta[i]):8.2f}  max {cta[i].max():8.2f}")
print("CTA 0:", [round(float(cta[i, 0]), 2) for i in range(4)], " slowest CTA:", int(cta[3].argmax()), [round(float(cta[i, int(cta[3].argmax())]), 2) for i in range(4)])
names = ["prod:empty", "mma:full", "mma:bfull", "mma:commit", "bld:empty", "bld:kfull", "cpy:kfull", "cpy:done"]
t0 = out[0, 0]
ns = 1.0 / 1.965          # cycles -> ns at 1965 MHz
print("step " + " ".join(f"{n:>10s}" for n in names) + "   (us since the first TMA issue)")
last = int((out[0] != 0).sum()) - 1
print('steps traced:', last + 1)
for it in list(range(0, 8)) + list(range(100, 108)) + list(range(max(0, last - 20), last + 1)):
    print(f"{it:4d} " + " ".join(f"{(out[r, it] - t0) * ns / 1e3:10.2f}" for r in range(8)))
d = np.diff(out[:, 20:160], axis=1) * ns / 1e3
print("mean step period (us) per role:", {names[r]: round(float(d[r].mean()), 3) for r in range(8)})
lat = (out[1, 20:160] - out[0, 20:160]) * ns / 1e3
print("TMA issue -> full seen by MMA (us): mean %.2f  p10 %.2f  p90 %.2f" % (lat.mean(), np.percentile(lat, 10), np.percentile(lat, 90)))
hold = (out[7, 20:160] - out[1, 20:160]) * ns / 1e3
print("full -> copier release (us): mean %.2f p90 %.2f" % (hold.mean(), np.percentile(hold, 90)))
rel = (out[0, 23:160] - out[7, 20:157]) * ns / 1e3
print("copier release(k-3) -> producer issue(k) (us): mean %.2f p90 %.2f" % (rel.mean(), np.percentile(rel, 90)))
relc = (out[0, 23:160] - out[3, 20:157]) * ns / 1e3
print("mma commit(k-3) -> producer issue(k) (us): mean %.2f p90 %.2f" % (relc.mean(), np.percentile(relc, 90)))

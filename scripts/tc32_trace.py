"""Debug: per-role clock stamps of proto_tc32_kernel CTA 0 (build with -DARCO_TC_TRACE into libarco_b200_trace.so).

    python -c "from arco_b200.build import build; build(extra=['-DARCO_TC_TRACE'], out='arco_b200/lib/libarco_b200_trace.so')"
    ARCO_B200_LIB=arco_b200/lib/libarco_b200_trace.so python scripts/tc32_trace.py [workload]
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
spec, x = bench_inputs(sys.argv[1] if len(sys.argv) > 1 else "cityscapes", dev)
bank, ptr, caps = bench_bank(spec)
for _ in range(3):
    rep = x["rep"].clone().requires_grad_(True)
    arco_b200.compute_contra_memobank_loss(rep, x["label_l"], x["label_u"], x["prob_l"], x["prob_u"], x["low_mask"],
                                           x["high_mask"], bank, ptr, caps, x["rep_teacher"], num_queries=256, num_negatives=512)
torch.cuda.synchronize()
out = np.zeros((8, 512), np.int64)
fn = _cabi.lib.arco_debug_tc32_trace
fn.restype = C.c_int
assert fn(out.ctypes.data_as(C.c_void_p)) == 0
names = ["prod:empty", "mma:hi_beg", "mma:hi_end", "mma:lo_rdy", "bld:empty", "bld:kfull", "cvt:hidone", "cvt:done"]
t0 = out[0, 0]
us = lambda v: (v - t0) / 1.965e3
print("step " + " ".join(f"{n:>10s}" for n in names))
for it in list(range(0, 10)) + list(range(300, 316)):
    print(f"{it:4d} " + " ".join(f"{us(out[r, it]):10.2f}" for r in range(8)))
d = np.diff(out[:, 100:480], axis=1) / 1.965e3
print("mean period (us):", {names[r]: round(float(d[r].mean()), 3) for r in range(8)})
print("hi pass issue + commit (us): %.3f" % float((us(out[2, 100:480]) - us(out[1, 100:480])).mean()))
print("issue->full %.2f | full->hidone(seen by cvt) %.2f | convert %.2f | cvt done->lo seen by mma %.2f" % (
    float((us(out[1, 100:480]) - us(out[0, 100:480])).mean()), float((us(out[6, 100:480]) - us(out[1, 100:480])).mean()),
    float((us(out[7, 100:480]) - us(out[6, 100:480])).mean()), float((us(out[3, 100:480]) - us(out[7, 100:480])).mean())))

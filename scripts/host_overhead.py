"""Where does the host time of one fwd+bwd go?  (small shapes are launch-bound)"""
import cProfile, pstats, sys, time
sys.path.insert(0, ".")
import torch
import arco_b200
from arco_b200.synth import bench_bank, bench_inputs
dev = torch.device("cuda", 0)
spec, x = bench_inputs(sys.argv[1] if len(sys.argv) > 1 else "la3d", dev)
bank, ptr, caps = bench_bank(spec)
rep = x["rep"].requires_grad_(True)
def step():
    rep.grad = None
    _, loss = arco_b200.compute_contra_memobank_loss(rep, x["label_l"], x["label_u"], x["prob_l"], x["prob_u"], x["low_mask"],
        x["high_mask"], bank, ptr, caps, x["rep_teacher"], delta_n=0.97, func="smc", seed=1)
    loss.backward()
for _ in range(20): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(300): step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("300 steps: issue ms/step %.3f   with drain %.3f (issue == drain: the launch queue filled, i.e. the GPU is the limit)" % ((t1 - t0) / 300 * 1e3, (t2 - t0) / 300 * 1e3))
# host cost alone: few enough steps that the launch queue never fills
hs = []
for _ in range(5):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20): step()
    hs.append((time.perf_counter() - t0) / 20 * 1e3)
    torch.cuda.synchronize()
print("20-step bursts: host issue ms/step %s" % ["%.3f" % h for h in hs])
pr = cProfile.Profile(); pr.enable()
for _ in range(300): step()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)

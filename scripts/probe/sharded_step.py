import os, sys, torch, torch.distributed as dist
sys.path.insert(0, ".")
import arco_b200
from arco_b200.synth import bench_inputs, bench_bank
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", local); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
spec, x = bench_inputs(sys.argv[1], dev, seed=1337 + rank)
bank, ptr, caps = bench_bank(spec, seed=1337 + rank)
rep = x["rep"].requires_grad_(True)
kw = dict(delta_n=0.97, func="asmc" if sys.argv[1] == "la3d" else "smc", num_queries=256, num_negatives=512, temp=0.5, process_group=dist.group.WORLD, seed=1337)
def step():
    rep.grad = None
    _, loss = arco_b200.compute_contra_memobank_loss(rep, x["label_l"], x["label_u"], x["prob_l"], x["prob_u"], x["low_mask"], x["high_mask"], bank, ptr, caps, x["rep_teacher"], **kw)
    loss.backward()
for _ in range(5): step()
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(30)]
for a, b in ev:
    a.record(); step(); b.record()
torch.cuda.synchronize()
ms = sorted(a.elapsed_time(b) for a, b in ev)
arco_b200.synchronize_bank(bank)
plan = bank[0].bank.last_plan
print(f"rank {rank} {sys.argv[1]} median {ms[15]:.4f} min {ms[0]:.4f} max {ms[-1]:.4f} replanned={int(plan.replanned)} n_valid={int(plan.n_valid)} plane={arco_b200.contra.EXCHANGE_PLANE}", flush=True)
dist.destroy_process_group()

"""Probe for the initcheck reports in the producers path (profiles/r02_sanitizer.md): the same chain of torch ops that produces
the flagged tensor -- fp32 rows -> bf16, three `a @ w.to(bf16).t()` products with autograd on, then `.float()` -- with no
arco_b200 kernel involved.  Run under `compute-sanitizer --tool initcheck`."""
import sys
import torch
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
rows = torch.randn(n, 80, device=dev, generator=g)
ws = [(torch.randn(80, 80, device=dev, generator=g) / 9).requires_grad_(True) for _ in range(3)]
a = rows.to(torch.bfloat16)
for w in ws:
    a = a @ w.to(torch.bfloat16).t()
f = a.detach().to(torch.float32).contiguous()
torch.cuda.synchronize()
print("ok", float(f.abs().sum()))

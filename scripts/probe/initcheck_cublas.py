"""Probe: does compute-sanitizer --tool initcheck see the stores of cuBLAS's sm_100 bf16 GEMM epilogue?  (It reported the
fp32 up-cast of such a product as an uninitialised read in the producers path; this repeats it with no arco_b200 code.)"""
import torch
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
a = torch.randn(64, 80, device=dev, generator=g).to(torch.bfloat16)
w = torch.randn(80, 80, device=dev, generator=g).to(torch.bfloat16)
c = a @ w.t()
f = c.float()
torch.cuda.synchronize()
print("ok", float(f.abs().sum()))

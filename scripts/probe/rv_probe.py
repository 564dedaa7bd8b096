import sys, torch
sys.path.insert(0, ".")
import arco_b200
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
bs, K, D, H, W = 12, 36, 496, 256, 256
L = D * H * W
pool = torch.nn.functional.normalize(torch.randn(K, L, device=dev, generator=g), dim=1)
for dt in (torch.bfloat16, torch.float32):
    rs = torch.randn(bs, D, H, W, device=dev, generator=g).to(dt)
    rt = torch.randn(bs, D, H, W, device=dev, generator=g).to(dt)
    f = lambda: arco_b200.get_revisiting_loss(pool, rs, rt, topk=5)
    for _ in range(3): f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(10): f()
    e1.record(); e1.synchronize()
    print(dt, e0.elapsed_time(e1) / 10)

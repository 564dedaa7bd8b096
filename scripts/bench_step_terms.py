"""time the fused CUDA versions of the 2-D trainer's other per-step loss terms against
the trainer's own op mix (the oracle restatements, run with CUDA tensors on the same GPU) at the trainer's shapes.
One JSON line per term."""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
sys.path.insert(0, "..")
import arco_b200
import oracle

dev = torch.device("cuda", 0)
PEAK = 6539.2
try:
    PEAK = float(json.load(open(__import__("os").path.join(__import__("os").path.dirname(__import__("os").path.abspath(__file__)), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def timed(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / n


def bench_revisit(dtype):
    g = torch.Generator(device=dev).manual_seed(1)
    bs, K, D, H, W = 12, 36, 496, 256, 256
    L = D * H * W
    pool = torch.nn.functional.normalize(torch.randn(K, L, device=dev, generator=g), dim=1)
    rs = torch.randn(bs, D, H, W, device=dev, generator=g).to(dtype)
    rt = torch.randn(bs, D, H, W, device=dev, generator=g).to(dtype)
    ptr = torch.zeros(1, dtype=torch.long)

    def ours():
        loss = arco_b200.get_revisiting_loss(pool, rs, rt, topk=5)
        arco_b200.revisit_enqueue(rt, pool, ptr)
        return loss

    pool_r = pool.clone()
    ptr_r = torch.zeros(1, dtype=torch.long)

    def ref():
        loss, _, _, _ = oracle.revisiting_loss(pool_r, rs, rt, topk=5)
        oracle.pool_enqueue(rt, pool_r, ptr_r)
        return loss

    ms_loss = timed(lambda: arco_b200.get_revisiting_loss(pool, rs, rt, topk=5))
    ms = timed(ours)
    ms_ref = timed(ref, n=3, warm=1)
    e = rs.element_size()
    byt_loss = K * L * 4 + 2 * bs * L * e
    byt_enq = bs * L * (e + 4)
    print(json.dumps(dict(term="revisiting loss + pool enqueue", rep_dtype=str(dtype), bs=bs, K=K, length=L,
                          ms_loss_pass=ms_loss, loss_pass_alg_bytes=byt_loss, loss_pass_gbs=byt_loss / ms_loss / 1e6,
                          loss_pass_frac_hbm=byt_loss / ms_loss / 1e6 / PEAK, ms_loss_plus_enqueue=ms,
                          total_alg_bytes=byt_loss + byt_enq, total_frac_hbm=(byt_loss + byt_enq) / ms / 1e6 / PEAK,
                          ms_reference_ops_on_gpu=ms_ref, speedup=ms_ref / ms,
                          note="reference ops: 3x F.normalize of [12, 32.5M] + 2 fp32 einsums against the 4.7 GB pool + slice copy")),
          flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["revisit", "unsup", "eqv"]
    if "revisit" in which:
        bench_revisit(torch.bfloat16)
        bench_revisit(torch.float32)
    if "unsup" in which or "eqv" in which:
        from bench_step_terms_rank4 import bench_eqv, bench_unsup      # noqa: E402
        if "unsup" in which:
            bench_unsup()
        if "eqv" in which:
            bench_eqv()

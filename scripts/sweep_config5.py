"""BASELINE.json configs[4]: similarity/InfoNCE kernel sweep, Q x N x D with a 30000-row bank.
Per point: the gather form on an fp32 ring (FFMA, forward + anchor gradient), the gather form on a bf16 ring (mma.sync,
forward + anchor gradient) and the dense tensor-core form (arco_similarity_dense: split-bf16 tcgen05 GEMM + scalar
gather, forward logits only).  One JSON line per point; profiles/summarize_config5.py turns them into the table."""
import ctypes as C, json, sys
sys.path.insert(0, ".")
import torch
from arco_b200 import _cabi
from arco_b200.bank import DeviceMemoryBank
from arco_b200.similarity import dense_similarity

dev = torch.device("cuda", 0)
lib = _cabi.lib
Cn, S, M = 4, 64 * 64, 30000
out = []
for D in (64, 128, 256):
    for Q in (256, 1024, 4096):
        for N in (512, 2048, 8192):
            if Q * N > 2 ** 25:
                continue
            res = {}
            for tag, dt in (("f32", torch.float32), ("bf16", torch.bfloat16)):
                dims = _cabi.Dims(1, 1, Cn, D, S, Q, N, _cabi.F32 if dt == torch.float32 else _cabi.BF16, 1)
                L = _cabi.workspace_layout(dims)
                ws = torch.empty(L.total_bytes, dtype=torch.uint8, device=dev)
                gen = torch.Generator(device=dev).manual_seed(7)
                lab = torch.randint(0, Cn, (2, S), device=dev, generator=gen)
                prob = torch.softmax(torch.randn(2, Cn, S, device=dev, generator=gen), 1)
                ones = torch.ones(2, S, device=dev)
                rep = torch.randn(2, D, S, device=dev, generator=gen).to(dt)
                cg = torch.Generator().manual_seed(11)
                memobank = [[torch.randn(M, D, generator=cg).to(dt).to(torch.float32)] for _ in range(Cn)]
                ptr = [torch.zeros(1, dtype=torch.long) for _ in range(Cn)]
                bank = DeviceMemoryBank(memobank, ptr, [M] * Cn, D, dev, prefer_bf16=dt == torch.bfloat16)
                sp = torch.cuda.current_stream().cuda_stream
                d, b = C.byref(dims), C.byref(bank.c_struct)
                proto = torch.empty(Cn, D + 1, dtype=torch.float64, device=dev)
                ia = torch.empty(Cn, Q, dtype=torch.int32, device=dev); inn = torch.empty(Cn, Q * N, dtype=torch.int32, device=dev)
                loss = torch.empty(1, device=dev); g = torch.empty(Cn, Q, D, device=dev); pix = torch.empty(Cn, Q, dtype=torch.int32, device=dev)
                _cabi.check(lib.arco_classify_count(d, lab[:1].contiguous().data_ptr(), lab[1:].contiguous().data_ptr(), prob[:1].contiguous().data_ptr(),
                                                    prob[1:].contiguous().data_ptr(), ones.data_ptr(), ones.data_ptr(), 0.3, 0.97, 3, 20, ws.data_ptr(), sp), "c")
                _cabi.check(lib.arco_scan_plan(d, b, ws.data_ptr(), sp), "s")
                _cabi.check(lib.arco_proto_enqueue(d, rep.data_ptr(), b, proto.data_ptr(), ws.data_ptr(), sp), "p")
                _cabi.check(lib.arco_sample(d, _cabi.FUNC_SMC, 1, 1, ia.data_ptr(), inn.data_ptr(), ws.data_ptr(), sp), "m")
                def run():
                    _cabi.check(lib.arco_infonce(d, rep.data_ptr(), b, proto.data_ptr(), ia.data_ptr(), inn.data_ptr(), 0.5, loss.data_ptr(),
                                                 g.data_ptr(), pix.data_ptr(), None, ws.data_ptr(), sp), "i")
                def timeit(fn, n=10):
                    for _ in range(3): fn()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(n): fn()
                    e1.record(); e1.synchronize()
                    return e0.elapsed_time(e1) / n
                res[tag] = timeit(run)
                if dt == torch.bfloat16:
                    anchors = torch.randn(Cn, Q, D, device=dev, generator=gen)
                    idx3 = inn.view(Cn, Q, N)
                    res["dense"] = timeit(lambda: dense_similarity(anchors, bank, list(range(Cn)), idx3), 5)
                    # forward + backward of the dense form (second GEMM: scattered logit gradients x transposed ring), the
                    # like-for-like comparison with the gather kernels, which emit the anchor gradient in the same pass
                    a_req = anchors.clone().requires_grad_(True)
                    g_up = torch.randn(Cn, Q, N, device=dev, generator=gen)
                    def fb():
                        a_req.grad = None
                        dense_similarity(a_req, bank, list(range(Cn)), idx3).backward(g_up)
                    res["dense_fb"] = timeit(fb, 5)
            ms = res["f32"]
            plan = _cabi.Plan.from_buffer_copy(ws[L.plan: L.plan + C.sizeof(_cabi.Plan)].cpu().numpy().tobytes())
            cv = sum(1 for j in range(Cn) if plan.slot_active[j])
            gathered = cv * Q * N * D * 4
            flops = 2 * 2 * cv * Q * (1 + N) * D
            dense_flops = 2 * cv * Q * M * D * 3 * 2      # two GEMMs (scores, gradient), 3-way bf16 split
            gemm_flops = 2 * Cn * Q * M * D * 3                      # what arco_similarity_dense issues (3 bf16 terms, all 4 classes)
            rec = dict(D=D, Q=Q, N=N, M=M, C_v=cv, ms=ms, ms_gather_bf16=res["bf16"], ms_dense_fwd=res["dense"], ms_dense_fwd_bwd=res["dense_fb"],
                       gather_gbs=gathered / ms / 1e6, tflops=flops / ms / 1e9, dense_tflops=gemm_flops / res["dense"] / 1e9,
                       reuse=Q * N / M, dense_tflop_needed=dense_flops / 1e12, dense_ms_at_1pf=dense_flops / 1e15 * 1e3 / 1.0)
            out.append(rec); print(json.dumps(rec), flush=True)
json.dump(out, open("gpurun_out/config5_sweep.json", "w"))

"""Probe: how much of the headline InfoNCE kernel (bf16 ring, D = 496, Q = 256, N = 512, mma.sync kernel) is gather latency?
Times arco_infonce with (a) the sampler's indices, (b) every negative = ring row 0 (always an L1/L2 hit), (c) consecutive rows per
query (streaming-friendly), on banks of 50000 / 30000 rows."""
import ctypes as C, json, sys
sys.path.insert(0, ".")
import torch
from arco_b200 import _cabi
from arco_b200.bank import DeviceMemoryBank

dev = torch.device("cuda", 0)
lib = _cabi.lib
Cn, S, D, Q, N = 4, 256 * 256, 496, 256, 512
caps = [50000, 30000, 30000, 30000]
dims = _cabi.Dims(1, 1, Cn, D, S, Q, N, _cabi.BF16, 1)
L = _cabi.workspace_layout(dims)
ws = torch.empty(L.total_bytes, dtype=torch.uint8, device=dev)
gen = torch.Generator(device=dev).manual_seed(7)
lab = torch.randint(0, Cn, (2, S), device=dev, generator=gen)
prob = torch.softmax(torch.randn(2, Cn, S, device=dev, generator=gen), 1)
ones = torch.ones(2, S, device=dev)
rep = torch.randn(2, D, S, device=dev, generator=gen).to(torch.bfloat16)
cg = torch.Generator().manual_seed(11)
memobank = [[torch.randn(caps[c], D, generator=cg).to(torch.bfloat16).to(torch.float32)] for c in range(Cn)]
ptr = [torch.zeros(1, dtype=torch.long) for _ in range(Cn)]
bank = DeviceMemoryBank(memobank, ptr, caps, D, dev, prefer_bf16=True)
sp = torch.cuda.current_stream().cuda_stream
d, b = C.byref(dims), C.byref(bank.c_struct)
proto = torch.empty(Cn, D + 1, dtype=torch.float64, device=dev)
ia = torch.empty(Cn, Q, dtype=torch.int32, device=dev)
inn = torch.empty(Cn, Q * N, dtype=torch.int32, device=dev)
loss = torch.empty(1, device=dev); g = torch.empty(Cn, Q, D, device=dev); pix = torch.empty(Cn, Q, dtype=torch.int32, device=dev)
_cabi.check(lib.arco_classify_count(d, lab[:1].contiguous().data_ptr(), lab[1:].contiguous().data_ptr(), prob[:1].contiguous().data_ptr(),
                                    prob[1:].contiguous().data_ptr(), ones.data_ptr(), ones.data_ptr(), 0.3, 0.97, 3, 20, ws.data_ptr(), sp), "c")
_cabi.check(lib.arco_scan_plan(d, b, ws.data_ptr(), sp), "s")
_cabi.check(lib.arco_proto_enqueue(d, rep.data_ptr(), b, proto.data_ptr(), ws.data_ptr(), sp), "p")
_cabi.check(lib.arco_sample(d, _cabi.FUNC_SMC, 1, 1, ia.data_ptr(), inn.data_ptr(), ws.data_ptr(), sp), "m")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def run():
    _cabi.check(lib.arco_infonce(d, rep.data_ptr(), b, proto.data_ptr(), ia.data_ptr(), inn.data_ptr(), 0.5, loss.data_ptr(),
                                 g.data_ptr(), pix.data_ptr(), None, ws.data_ptr(), sp), "i")

def timeit(n=10):
    for _ in range(3): run()
    tot = 0.0
    for i in range(n):
        flush.fill_(i)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); e1.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n

res = {"sampled": timeit()}
keep = inn.clone()
inn.zero_()
res["all_row_0"] = timeit()
inn.copy_((torch.arange(Q * N, device=dev, dtype=torch.int32) % 20000).repeat(Cn, 1))
res["consecutive_rows"] = timeit()
print(json.dumps(res))

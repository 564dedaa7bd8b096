"""the loss from the producers' inputs (arco_b200.producers) against the reference composition
-- materialise rep / rep_teacher with the 1x1 convolutions (cuDNN/cuBLAS through torch), then arco_b200's plain op -- at the
2-D trainer's shape (12 + 12 images, 256 x 256, D = 496, bf16 like BASELINE config 2's autocast, Q = 256, N = 512, banks
pre-filled to capacity).  Forward + backward down to the gradient of the student features and of the three student weights.
Also times the new kernels alone.  One JSON line per measurement."""
import ctypes as C
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
sys.path.insert(0, "..")
import arco_b200
from arco_b200 import _cabi, producers
from arco_b200.synth import bench_bank, bench_inputs

dev = torch.device("cuda", 0)
HERE = os.path.dirname(os.path.abspath(__file__))
PEAKS = {}
try:
    PEAKS = json.load(open(os.path.join(HERE, "..", "MEASURED_PEAKS.json")))
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


def run(workload="acdc2d_trainstep", emit=None):
    """Returns the list of result dicts (also handed to ``emit`` one by one as they are produced)."""
    results = []

    def out(d):
        results.append(d)
        if emit is not None:
            emit(d)

    spec, x = bench_inputs(workload, dev)
    D = spec.feat
    g = torch.Generator(device=dev).manual_seed(3)
    cdt = x["rep"].dtype
    ws = [(torch.randn(D, D, device=dev, generator=g) / D ** 0.5).requires_grad_(True) for _ in range(3)]
    wk = torch.randn(D, D, device=dev, generator=g) / D ** 0.5
    xs = x["rep"].clone().requires_grad_(True)          # FeatureExtractor.trunk output of the student
    xt = x["rep_teacher"]
    common = dict(delta_n=0.97, func="smc", num_queries=256, num_negatives=512, temp=0.5)

    bank_a, ptr_a, caps = bench_bank(spec)
    bank_b, ptr_b, _ = bench_bank(spec)

    def reference_composition():
        for t in ws + [xs]:
            t.grad = None
        rep = xs
        for w in ws:
            rep = F.conv2d(rep, w.to(cdt).view(D, D, 1, 1))
        with torch.no_grad():
            rep_t = F.conv2d(xt, wk.to(cdt).view(D, D, 1, 1))
        _, loss = arco_b200.compute_contra_memobank_loss(rep, x["label_l"], x["label_u"], x["prob_l"], x["prob_u"], x["low_mask"],
                                                         x["high_mask"], bank_a, ptr_a, caps, rep_t, **common)
        loss.backward()
        return loss

    def fused(sparse=False):
        for t in ws + [xs]:
            t.grad = None
        _, loss = producers.compute_contra_memobank_loss_from_features(
            xs, xt, ws, wk, x["label_l"], x["label_u"], x["prob_l"], x["prob_u"], x["low_mask"], x["high_mask"], bank_b, ptr_b, caps,
            sparse_grad=sparse, **common)
        loss.backward()
        return loss

    ms_ref = timed(reference_composition, n=5, warm=2)
    ms_fused = timed(fused, n=20, warm=3)
    ms_fused_sparse = timed(lambda: fused(True), n=20, warm=3)
    pixels = spec.pixels
    conv_flop = 2.0 * pixels * D * D
    out(({
        "what": "contrastive loss incl. its producers (teacher fea4; student fea4 + q_representation), fwd+bwd", "workload": workload,
        "dtype": str(cdt), "pixels": pixels, "D": D,
        "ms_reference_composition": ms_ref, "ms_fused_producers": ms_fused, "ms_fused_producers_sparse_grad_leaf_input": ms_fused_sparse,
        "speedup": ms_ref / ms_fused,
        "note": "reference composition = 4 forward + 6 backward 1x1-conv GEMMs of %.0f GFLOP each through torch (cuDNN/cuBLAS) "
                "around arco_b200.compute_contra_memobank_loss; fused = no rep / rep_teacher tensor, weights applied to the "
                "K key rows (tcgen05), the C x D class sums (fp64) and the C*Q anchor rows" % (conv_flop / 1e9)}))

    # ---- the new kernels alone (staged C-ABI calls on a prepared workspace) ----
    bank, ptr, _ = bench_bank(spec)
    dbg = {}
    with torch.no_grad():
        _, loss = producers.compute_contra_memobank_loss_from_features(
            xs, xt, ws, wk, x["label_l"], x["label_u"], x["prob_l"], x["prob_u"], x["low_mask"], x["high_mask"], bank, ptr, caps,
            seed=1, _debug=dbg, **common)
    torch.cuda.synchronize()
    arco_b200.synchronize_bank(bank)
    plan = bank[0].bank.last_plan
    K = sum(min(int(plan.n_key[c]), caps[c]) for c in range(spec.classes))
    dims, wsbuf = dbg["dims"], dbg["ws"]
    b = bank[0].bank
    ring_f32 = b.c_struct.row_dtype == _cabi.F32
    wk_ring = (wk.float() if ring_f32 else wk.to(torch.bfloat16)).contiguous()
    scratch = torch.empty(max(1, _cabi.lib.arco_keys_transform_scratch_bytes(D, b.c_struct.row_dtype)), dtype=torch.uint8, device=dev)
    sp = torch.cuda.current_stream().cuda_stream

    def keys_transform():
        _cabi.check(_cabi.lib.arco_keys_transform(C.byref(dims), C.byref(b.c_struct), wk_ring.data_ptr(), scratch.data_ptr(),
                                                  wsbuf.data_ptr(), sp), "arco_keys_transform")

    ms_kt = timed(keys_transform, n=50, warm=5)
    e = 4 if ring_f32 else 2
    flop = 2.0 * K * D * D * (3 if ring_f32 else 1)
    byts = 2.0 * K * D * e + D * D * e
    tf_peak = float(PEAKS.get("bf16_tflops_sustained", PEAKS.get("bf16_tflops", 1378.6)))
    hbm = float(PEAKS.get("hbm_gbs", 6539.2))
    out(({
        "kernel": "keys_transform_kernel<%s>" % ("tf32 x3" if ring_f32 else "bf16"), "K_rows": K, "D": D, "ms": ms_kt,
        "issued_tflops": flop / ms_kt / 1e9, "alg_bytes": byts, "gbs": byts / ms_kt / 1e6,
        "frac_of_hbm_peak": byts / ms_kt / 1e6 / hbm, "frac_of_bf16_tensor_peak": flop / ms_kt / 1e9 / tf_peak,
        "note": "%d tiles of 128 rows over 148 persistent CTAs (<= 2 per CTA): start-up + one tile's latency, neither roofline binds"
                % ((K + 127) // 128)}))
    proto_x, proto = dbg["proto_sums_x"], torch.empty_like(dbg["proto_sums_x"])
    ms_pt = timed(lambda: _cabi.check(_cabi.lib.arco_proto_transform(spec.classes, D, wk_ring.data_ptr(), _cabi.F32 if ring_f32 else _cabi.BF16,
                                                                      proto_x.data_ptr(), proto.data_ptr(), sp), "pt"), n=50, warm=5)
    rows = torch.empty((spec.classes * 256, D), dtype=torch.float32, device=dev)
    pix = torch.empty((spec.classes * 256,), dtype=torch.int32, device=dev)
    ms_ag = timed(lambda: _cabi.check(_cabi.lib.arco_anchor_gather(C.byref(dims), xs.detach().data_ptr(), dbg["idx_anchor"].data_ptr(),
                                                                    rows.data_ptr(), pix.data_ptr(), wsbuf.data_ptr(), sp), "ag"), n=50, warm=5)
    out({"kernel": "proto_transform_kernel", "ms": ms_pt})
    out({"kernel": "anchor_gather_kernel", "ms": ms_ag, "rows": spec.classes * 256})
    return results


def main():
    run(sys.argv[1] if len(sys.argv) > 1 else "acdc2d_trainstep", emit=lambda d: print(json.dumps(d), flush=True))


if __name__ == "__main__":
    main()

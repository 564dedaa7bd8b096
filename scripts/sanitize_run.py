"""Workload driven under compute-sanitizer (scripts/sanitize.sh): smoke() plus one small case per kernel variant --
tcgen05 bf16 prototype kernel (TMA + mbarrier ring + tcgen05.commit hand-offs), TF32 prototype kernel (TMEM A operand),
pipelined / register / scalar CUDA-core variants, the cluster (DSMEM) sampler, the mma.sync and FFMA InfoNCE kernels with
cp.async.bulk gathers, ring overflow, mask/threshold preparation and the dense tcgen05 similarity.
Sizes are tiny: the sanitizer serialises and instruments every access."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import __graft_entry__ as entry  # noqa: E402
import arco_b200  # noqa: E402
from arco_b200.synth import CaseSpec, exact_case, make_bank  # noqa: E402

SPECS = [
    CaseSpec("tc_bf16", 1, 2, 4, (24, 24), 496, queries=16, negatives=12, dtype="bf16", bank_init="fill:80", caps=[120] * 4),
    CaseSpec("tc32_c19", 1, 1, 19, (32, 32), 256, queries=8, negatives=16, bank_init="fill:40", caps=[64] * 19),
    CaseSpec("tc32_d512", 1, 1, 5, (16, 32), 512, queries=8, negatives=8, bank_init="fill:40", caps=[64] * 5),
    CaseSpec("pipe8", 1, 1, 4, (32, 32), 64, queries=16, negatives=16, bank_init="fill:100", caps=[150] * 4),
    CaseSpec("pipe4_bf16_c19", 1, 1, 19, (32, 32), 48, queries=8, negatives=5, dtype="bf16", bank_init="fill:40", caps=[64] * 19),
    CaseSpec("small_la", 1, 1, 2, (16, 16, 12), 16, queries=16, negatives=8, bank_init="randn1", func="asmc"),
    CaseSpec("scalar_odd", 1, 2, 5, (9, 7, 5), 24, queries=12, negatives=3, bank_init="fill:50", caps=[64] * 5),
    CaseSpec("overflow_tc", 1, 2, 5, (32, 32), 128, queries=8, negatives=8, dtype="bf16", bank_init="fill:10",
             caps=[16, 12, 12, 12, 12], mask_frac=0.9, steps=2),
    # coherent entropy masks: the tensor-core prototype kernels skip the 32- / 64-pixel steps without a needed pixel
    CaseSpec("tc_bf16_coherent", 1, 2, 4, (64, 64), 128, queries=16, negatives=8, dtype="bf16", bank_init="fill:60", caps=[90] * 4,
             mask_mode="coherent", mask_frac=0.1),
    CaseSpec("tc32_coherent", 0, 2, 19, (36, 44), 132, queries=8, negatives=8, bank_init="fill:40", caps=[64] * 19,
             mask_mode="coherent", mask_frac=0.15),
    CaseSpec("grid_sampler", 2, 2, 4, (48, 48), 32, queries=64, negatives=64, bank_init="fill:400", caps=[500] * 4, func="asmc"),
]


def main():
    dev = torch.device("cuda", 0)
    only = sys.argv[1:] or None
    if only is None or "smoke" in only:
        entry.smoke()
    for spec in SPECS:
        if only is not None and spec.name not in only:
            continue
        bank, ptr, caps = make_bank(spec)
        if spec.dtype == "bf16":
            for m in bank:
                m[0] = m[0].to(torch.bfloat16).to(torch.float32)
        for step in range(spec.steps):
            g = {k: v.to(dev) for k, v in exact_case(spec, step).items()}
            for fused in (False, True):
                rep = g["rep"].clone().requires_grad_(True)
                kw = {} if fused else {"_debug": {}}
                _, loss = arco_b200.compute_contra_memobank_loss(
                    rep, g["label_l"], g["label_u"], g["prob_l"], g["prob_u"], g["low_mask"], g["high_mask"], bank, ptr, caps,
                    g["rep_teacher"], delta_n=spec.delta_n, func=spec.func, num_queries=spec.queries,
                    num_negatives=spec.negatives, temp=spec.temp, seed=77, **kw)
                loss.backward()
                torch.cuda.synchronize()
                assert torch.isfinite(loss)
        print("ok", spec.name, float(loss))
    if only is None or "replay" in only:
        # arco_forward's replay cache: direct launches, capture on the second sighting, graph launches afterwards
        spec = CaseSpec("replay", 1, 2, 4, (32, 32), 64, queries=16, negatives=8, bank_init="fill:60", caps=[80] * 4)
        bank, ptr, caps = make_bank(spec)
        g = {k: v.to(dev) for k, v in exact_case(spec, 0).items()}
        rep = g["rep"].clone().requires_grad_(True)
        for step in range(8):
            rep.grad = None
            _, loss = arco_b200.compute_contra_memobank_loss(
                rep, g["label_l"], g["label_u"], g["prob_l"], g["prob_u"], g["low_mask"], g["high_mask"], bank, ptr, caps,
                g["rep_teacher"], delta_n=spec.delta_n, func=spec.func, num_queries=spec.queries, num_negatives=spec.negatives,
                temp=spec.temp, seed=9)
            loss.backward()
            torch.cuda.synchronize()
        import ctypes
        st = (ctypes.c_int64 * 3)()
        arco_b200._cabi.lib.arco_forward_replay_stats(st)
        # (under compute-sanitizer the allocator never hands back the same workspace address, so every step is a first
        # sighting and nothing is captured here; tests/test_gpu_replay.py asserts the captures and replays)
        print("ok replay", float(loss), "replayed / captured / direct:", list(st))
    if only is None or "producers" in only:
        # in-place tcgen05 key-row transform (bf16 kind::f16 and fp32 kind::tf32 x3), anchors as rows
        from arco_b200 import producers
        for dt in ("bf16", "f32"):
            spec = CaseSpec("producers_" + dt, 1, 2, 4, (32, 32), 80, queries=16, negatives=12, dtype=dt, bank_init="fill:150",
                            caps=[200] * 4, mask_frac=0.9)
            bank, ptr, caps = make_bank(spec)
            if dt == "bf16":
                for m in bank:
                    m[0] = m[0].to(torch.bfloat16).to(torch.float32)
            gen = torch.Generator(device=dev).manual_seed(5)
            ws = [(torch.randn(80, 80, device=dev, generator=gen) / 9).requires_grad_(True) for _ in range(3)]
            wk = torch.randn(80, 80, device=dev, generator=gen) / 9
            for step in range(2):
                g = {k: v.to(dev) for k, v in exact_case(spec, step).items()}
                xs = g["rep"].clone().requires_grad_(True)
                _, loss = producers.compute_contra_memobank_loss_from_features(
                    xs, g["rep_teacher"], ws, wk, g["label_l"], g["label_u"], g["prob_l"], g["prob_u"], g["low_mask"], g["high_mask"],
                    bank, ptr, caps, delta_n=spec.delta_n, func=spec.func, num_queries=spec.queries, num_negatives=spec.negatives,
                    temp=spec.temp, seed=7)
                loss.backward()
                torch.cuda.synchronize()
                assert torch.isfinite(loss)
            print("ok", spec.name, float(loss))
    if only is None or "revisit" in only:
        # TMA-box ring (three 2-D boxes per chunk, mbarrier full / empty), ragged tail of L
        gen = torch.Generator(device=dev).manual_seed(2)
        bs, K, L = 3, 9, 5 * 4 * 12 * 8
        pool = torch.nn.functional.normalize(torch.randn(K, L, device=dev, generator=gen), dim=1)
        for dt in (torch.float32, torch.bfloat16):
            rs = torch.randn(bs, 5, 4, 12 * 8, device=dev, generator=gen).to(dt)
            rt = torch.randn(bs, 5, 4, 12 * 8, device=dev, generator=gen).to(dt)
            loss = arco_b200.get_revisiting_loss(pool, rs, rt, topk=3)
            arco_b200.revisit_enqueue(rt, pool, torch.zeros(1, dtype=torch.long))
            torch.cuda.synchronize()
            assert torch.isfinite(loss)
        print("ok revisit", float(loss))
    if only is None or "logits" in only:
        # logits-in classify (teacher softmax + entropy masks in registers) + thresholds-only select
        gen = torch.Generator(device=dev).manual_seed(4)
        n, c = 2, 4
        pl, pu, ps = (torch.randn(n, c, 32, 32, device=dev, generator=gen) for _ in range(3))
        ll = torch.randint(0, c, (n, 32, 32), device=dev, generator=gen)
        lu = torch.randint(-1, c, (n, 32, 32), device=dev, generator=gen)
        spec = CaseSpec("logits", n, n, c, (32, 32), 32, queries=16, negatives=8, bank_init="fill:40", caps=[64] * c)
        bank, ptr, caps = make_bank(spec)
        rep = torch.randn(2 * n, 32, 32, 32, device=dev, generator=gen).requires_grad_(True)
        _, loss = arco_b200.compute_contra_memobank_loss_from_logits(rep, ll, lu, pl, pu, ps, 20.0, bank, ptr, caps, rep.detach() * 0.5,
                                                                     delta_n=0.97, func="smc", num_queries=16, num_negatives=8, seed=3)
        loss.backward()
        torch.cuda.synchronize()
        print("ok logits", float(loss))
    if only is None or "sharded" in only:
        # the exchange block inside the InfoNCE launch + gated redo launches, with a fake peer buffer on the same GPU
        from arco_b200 import contra
        spec = CaseSpec("shard1", 2, 2, 4, (32, 32), 16, queries=32, negatives=8, bank_init="fill:60", caps=[80, 70, 70, 70],
                        label_mode="absent:1", seed=41)
        g = {k: v.to(dev) for k, v in exact_case(spec, 0).items()}
        n_sum = 4 * 17
        slot = (n_sum + 63) // 64 * 64
        mine = torch.zeros(2 * slot + 64 + 8, dtype=torch.float64, device=dev)      # two slots, 64 flags, the step word
        peer = torch.zeros(2 * slot + 64 + 8, dtype=torch.float64, device=dev)
        ps = torch.rand(4, 17, dtype=torch.float64) * 5 + 1
        for sl in range(2):
            peer[sl * slot: sl * slot + n_sum] = ps.flatten().to(dev)
        mine.view(torch.int64)[2 * slot + 1] = 1 << 60
        state = dict(buf=mine, peer_buf=peer, hdl=None, rank=0, world=2, slot=slot, seq=0, seq_flags=1 << 63,
                     peers=torch.tensor([mine.data_ptr(), peer.data_ptr()], dtype=torch.int64, device=dev))
        orig = contra._p2p_exchange
        contra._p2p_exchange = lambda group, d, n: state
        try:
            bank, ptr, caps = make_bank(spec)
            rep = g["rep"].clone().requires_grad_(True)
            _, loss = arco_b200.compute_contra_memobank_loss(rep, g["label_l"], g["label_u"], g["prob_l"], g["prob_u"], g["low_mask"],
                                                             g["high_mask"], bank, ptr, caps, g["rep_teacher"], delta_n=0.97, func="smc",
                                                             num_queries=32, num_negatives=8, seed=5, process_group=object())
            loss.backward()
            torch.cuda.synchronize()
            assert int(mine.view(torch.int64)[2 * slot + 64]) == 1          # the exchange stored its sequence number in the step word
        finally:
            contra._p2p_exchange = orig
        print("ok sharded", float(loss))
    if only is None or "dense" in only:
        from arco_b200.bank import DeviceMemoryBank
        from arco_b200.similarity import dense_similarity
        gen = torch.Generator().manual_seed(5)
        rows = [torch.randn(200, 72, generator=gen).to(torch.bfloat16).to(torch.float32) for _ in range(2)]
        bank = DeviceMemoryBank([[r.clone()] for r in rows], [torch.zeros(1, dtype=torch.long) for _ in range(2)], [257, 200], 72, dev,
                                prefer_bf16=True)
        a = torch.randn(2, 128, 72, generator=gen).to(dev).requires_grad_(True)
        idx = torch.randint(0, 200, (2, 128, 8), generator=gen).to(torch.int32).to(dev)
        out = dense_similarity(a, bank, [1, 0], idx)
        out.sum().backward()
        torch.cuda.synchronize()
        print("ok dense", float(out.sum()))
    if only is None or "prepare" in only:
        n, c, s = 2, 4, 64 * 64
        gen = torch.Generator(device=dev).manual_seed(3)
        pu, plt, put = (torch.randn(n, c, 64, 64, device=dev, generator=gen) for _ in range(3))
        ll = torch.randint(0, c, (n, 64, 64), device=dev, generator=gen)
        lu = torch.randint(-1, c, (n, 64, 64), device=dev, generator=gen)
        out = arco_b200.prepare_contrast_inputs(pu, plt, put, ll, lu, 20.0)
        torch.cuda.synchronize()
        print("ok prepare", sorted(out)[:4])


if __name__ == "__main__":
    main()

"""GPU: shape / dtype sweep of the CUDA path against the oracle's ops (run on the GPU), with the device sampler's own
indices replayed into the oracle.  Exercises every prototype-kernel variant (tcgen05 bf16, pipelined EPL 4/8, small
register kernel, scalar fallback), ragged tiles, empty halves of the batch and odd Q/N."""
import pytest
import torch

import oracle
from arco_b200.synth import CaseSpec, exact_case, make_bank

pytestmark = pytest.mark.gpu

SPECS = [
    # name, n_lab, n_unlab, C, spatial, D, kwargs
    CaseSpec("tc_bf16_ragged", 2, 2, 4, (24, 24), 496, queries=16, negatives=12, dtype="bf16", bank_init="fill:80", caps=[120] * 4),
    CaseSpec("tc_bf16_d64_c16", 1, 2, 16, (40, 40), 64, queries=8, negatives=8, dtype="bf16", bank_init="fill:50", caps=[70] * 16),
    CaseSpec("tc_bf16_3d", 1, 1, 5, (16, 16, 8), 128, queries=16, negatives=9, dtype="bf16", bank_init="fill:60", caps=[90] * 5, func="asmc"),
    CaseSpec("pipe8_f32", 2, 2, 4, (48, 48), 64, queries=32, negatives=16, bank_init="fill:100", caps=[150] * 4),
    CaseSpec("pipe8_f32_d200", 1, 1, 8, (32, 32), 200, queries=16, negatives=7, bank_init="fill:64", caps=[100] * 8),
    CaseSpec("pipe4_f32_c19", 1, 1, 19, (32, 64), 256, queries=8, negatives=16, bank_init="fill:40", caps=[64] * 19),
    CaseSpec("tc32_f32_d512", 1, 2, 5, (32, 48), 512, queries=8, negatives=8, bank_init="fill:40", caps=[64] * 5),
    CaseSpec("tc32_f32_d384_c16", 2, 1, 16, (40, 40), 384, queries=8, negatives=8, bank_init="fill:40", caps=[64] * 16, mask_frac=0.6),
    CaseSpec("tc32_f32_d132_ragged", 1, 1, 3, (24, 28), 132, queries=8, negatives=8, bank_init="fill:40", caps=[64] * 3),
    CaseSpec("pipe4_bf16_c19", 1, 1, 19, (32, 32), 48, queries=8, negatives=5, dtype="bf16", bank_init="fill:40", caps=[64] * 19),
    CaseSpec("small_la", 1, 1, 2, (16, 16, 12), 16, queries=16, negatives=8, bank_init="randn1", func="asmc"),
    CaseSpec("small_c3_d32_bf16", 1, 2, 3, (32, 40), 32, queries=8, negatives=8, dtype="bf16", bank_init="fill:30", caps=[40] * 3),
    CaseSpec("scalar_odd_s", 1, 2, 5, (9, 7, 5), 24, queries=12, negatives=3, bank_init="fill:50", caps=[64] * 5, label_mode="blocky"),
    CaseSpec("scalar_bf16_odd_s", 1, 1, 4, (21, 13), 72, queries=12, negatives=4, dtype="bf16", bank_init="fill:50", caps=[64] * 4),
    CaseSpec("only_unlabelled", 0, 3, 5, (32, 32), 64, queries=16, negatives=8, bank_init="fill:60", caps=[80] * 5),
    CaseSpec("only_labelled", 3, 0, 4, (32, 32), 64, queries=16, negatives=8, bank_init="fill:60", caps=[80] * 4),
    CaseSpec("one_negative", 2, 2, 4, (32, 32), 32, queries=7, negatives=1, bank_init="fill:60", caps=[80] * 4),
    # spatially coherent entropy masks: whole 32- / 64-pixel steps of the unlabelled images hold nothing the prototype pass needs
    # and are skipped by the tensor-core kernels (tile_flagged is a per-32-pixel bit mask)
    CaseSpec("tc_bf16_coherent", 1, 2, 4, (64, 64), 128, queries=16, negatives=8, dtype="bf16", bank_init="fill:60", caps=[90] * 4,
             mask_mode="coherent", mask_frac=0.1),
    CaseSpec("tc_bf16_coherent_ragged", 1, 2, 5, (40, 52), 496, queries=16, negatives=8, dtype="bf16", bank_init="fill:60", caps=[90] * 5,
             mask_mode="coherent", mask_frac=0.15),
    CaseSpec("tc_bf16_coherent_3d", 0, 2, 5, (32, 32, 16), 64, queries=16, negatives=8, dtype="bf16", bank_init="fill:60", caps=[90] * 5,
             mask_mode="coherent"),
    CaseSpec("tc32_f32_coherent", 1, 2, 5, (64, 64), 256, queries=8, negatives=8, bank_init="fill:40", caps=[64] * 5,
             mask_mode="coherent", mask_frac=0.1),
    CaseSpec("tc32_f32_coherent_ragged", 0, 2, 19, (36, 44), 132, queries=8, negatives=8, bank_init="fill:40", caps=[64] * 19,
             mask_mode="coherent", mask_frac=0.15),
    CaseSpec("overflow_tc", 1, 2, 5, (32, 32), 128, queries=8, negatives=8, dtype="bf16", bank_init="fill:10", caps=[16, 12, 12, 12, 12],
             mask_frac=0.9, steps=2),
]


def _rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _check_step_masks(dbg, spec):
    """tile_flagged[t] bit g  <=>  the 32-pixel group g of tile t holds a low-valid (0x20) or key (0x80) code: the bits the
    tensor-core prototype kernels walk instead of every step (classify.cu, all three classify kernels)."""
    L, ws = dbg["layout"], dbg["ws"]
    S = 1
    for s_ in spec.spatial:
        S *= s_
    B, tpi = spec.batch, int(L.tiles_per_image)
    codes = ws[L.codes: L.codes + B * S].view(B, S)
    flagged = ws[L.tile_flagged: L.tile_flagged + 4 * B * tpi].view(torch.int32).view(B, tpi).cpu()
    need = ((codes & 0xA0) != 0)
    pad = tpi * 1024 - S
    if pad:
        need = torch.nn.functional.pad(need, (0, pad))
    groups = need.view(B, tpi, 32, 32).any(dim=-1).cpu()                                   # [B, tile, group]
    weights = (torch.ones(32, dtype=torch.int64) << torch.arange(32))
    want = (groups.long() * weights).sum(dim=-1)
    got = flagged.long() & 0xFFFFFFFF
    assert torch.equal(got, want), "per-tile step masks disagree with the code bytes"


@pytest.mark.parametrize("spec", SPECS, ids=lambda s: s.name)
@pytest.mark.parametrize("index_labels", [False, True], ids=["onehot", "indexlabels"])
def test_shape(spec, index_labels):
    import arco_b200
    dev = torch.device("cuda", 0)
    bank_g, ptr_g, caps = make_bank(spec)
    bank_c, ptr_c, _ = make_bank(spec)
    if spec.dtype == "bf16":
        # a bf16 run only ever enqueues bf16 teacher rows: bf16-exact banks are stored as a bf16 ring and take the
        # tensor-core InfoNCE path (the fp32 ring under a bf16 head is covered by the golden bf16_rep case)
        for bank in (bank_g, bank_c):
            for m in bank:
                m[0] = m[0].to(torch.bfloat16).to(torch.float32)
    Q, N = spec.queries, spec.negatives
    for step in range(spec.steps):
        x = exact_case(spec, step)
        g = {k: v.to(dev) for k, v in x.items()}
        rep_g = g["rep"].clone().requires_grad_(True)
        dbg = {}
        ll, lu = (g["labels"][: spec.n_lab].contiguous(), g["labels"][spec.n_lab:].contiguous()) if index_labels \
            else (g["label_l"], g["label_u"])
        new_keys, loss = arco_b200.compute_contra_memobank_loss(
            rep_g, ll, lu, g["prob_l"], g["prob_u"], g["low_mask"], g["high_mask"], bank_g, ptr_g, caps, g["rep_teacher"],
            delta_n=spec.delta_n, func=spec.func, num_queries=Q, num_negatives=N, temp=spec.temp, seed=77, _debug=dbg)
        loss.backward()
        torch.cuda.synchronize()
        arco_b200.synchronize_bank(bank_g)
        plan = bank_g[0].bank.last_plan
        active = [j for j in range(spec.classes) if plan.slot_active[j]]
        replay = []
        for j in active:
            replay += [dbg["idx_anchor"][j].long().cpu(), dbg["idx_neg"][j, : Q * N].long().cpu()]
        it = iter(replay)
        rep_c = g["rep"].float().clone().requires_grad_(True)
        res = oracle.contra_memobank_loss(
            rep_c, g["label_l"], g["label_u"], g["prob_l"], g["prob_u"], g["low_mask"], g["high_mask"], bank_c, ptr_c, caps,
            g["rep_teacher"].float(), delta_n=spec.delta_n, sampler=lambda h, s: next(it), num_queries=Q, num_negatives=N,
            temp=spec.temp)
        res.loss.backward()
        Cn = spec.classes
        _check_step_masks(dbg, spec)
        assert list(new_keys) == res.new_keys
        assert [int(plan.lv_count[c]) for c in range(Cn)] == res.low_valid_counts
        assert [int(plan.n_anchor[c]) for c in range(Cn)] == [len(a) for a in res.anchor_lists]
        assert [int(plan.valid_class[i]) for i in range(int(plan.n_valid))] == res.valid_classes
        assert [int(q) for q in ptr_g] == [int(q) for q in ptr_c]
        for c in range(Cn):
            assert torch.equal(bank_g[c][0].cpu(), bank_c[c][0].float()), f"bank {c} (step {step})"
        ok = torch.tensor([n > 0 for n in res.low_valid_counts], device=dev)
        proto_g = (dbg["proto_sums"][:, :-1] / dbg["proto_sums"][:, -1:]).float()
        if ok.any():
            assert _rel(proto_g[ok], res.proto[ok]) <= 2e-5
        tol = 2e-2 if spec.dtype == "bf16" else 1e-5
        lo = float(res.loss.detach())
        assert abs(float(loss.detach()) - lo) <= min(tol, 1e-4) * max(1.0, abs(lo))     # the loss itself is fp32 either way
        if spec.dtype == "bf16" and spec.feat % 8 == 0:
            assert bank_g[0].bank.row_dtype == torch.bfloat16
        if float(rep_c.grad.abs().max()) > 0:
            assert _rel(rep_g.grad.float(), rep_c.grad) <= tol
        else:
            assert float(rep_g.grad.float().abs().max()) == 0.0

"""GPU: the device ring-buffer bank keeps the reference's list protocol (memobank[c][0], queue_prtlis, new_keys,
dequeue_and_enqueue -- loss_helper_3d.py:12-32, train_arco_2d.py:147-154)."""
import pytest
import torch

import oracle
from arco_b200.synth import CaseSpec, exact_case, make_bank

pytestmark = pytest.mark.gpu


def _step(spec, bank, ptr, caps, step=0, **kw):
    import arco_b200
    dev = torch.device("cuda", 0)
    x = {k: v.to(dev) for k, v in exact_case(spec, step).items()}
    rep = x["rep"].clone().requires_grad_(True)
    return arco_b200.compute_contra_memobank_loss(
        rep, x["label_l"], x["label_u"], x["prob_l"], x["prob_u"], x["low_mask"], x["high_mask"], bank, ptr, caps,
        x["rep_teacher"], delta_n=spec.delta_n, func=spec.func, num_queries=spec.queries, num_negatives=spec.negatives, **kw)


def test_adoption_keeps_list_protocol():
    import arco_b200
    spec = CaseSpec("bank", 2, 2, 5, (16, 16), 8, bank_init="fill:9", caps=[12, 10, 10, 10, 10], mask_frac=0.9, seed=3)
    bank, ptr, caps = make_bank(spec)
    ref_bank, ref_ptr, _ = make_bank(spec)
    before = [b[0].clone() for b in bank]
    new_keys, loss = _step(spec, bank, ptr, caps)
    assert isinstance(bank[0], arco_b200.BankSlot) and isinstance(bank[0], list) and len(bank[0]) == 1
    # the same step on the CPU lists with the reference semantics
    x = exact_case(spec, 0)
    res = oracle.contra_memobank_loss(x["rep"].clone().requires_grad_(True), x["label_l"], x["label_u"], x["prob_l"],
                                      x["prob_u"], x["low_mask"], x["high_mask"], ref_bank, ref_ptr, caps, x["rep_teacher"],
                                      delta_n=spec.delta_n, num_queries=spec.queries, num_negatives=spec.negatives)
    assert len(new_keys) == spec.classes and new_keys == res.new_keys and new_keys[2] == res.new_keys[2]
    assert [int(k) for k in new_keys] == res.new_keys
    for c in range(spec.classes):
        rows = bank[c][0]
        assert rows.is_cuda and rows.shape[0] == ref_bank[c][0].shape[0] == min(caps[c], before[c].shape[0] + res.new_keys[c])
        assert torch.equal(rows.cpu(), ref_bank[c][0])
        assert torch.equal(bank[c][0][-1].cpu(), ref_bank[c][0][-1])           # row indexing, newest last
    arco_b200.synchronize_bank(bank)
    assert [int(p) for p in ptr] == [int(p) for p in ref_ptr]                    # queue_prtlis follows the reference rule


def test_bank_slot_assignment_and_dequeue_and_enqueue_dropin():
    import arco_b200
    spec = CaseSpec("bank2", 2, 2, 4, (16, 16), 8, bank_init="fill:5", caps=[9, 9, 9, 9], seed=4)
    bank, ptr, caps = make_bank(spec)
    _step(spec, bank, ptr, caps)
    dev = torch.device("cuda", 0)
    fresh = torch.arange(3 * 8, dtype=torch.float32).reshape(3, 8)
    bank[1][0] = fresh                                                           # a trainer resetting one class
    assert torch.equal(bank[1][0].cpu(), fresh)
    cpu_slot, cpu_ptr = [fresh.clone()], torch.zeros(1, dtype=torch.long)
    dev_ptr = torch.zeros(1, dtype=torch.long)
    keys = torch.randn(8, 8)
    n_ref = oracle.fifo_enqueue(keys, cpu_slot, cpu_ptr, 9)
    n_dev = arco_b200.dequeue_and_enqueue(keys.to(dev), bank[1], dev_ptr, 9)
    assert n_ref == n_dev == 8 and int(cpu_ptr) == int(dev_ptr)
    assert torch.equal(bank[1][0].cpu(), cpu_slot[0])
    # the adopted bank keeps working in the loss after manual edits
    nk, loss = _step(spec, bank, ptr, caps, step=1)
    assert torch.isfinite(loss) and len(list(nk)) == 4


def test_mismatched_bank_is_rejected():
    import arco_b200
    spec = CaseSpec("bank3", 2, 2, 4, (16, 16), 8, bank_init="fill:5", caps=[9, 9, 9, 9], seed=5)
    bank, ptr, caps = make_bank(spec)
    _step(spec, bank, ptr, caps)
    with pytest.raises(ValueError, match="queue_size changed"):
        _step(spec, bank, ptr, [9, 9, 9, 10])
    bad, ptr2, caps2 = make_bank(spec)
    bad[2] = [torch.zeros(3, 7)]                                                 # wrong feature size
    with pytest.raises(ValueError, match="must be"):
        _step(spec, bad, ptr2, caps2)


@pytest.mark.parametrize("feat,spatial", [(16, (16, 16)), (72, (32, 32)), (264, (32, 32))])
def test_bf16_ring_is_value_identical_to_fp32_ring(feat, spatial, monkeypatch):
    """A bf16 representation head with a bf16-exact bank is stored as a bf16 ring: the integer artefacts and the bank
    rows (after eviction) are bit-identical to the fp32 ring fed the same sampler seed; loss and gradient agree to
    fp32 summation-order noise (the 16-byte chunks hold 8 dims instead of 4, so the dot products associate
    differently)."""
    import arco_b200
    spec = CaseSpec("bankbf", 2, 2, 5, spatial, feat, bank_init="fill:30", caps=[40, 36, 36, 36, 36], mask_frac=0.9,
                    dtype="bf16", queries=32, negatives=24, seed=31)
    dev = torch.device("cuda", 0)

    def run(narrow):
        monkeypatch.setenv("ARCO_BANK_BF16", "1" if narrow else "0")
        bank, ptr, caps = make_bank(spec)
        for b in bank:
            b[0] = b[0].to(torch.bfloat16).to(torch.float32)
        outs = []
        for step in range(3):                                                    # third step wraps the rings
            x = {k: v.to(dev) for k, v in exact_case(spec, step).items()}
            rep = x["rep"].clone().requires_grad_(True)
            nk, loss = arco_b200.compute_contra_memobank_loss(
                rep, x["label_l"], x["label_u"], x["prob_l"], x["prob_u"], x["low_mask"], x["high_mask"], bank, ptr,
                caps, x["rep_teacher"], delta_n=spec.delta_n, func=spec.func, num_queries=spec.queries,
                num_negatives=spec.negatives, seed=1234 + step)
            loss.backward()
            outs.append((list(nk), loss.detach().clone(), rep.grad.clone()))
        assert bank[0].bank.row_dtype == (torch.bfloat16 if narrow else torch.float32)
        return outs, [bank[c][0].clone() for c in range(spec.classes)], bank

    a, rows_a, bank_a = run(True)
    b, rows_b, _ = run(False)
    for (nk_a, loss_a, g_a), (nk_b, loss_b, g_b) in zip(a, b):
        assert nk_a == nk_b
        assert abs(float(loss_a) - float(loss_b)) <= 1e-5 * max(1.0, abs(float(loss_b)))
        # grad_rep is bf16 and duplicate anchors (sampling with replacement) accumulate through bf16 atomics whose
        # order is not fixed: allow a few bf16 ulps of the largest entry, like two runs of the same configuration
        gmax = max(float(g_b.float().abs().max()), 1e-30)
        assert torch.allclose(g_a.float(), g_b.float(), rtol=2.0 ** -6, atol=2.0 ** -7 * gmax)
    assert sum(a[-1][0]) > 0
    for ra, rb in zip(rows_a, rows_b):
        assert ra.dtype == torch.float32 and torch.equal(ra, rb)
    # assigning a row that bf16 cannot hold widens the ring; the other classes keep their values
    odd = torch.full((2, feat), 1.0 + 2.0 ** -12)
    bank_a[1][0] = odd
    assert bank_a[0].bank.row_dtype == torch.float32
    assert torch.equal(bank_a[1][0].cpu(), odd) and torch.equal(bank_a[2][0], rows_a[2])


def test_bf16_ring_needs_exact_rows_and_bf16_rep():
    import arco_b200
    spec = CaseSpec("bankbf2", 2, 2, 4, (16, 16), 16, bank_init="fill:9", caps=[12, 12, 12, 12], dtype="bf16", seed=33)
    bank, ptr, caps = make_bank(spec)
    bank[3][0] = torch.full((1, 16), 1.0 + 2.0 ** -12)                           # not a bf16 value
    _step(spec, bank, ptr, caps)
    assert bank[0].bank.row_dtype == torch.float32
    spec32 = CaseSpec("bankbf3", 2, 2, 4, (16, 16), 16, bank_init="fill:9", caps=[12, 12, 12, 12], seed=33)
    bank, ptr, caps = make_bank(spec32)
    for b in bank:
        b[0] = b[0].to(torch.bfloat16).to(torch.float32)
    _step(spec32, bank, ptr, caps)
    assert bank[0].bank.row_dtype == torch.float32                               # fp32 keys are not bf16-exact

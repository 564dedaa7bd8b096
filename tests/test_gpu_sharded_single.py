"""GPU (ONE device): the batch-shard machinery of the production path -- the exchange block inside the InfoNCE launch
(arco_infonce_sharded), the re-derived plan, and the gated redo launches -- driven through the public op with a FAKE PEER:
a second exchange buffer on the same GPU whose flag is already raised and whose slots hold preset class sums.  tests/test_gpu_dist.py
covers real peers over NVLink but is skipped on a 1-GPU box; this one is not.

Scenario "same": the peer contributes nothing -> the plan does not change, results must equal the single-GPU op bit for bit.
Scenario "replanned": class 1 is absent from the local batch but present on the "peer" -> the global valid-class list differs
from the local one, the speculative pass is discarded and the gated sampler + InfoNCE launches redo it; checked against the
oracle's multi-GPU restatement (global prototype sums, rank-local everything else)."""
import pytest
import torch

import arco_b200
import oracle
from arco_b200 import contra
from arco_b200.synth import CaseSpec, exact_case, make_bank

pytestmark = pytest.mark.gpu


def _fake_exchange(dev, n, peer_sums):
    slot = (n + 63) // 64 * 64
    mine = torch.zeros(2 * slot + 64, dtype=torch.float64, device=dev)
    peer = torch.zeros(2 * slot + 64, dtype=torch.float64, device=dev)
    for s in range(2):                                        # the slot alternates with the step's sequence number
        peer[s * slot: s * slot + n] = peer_sums.flatten().to(dev)
    mine.view(torch.int64)[2 * slot + 1] = 1 << 60            # "peer 1 has raised every sequence number already"
    return dict(buf=mine, peer_buf=peer, hdl=None, rank=0, world=2, slot=slot, seq=0,
                peers=torch.tensor([mine.data_ptr(), peer.data_ptr()], dtype=torch.int64, device=dev))


@pytest.mark.parametrize("scenario", ["same", "replanned"])
def test_sharded_forward_with_a_fake_peer(scenario, monkeypatch):
    dev = torch.device("cuda", 0)
    spec = CaseSpec("shard1", 2, 2, 4, (32, 32), 16, queries=32, negatives=8, bank_init="fill:60", caps=[80, 70, 70, 70],
                    label_mode="absent:1" if scenario == "replanned" else "iid", seed=41)
    x = exact_case(spec, 0)
    g = {k: v.to(dev) for k, v in x.items()}
    Cn, D = spec.classes, spec.feat
    peer_sums = torch.zeros(Cn, D + 1, dtype=torch.float64)
    if scenario == "replanned":
        gen = torch.Generator().manual_seed(3)
        peer_sums[:, :D] = torch.randn(Cn, D, generator=gen, dtype=torch.float64) * 5
        peer_sums[:, D] = torch.tensor([7.0, 5.0, 9.0, 4.0], dtype=torch.float64)
    state = _fake_exchange(dev, Cn * (D + 1), peer_sums)
    monkeypatch.setattr(contra, "_p2p_exchange", lambda group, d, n: state)
    kw = dict(delta_n=0.97, func="smc", num_queries=spec.queries, num_negatives=spec.negatives, temp=0.5, seed=5)

    bank_s, ptr_s, caps = make_bank(spec)
    rep_s = g["rep"].clone().requires_grad_(True)
    dbg = {"fused": True}
    nk, loss = arco_b200.compute_contra_memobank_loss(rep_s, g["label_l"], g["label_u"], g["prob_l"], g["prob_u"], g["low_mask"],
                                                      g["high_mask"], bank_s, ptr_s, caps, g["rep_teacher"], process_group=object(),
                                                      _debug=dbg, **kw)
    loss.backward()
    torch.cuda.synchronize()
    arco_b200.synchronize_bank(bank_s)
    plan = bank_s[0].bank.last_plan
    assert state["seq"] == 1

    if scenario == "same":
        bank_p, ptr_p, _ = make_bank(spec)
        rep_p = g["rep"].clone().requires_grad_(True)
        nk_p, loss_p = arco_b200.compute_contra_memobank_loss(rep_p, g["label_l"], g["label_u"], g["prob_l"], g["prob_u"], g["low_mask"],
                                                              g["high_mask"], bank_p, ptr_p, caps, g["rep_teacher"], **kw)
        loss_p.backward()
        torch.cuda.synchronize()
        assert int(plan.replanned) == 0
        assert list(nk) == list(nk_p)
        assert float(loss.detach()) == float(loss_p.detach())
        assert float((rep_s.grad - rep_p.grad).abs().max()) <= 1e-6 * float(rep_p.grad.abs().max())    # float atomics for duplicates
        return

    assert int(plan.replanned) == 1, "the fake peer owns a class the local batch lacks: the plan must change"
    nv = int(plan.n_valid)
    assert [int(plan.valid_class[i]) for i in range(nv)] == [0, 1, 2, 3]
    active = [j for j in range(Cn) if plan.slot_active[j]]
    replay = []
    for j in active:
        replay += [dbg["idx_anchor"][j].long().cpu(), dbg["idx_neg"][j, : spec.queries * spec.negatives].long().cpu()]
    it = iter(replay)
    bank_c, ptr_c, _ = make_bank(spec)
    rep_c = x["rep"].clone().requires_grad_(True)
    res = oracle.contra_memobank_loss(rep_c, x["label_l"], x["label_u"], x["prob_l"], x["prob_u"], x["low_mask"], x["high_mask"],
                                      bank_c, ptr_c, caps, x["rep_teacher"], delta_n=0.97, sampler=lambda h, s: next(it),
                                      num_queries=spec.queries, num_negatives=spec.negatives,
                                      proto_sum_hook=lambda local: local + peer_sums)
    res.loss.backward()
    assert res.valid_classes == [0, 1, 2, 3]
    assert list(nk) == res.new_keys
    glob = dbg["proto_sums"].cpu()
    want_cnt = torch.tensor(res.low_valid_counts, dtype=torch.float64) + peer_sums[:, D]
    assert torch.equal(glob[:, D], want_cnt), "global low-valid counts"
    lo = float(res.loss.detach())
    assert abs(float(loss.detach()) - lo) <= 1e-5 * max(1.0, abs(lo)), (float(loss.detach()), lo)
    assert float((rep_s.grad.cpu() - rep_c.grad).norm() / rep_c.grad.norm()) <= 1e-5


@pytest.mark.parametrize("scenario", ["same", "replanned"])
def test_sharded_steps_replay_as_graphs(scenario, monkeypatch):
    """The multi-GPU step through arco_forward's replay cache: the exchange buffer carries a step word, replayed steps take
    their sequence number from it (no launch parameter changes from step to step).  Ten steps with fresh contents under
    fixed addresses, cache on vs cache off: bit-identical loss, gradient, keys, ring rows; the sequence word follows."""
    import ctypes as C
    from arco_b200 import _cabi
    dev = torch.device("cuda", 0)
    spec = CaseSpec("shard_replay", 2, 2, 4, (32, 32), 16, queries=32, negatives=8, bank_init="fill:60", caps=[80, 70, 70, 70],
                    label_mode="absent:1" if scenario == "replanned" else "iid", seed=43)
    Cn, D = spec.classes, spec.feat
    peer_sums = torch.zeros(Cn, D + 1, dtype=torch.float64)
    if scenario == "replanned":
        gen = torch.Generator().manual_seed(3)
        peer_sums[:, :D] = torch.randn(Cn, D, generator=gen, dtype=torch.float64) * 5
        peer_sums[:, D] = torch.tensor([7.0, 5.0, 9.0, 4.0], dtype=torch.float64)
    kw = dict(delta_n=0.97, func="smc", num_queries=spec.queries, num_negatives=spec.negatives, temp=0.5, seed=5)

    def stats():
        buf = (C.c_int64 * 3)()
        _cabi.check(_cabi.lib.arco_forward_replay_stats(buf), "arco_forward_replay_stats")
        return list(buf)

    def run(replay_on, steps=10):
        n = Cn * (D + 1)
        slot = (n + 63) // 64 * 64
        mine = torch.zeros(2 * slot + 64 + 8, dtype=torch.float64, device=dev)        # two slots, 64 flags, the step word
        peer = torch.zeros(2 * slot + 64 + 8, dtype=torch.float64, device=dev)
        for s in range(2):
            peer[s * slot: s * slot + n] = peer_sums.flatten().to(dev)
        mine.view(torch.int64)[2 * slot + 1] = 1 << 60
        state = dict(buf=mine, peer_buf=peer, hdl=None, rank=0, world=2, slot=slot, seq=0, seq_flags=1 << 63,
                     peers=torch.tensor([mine.data_ptr(), peer.data_ptr()], dtype=torch.int64, device=dev))
        monkeypatch.setattr(contra, "_p2p_exchange", lambda group, d, n_: state)
        prev = _cabi.lib.arco_forward_replay(1 if replay_on else 0)
        try:
            bank, ptr, caps = make_bank(spec)
            g = {k: v.to(dev) for k, v in exact_case(spec, 0).items()}
            rep = g["rep"].clone().requires_grad_(True)
            before = stats()
            out = []
            for step in range(steps):
                x = exact_case(spec, step)
                with torch.no_grad():
                    for k, v in x.items():
                        (rep if k == "rep" else g[k]).copy_(v)
                rep.grad = None
                nk, loss = arco_b200.compute_contra_memobank_loss(rep, g["label_l"], g["label_u"], g["prob_l"], g["prob_u"],
                                                                  g["low_mask"], g["high_mask"], bank, ptr, caps, g["rep_teacher"],
                                                                  process_group=object(), **kw)
                loss.backward()
                torch.cuda.synchronize()
                arco_b200.synchronize_bank(bank)
                assert int(mine.view(torch.int64)[2 * slot + 64]) == step + 1          # the device's step word follows
                out.append((loss.detach().clone(), rep.grad.clone(), [int(k) for k in nk],
                            [bank[c][0].cpu().clone() for c in range(Cn)], int(bank[0].bank.last_plan.replanned)))
            return out, [a - b for a, b in zip(stats(), before)]
        finally:
            _cabi.lib.arco_forward_replay(prev)

    on, d_on = run(True)
    off, d_off = run(False)
    assert d_off[0] == 0 and d_off[1] == 0
    assert d_on[0] >= 2, d_on                                    # (graphs captured by an earlier test at the same addresses count too)
    for a, b in zip(on, off):
        assert torch.equal(a[0], b[0]) and a[2] == b[2] and a[4] == b[4]
        assert float((a[1] - b[1]).abs().max()) <= 1e-6 * max(1e-30, float(b[1].abs().max()))   # float atomics for duplicates
        for ra, rb in zip(a[3], b[3]):
            assert torch.equal(ra, rb)
    if scenario == "replanned":
        assert all(o[4] == 1 for o in on)

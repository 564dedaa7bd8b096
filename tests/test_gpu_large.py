"""GPU: BASELINE.json-sized shapes.  The oracle's torch ops are run ON THE GPU here (they are device
agnostic) so that full-size inputs finish in seconds; the CUDA path gets the very same sample indices
(read back from its own Philox sampler and replayed into the oracle).  Plus size-independent properties:
FIFO bank contents, gradient support == sampled anchor pixels, run-to-run bit reproducibility."""
import ctypes as C

import pytest
import torch

import oracle
from arco_b200.synth import bench_bank, bench_inputs

pytestmark = pytest.mark.gpu

SHAPES = [
    # workload, n_lab, n_unlab, blocky
    ("acdc2d_loss", 4, 4, False),          # config 1 at full size
    ("acdc2d_loss", 4, 4, True),
    ("acdc2d_trainstep", 2, 2, False),     # config 2 shape (D=496, bf16), batch reduced to keep the oracle quick
    ("la3d", 2, 2, False),                 # config 3 at full size: C=2 -> no key is ever enqueued (trap 3)
    ("cityscapes", 1, 1, False),           # config 4 shape (C=19, D=256), batch reduced
]


def _rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


@pytest.mark.parametrize("workload,n_lab,n_unlab,blocky", SHAPES)
def test_full_size_against_oracle_ops(workload, n_lab, n_unlab, blocky):
    import arco_b200
    dev = torch.device("cuda", 0)
    spec, x = bench_inputs(workload, dev, seed=7, blocky=blocky, n_lab=n_lab, n_unlab=n_unlab)
    bank_g, ptr_g, caps = bench_bank(spec, seed=3)
    bank_c, ptr_c, _ = bench_bank(spec, seed=3)
    Q, N = 64, 128
    rep_g = x["rep"].clone().requires_grad_(True)
    dbg = {}
    new_keys, loss = arco_b200.compute_contra_memobank_loss(
        rep_g, x["label_l"], x["label_u"], x["prob_l"], x["prob_u"], x["low_mask"], x["high_mask"],
        bank_g, ptr_g, caps, x["rep_teacher"], delta_n=0.97, func="smc", num_queries=Q, num_negatives=N,
        seed=99, _debug=dbg)
    loss.backward()
    torch.cuda.synchronize()
    arco_b200.synchronize_bank(bank_g)
    plan = bank_g[0].bank.last_plan
    active = [j for j in range(spec.classes) if plan.slot_active[j]]
    replay = []
    for j in active:
        replay.append(dbg["idx_anchor"][j].long().cpu())
        replay.append(dbg["idx_neg"][j, : Q * N].long().cpu())
    pos = [0]

    def sampler(high, shape):
        idx = replay[pos[0]]
        pos[0] += 1
        assert idx.numel() == shape and int(idx.max()) < high
        return idx

    rep_c = x["rep"].float().clone().requires_grad_(True)
    ores = oracle.contra_memobank_loss(
        rep_c, x["label_l"], x["label_u"], x["prob_l"], x["prob_u"], x["low_mask"], x["high_mask"],
        bank_c, ptr_c, caps, x["rep_teacher"].float(), delta_n=0.97, sampler=sampler, num_queries=Q,
        num_negatives=N, temp=0.5)
    ores.loss.backward()
    assert pos[0] == len(replay)
    Cn = spec.classes
    # integers: exact
    assert list(new_keys) == ores.new_keys
    assert [int(plan.lv_count[c]) for c in range(Cn)] == ores.low_valid_counts
    assert [int(plan.n_anchor[c]) for c in range(Cn)] == [len(a) for a in ores.anchor_lists]
    assert [int(q) for q in ptr_g] == [int(q) for q in ptr_c]
    for c in range(Cn):
        assert torch.equal(bank_g[c][0].cpu(), bank_c[c][0].float()), f"bank {c}"      # FIFO order, verbatim rows
    # floats
    tol = 2e-2 if spec.dtype == "bf16" else 1e-5
    proto_g = (dbg["proto_sums"][:, :-1] / dbg["proto_sums"][:, -1:]).float()
    assert _rel(proto_g, ores.proto) <= 2e-5
    lo = float(ores.loss.detach())
    assert abs(float(loss.detach()) - lo) <= tol * max(1.0, abs(lo))
    assert _rel(rep_g.grad.float(), rep_c.grad) <= tol
    # gradient support == the sampled anchor pixels
    pix = dbg["anchor_pix"]
    pix = pix[pix >= 0].long().unique()
    S = x["rep"][0, 0].numel()
    touched = (rep_g.grad.float().flatten(2) != 0).any(dim=1).flatten().nonzero().flatten()
    assert set(touched.tolist()) <= set(pix.tolist())


def test_bit_reproducible_run_to_run():
    import arco_b200
    dev = torch.device("cuda", 0)
    outs = []
    for _ in range(2):
        spec, x = bench_inputs("acdc2d_loss", dev, seed=5)
        bank, ptr, caps = bench_bank(spec, seed=5)
        rep = x["rep"].clone().requires_grad_(True)
        dbg = {}
        _, loss = arco_b200.compute_contra_memobank_loss(
            rep, x["label_l"], x["label_u"], x["prob_l"], x["prob_u"], x["low_mask"], x["high_mask"],
            bank, ptr, caps, x["rep_teacher"], delta_n=0.97, func="asmc", seed=1234, _debug=dbg)
        loss.backward()
        outs.append((loss.detach().clone(), rep.grad.clone(), dbg["proto_sums"].clone(), bank[1][0].clone()))
    for a, b in zip(*outs):
        assert torch.equal(a, b)

"""GPU: SURVEY.md section 8(f) rank 2 -- the loss taken from the producers' INPUTS (arco_b200.producers) against

* golden vectors made by executing the reference's own FeatureExtractor class, trainer statements and loss
  (tests/golden/make_golden_producers.py), with the reference's sampled indices replayed, and
* the oracle composition (oracle/producers_oracle.py: materialise rep / rep_teacher with F.conv2d, then the loss oracle) run
  on the GPU in fp32, at a trainer-sized D = 496 bf16 shape where the device sampler's own indices are replayed into it.

Bars: new_keys / pointers exact; fp32 loss, ring rows, gradients of the student feature maps and of the three student weights
within 1e-5 relative (norm-wise) -- the weights are applied AFTER the selections, so ring rows are no longer verbatim copies
and agree to summation order, not bit for bit; bf16: one bf16 ulp (2^-8) on ring rows, 2e-2 on loss / gradients against the
same composition evaluated in bf16."""
import numpy as np
import pytest
import torch

import arco_b200
from arco_b200 import producers
from arco_b200.synth import CaseSpec, exact_case, make_bank
from cases import PRODUCER_CASES, producer_inputs
from oracle import producers_oracle as po
from util import Replay, load_golden, rel_err

pytestmark = pytest.mark.gpu


def _dev():
    return torch.device("cuda", 0)


@pytest.mark.parametrize("case", PRODUCER_CASES, ids=lambda c: c["name"])
def test_fused_producers_match_reference(case):
    dev = _dev()
    spec = case["spec"]
    gold = load_golden(case["name"])
    D = sum(case["fea_dim"])
    bank_g, ptr_g, caps = make_bank(spec)
    q_fe, k_fe = producers.FeatureExtractor(case["fea_dim"], D).to(dev), producers.FeatureExtractor(case["fea_dim"], D).to(dev)
    q_rep = producers.make_q_representation(D).to(dev)
    prev = torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    try:
        for step in range(spec.steps):
            x = {k: v.to(dev) for k, v in exact_case(spec, step).items()}
            pin = producer_inputs(case, step)
            q_fe.load_state_dict({f"fea{i}.weight": pin["w_q_fe"][i][:, :, None, None] for i in range(5)})
            k_fe.load_state_dict({f"fea{i}.weight": pin["w_k_fe"][i][:, :, None, None] for i in range(5)})
            q_rep.load_state_dict({f"{i}.weight": pin["w_q_rep"][i][:, :, None, None] for i in range(2)})
            for m in (q_fe, k_fe, q_rep):
                m.zero_grad()
            maps_l = [t.to(dev).requires_grad_(True) for t in pin["maps_l"]]
            maps_u = [t.to(dev).requires_grad_(True) for t in pin["maps_u"]]
            x_s = torch.cat((q_fe.trunk(maps_l), q_fe.trunk(maps_u)))
            with torch.no_grad():
                x_t = torch.cat((k_fe.trunk([t.to(dev) for t in pin["maps_l_teacher"]]),
                                 k_fe.trunk([t.to(dev) for t in pin["maps_u_teacher"]])))
            anchors, negs = Replay(gold, step).split()
            dbg = {}
            new_keys, loss = producers.compute_contra_memobank_loss_from_features(
                x_s, x_t, [q_fe.fea4.weight, q_rep[0].weight, q_rep[1].weight], k_fe.fea4.weight,
                x["label_l"], x["label_u"], x["prob_l"], x["prob_u"], x["low_mask"], x["high_mask"], bank_g, ptr_g, caps,
                delta_n=spec.delta_n, func=spec.func, num_queries=spec.queries, num_negatives=spec.negatives, temp=spec.temp,
                _inject={"anchor": anchors, "neg": negs}, _debug=dbg)
            loss.backward()
            torch.cuda.synchronize()
            p = f"s{step}_"
            assert list(new_keys) == gold[p + "new_keys"].tolist()
            arco_b200.synchronize_bank(bank_g)
            assert [int(q) for q in ptr_g] == gold[p + "ptrs"].tolist()
            gl = float(gold[p + "loss"])
            assert abs(float(loss.detach()) - gl) <= 1e-5 * max(1.0, abs(gl)), f"loss {float(loss)} vs reference {gl}"
            for c in range(spec.classes):
                got = bank_g[c][0].cpu().numpy()
                assert got.shape == gold[p + f"bank{c}"].shape, f"bank {c} length"
                assert rel_err(torch.from_numpy(got), torch.from_numpy(gold[p + f"bank{c}"])) <= 1e-5, f"bank {c} rows"
                # rows enqueued by EARLIER steps (and the initial fill) are untouched by this step's transform: bit-exact
            assert rel_err(maps_l[4].grad.cpu(), torch.from_numpy(gold[p + "grad_map4_l"])) <= 1e-5
            assert rel_err(maps_u[4].grad.cpu(), torch.from_numpy(gold[p + "grad_map4_u"])) <= 1e-5
            assert rel_err(maps_u[0].grad.cpu(), torch.from_numpy(gold[p + "grad_map0_u"])) <= 1e-5
            for name, w in (("fea4", q_fe.fea4.weight), ("qrep0", q_rep[0].weight), ("qrep1", q_rep[1].weight),
                            ("fea3", q_fe.fea3.weight)):
                g = w.grad.reshape(w.shape[0], w.shape[1]).cpu()
                assert rel_err(g.sum(dim=1), torch.from_numpy(gold[p + f"gw_{name}_rowsum"])) <= 2e-5, name
                assert rel_err(g.sum(dim=0), torch.from_numpy(gold[p + f"gw_{name}_colsum"])) <= 2e-5, name
                gn = float(gold[p + f"gw_{name}_norm"])
                assert abs(float(g.norm()) - gn) <= 1e-5 * gn, name
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = prev


def _trainer_like(dtype, dev, seed=77, n_lab=2, n_unlab=2, spatial=(128, 128), D=496, C=4, Q=64, N=128, fill=600, caps=(700, 650, 650, 650)):
    spec = CaseSpec("producers_big", n_lab, n_unlab, C, spatial, D, queries=Q, negatives=N, func="smc", bank_init=f"fill:{fill}",
                    caps=list(caps), seed=seed, dtype="bf16" if dtype == torch.bfloat16 else "f32")
    x = {k: v.to(dev) for k, v in exact_case(spec, 0).items()}
    g = torch.Generator(device="cpu").manual_seed(seed)
    ws = [(torch.randn(D, D, generator=g) / D ** 0.5).to(dev) for _ in range(4)]
    return spec, x, ws


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
def test_fused_producers_match_oracle_composition(dtype):
    """D = 496, thousands of keys (several 128-row tiles per class, ring wrap), device sampler: the oracle composition on the
    GPU materialises rep / rep_teacher with F.conv2d in the same dtype and replays the indices the device sampler drew."""
    dev = _dev()
    spec, x, ws = _trainer_like(dtype, dev)
    bank_g, ptr_g, caps = make_bank(spec)
    bank_o, ptr_o, _ = make_bank(spec)
    if dtype == torch.bfloat16:                       # bf16-exact initial rows: the device ring stays bf16 (the tcgen05 kind::f16 path)
        for b in (bank_g, bank_o):
            for c in range(spec.classes):
                b[c][0] = b[c][0].to(torch.bfloat16).float()
    prev = torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    try:
        for step in range(2):
            xs = x["rep"].clone().requires_grad_(True)             # the producers' input tensors
            xt = x["rep_teacher"].roll(step, dims=0).contiguous()
            w_s = [w.clone().requires_grad_(True) for w in ws[:3]]
            dbg = {}
            new_keys, loss = producers.compute_contra_memobank_loss_from_features(
                xs, xt, w_s, ws[3], x["label_l"], x["label_u"], x["prob_l"], x["prob_u"], x["low_mask"], x["high_mask"],
                bank_g, ptr_g, caps, delta_n=spec.delta_n, func=spec.func, num_queries=spec.queries, num_negatives=spec.negatives,
                temp=spec.temp, seed=1234, _debug=dbg)
            loss.backward()
            torch.cuda.synchronize()
            arco_b200.synchronize_bank(bank_g)
            plan = bank_g[0].bank.last_plan
            active = [j for j in range(spec.classes) if plan.slot_active[j]]
            calls = []
            for j in active:
                calls += [dbg["idx_anchor"][j].long().cpu(), dbg["idx_neg"][j, : spec.queries * spec.negatives].long().cpu()]
            it = iter(calls)
            xo = x["rep"].clone().requires_grad_(True)
            w_o = [w.clone().requires_grad_(True) for w in ws[:3]]
            res = po.contra_from_features(xo, xt, w_o, ws[3], x["label_l"], x["label_u"], x["prob_l"], x["prob_u"], x["low_mask"],
                                          x["high_mask"], bank_o, ptr_o, caps, delta_n=spec.delta_n,
                                          sampler=lambda high, shape: next(it), num_queries=spec.queries,
                                          num_negatives=spec.negatives, temp=spec.temp)
            res.loss.backward()
            assert list(new_keys) == list(res.new_keys)
            assert max(res.new_keys) > 128, "the case must span several row tiles"
            tol_rows, tol = (2.0 ** -7, 2e-2) if dtype == torch.bfloat16 else (1e-5, 1e-5)
            for c in range(spec.classes):
                got, want = bank_g[c][0].cpu().float(), bank_o[c][0]
                assert got.shape == want.shape
                if dtype == torch.bfloat16:
                    err = ((got - want).abs() / want.abs().clamp_min(2.0 ** -6)).max()
                    assert float(err) <= tol_rows, f"bank {c}: {float(err)}"
                else:
                    assert rel_err(got, want) <= tol_rows, f"bank {c}"
            proto_g = (dbg["proto_sums"][:, :-1] / dbg["proto_sums"][:, -1:]).float().cpu()
            assert rel_err(proto_g, res.proto.float().cpu()) <= (4e-3 if dtype == torch.bfloat16 else 1e-5)
            lo = float(res.loss.detach())
            assert abs(float(loss.detach()) - lo) <= tol * max(1.0, abs(lo)), f"loss {float(loss)} vs oracle {lo}"
            assert rel_err(xs.grad.float().cpu(), xo.grad.float().cpu()) <= tol
            for a, b in zip(w_s, w_o):
                assert rel_err(a.grad.cpu(), b.grad.cpu()) <= tol
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = prev


def test_sparse_gradient_contract_and_no_grad():
    dev = _dev()
    spec, x, ws = _trainer_like(torch.float32, dev, spatial=(32, 32), D=64, Q=32, N=16, fill=80, caps=(120, 100, 100, 100))
    outs = []
    for sparse in (False, True):
        bank, ptr, caps = make_bank(spec)
        xs = x["rep"].clone().requires_grad_(True)
        w_s = [w.clone().requires_grad_(True) for w in ws[:3]]
        _, loss = producers.compute_contra_memobank_loss_from_features(
            xs, x["rep_teacher"], w_s, ws[3], x["label_l"], x["label_u"], x["prob_l"], x["prob_u"], x["low_mask"], x["high_mask"],
            bank, ptr, caps, delta_n=spec.delta_n, func=spec.func, num_queries=spec.queries, num_negatives=spec.negatives,
            temp=spec.temp, seed=5, sparse_grad=sparse)
        loss.backward()
        outs.append((float(loss), xs.grad.clone(), [w.grad.clone() for w in w_s]))
    assert outs[0][0] == outs[1][0]
    assert torch.equal(outs[0][1], outs[1][1])
    for a, b in zip(outs[0][2], outs[1][2]):
        assert torch.equal(a, b)
    bank, ptr, caps = make_bank(spec)
    with torch.no_grad():
        _, loss = producers.compute_contra_memobank_loss_from_features(
            x["rep"], x["rep_teacher"], ws[:3], ws[3], x["label_l"], x["label_u"], x["prob_l"], x["prob_u"], x["low_mask"],
            x["high_mask"], bank, ptr, caps, delta_n=spec.delta_n, func=spec.func, num_queries=spec.queries,
            num_negatives=spec.negatives, temp=spec.temp, seed=5)
    assert float(loss) == outs[0][0]

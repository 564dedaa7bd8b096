"""GPU: the CUDA path (through the C ABI) against the golden vectors of the real reference and
against the CPU oracle, on identical inputs with identical injected sample indices.

Bars (BASELINE.json north_star): masks / counts / index lists / bank contents bit-exact; fp32 loss and
gradients within 1e-5 relative (norm-wise for tensors)."""
import ctypes as C

import numpy as np
import pytest
import torch

import oracle
from arco_b200.synth import exact_case, make_bank
from cases import BY_NAME, CASES
from util import Replay, load_golden, rel_err

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-5


def _dev():
    return torch.device("cuda", 0)


def _to_dev(x, dev):
    return {k: v.to(dev) for k, v in x.items()}


def _export_list(dbg, kind, cls, cap):
    from arco_b200 import _cabi
    out = torch.full((max(cap, 1),), -1, dtype=torch.int32, device=dbg["ws"].device)
    cnt = torch.zeros(1, dtype=torch.int32, device=out.device)
    _cabi.check(_cabi.lib.arco_export_list(C.byref(dbg["dims"]), kind, cls, out.data_ptr(), out.numel(),
                                           cnt.data_ptr(), dbg["ws"].data_ptr(),
                                           torch.cuda.current_stream().cuda_stream), "arco_export_list")
    n = int(cnt.item())
    return out[:n].cpu().long()


def _run_case(spec, use_index_labels=False):
    import arco_b200
    from arco_b200 import _cabi
    dev = _dev()
    gold = load_golden(spec.name)
    bank_gpu, ptr_gpu, caps = make_bank(spec)
    bank_cpu, ptr_cpu, _ = make_bank(spec)
    tol = 2e-2 if spec.dtype == "bf16" else FP32_TOL
    for step in range(spec.steps):
        x = exact_case(spec, step)
        xg = _to_dev(x, dev)
        p = f"s{step}_"
        # ---------------- CPU oracle on the same inputs with the reference's indices ----------------
        rep_c = x["rep"].float().clone().requires_grad_(True)       # bf16 case: exact upcast of the same values
        replay_c = Replay(gold, step)
        ores = oracle.contra_memobank_loss(
            rep_c, x["label_l"], x["label_u"], x["prob_l"], x["prob_u"], x["low_mask"], x["high_mask"],
            bank_cpu, ptr_cpu, caps, x["rep_teacher"].float(), momentum_prototype=x.get("momentum_prototype"),
            i_iter=spec.i_iter, delta_n=spec.delta_n, sampler=replay_c,
            num_queries=spec.queries, num_negatives=spec.negatives, temp=spec.temp)
        ores.loss.backward()
        # ---------------- CUDA path ----------------
        replay_g = Replay(gold, step)
        anchors, negs = replay_g.split()
        rep_g = xg["rep"].clone().requires_grad_(True)
        dbg = {}
        if use_index_labels:
            lab = xg["labels"]
            ll, lu = lab[: spec.n_lab].contiguous(), lab[spec.n_lab:].contiguous()
        else:
            ll, lu = xg["label_l"], xg["label_u"]
        ret = arco_b200.compute_contra_memobank_loss(
            rep_g, ll, lu, xg["prob_l"], xg["prob_u"], xg["low_mask"], xg["high_mask"],
            bank_gpu, ptr_gpu, caps, xg["rep_teacher"], momentum_prototype=xg.get("momentum_prototype"),
            i_iter=spec.i_iter, delta_n=spec.delta_n, func=spec.func,
            num_queries=spec.queries, num_negatives=spec.negatives, temp=spec.temp,
            _inject={"anchor": anchors, "neg": negs}, _debug=dbg)
        new_keys, loss = ret[-2], ret[-1]
        assert len(ret) == (3 if spec.momentum else 2)
        loss.backward()
        torch.cuda.synchronize()

        # ---- integer artefacts: bit-exact against the reference (golden) and the oracle ----
        assert list(new_keys) == gold[p + "new_keys"].tolist() == ores.new_keys
        arco_b200.synchronize_bank(bank_gpu)
        assert [int(q) for q in ptr_gpu] == gold[p + "ptr"].tolist()
        plan = bank_gpu[0].bank.last_plan
        Cn = spec.classes
        assert [int(plan.bank_len[c]) for c in range(Cn)] == gold[p + "bank_len"].tolist()
        assert [int(plan.lv_count[c]) for c in range(Cn)] == ores.low_valid_counts
        assert [int(plan.n_anchor[c]) for c in range(Cn)] == [len(a) for a in ores.anchor_lists]
        assert [int(plan.n_key[c]) for c in range(Cn)] == [len(k) for k in ores.key_lists]
        nv = int(plan.n_valid)
        assert [int(plan.valid_class[i]) for i in range(nv)] == ores.valid_classes
        if nv > 1:
            assert [bool(plan.slot_active[j]) for j in range(nv)] == [s["active"] for s in ores.slots]
        for c in range(Cn):
            assert torch.equal(_export_list(dbg, 0, c, spec.pixels), ores.anchor_lists[c]), f"anchor list {c}"
            assert torch.equal(_export_list(dbg, 1, c, spec.pixels), ores.key_lists[c]), f"key list {c}"
            # bank rows are verbatim copies of teacher rows in FIFO order: exact
            got = bank_gpu[c][0].cpu().numpy()
            assert got.shape == gold[p + f"bank{c}"].shape, f"bank {c} length"
            assert np.array_equal(got, gold[p + f"bank{c}"]), f"bank {c} rows"
        # ---- floats ----
        if nv > 1:
            for j, s in enumerate(ores.slots):
                if s["active"]:
                    assert rel_err(dbg["logits"][j].cpu(), s["logits"]) <= tol, f"logits of position {j}"
            proto_g = (dbg["proto_sums"][:, :-1] / dbg["proto_sums"][:, -1:]).float().cpu()
            ok = torch.tensor([c > 0 for c in ores.low_valid_counts])
            assert rel_err(proto_g[ok], ores.proto[ok]) <= tol
        if spec.momentum:                      # a11: the returned `prototype` tensor [C,Q,1,D]
            assert ret[0].shape == (spec.classes, spec.queries, 1, spec.feat)
            assert rel_err(ret[0].cpu(), torch.from_numpy(gold[p + "prototype"])) <= tol
            assert rel_err(ret[0].cpu(), ores.prototype) <= tol
        gl, ol = float(gold[p + "loss"]), float(ores.loss.detach())
        assert abs(float(loss.detach()) - gl) <= tol * max(1.0, abs(gl)), f"loss {float(loss)} vs reference {gl}"
        assert abs(float(loss.detach()) - ol) <= tol * max(1.0, abs(ol)), f"loss {float(loss)} vs oracle {ol}"
        g_gpu = rep_g.grad.float().cpu()
        g_gold = torch.from_numpy(gold[p + "grad"])
        assert g_gpu.shape == g_gold.shape
        if spec.dtype != "bf16":
            assert torch.equal(g_gpu != 0, g_gold != 0), "gradient support differs from the reference"
        assert rel_err(g_gpu, g_gold) <= tol, f"grad vs reference: {rel_err(g_gpu, g_gold)}"
        assert rel_err(g_gpu, rep_c.grad) <= tol, f"grad vs oracle: {rel_err(g_gpu, rep_c.grad)}"


@pytest.mark.parametrize("spec", CASES, ids=lambda s: s.name)
def test_case_matches_reference(spec):
    _run_case(spec)


@pytest.mark.parametrize("name", ["acdc_smc", "ignore_heavy", "city19", "la3d", "blocky_odd"])
def test_integer_label_maps_match_onehot_path(name):
    """The fused in-kernel one-hot decode (a1) gives the same results as the materialised one-hot."""
    _run_case(BY_NAME[name], use_index_labels=True)


def test_label_onehot_dropin():
    import arco_b200
    lab = torch.randint(-1, 5, (3, 17, 9), device=_dev())
    got = arco_b200.label_onehot(lab, 5)
    assert torch.equal(got.cpu(), oracle.label_onehot(lab.cpu(), 5))


def test_multi_hot_labels_are_reported():
    import arco_b200
    spec = BY_NAME["acdc_smc"]
    x = _to_dev(exact_case(spec, 0), _dev())
    x["label_u"][0, :, 0, 0] = 1                      # every class set on one pixel
    bank, ptr, caps = make_bank(spec)
    rep = x["rep"].clone().requires_grad_(True)
    nk, loss = arco_b200.compute_contra_memobank_loss(
        rep, x["label_l"], x["label_u"], x["prob_l"], x["prob_u"], x["low_mask"], x["high_mask"],
        bank, ptr, caps, x["rep_teacher"], delta_n=0.97, func="smc", num_queries=8, num_negatives=4)
    with pytest.raises(ValueError, match="one-hot"):
        list(nk)


def test_cpu_tensors_are_rejected():
    import arco_b200
    spec = BY_NAME["acdc_smc"]
    x = exact_case(spec, 0)
    bank, ptr, caps = make_bank(spec)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        arco_b200.compute_contra_memobank_loss(
            x["rep"], x["label_l"], x["label_u"], x["prob_l"], x["prob_u"], x["low_mask"], x["high_mask"],
            bank, ptr, caps, x["rep_teacher"])

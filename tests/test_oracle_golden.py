"""CPU: the oracle restatement against golden vectors produced by the real reference."""
import random

import numpy as np
import pytest
import torch

import oracle
from arco_b200.synth import exact_case, make_bank
from cases import CASES, SAMPLER_CASES
from util import Replay, load_golden, rel_err


@pytest.mark.parametrize("spec", CASES, ids=lambda s: s.name)
def test_loss_matches_reference(spec):
    gold = load_golden(spec.name)
    memobank, ptrs, caps = make_bank(spec)
    tol = 2e-2 if spec.dtype == "bf16" else 1e-5
    for step in range(spec.steps):
        x = exact_case(spec, step)
        rep = x["rep"].clone().requires_grad_(True)
        replay = Replay(gold, step)
        res = oracle.contra_memobank_loss(
            rep, x["label_l"], x["label_u"], x["prob_l"], x["prob_u"], x["low_mask"], x["high_mask"],
            memobank, ptrs, caps, x["rep_teacher"], momentum_prototype=x.get("momentum_prototype"), i_iter=spec.i_iter,
            delta_n=spec.delta_n, sampler=replay,
            num_queries=spec.queries, num_negatives=spec.negatives, temp=spec.temp)
        res.loss.backward()
        p = f"s{step}_"
        assert replay.done()
        # integer artefacts: bit-exact
        assert res.new_keys == gold[p + "new_keys"].tolist()
        assert [int(q) for q in ptrs] == gold[p + "ptr"].tolist()
        assert [m[0].shape[0] for m in memobank] == gold[p + "bank_len"].tolist()
        for c, m in enumerate(memobank):          # bank rows are copies of teacher rows: exact, ordered
            assert np.array_equal(m[0].float().numpy(), gold[p + f"bank{c}"]), f"bank {c}"
        # floats
        assert abs(float(res.loss.detach()) - float(gold[p + "loss"])) <= tol * max(1.0, abs(float(gold[p + "loss"])))
        g = torch.from_numpy(gold[p + "grad"])
        assert torch.equal(rep.grad.float() != 0, g != 0) or spec.dtype == "bf16"
        assert rel_err(rep.grad.float(), g) <= tol
        if spec.momentum:
            assert rel_err(res.prototype, torch.from_numpy(gold[p + "prototype"])) <= 1e-6


@pytest.mark.parametrize("func,high,shape,seed", SAMPLER_CASES)
def test_sampler_bitexact(func, high, shape, seed):
    gold = load_golden("samplers")[f"{func}_{high}_{shape}_{seed}"]
    torch.manual_seed(seed)
    random.seed(seed)
    np.random.seed(seed)
    fn = oracle.grid_strata_sample if func == "smc" else oracle.grid_antithetic_sample
    got = fn(high, shape).numpy()
    assert got.shape == (shape,)
    assert np.array_equal(got, gold)


def test_label_onehot_ignore_label():
    lab = torch.tensor([[[0, 2, -1], [1, -1, 3]]])
    oh = oracle.label_onehot(lab, 4)
    assert oh.shape == (1, 4, 2, 3)
    assert oh[0, 0, 0, 2] == 1 and oh[0, 0, 1, 1] == 1      # -1 -> class 0 (trap 4)
    assert oh.sum(1).eq(1).all()

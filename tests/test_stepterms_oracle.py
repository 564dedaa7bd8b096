"""CPU: oracles of the trainer's other per-pixel loss terms against golden vectors made by
running the reference's own source (tests/golden/make_golden_step.py), plus the host-side RNG restatement of RandTPS."""
import itertools
import random

import numpy as np
import pytest
import torch

import oracle
from cases import EQV_CASES, UNSUP_CASES, eqv_inputs, unsup_inputs
from util import load_golden, rel_err


@pytest.mark.parametrize("case", UNSUP_CASES, ids=lambda c: c["name"])
def test_unsupervised_loss_oracle(case):
    gold = load_golden(case["name"])
    x = unsup_inputs(case)
    pred = x["predict"].clone().requires_grad_(True)
    loss = oracle.unsupervised_loss(pred, x["target"], x["logits"], case["strong_threshold"])
    loss.backward()
    assert abs(float(loss.detach()) - float(gold["loss"])) <= 1e-6 * abs(float(gold["loss"]))
    assert rel_err(pred.grad, gold["grad"]) <= 1e-6


def _draw(case):
    """Same seeding and construction order as the golden generator: RandTPS.__init__ draws once, reset draws again."""
    from arco_b200 import stepterms          # only its host-side drawing function is used here (no GPU call)
    torch.manual_seed(case["seed"])
    random.seed(case["seed"])
    np.random.seed(case["seed"])
    ctrl = torch.Tensor(list(itertools.product(torch.arange(-1.0, 1.00001, 2.0 / 4), torch.arange(-1.0, 1.00001, 2.0 / 4))))
    inv_scale = (1.0 / 1.2, 1.0 / 0.8)
    stepterms.draw_source_control_points(ctrl, case["B"], case["sigma"], inv_scale, "affine", True)     # __init__ (:110)
    return stepterms.draw_source_control_points(ctrl, case["B"], case["sigma"], inv_scale, "affine", True)   # reset (:412)


@pytest.mark.parametrize("case", EQV_CASES, ids=lambda c: c["name"])
def test_tps_grid_and_equivariance_oracle(case):
    gold = load_golden(case["name"])
    x = eqv_inputs(case)
    src = _draw(case)
    grid = oracle.tps_grid(src, case["H"], case["W"])
    assert np.allclose(grid.numpy(), gold["grid"], rtol=0, atol=2e-6), float(np.abs(grid.numpy() - gold["grid"]).max())
    g = torch.from_numpy(gold["grid"])
    pred_tps = x["pred_tps"].clone().requires_grad_(True)
    loss, mask_tps, org = oracle.equivariance_loss(pred_tps, x["pred_all"], g, x["labels"], x["logits"], case["weak_threshold"])
    loss.backward()
    assert np.allclose(mask_tps.numpy(), gold["mask_tps"], atol=1e-6) and np.allclose(org.numpy(), gold["pred_tps_org"], atol=1e-5)
    assert abs(float(loss) - float(gold["loss"])) <= 1e-6 * abs(float(gold["loss"]))
    assert rel_err(pred_tps.grad, gold["grad"]) <= 1e-6
    assert np.allclose(oracle.warp(x["images"], g).numpy(), gold["images_tps"], atol=1e-6)

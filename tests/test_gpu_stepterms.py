"""GPU: the fused per-pixel loss terms (compute_unsupervised_loss, RandTPS, the
equivariance loss) against the reference's golden vectors and, at the trainer's size, against the oracle's ops on the GPU."""
import random

import numpy as np
import pytest
import torch

import oracle
from cases import EQV_CASES, UNSUP_CASES, eqv_inputs, unsup_inputs
from util import load_golden, rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", UNSUP_CASES, ids=lambda c: c["name"])
def test_unsupervised_loss_golden(case):
    import arco_b200
    dev = torch.device("cuda", 0)
    gold = load_golden(case["name"])
    x = {k: v.to(dev) for k, v in unsup_inputs(case).items()}
    pred = x["predict"].clone().requires_grad_(True)
    loss = arco_b200.compute_unsupervised_loss(pred, x["target"], x["logits"], case["strong_threshold"])
    (3.0 * loss).backward()
    assert abs(float(loss) - float(gold["loss"])) <= 1e-5 * abs(float(gold["loss"]))
    assert rel_err(pred.grad.cpu() / 3.0, gold["grad"]) <= 1e-5
    assert torch.equal(pred.grad == 0, torch.from_numpy(gold["grad"]).to(dev) == 0)              # identical support


@pytest.mark.parametrize("case", EQV_CASES, ids=lambda c: c["name"])
def test_rand_tps_and_equivariance_golden(case):
    import arco_b200
    dev = torch.device("cuda", 0)
    gold = load_golden(case["name"])
    x = {k: v.to(dev) for k, v in eqv_inputs(case).items()}
    torch.manual_seed(case["seed"])
    random.seed(case["seed"])
    np.random.seed(case["seed"])
    tps = arco_b200.RandTPS(case["W"], case["H"], batch_size=case["B"], sigma=case["sigma"], border_padding=False, random_mirror=True,
                            random_scale=(0.8, 1.2), mode="affine")
    tps.reset_control_points()                                     # same RNG consumption as the reference -> same warp
    grid = tps.grid.data
    assert float((grid.cpu() - torch.from_numpy(gold["grid"])).abs().max()) <= 5e-6
    assert np.allclose(tps(x["images"]).cpu().numpy(), gold["images_tps"], atol=2e-4)
    # the loss on the reference's own grid (isolates the fused kernel from grid rounding)
    tps.grid.data.copy_(torch.from_numpy(gold["grid"]))
    assert np.allclose(tps(x["images"]).cpu().numpy(), gold["images_tps"], atol=2e-6)
    assert np.allclose(tps(x["pred_all"], padding_mode="zeros").cpu().numpy(), gold["pred_tps_org"], atol=2e-5)
    pred_tps = x["pred_tps"].clone().requires_grad_(True)
    loss = arco_b200.tps_equivariance_loss(pred_tps, x["pred_all"], tps, x["labels"], x["logits"], case["weak_threshold"])
    (0.5 * loss).backward()
    assert abs(float(loss) - float(gold["loss"])) <= 1e-5 * abs(float(gold["loss"]))
    assert rel_err(pred_tps.grad.cpu() * 2.0, gold["grad"]) <= 1e-5


def test_trainer_size_against_oracle_ops():
    """batch 24 (12 + 12), 4 classes, 256 x 256: the 2-D trainer's shapes (train_arco_2d.py:41,255-261)."""
    import arco_b200
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(9)
    B, C, H, W = 24, 4, 256, 256
    pred = torch.randn(B, C, H, W, device=dev, generator=g) * 2
    target = torch.randint(0, C, (B, H, W), device=dev, generator=g)
    target[torch.rand(B, H, W, device=dev, generator=g) < 0.2] = -1
    conf = torch.rand(B, H, W, device=dev, generator=g)
    p1 = pred.clone().requires_grad_(True)
    p2 = pred.clone().requires_grad_(True)
    l1 = arco_b200.compute_unsupervised_loss(p1, target, conf, 0.97)
    l2 = oracle.unsupervised_loss(p2, target, conf, 0.97)
    l1.backward()
    l2.backward()
    assert abs(float(l1) - float(l2)) <= 1e-5 * abs(float(l2))
    assert rel_err(p1.grad, p2.grad) <= 1e-5

    torch.manual_seed(3)
    random.seed(3)
    np.random.seed(3)
    tps = arco_b200.RandTPS(W, H, batch_size=B, sigma=0.01, random_scale=(0.8, 1.2), mode="affine")
    tps.reset_control_points()
    labels = torch.randint(0, C, (B, H, W), device=dev, generator=g)
    pred_all = torch.randn(B, C, H, W, device=dev, generator=g) * 2
    q1 = pred.clone().requires_grad_(True)
    q2 = pred.clone().requires_grad_(True)
    e1 = arco_b200.tps_equivariance_loss(q1, pred_all, tps, labels, conf, 0.7)
    e2, mask_tps, _ = oracle.equivariance_loss(q2, pred_all, tps.grid.data, labels, conf, 0.7)
    e1.backward()
    e2.backward()
    assert abs(float(e1) - float(e2)) <= 1e-5 * abs(float(e2)), (float(e1), float(e2))
    assert rel_err(q1.grad, q2.grad) <= 1e-5
    assert float((tps(labels.ne(0).float().unsqueeze(1)) - oracle.warp(labels.ne(0).float().unsqueeze(1), tps.grid.data)).abs().max()) <= 1e-5

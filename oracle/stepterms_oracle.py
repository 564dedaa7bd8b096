"""Restatements of the 2-D trainer's other per-pixel loss terms -- TEST INFRASTRUCTURE ONLY.

``/root/reference/code/train_arco_2d.py``: ``compute_unsupervised_loss`` :482-489, the equivariance block :404-423;
``tps/rand_tps.py:82-153`` + ``tps_stn_pytorch/tps_grid_gen.py:9-75`` (grid from source control points) and
``tps/grid_sample.py:11-12``.  Pinned by ``tests/golden/unsup_*.npz`` / ``eqv_*.npz`` (tests/golden/make_golden_step.py runs
the reference's own source).
"""
import itertools

import torch
import torch.nn.functional as F


def unsupervised_loss(predict, target, logits, strong_threshold):
    batch_size = predict.shape[0]
    valid_mask = (target >= 0).float()                                                           # :484
    weighting = logits.view(batch_size, -1).ge(strong_threshold).sum(-1) / valid_mask.view(batch_size, -1).sum(-1)   # :486
    loss = F.cross_entropy(predict, target, reduction="none", ignore_index=-1)                   # :487
    return torch.mean(torch.masked_select(weighting[:, None, None] * loss, loss > 0))            # :488


def _partial_repr(points, control):
    diff = points.view(-1, 1, 2) - control.view(1, -1, 2)
    d2 = diff[:, :, 0] * diff[:, :, 0] + diff[:, :, 1] * diff[:, :, 1]
    rep = 0.5 * d2 * torch.log(d2)
    rep.masked_fill_(rep != rep, 0)
    return rep


def tps_grid(source_control_points, height, width):
    """TPSGridGen (tps_grid_gen.py:23-75) for the 5 x 5 control lattice of RandTPS (rand_tps.py:103-106)."""
    ctrl = torch.Tensor(list(itertools.product(torch.arange(-1.0, 1.00001, 2.0 / 4), torch.arange(-1.0, 1.00001, 2.0 / 4))))
    n = ctrl.shape[0]
    fk = torch.zeros(n + 3, n + 3)
    fk[:n, :n].copy_(_partial_repr(ctrl, ctrl))
    fk[:n, -3].fill_(1)
    fk[-3, :n].fill_(1)
    fk[:n, -2:].copy_(ctrl)
    fk[-2:, :n].copy_(ctrl.transpose(0, 1))
    inv = torch.inverse(fk)
    coords = torch.Tensor(list(itertools.product(range(height), range(width))))
    Y, X = coords.split(1, dim=1)
    Y = Y * 2 / (height - 1) - 1
    X = X * 2 / (width - 1) - 1
    tc = torch.cat([X, Y], dim=1)
    rep = torch.cat([_partial_repr(tc, ctrl), torch.ones(height * width, 1), tc], dim=1)
    B = source_control_points.shape[0]
    mapping = torch.matmul(inv, torch.cat([source_control_points, torch.zeros(B, 3, 2)], 1))
    return torch.matmul(rep, mapping).view(-1, height, width, 2)


def warp(x, grid, padding_mode="zeros"):
    return F.grid_sample(x, grid, mode="bilinear", padding_mode=padding_mode, align_corners=True)   # grid_sample.py:12


def equivariance_loss(pred_tps, pred_all, grid, labels, logits, weak_threshold):
    """(:404-423) with ``pred_tps = model(tps(images_cj2))[0]`` given."""
    B, _, H, W = pred_all.shape
    mask = torch.ones((B, H, W), device=pred_all.device)
    neg = torch.zeros((B, H, W), device=pred_all.device)
    mask = torch.where(labels == 0, neg, mask)                                                   # :408
    mask = torch.where(logits < weak_threshold, neg, mask)                                       # :409
    mask_tps = warp(mask.unsqueeze(1).float(), grid)                                             # :414
    pred_tps_org = warp(pred_all.detach(), grid)                                                 # :418
    loss = F.kl_div(F.log_softmax(pred_tps, dim=1), F.softmax(pred_tps_org, dim=1), reduction="none")   # :419-421
    loss = (loss * mask_tps).flatten(1).sum(1) / (mask_tps.flatten(1).sum(1) + 1e-7)              # :422
    return loss.mean(), mask_tps, pred_tps_org                                                   # :423

#!/usr/bin/env python
"""Golden vectors for the SURVEY.md section 8(f) rows 3 and 4 (the other per-step loss terms of the 2-D trainer), made by
executing the REFERENCE's own source: the function definitions are cut out of ``/root/reference/code/train_arco_2d.py``
as text and exec'ed (the trainer itself cannot be imported: it parses argv and imports tensorboardX / h5py datasets at
module level), and ``RandTPS`` is imported from ``/root/reference/code/tps`` unmodified.  Build container only:

    python tests/golden/make_golden_step.py

Outputs: ``revisit_*.npz`` (get_revisiting_loss + the pool enqueue, :126-136, :109-120, :400-402),
``unsup_*.npz`` (compute_unsupervised_loss, :482-489), ``eqv_*.npz`` (RandTPS grid + the equivariance loss, :404-423).
"""
import argparse
import os
import random
import sys
import textwrap

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, "/root/reference/code")
TRAINER = "/root/reference/code/train_arco_2d.py"

from cases import EQV_CASES, REVISIT_CASES, UNSUP_CASES, eqv_inputs, revisit_inputs, unsup_inputs   # noqa: E402


def _def_source(name):
    """Source text of a top-level ``def name`` of the reference trainer (decorators included)."""
    lines = open(TRAINER).read().split("\n")
    d0 = next(i for i, ln in enumerate(lines) if ln.startswith(f"def {name}("))
    while d0 > 0 and lines[d0 - 1].startswith("@"):
        d0 -= 1
    d1 = next(i for i in range(d0 + 2, len(lines)) if lines[i] and not lines[i][0].isspace() and not lines[i].startswith("@")
              and not lines[i].startswith(")"))
    return "\n".join(lines[d0:d1])


def _ns(**kw):
    ns = dict(torch=torch, np=np, F=F, nn=nn)
    ns.update(kw)
    return ns


def run_revisit(case):
    x = revisit_inputs(case)
    ns = _ns(args=argparse.Namespace(K=x["pool"].shape[0]))
    exec(compile(_def_source("get_revisiting_loss") + "\n" + _def_source("_dequeue_and_enqueue"), "<reference trainer>", "exec"), ns)
    pool = x["pool"].clone()
    ptr = torch.zeros(1, dtype=torch.long)
    out = {}
    for step in range(case["steps"]):
        rs, rt = x["rep_u"][step], x["rep_u_teacher"][step]
        loss = ns["get_revisiting_loss"](random_pool=pool, rep_u=rs, rep_u_teacher=rt, topk=case["topk"])
        # train_arco_2d.py:400-402, verbatim
        k = rt.view(rt.shape[0], -1)
        k = torch.nn.functional.normalize(k, dim=-1)
        ns["_dequeue_and_enqueue"](keys=k, queue=pool, queue_ptr=ptr)
        out[f"s{step}_loss"] = loss.numpy()
        out[f"s{step}_ptr"] = ptr.clone().numpy()
    out["pool_after"] = pool.numpy()
    return out


def run_unsup(case):
    x = unsup_inputs(case)
    ns = _ns()
    exec(compile(_def_source("compute_unsupervised_loss"), "<reference trainer>", "exec"), ns)
    pred = x["predict"].clone().requires_grad_(True)
    loss = ns["compute_unsupervised_loss"](pred, x["target"], x["logits"], case["strong_threshold"])
    loss.backward()
    return dict(loss=loss.detach().numpy(), grad=pred.grad.numpy())


def run_eqv(case):
    from tps.rand_tps import RandTPS
    x = eqv_inputs(case)
    B, C, H, W = x["pred_all"].shape
    torch.manual_seed(case["seed"])
    random.seed(case["seed"])
    np.random.seed(case["seed"])
    tps = RandTPS(W, H, batch_size=B, sigma=case["sigma"], border_padding=False, random_mirror=True, random_scale=(0.8, 1.2),
                  mode="affine")
    tps.reset_control_points()
    grid = tps.grid.data.clone()
    # train_arco_2d.py:404-423 with the tensors of the case (model(images_tps)[0] is an input: pred_tps)
    labels, logits = x["labels"], x["logits"]
    mask = torch.ones((B, H, W))
    neg = torch.zeros((B, H, W))
    mask = torch.where(labels == 0, neg, mask)
    mask = torch.where(logits < case["weak_threshold"], neg, mask)
    mask = mask.unsqueeze(1)
    images_tps = tps(x["images"])
    mask_tps = tps(mask.float(), padding_mode="zeros")
    pred_tps = x["pred_tps"].clone().requires_grad_(True)
    pred_d = x["pred_all"].detach()
    pred_tps_org = tps(pred_d, padding_mode="zeros")
    kl = nn.KLDivLoss(reduction="none")
    loss_eqv = kl(F.log_softmax(pred_tps, dim=1), F.softmax(pred_tps_org, dim=1))
    loss_eqv = (loss_eqv * mask_tps).flatten(1).sum(1) / (mask_tps.flatten(1).sum(1) + 1e-7)
    loss_eqv = loss_eqv.mean()
    loss_eqv.backward()
    return dict(grid=grid.numpy(), images_tps=images_tps.detach().numpy(), mask_tps=mask_tps.numpy(),
                pred_tps_org=pred_tps_org.numpy(), loss=loss_eqv.detach().numpy(), grad=pred_tps.grad.numpy())


def main():
    for case in REVISIT_CASES:
        res = run_revisit(case)
        np.savez_compressed(os.path.join(HERE, case["name"] + ".npz"), **res)
        print(case["name"], [float(res[f"s{t}_loss"]) for t in range(case["steps"])])
    for case in UNSUP_CASES:
        res = run_unsup(case)
        np.savez_compressed(os.path.join(HERE, case["name"] + ".npz"), **res)
        print(case["name"], float(res["loss"]))
    for case in EQV_CASES:
        res = run_eqv(case)
        np.savez_compressed(os.path.join(HERE, case["name"] + ".npz"), **res)
        print(case["name"], float(res["loss"]))


if __name__ == "__main__":
    main()

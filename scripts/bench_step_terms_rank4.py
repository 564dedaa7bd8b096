"""Timing of the rank-4 step terms (imported by scripts/bench_step_terms.py)."""
import json
import random

import numpy as np
import torch

import arco_b200
import oracle
from bench_step_terms import PEAK, dev, timed


def bench_unsup():
    g = torch.Generator(device=dev).manual_seed(2)
    B, C, H, W = 12, 4, 256, 256
    pred = torch.randn(B, C, H, W, device=dev, generator=g)
    target = torch.randint(0, C, (B, H, W), device=dev, generator=g)
    target[torch.rand(B, H, W, device=dev, generator=g) < 0.1] = -1
    conf = torch.rand(B, H, W, device=dev, generator=g)

    def run(fn):
        p = pred.clone().requires_grad_(True)

        def step():
            p.grad = None
            fn(p, target, conf, 0.97).backward()
        return timed(step, n=20)

    ms, ms_ref = run(arco_b200.compute_unsupervised_loss), run(oracle.unsupervised_loss)
    byt = B * H * W * (2 * C * 4 + 8 + 4) + B * H * W * (C * 4 + 8) + B * C * H * W * 4
    print(json.dumps(dict(term="compute_unsupervised_loss fwd+bwd", shape=[B, C, H, W], ms=ms, alg_bytes=byt,
                          frac_hbm=byt / ms / 1e6 / PEAK, ms_reference_ops_on_gpu=ms_ref, speedup=ms_ref / ms)), flush=True)


def bench_eqv():
    g = torch.Generator(device=dev).manual_seed(3)
    B, C, H, W = 24, 4, 256, 256
    torch.manual_seed(3)
    random.seed(3)
    np.random.seed(3)
    tps = arco_b200.RandTPS(W, H, batch_size=B, sigma=0.01, random_scale=(0.8, 1.2), mode="affine")
    pred_tps = torch.randn(B, C, H, W, device=dev, generator=g)
    pred_all = torch.randn(B, C, H, W, device=dev, generator=g)
    images = torch.rand(B, 1, H, W, device=dev, generator=g)
    labels = torch.randint(0, C, (B, H, W), device=dev, generator=g)
    conf = torch.rand(B, H, W, device=dev, generator=g)

    def ours():
        p = pred_tps.detach().requires_grad_(True)
        tps(images)
        arco_b200.tps_equivariance_loss(p, pred_all, tps, labels, conf, 0.7).backward()

    def ref():
        p = pred_tps.detach().requires_grad_(True)
        oracle.warp(images, tps.grid.data)
        oracle.equivariance_loss(p, pred_all, tps.grid.data, labels, conf, 0.7)[0].backward()

    ms_grid = timed(lambda: tps.reset_control_points(), n=10)
    ms, ms_ref = timed(ours, n=20), timed(ref, n=20)
    byt = B * H * W * (8 + 2 * 4 + 8 + 4 + 2 * C * 4 + C * 4 + 2 * C * 4)
    print(json.dumps(dict(term="equivariance loss fwd+bwd (+ image warp)", shape=[B, C, H, W], ms=ms, alg_bytes=byt,
                          frac_hbm=byt / ms / 1e6 / PEAK, ms_reference_ops_on_gpu=ms_ref, speedup=ms_ref / ms,
                          ms_reset_control_points_incl_host_rng=ms_grid)), flush=True)

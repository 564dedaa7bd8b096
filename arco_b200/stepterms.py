"""The other per-pixel loss terms of the 2-D trainer's step, fused.

Drop-ins for ``/root/reference/code/train_arco_2d.py``:

* :func:`compute_unsupervised_loss` -- same name / arguments as the trainer's helper (:482-489), differentiable w.r.t.
  ``predict``: one forward pass + one backward pass instead of cross_entropy + masked_select + broadcasts.
* :class:`RandTPS` -- same constructor and methods as ``tps/rand_tps.py:82-153`` (``reset_control_points`` consumes the
  host RNG streams exactly like the reference -- torch CPU ``uniform_``, NumPy ``uniform`` x4, ``random.randint`` -- so a
  seeded run draws the same warps); the sampling grid is computed and kept on the GPU (``arco_tps_grid``), ``tps(x)`` is
  ``arco_grid_sample`` (bilinear, ``align_corners=True``; forward only, as the trainer uses it).
* :func:`tps_equivariance_loss` -- the trainer's block :404-423 minus the model forward: mask construction, its warp,
  the warp of the detached predictions, both softmaxes and the masked KL in ONE kernel; the backward pass rescales the
  gradient the forward kernel already wrote.

No CPU path: every function needs CUDA tensors.
"""
from __future__ import annotations

import itertools
import random
from typing import Optional

import numpy as np
import torch

from . import _cabi


def _need_cuda(*ts):
    for t in ts:
        if not (torch.is_tensor(t) and t.is_cuda):
            raise RuntimeError("arco_b200 step terms need CUDA tensors: there is no CPU fallback")


def _scratch(batch: int, space: int, dev) -> torch.Tensor:
    return torch.empty(int(_cabi.lib.arco_step_scratch_bytes(batch, space)), dtype=torch.uint8, device=dev)


class _UnsupLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, predict, target, logits, thr):
        B, C = predict.shape[0], predict.shape[1]
        S = predict[0, 0].numel()
        dev = predict.device
        p = predict.detach().contiguous()
        t = target.contiguous()
        lg = logits.detach().to(torch.float32).contiguous()
        stats = torch.empty(B + 2, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _cabi.check(_cabi.lib.arco_unsup_loss(p.data_ptr(), t.data_ptr(), lg.data_ptr(), float(thr), B, C, S, stats.data_ptr(),
                                                  _scratch(B, S, dev).data_ptr(), torch.cuda.current_stream().cuda_stream),
                        "arco_unsup_loss")
        ctx.save_for_backward(p, t, stats)
        return stats[B + 1].clone()

    @staticmethod
    def backward(ctx, grad_out):
        p, t, stats = ctx.saved_tensors
        B, C = p.shape[0], p.shape[1]
        S = p[0, 0].numel()
        grad = torch.empty_like(p)
        go = grad_out.detach().to(torch.float32).reshape(1).contiguous()
        with torch.cuda.device(p.device):
            _cabi.check(_cabi.lib.arco_unsup_loss_backward(p.data_ptr(), t.data_ptr(), stats.data_ptr(), go.data_ptr(), B, C, S,
                                                           grad.data_ptr(), torch.cuda.current_stream().cuda_stream),
                        "arco_unsup_loss_backward")
        return grad, None, None, None


def compute_unsupervised_loss(predict, target, logits, strong_threshold):
    """``compute_unsupervised_loss(predict [B,C,*S], target [B,*S] int64 (ignore -1), logits [B,*S], strong_threshold)``
    (train_arco_2d.py:482-489)."""
    _need_cuda(predict, target, logits)
    if predict.dtype != torch.float32 or target.dtype != torch.int64:
        raise ValueError("predict must be float32 and target int64")
    if tuple(target.shape) != (predict.shape[0],) + tuple(predict.shape[2:]) or logits.shape != target.shape:
        raise ValueError("target / logits must be [B, *S] matching predict [B, C, *S]")
    return _UnsupLoss.apply(predict, target, logits, float(strong_threshold))


# ----------------------------------------------------------------------------------------------------------------------
def _partial_repr(points: torch.Tensor, control: torch.Tensor) -> torch.Tensor:
    """phi(x1, x2) = r^2 log r with 0 log 0 := 0 (tps_stn_pytorch/tps_grid_gen.py:9-21); host side, 25 x 25."""
    diff = points.view(-1, 1, 2) - control.view(1, -1, 2)
    d2 = diff[:, :, 0] * diff[:, :, 0] + diff[:, :, 1] * diff[:, :, 1]
    rep = 0.5 * d2 * torch.log(d2)
    rep.masked_fill_(rep != rep, 0)
    return rep


def _similarity_matrices(batch_size, img_sz, translate=0.1, random_scale=(0.7, 1.1), rotate=60):
    """generate_transformer_matrices (tps/rand_tps.py:56-80): same NumPy draws in the same order."""
    angle = np.random.uniform(size=[batch_size, ], low=-rotate, high=rotate) / 180.0 * np.pi
    scale = np.random.uniform(size=[batch_size, ], low=random_scale[0], high=random_scale[1])
    shift_x = np.random.uniform(size=(batch_size,), low=-translate, high=translate).reshape((-1, 1))
    shift_y = np.random.uniform(size=(batch_size,), low=-translate, high=translate).reshape((-1, 1))
    img_sz_f = np.float32(img_sz)
    cos_v = (scale * np.cos(angle)).reshape((-1, 1))
    sin_v = (scale * np.sin(angle)).reshape((-1, 1))
    return np.concatenate([cos_v, -sin_v, shift_x * np.float32(img_sz_f / 2.0), sin_v, cos_v,
                           shift_y * np.float32(img_sz_f / 2.0)], axis=1)


def _perspective_matrices(batch_size, random_scale=(0.7, 1.1), rotate=(10, 10, 60)):
    """generate_perspective_matrices (tps/rand_tps.py:18-54)."""
    ax = np.random.uniform(size=[batch_size, ], low=-rotate[0], high=rotate[0]) / 180.0 * np.pi
    ay = np.random.uniform(size=[batch_size, ], low=-rotate[1], high=rotate[1]) / 180.0 * np.pi
    az = np.random.uniform(size=[batch_size, ], low=-rotate[2], high=rotate[2]) / 180.0 * np.pi
    ones, zeros = np.ones(batch_size).reshape((-1, 1)), np.zeros(batch_size).reshape((-1, 1))
    scale = np.random.uniform(size=[batch_size, ], low=random_scale[0], high=random_scale[1])
    c, s = np.cos(ax).reshape((-1, 1)), np.sin(ax).reshape((-1, 1))
    rx = torch.from_numpy(np.concatenate([ones, zeros, zeros, zeros, c, -s, zeros, s, c], axis=1).reshape(batch_size, 3, 3)).transpose(1, 2)
    c, s = np.cos(ay).reshape((-1, 1)), np.sin(ay).reshape((-1, 1))
    ry = torch.from_numpy(np.concatenate([c, zeros, s, zeros, ones, zeros, -s, zeros, c], axis=1).reshape(batch_size, 3, 3)).transpose(1, 2)
    c, s = (scale * np.cos(az)).reshape((-1, 1)), (scale * np.sin(az)).reshape((-1, 1))
    rz = torch.from_numpy(np.concatenate([c, -s, zeros, s, c, zeros, zeros, zeros, ones], axis=1).reshape(batch_size, 3, 3)).transpose(1, 2)
    return torch.matmul(rz, torch.matmul(ry, rx))


def draw_source_control_points(target_control_points, batch_size, sigma, random_scale, mode="affine", rand_mirror=True):
    """The host-side part of ``RandTPS.reset_control_points`` (tps/rand_tps.py:114-141): consumes the same RNG streams in
    the same order as the reference (torch CPU ``uniform_``, NumPy ``uniform`` x4 or x4 for the projective mode, Python
    ``random.randint`` for the mirror).  ``random_scale`` is the already inverted pair RandTPS stores (:89).  CPU only."""
    src = target_control_points.unsqueeze(0).repeat(batch_size, 1, 1)
    src += torch.Tensor(src.size()).uniform_(-sigma, sigma)
    if mode == "affine":
        theta = _similarity_matrices(batch_size, 2.0, random_scale=random_scale)
        m = torch.from_numpy(theta.reshape((-1, 2, 3)).copy()).type(torch.FloatTensor).transpose(1, 2)
        src = torch.matmul(torch.cat((src, torch.ones(*src.shape[0:2], 1)), dim=2), m)
    elif mode == "projective":
        R = _perspective_matrices(batch_size, random_scale=random_scale).type(torch.FloatTensor).detach()
        src = torch.matmul(torch.cat((src, torch.ones(*src.shape[0:2], 1)), dim=2), R)
        src[:, :, 0] = src[:, :, 0] / src[:, :, 2]
        src[:, :, 1] = src[:, :, 1] / src[:, :, 2]
        src = src[:, :, :2]
    if rand_mirror and random.randint(0, 1):
        src[:, :, 0] = -src[:, :, 0]
    return src


class RandTPS(torch.nn.Module):
    """Twin of ``tps.rand_tps.RandTPS`` (same arguments, attributes ``grid`` / ``target_control_points`` and methods)."""

    def __init__(self, width, height, batch_size=16, sigma=0.01, border_padding=False, random_mirror=True,
                 random_scale=(0.7, 1.1), mode="affine", device=None):
        super().__init__()
        self.width, self.height, self.batch_size, self.sigma = int(width), int(height), int(batch_size), sigma
        self.random_scale = (1.0 / random_scale[1], 1.0 / random_scale[0])     # applied target -> source (:89)
        self.padding_mode = "border" if border_padding else "zeros"
        self.rand_mirror, self.mode = random_mirror, mode
        self.target_control_points = torch.Tensor(list(itertools.product(torch.arange(-1.0, 1.00001, 2.0 / 4),
                                                                         torch.arange(-1.0, 1.00001, 2.0 / 4))))
        n = self.target_control_points.shape[0]
        fk = torch.zeros(n + 3, n + 3)                                           # TPSGridGen.__init__ (:31-40)
        fk[:n, :n].copy_(_partial_repr(self.target_control_points, self.target_control_points))
        fk[:n, -3].fill_(1)
        fk[-3, :n].fill_(1)
        fk[:n, -2:].copy_(self.target_control_points)
        fk[-2:, :n].copy_(self.target_control_points.transpose(0, 1))
        self._inverse_kernel = torch.inverse(fk)
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.grid = torch.nn.Parameter(torch.zeros(self.batch_size, self.height, self.width, 2, device=dev), requires_grad=False)
        self._ctrl_dev = self.target_control_points.to(dev).contiguous()
        self.reset_control_points()

    def source_control_points(self) -> torch.Tensor:
        return draw_source_control_points(self.target_control_points, self.batch_size, self.sigma, self.random_scale, self.mode,
                                          self.rand_mirror)

    def reset_control_points(self, source_control_points: Optional[torch.Tensor] = None):
        src = self.source_control_points() if source_control_points is None else source_control_points.float().cpu()
        # TPSGridGen.forward (:62-73): mapping = inverse_kernel @ [source; 0]  (28 x 28, host), grid on the device
        Y = torch.cat([src, torch.zeros(self.batch_size, 3, 2)], 1)
        mapping = torch.matmul(self._inverse_kernel, Y).contiguous()
        dev = self.grid.device
        with torch.cuda.device(dev):
            mp = mapping.to(dev)
            n = self.target_control_points.shape[0]
            _cabi.check(_cabi.lib.arco_tps_grid(mp.data_ptr(), self._ctrl_dev.data_ptr(), n, self.batch_size, self.height, self.width,
                                                self.grid.data.data_ptr(), torch.cuda.current_stream().cuda_stream), "arco_tps_grid")
        self.grid.requires_grad = False

    def forward(self, x, padding_mode=None, mode="bilinear"):
        _need_cuda(x)
        if mode != "bilinear":
            raise ValueError("only bilinear sampling is implemented (the trainers use nothing else)")
        if x.requires_grad and torch.is_grad_enabled():
            raise RuntimeError("arco_b200.RandTPS is forward-only: the trainers warp inputs, masks and DETACHED predictions")
        pm = self.padding_mode if padding_mode is None else padding_mode
        B, Cc, H, W = x.shape
        if (B, H, W) != (self.batch_size, self.height, self.width):
            raise ValueError(f"expected [{self.batch_size}, C, {self.height}, {self.width}], got {tuple(x.shape)}")
        xin = x.detach().to(torch.float32).contiguous()
        out = torch.empty_like(xin)
        with torch.cuda.device(x.device):
            _cabi.check(_cabi.lib.arco_grid_sample(xin.data_ptr(), self.grid.data.data_ptr(), B, Cc, H, W, 1 if pm == "border" else 0,
                                                   out.data_ptr(), torch.cuda.current_stream().cuda_stream), "arco_grid_sample")
        return out


class _EqvLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred_tps, pred_d, grid, labels, logits, weak_thr):
        B, Cc, H, W = pred_tps.shape
        dev = pred_tps.device
        pt = pred_tps.detach().contiguous()
        need_grad = pred_tps.requires_grad
        g = torch.empty_like(pt) if need_grad else None
        stats = torch.empty(B + 1, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _cabi.check(_cabi.lib.arco_eqv_loss(pt.data_ptr(), pred_d.data_ptr(), grid.data_ptr(), labels.data_ptr(), logits.data_ptr(),
                                                float(weak_thr), B, Cc, H, W, stats.data_ptr(), g.data_ptr() if g is not None else None,
                                                _scratch(B, H * W, dev).data_ptr(), torch.cuda.current_stream().cuda_stream),
                        "arco_eqv_loss")
        ctx.g, ctx.stats, ctx.B = g, stats, B
        return stats[B].clone()

    @staticmethod
    def backward(ctx, grad_out):
        g, stats, B = ctx.g, ctx.stats, ctx.B
        go = grad_out.detach().to(torch.float32).reshape(1).contiguous()
        with torch.cuda.device(g.device):
            _cabi.check(_cabi.lib.arco_scale_rows(g.data_ptr(), stats.data_ptr(), go.data_ptr(), B, g[0].numel(), g.data_ptr(),
                                                  torch.cuda.current_stream().cuda_stream), "arco_scale_rows")
        ctx.g = None
        return g, None, None, None, None, None


def tps_equivariance_loss(pred_tps, pred_all, tps: RandTPS, labels, logits, weak_threshold):
    """The trainer's equivariance term (train_arco_2d.py:404-423) given ``pred_tps = model(tps(images_cj2))[0]``:

        mask = (labels != 0) & ~(logits < weak_threshold);  mask_tps = tps(mask, 'zeros')
        loss = mean_b  sum(KL(log_softmax(pred_tps) || softmax(tps(pred_all.detach(), 'zeros'))) * mask_tps) / (sum(mask_tps) + 1e-7)

    ``labels`` int64 ``[B,H,W]``, ``logits`` float ``[B,H,W]``; differentiable w.r.t. ``pred_tps`` only (as in the reference)."""
    _need_cuda(pred_tps, pred_all, labels, logits)
    if pred_tps.dtype != torch.float32 or pred_all.shape != pred_tps.shape:
        raise ValueError("pred_tps / pred_all must be float32 [B, C, H, W] of equal shape")
    if pred_tps.shape[1] > 32:
        raise ValueError("at most 32 classes")
    return _EqvLoss.apply(pred_tps, pred_all.detach().to(torch.float32).contiguous(), tps.grid.data.contiguous(),
                          labels.to(torch.int64).contiguous(), logits.detach().to(torch.float32).contiguous(), float(weak_threshold))

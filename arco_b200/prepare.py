"""Device-side mask / threshold preparation feeding the loss.

Replaces the ``with torch.no_grad():`` block of the trainers -- ``train_arco_2d.py:345-393``, ``train_arco_3d.py:315-353``:
teacher softmax, student entropy, two ``np.percentile`` calls (each a GPU -> CPU -> GPU round trip with a host sync in the
reference), the entropy masks and the CPU one-hot.  Everything stays on the device and on the current stream; the
one-hot is not built at all: :func:`arco_b200.compute_contra_memobank_loss` takes the integer label maps.

Trainer edit::

    p = arco_b200.prepare_contrast_inputs(pred_u, pred_l_teacher, pred_u_teacher, train_l_label, train_u_aug_label, alpha_t)
    reco_loss = compute_contra_memobank_loss(rep_all, p["label_l"], p["label_u"], p["prob_l_teacher"], p["prob_u_teacher"],
                                             p["low_mask_all"], p["high_mask_all"], memobank, queue_ptrlis, queue_size, ...)[-1]

Predictions and labels must have the same spatial size (the 2-D trainer's ``F.interpolate(..., mode='nearest')`` is the
identity in that case, ``train_arco_2d.py:348-350``).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _cabi


def _q32(percent: float) -> float:
    """numpy 2.x forms the quantile of float32 data as float32(percent) / float32(100) (np.percentile, method 'linear')."""
    return float(np.float32(percent) / np.float32(100))


def softmax_entropy(logits: torch.Tensor, want_prob: bool = True, want_entropy: bool = False):
    """softmax over dim 1 of ``[B, C, *S]`` float32 logits and/or ``-sum(p * log(p + 1e-10), dim=1)``."""
    if not (logits.is_cuda and logits.dtype == torch.float32 and logits.dim() >= 3):
        raise ValueError("logits must be a CUDA float32 tensor [B, C, *S]")
    x = logits.detach().contiguous()
    B, Cn = x.shape[0], x.shape[1]
    S = int(np.prod(x.shape[2:]))
    prob = torch.empty_like(x) if want_prob else None
    ent = torch.empty((B,) + tuple(x.shape[2:]), dtype=torch.float32, device=x.device) if want_entropy else None
    with torch.cuda.device(x.device):
        _cabi.check(_cabi.lib.arco_softmax_rows(x.data_ptr(), B, Cn, S, prob.data_ptr() if want_prob else None,
                                                ent.data_ptr() if want_entropy else None,
                                                torch.cuda.current_stream().cuda_stream), "arco_softmax_rows")
    return prob, ent


def entropy_masks(entropy: torch.Tensor, train_l_label: torch.Tensor, train_u_aug_label: torch.Tensor, alpha_t: float):
    """``low_mask_all, high_mask_all`` f32 ``[B_l + B_u, 1, *S]`` and the two thresholds (device ``float32[2]``) from the
    student entropy of the unlabelled images, exactly as ``train_arco_2d.py:360-392`` (``np.percentile`` of numpy 2.x)."""
    if not (entropy.is_cuda and entropy.dtype == torch.float32):
        raise ValueError("entropy must be a CUDA float32 tensor [B_u, *S]")
    if train_l_label.dtype != torch.int64 or train_u_aug_label.dtype != torch.int64:
        raise ValueError("label maps must be int64 (ignore label negative)")
    if entropy.shape != train_u_aug_label.shape or train_l_label.shape[1:] != train_u_aug_label.shape[1:]:
        raise ValueError("entropy / label shapes disagree")
    dev = entropy.device
    e = entropy.contiguous()
    ll = train_l_label.to(dev).contiguous()
    lu = train_u_aug_label.to(dev).contiguous()
    spatial = tuple(lu.shape[1:])
    n_l, n_u = ll.numel(), lu.numel()
    low = torch.empty((ll.shape[0] + lu.shape[0], 1) + spatial, dtype=torch.float32, device=dev)
    high = torch.empty_like(low)
    thr = torch.empty(2, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        scratch = torch.empty(int(_cabi.lib.arco_entropy_masks_scratch()), dtype=torch.uint8, device=dev)
        _cabi.check(_cabi.lib.arco_entropy_masks(e.data_ptr(), ll.data_ptr(), lu.data_ptr(), n_l, n_u, _q32(alpha_t),
                                                 _q32(100 - alpha_t), low.data_ptr(), high.data_ptr(), thr.data_ptr(),
                                                 scratch.data_ptr(), torch.cuda.current_stream().cuda_stream),
                    "arco_entropy_masks")
    return low, high, thr


def prepare_contrast_inputs(pred_u, pred_l_teacher, pred_u_teacher, train_l_label, train_u_aug_label, alpha_t: float) -> dict:
    """Everything ``compute_contra_memobank_loss`` needs besides the representations, from the trainers' raw tensors:
    student logits of the unlabelled images, teacher logits of both halves, the integer label maps and ``alpha_t``."""
    for t in (pred_u, pred_l_teacher, pred_u_teacher):
        if not (t.is_cuda and t.dtype == torch.float32 and t.dim() >= 3):
            raise ValueError("predictions must be CUDA float32 tensors [B, C, *S]")
    dev = pred_u.device
    pu, plt_, put = pred_u.detach().contiguous(), pred_l_teacher.detach().contiguous(), pred_u_teacher.detach().contiguous()
    ll, lu = train_l_label.to(dev).contiguous(), train_u_aug_label.to(dev).contiguous()
    if ll.dtype != torch.int64 or lu.dtype != torch.int64:
        raise ValueError("label maps must be int64 (ignore label negative)")
    n_l, n_u, Cn = plt_.shape[0], pu.shape[0], pu.shape[1]
    spatial = tuple(pu.shape[2:])
    if put.shape != pu.shape or plt_.shape[1:] != pu.shape[1:] or tuple(ll.shape) != (n_l,) + spatial or tuple(lu.shape) != (n_u,) + spatial:
        raise ValueError("prediction / label shapes disagree (labels must have the predictions' spatial size)")
    S = int(np.prod(spatial))
    prob_l = torch.empty_like(plt_)
    prob_u = torch.empty_like(put)
    entropy = torch.empty((n_u,) + spatial, dtype=torch.float32, device=dev)
    low = torch.empty((n_l + n_u, 1) + spatial, dtype=torch.float32, device=dev)
    high = torch.empty_like(low)
    thr = torch.empty(2, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        scratch = torch.empty(int(_cabi.lib.arco_entropy_masks_scratch()), dtype=torch.uint8, device=dev)
        _cabi.check(_cabi.lib.arco_prepare_contrast(
            pu.data_ptr(), plt_.data_ptr(), put.data_ptr(), ll.data_ptr(), lu.data_ptr(), n_l, n_u, Cn, S, _q32(alpha_t),
            _q32(100 - alpha_t), prob_l.data_ptr(), prob_u.data_ptr(), entropy.data_ptr(), low.data_ptr(), high.data_ptr(),
            thr.data_ptr(), scratch.data_ptr(), torch.cuda.current_stream().cuda_stream), "arco_prepare_contrast")
    return dict(label_l=ll, label_u=lu, prob_l_teacher=prob_l, prob_u_teacher=prob_u, low_mask_all=low, high_mask_all=high,
                entropy=entropy, thresholds=thr)

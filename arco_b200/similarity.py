"""Config-5 dense similarity: cosine logits of every query against its N sampled bank
rows as ONE tcgen05 GEMM per class plus a scalar gather, instead of Q*N row gathers (loss_helper_3d.py:466-486).
Forward AND backward (a second tcgen05 GEMM ``[Q, M] x [M, D]`` of the scattered logit gradients against the transposed
ring): it exists to measure where the dense form overtakes the gather form (``scripts/sweep_config5.py``)."""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import torch

from . import _cabi
from .bank import BankSlot, DeviceMemoryBank


class _DenseSim(torch.autograd.Function):
    @staticmethod
    def forward(ctx, anchors, bank, cls, idx_neg):
        n_slots, Q, D = anchors.shape
        N = idx_neg.shape[2]
        a = anchors.detach().contiguous()
        with torch.cuda.device(a.device):
            need = _cabi.lib.arco_similarity_dense_scratch(D, Q, n_slots, C.byref(bank.c_struct), cls)
            if need < 0:
                raise ValueError("bad arguments for arco_similarity_dense_scratch")
            scratch = torch.empty(need, dtype=torch.uint8, device=a.device)
            out = torch.empty((n_slots, Q, N), dtype=torch.float32, device=a.device)
            _cabi.check(_cabi.lib.arco_similarity_dense(D, Q, N, n_slots, cls, a.data_ptr(), C.byref(bank.c_struct),
                                                        idx_neg.data_ptr(), out.data_ptr(), scratch.data_ptr(),
                                                        torch.cuda.current_stream().cuda_stream), "arco_similarity_dense")
        ctx.save_for_backward(a, idx_neg)
        ctx.bank, ctx.cls = bank, cls
        return out

    @staticmethod
    def backward(ctx, grad_logits):
        a, idx_neg = ctx.saved_tensors
        bank, cls = ctx.bank, ctx.cls
        n_slots, Q, D = a.shape
        N = idx_neg.shape[2]
        g = grad_logits.detach().to(torch.float32).contiguous()
        with torch.cuda.device(a.device):
            need = _cabi.lib.arco_similarity_dense_backward_scratch(D, Q, n_slots, C.byref(bank.c_struct), cls)
            scratch = torch.empty(max(int(need), 1), dtype=torch.uint8, device=a.device)
            g_hat = torch.empty((n_slots, Q, D), dtype=torch.float32, device=a.device)
            _cabi.check(_cabi.lib.arco_similarity_dense_backward(D, Q, N, n_slots, cls, g.data_ptr(), C.byref(bank.c_struct),
                                                                 idx_neg.data_ptr(), g_hat.data_ptr(), scratch.data_ptr(),
                                                                 torch.cuda.current_stream().cuda_stream),
                        "arco_similarity_dense_backward")
        # through the normalisation a_hat = a / max(|a|, 1e-8) (torch.cosine_similarity's per-norm eps)
        na = a.norm(dim=2, keepdim=True)
        big = na > 1e-8
        a_hat = a / na.clamp_min(1e-8)
        grad_a = torch.where(big, (g_hat - (g_hat * a_hat).sum(dim=2, keepdim=True) * a_hat) / na.clamp_min(1e-8), g_hat / 1e-8)
        return grad_a, None, None, None


def dense_similarity(anchors: torch.Tensor, memobank, slot_classes: Sequence[int], idx_neg: torch.Tensor) -> torch.Tensor:
    """``anchors`` f32 ``[n_slots, Q, D]`` (raw rows), ``memobank`` an adopted bank (``memobank[c]`` is a
    :class:`BankSlot`) or a :class:`DeviceMemoryBank` with bf16 ring storage, ``slot_classes[j]`` the bank class slot j
    is contrasted against, ``idx_neg`` int32 ``[n_slots, Q, N]`` logical ring rows.  Returns f32 cosines
    ``[n_slots, Q, N]`` (not divided by the temperature); differentiable with respect to ``anchors``."""
    bank = memobank if isinstance(memobank, DeviceMemoryBank) else memobank[0].bank
    if not isinstance(bank, DeviceMemoryBank):
        raise TypeError("memobank must be adopted by arco_b200 (call compute_contra_memobank_loss once, or build a DeviceMemoryBank)")
    if not (anchors.is_cuda and anchors.dtype == torch.float32 and anchors.dim() == 3):
        raise ValueError("anchors must be a CUDA float32 tensor [n_slots, Q, D]")
    n_slots, Q, D = anchors.shape
    if idx_neg.dtype != torch.int32 or idx_neg.dim() != 3 or idx_neg.shape[:2] != (n_slots, Q) or idx_neg.device != anchors.device:
        raise ValueError("idx_neg must be int32 [n_slots, Q, N] on the anchors' device")
    if bank.row_dtype != torch.bfloat16:
        raise ValueError("dense_similarity needs a bf16 ring (a bf16 representation head with bf16-exact rows)")
    if D != bank.feat or len(slot_classes) != n_slots:
        raise ValueError("feature size / slot count mismatch")
    bank.settle()
    cls = (C.c_int32 * n_slots)(*[int(c) for c in slot_classes])
    return _DenseSim.apply(anchors, bank, cls, idx_neg.contiguous())
